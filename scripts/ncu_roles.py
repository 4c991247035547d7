#!/usr/bin/env python
"""Per-role view of an ncu capture of a warp-specialised kernel.

    python scripts/ncu_roles.py gpurun_out/prof_glm_fp8_v4.ncu-rep [--tiles 125000]
                                [--regions 0xd700:0xe200:tma,0xe200:0x10f00:epilogue,...]

Prints (1) the instructions with the most stall samples, (2) every barrier wait / TMA / MMA / commit
instruction with its sample and execution counts (spin counts show which barrier a role waits on), and
(3) for the given address regions: warp-instructions per tile and samples per role.  Regions are the SASS
address ranges of the role branches (find them from the marker list of (2)).  This is how
profiles/issue_loop_analysis.txt was produced.
"""
import argparse
import csv
import io
import subprocess


def load(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr)]
    return rows[0], hdr, ix, body


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tiles", type=int, default=0, help="tiles processed in the capture (for per-tile counts)")
    ap.add_argument("--regions", default="", help="lo:hi:name,... (hex SASS addresses, low 20 bits)")
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()
    title, hdr, ix, body = load(args.report)
    print(title[1] if len(title) > 1 else title)

    def num(r, key):
        try:
            return int(r[ix[key]] or 0)
        except ValueError:
            return 0

    def addr(r):
        return int(r[ix["Address"]], 16) & 0xFFFFF

    total = sum(num(r, "# Samples") for r in body)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "(" not in h]
    print(f"\n== top {args.top} instructions by stall samples (total {total})")
    for r in sorted(body, key=lambda r: -num(r, "# Samples"))[: args.top]:
        top = sorted(((k, num(r, k)) for k in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"{num(r, '# Samples'):7d} {100 * num(r, '# Samples') / max(total, 1):5.1f}%  {addr(r):#07x}  "
              f"{r[ix['Source']].strip()[:72]:72s} {top}")
    print("\n== barrier waits / TMA / MMA / commits: address, samples (incl. the following branch), executions")
    markers = ("SYNCS.PHASECHK", "UTCQMMA", "UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "STTM")
    for i, r in enumerate(body):
        src = r[ix["Source"]]
        if any(m in src for m in markers):
            n = num(r, "# Samples") + (num(body[i + 1], "# Samples") if i + 1 < len(body) else 0)
            print(f"{addr(r):#07x}  {n:7d}  {num(r, 'Instructions Executed'):10d}  {src.strip()[:80]}")
    if args.regions:
        print("\n== roles")
        for spec in args.regions.split(","):
            lo, hi, name = spec.split(":")
            sel = [r for r in body if int(lo, 16) <= addr(r) < int(hi, 16)]
            instr = sum(num(r, "Instructions Executed") for r in sel)
            samples = sum(num(r, "# Samples") for r in sel)
            stalls = sorted(((k, sum(num(r, k) for r in sel)) for k in stall_cols), key=lambda kv: -kv[1])[:3]
            per_tile = f"{instr / args.tiles:9.1f} instr/tile" if args.tiles else f"{instr} instr"
            print(f"{name:>14s}: {per_tile}, {samples:7d} samples, {stalls}")


if __name__ == "__main__":
    main()
