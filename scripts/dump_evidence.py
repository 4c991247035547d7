#!/usr/bin/env python
"""Regenerates the static evidence under profiles/ from the current sources (no GPU needed):

* ``profiles/sass/<kernel>.sass``  — per-kernel SASS listings of the tensor-core kernels (encodings stripped)
* ``profiles/sass_mnemonics.txt``   — per-kernel counts of the mnemonics that prove the Blackwell path
  (UTC*MMA = tcgen05.mma, UTMALDG = TMA, LDTM/STTM = TMEM, SYNCS = mbarrier, MEMBAR/STRONG.SYS = federation)
* ``profiles/ptx_evidence.txt``     — PTX-level counts (tcgen05.*, cp.async.bulk.tensor, multimem.*, griddepcontrol)

    python scripts/dump_evidence.py
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pytensor_federated_b200 import build as native_build  # noqa: E402

PROFILES = ROOT / "profiles"
INTERESTING = re.compile(r"^(UTC\w*MMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|SYNCS|UTCBAR|MEMBAR|CCTL|ELECT|LDGSTS|HMMA|"
                         r"ATOM\w*|RED)|STRONG\.SYS|ACQUIRE|\.MMIO")
PTX_PATTERNS = ("tcgen05.", "cp.async.bulk.tensor", "multimem.", "griddepcontrol", "ld.acquire.sys", "st.release.sys",
                "mbarrier.try_wait", "mbarrier.arrive.expect_tx", "elect.sync", "fence.proxy.async", "setmaxnreg")
LISTED = {   # kernel-name fragment -> listing file
    "fed_glm_tc_kernelILi1E": "glm_tc_k1", "fed_glm_tc_kernelILi16E": "glm_tc_k16",
    "fed_glm_fp8_kernelILi1ELb0E": "glm_fp8_k1", "fed_ode_generic_kernel": "ode_generic_lv",
}


def main() -> None:
    lib = native_build.build()
    res = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True)
    kernels, name = collections.OrderedDict(), None
    for line in res.stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);?\s*/\*", line)
        if m:
            kernels[name].append((m.group(1), m.group(2).strip()))
    (PROFILES / "sass").mkdir(parents=True, exist_ok=True)
    with open(PROFILES / "sass_mnemonics.txt", "w") as out:
        out.write("# per-kernel SASS mnemonic counts (cuobjdump -sass of libb200fed.so; scripts/dump_evidence.py)\n")
        for kname, ins in kernels.items():
            counts = collections.Counter()
            for _, text in ins:
                body = re.sub(r"^@!?U?P\d+\s+", "", text)
                op = body.split()[0] if body else ""
                if INTERESTING.search(op):
                    counts[op] += 1
            generic_smem = sum(1 for _, t in ins if re.search(r"\b(LD|ST)\.E\b", t) and "desc[" not in t)
            out.write(f"{kname}\n")
            for op, n in sorted(counts.items()):
                out.write(f"   {op:40s} {n}\n")
            if "tc_kernel" in kname or "fp8_kernel" in kname:
                lds = sum(1 for _, t in ins if re.match(r"(@!?U?P\d+\s+)?LDS", t))
                sts = sum(1 for _, t in ins if re.match(r"(@!?U?P\d+\s+)?STS", t))
                out.write(f"   {'(shared-space LDS / STS)':40s} {lds} / {sts}\n")
                out.write(f"   {'(generic LD.E / ST.E without descriptor)':40s} {generic_smem}\n")
            for frag, fname in LISTED.items():
                if frag in kname:
                    with open(PROFILES / "sass" / f"{fname}.sass", "w") as fh:
                        fh.write(f"// {kname}\n")
                        for addr, text in ins:
                            fh.write(f"/*{addr}*/ {text}\n")
    with open(PROFILES / "ptx_evidence.txt", "w") as out:
        out.write("# PTX-level evidence: nvcc -ptx of every source, instruction templates counted (scripts/dump_evidence.py)\n")
        for src in native_build.SOURCES:
            cmd = [native_build.nvcc_path(), "-arch=compute_100a", "-std=c++17", "--expt-relaxed-constexpr", "-I", str(native_build.CSRC),
                   "-ptx", str(native_build.CSRC / src), "-o", "/dev/stdout"]
            ptx = subprocess.run(cmd, capture_output=True, text=True)
            if ptx.returncode != 0:
                out.write(f"== {src}: nvcc -ptx failed\n")
                continue
            counts = collections.Counter()
            for line in ptx.stdout.splitlines():
                line = line.strip()
                if any(p in line for p in PTX_PATTERNS):
                    line = re.sub(r"%\w+", "%r", line)
                    line = re.sub(r"\[[^\]]*\]", "[..]", line)
                    line = re.sub(r"\{[^}]*\}", "{..}", line)
                    counts[line] += 1
            if counts:
                out.write(f"== {src}\n")
                for line, n in counts.most_common():
                    out.write(f"{n:7d} {line}\n")
    print("wrote", PROFILES / "sass_mnemonics.txt", PROFILES / "ptx_evidence.txt", "and profiles/sass/")


if __name__ == "__main__":
    main()
