"""The operand-splitting schemes of the tensor-core kernels, emulated with torch dtypes on the CPU.

The kernels feed low-precision operands to the MMA and recombine in fp32 (csrc/glm_tc.cu: theta as three
bf16 terms, residuals as two; csrc/glm_fp8.cu: radix-16 e4m3 expansions, residuals relative to a per-32-row
power-of-two scale).  These tests pin the accuracy those schemes deliver, independent of any GPU.
"""
import numpy as np
import pytest
import torch


def bf16_terms(x: torch.Tensor, n: int):
    """x ~= sum of n bf16 terms (hi, mid, lo ...), as the kernels build them."""
    terms, rem = [], x.clone()
    for _ in range(n):
        t = rem.to(torch.bfloat16)
        terms.append(t)
        rem = rem - t.to(torch.float32)
    return terms


def e4m3_expansion(x: torch.Tensor, n: int):
    """x ~= sum_k t_k 16^-k with e4m3 terms (expand16 in glm_fp8.cu)."""
    terms, rem = [], x.clone()
    for _ in range(n):
        t = rem.clamp(-448, 448).to(torch.float8_e4m3fn)
        terms.append(t)
        rem = (rem - t.to(torch.float32)) * 16.0
    return terms


def test_three_bf16_terms_carry_fp32_coefficients():
    torch.manual_seed(0)
    theta = torch.randn(4096) * torch.exp(2 * torch.randn(4096))
    hi, mid, lo = bf16_terms(theta, 3)
    back = hi.float() + mid.float() + lo.float()
    rel = ((back - theta).abs() / theta.abs()).max().item()
    assert rel < 2.0 ** -22                      # 3 x 8 mantissa bits: fp32-grade
    two = bf16_terms(theta, 2)
    assert ((two[0].float() + two[1].float() - theta).abs() / theta.abs()).max().item() < 2.0 ** -15


def test_split_operand_gemm_matches_fp32_gemm():
    """eta = X . theta with bf16 X and a 3-term theta, accumulated in fp32, equals the fp32 product of the
    same bf16 X with the fp32 theta to ~1e-6 relative — the design matrix is the only quantised quantity."""
    torch.manual_seed(1)
    X = torch.randn(512, 256).to(torch.bfloat16)
    theta = torch.randn(256) * 0.05
    want = X.double() @ theta.double()
    got = sum(X.float() @ t.float() for t in bf16_terms(theta, 3)).double()
    scale = (X.double().abs() @ theta.double().abs()).max()
    assert ((got - want).abs().max() / scale).item() < 1e-6
    one_term = (X.float() @ theta.to(torch.bfloat16).float()).double()
    assert ((one_term - want).abs().max() / scale).item() > 1e-4       # what the split buys


@pytest.mark.parametrize("terms,bound", [(4, 2.0 ** -15), (5, 2.0 ** -19)])
def test_radix16_e4m3_expansion_error(terms, bound):
    torch.manual_seed(2)
    v = (torch.rand(8192) * 2 - 1) * 256.0       # the kernels normalise to |v| <= 256 before expanding
    parts = e4m3_expansion(v, terms)
    back = sum(p.float() * 16.0 ** -k for k, p in enumerate(parts))
    assert ((back - v).abs().max() / 256.0).item() < bound


def test_block_scaled_residuals_keep_relative_accuracy_over_orders_of_magnitude():
    """Gaussian / Poisson residuals: one power-of-two scale per 32 rows (warp max), 4-term expansion.
    The reconstruction error is relative to the block maximum whatever the magnitude."""
    torch.manual_seed(3)
    r = torch.randn(64, 32) * torch.exp(4 * torch.randn(64, 1))          # block magnitudes over ~e^12
    m = r.abs().amax(dim=1, keepdim=True)
    e = torch.ceil(torch.log2(m))                                        # m / 2^e in (0.5, 1]
    scaled = r / torch.exp2(e)
    parts = e4m3_expansion(scaled * 256.0, 4)
    back = sum(p.float() * 16.0 ** -k for k, p in enumerate(parts)) / 256.0 * torch.exp2(e)
    rel_to_block = ((back - r).abs() / m).max().item()
    assert rel_to_block < 2.0 ** -14
    # without the block scale the small blocks are lost in the representation of the big ones
    global_scale = r.abs().max()
    naive = e4m3_expansion(r / global_scale * 256.0, 4)
    naive_back = sum(p.float() * 16.0 ** -k for k, p in enumerate(naive)) / 256.0 * global_scale
    small = m.squeeze(1) < 1e-3 * global_scale
    assert small.any() and (((naive_back - r).abs() / m)[small]).max().item() > 1e-2


def test_fixed_point_accumulation_is_order_independent():
    """Intercept gradients use 40.24 fixed-point integer atomics (fed::fix_add): any order, same bits."""
    rng = np.random.default_rng(4)
    vals = rng.normal(size=10000) * 50
    fixed = np.rint(vals * 2.0 ** 24).astype(np.int64)
    orders = [rng.permutation(len(vals)) for _ in range(5)]
    sums = {int(fixed[o].sum()) for o in orders}
    assert len(sums) == 1
    assert abs(sums.pop() / 2.0 ** 24 - vals.sum()) < len(vals) * 2.0 ** -25
    float_sums = {float(np.sum(vals[o].astype(np.float32), dtype=np.float32)) for o in orders}
    assert len(float_sums) > 1                                           # what the integers avoid


def test_block_fp8_quantisation_roundtrip_on_cpu():
    """quantize_block_fp8 / dequantize_block_fp8: per-block power-of-two scales, e4m3 relative error,
    rows padded to whole 128-row tiles with 2^0, all-zero blocks stay exactly zero."""
    from pytensor_federated_b200.models.glm import dequantize_block_fp8, quantize_block_fp8

    torch.manual_seed(5)
    n, P = 300, 128
    X = torch.randn(n, P) * torch.exp(2 * torch.randn(1, P)) * torch.exp(torch.randn(n, 1))
    X[64:96, 32:64] = 0.0
    Xq, scales = quantize_block_fp8(X)
    assert Xq.dtype == torch.float8_e4m3fn and Xq.shape == (n, P)
    assert scales.dtype == torch.uint8 and scales.shape == (((n + 127) // 128) * 4, P // 32)
    assert int(scales[2, 1]) == 127 and torch.all(scales[10:] == 127)          # zero block, padding rows
    back = dequantize_block_fp8(Xq, scales)
    assert torch.all(back[64:96, 32:64] == 0)
    # per 32 x 32 block: the scale is the smallest power of two with amax / scale <= 448 ...
    for rb in range(0, n // 32):
        for fb in range(P // 32):
            blk = X[rb * 32:(rb + 1) * 32, fb * 32:(fb + 1) * 32]
            amax = blk.abs().max().item()
            if amax == 0:
                continue
            s = 2.0 ** (int(scales[rb, fb]) - 127)
            assert amax / s <= 448.0 and amax / (s / 2) > 448.0
            # ... and the element error is e4m3's half-ulp relative to the block's range
            err = (back[rb * 32:(rb + 1) * 32, fb * 32:(fb + 1) * 32] - blk).abs().max().item()
            assert err <= amax * 2.0 ** -4
    with pytest.raises(ValueError):
        quantize_block_fp8(torch.zeros(4, 48))
