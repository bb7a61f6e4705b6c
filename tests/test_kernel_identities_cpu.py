"""Arithmetic identities the CUDA kernels rely on, checked exhaustively / on dense samples with exact rational arithmetic (no GPU):

* fused cost kernels: (u8)(1.0*s/wp + 0.5) (calc_cost_sgm.cpp:404, calc_pyd_cost_sgm.cpp:428) as ONE fp16 fma on the subnormal bit
  pattern of the integer sum (csrc/cost_epi.cu box phase, csrc/pyd.cu pyd_cost_sep_kernel);
* epipolar cost kernel: clamp((int)round(v), 0, hi) (calc_cost_sgm.cpp:366-375) as the low mantissa word of 2v + (1.5*2^52 + 1)
  rounded DOWN, clamped to [0, 2*hi + 1], shifted right by one (csrc/cost_epi.cu ref_round_clamp_k);
* pyd path: (int)(1.0*n + mv + 0.5) for an integer-valued prior (calc_pyd_cost_sgm.cpp:46-47, :413-414) = n + mv, plus one where that is
  negative — the closed form behind the shift descriptors of csrc/pydl.cu (pydl_desc_kernel)."""
import math
import struct
from fractions import Fraction

import numpy as np
import pytest


def _fp16_from_bits(b):
    return float(np.array([b], dtype=np.uint16).view(np.float16)[0])


def _fma_fp16_bits(a_bits, b_bits, c_bits):
    """fma.rn.f16: exact a*b + c, one rounding to fp16 (nearest even); operands may be subnormal.  Returns the result's bits."""
    exact = Fraction(_fp16_from_bits(a_bits)) * Fraction(_fp16_from_bits(b_bits)) + Fraction(_fp16_from_bits(c_bits))
    assert exact > 0
    e = math.floor(math.log2(exact))
    while Fraction(2) ** e > exact:
        e -= 1
    while Fraction(2) ** (e + 1) <= exact:
        e += 1
    e = max(e, -14)                                   # subnormal spacing below 2^-14
    ulp = Fraction(2) ** (e - 10)
    n = exact / ulp
    fl = n.numerator // n.denominator
    fr = n - fl
    r = fl + (1 if fr > Fraction(1, 2) or (fr == Fraction(1, 2) and fl % 2 == 1) else 0)
    val = float(r * ulp)
    return int(np.array([val], dtype=np.float16).view(np.uint16)[0])


@pytest.mark.parametrize("wp,mul_bits,add_bits", [(25, 0x791F, 0x5400), (9, 0x7B1C, 0x5000)])
def test_box_normalisation_by_one_fp16_fma(wp, mul_bits, add_bits):
    for s in range(1024):                             # sums reach 24 * 25 = 600 (5x5) and 24 * 9 = 216 (3x3); a half holds up to 1023
        want = int(1.0 * s / wp + 0.5)                # the reference's expression, truncated by the store into unsigned char
        got = _fma_fp16_bits(s, mul_bits, add_bits)   # the integer s as bits = the fp16 subnormal s * 2^-24
        assert got == (add_bits | want), (s, hex(got), want)


def _round_clamp_reference(v, hi):
    """clamp((int)round(v), 0, hi) with C round() (half away from zero) and the x86 conversion (out of range / NaN -> INT_MIN)."""
    if math.isnan(v) or math.isinf(v):
        return 0
    r = math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)     # exact: Python integers
    if not (-2 ** 31 <= r < 2 ** 31):
        r = -2 ** 31
    return min(max(r, 0), hi)


def _magic_low_word(w):
    """low 32 bits of the mantissa of RD(w + (1.5 * 2^52 + 1)), as a signed 32-bit integer."""
    magic = Fraction(6755399441055745)
    exact = Fraction(w) + magic
    rd = exact.numerator // exact.denominator       # spacing 1 in [2^52, 2^53): rounding down = floor
    assert 2 ** 52 <= rd < 2 ** 53
    bits = struct.unpack("<Q", struct.pack("<d", float(rd)))[0]
    lo = bits & 0xFFFFFFFF
    return lo - (1 << 32) if lo & 0x80000000 else lo


def test_round_clamp_by_magic_number_add():
    rng = np.random.default_rng(5)
    hi_vals = [0, 1, 374, 1241, 4095]
    samples = list(rng.uniform(-3000, 3000, 4000)) + list(rng.uniform(-2.0 ** 29, 2.0 ** 29, 2000))
    samples += [k + f for k in range(-6, 1300, 7) for f in (0.0, 0.5, -0.5, 0.49999999999999994, 0.5000000000000001, 0.25)]
    samples += [0.0, -0.0, 1241.5, 1241.4999999999998, 374.5, -0.5, -0.49999999999999994, 2.0 ** 29 - 0.5, -(2.0 ** 29)]
    for v in samples:
        w = 2.0 * v                                   # exact doubling
        L = _magic_low_word(w)
        assert L == math.floor(w) + 1
        for hi in hi_vals:
            k = min(max(L, 0), 2 * hi + 1)            # VIMNMX.RELU against 2*hi + 1
            assert k >> 1 == _round_clamp_reference(v, hi), (v, hi)


def test_integer_prior_sample_coordinate():
    for n in range(-40, 40):
        for mv in range(-30, 31):
            want = int(1.0 * n + float(mv) + 0.5)     # C conversion: truncation toward zero
            t = n + mv
            assert want == t + (1 if t < 0 else 0)
