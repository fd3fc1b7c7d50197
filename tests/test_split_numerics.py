"""CPU: the numerical claim behind RP_PREC_3XF16 (rp_gemm_tc.cuh / rp_kernels.cuh), restated in numpy.

x = hi + lo with hi = rn16(x * 2^e), lo = rn16(x * 2^e - hi) carries x to ~2^-22 |x| as long as both words are normal
binary16 numbers (binary16 and tf32 share the 11-bit significand); below that the error is the absolute 2^-25 of the
subnormal grid.  With the scale 2^e chosen from the operand's maximum (expo_for), hi*hi + hi*lo + lo*hi accumulated in fp32
reproduces an fp32 dot product.  These tests pin the scale rule and the error bounds the DESIGN quotes."""
import numpy as np
import pytest


def expo_for(bound: float, H: int) -> int:
    """Mirror of rp::expo_for: bound * 2^e lands in [2^(H-1), 2^H)."""
    if not (bound > 0.0) or not np.isfinite(bound):
        return 0
    e = H - 1 - int(np.floor(np.log2(bound)))
    return max(-100, min(100, e))


def split16(x: np.ndarray, e: int):
    xs = (x.astype(np.float32) * np.float32(2.0 ** e)).astype(np.float32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


@pytest.mark.parametrize("H", [4, 12, 14])
@pytest.mark.parametrize("scale", [1e-20, 1.0, 3e7, 1e15])
def test_scale_rule_keeps_operands_in_range(H, scale):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(4096) * scale).astype(np.float32)
    e = expo_for(float(np.abs(x).max()), H)
    hi, lo = split16(x, e)
    assert np.isfinite(hi.astype(np.float32)).all() and np.isfinite(lo.astype(np.float32)).all()
    top = np.abs(x).max() * 2.0 ** e
    assert 2.0 ** (H - 1) <= top < 2.0 ** H


def test_split_representation_error():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(100000) * np.exp(3.0 * rng.standard_normal(100000))).astype(np.float32)
    e = expo_for(float(np.abs(x).max()), 14)
    hi, lo = split16(x, e)
    xs = x.astype(np.float64) * 2.0 ** e
    err = np.abs(xs - hi.astype(np.float64) - lo.astype(np.float64))
    # relative 2^-22 where both words are normal, absolute 2^-25 (half a subnormal step) otherwise
    assert np.all(err <= np.maximum(2.0 ** -22 * np.abs(xs), 2.0 ** -25) * 1.0001)
    # the range-induced part (the absolute floor of small elements) is negligible against the operand's maximum
    assert 2.0 ** -25 / np.abs(xs).max() < 2.0 ** -37
    small = np.abs(xs) < 2.0 ** -3
    assert small.any() and err[small].max() / np.abs(xs).max() < 2.0 ** -37


@pytest.mark.parametrize("case", ["randn", "positive", "wide", "spikes"])
def test_three_term_product_matches_fp32_dot(case):
    rng = np.random.default_rng(2)
    K, n = 1024, 64
    a = rng.standard_normal((n, K)).astype(np.float32)
    b = rng.standard_normal((n, K)).astype(np.float32)
    if case == "positive":
        a, b = np.abs(a) + 0.5, np.abs(b) + 0.25
    elif case == "wide":
        a = a * np.exp(4.0 * rng.standard_normal((n, K))).astype(np.float32)
    elif case == "spikes":
        b = (rng.uniform(size=(n, K)) < 0.02).astype(np.float32) + 0.3 * rng.uniform(size=(n, K)).astype(np.float32) ** 8
    ea, eb = expo_for(float(np.abs(a).max()), 14), expo_for(float(np.abs(b).max()), 14)
    ah, al = split16(a, ea)
    bh, bl = split16(b, eb)
    f = lambda t: t.astype(np.float64)
    # products of two binary16 numbers are exact in fp32; accumulate the three terms (small ones first) in fp64 here and
    # compare with the exact dot product: what is left is the split error alone
    got = ((f(al) * f(bh)).sum(1) + (f(ah) * f(bl)).sum(1) + (f(ah) * f(bh)).sum(1)) * 2.0 ** -(ea + eb)
    ref = (f(a) * f(b)).sum(1)
    norm = (np.abs(f(a)) * np.abs(f(b))).sum(1)
    assert np.max(np.abs(got - ref) / norm) < 2.0 ** -21        # dropped lo*lo term + representation error
    fp32 = (a * b).astype(np.float32).sum(1, dtype=np.float32)
    # ... which is the size of the rounding error of a plain fp32 dot product of the same data
    assert np.max(np.abs(got - ref) / norm) <= 4.0 * max(np.max(np.abs(f(fp32) - ref) / norm), 2.0 ** -23)
