"""The filter scans are only as exact as their error bounds are rigorous.  This CPU test restates the two bounds of
DESIGN.md §3.2 / §3.3 in numpy and checks the inequalities |s_hat - s| <= bound on random, same-sign (worst case for
Cauchy-Schwarz slack) and tiny / huge magnitude data:
    TF32 (scan_tensor.cu)         |s_hat - s| <= c |x||q|,  c = 2^-10 + d 2^-22 + 4e-6
    BF16 shadow (scan_tensor_shadow.cu)   |s_hat - s| <= |dx||qb| + |x||dq| + acc_c |x||q|,  acc_c = d 2^-22 + 4e-6
s_hat is the product of the ROUNDED operands accumulated in fp32 (the tensor cores accumulate in fp32 with truncating
adders: error <= d 2^-23 relative to sum|x_i q_i|, inside the d 2^-22 the bounds allow); s is the fp64 value."""
import numpy as np
import pytest


def bf16_round(a):
    """Round-to-nearest-even to 8 significant bits, like cvt.rn.bf16x2.f32."""
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def tf32_round(a):
    """(x + 0x1000) & 0xffffe000: nearest, ties away (ptx::tf32_round_bits)."""
    u = np.asarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def fp32_dot(a, b):
    acc = np.float32(0.0)
    for p in (a.astype(np.float32) * b.astype(np.float32)):  # products of 8/11-bit operands are exact in fp32
        acc = np.float32(acc + p)
    return float(acc)


CASES = ["uniform", "same_sign", "tiny", "huge", "sparse"]


def make(case, d, rng):
    x = rng.uniform(-1, 1, d).astype(np.float32)
    q = rng.uniform(-1, 1, d).astype(np.float32)
    if case == "same_sign":
        x, q = np.abs(x), np.abs(q)
    elif case == "tiny":
        x, q = x * np.float32(1e-12), q * np.float32(3e-9)
    elif case == "huge":
        x, q = x * np.float32(1e6), q * np.float32(2e5)
    elif case == "sparse":
        x[rng.random(d) < 0.9] = 0
        q[rng.random(d) < 0.5] = 0
    return x, q


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("d", [32, 100, 1024])
def test_bf16_shadow_bound_is_rigorous(case, d):
    rng = np.random.default_rng(d * 7 + CASES.index(case))
    acc_c = d * 2.0 ** -22 + 4e-6
    for _ in range(40):
        x, q = make(case, d, rng)
        xb, qb = bf16_round(x), bf16_round(q)
        s = float(x.astype(np.float64) @ q.astype(np.float64))
        s_hat = fp32_dot(xb, qb)
        n = lambda v: float(np.sqrt((v.astype(np.float64) ** 2).sum()))
        ra, rb = n(x - xb) * 1.0001, n(x)
        qa, qbnd = n(qb) * 1.0001, (n(q - qb) + acc_c * n(q)) * 1.0001
        bound = ra * qa + rb * qbnd
        assert abs(s_hat - s) <= bound, (case, d, s, s_hat, bound)
        # and it is tight enough to be useful: well under the a-priori 2^-7 |x||q|
        if case == "uniform" and d == 1024:
            assert bound < 0.6 * 2.0 ** -7 * n(x) * n(q)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("d", [32, 100, 1024])
def test_tf32_bound_is_rigorous(case, d):
    rng = np.random.default_rng(d * 11 + CASES.index(case))
    c = 2.0 ** -10 + d * 2.0 ** -22 + 4e-6
    for _ in range(40):
        x, q = make(case, d, rng)
        s = float(x.astype(np.float64) @ q.astype(np.float64))
        s_hat = fp32_dot(tf32_round(x), tf32_round(q))
        bound = c * float(np.linalg.norm(x.astype(np.float64))) * float(np.linalg.norm(q.astype(np.float64)))
        assert abs(s_hat - s) <= bound, (case, d, s, s_hat, bound)


def test_rounding_helpers_match_the_device_definitions():
    # bf16: 1 + 2^-8 is a tie -> even (1.0); 1 + 3*2^-9 rounds up to 1 + 2^-7
    assert bf16_round(np.float32(1.0 + 2.0 ** -8)) == np.float32(1.0)
    assert bf16_round(np.float32(1.0 + 3 * 2.0 ** -9)) == np.float32(1.0 + 2.0 ** -7)
    assert bf16_round(np.float32(-2.5)) == np.float32(-2.5)
    # tf32: 10 explicit mantissa bits, ties away from zero
    assert tf32_round(np.float32(1.0 + 2.0 ** -11)) == np.float32(1.0 + 2.0 ** -10)
    assert tf32_round(np.float32(1.0 + 2.0 ** -12)) == np.float32(1.0)
