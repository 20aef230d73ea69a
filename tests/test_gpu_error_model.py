"""The filters' error model, MEASURED on the device (VERDICT r1 "weak" #1 / "next" #6).

The BF16-shadow / FP16 / TF32 filter scans are sound iff, for every (row, query) pair, the score estimate the tensor
cores produce is within the bound the kernel adds to it:   |est - exact| <= bnd,   where bnd is built from the kernel's
own per-row and per-query factors (ra, rb, qa, qb / margin_c) and includes the allowance acc_c = d * 2^-22 + 4e-6 for
tcgen05's fp32 accumulation — a constant round 1 only checked against a sequential numpy sum.  Here the DUMP
instantiation of the scan kernels writes (est, bnd) for every pair straight from the TMEM accumulator, and the test
compares with the fp64 score on uniform, same-sign (no cancellation), huge / tiny scaled, sparse and heavy-cancellation
data at d = 32 / 1024 / 4096: > 10^7 pairs per run.  The worst observed err / bnd is printed and must stay below 1 (it
sits well below: the Cauchy-Schwarz bounds are loose by construction)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _exact(rows, q, metric):
    x = rows.astype(np.float64)
    qq = q.astype(np.float64)
    dots = x @ qq.T
    if metric == "dot":
        return dots
    xn = np.sqrt((x * x).sum(1))[:, None]
    qn = np.sqrt((qq * qq).sum(1))[None, :]
    if metric == "cosine":
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(xn * qn > 0, dots / (xn * qn), 0.0)
    return -(xn ** 2 + qn ** 2 - 2.0 * dots)  # key space of the euclidean filter: -dist^2


def _datasets(rng, n, d):
    u = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    yield "uniform", u
    yield "same-sign", np.abs(u)                                   # no cancellation: accumulation error is maximal
    yield "scaled-1e6", (u * np.float32(1e6)).astype(np.float32)
    yield "scaled-1e-6", (u * np.float32(1e-6)).astype(np.float32)
    sp = u.copy()
    sp[rng.random((n, d)) < 0.95] = 0.0                            # sparse rows (some all-zero)
    yield "sparse", sp
    c = np.abs(u)
    c[:, 1::2] = -c[:, ::2][:, : c[:, 1::2].shape[1]] * np.float32(1.0 + 2.0 ** -9)  # x_{2i+1} ~ -x_{2i}: heavy cancellation
    yield "cancelling", c.astype(np.float32)


def _queries(rng, data, Q):
    n, d = data.shape
    qs = data[rng.integers(0, n, Q)] * np.float32(1.0) + rng.uniform(-1, 1, (Q, d)).astype(np.float32) * np.float32(np.abs(data).mean() * 0.1)
    qs[0] = rng.uniform(-1, 1, d).astype(np.float32) * np.float32(np.abs(data).max())  # unrelated direction
    qs[1] = 0.0                                                                       # zero query
    return qs.astype(np.float32)


@pytest.mark.parametrize("which,dtype", [("shadow", "f32"), ("filter", "f32"), ("shadow", "f16"), ("shadow", "bf16")])
@pytest.mark.parametrize("d,n", [(32, 60_000), (1024, 40_000), (4096, 6000)])
def test_filter_estimates_stay_inside_their_bounds(knn_lib, which, dtype, d, n, capsys):
    from nornicdb_b200.knn import KnnIndex, from_bf16_bits, to_bf16_bits
    rng = np.random.default_rng(d * 7 + n)
    Q = 64
    worst, pairs = {}, 0
    for name, data in _datasets(rng, n, d):
        if dtype == "f16":
            if name == "scaled-1e6":
                continue  # outside fp16's range: such a corpus cannot be stored as fp16 in the first place
            stored = data.astype(np.float16)
            exact_rows = stored.astype(np.float32)
        elif dtype == "bf16":
            stored = to_bf16_bits(data)
            exact_rows = from_bf16_bits(stored)
        else:
            stored = data
            exact_rows = data
        q = _queries(rng, exact_rows, Q)
        for metric in ("cosine", "dot", "euclidean"):
            ix = KnnIndex(d, metric=metric, dtype=dtype)
            ix.upload(stored)
            est, bnd = ix.filter_dump(q, which)
            ix.release()
            s64 = _exact(exact_rows, q, metric)
            err = np.abs(est.astype(np.float64) - s64)
            fin = np.isfinite(est) & np.isfinite(bnd)
            assert fin.all(), (name, metric, "non-finite estimate / bound on finite data")
            viol = err > bnd.astype(np.float64)
            assert not viol.any(), (which, dtype, name, metric, d, "bound violated", float((err / np.maximum(bnd, 1e-300)).max()),
                                    np.argwhere(viol)[:4].tolist())
            with np.errstate(invalid="ignore", divide="ignore"):
                ratio = np.where(bnd > 0, err / bnd, 0.0)
            worst[(name, metric)] = float(ratio.max())
            pairs += est.size
    top = sorted(worst.items(), key=lambda t: -t[1])[:3]
    with capsys.disabled():
        print(f"\n[error model] {which}/{dtype} d={d}: {pairs:.3g} pairs, worst err/bound = {top[0][1]:.3f} "
              f"({top[0][0][0]}, {top[0][0][1]}); next {[(k[0], k[1], round(v, 3)) for k, v in top[1:]]}")
    assert top[0][1] < 1.0
