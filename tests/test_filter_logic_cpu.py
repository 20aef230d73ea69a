"""Soundness of the filter scan's selection logic, restated in numpy: if every approximate score
is within its row's bound B_i of the true score, then
  * per-CTA pruning at (k-th upper bound - 2 max B),
  * the cross-CTA shared threshold (any CTA's pruned threshold is a floor for every other CTA, in any order),
  * the finish step's gather of everything above (k-th upper bound of the merged list - 2 max B) and exact re-scoring
return exactly the true top-k (score desc, row asc) — for random data, near-ties and adversarial error patterns."""
import numpy as np
import pytest


def filter_topk(true_s, approx_s, B, k, n_cta, rng, prune_every=97):
    n = len(true_s)
    upper = approx_s + B
    order = rng.permutation(n_cta)                   # CTAs publish / adopt thresholds in arbitrary interleavings
    chunks = np.array_split(np.arange(n), n_cta)
    gtau = -np.inf
    lists = []
    state = {c: {"buf": [], "tau": -np.inf, "maxB": 0.0, "pos": 0} for c in range(n_cta)}
    live = True
    while live:                                      # round-robin: each CTA advances by one "tile" per turn
        live = False
        for c in order:
            st, rows = state[c], chunks[c]
            if st["pos"] >= len(rows):
                continue
            live = True
            tile = rows[st["pos"]:st["pos"] + prune_every]
            st["pos"] += prune_every
            st["maxB"] = max(st["maxB"], float(B[tile].max()))
            st["tau"] = max(st["tau"], gtau)         # adopt the shared threshold
            st["buf"].extend(int(r) for r in tile if upper[r] >= st["tau"])
            if len(st["buf"]) >= k:                  # prune: k-th upper bound minus twice the largest bound seen
                u = np.sort(upper[st["buf"]])[::-1]
                tau = max(u[k - 1] - 2.0 * st["maxB"], gtau)
                st["buf"] = [r for r in st["buf"] if upper[r] >= tau]
                st["tau"] = tau
                gtau = max(gtau, tau)                # publish
    for c in range(n_cta):                           # emission: everything above the final shared threshold
        lists.extend(r for r in state[c]["buf"] if upper[r] >= gtau)
    lists = np.array(sorted(set(lists)), dtype=np.int64)
    u = np.sort(upper[lists])[::-1]
    thr = u[k - 1] - 2.0 * float(B.max()) if len(lists) >= k else -np.inf
    cand = lists[upper[lists] >= thr]
    exact = sorted(cand, key=lambda r: (-true_s[r], r))[:k]
    return np.array(exact), len(cand)


@pytest.mark.parametrize("case", ["random", "near_ties", "adversarial_errors", "varying_bounds"])
def test_filter_selection_returns_the_true_topk(case):
    rng = np.random.default_rng(["random", "near_ties", "adversarial_errors", "varying_bounds"].index(case))
    for trial in range(12):
        n, k, n_cta = 4000, int(rng.choice([1, 10, 37])), int(rng.choice([1, 5, 16]))
        true_s = rng.standard_normal(n)
        B = np.full(n, 0.02)
        if case == "near_ties":
            true_s = np.round(true_s, 2)            # many exact and near ties around every rank
        if case == "varying_bounds":
            B = rng.uniform(0.001, 0.08, n)
        err = rng.uniform(-1, 1, n) * B
        if case == "adversarial_errors":            # push the true winners down and everybody else up, as far as allowed
            top = np.argsort(-true_s)[:k * 3]
            err = B.copy()
            err[top] = -B[top]
        approx = true_s + err
        got, n_cand = filter_topk(true_s, approx, B, k, n_cta, rng)
        want = np.array(sorted(range(n), key=lambda r: (-true_s[r], r))[:k])
        assert (got == want).all(), (case, trial, k, n_cta)
        assert n_cand < n / 4                       # and the filter does filter


def test_assignment_candidate_rule_contains_the_true_nearest_centroid():
    """assign_tensor.cu: a row keeps its 4 best UPPER bounds and its best LOWER bound over all centroids; the true winner
    (highest true score, lowest index on ties) is always among {upper >= best lower}, and if the 4th best upper also
    reaches the best lower bound the row must fall back to "all centroids"."""
    rng = np.random.default_rng(5)
    for trial in range(300):
        K = int(rng.choice([1, 2, 3, 7, 64, 300]))
        true_s = np.round(rng.standard_normal(K), int(rng.choice([1, 2, 6])))  # exact ties at low precision
        B = rng.uniform(0.0, 0.3, K) if trial % 3 else np.full(K, 0.05)
        err = rng.uniform(-1, 1, K) * B
        if trial % 5 == 0:                          # adversarial: winner pushed down, the rest up
            err = B.copy()
            err[int(np.argmax(true_s))] = -B[int(np.argmax(true_s))]
        approx = true_s + err
        up, lo = approx + B, approx - B
        top = []                                    # (upper, idx) kept sorted desc with strict insertion, like the kernel
        for ci in range(K):
            if len(top) < 4 or up[ci] > top[-1][0]:
                top.append((up[ci], ci))
                top.sort(key=lambda t: (-t[0], t[1]))
                top = top[:4]
        maxlo = lo.max()
        want = int(min(np.flatnonzero(true_s == true_s.max())))
        if len(top) == 4 and top[3][0] >= maxlo:
            cands = list(range(K))                  # "ALL"
        else:
            cands = [ci for u, ci in top[:3] if u >= maxlo]
        assert want in cands, (trial, K, want, cands)
        decided = len(top) < 2 or not (top[1][0] >= maxlo)
        if decided:
            assert top[0][1] == want
        got = min(cands, key=lambda ci: (-true_s[ci], ci))  # exact re-scoring in ascending order, strict comparison
        assert got == want
