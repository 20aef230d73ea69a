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


# ---- the truncated radix select of filter_finish_kernel / merge_keys_kernel (scan_tensor.cu, merge.cu) -----------------
def _ord_bits(x):
    """common.cuh ord_bits: order-preserving map float32 -> uint32."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)


def _device_select(words, k, stop_bits=10):
    """The kernels' select, statement for statement: common prefix of [umin, umax], then 8-bit histogram passes from the
    top of the varying bits, stopping once <= stop_bits low bits are unresolved.  Returns the prefix (low bits zero)."""
    words = np.asarray(words, dtype=np.uint64)
    umax, umin = int(words.max()), int(words.min())
    rem = (umax ^ umin).bit_length()                 # 32 - clz
    prefix = 0 if rem >= 32 else (umax >> rem) << rem
    krem = k
    while rem > stop_bits:
        w = min(rem, 8)
        shift = rem - w
        sel = words if rem >= 32 else words[(words >> rem) == (prefix >> rem)]
        hist = np.bincount(((sel >> shift) & ((1 << w) - 1)).astype(np.int64), minlength=1 << w)
        c = 0
        for digit in range((1 << w) - 1, -1, -1):    # the digit holding the krem-th largest
            if c + hist[digit] >= krem:
                prefix |= digit << shift
                krem -= c
                break
            c += hist[digit]
        rem = shift
    return prefix, rem


@pytest.mark.parametrize("case", ["scores", "near_ties", "all_equal", "sign_span", "tiny_range"])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_truncated_radix_select_is_a_tight_lower_bound_of_the_kth_largest(case, k):
    rng = np.random.default_rng(hash((case, k)) & 0xFFFF)
    n = 3000
    if case == "scores":
        x = rng.uniform(0.05, 0.4, n)
    elif case == "near_ties":
        x = 0.25 + rng.standard_normal(n) * 1e-6
    elif case == "all_equal":
        x = np.full(n, 0.125)
    elif case == "sign_span":
        x = rng.uniform(-1.0, 1.0, n)
    else:
        x = np.float32(0.3) + np.arange(n, dtype=np.float32) * np.float32(2e-8)
    words = _ord_bits(x.astype(np.float32))
    prefix, unresolved = _device_select(words, k)
    kth = int(np.sort(words)[::-1][k - 1])
    assert unresolved <= 10
    assert prefix <= kth                                   # a lower bound: the threshold derived from it is sound
    assert kth - prefix < (1 << 10)                        # ... and within 2^10 ulps of the exact k-th largest
    assert int((words >= prefix).sum()) >= k               # at least k entries survive the cut (merge pre-filter)
    exact, _ = _device_select(words, k, stop_bits=0)       # the untruncated select finds the k-th largest itself
    assert exact == kth


# ---- emission layout of knn_scan_shadow_kernel (scan_tensor_shadow.cu): chunks staged in the operand rings -------------
@pytest.mark.parametrize("QT,ring_bytes", [(64, 4 * 32768 + 8 * 8192), (128, 4 * 32768 + 5 * 16384)])
def test_emission_stages_every_live_entry_exactly_once(QT, ring_bytes):
    rng = np.random.default_rng(QT)
    NW, k_emit, stage_chunks = 12, 512, ring_bytes // 256
    for trial in range(50):
        nq = int(rng.integers(1, QT + 1))
        cnt = rng.integers(0, 1024, QT)
        cnt[rng.random(QT) < 0.5] //= 8                    # mostly short buffers, some long, some beyond k_emit
        chunks = np.where((np.arange(QT) < nq) & (cnt <= k_emit), (cnt + 31) >> 5, 0)
        eoff = np.concatenate([[0], np.cumsum(chunks)])
        staged = {}                                        # stage slot -> (query, buffer slot), as the kernel fills them
        slow = []
        for warp in range(NW):
            for qi in range(warp, nq, NW):
                c0, c1 = int(eoff[qi]), int(eoff[qi + 1])
                if c1 > stage_chunks:
                    continue
                for ch in range(c1 - c0):
                    for lane in range(32):
                        slot = ch * 32 + lane
                        key = (c0 + ch) * 32 + lane
                        assert key not in staged and key < stage_chunks * 32
                        staged[key] = (qi, slot) if slot < cnt[qi] else None
        for qi in range(nq):                               # the per-query path takes exactly the others
            if not (cnt[qi] <= k_emit and eoff[qi + 1] <= stage_chunks):
                slow.append(qi)
        for qi in range(nq):
            live = {(q, s) for v in staged.values() if v is not None for (q, s) in [v] if q == qi}
            if qi in slow:
                assert not live
            else:
                assert live == {(qi, s) for s in range(int(cnt[qi]))}
