"""Host-side logic of the row-sharded multi-GPU path: shard plan, candidate-key codec and the
host k-way merge.  Pure numpy — this is orchestration around the CUDA scan, not a CPU implementation of it.

Keys are the 64-bit packing the kernels emit (csrc/common.cuh make_key): order-preserving score bits << 32 |
(0xffffffff - row), so a descending unsigned sort is exactly (score desc, row asc); euclidean search carries
-dist^2 so the same sort is (distance asc, row asc).  Key 0 marks an empty slot."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_total: int, n_shards: int, shard: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of shard `shard` (same split as nk_index_upload: g*N/G .. (g+1)*N/G)."""
    return n_total * shard // n_shards, n_total * (shard + 1) // n_shards


def pack_keys(scores: np.ndarray, rows: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(scores, dtype=np.float32)
    b = s.view(np.uint32).astype(np.uint64)
    neg = (b & np.uint64(0x80000000)) != 0
    o = np.where(neg, (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    return (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(rows, dtype=np.uint64))


def unpack_keys(keys: np.ndarray, euclidean: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    k = np.asarray(keys, dtype=np.uint64)
    o = (k >> np.uint64(32)).astype(np.uint32)
    neg = (o & np.uint32(0x80000000)) == 0
    b = np.where(neg, ~o, o & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    scores = b.view(np.float32).copy()
    rows = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.uint32)
    if euclidean:
        scores = np.sqrt(np.maximum(-scores, 0.0))
    empty = k == 0
    scores[empty] = 0.0
    rows[empty] = 0xFFFFFFFF
    return rows, scores


def merge_keys_host(keys: np.ndarray, k: int) -> np.ndarray:
    """keys [n_lists, Q, k_in] -> best k per query, sorted descending ([Q, k]); what nk_merge_keys_device does."""
    L, Q, kin = keys.shape
    flat = np.transpose(keys, (1, 0, 2)).reshape(Q, L * kin)
    order = np.sort(flat, axis=1)[:, ::-1]
    out = np.zeros((Q, k), dtype=np.uint64)
    take = min(k, order.shape[1])
    out[:, :take] = order[:, :take]
    return out
