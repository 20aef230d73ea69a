"""ctypes binding of the CPU oracle (oracle/knn_oracle.c + oracle/simd_baseline.c).

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing under nornicdb_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

COSINE, DOT, EUCLIDEAN = 0, 1, 2
METRICS = {"cosine": 0, "dot": 1, "euclidean": 2}
F32, F16 = 0, 1

_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("knn_oracle.c", "simd_baseline.c", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "clean", "all"], check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        fp, vp, sz, u64, u32, i = C.POINTER(C.c_float), C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
        for name in ("orc_dot", "orc_cosine", "orc_euclid", "orc_cosine_flat"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [vp, sz, vp, sz]
        L.orc_norm.restype = C.c_float
        L.orc_norm.argtypes = [vp, sz]
        L.orc_normalize_inplace.restype = None
        L.orc_normalize_inplace.argtypes = [vp, sz]
        for name in ("orc_batch_cosine", "orc_batch_dot", "orc_batch_euclid"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, sz, vp, sz, vp, sz]
        L.orc_batch_normalize.restype = None
        L.orc_batch_normalize.argtypes = [vp, sz, sz, sz]
        for name in ("orc_vec_cosine64", "orc_vec_dot", "orc_vec_euclid_sim"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [vp, sz, vp, sz]
        L.orc_vec_normalize.restype = None
        L.orc_vec_normalize.argtypes = [vp, sz, vp]
        L.orc_partial_sort.restype = None
        L.orc_partial_sort.argtypes = [vp, sz, vp, sz]
        L.orc_topk_insertion.restype = C.c_uint
        L.orc_topk_insertion.argtypes = [vp, C.c_uint, C.c_uint, vp, vp]
        L.orc_fill_uniform.restype = None
        L.orc_fill_uniform.argtypes = [vp, u64, u64, u64, u64]
        L.orc_fill_uniform_f16.restype = None
        L.orc_fill_uniform_f16.argtypes = [vp, u64, u64, u64, u64]
        L.orc_knn_exact64.restype = C.c_uint
        L.orc_knn_exact64.argtypes = [vp, i, u64, u32, u64, vp, u32, u32, i, vp, vp]
        L.orc_scores_exact64.restype = None
        L.orc_scores_exact64.argtypes = [vp, i, u64, u32, vp, i, vp]
        L.orc_num_threads.restype = i
        L.orc_sq_euclid64.restype = C.c_double
        L.orc_sq_euclid64.argtypes = [vp, vp, sz]
        L.orc_kmeans_assign.restype = u64
        L.orc_kmeans_assign.argtypes = [vp, u64, u32, vp, u32, i, vp]
        L.orc_kmeans_update.restype = None
        L.orc_kmeans_update.argtypes = [vp, u64, u32, vp, u32, vp, vp]
        for name in ("sb_dot", "sb_cosine", "sb_euclid"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [vp, vp, sz]
        L.sb_knn.restype = C.c_uint
        L.sb_knn.argtypes = [vp, u64, u32, vp, u32, u32, i, i, vp, vp]
        L.sb_max_threads.restype = i
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _pair(fn, a, b):
    a, b = _f32(a), _f32(b)
    return fn(_p(a), a.size, _p(b), b.size)


def dot(a, b): return float(_pair(lib().orc_dot, a, b))
def cosine(a, b): return float(_pair(lib().orc_cosine, a, b))
def euclid(a, b): return float(_pair(lib().orc_euclid, a, b))
def cosine_flat(a, b): return float(_pair(lib().orc_cosine_flat, a, b))
def vec_cosine64(a, b): return float(_pair(lib().orc_vec_cosine64, a, b))
def vec_dot(a, b): return float(_pair(lib().orc_vec_dot, a, b))
def vec_euclid_sim(a, b): return float(_pair(lib().orc_vec_euclid_sim, a, b))


def norm(v):
    v = _f32(v)
    return float(lib().orc_norm(_p(v), v.size))


def normalize_inplace(v: np.ndarray) -> np.ndarray:
    assert v.dtype == np.float32 and v.flags.c_contiguous
    lib().orc_normalize_inplace(_p(v), v.size)
    return v


def vec_normalize(v):
    v = _f32(v)
    out = np.empty_like(v)
    lib().orc_vec_normalize(_p(v), v.size, _p(out))
    return out


def batch(kind: str, emb, q, n_scores=None):
    emb, q = _f32(emb), _f32(q)
    n = emb.size // q.size if q.size else 0
    scores = np.full(n if n_scores is None else n_scores, np.nan, dtype=np.float32)
    getattr(lib(), f"orc_batch_{kind}")(_p(emb), emb.size, _p(q), q.size, _p(scores), scores.size)
    return scores


def batch_normalize(vectors: np.ndarray, n: int, dims: int) -> np.ndarray:
    assert vectors.dtype == np.float32 and vectors.flags.c_contiguous
    lib().orc_batch_normalize(_p(vectors), vectors.size, n, dims)
    return vectors


def partial_sort(scores, k):
    scores = _f32(scores)
    idx = np.arange(scores.size, dtype=np.int64)
    lib().orc_partial_sort(_p(idx), idx.size, _p(scores), k)
    return idx


def topk_insertion(scores, k):
    scores = _f32(scores)
    kk = max(int(k), 1)
    idx = np.zeros(kk, dtype=np.uint32)
    out = np.zeros(kk, dtype=np.float32)
    ke = lib().orc_topk_insertion(_p(scores), scores.size, int(k), _p(idx), _p(out))
    return idx[:ke], out[:ke]


def fill_uniform(n_rows, dim, seed, row_base=0, dtype="f32"):
    if dtype in ("f16", F16):
        out = np.empty((n_rows, dim), dtype=np.uint16)
        lib().orc_fill_uniform_f16(_p(out), n_rows, dim, seed, row_base)
        return out.view(np.float16)
    out = np.empty((n_rows, dim), dtype=np.float32)
    lib().orc_fill_uniform(_p(out), n_rows, dim, seed, row_base)
    return out


def _rows_dtype(rows):
    rows = np.ascontiguousarray(rows)
    if rows.dtype == np.float16:
        return rows, F16
    return np.ascontiguousarray(rows, dtype=np.float32), F32


def knn_exact64(rows, queries, k, metric, row_base=0):
    """fp64 brute force, (score desc, row asc) / (distance asc, row asc).  -> (idx [Q,k'], score64 [Q,k'])."""
    rows, dt = _rows_dtype(rows)
    q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
    if q.ndim == 1:
        q = q.reshape(1, -1)
    n, d = rows.shape
    Q = q.shape[0]
    kk = max(int(k), 1)
    idx = np.zeros((Q, kk), dtype=np.uint32)
    sc = np.zeros((Q, kk), dtype=np.float64)
    m = METRICS[metric] if isinstance(metric, str) else metric
    ke = lib().orc_knn_exact64(_p(rows), dt, n, d, row_base, _p(q), Q, int(k), m, _p(idx), _p(sc))
    return idx[:, :ke], sc[:, :ke]


def scores_exact64(rows, query, metric):
    rows, dt = _rows_dtype(rows)
    q = _f32(query)
    out = np.empty(rows.shape[0], dtype=np.float64)
    m = METRICS[metric] if isinstance(metric, str) else metric
    lib().orc_scores_exact64(_p(rows), dt, rows.shape[0], rows.shape[1], _p(q), m, _p(out))
    return out


def sq_euclid64(a, b) -> float:
    """squaredEuclidean, kmeans.go:430-454."""
    a, b = _f32(a), _f32(b)
    assert a.size == b.size
    return float(lib().orc_sq_euclid64(_p(a), _p(b), a.size))


def kmeans_assign(rows, centroids, assign, by_cosine=False):
    """assignToCentroids / assignToCentroidsGPU (kmeans.go:458-546).  `assign` (int32 [n]) is updated in place; returns
    the number of changed assignments."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    cen = np.ascontiguousarray(centroids, dtype=np.float32)
    assert assign.dtype == np.int32 and assign.flags.c_contiguous and assign.size == rows.shape[0]
    return int(lib().orc_kmeans_assign(_p(rows), rows.shape[0], rows.shape[1], _p(cen), cen.shape[0], int(bool(by_cosine)), _p(assign)))


def kmeans_update(rows, assign, centroids):
    """updateCentroidsWithBuffer (kmeans.go:585-618): returns (new centroids, counts); empty clusters keep their position."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    cen = np.array(centroids, dtype=np.float32, order="C", copy=True)
    a = np.ascontiguousarray(assign, dtype=np.int32)
    counts = np.zeros(cen.shape[0], dtype=np.uint32)
    lib().orc_kmeans_update(_p(rows), rows.shape[0], rows.shape[1], _p(a), cen.shape[0], _p(cen), _p(counts))
    return cen, counts


def simd_knn(rows, queries, k, metric, threads=1):
    """Reference-shaped AVX2 fp32 brute force (simd_baseline.c).  -> (idx, score32)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
    if q.ndim == 1:
        q = q.reshape(1, -1)
    n, d = rows.shape
    Q = q.shape[0]
    kk = max(int(k), 1)
    idx = np.zeros((Q, kk), dtype=np.uint32)
    sc = np.zeros((Q, kk), dtype=np.float32)
    m = METRICS[metric] if isinstance(metric, str) else metric
    ke = lib().sb_knn(_p(rows), n, d, _p(q), Q, int(k), m, int(threads), _p(idx), _p(sc))
    return idx[:, :ke], sc[:, :ke]


def max_threads() -> int:
    return int(lib().sb_max_threads())
