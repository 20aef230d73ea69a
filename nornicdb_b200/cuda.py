"""Host-side mirror of the reference's Go package pkg/gpu/cuda (cuda_bridge.go:379-723) over the C ABI.

Same names, argument meaning and error behaviour as the Go wrappers, so the parity tests read like
pkg/gpu/cuda/cuda_test.go.  (Go is not installed in this image; INTEGRATION.md shows the cgo stub that
binds the same symbols.)  Device.Search is the one functional change: instead of NewBuffer(query) +
NewEmptyBuffer(n) + CosineSimilarity + TopK (cuda_bridge.go:643-686) it makes ONE call into the fused
batched kernel (nk_search) on the caller's device buffer — no per-query allocation, no n-float D2H."""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


# Sentinel errors, cuda_bridge.go:387-393.
class CudaError(RuntimeError):
    pass


class ErrCUDANotAvailable(CudaError):
    def __init__(self, msg="cuda: CUDA is not available on this system"):
        super().__init__(msg)


class ErrDeviceCreation(CudaError):
    pass


class ErrBufferCreation(CudaError):
    pass


class ErrKernelExecution(CudaError):
    pass


class ErrInvalidBuffer(CudaError):
    pass


# MemoryType, cuda_bridge.go:396-404.
MemoryDevice = 0
MemoryPinned = 1

# vectorspace.DistanceMetric, pkg/vectorspace/registry.go:27-31.
METRICS = {"cosine": 0, "dot": 1, "euclidean": 2}


@dataclass
class SearchResult:  # cuda_bridge.go:425-428
    Index: int
    Score: float


def _take_error() -> str:
    lib = _lib.load()
    msg = lib.cuda_get_last_error()
    text = msg.decode("utf-8", "replace") if msg else ""
    lib.cuda_clear_error()  # cuda_bridge.go:452-453
    return text


def IsAvailable() -> bool:  # cuda_bridge.go:431
    return _lib.load().cuda_is_available() != 0


def DeviceCount() -> int:  # cuda_bridge.go:436
    c = _lib.load().cuda_get_device_count()
    return 0 if c < 0 else int(c)


def _as_f32(data) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(data, dtype=np.float32).reshape(-1))


class Buffer:  # cuda_bridge.go:417-422
    def __init__(self, ptr, size: int, device: "Device"):
        self.ptr = ptr
        self.size = size
        self.device = device
        self._index = None  # cached fused-search handle over this buffer
        self._index_key = None

    def _drop_index(self):
        if self._index is not None:
            _lib.load().nk_index_release(self._index)
            self._index = None
            self._index_key = None

    def Release(self) -> None:  # cuda_bridge.go:560-565 (idempotent, no lock)
        self._drop_index()
        if self.ptr:
            _lib.load().cuda_release_buffer(self.ptr)
            self.ptr = None

    def Size(self) -> int:
        return self.size

    def DataPtr(self) -> int:
        return int(_lib.load().cuda_buffer_data(self.ptr) or 0) if self.ptr else 0

    def ReadFloat32(self, count: int) -> Optional[np.ndarray]:  # cuda_bridge.go:573-585
        if count <= 0 or count * 4 > self.size or not self.ptr:
            return None
        out = np.empty(count, dtype=np.float32)
        ret = _lib.load().cuda_buffer_copy_to_host(self.ptr, out.ctypes.data_as(C.c_void_p), count)
        return None if ret != 0 else out


class Device:  # cuda_bridge.go:407-415
    def __init__(self, ptr, device_id: int):
        lib = _lib.load()
        self.ptr = ptr
        self.id = device_id
        cc = int(lib.cuda_device_compute_capability(device_id))
        name = lib.cuda_device_name(device_id)
        self.name = name.decode("utf-8", "replace") if name else ""
        self.memory = int(lib.cuda_device_memory(device_id))
        self.ccMajor, self.ccMinor = cc // 10, cc % 10
        self.mu = threading.Lock()

    def Release(self) -> None:  # cuda_bridge.go:470-478
        with self.mu:
            if self.ptr:
                _lib.load().cuda_release_device(self.ptr)
                self.ptr = None

    def ID(self) -> int:
        return self.id

    def Name(self) -> str:
        return self.name

    def MemoryBytes(self) -> int:
        return self.memory

    def MemoryMB(self) -> int:
        return self.memory // (1024 * 1024)

    def ComputeCapability(self) -> Tuple[int, int]:
        return self.ccMajor, self.ccMinor

    def NewBuffer(self, data, memType: int = MemoryDevice) -> Buffer:  # cuda_bridge.go:506-532
        arr = _as_f32(data)
        if arr.size == 0:
            raise CudaError("cuda: cannot create empty buffer")
        with self.mu:
            ptr = _lib.load().cuda_create_buffer(self.ptr, arr.ctypes.data_as(C.c_void_p), arr.size, memType)
        if not ptr:
            raise ErrBufferCreation(f"cuda: failed to create buffer: {_take_error()}")
        return Buffer(ptr, arr.size * 4, self)

    def NewEmptyBuffer(self, count: int, memType: int = MemoryDevice) -> Buffer:  # cuda_bridge.go:535-557
        with self.mu:
            ptr = _lib.load().cuda_create_buffer(self.ptr, None, int(count), memType)
        if not ptr:
            raise ErrBufferCreation(f"cuda: failed to create buffer: {_take_error()}")
        return Buffer(ptr, int(count) * 4, self)

    def ComputeNorms(self, vectors: Buffer, norms: Buffer, n: int, dimensions: int) -> None:
        with self.mu:
            ret = _lib.load().cuda_compute_norms(self.ptr, vectors.ptr, norms.ptr, n, dimensions)
        if ret != 0:
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_take_error()}")

    def NormalizeVectors(self, vectors: Buffer, n: int, dimensions: int) -> None:  # cuda_bridge.go:587-598
        with self.mu:
            ret = _lib.load().cuda_normalize_vectors(self.ptr, vectors.ptr, n, dimensions)
        if ret != 0:
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_take_error()}")
        vectors._drop_index()

    def CosineSimilarity(self, embeddings: Buffer, query: Buffer, scores: Buffer, n: int, dimensions: int,
                         normalized: bool) -> None:  # cuda_bridge.go:600-618
        with self.mu:
            ret = _lib.load().cuda_cosine_similarity(self.ptr, embeddings.ptr, query.ptr, scores.ptr, n, dimensions,
                                                     1 if normalized else 0)
        if ret != 0:
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_take_error()}")

    def TopK(self, scores: Buffer, n: int, k: int) -> Tuple[np.ndarray, np.ndarray]:  # cuda_bridge.go:620-640
        indices = np.zeros(k, dtype=np.uint32)
        top = np.zeros(k, dtype=np.float32)
        with self.mu:
            ret = _lib.load().cuda_topk(self.ptr, scores.ptr, indices.ctypes.data_as(C.c_void_p),
                                        top.ctypes.data_as(C.c_void_p), n, k)
        if ret != 0:
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_take_error()}")
        return indices, top

    # -- fused path -------------------------------------------------------------------------------
    def _index_for(self, embeddings: Buffer, n: int, dimensions: int, metric: int):
        lib = _lib.load()
        key = (n, dimensions, metric, embeddings.DataPtr())
        if embeddings._index is not None and embeddings._index_key == key:
            return embeddings._index
        embeddings._drop_index()
        if n * dimensions * 4 > embeddings.size:
            raise ErrInvalidBuffer("cuda: invalid buffer: embeddings buffer smaller than n*dimensions")
        ids = (C.c_int * 1)(self.id)
        ix = lib.nk_index_create(ids, 1, dimensions, 0, metric)
        if not ix:
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_lib.last_error()}")
        if lib.nk_index_attach_device_rows(ix, embeddings.DataPtr(), n) != 0:
            lib.nk_index_release(ix)
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_lib.last_error()}")
        # The rows are final by the time the reference searches them (syncToCUDA = NewBuffer + NormalizeVectors, then
        # Search, gpu.go:2073-2118): build the BF16 shadow now so the drop-in route runs the fast filter path.
        if lib.nk_index_refresh_shadow(ix) != 0:
            lib.nk_index_release(ix)
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_lib.last_error()}")
        embeddings._index, embeddings._index_key = ix, key
        return ix

    def SearchBatch(self, embeddings: Buffer, queries, n: int, dimensions: int, k: int, normalized: bool = True,
                    metric: Optional[str] = None) -> Optional[List[List[SearchResult]]]:
        """Q queries in one fused launch (the batched form the Go host gains; Qdrant SearchBatch is a
        serial loop today, pkg/qdrantgrpc/points_service.go:697-725).  metric=None keeps Device.Search's
        convention: normalized=True -> raw dot (what the reference's sgemv returns), False -> true cosine."""
        if k <= 0:
            return None  # cuda_bridge.go:644-646
        if k > n:
            k = n  # cuda_bridge.go:647-649
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32).reshape(-1, dimensions))
        Q = q.shape[0]
        m = METRICS[metric] if metric is not None else (METRICS["dot"] if normalized else METRICS["cosine"])
        with self.mu:
            ix = self._index_for(embeddings, n, dimensions, m)
            idx = np.empty((Q, k), dtype=np.uint32)
            sc = np.empty((Q, k), dtype=np.float32)
            ret = _lib.load().nk_search(ix, q.ctypes.data_as(C.c_void_p), Q, k, idx.ctypes.data_as(C.c_void_p),
                                        sc.ctypes.data_as(C.c_void_p))
        if ret < 0:
            raise ErrKernelExecution(f"cuda: kernel execution failed: {_lib.last_error()}")
        return [[SearchResult(int(idx[i, j]), float(sc[i, j])) for j in range(ret)] for i in range(Q)]

    def Search(self, embeddings: Buffer, query: Sequence[float], n: int, dimensions: int, k: int,
               normalized: bool) -> Optional[List[SearchResult]]:  # cuda_bridge.go:643-686
        res = self.SearchBatch(embeddings, [query], n, dimensions, k, normalized)
        return None if res is None else res[0]


def NewDevice(deviceID: int) -> Device:  # cuda_bridge.go:445-467
    if not IsAvailable():
        raise ErrCUDANotAvailable()
    ptr = _lib.load().cuda_create_device(deviceID)
    if not ptr:
        raise ErrDeviceCreation(f"cuda: failed to create CUDA device: {_take_error()}")
    return Device(ptr, deviceID)


def HasGPUHardware() -> bool:  # cuda_bridge.go:689
    return IsAvailable()


def IsCUDACapable() -> bool:  # cuda_bridge.go:695
    return IsAvailable()


def GPUName() -> str:  # cuda_bridge.go:700-710
    if not IsAvailable():
        return ""
    try:
        d = NewDevice(0)
    except CudaError:
        return ""
    try:
        return d.Name()
    finally:
        d.Release()


def GPUMemoryMB() -> int:  # cuda_bridge.go:713-723
    if not IsAvailable():
        return 0
    try:
        d = NewDevice(0)
    except CudaError:
        return 0
    try:
        return d.MemoryMB()
    finally:
        d.Release()
