"""ctypes binding of the C-ABI library libsangnom_cuda.so (include/sangnom_cuda.h).

This is the Python face of the product path: there is no fallback - if the library is missing or
no B200 is present, `Context(...)` raises. numpy arrays are host buffers; `process_device` takes raw
device pointers (e.g. torch tensors' data_ptr()) for the device-resident path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_LIB = os.path.join(_PKG_DIR, "libsangnom_cuda.so")

SN_OK, SN_ERR_INVALID, SN_ERR_CUDA, SN_ERR_UNSUPPORTED, SN_ERR_NOMEM = range(5)
MODE_COPY, MODE_FIELD, MODE_DH, MODE_INPLACE = range(4)
ABI_VERSION = 2
DEVICE_ALL = -1
FLAG_PERSISTENT_POOL = 1
FLAG_SATURATE = 2

EXPORTS = [
    "sangnom_cuda_abi_version", "sangnom_cuda_create", "sangnom_cuda_destroy", "sangnom_cuda_process_planes",
    "sangnom_cuda_process_planes_device", "sangnom_cuda_synchronize", "sangnom_cuda_threshold",
    "sangnom_cuda_get_limits", "sangnom_cuda_get_stats", "sangnom_cuda_reset_stats", "sangnom_cuda_host_alloc",
    "sangnom_cuda_host_free", "sangnom_cuda_last_error", "sangnom_cuda_submit", "sangnom_cuda_wait",
    "sangnom_cuda_chain_create", "sangnom_cuda_chain_destroy", "sangnom_cuda_chain_process", "sangnom_cuda_chain_get_stats",
    "sangnom_cuda_chain_last_error", "sangnom_cuda_turn_planes_device",
    "sangnom_cuda_host_pin", "sangnom_cuda_host_unpin", "sangnom_cuda_device_count",
]
TURN_TRANSPOSE, TURN_RIGHT_LEFT, TURN_LEFT_RIGHT = range(3)
TURN_DST_PADDING_WRITABLE = 1


class SnConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("sample_type", C.c_int), ("pool_width", C.c_int),
                ("pool_height", C.c_int), ("max_frames_in_flight", C.c_int), ("flags", C.c_int),
                ("device_mask", C.c_ulonglong), ("copy_threads", C.c_int)]


class SnPlaneJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_pitch", C.c_ssize_t), ("dst", C.c_void_p), ("dst_pitch", C.c_ssize_t),
                ("width", C.c_int), ("dst_height", C.c_int), ("offset", C.c_int), ("mode", C.c_int),
                ("threshold", C.c_float), ("plane", C.c_int), ("frame", C.c_int)]


class SnChainConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("sample_type", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("turn", C.c_int), ("max_frames_in_flight", C.c_int), ("flags", C.c_int)]


class SnChainJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_pitch", C.c_ssize_t), ("dst", C.c_void_p), ("dst_pitch", C.c_ssize_t),
                ("width", C.c_int), ("height", C.c_int), ("offset1", C.c_int), ("offset2", C.c_int),
                ("threshold", C.c_float), ("plane", C.c_int), ("frame", C.c_int)]


class SnChainStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pass_kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("frames", C.c_uint64)]


class SnTurnPlane(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_pitch", C.c_ssize_t), ("dst", C.c_void_p), ("dst_pitch", C.c_ssize_t),
                ("width", C.c_int), ("height", C.c_int), ("flags", C.c_int)]


class SnLimits(C.Structure):
    _fields_ = [("max_pool_width", C.c_int * 5), ("sm_count", C.c_int), ("compute_major", C.c_int), ("compute_minor", C.c_int)]


class SnStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("planes_processed", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("frames", C.c_uint64), ("host_copy_bytes", C.c_uint64)]


class SangNomCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsangnom_cuda error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the C-ABI library. Raises if it was not built - there is no other compute path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(CUDA_LIB):
        raise FileNotFoundError(f"{CUDA_LIB} missing - run __graft_entry__.build()")
    L = C.CDLL(CUDA_LIB)
    L.sangnom_cuda_abi_version.restype = C.c_int
    L.sangnom_cuda_create.restype = C.c_int
    L.sangnom_cuda_create.argtypes = [C.POINTER(SnConfig), C.POINTER(C.c_void_p)]
    L.sangnom_cuda_destroy.restype = None
    L.sangnom_cuda_destroy.argtypes = [C.c_void_p]
    L.sangnom_cuda_process_planes.restype = C.c_int
    L.sangnom_cuda_process_planes.argtypes = [C.c_void_p, C.POINTER(SnPlaneJob), C.c_int]
    L.sangnom_cuda_submit.restype = C.c_int
    L.sangnom_cuda_submit.argtypes = [C.c_void_p, C.POINTER(SnPlaneJob), C.c_int, C.POINTER(C.c_uint64)]
    L.sangnom_cuda_wait.restype = C.c_int
    L.sangnom_cuda_wait.argtypes = [C.c_void_p, C.c_uint64]
    L.sangnom_cuda_chain_create.restype = C.c_int
    L.sangnom_cuda_chain_create.argtypes = [C.POINTER(SnChainConfig), C.POINTER(C.c_void_p)]
    L.sangnom_cuda_chain_destroy.restype = None
    L.sangnom_cuda_chain_destroy.argtypes = [C.c_void_p]
    L.sangnom_cuda_chain_process.restype = C.c_int
    L.sangnom_cuda_chain_process.argtypes = [C.c_void_p, C.POINTER(SnChainJob), C.c_int]
    L.sangnom_cuda_chain_get_stats.restype = C.c_int
    L.sangnom_cuda_chain_get_stats.argtypes = [C.c_void_p, C.POINTER(SnChainStats)]
    L.sangnom_cuda_chain_last_error.restype = C.c_char_p
    L.sangnom_cuda_chain_last_error.argtypes = [C.c_void_p]
    L.sangnom_cuda_turn_planes_device.restype = C.c_int
    L.sangnom_cuda_turn_planes_device.argtypes = [C.c_int, C.c_int, C.POINTER(SnTurnPlane), C.c_int, C.c_void_p]
    L.sangnom_cuda_process_planes_device.restype = C.c_int
    L.sangnom_cuda_process_planes_device.argtypes = [C.c_void_p, C.POINTER(SnPlaneJob), C.c_int, C.c_void_p]
    L.sangnom_cuda_synchronize.restype = C.c_int
    L.sangnom_cuda_synchronize.argtypes = [C.c_void_p]
    L.sangnom_cuda_threshold.restype = C.c_float
    L.sangnom_cuda_threshold.argtypes = [C.c_int, C.c_int, C.c_int]
    L.sangnom_cuda_get_limits.restype = C.c_int
    L.sangnom_cuda_get_limits.argtypes = [C.c_int, C.POINTER(SnLimits)]
    L.sangnom_cuda_get_stats.restype = C.c_int
    L.sangnom_cuda_get_stats.argtypes = [C.c_void_p, C.POINTER(SnStats)]
    L.sangnom_cuda_reset_stats.restype = None
    L.sangnom_cuda_reset_stats.argtypes = [C.c_void_p]
    L.sangnom_cuda_host_alloc.restype = C.c_void_p
    L.sangnom_cuda_host_alloc.argtypes = [C.c_size_t]
    L.sangnom_cuda_host_free.restype = None
    L.sangnom_cuda_host_free.argtypes = [C.c_void_p]
    L.sangnom_cuda_last_error.restype = C.c_char_p
    L.sangnom_cuda_last_error.argtypes = [C.c_void_p]
    L.sangnom_cuda_host_pin.restype = C.c_int
    L.sangnom_cuda_host_pin.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.sangnom_cuda_host_unpin.restype = C.c_int
    L.sangnom_cuda_host_unpin.argtypes = [C.c_void_p, C.c_void_p]
    L.sangnom_cuda_device_count.restype = C.c_int
    L.sangnom_cuda_device_count.argtypes = [C.c_void_p]
    if L.sangnom_cuda_abi_version() != ABI_VERSION:
        raise RuntimeError("libsangnom_cuda ABI version mismatch")
    _lib = L
    return L


def threshold(aa, bits, sample_bytes):
    return float(load().sangnom_cuda_threshold(int(aa), int(bits), int(sample_bytes)))


def resolve_offset(order, parity):
    """order 0: keep the field the frame's parity names; 1: keep top; 2: keep bottom
    (reference SangNom2.cpp:336-341)."""
    if order == 0:
        return 0 if parity else 1
    return 0 if order == 1 else 1


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc memory (freed when the array is garbage collected)."""
    L = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = L.sangnom_cuda_host_alloc(max(nbytes, 1))
    if not ptr:
        raise MemoryError("cudaHostAlloc failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr)

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            try:
                L.sangnom_cuda_host_free(self.p)
            except Exception:
                pass
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _owners[arr.__array_interface__["data"][0]] = _Owner(ptr)
    return arr


_owners = {}


class PinnedArena:
    """One cudaHostAlloc allocation handed out as consecutive numpy planes - how a batching host layer
    stages frames: planes that are neighbours in host memory travel in one DMA transfer."""

    def __init__(self, nbytes):
        self.buf = pinned_empty((max(int(nbytes), 1),), np.uint8)
        self.pos = 0

    def take(self, shape, dtype, align=16):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        start = (self.pos + align - 1) // align * align
        if start + nbytes > self.buf.size:
            raise MemoryError("pinned arena exhausted")
        self.pos = start + nbytes
        return self.buf[start:start + nbytes].view(dtype).reshape(shape)


def make_job(src_ptr, src_pitch, dst_ptr, dst_pitch, width, dst_height, offset, mode, thr, plane, frame):
    return SnPlaneJob(src_ptr, src_pitch, dst_ptr, dst_pitch, width, dst_height, offset, mode, thr, plane, frame)


class Context:
    """sn_ctx wrapper. pool_width/pool_height are the OUTPUT luma dims (after dh).
    device: ordinal, DEVICE_ALL, or a list of ordinals (one host pipeline per device behind one context)."""

    def __init__(self, sample_bytes, pool_width, pool_height, device=0, max_frames_in_flight=0, flags=0, copy_threads=0):
        L = load()
        mask = 0
        if isinstance(device, (list, tuple)):
            for d in device:
                mask |= 1 << int(d)
            device = int(device[0])
        cfg = SnConfig(ABI_VERSION, device, sample_bytes, pool_width, pool_height, max_frames_in_flight, flags, mask, copy_threads)
        h = C.c_void_p()
        rc = L.sangnom_cuda_create(C.byref(cfg), C.byref(h))
        if rc != SN_OK:
            raise SangNomCudaError(rc, L.sangnom_cuda_last_error(None).decode())
        self._h = h
        self.sample_bytes = sample_bytes
        self.pool_width, self.pool_height = pool_width, pool_height

    def close(self):
        if self._h:
            load().sangnom_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != SN_OK:
            raise SangNomCudaError(rc, load().sangnom_cuda_last_error(self._h).decode())

    def process_jobs(self, jobs):
        arr = (SnPlaneJob * len(jobs))(*jobs)
        self._check(load().sangnom_cuda_process_planes(self._h, arr, len(jobs)))

    def submit(self, jobs):
        """Queue a batch (host buffers) and return its ticket; the buffers must stay alive until wait(ticket)."""
        arr = jobs if isinstance(jobs, C.Array) else (SnPlaneJob * len(jobs))(*jobs)
        t = C.c_uint64()
        self._check(load().sangnom_cuda_submit(self._h, arr, len(arr), C.byref(t)))
        return int(t.value)

    def wait(self, ticket):
        self._check(load().sangnom_cuda_wait(self._h, int(ticket)))

    def process_jobs_device(self, jobs, stream=None):
        """stream: a cudaStream_t handle as int (0 = CUDA's legacy default stream); None = the context's own stream."""
        arr = jobs if isinstance(jobs, C.Array) else (SnPlaneJob * len(jobs))(*jobs)
        handle = C.c_void_p(-1) if stream is None else C.c_void_p(stream)
        self._check(load().sangnom_cuda_process_planes_device(self._h, arr, len(arr), handle))

    def synchronize(self):
        self._check(load().sangnom_cuda_synchronize(self._h))

    def device_count(self):
        return int(load().sangnom_cuda_device_count(self._h))

    def host_pin(self, array):
        """Pin the memory of a (long-lived, contiguous) numpy array the caller owns; planes inside it then travel by DMA."""
        self._check(load().sangnom_cuda_host_pin(self._h, array.ctypes.data, array.nbytes))

    def host_unpin(self, array):
        self._check(load().sangnom_cuda_host_unpin(self._h, array.ctypes.data))

    def stats(self):
        s = SnStats()
        self._check(load().sangnom_cuda_get_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in SnStats._fields_}

    def reset_stats(self):
        load().sangnom_cuda_reset_stats(self._h)

    # ---- convenience: whole frames as numpy planes (host path) ------------------------------
    def frame_jobs(self, src_planes, dst_planes, bits, order=1, aa=48, aac=0, dh=False, luma=True, chroma=True,
                   parity=True, frame_key=0):
        """Jobs for one frame, mirroring the reference's GetFrame plane loop (SangNom2.cpp:346-394)."""
        off = resolve_offset(order, parity)
        sb = self.sample_bytes
        jobs = []
        for p, (s, d) in enumerate(zip(src_planes, dst_planes)):
            if p == 3:
                # alpha: the reference never writes it (:346-348); we copy it (row-doubled for dh)
                jobs.append(make_job(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0] * (2 if dh else 1), s.shape[1],
                                     s.shape[0], 0, MODE_COPY, 0.0, 3, frame_key))
                if dh:
                    jobs.append(make_job(s.ctypes.data, s.strides[0], d.ctypes.data + d.strides[0], d.strides[0] * 2, s.shape[1],
                                         s.shape[0], 0, MODE_COPY, 0.0, 3, frame_key))
                continue
            enabled = dh or (luma if p == 0 else chroma)
            thr = threshold(aa if p == 0 else aac, bits, sb)
            mode = MODE_DH if dh else (MODE_FIELD if enabled else MODE_COPY)
            jobs.append(make_job(s.ctypes.data, s.strides[0], d.ctypes.data, d.strides[0], s.shape[1], d.shape[0], off, mode, thr,
                                 p, frame_key))
        return jobs

    def process_frames(self, frames, bits, order=1, aa=48, aac=0, dh=False, luma=True, chroma=True, parities=None):
        """frames: list of plane lists (numpy). Returns list of output plane lists."""
        outs, jobs, keep = [], [], []
        for k, planes in enumerate(frames):
            srcs = [np.ascontiguousarray(p) for p in planes]
            dsts = [np.empty((p.shape[0] * (2 if dh else 1), p.shape[1]), dtype=p.dtype) for p in srcs]
            keep.append(srcs)
            par = True if parities is None else parities[k]
            jobs += self.frame_jobs(srcs, dsts, bits, order, aa, aac, dh, luma, chroma, par, frame_key=k)
            outs.append(dsts)
        self.process_jobs(jobs)
        return outs


class Chain:
    """sn_chain wrapper: SangNom2(dh=true) -> turn -> SangNom2(dh=true) -> turn back on the device.
    width/height: INPUT luma size; every frame comes back as 2*width x 2*height."""

    def __init__(self, sample_bytes, width, height, turn=TURN_TRANSPOSE, device=0, max_frames_in_flight=0):
        L = load()
        cfg = SnChainConfig(ABI_VERSION, device, sample_bytes, width, height, turn, max_frames_in_flight, 0)
        h = C.c_void_p()
        rc = L.sangnom_cuda_chain_create(C.byref(cfg), C.byref(h))
        if rc != SN_OK:
            raise SangNomCudaError(rc, L.sangnom_cuda_chain_last_error(None).decode())
        self._h = h
        self.sample_bytes = sample_bytes

    def close(self):
        if self._h:
            load().sangnom_cuda_chain_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process_jobs(self, jobs):
        arr = jobs if isinstance(jobs, C.Array) else (SnChainJob * len(jobs))(*jobs)
        rc = load().sangnom_cuda_chain_process(self._h, arr, len(arr))
        if rc != SN_OK:
            raise SangNomCudaError(rc, load().sangnom_cuda_chain_last_error(self._h).decode())

    def stats(self):
        s = SnChainStats()
        load().sangnom_cuda_chain_get_stats(self._h, C.byref(s))
        return {k: int(getattr(s, k)) for k, _ in SnChainStats._fields_}

    def process_frames(self, frames, bits, aa=48, aac=0, offset1=0, offset2=0):
        """frames: list of plane lists (numpy, Y[,U,V]). Returns the 2W x 2H output planes per frame."""
        outs, jobs, keep = [], [], []
        for k, planes in enumerate(frames):
            srcs = [np.ascontiguousarray(p) for p in planes[:3]]
            dsts = [np.empty((2 * p.shape[0], 2 * p.shape[1]), dtype=p.dtype) for p in srcs]
            keep.append(srcs)
            for p, (s_, d) in enumerate(zip(srcs, dsts)):
                thr = threshold(aa if p == 0 else aac, bits, self.sample_bytes)
                jobs.append(SnChainJob(s_.ctypes.data, s_.strides[0], d.ctypes.data, d.strides[0], s_.shape[1], s_.shape[0],
                                       offset1, offset2, thr, p, k))
            outs.append(dsts)
        self.process_jobs(jobs)
        return outs
