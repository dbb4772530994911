"""ctypes binding of the C-ABI CUDA library (``include/wfagpu.h``).

The library is the only compute path of this package: if ``libwfagpu.so`` is missing or no
B200 is visible, calls raise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WFAGPU_LIB") or os.path.join(HERE, "libwfagpu.so")   # env: tuning experiments

CONFIG_FIELDS = [
    "distance", "scope", "span",
    "pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free",
    "heuristic", "min_wavefront_length", "max_distance_threshold",
    "steps_between_cutoffs", "xdrop",
    "match", "mismatch", "gap_opening1", "gap_extension1", "gap_opening2", "gap_extension2",
    "max_steps", "wildcard",
]


class Config(C.Structure):
    """``wfagpu_config_t``"""
    _fields_ = [(f, C.c_int32) for f in CONFIG_FIELDS]


class BatchStats(C.Structure):
    """``wfagpu_batch_stats_t``"""
    _fields_ = [(f, C.c_int64) for f in (
        "n_pairs", "kernel_launches", "packed_bytes", "h2d_bytes", "d2h_bytes", "cells",
        "history_bytes", "retried_pairs")]


OK, EINVAL, ECUDA, ENOMEM, ENODEVICE, EUNSUPPORTED = 0, -1, -2, -3, -4, -5


class WfaGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"wfagpu error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load ``libwfagpu.so`` (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pywfa_b200.build` "
            "(pywfa_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, cp = C.c_void_p, C.c_int64, C.c_char_p
    cfgp = vp          # any structure with wfagpu_config_t's layout, passed by address
    L.wfagpu_config_default.argtypes = [cfgp]
    L.wfagpu_config_default.restype = None
    L.wfagpu_config_check.argtypes = [cfgp, i64, i64, cp, C.c_size_t]
    L.wfagpu_config_check.restype = C.c_int
    L.wfagpu_device_count.argtypes = []
    L.wfagpu_device_count.restype = C.c_int
    L.wfagpu_create.argtypes = [C.POINTER(vp), C.c_int, cp, C.c_size_t]
    L.wfagpu_create.restype = C.c_int
    L.wfagpu_destroy.argtypes = [vp]
    L.wfagpu_destroy.restype = None
    L.wfagpu_last_error.argtypes = [vp]
    L.wfagpu_last_error.restype = cp
    L.wfagpu_strerror.argtypes = [C.c_int]
    L.wfagpu_strerror.restype = cp
    seq_args = [vp, vp, vp, vp, vp, i64]          # seq, p_off, p_len, t_off, t_len, n
    out_args = [vp, vp, vp, vp, C.POINTER(vp)]    # score, status, locs, cig_off, cig_runs
    L.wfagpu_align_batch.argtypes = [vp, cfgp] + seq_args + out_args
    L.wfagpu_align_batch.restype = C.c_int
    L.wfagpu_batch_prepare.argtypes = [vp, cfgp] + seq_args + [C.POINTER(vp)]
    L.wfagpu_batch_prepare.restype = C.c_int
    L.wfagpu_batch_run.argtypes = [vp, vp, vp]
    L.wfagpu_batch_run.restype = C.c_int
    L.wfagpu_batch_fetch.argtypes = [vp, vp] + out_args
    L.wfagpu_batch_fetch.restype = C.c_int
    L.wfagpu_batch_free.argtypes = [vp, vp]
    L.wfagpu_batch_free.restype = None
    L.wfagpu_batch_get_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.wfagpu_batch_get_stats.restype = C.c_int
    L.wfagpu_last_launches.argtypes = [vp]
    L.wfagpu_last_launches.restype = C.c_int64
    i32p = C.POINTER(C.c_int32)
    L.wfagpu_align_pair.argtypes = [vp, cfgp, cp, C.c_int32, cp, C.c_int32, i32p, i32p, i32p,
                                    C.POINTER(C.POINTER(C.c_uint32)), i32p]
    L.wfagpu_align_pair.restype = C.c_int
    L.wfagpu_set_run_buffer.argtypes = [vp, vp, i64]
    L.wfagpu_set_run_buffer.restype = C.c_int
    L.wfagpu_host_alloc.argtypes = [C.c_size_t]
    L.wfagpu_host_alloc.restype = vp
    L.wfagpu_host_free.argtypes = [vp]
    L.wfagpu_host_free.restype = None
    L.wfagpu_host_register.argtypes = [vp, C.c_size_t]
    L.wfagpu_host_register.restype = C.c_int
    L.wfagpu_host_unregister.argtypes = [vp]
    L.wfagpu_host_unregister.restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    if isinstance(a, _DevicePtr):
        return C.c_void_p(a.ptr)
    return a.ctypes.data_as(C.c_void_p)


class _DevicePtr:
    """Address of a tensor that exposes ``data_ptr()`` (kept alive for the duration of the call)."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.ptr = int(tensor.data_ptr())

    def __len__(self):
        return int(self.tensor.numel())


class _PinnedBlock:
    """Owner of one ``wfagpu_host_alloc`` allocation; freed when the last array viewing it dies."""

    def __init__(self, nbytes: int):
        self.nbytes = max(int(nbytes), 1)
        self.ptr = lib().wfagpu_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"wfagpu_host_alloc({self.nbytes}) failed")

    def __del__(self):
        try:
            if self.ptr:
                lib().wfagpu_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """A numpy array in pinned host memory (``wfagpu_host_alloc``): batches built in such arrays
    and results written to them move by DMA, without a host-side staging copy."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    block = _PinnedBlock(n * dtype.itemsize)
    buf = (C.c_uint8 * block.nbytes).from_address(block.ptr)
    buf._wfagpu_owner = block           # arr.base -> buf -> block: the allocation lives as long as any view
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pinned_copy(a) -> np.ndarray:
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


class Context:
    """One ``wfagpu_ctx`` (= one CUDA device)."""

    def __init__(self, device: int = 0):
        L = lib()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = L.wfagpu_create(C.byref(h), device, err, len(err))
        if rc != OK:
            raise WfaGpuError(rc, err.value.decode() or L.wfagpu_strerror(rc).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().wfagpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise WfaGpuError(rc, lib().wfagpu_last_error(self._h).decode())

    @staticmethod
    def _inputs(seq, p_off, p_len, t_off, t_len, check=True):
        if hasattr(seq, "data_ptr"):
            # a torch tensor (uint8, contiguous): bases already resident on the device are packed in place
            if not seq.is_contiguous() or seq.element_size() != 1:
                raise ValueError("sequence tensor must be contiguous uint8")
            seq = _DevicePtr(seq)
            check = False
        else:
            seq = np.ascontiguousarray(seq, np.uint8)
        p_off = np.ascontiguousarray(p_off, np.int64)
        t_off = np.ascontiguousarray(t_off, np.int64)
        p_len = np.ascontiguousarray(p_len, np.int32)
        t_len = np.ascontiguousarray(t_len, np.int32)
        n = len(p_len)
        if not (len(p_off) == len(t_off) == len(t_len) == n):
            raise ValueError("offset/length arrays differ in length")
        if n and check:
            if (p_off < 0).any() or (t_off < 0).any() or int((p_off + p_len).max()) > len(seq) \
                    or int((t_off + t_len).max()) > len(seq):
                raise ValueError("a pair lies outside the sequence buffer")
        return seq, p_off, p_len, t_off, t_len, n

    def align_batch(self, cfg: Config, seq, p_off, p_len, t_off, t_len, copy_runs=True, check=True, out=None):
        """``wfagpu_align_batch`` on host arrays; returns a dict of numpy arrays.  With
        ``copy_runs=False`` the CIGAR run array is a view of library-owned pinned memory that
        stays valid until the next align call on this context.  ``out`` may hold preallocated
        ``score``/``status``/``locs``/``cig_off`` arrays of the right shape to be filled in place."""
        seq, p_off, p_len, t_off, t_len, n = self._inputs(seq, p_off, p_len, t_off, t_len, check)
        if out is not None:
            score, status, locs, cig_off = out["score"], out["status"], out["locs"], out["cig_off"]
            if not (score.shape == (n,) and status.shape == (n,) and locs.shape == (n, 4) and cig_off.shape == (n + 1,)
                    and score.dtype == np.int32 and status.dtype == np.int32 and locs.dtype == np.int32
                    and cig_off.dtype == np.int64 and all(a.flags.c_contiguous for a in (score, status, locs, cig_off))):
                raise ValueError("out arrays have the wrong shape or dtype")
        else:
            score = np.empty(n, np.int32)
            status = np.empty(n, np.int32)
            locs = np.empty((n, 4), np.int32)
            cig_off = np.empty(n + 1, np.int64)
        runs_p = C.c_void_p()
        rc = lib().wfagpu_align_batch(self._h, C.addressof(cfg), _ptr(seq), _ptr(p_off), _ptr(p_len),
                                      _ptr(t_off), _ptr(t_len), n, _ptr(score), _ptr(status),
                                      _ptr(locs), _ptr(cig_off), C.byref(runs_p))
        self._check(rc)
        total = int(cig_off[n])
        if total:
            runs = np.ctypeslib.as_array(C.cast(runs_p, C.POINTER(C.c_uint32)), shape=(total,))
            if copy_runs:
                runs = runs.copy()
        else:
            runs = np.zeros(0, np.uint32)
        return dict(score=score, status=status, locs=locs, cig_off=cig_off, runs=runs)

    def set_run_buffer(self, buf=None):
        """CIGAR runs of later ``align_batch`` calls land in ``buf`` (uint32, ideally pinned: e.g. this
        context's slice of a gathered array) instead of library-owned memory; ``None`` restores it."""
        if buf is None:
            self._run_buf = None
            self._check(lib().wfagpu_set_run_buffer(self._h, None, 0))
            return
        if buf.dtype != np.uint32 or not buf.flags.c_contiguous:
            raise ValueError("run buffer must be a contiguous uint32 array")
        self._run_buf = buf
        self._check(lib().wfagpu_set_run_buffer(self._h, _ptr(buf), buf.size))

    def align_pair(self, cfg: Config, pattern: bytes, text: bytes):
        """``wfagpu_align_pair``: one pair, low latency.  Returns ``(score, status, locs, runs)`` with
        ``runs`` a list of ``length << 4 | op`` words."""
        score, status, n = C.c_int32(), C.c_int32(), C.c_int32()
        locs = (C.c_int32 * 4)()
        runs = C.POINTER(C.c_uint32)()
        rc = lib().wfagpu_align_pair(self._h, C.addressof(cfg), pattern, len(pattern), text, len(text),
                                     C.byref(score), C.byref(status), locs, C.byref(runs), C.byref(n))
        self._check(rc)
        return score.value, status.value, list(locs), runs[:n.value]

    def last_launches(self) -> int:
        return int(lib().wfagpu_last_launches(self._h))

    def prepare(self, cfg: Config, seq, p_off, p_len, t_off, t_len) -> "Batch":
        seq, p_off, p_len, t_off, t_len, n = self._inputs(seq, p_off, p_len, t_off, t_len)
        h = C.c_void_p()
        rc = lib().wfagpu_batch_prepare(self._h, C.addressof(cfg), _ptr(seq), _ptr(p_off), _ptr(p_len),
                                        _ptr(t_off), _ptr(t_len), n, C.byref(h))
        self._check(rc)
        return Batch(self, h, n)


class Batch:
    """A packed, HBM-resident batch (``wfagpu_batch``): ``run`` launches kernels only."""

    def __init__(self, ctx: Context, h, n: int):
        self.ctx, self._h, self.n = ctx, h, n

    def run(self, stream=None):
        self.ctx._check(lib().wfagpu_batch_run(self.ctx._h, self._h, C.c_void_p(stream) if stream else None))

    def fetch(self, cigars: bool = True):
        n = self.n
        score = np.empty(n, np.int32)
        status = np.empty(n, np.int32)
        locs = np.empty((n, 4), np.int32)
        cig_off = np.empty(n + 1, np.int64)
        runs_p = C.c_void_p()
        rc = lib().wfagpu_batch_fetch(self.ctx._h, self._h, _ptr(score), _ptr(status), _ptr(locs),
                                      _ptr(cig_off), C.byref(runs_p) if cigars else None)
        self.ctx._check(rc)
        total = int(cig_off[n])
        if cigars and total:
            runs = np.ctypeslib.as_array(C.cast(runs_p, C.POINTER(C.c_uint32)), shape=(total,)).copy()
        else:
            runs = np.zeros(0, np.uint32)
        return dict(score=score, status=status, locs=locs, cig_off=cig_off, runs=runs)

    def stats(self) -> dict:
        s = BatchStats()
        lib().wfagpu_batch_get_stats(self._h, C.byref(s))
        return {f: int(getattr(s, f)) for f, _ in BatchStats._fields_}

    def free(self):
        if self._h:
            lib().wfagpu_batch_free(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass
