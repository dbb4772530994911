"""ctypes bridge to the parity checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
``pywfa_b200`` never does.

Two checkers share one interface (``align_batch``):

* ``kind="port"``       -- ``oracle/liboracle.so``, our CPU restatement
  (``oracle/wfa_oracle.c``).
* ``kind="reference"``  -- ``oracle/_ref/libwfa_ref.so``, the unmodified
  reference WFA2-lib compiled from /root/reference plus ``ref_harness.c``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libwfa_ref.so")
REF_PYWFA_DIR = os.path.join(HERE, "_ref")

_FIELDS = [
    "distance", "scope", "span",
    "pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free",
    "heuristic", "min_wavefront_length", "max_distance_threshold",
    "steps_between_cutoffs", "xdrop",
    "match", "mismatch", "gap_opening1", "gap_extension1", "gap_opening2", "gap_extension2",
    "max_steps", "wildcard",
]


class Config(C.Structure):
    """Mirror of ``wfagpu_config_t`` (include/wfagpu.h)."""
    _fields_ = [(f, C.c_int32) for f in _FIELDS]


_DIST = {"affine": 0, "affine2p": 1, "linear": 2, "levenshtein": 3, "indel": 4}
_SCOPE = {"score": 0, "full": 1}
_SPAN = {"end-to-end": 0, "ends-free": 1}
_HEUR = {None: 0, "adaptive": 1, "X-drop": 2}


def make_config(distance="affine", match=0, mismatch=4, gap_opening=6, gap_extension=2,
                gap_opening2=24, gap_extension2=1, scope="full", span="ends-free",
                pattern_begin_free=0, pattern_end_free=0, text_begin_free=0, text_end_free=0,
                heuristic=None, min_wavefront_length=10, max_distance_threshold=50,
                steps_between_cutoffs=1, xdrop=20, max_steps=0, wildcard=None) -> Config:
    """kwargs with the names/defaults of pywfa's constructor (pywfa/align.pyx:309-334)."""
    return Config(
        distance=_DIST[distance], scope=_SCOPE[scope], span=_SPAN[span],
        pattern_begin_free=pattern_begin_free, pattern_end_free=pattern_end_free,
        text_begin_free=text_begin_free, text_end_free=text_end_free,
        heuristic=_HEUR[heuristic], min_wavefront_length=min_wavefront_length,
        max_distance_threshold=max_distance_threshold,
        steps_between_cutoffs=steps_between_cutoffs, xdrop=xdrop,
        match=match, mismatch=mismatch, gap_opening1=gap_opening, gap_extension1=gap_extension,
        gap_opening2=gap_opening2, gap_extension2=gap_extension2, max_steps=max_steps,
        wildcard=ord(wildcard.upper()) if wildcard else 0)


def build(ref: bool | None = None) -> None:
    """Compile the checkers (``make -C oracle``).  ``oracle/_ref`` is only (re)built when
    /root/reference is present; elsewhere the prebuilt files are used as they are."""
    targets = ["oracle"]
    if ref is None:
        ref = os.path.isdir("/root/reference/pywfa/WFA2_lib")
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True,
                   stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def checker_kind() -> str:
    """The strongest checker available: the unmodified reference when its build travelled with
    the snapshot (oracle/_ref), else the restatement."""
    return "reference" if have_ref() else "port"


_i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")

_libs: dict = {}


def _load(kind: str):
    if kind in _libs:
        return _libs[kind]
    if kind == "port":
        if not os.path.exists(PORT_SO):
            build(ref=False)
        lib = C.CDLL(PORT_SO)
        lib.oracle_align_batch.restype = C.c_int
        lib.oracle_align_batch.argtypes = [
            C.POINTER(Config), C.c_int, _u8p, _i64p, _i32p, _i64p, _i32p, C.c_int64,
            _i32p, _i32p, _i64p, _u8p, C.c_int64, _i64p]
    elif kind == "reference":
        lib = C.CDLL(REF_SO)
        lib.ref_new.restype = C.c_void_p
        lib.ref_new.argtypes = [C.POINTER(Config), C.c_int]
        lib.ref_delete.argtypes = [C.c_void_p]
        lib.ref_align_batch.restype = C.c_int
        lib.ref_align_batch.argtypes = [
            C.c_void_p, _u8p, _i64p, _i32p, _i64p, _i32p, C.c_int64,
            _i32p, _i32p, _i64p, _u8p, C.c_int64, _i64p]
    else:
        raise ValueError(kind)
    _libs[kind] = lib
    return lib


_MEMORY_MODE = {"high": 0, "medium": 1, "low": 2, "biwfa": 3}

# op character -> SAM code, pywfa/align.pyx:11-14
_OPCODE = np.zeros(256, np.uint8)
for _ch, _code in (("M", 0), ("I", 1), ("D", 2), ("X", 8)):
    _OPCODE[ord(_ch)] = _code


def ops_to_runs(ops: np.ndarray, ops_off: np.ndarray):
    """Raw per-base operation characters -> (cig_off, runs) with runs = len<<4 | SAM op,
    the run-length encoding pywfa's ``cigartuples`` performs (pywfa/align.pyx:759-786)."""
    n = len(ops_off) - 1
    total = int(ops_off[-1])
    if total == 0:
        return np.zeros(n + 1, np.int64), np.zeros(0, np.uint32)
    o = ops[:total]
    start = np.ones(total, bool)
    start[1:] = o[1:] != o[:-1]
    firsts = ops_off[:-1][ops_off[:-1] < total]
    start[firsts] = True
    idx = np.flatnonzero(start)
    lens = np.diff(np.append(idx, total)).astype(np.uint32)
    runs = (lens << np.uint32(4)) | _OPCODE[o[idx]].astype(np.uint32)
    cig_off = np.searchsorted(idx, ops_off, side="left").astype(np.int64)
    return cig_off, runs


def locations_from_runs(cig_off, runs, p_len, t_len, scope_full=True) -> np.ndarray:
    """pattern_start, pattern_end, text_start, text_end as pywfa's ``locations`` property
    computes them (pywfa/align.pyx:788-833).  Small-batch Python loop (checker only)."""
    n = len(p_len)
    locs = np.zeros((n, 4), np.int32)
    if not scope_full:
        return locs
    for i in range(n):
        r = runs[cig_off[i]:cig_off[i + 1]]
        if len(r) == 0 or t_len[i] == 0 or p_len[i] == 0:
            continue
        ps = ts = 0
        for w in r:
            op, ln = int(w) & 15, int(w) >> 4
            if op == 0:
                if ln >= 1:
                    break
            elif op == 2:
                ps += ln
            elif op == 8:
                ts += ln
                ps += ln
            elif op == 1:
                ts += ln
        pe, te = int(p_len[i]), int(t_len[i])
        for w in r[::-1]:
            op, ln = int(w) & 15, int(w) >> 4
            if op == 0:
                if ln >= 1:
                    break
            elif op == 2:
                pe -= ln
            elif op == 8:
                pe -= ln
                te -= ln
            elif op == 1:
                te -= ln
        locs[i] = (ps, pe, ts, te)
    return locs


def align_batch(cfg: Config, seq: np.ndarray, p_off, p_len, t_off, t_len, kind="port",
                bt_mode=0, memory_mode="high"):
    """Run a checker over a batch.  Returns a dict with score, status, cig_off, runs,
    locs (n,4), cells (per pair), ops/ops_off (raw op characters)."""
    lib = _load(kind)
    seq = np.ascontiguousarray(seq, np.uint8)
    p_off = np.ascontiguousarray(p_off, np.int64)
    t_off = np.ascontiguousarray(t_off, np.int64)
    p_len = np.ascontiguousarray(p_len, np.int32)
    t_len = np.ascontiguousarray(t_len, np.int32)
    n = len(p_len)
    if cfg.span == 1 and n:
        # the reference exit(1)s on free ends longer than a sequence (W/wavefront/wavefront_align.c:89-100);
        # a checker must refuse them too instead of reading outside the DP matrix
        if (max(cfg.pattern_begin_free, cfg.pattern_end_free) > int(p_len.min())
                or max(cfg.text_begin_free, cfg.text_end_free) > int(t_len.min())):
            raise ValueError("ends-free parameters larger than a sequence of the batch")
    score = np.zeros(n, np.int32)
    status = np.zeros(n, np.int32)
    cells = np.zeros(n, np.int64)
    ops_off = np.zeros(n + 1, np.int64)
    cap = int(p_len.astype(np.int64).sum() + t_len.astype(np.int64).sum()) + 16
    ops = np.zeros(cap, np.uint8)
    if kind == "port":
        rc = lib.oracle_align_batch(C.byref(cfg), bt_mode, seq, p_off, p_len, t_off, t_len, n,
                                    score, status, ops_off, ops, cap, cells)
    else:
        h = lib.ref_new(C.byref(cfg), _MEMORY_MODE[memory_mode])
        try:
            rc = lib.ref_align_batch(h, seq, p_off, p_len, t_off, t_len, n,
                                     score, status, ops_off, ops, cap, cells)
        finally:
            lib.ref_delete(h)
    if rc != 0:
        raise RuntimeError(f"{kind} checker failed rc={rc}")
    cig_off, runs = ops_to_runs(ops, ops_off)
    locs = locations_from_runs(cig_off, runs, p_len, t_len, cfg.scope == 1) if n <= 20000 else None
    return dict(score=score, status=status, cig_off=cig_off, runs=runs, locs=locs,
                cells=cells, ops=ops[:int(ops_off[-1])], ops_off=ops_off)


def runs_to_cigarstring(runs) -> str:
    return "".join(f"{int(w) >> 4}{'MIDNSHP=XB'[int(w) & 15]}" for w in runs)
