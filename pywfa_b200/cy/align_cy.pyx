# cython: language_level=3
# align_cy.pyx -- the Cython host layer over the C ABI: the shape a pywfa maintainer's align.pyx takes
# when `wavefront_align` (pywfa/align.pyx:421-443) calls libwfagpu instead of WFA2-lib.  One cdef class
# with pywfa's constructor kwargs; the alignment is one nogil call of wfagpu_align_pair (same arguments as
# wavefront_align) and the result properties read what that call filled (pywfa/align.pyx:731-833 read aligner->cigar).
# The configuration checks, AlignmentResult and the CIGAR post-processing are shared with the ctypes
# host layer (pywfa_b200/align.py), which mirrors pywfa/align.pyx:17-295,309-419.
from libc.stdint cimport int32_t, int64_t, uint8_t, uint32_t
from libc.string cimport memcpy

cimport wfagpu as gpu

from pywfa_b200.align import (AlignmentResult, WavefrontAligner as _HostAligner, cigartuples_to_str,
                              clip_cigartuples, elide_mismatches_from_cigar)


cdef class WavefrontAligner:
    """pywfa.WavefrontAligner on the B200: same constructor, ``wavefront_align`` / ``__call__``,
    ``score`` / ``status`` / ``cigartuples`` / ``cigarstring`` / ``locations``."""
    cdef gpu.wfagpu_ctx* _ctx
    cdef gpu.wfagpu_config_t _cfg
    cdef object _host                 # configuration surface (properties, validation) of the ctypes layer
    cdef bytes _pattern
    cdef str _text
    cdef readonly int pattern_len, text_len
    cdef int32_t _score, _status
    cdef int32_t _locs[4]
    cdef list _cigartuples

    def __cinit__(self):
        self._ctx = NULL

    def __init__(self, pattern=None, device=0, **kwargs):
        cdef char err[512]
        self._host = _HostAligner(pattern, device=device, **kwargs)      # raises like pywfa on bad kwargs
        raw = bytes(self._host._cfg)
        assert len(raw) == sizeof(gpu.wfagpu_config_t)
        memcpy(&self._cfg, <const char*>raw, sizeof(gpu.wfagpu_config_t))
        self._pattern = pattern.upper().encode("ascii") if pattern else None
        self.pattern_len = len(self._pattern) if self._pattern else 0
        self.text_len = 0
        self._score = -(2 ** 31)
        self._status = 0
        self._cigartuples = []
        self._locs[0] = self._locs[1] = self._locs[2] = self._locs[3] = 0
        cdef int rc = gpu.wfagpu_create(&self._ctx, device, err, sizeof(err))
        if rc != 0:
            raise RuntimeError(err.decode())

    def __dealloc__(self):
        if self._ctx != NULL:
            gpu.wfagpu_destroy(self._ctx)
            self._ctx = NULL

    def wavefront_align(self, text, pattern=None):
        """Align ``text`` to ``pattern`` (or the cached pattern); returns the score."""
        if pattern is not None:
            self._pattern = pattern.upper().encode("ascii")
        if self._pattern is None:
            raise ValueError("pattern is None")
        cdef bytes p = self._pattern
        cdef bytes t = text.upper().encode("ascii")
        self._text = text
        self.pattern_len = len(p)
        self.text_len = len(t)
        self._host._validate(len(p), len(t))
        cdef const char* pp = p
        cdef const char* tp = t
        cdef int32_t p_len = len(p), t_len = len(t), n_runs = 0
        cdef const uint32_t* runs = NULL
        cdef int rc
        with nogil:      # the one call that replaces wavefront_align(aligner, pattern, plen, text, tlen), align.pyx:439
            rc = gpu.wfagpu_align_pair(self._ctx, &self._cfg, pp, p_len, tp, t_len,
                                       &self._score, &self._status, self._locs, &runs, &n_runs)
        if rc == -5:
            raise NotImplementedError(gpu.wfagpu_last_error(self._ctx).decode())
        if rc == -1:
            raise ValueError(gpu.wfagpu_last_error(self._ctx).decode())
        if rc != 0:
            raise RuntimeError(gpu.wfagpu_last_error(self._ctx).decode())
        cdef int32_t i
        self._cigartuples = [(runs[i] & 15, runs[i] >> 4) for i in range(n_runs)]
        return self._score

    @property
    def score(self):
        return self._score

    @property
    def status(self):
        return self._status

    @property
    def cigartuples(self):
        return list(self._cigartuples)

    @property
    def cigarstring(self):
        return cigartuples_to_str(self._cigartuples)

    @property
    def locations(self):
        if self._cfg.scope == 0 or not self._cigartuples or self.text_len == 0 or self.pattern_len == 0:
            return [0, 0, 0, 0]
        return (self._locs[0], self._locs[1], self._locs[2], self._locs[3])

    def __call__(self, text, pattern=None, clip_cigar=False, min_aligned_bases_left=1,
                 min_aligned_bases_right=1, elide_mismatches=False, supress_sequences=False):
        """pywfa/align.pyx:835-879, incl. its scope gate around the post-processing."""
        if pattern is None:
            if not self._pattern:
                raise ValueError("pattern is None")
            p = self._pattern.decode("ascii")
            score = self.wavefront_align(text)
        else:
            p = pattern
            score = self.wavefront_align(text, pattern)
        locs = self.locations
        seqs = ("", "") if supress_sequences else (p, text)
        res = AlignmentResult(len(p), len(text), locs[0], locs[1], locs[2], locs[3], self.cigartuples, score,
                              seqs[0], seqs[1], self._status)
        if self._cfg.scope != 1:
            if clip_cigar:
                res = clip_cigartuples(res, min_aligned_bases_left, min_aligned_bases_right)
            if elide_mismatches:
                res.cigartuples = elide_mismatches_from_cigar(res.cigartuples)
        return res
