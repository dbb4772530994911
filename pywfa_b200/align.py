"""Host-side mirror of pywfa's public interface on top of the B200 C-ABI library.

Same names, keyword arguments, defaults, return values and error behaviour as
``pywfa/align.pyx`` (reference, pywfa 0.5.1) for the accelerated path -- ``WavefrontAligner``
(ctor ``align.pyx:309-419``, ``wavefront_align`` ``:421-443``, ``__call__`` ``:835-879``, the
properties ``:461-833``), ``AlignmentResult`` (``:17-180``), ``clip_cigartuples`` (``:183-250``),
``elide_mismatches_from_cigar`` (``:253-277``), ``cigartuples_to_str`` (``:280-295``) -- plus the
batched entry points ``WavefrontAligner.align_batch`` / ``align_arrays`` that the reference does
not have.  Every alignment, including a single ``aligner(text, pattern)`` call, runs on the GPU
through ``libwfagpu.so``; there is no CPU fallback.

Intentional deviations from the reference:
  * configurations the reference ``exit(1)``s on raise ``ValueError`` before any launch;
  * ``memory_mode="biwfa"`` is accepted for ``scope="score"`` without cut-offs and step limits
    (BiWFA is exact: the reference returns the same score and status as its other memory modes
    there, pinned by ``tests/golden/metrics.json``) and runs on the same kernels; with
    ``scope="full"`` BiWFA picks other co-optimal CIGARs (7 % of 200 bp pairs at 10 % divergence),
    with cut-offs / ``max_steps`` it stops elsewhere -- those raise ``NotImplementedError``;
    non-ACGT bases and the wildcard are aligned in the library's byte mode;
  * ``distance="linear" | "levenshtein" | "indel"`` are aligned by the scalar tiers only (no
    register / packed-halfword kernels for them).  Changing ``distance`` after construction keeps
    the constructor's penalties (the reference's setter re-reads per-metric copies that it only
    initialised for the constructor's metric, ``align.pyx:610-620``);
  * ``span="end-to-end"`` ignores the ``*_begin_free`` / ``*_end_free`` values.  pywfa copies them
    into ``alignment_form`` regardless of the span, and WFA2-lib then still seeds wavefront 0 with
    ``lo = -pattern_begin_free, hi = text_begin_free`` (``wavefront_aligner.c:260-261``) while
    initialising only offset 0 -- the other seeds are uninitialised memory, so the reference's
    result for that combination is undefined; here the seed is the single cell ``k = 0``;
  * ``match < 0`` together with ends-free *begin* gaps (score-dependent seeding,
    ``wavefront_compute.c:124-254``) raises ``NotImplementedError``: the reference itself does not
    survive that combination on ordinary reads (r02, the unmodified library on 200 bp pairs at
    10 % divergence: ``match=-1, text_begin_free=20, text_end_free=20, distance="affine2p"`` ends
    in ``[WFA::Backtrace] I?/D?-Beginning backtrace error`` and ``exit(1)``), so there is no
    behaviour to be bit-exact with.

Thread safety: all aligners of one device share one library context; calls from several threads
are serialised inside the library (pywfa's aligners are independent objects, one per thread works
there and here).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _ffi

__all__ = ["WavefrontAligner", "AlignmentResult", "BatchResult", "clip_cigartuples",
           "cigartuples_to_str", "elide_mismatches_from_cigar"]

INT_MAX = 2 ** 31 - 1
_CIGAR_LETTERS = "MIDNSHP=XB"          # SAM op codes, pywfa/align.pyx:11-14 / :290

_DISTANCES = {"affine": 0, "affine2p": 1, "linear": 2, "levenshtein": 3, "indel": 4}     # wfagpu.h WFAGPU_DISTANCE_*
_DISTANCE_NAMES = {v: k for k, v in _DISTANCES.items()}
_SCOPES = {"score": 0, "full": 1}
_SPANS = {"end-to-end": 0, "ends-free": 1}
_HEURISTICS = {None: 0, "adaptive": 1, "X-drop": 2}
_MEMORY_MODES = ("high", "medium", "low", "biwfa")


@dataclass
class AlignmentResult:
    """Result of one alignment (field-for-field the reference's dataclass, align.pyx:17-46)."""
    pattern_length: int
    text_length: int
    pattern_start: int
    pattern_end: int
    text_start: int
    text_end: int
    cigartuples: object
    score: int
    pattern: str
    text: str
    status: int

    def __init__(self, pl, tl, ps, pe, ts, te, ct, s, p, t, status):
        self.pattern_length, self.text_length = pl, tl
        self.pattern_start, self.pattern_end = ps, pe
        self.text_start, self.text_end = ts, te
        self.cigartuples = ct
        self.score = s
        self.pattern, self.text = p, t
        self.status = status

    def __repr__(self):
        keys = ("score", "pattern_start", "pattern_end", "text_start", "text_end",
                "cigartuples", "pattern", "text")
        return "".join(f"    {k}: {getattr(self, k)}\n" for k in keys)

    def __str__(self):
        head = "Score: %d" % self.score
        if not (self.pattern and self.cigartuples):
            return head
        t, p = self.aligned_text, self.aligned_pattern
        n = len(t)
        if n > 30:
            t, p = t[:30] + "...", p[:30] + "..."
        return "\n".join([p, t, self.cigarstring[:30], head, "Length: %d" % len(t)])

    # The reference unpacks each (op, length) tuple as ``length, mid`` and compares ``mid`` with
    # the letter "D"/"I" (align.pyx:168-180); the comparison is therefore never true and the
    # op code is used as a slice length.  Kept as is: "post-processing unchanged".
    def _get_aligned_sequence(self, sequence, tuple_cigar, begin, end, gap_type):
        window = sequence[begin:end]
        parts, pos = [], 0
        for length, mid in tuple_cigar:
            if mid == gap_type:
                parts.append("-" * length)
            else:
                parts.append(window[pos:pos + length])
                pos += length
        parts.append(window[pos:end - begin])
        return "".join(parts)

    @property
    def aligned_pattern(self):
        if self.pattern:
            return self._get_aligned_sequence(self.pattern, self.cigartuples, self.pattern_start,
                                              self.pattern_end, "D")

    @property
    def aligned_text(self):
        if self.text:
            return self._get_aligned_sequence(self.text, self.cigartuples, self.text_start,
                                              self.text_end, "I")

    @property
    def cigarstring(self):
        return cigartuples_to_str(self.cigartuples)

    @property
    def pretty(self):
        """Three-line rendering of the alignment (align.pyx:121-165)."""
        out = f"{self.cigarstring}      ALIGNMENT\n"
        compact = [c for c in self.cigartuples if c[0] != 0 and c[0] != [8]]
        out += f"{cigartuples_to_str(compact)}      ALIGNMENT.COMPACT\n"
        rows = {"p": "      PATTERN    ", "g": "                 ", "t": "      TEXT       "}
        pi = ti = 0
        for op, ln in self.cigartuples:
            if op in (1, 4, 5):
                rows["t"] += self.text[ti:ti + ln]; ti += ln
                rows["p"] += "-" * ln; rows["g"] += " " * ln
            elif op in (0, 7, 8):
                rows["t"] += self.text[ti:ti + ln]; ti += ln
                rows["p"] += self.pattern[pi:pi + ln]; pi += ln
                rows["g"] += ("*" if op == 8 else "|") * ln
            elif op == 2:
                rows["t"] += "-" * ln
                rows["p"] += self.pattern[pi:pi + ln]; pi += ln
                rows["g"] += " " * ln
            else:
                raise ValueError(f"Cigar operation not available for pretty print - {op}")
        return out + rows["p"] + "\n" + rows["g"] + "\n" + rows["t"] + "\n"


def _flank(ct, order, min_match):
    """Walk CIGAR runs from one end until an M run of at least ``min_match``; return the index
    where the walk stopped and the (text, pattern) bases consumed before it."""
    text = pattern = 0
    stop = order[-1] if len(order) else 0
    for idx in order:
        stop = idx
        op, ln = ct[idx][0], ct[idx][1]
        if op == 0:
            if ln >= min_match:
                break
            text += ln; pattern += ln
        elif op == 2:
            pattern += ln
        elif op == 8:
            text += ln; pattern += ln
        elif op == 1:
            text += ln
    return stop, text, pattern


def clip_cigartuples(align_result, min_aligned_bases_left=5, min_aligned_bases_right=5):
    """Soft-clip flanks whose aligned blocks are shorter than the thresholds
    (behaviour of align.pyx:183-250)."""
    ct = align_result.cigartuples
    if not ct:
        return align_result
    i, t_lead, p_lead = _flank(ct, range(len(ct)), int(min_aligned_bases_left))
    j, t_tail, p_tail = _flank(ct, range(len(ct) - 1, -1, -1), int(min_aligned_bases_right))
    text_end = align_result.text_length - t_tail
    pattern_end = align_result.pattern_length - p_tail
    clipped = []
    if align_result.text_start + t_lead > 0:
        clipped.append((4, t_lead))
    clipped += ct[i:j + 1]
    if align_result.text_length - text_end > 0:
        clipped.append((4, align_result.text_length - text_end))
    align_result.cigartuples = clipped
    align_result.text_start, align_result.text_end = t_lead, text_end
    align_result.pattern_start, align_result.pattern_end = p_lead, pattern_end
    return align_result


def elide_mismatches_from_cigar(cigartuples):
    """Merge X runs into the neighbouring M blocks (behaviour of align.pyx:253-277)."""
    if not cigartuples:
        return []
    out, block = [], 0
    for op, ln in cigartuples:
        if op in (0, 8):
            block += int(ln)
            continue
        if block:
            out.append((0, block))
            block = 0
        out.append((op, ln))
    if block:
        out.append((0, block))
    return out


def cigartuples_to_str(cigartuples):
    """``[(op, length), ...]`` -> CIGAR string (align.pyx:280-295)."""
    if not cigartuples:
        return ""
    return "".join(f"{int(ln)}{_CIGAR_LETTERS[op]}" for op, ln in cigartuples)


def _runs_to_tuples(runs):
    return [(int(w) & 15, int(w) >> 4) for w in runs]


class BatchResult:
    """Arrays returned by the batched entry point.

    ``score``/``status``: int32[n]; ``locations``: int32[n,4] = pattern_start, pattern_end,
    text_start, text_end (as the reference's ``locations`` property); CIGAR of pair ``i`` =
    ``cig_runs[cig_off[i]:cig_off[i+1]]`` with each word ``length << 4 | op`` (SAM op codes).
    """

    def __init__(self, d, p_len, t_len):
        self.score = d["score"]
        self.status = d["status"]
        self.locations = d["locs"]
        self.cig_off = d["cig_off"]
        self.cig_runs = d["runs"]
        self.pattern_length = p_len
        self.text_length = t_len

    def __len__(self):
        return len(self.score)

    def cigartuples(self, i):
        return _runs_to_tuples(self.cig_runs[self.cig_off[i]:self.cig_off[i + 1]])

    def cigarstring(self, i):
        return cigartuples_to_str(self.cigartuples(i))

    def result(self, i, pattern="", text=""):
        ps, pe, ts, te = (int(v) for v in self.locations[i])
        return AlignmentResult(int(self.pattern_length[i]), int(self.text_length[i]), ps, pe, ts, te,
                               self.cigartuples(i), int(self.score[i]), pattern, text,
                               int(self.status[i]))


_contexts = {}


def _run(ctx, *args, **kw):
    """Map library error codes onto the exception types of the pywfa surface."""
    try:
        return ctx.align_batch(*args, **kw)
    except _ffi.WfaGpuError as e:
        if e.code == _ffi.EUNSUPPORTED:
            raise NotImplementedError(str(e)) from None
        if e.code == _ffi.EINVAL:
            raise ValueError(str(e)) from None
        raise


def _context(device):
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = _ffi.Context(device)
    return ctx


class WavefrontAligner:
    """Drop-in for ``pywfa.WavefrontAligner`` on the gap-affine / gap-affine-2p path, executed
    by hand-written sm_100a kernels.  If a pattern is supplied it is cached for re-use."""

    def __init__(self, pattern=None, distance="affine", memory_mode="high", match=0, mismatch=4,
                 gap_opening=6, gap_extension=2, gap_opening2=24, gap_extension2=1, scope="full",
                 span="ends-free", pattern_begin_free=0, pattern_end_free=0, text_begin_free=0,
                 text_end_free=0, heuristic=None, min_wavefront_length=10,
                 max_distance_threshold=50, steps_between_cutoffs=1, xdrop=20, wildcard=None,
                 max_steps=0, device=0):
        self.pattern_len = 0
        self.text_len = 0
        self._pattern = None
        self._text = None
        if pattern:
            self._pattern = pattern.upper()
            self.pattern_len = len(self._pattern.encode("ascii"))
        self._cfg = _ffi.Config()
        self._device = device
        self.wildcard = wildcard
        if distance not in _DISTANCES:
            raise NotImplementedError(f'{distance} distance not implemented')
        self._cfg.distance = _DISTANCES[distance]
        # the caller's penalties; gap-linear's single indel penalty is the constructor's
        # gap_extension (align.pyx:351-355) and follows both gap setters afterwards (:675, :684)
        self._pen = dict(match=int(match), mismatch=int(mismatch), gap_opening1=int(gap_opening),
                         gap_extension1=int(gap_extension), gap_opening2=int(gap_opening2),
                         gap_extension2=int(gap_extension2), indel=int(gap_extension))
        self._sync_penalties()
        if scope not in _SCOPES:
            raise ValueError(f'{scope} scope not understood')
        self._cfg.scope = _SCOPES[scope]
        if memory_mode not in _MEMORY_MODES:
            raise ValueError("memory_mode must be one of 'high', 'medium', 'low', 'biwfa'")
        self._memory_mode = memory_mode        # "biwfa": see _validate
        self._cfg.pattern_begin_free, self._cfg.pattern_end_free = int(pattern_begin_free), int(pattern_end_free)
        self._cfg.text_begin_free, self._cfg.text_end_free = int(text_begin_free), int(text_end_free)
        if span not in _SPANS:
            raise NotImplementedError(f'{span} span not implemented')
        self._cfg.span = _SPANS[span]
        if heuristic not in _HEURISTICS:
            raise NotImplementedError(f'{heuristic} heuristic not implemented')
        self._cfg.heuristic = _HEURISTICS[heuristic]
        # Only the parameters of the chosen heuristic reach the aligner
        # (wavefront_aligner_init_heuristic, W/wavefront/wavefront_aligner.c:188-224); the others
        # stay zero, which is what the property getters of the reference then return.
        self._cfg.min_wavefront_length = self._cfg.max_distance_threshold = 0
        self._cfg.steps_between_cutoffs = self._cfg.xdrop = 0
        if heuristic == "adaptive":
            self._cfg.min_wavefront_length = int(min_wavefront_length)
            self._cfg.max_distance_threshold = int(max_distance_threshold)
            self._cfg.steps_between_cutoffs = int(steps_between_cutoffs)
        elif heuristic == "X-drop":
            self._cfg.xdrop = int(xdrop)
            self._cfg.steps_between_cutoffs = int(steps_between_cutoffs)
        self._cfg.max_steps = int(max_steps) if max_steps > 0 else 0
        self._validate()
        # last single-pair result
        self._score = -(2 ** 31)
        self._status = 0
        self._cigartuples = []
        self._locations = [0, 0, 0, 0]

    # ---- configuration -------------------------------------------------------------------
    def _validate(self, plen=-1, tlen=-1):
        import ctypes as C
        if getattr(self, "_memory_mode", "high") == "biwfa":
            c = self._cfg
            if c.span == 1 and (c.pattern_begin_free > 0 or c.pattern_end_free > 0 or c.text_begin_free > 0
                                or c.text_end_free > 0):
                # wavefront_align_presets__checks, W/wavefront/wavefront_align.c:66-76 (exit(1) there)
                raise ValueError("[WFA] BiWFA ends-free has not been tested properly yet")
            if c.scope != 0 or c.heuristic != 0 or c.max_steps > 0:
                raise NotImplementedError("memory_mode='biwfa' is on the accelerated path for scope='score' without "
                                          "heuristic and max_steps only (elsewhere BiWFA breaks ties / stops differently)")
        err = C.create_string_buffer(512)
        rc = _ffi.lib().wfagpu_config_check(C.addressof(self._cfg), plen, tlen, err, len(err))
        if rc == _ffi.EUNSUPPORTED:
            raise NotImplementedError(err.value.decode())
        if rc != _ffi.OK:
            raise ValueError(err.value.decode())

    def _sync_penalties(self):
        """caller's penalties -> the configuration the library reads (gap-linear: the indel
        penalty travels in gap_extension1, include/wfagpu.h)"""
        c, u = self._cfg, self._pen
        c.match, c.mismatch = u["match"], u["mismatch"]
        c.gap_opening1 = u["gap_opening1"]
        c.gap_extension1 = u["indel"] if c.distance == 2 else u["gap_extension1"]
        c.gap_opening2, c.gap_extension2 = u["gap_opening2"], u["gap_extension2"]

    def _normalised(self):
        """(match, mismatch, o1, e1, o2, e2) as WFA2-lib stores them after
        wavefront_penalties_set_* (Eizenga's transform when match < 0; -1 for the unused
        penalties of the metric, W/wavefront/wavefront_penalties.c:38-173), which is what the
        reference's getters return."""
        c = self._cfg
        if c.distance == 4:
            return [0, -1, 1, -1, -1, -1]
        if c.distance == 3:
            return [0, 1, 1, -1, -1, -1]
        if c.distance == 2:
            if c.match < 0:
                return [c.match, 2 * c.mismatch - 2 * c.match, 2 * c.gap_extension1 - c.match, -1, -1, -1]
            return [0, c.mismatch, c.gap_extension1, -1, -1, -1]
        if c.match < 0:
            vals = [c.match, 2 * c.mismatch - 2 * c.match, 2 * c.gap_opening1,
                    2 * c.gap_extension1 - c.match, 2 * c.gap_opening2, 2 * c.gap_extension2 - c.match]
        else:
            vals = [0, c.mismatch, c.gap_opening1, c.gap_extension1, c.gap_opening2, c.gap_extension2]
        if c.distance == 0:
            vals[4] = vals[5] = -1
        return vals

    @property
    def status(self):
        return self._status

    @property
    def score(self):
        return self._score

    def _int_prop(field):                                    # noqa: N805 (property factory)
        def get(self):
            return getattr(self._cfg, field)

        def set_(self, value):
            setattr(self._cfg, field, int(value))
        return property(get, set_)

    pattern_begin_free = _int_prop("pattern_begin_free")
    pattern_end_free = _int_prop("pattern_end_free")
    text_begin_free = _int_prop("text_begin_free")
    text_end_free = _int_prop("text_end_free")
    min_wavefront_length = _int_prop("min_wavefront_length")
    max_distance_threshold = _int_prop("max_distance_threshold")
    steps_between_cutoffs = _int_prop("steps_between_cutoffs")
    xdrop = _int_prop("xdrop")
    del _int_prop

    @property
    def scope(self):
        return "full" if self._cfg.scope == 1 else "score"

    @scope.setter
    def scope(self, scope):
        if scope not in _SCOPES:
            raise ValueError(f'{scope} scope not understood')
        self._cfg.scope = _SCOPES[scope]

    @property
    def span(self):
        return "ends-free" if self._cfg.span == 1 else "end-to-end"

    @span.setter
    def span(self, span):
        if span not in _SPANS:
            raise NotImplementedError(f'{span} span not implemented')
        self._cfg.span = _SPANS[span]

    @property
    def memory_mode(self):
        return self._memory_mode

    @memory_mode.setter
    def memory_mode(self, memory_mode):
        # the reference's setter spells medium "med" (align.pyx:545)
        if memory_mode not in ("high", "med", "medium", "low"):
            raise NotImplementedError(f'{memory_mode} memory_mode not implemented')
        self._memory_mode = "medium" if memory_mode == "med" else memory_mode

    @property
    def heuristic(self):
        return {0: None, 1: "adaptive", 2: "X-drop"}[self._cfg.heuristic]

    @heuristic.setter
    def heuristic(self, heuristic):
        if heuristic not in _HEURISTICS:
            raise NotImplementedError(f'{heuristic} heuristic not implemented')
        self._cfg.heuristic = _HEURISTICS[heuristic]

    @property
    def distance(self):
        return _DISTANCE_NAMES[self._cfg.distance]

    @distance.setter
    def distance(self, distance):
        if distance not in _DISTANCES:
            raise NotImplementedError(f'{distance} distance not implemented')
        old = self._cfg.distance
        self._cfg.distance = _DISTANCES[distance]
        self._sync_penalties()
        try:
            self._validate()
        except Exception:
            self._cfg.distance = old
            self._sync_penalties()
            raise

    def _penalty_prop(field, index):                          # noqa: N805
        def get(self):
            return self._normalised()[index]

        def set_(self, value):
            old = dict(self._pen)
            self._pen[field] = int(value)
            if field in ("gap_opening1", "gap_extension1"):
                self._pen["indel"] = int(value)
            self._sync_penalties()
            try:
                self._validate()
            except Exception:
                self._pen = old
                self._sync_penalties()
                raise
        return property(get, set_)

    match_score = _penalty_prop("match", 0)
    mismatch_penalty = _penalty_prop("mismatch", 1)
    gap_opening_penalty = _penalty_prop("gap_opening1", 2)
    gap_extension_penalty = _penalty_prop("gap_extension1", 3)
    gap_opening2_penalty = _penalty_prop("gap_opening2", 4)
    gap_extension2_penalty = _penalty_prop("gap_extension2", 5)
    del _penalty_prop

    @property
    def wildcard(self):
        return self._wildcard

    @wildcard.setter
    def wildcard(self, wildcard):
        # pywfa/align.pyx:709-719; the byte goes to the C ABI, whose extension then treats it as
        # "matches every base" (wildcard_match_fun, pywfa/align.pyx:302-304)
        if wildcard is None:
            self._wildcard = None
            self._cfg.wildcard = 0
            return
        if not isinstance(wildcard, str):
            raise TypeError(f"expected wildcard to be a string, but it is {type(wildcard)}")
        if len(wildcard) > 1:
            raise ValueError(f"wildcard must have length 1, but has length {len(wildcard)}")
        self._wildcard = wildcard
        self._cfg.wildcard = wildcard.upper().encode("ascii")[0] if wildcard else 0

    @property
    def max_steps(self):
        return self._cfg.max_steps if self._cfg.max_steps > 0 else INT_MAX

    @max_steps.setter
    def max_steps(self, steps):
        steps = int(steps)
        self._cfg.max_steps = 0 if (steps <= 0 or steps >= INT_MAX) else steps

    # ---- single pair (a batch of one through the same kernels) ---------------------------
    def wavefront_align(self, text, pattern=None):
        """Align ``text`` to ``pattern`` (or the cached pattern); returns the score."""
        if pattern is not None:
            self._pattern = pattern.upper()
        if self._pattern is None:
            raise ValueError("pattern is None")
        pb = self._pattern.encode("ascii")
        self.pattern_len = len(pb)
        tb = text.upper().encode("ascii")
        self._text = text
        self.text_len = len(tb)
        self._validate(len(pb), len(tb))
        try:
            self._score, self._status, self._locations, runs = _context(self._device).align_pair(self._cfg, pb, tb)
        except _ffi.WfaGpuError as e:
            if e.code == _ffi.EUNSUPPORTED:
                raise NotImplementedError(str(e)) from None
            if e.code == _ffi.EINVAL:
                raise ValueError(str(e)) from None
            raise
        self._cigartuples = [(w & 15, w >> 4) for w in runs]
        return self._score

    @property
    def cigartuples(self):
        return list(self._cigartuples)

    @property
    def cigarstring(self):
        return cigartuples_to_str(self._cigartuples)

    @property
    def locations(self):
        if self.scope == "score" or not self._cigartuples or self.text_len == 0 or self.pattern_len == 0:
            return [0, 0, 0, 0]
        return tuple(self._locations)

    def cigar_print_pretty(self, file_name=None):
        """Pretty-print the last alignment (the reference calls WFA2-lib's
        cigar_print_pretty, align.pyx:445-459; rendered here from the CIGAR runs)."""
        res = AlignmentResult(self.pattern_len, self.text_len, 0, 0, 0, 0, self.cigartuples,
                              self._score, self._pattern, (self._text or "").upper(), self._status)
        out = res.pretty
        if file_name:
            with open(file_name, "w") as fh:
                fh.write(out)
        else:
            print(out, end="")

    def __call__(self, text, pattern=None, clip_cigar=False, min_aligned_bases_left=1,
                 min_aligned_bases_right=1, elide_mismatches=False, supress_sequences=False):
        """Align ``text`` to ``pattern``; returns an ``AlignmentResult`` (align.pyx:835-879)."""
        if pattern is None:
            p = self._pattern
            if not p:
                raise ValueError("pattern is None")
            score = self.wavefront_align(text)
        else:
            p = pattern
            score = self.wavefront_align(text, pattern)
        lp = len(p)
        ct, locs, status = self.cigartuples, self.locations, self.status
        seqs = ("", "") if supress_sequences else (p, text)
        res = AlignmentResult(lp, len(text), locs[0], locs[1], locs[2], locs[3], ct, score,
                              seqs[0], seqs[1], status)
        # As shipped, the reference only post-processes when scope is NOT "full"
        # (align.pyx:874), where the CIGAR is empty and both steps are no-ops.  Kept verbatim
        # for bit-exact results; call clip_cigartuples / elide_mismatches_from_cigar directly.
        if not self.scope == "full":
            if clip_cigar:
                res = clip_cigartuples(res, min_aligned_bases_left, min_aligned_bases_right)
            if elide_mismatches:
                res.cigartuples = elide_mismatches_from_cigar(res.cigartuples)
        return res

    # ---- batched entry points (new) -------------------------------------------------------
    def align_arrays(self, seq, p_off, p_len, t_off, t_len) -> BatchResult:
        """Align ``n`` pairs given as one uint8 ASCII buffer plus offset/length arrays
        (pattern ``i`` = ``seq[p_off[i]:p_off[i]+p_len[i]]``, text likewise)."""
        p_len = np.ascontiguousarray(p_len, np.int32)
        t_len = np.ascontiguousarray(t_len, np.int32)
        if self._cfg.span == 1 and len(p_len):
            self._validate(int(p_len.min()), int(t_len.min()))
        d = _run(_context(self._device), self._cfg, seq, p_off, p_len, t_off, t_len)
        return BatchResult(d, p_len, t_len)

    def align_batch(self, texts, patterns=None) -> BatchResult:
        """Align ``texts[i]`` to ``patterns[i]`` (or to the cached pattern) for all ``i``."""
        texts = list(texts)
        if patterns is None:
            if not self._pattern:
                raise ValueError("pattern is None")
            patterns = [self._pattern] * len(texts)
        else:
            patterns = list(patterns)
            if len(patterns) != len(texts):
                raise ValueError("texts and patterns differ in length")
        pb = [p.upper().encode("ascii") for p in patterns]
        tb = [t.upper().encode("ascii") for t in texts]
        p_len = np.fromiter((len(b) for b in pb), np.int32, len(pb))
        t_len = np.fromiter((len(b) for b in tb), np.int32, len(tb))
        rec = p_len.astype(np.int64) + t_len
        p_off = np.zeros(len(pb), np.int64)
        np.cumsum(rec[:-1], out=p_off[1:])
        t_off = p_off + p_len
        seq = np.frombuffer(b"".join(x for pair in zip(pb, tb) for x in pair) + b"\0", np.uint8)
        return self.align_arrays(seq, p_off, p_len, t_off, t_len)
