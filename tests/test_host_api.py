"""Host logic without a GPU: the C-ABI library loads and exports every symbol of
include/wfagpu.h, configuration checks mirror the reference's exit(1) conditions, the Python
mirror of pywfa's interface behaves like the reference (golden vectors), and the product path
fails loudly -- never silently on a CPU -- when no device is present."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import pywfa_b200
from pywfa_b200 import _ffi
from pywfa_b200.align import AlignmentResult
from pywfa_b200.build import build_library

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
POST = json.load(open(os.path.join(HERE, "golden", "postprocess.json")))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _ffi.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "wfagpu.h")).read()
    names = set(re.findall(r"\b(wfagpu_[a-z_]+)\s*\(", header))
    assert len(names) >= 13
    for name in sorted(names):
        assert hasattr(lib, name), f"{name} declared in include/wfagpu.h but not exported"


def test_config_default_and_layout(lib):
    cfg = _ffi.Config()
    lib.wfagpu_config_default(C.addressof(cfg))
    assert C.sizeof(cfg) == 80
    assert (cfg.distance, cfg.scope, cfg.span, cfg.heuristic) == (0, 1, 1, 0)
    assert (cfg.match, cfg.mismatch, cfg.gap_opening1, cfg.gap_extension1, cfg.gap_opening2, cfg.gap_extension2) == (0, 4, 6, 2, 24, 1)


@pytest.mark.parametrize("field,value,code", [
    ("match", 1, _ffi.EINVAL), ("mismatch", 0, _ffi.EINVAL), ("gap_opening1", -1, _ffi.EINVAL),
    ("gap_extension1", 0, _ffi.EINVAL), ("scope", 7, _ffi.EINVAL), ("span", 3, _ffi.EINVAL),
    ("heuristic", 9, _ffi.EINVAL), ("distance", 5, _ffi.EINVAL), ("distance", -1, _ffi.EINVAL),
])
def test_config_check_rejects_what_the_reference_exits_on(lib, field, value, code):
    cfg = _ffi.Config()
    lib.wfagpu_config_default(C.addressof(cfg))
    setattr(cfg, field, value)
    err = C.create_string_buffer(256)
    assert lib.wfagpu_config_check(C.addressof(cfg), -1, -1, err, 256) == code
    assert err.value


def test_config_check_endsfree_bounds(lib):
    cfg = _ffi.Config()
    lib.wfagpu_config_default(C.addressof(cfg))
    cfg.text_begin_free = 20
    err = C.create_string_buffer(256)
    assert lib.wfagpu_config_check(C.addressof(cfg), 100, 20, err, 256) == _ffi.OK
    assert lib.wfagpu_config_check(C.addressof(cfg), 100, 19, err, 256) == _ffi.EINVAL
    assert b"Ends-free parameters" in err.value


def test_strerror(lib):
    for code in range(0, -6, -1):
        assert lib.wfagpu_strerror(code)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must raise (this test is skipped on a GPU box)."""
    if lib.wfagpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_ffi.WfaGpuError) as ei:
        _ffi.Context(0)
    assert ei.value.code == _ffi.ENODEVICE
    a = pywfa_b200.WavefrontAligner("ACGT")
    with pytest.raises(_ffi.WfaGpuError):
        a("ACGT")
    with pytest.raises(_ffi.WfaGpuError):
        a.align_batch(["ACGT", "AC"])


def test_product_package_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pywfa_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "wfa_core.cuh", f"{f} mentions the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "pywfa_b200", "csrc", "wfa_core.cuh")).read().lower()


# ---- the Python mirror of pywfa's interface ---------------------------------------------------
def test_constructor_defaults_and_errors():
    a = pywfa_b200.WavefrontAligner()
    assert (a.distance, a.scope, a.span, a.heuristic, a.memory_mode) == ("affine", "full", "ends-free", None, "high")
    assert (a.match_score, a.mismatch_penalty, a.gap_opening_penalty, a.gap_extension_penalty) == (0, 4, 6, 2)
    assert (a.gap_opening2_penalty, a.gap_extension2_penalty) == (-1, -1)      # unused piece-2 penalties read -1
    assert a.max_steps == 2 ** 31 - 1
    b = pywfa_b200.WavefrontAligner(distance="affine2p", match=-1)
    assert (b.match_score, b.mismatch_penalty, b.gap_opening_penalty, b.gap_extension_penalty,
            b.gap_opening2_penalty, b.gap_extension2_penalty) == (-1, 10, 12, 5, 48, 3)   # Eizenga transform
    with pytest.raises(NotImplementedError):
        pywfa_b200.WavefrontAligner(distance="hamming")
    # the M-only metrics: getters return what wavefront_penalties_set_* leaves (penalties.c:38-93)
    lv = pywfa_b200.WavefrontAligner(distance="levenshtein")
    assert (lv.distance, lv.match_score, lv.mismatch_penalty, lv.gap_opening_penalty, lv.gap_extension_penalty,
            lv.gap_opening2_penalty, lv.gap_extension2_penalty) == ("levenshtein", 0, 1, 1, -1, -1, -1)
    ind = pywfa_b200.WavefrontAligner(distance="indel", mismatch=0)          # penalties are ignored
    assert (ind.distance, ind.mismatch_penalty, ind.gap_opening_penalty) == ("indel", -1, 1)
    lin = pywfa_b200.WavefrontAligner(distance="linear", mismatch=3, gap_opening=0, gap_extension=5)
    assert (lin.distance, lin.mismatch_penalty, lin.gap_opening_penalty, lin.gap_extension_penalty) == ("linear", 3, 5, -1)
    lin = pywfa_b200.WavefrontAligner(distance="linear", match=-1, mismatch=3, gap_extension=5)
    assert (lin.match_score, lin.mismatch_penalty, lin.gap_opening_penalty) == (-1, 8, 11)      # Eizenga transform
    lin.gap_opening_penalty = 2                                               # align.pyx:675: the indel penalty
    assert lin.gap_opening_penalty == 5
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(distance="linear", gap_extension=0)
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(distance="indel", heuristic="X-drop")     # the reference exit(1)s at the first alignment
    sw = pywfa_b200.WavefrontAligner(distance="levenshtein", mismatch=0)
    with pytest.raises(ValueError):
        sw.distance = "affine"                                                # mismatch=0 is not a gap-affine penalty
    assert sw.distance == "levenshtein"
    # BiWFA: exact, so score-only results equal the other memory modes'; its CIGAR tie-breaks / stops are not reproduced
    assert pywfa_b200.WavefrontAligner(memory_mode="biwfa", scope="score", span="end-to-end").memory_mode == "biwfa"
    with pytest.raises(NotImplementedError):
        pywfa_b200.WavefrontAligner(memory_mode="biwfa", span="end-to-end")
    with pytest.raises(NotImplementedError):
        pywfa_b200.WavefrontAligner(memory_mode="biwfa", scope="score", span="end-to-end", heuristic="adaptive")
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(memory_mode="biwfa", scope="score", text_end_free=3)    # the reference exit(1)s
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(scope="partial")
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(memory_mode="tiny")
    with pytest.raises(NotImplementedError):
        pywfa_b200.WavefrontAligner(span="local")
    with pytest.raises(NotImplementedError):
        pywfa_b200.WavefrontAligner(heuristic="z-drop")
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(mismatch=0)                 # the reference exit(1)s here
    with pytest.raises(TypeError):
        pywfa_b200.WavefrontAligner(wildcard=5)
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner(wildcard="NN")
    with pytest.raises(ValueError):
        pywfa_b200.WavefrontAligner()("ACGT")                   # pattern is None


def test_property_setters():
    a = pywfa_b200.WavefrontAligner("ACGT", heuristic="adaptive", min_wavefront_length=7, max_steps=5)
    assert (a.min_wavefront_length, a.max_distance_threshold, a.steps_between_cutoffs, a.max_steps) == (7, 50, 1, 5)
    a.scope = "score"; a.span = "end-to-end"; a.heuristic = "X-drop"; a.xdrop = 33; a.max_steps = 0
    assert (a.scope, a.span, a.heuristic, a.xdrop, a.max_steps) == ("score", "end-to-end", "X-drop", 33, 2 ** 31 - 1)
    a.gap_opening_penalty = 5; a.mismatch_penalty = 3
    assert (a.gap_opening_penalty, a.mismatch_penalty) == (5, 3)
    with pytest.raises(ValueError):
        a.gap_extension_penalty = 0
    assert a.gap_extension_penalty == 2
    with pytest.raises(ValueError):
        a.scope = "nope"
    a.text_end_free = 9
    assert a.text_end_free == 9


def _result(case):
    ct = [tuple(c) for c in case["cigartuples"]]
    return AlignmentResult(case["pattern_length"], case["text_length"], 0, case["pattern_length"],
                           case["text_start"], case["text_length"], ct, -1, "", "", 0)


def test_postprocessing_matches_reference_golden():
    n = {"clip": 0, "elide": 0, "str": 0}
    for case in POST:
        ct = [tuple(c) for c in case["cigartuples"]]
        if case["kind"] == "clip":
            res = pywfa_b200.clip_cigartuples(_result(case), case["left"], case["right"])
            assert [list(c) for c in res.cigartuples] == case["expect"]["cigartuples"]
            assert [res.pattern_start, res.pattern_end, res.text_start, res.text_end] == case["expect"]["locations"]
        elif case["kind"] == "elide":
            assert [list(c) for c in pywfa_b200.elide_mismatches_from_cigar(ct)] == case["expect"]
        else:
            assert pywfa_b200.cigartuples_to_str(ct) == case["expect"]
        n[case["kind"]] += 1
    assert min(n.values()) > 20


def test_alignment_result_helpers():
    r = AlignmentResult(8, 8, 0, 8, 0, 8, [(0, 3), (8, 1), (0, 4)], -4, "ACGTACGT", "ACGAACGT", 0)
    assert r.cigarstring == "3M1X4M"
    assert "ALIGNMENT" in r.pretty and "|||*||||" in r.pretty
    assert "score: -4" in repr(r)
    assert str(AlignmentResult(0, 0, 0, 0, 0, 0, [], -3, "", "", 0)) == "Score: -3"


def test_fastx_reader(tmp_path):
    """FASTA (wrapped lines, comments, blank lines), FASTQ and gzip through pywfa_b200.fastx.read_fastx
    (what pysam.FastxFile does for the reference's tests, pywfa/tests/test.py:198-232)."""
    import gzip

    from pywfa_b200.fastx import _batch_arrays, read_fastx
    fa = tmp_path / "a.fa"
    fa.write_text(">r1 first record\nACGT\nacgtn\n\n>r2\nTTTT\n>empty\n")
    recs = list(read_fastx(fa))
    assert [(r.name, r.sequence, r.comment) for r in recs] == [("r1", "ACGTacgtn", "first record"), ("r2", "TTTT", None), ("empty", "", None)]
    fq = tmp_path / "b.fq.gz"
    with gzip.open(fq, "wt") as fh:
        fh.write("@q1 c\nACGTN\n+\nIIII#\n@q2\nGG\n+q2\n!!\n")
    recs = list(read_fastx(fq))
    assert [(r.name, r.sequence, r.quality) for r in recs] == [("q1", "ACGTN", "IIII#"), ("q2", "GG", "!!")]
    bad = tmp_path / "bad.fq"
    bad.write_text("@q1\nACGT\n+\nII\n")
    with pytest.raises(ValueError):
        list(read_fastx(bad))
    seq, po, pl, to, tl = _batch_arrays([b"ACG", b"T"], [b"AC", b"TTTT"])
    assert seq.tobytes() == b"ACGACTTTTT\0" and po.tolist() == [0, 5] and to.tolist() == [3, 6] and pl.tolist() == [3, 1] and tl.tolist() == [2, 4]


def test_sam_reader(tmp_path):
    """Text SAM through pywfa_b200.fastx.read_sam / read_seqs: header lines skipped, QNAME / SEQ / QUAL taken,
    records without a stored sequence skipped, reverse-strand records restored on request, gzip, and the
    format sniffing that lets align_fastx take FASTA, FASTQ or SAM."""
    import gzip

    from pywfa_b200.fastx import read_sam, read_seqs
    sam = tmp_path / "r.sam"
    sam.write_text("@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:1000\n"
                   "r1\t0\tchr1\t10\t60\t8M\t*\t0\t0\tACGTTGCA\tIIIIHHHH\tNM:i:0\n"
                   "r2\t16\tchr1\t50\t60\t6M\t*\t0\t0\tAACCGT\t123456\n"
                   "r3\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n"
                   "r4\t0\tchr1\t70\t60\t4M\t*\t0\t0\tacgn\t*\n")
    recs = list(read_sam(sam))
    assert [(r.name, r.sequence, r.quality) for r in recs] == [("r1", "ACGTTGCA", "IIIIHHHH"), ("r2", "AACCGT", "123456"), ("r4", "acgn", None)]
    assert recs[0].comment == "0 chr1 10 8M"
    orig = list(read_sam(sam, original_orientation=True))
    assert (orig[1].sequence, orig[1].quality) == ("ACGGTT", "654321") and orig[0].sequence == "ACGTTGCA"
    gz = tmp_path / "r.sam.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write(sam.read_text())
    assert [r.sequence for r in read_seqs(gz)] == ["ACGTTGCA", "AACCGT", "acgn"]
    headless = tmp_path / "h.sam"
    headless.write_text("r1\t0\tchr1\t10\t60\t8M\t*\t0\t0\tACGTTGCA\tIIIIHHHH\n")
    assert [r.name for r in read_seqs(headless)] == ["r1"]
    fq = tmp_path / "r.fq"
    fq.write_text("@r1 c\nACGT\n+\nIIII\n")
    assert [r.sequence for r in read_seqs(fq)] == ["ACGT"]
    short = tmp_path / "bad.sam"
    short.write_text("@HD\tVN:1.6\nr1\t0\tchr1\n")
    with pytest.raises(ValueError):
        list(read_seqs(short))
