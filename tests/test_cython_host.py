"""The Cython host layer (pywfa_b200/cy): INTEGRATION.md's binding compiled for real.  Without a GPU
the extension must build, import and fail loudly at construction; on the GPU the reference's own
known-answer tests (pywfa/tests/test.py + README, tests/golden/reference_kat.json) run through it."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def cy():
    from pywfa_b200.build import build_cython
    build_cython()
    from pywfa_b200.cy import align_cy
    return align_cy


def test_cython_layer_builds_and_has_no_cpu_path(cy):
    from pywfa_b200 import _ffi
    assert cy.WavefrontAligner.__doc__
    with pytest.raises(NotImplementedError):
        cy.WavefrontAligner("ACGT", distance="hamming")           # same kwargs checks as pywfa
    with pytest.raises(ValueError):
        cy.WavefrontAligner("ACGT", scope="everything")
    if _ffi.lib().wfagpu_device_count() == 0:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            cy.WavefrontAligner("ACGT")


@pytest.mark.gpu
def test_reference_known_answers_through_the_cython_layer(cy):
    kat = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))
    kat += json.load(open(os.path.join(HERE, "golden", "metrics.json")))["kat"]      # linear / levenshtein / indel
    assert len(kat) >= 56
    for case in kat:
        a = cy.WavefrontAligner(**case["ctor"])
        res = a(case["text"], case["pattern"], **case["call"])
        e = case["expect"]
        assert (res.score, res.status, res.cigarstring) == (e["score"], e["status"], e["cigarstring"]), case["name"]
        assert [res.pattern_start, res.pattern_end, res.text_start, res.text_end] == e["locations"], case["name"]
        assert (a.score, a.status, a.cigarstring) == (e["aligner_score"], e["aligner_status"], e["aligner_cigarstring"])
    # the README pair, cached pattern, repeated calls on one aligner
    a = cy.WavefrontAligner("TCTTTACTCGCGCGTTGGAGAAATACAATAGT")
    for _ in range(3):
        assert a.wavefront_align("TCTATACTGCGCGTTTGGAGAAATAAAATAGT") == -24
        assert a.cigarstring == "3M1X4M1D7M1I9M1X6M" and a.status == 0
        assert a.locations == (0, 32, 0, 32)
