"""The oracle (oracle/wfa_oracle.c, a CPU restatement of the reference algorithm) is pinned
against (a) the golden vectors recorded from the unmodified reference (tests/golden/, made by
tests/golden/make_golden.py) and (b) the reference itself when oracle/_ref is built."""
import json
import os

import numpy as np
import pytest

from pywfa_b200.synth import generate_pairs, pairs_from_strings

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLD, "reference_kat.json")))
SYN = json.load(open(os.path.join(GOLD, "synthetic.json")))
MET = json.load(open(os.path.join(GOLD, "metrics.json")))          # distance = linear / levenshtein / indel
KAT = KAT + MET["kat"]
SYN = SYN + MET["synthetic"]

_CTOR_MAP = dict(gap_opening="gap_opening", gap_extension="gap_extension")


def config_from_ctor(oracle, ctor):
    """pywfa constructor kwargs -> wfagpu_config_t (pattern is not a config field)."""
    kw = {k: v for k, v in ctor.items() if k not in ("pattern", "memory_mode")}
    h = kw.get("heuristic")
    cfg = oracle.make_config(**kw)
    # only the chosen heuristic's parameters reach the aligner (wavefront_aligner.c:188-224)
    if h is None:
        cfg.min_wavefront_length = cfg.max_distance_threshold = cfg.steps_between_cutoffs = cfg.xdrop = 0
    elif h == "adaptive":
        cfg.xdrop = 0
    else:
        cfg.min_wavefront_length = cfg.max_distance_threshold = 0
    return cfg


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_oracle_matches_reference_kat(oracle, case):
    cfg = config_from_ctor(oracle, case["ctor"])
    batch = pairs_from_strings([(case["pattern"], case["text"])])
    r = oracle.align_batch(cfg, *batch, kind="port")
    e = case["expect"]
    assert int(r["score"][0]) == e["aligner_score"]
    assert int(r["status"][0]) == e["aligner_status"]
    assert oracle.runs_to_cigarstring(r["runs"]) == e["aligner_cigarstring"]
    if cfg.scope == 1 and not case["call"].get("clip_cigar"):
        assert r["locs"][0].tolist() == e["locations"]


@pytest.mark.parametrize("case", SYN, ids=[c["name"] for c in SYN])
@pytest.mark.parametrize("bt_mode", [0, 1], ids=["bt-reference", "bt-origin-codes"])
def test_oracle_matches_reference_synthetic(oracle, case, bt_mode):
    """bt_mode 1 = backtrace from forward-recorded origin codes, the scheme the CUDA kernels use."""
    batch = generate_pairs(case["n"], case["length"], case["div"], case["seed"], text_flank=case["flank"])
    assert int(batch[0].astype(np.uint64).sum()) == case["input_checksum"], "synthetic generator drifted"
    cfg = oracle.make_config(**case["config"])
    r = oracle.align_batch(cfg, *batch, kind="port", bt_mode=bt_mode)
    assert r["score"].tolist() == case["score"]
    assert r["status"].tolist() == case["status"]
    cig = [oracle.runs_to_cigarstring(r["runs"][r["cig_off"][j]:r["cig_off"][j + 1]]) for j in range(case["n"])]
    assert cig == case["cigars"]
    assert r["locs"].tolist() == case["locations"]
    if case["config"].get("scope", "full") == "full":
        assert r["cells"].tolist() == case["cells"]


def prune_pairs(seed):
    """the pairs of metrics.json's "prune" section (tests/golden/make_golden.py:metric_cases)"""
    rng = np.random.default_rng(seed)
    rs = lambda m: "".join("ACGT"[i] for i in rng.integers(0, 4, m))      # noqa: E731
    pairs = []
    for pl, tl in ((300, 3000), (3000, 300), (1500, 2500), (100, 2500)):
        p = rs(pl)
        pairs += [(p, rs(tl)), (p, (p * (tl // pl + 1))[:tl])]
    return pairs_from_strings(pairs)


@pytest.mark.parametrize("case", MET["prune"], ids=[c["config"].get("scope", "full") for c in MET["prune"]])
def test_oracle_matches_reference_edit_exact_prune(oracle, case):
    """levenshtein end-to-end drops provably useless ends of wavefronts of >= 1000 diagonals"""
    import hashlib
    r = oracle.align_batch(oracle.make_config(**case["config"]), *prune_pairs(case["seed"]), kind="port")
    assert r["score"].tolist() == case["score"] and r["status"].tolist() == case["status"]
    assert hashlib.sha256(r["runs"].tobytes()).hexdigest() == case["cigar_sha256"] and len(r["runs"]) == case["n_runs"]
    if case["config"].get("scope", "full") == "full":
        assert r["cells"].tolist() == case["cells"]


LIVE = [
    ("affine-e2e", dict(span="end-to-end"), 600, 150, 0.08, 0),
    ("affine-score", dict(span="end-to-end", scope="score"), 300, 250, 0.10, 0),
    ("2p-endsfree", dict(distance="affine2p", pattern_end_free=15, text_begin_free=9), 200, 300, 0.12, 6),
    ("adaptive", dict(heuristic="adaptive", min_wavefront_length=5, max_distance_threshold=10), 200, 300, 0.2, 0),
    ("xdrop", dict(heuristic="X-drop", xdrop=60, steps_between_cutoffs=2), 200, 300, 0.1, 0),
    ("match-2", dict(span="end-to-end", match=-2, distance="affine2p"), 200, 150, 0.1, 0),
    ("linear", dict(distance="linear", pattern_end_free=12, text_end_free=7, mismatch=3, gap_extension=2), 200, 200, 0.12, 0),
    ("linear-match-1-xdrop", dict(distance="linear", span="end-to-end", match=-1, heuristic="X-drop", xdrop=40), 200, 200, 0.12, 0),
    ("levenshtein-adaptive", dict(distance="levenshtein", heuristic="adaptive", min_wavefront_length=5, max_distance_threshold=8), 200, 300, 0.2, 0),
    ("indel-max-steps", dict(distance="indel", span="end-to-end", max_steps=30), 200, 200, 0.1, 0),
]


def test_oracle_matches_reference_live_wildcard(oracle):
    """non-ACGT bytes and the wildcard (pywfa/align.pyx:297-304,438-442): the port against the
    reference's wavefront_align_lambda path"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from test_emu import _pairs_with_n, BYTE_KW
    from pywfa_b200.synth import pairs_from_strings
    batch = pairs_from_strings(_pairs_with_n(11, 400, 20, 260))
    for kw in BYTE_KW:
        cfg = oracle.make_config(**kw)
        ref = oracle.align_batch(cfg, *batch, kind="reference")
        port = oracle.align_batch(cfg, *batch, kind="port")
        for k in ("score", "status", "cig_off", "runs"):
            assert np.array_equal(ref[k], port[k]), (kw, k)


@pytest.mark.parametrize("name,kw,n,length,div,flank", LIVE, ids=[c[0] for c in LIVE])
def test_oracle_matches_reference_live(oracle, name, kw, n, length, div, flank):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    batch = generate_pairs(n, length, div, seed=31, text_flank=flank)
    cfg = oracle.make_config(**kw)
    ref = oracle.align_batch(cfg, *batch, kind="reference")
    for bt in (0, 1):
        port = oracle.align_batch(cfg, *batch, kind="port", bt_mode=bt)
        for k in ("score", "status", "cig_off", "runs"):
            assert np.array_equal(ref[k], port[k]), (name, bt, k)
        if kw.get("scope", "full") == "full":
            assert np.array_equal(ref["cells"], port["cells"])


def test_oracle_matches_long_read_golden(oracle):
    """The port against the reference's low-memory mode on 20 kbp affine2p pairs
    (tests/golden/long_reads_20kbp.json, made by make_golden_long.py); the 100 kbp vectors are
    checked on the GPU only (the port keeps the full history in host memory)."""
    import hashlib
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long_reads_20kbp.json")
    g = json.load(open(path))
    gen = g["generator"]
    batch = generate_pairs(gen["n"], gen["length"], gen["div"], seed=gen["seed"])
    r = oracle.align_batch(oracle.make_config(**g["config"]), *batch, kind="port")
    for i, want in enumerate(g["pairs"]):
        runs = np.ascontiguousarray(r["runs"][r["cig_off"][i]:r["cig_off"][i + 1]], np.uint32)
        assert (int(r["score"][i]), int(r["status"][i]), len(runs)) == (want["score"], want["status"], want["nruns"])
        assert hashlib.sha256(runs.tobytes()).hexdigest() == want["runs_sha256"]
