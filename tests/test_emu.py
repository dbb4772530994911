"""CPU check of the DEVICE source: pywfa_b200/csrc/wfa_core.cuh compiled for the host with a
one-thread group (tests/emu/, test infrastructure) must reproduce the oracle bit-exactly,
report capacity overflows instead of wrong answers, and share the host packer with the library."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from pywfa_b200.synth import generate_pairs, pairs_from_strings

HERE = os.path.dirname(os.path.abspath(__file__))
SYN = json.load(open(os.path.join(HERE, "golden", "synthetic.json")))
SYN = SYN + json.load(open(os.path.join(HERE, "golden", "metrics.json")))["synthetic"]

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", os.path.join(HERE, "emu")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HERE, "emu", "libwfaemu.so"))
    lib.emu_align_batch.argtypes = [C.c_void_p, _u8p, _i64p, _i32p, _i64p, _i32p, C.c_int64, C.c_int,
                                    C.c_longlong, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i64p, _u32p,
                                    C.c_int64, _i32p, _i64p]
    lib.emu_align_batch.restype = C.c_int
    lib.emu_pack.argtypes = [_u8p, C.c_int, _u32p]
    lib.emu_pack.restype = C.c_int

    def run(cfg, batch, wcap=1 << 15, hcap=1 << 25, scap=1 << 18, off16=0):
        seq, po, pl, to, tl = batch
        n = len(pl)
        out = dict(score=np.zeros(n, np.int32), status=np.zeros(n, np.int32), locs=np.zeros((n, 4), np.int32),
                   cig_off=np.zeros(n + 1, np.int64), ovf=np.zeros(n, np.int32), cells=np.zeros(n, np.int64))
        cap = int(pl.sum() + tl.sum()) + 16
        runs = np.zeros(cap, np.uint32)
        rc = lib.emu_align_batch(C.addressof(cfg), np.ascontiguousarray(seq), po, pl, to, tl, n, wcap, hcap, scap, off16,
                                 out["score"], out["status"], out["locs"], out["cig_off"], runs, cap,
                                 out["ovf"], out["cells"])
        assert rc == 0, rc
        out["runs"] = runs[:out["cig_off"][-1]]
        return out
    run.lib = lib
    return run


@pytest.mark.parametrize("case", SYN, ids=[c["name"] for c in SYN])
def test_device_source_matches_golden(emu, oracle, case):
    batch = generate_pairs(case["n"], case["length"], case["div"], case["seed"], text_flank=case["flank"])
    cfg = oracle.make_config(**case["config"])
    r = emu(cfg, batch)
    assert not r["ovf"].any()
    assert r["score"].tolist() == case["score"]
    assert r["status"].tolist() == case["status"]
    cig = [oracle.runs_to_cigarstring(r["runs"][r["cig_off"][j]:r["cig_off"][j + 1]]) for j in range(case["n"])]
    assert cig == case["cigars"]
    assert r["locs"].tolist() == case["locations"]
    if case["config"].get("scope", "full") == "full":
        assert r["cells"].tolist() == case["cells"]


@pytest.mark.parametrize("case", SYN, ids=[c["name"] for c in SYN])
def test_device_source_grid_group_path_matches_golden(emu, oracle, case):
    """the code path of the several-CTAs-per-pair group (G::kGrid: ring loads that bypass L1, int32 rings)"""
    batch = generate_pairs(case["n"], case["length"], case["div"], case["seed"], text_flank=case["flank"])
    cfg = oracle.make_config(**case["config"])
    r = emu(cfg, batch, off16=2)
    assert not r["ovf"].any()
    assert r["score"].tolist() == case["score"] and r["status"].tolist() == case["status"]
    cig = [oracle.runs_to_cigarstring(r["runs"][r["cig_off"][j]:r["cig_off"][j + 1]]) for j in range(case["n"])]
    assert cig == case["cigars"] and r["locs"].tolist() == case["locations"]
    if case["config"].get("scope", "full") == "full":
        assert r["cells"].tolist() == case["cells"]


@pytest.mark.parametrize("case", [c for c in SYN if c["length"] <= 2000], ids=[c["name"] for c in SYN if c["length"] <= 2000])
def test_device_source_int16_rings_match_golden(emu, oracle, case):
    """the short-read tiers keep offsets as int16 (nulls = any negative value)"""
    batch = generate_pairs(case["n"], case["length"], case["div"], case["seed"], text_flank=case["flank"])
    cfg = oracle.make_config(**case["config"])
    r = emu(cfg, batch, wcap=4096, off16=1)
    assert not r["ovf"].any()
    assert r["score"].tolist() == case["score"] and r["status"].tolist() == case["status"]
    cig = [oracle.runs_to_cigarstring(r["runs"][r["cig_off"][j]:r["cig_off"][j + 1]]) for j in range(case["n"])]
    assert cig == case["cigars"] and r["locs"].tolist() == case["locations"]


GCD_CASES = [
    dict(span="end-to-end", max_steps=11), dict(span="end-to-end", max_steps=12), dict(span="end-to-end", max_steps=37),
    dict(heuristic="X-drop", xdrop=10, max_steps=13), dict(heuristic="X-drop", xdrop=15, scope="score"),
    dict(span="end-to-end", mismatch=6, gap_opening=9, gap_extension=3),
    dict(mismatch=6, gap_opening=9, gap_extension=3, heuristic="X-drop", xdrop=12, max_steps=100),
    dict(distance="affine2p", gap_extension2=2), dict(span="end-to-end", mismatch=3, gap_opening=1, gap_extension=1),
]


@pytest.mark.parametrize("kw", GCD_CASES, ids=[str(i) for i in range(len(GCD_CASES))])
def test_score_unit_gcd_edge_cases(emu, oracle, kw):
    """scores advance in units of gcd(penalties); step limits and the unreachable test are
    decided in original units (odd max_steps, drops between two multiples of g, ...)"""
    batch = generate_pairs(300, 150, 0.15, seed=23)
    cfg = oracle.make_config(**kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    for off16 in (0, 1):
        got = emu(cfg, batch, wcap=512, off16=off16)
        for k in ("score", "status", "cig_off", "runs", "locs", "cells"):
            assert np.array_equal(got[k], want[k]), (kw, off16, k)


def test_device_source_matches_oracle_ragged(emu, oracle):
    rng = np.random.default_rng(17)
    pairs = [("", ""), ("ACGT", ""), ("", "ACGT"), ("A", "A"), ("A", "C")]
    for _ in range(300):
        lp, lt = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        pairs.append(("".join("ACGT"[i] for i in rng.integers(0, 4, lp)),
                      "".join("ACGT"[i] for i in rng.integers(0, 4, lt))))
    batch = pairs_from_strings(pairs)
    for kw in (dict(span="end-to-end"), dict(), dict(distance="affine2p"),
               dict(heuristic="adaptive", min_wavefront_length=3, max_distance_threshold=5),
               dict(heuristic="X-drop", xdrop=30), dict(span="end-to-end", scope="score", max_steps=50)):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind="port")
        for mode in (0, 2):
            got = emu(cfg, batch, off16=mode)
            for k in ("score", "status", "cig_off", "runs", "locs", "cells"):
                assert np.array_equal(got[k], want[k]), (kw, mode, k)


M_ONLY_CASES = [
    dict(distance=d, **kw)
    for d in ("linear", "levenshtein", "indel")
    for kw in (dict(span="end-to-end"), dict(), dict(span="end-to-end", scope="score"),
               dict(pattern_begin_free=3, pattern_end_free=5, text_begin_free=4, text_end_free=6),
               dict(span="end-to-end", heuristic="adaptive", min_wavefront_length=5, max_distance_threshold=10,
                    steps_between_cutoffs=2),
               dict(span="end-to-end", max_steps=25))
] + [
    dict(distance="linear", span="end-to-end", match=-2, mismatch=4, gap_extension=3),
    dict(distance="linear", match=-2, mismatch=4, gap_extension=3, pattern_end_free=10, text_end_free=10),
    dict(distance="linear", span="end-to-end", mismatch=2, gap_extension=5),
    dict(distance="linear", span="end-to-end", mismatch=6, gap_extension=4, max_steps=31),
    dict(distance="linear", span="end-to-end", heuristic="X-drop", xdrop=30, steps_between_cutoffs=2),
]


@pytest.mark.parametrize("kw", M_ONLY_CASES, ids=[str(i) for i in range(len(M_ONLY_CASES))])
def test_gap_linear_edit_indel_match_oracle(emu, oracle, kw):
    """the M-only metrics (compute_linear.c / compute_edit.c) run on the scalar tier's source"""
    pairs = [("", ""), ("ACGT", ""), ("", "ACGT"), ("A", "A"), ("A", "C")]
    rng = np.random.default_rng(29)
    for _ in range(150):
        lp, lt = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        pairs.append(("".join("ACGT"[i] for i in rng.integers(0, 4, lp)),
                      "".join("ACGT"[i] for i in rng.integers(0, 4, lt))))
    if kw.get("span") != "end-to-end":
        pairs = [pt for pt in pairs if min(len(pt[0]), len(pt[1])) >= 10]
    cfg = oracle.make_config(**kw)
    for batch in (pairs_from_strings(pairs), generate_pairs(200, 200, 0.1, seed=3)):
        want = oracle.align_batch(cfg, *batch, kind="port")
        for off16 in (0, 1, 2):
            got = emu(cfg, batch, wcap=1024, off16=off16)
            assert not got["ovf"].any()
            for k in ("score", "status", "cig_off", "runs", "locs", "cells"):
                assert np.array_equal(got[k], want[k]), (kw, off16, k)


@pytest.mark.parametrize("metric,affine,sign", [
    (dict(distance="linear", mismatch=2, gap_extension=5), dict(mismatch=2, gap_opening=0, gap_extension=5), 1),
    (dict(distance="linear", match=-2, mismatch=4, gap_extension=3), dict(match=-2, mismatch=4, gap_opening=0, gap_extension=3), 1),
    (dict(distance="levenshtein"), dict(mismatch=1, gap_opening=0, gap_extension=1), -1),
    (dict(distance="indel"), dict(mismatch=2, gap_opening=0, gap_extension=1), -1),
], ids=["linear-2-5", "linear-match-2", "levenshtein", "indel"])
@pytest.mark.parametrize("form", [dict(span="end-to-end"), dict(pattern_end_free=7, text_end_free=12),
                                  dict(pattern_begin_free=3, pattern_end_free=5, text_begin_free=4, text_end_free=6),
                                  dict(span="end-to-end", max_steps=33)], ids=["e2e", "end-free", "four-free", "max-steps"])
def test_score_only_metrics_equal_zero_opening_affine(emu, oracle, metric, affine, sign, form):
    """what metric_as_affine (wfa_params.h) relies on: without a cut-off the M-only recurrences and gap-affine with a
    zero-cost opening reach the same optimum -- same score (edit / indel: sign flipped) and status"""
    if metric.get("match", 0) < 0 and (form.get("pattern_begin_free") or form.get("text_begin_free")):
        pytest.skip("match < 0 with begin-free ends is rejected")
    rng = np.random.default_rng(43)
    pairs = []
    for _ in range(300):
        lp, lt = int(rng.integers(12, 160)), int(rng.integers(12, 160))
        pairs.append(("".join("ACGT"[i] for i in rng.integers(0, 4, lp)), "".join("ACGT"[i] for i in rng.integers(0, 4, lt))))
    for batch in (pairs_from_strings(pairs), generate_pairs(300, 180, 0.12, seed=7, text_flank=8)):
        want = oracle.align_batch(oracle.make_config(scope="score", **metric, **form), *batch, kind="port")
        got = emu(oracle.make_config(scope="score", **affine, **form), batch, wcap=1024)
        assert not got["ovf"].any()
        done = want["status"] == 0
        assert np.array_equal(got["status"], want["status"])
        assert np.array_equal(sign * got["score"][done], want["score"][done])
        assert np.array_equal(got["score"][~done], want["score"][~done])        # -max_steps either way


def test_edit_exact_prune_matches_oracle(emu, oracle):
    """levenshtein end-to-end prunes wavefronts of >= 1000 diagonals (compute_edit.c:219-275)"""
    rng = np.random.default_rng(5)
    rs = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    pairs = []
    for pl, tl in ((300, 3000), (3000, 300), (1500, 2500), (100, 2500)):
        p = rs(pl)
        pairs += [(p, rs(tl)), (p, (p * (tl // pl + 1))[:tl])]
    batch = pairs_from_strings(pairs)
    for kw in (dict(distance="levenshtein", span="end-to-end"), dict(distance="levenshtein", span="end-to-end", scope="score")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind="port")
        got = emu(cfg, batch)
        assert not got["ovf"].any()
        for k in ("score", "status", "cig_off", "runs", "locs", "cells"):
            assert np.array_equal(got[k], want[k]), (kw, k)


def test_capacity_overflow_is_reported_not_wrong(emu, oracle):
    batch = generate_pairs(200, 250, 0.10, seed=2)
    cfg = oracle.make_config(span="end-to-end")
    want = oracle.align_batch(cfg, *batch, kind="port")
    small = emu(cfg, batch, wcap=64)
    assert small["ovf"].any(), "250 bp / 10 % pairs need wavefronts wider than 64"
    ok = small["ovf"] == 0
    assert np.array_equal(small["score"][ok], want["score"][ok])
    tiny_hist = emu(cfg, batch, hcap=2000)
    ok = tiny_hist["ovf"] == 0
    assert (~ok).any() and np.array_equal(tiny_hist["score"][ok], want["score"][ok])
    few_scores = emu(cfg, batch, scap=64)
    ok = few_scores["ovf"] == 0
    assert (~ok).any() and np.array_equal(few_scores["score"][ok], want["score"][ok])


def test_host_packer(emu):
    rng = np.random.default_rng(3)
    for n in (0, 1, 15, 16, 17, 31, 32, 33, 64, 150, 1000):
        codes = rng.integers(0, 4, n)
        s = np.frombuffer("".join("ACTG"[c] for c in codes).encode(), np.uint8).copy()   # A=0 C=1 T=2 G=3
        if n % 2:
            s[::3] |= 0x20                                   # lower case packs alike
        out = np.zeros(n // 16 + 2, np.uint32)
        assert emu.lib.emu_pack(np.ascontiguousarray(s), n, out) == 1
        got = [(int(out[i // 16]) >> (2 * (i % 16))) & 3 for i in range(n)]
        assert got == codes.tolist()
    bad = np.frombuffer(b"ACGTNACGT" * 8, np.uint8).copy()
    assert emu.lib.emu_pack(bad, len(bad), np.zeros(8, np.uint32)) == 0


def _pairs_with_n(seed, n, lo, hi, p_n=0.04, t_n=0.05):
    """Mutated pairs sprinkled with N (and a few lower-case / IUPAC bytes)."""
    rng = np.random.default_rng(seed)
    rnd = lambda m, alpha="ACGT": "".join(alpha[i] for i in rng.integers(0, len(alpha), m))
    pairs = []
    for _ in range(n):
        p = list(rnd(int(rng.integers(lo, hi))))
        t = list(p)
        for j in range(len(t)):
            r = rng.random()
            if r < 0.05: t[j] = rnd(1)
            elif r < 0.08: t[j] = ""
            elif r < 0.11: t[j] += rnd(2)
            elif r < 0.11 + t_n: t[j] = "N"
            elif r < 0.11 + 1.2 * t_n: t[j] = "r"
        for j in range(len(p)):
            if rng.random() < p_n: p[j] = "N"
        pairs.append(("".join(p), "".join(t)))
    return pairs


BYTE_KW = [dict(span="end-to-end"), dict(span="end-to-end", wildcard="N"), dict(distance="affine2p", wildcard="N"),
           dict(heuristic="adaptive", wildcard="N", span="end-to-end"), dict(span="end-to-end", wildcard="A", scope="score"),
           dict(pattern_begin_free=3, text_end_free=5, wildcard="N")]


@pytest.mark.parametrize("kw", BYTE_KW, ids=[str(i) for i in range(len(BYTE_KW))])
def test_byte_mode_non_acgt_and_wildcard(emu, oracle, kw):
    """non-ACGT bytes and pywfa's wildcard= (wildcard_match_fun, pywfa/align.pyx:302-304): the device
    source in byte mode (4 bases per word) against the checker, which itself is checked against the
    reference's wavefront_align_lambda path in test_oracle.py"""
    batch = pairs_from_strings(_pairs_with_n(5, 300, 20, 220))
    cfg = oracle.make_config(**kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = emu(cfg, batch, wcap=2048)
    assert not got["ovf"].any()
    for k in ("score", "status", "cig_off", "runs", "locs"):
        assert np.array_equal(got[k], want[k]), (kw, k)


def test_host_staging_scan_and_gather(emu):
    """The host side of the batch staging (pack.cpp): one parallel pass over the offset / length arrays
    yields what the planner needs (byte range, longest sequences, word count, length-class histogram,
    whether the pairs lie back to back), and scattered pairs are gathered back to back."""
    L = emu.lib
    i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
    L.emu_scan_pairs.argtypes = [i64p, i32p, i64p, i32p, C.c_int64, C.c_int, i64p]
    L.emu_gather_pairs.argtypes = [u8p, i64p, i32p, i64p, i32p, C.c_int64, u8p, i64p, i64p]
    L.emu_length_class_limit.restype = C.c_int
    limits = [L.emu_length_class_limit(c) for c in range(8)]
    assert limits == sorted(limits) and limits[-1] == 2**31 - 1
    rng = np.random.default_rng(8)
    for n in (1, 7, 5000, 300_000):
        p_len = rng.integers(0, 400, n).astype(np.int32)
        t_len = rng.integers(0, 3000, n).astype(np.int32)
        rec = p_len.astype(np.int64) + t_len
        p_off = np.concatenate(([0], np.cumsum(rec)[:-1])).astype(np.int64) + 17
        t_off = p_off + p_len
        for scattered in (False, True):
            if scattered:
                t_off = t_off + 1_000_000
            out = np.zeros(9 + 24, np.int64)
            for bpw in (16, 4):
                L.emu_scan_pairs(p_off, p_len, t_off, t_len, n, bpw, out)
                assert out[0] == -1 and out[1] == (0 if scattered else 1)
                assert out[2] == int(rec.sum())
                assert out[3] == int(((p_len.astype(np.int64) + bpw - 1) // bpw + (t_len.astype(np.int64) + bpw - 1) // bpw).sum())
                nz_p, nz_t = p_len > 0, t_len > 0
                lo = min(int(p_off[nz_p].min()) if nz_p.any() else 2**62, int(t_off[nz_t].min()) if nz_t.any() else 2**62)
                hi = max(int((p_off + p_len)[nz_p].max()) if nz_p.any() else -1, int((t_off + t_len)[nz_t].max()) if nz_t.any() else -1)
                if hi >= 0:
                    assert (out[4], out[5]) == (lo, hi)
                assert (out[6], out[7]) == (int(p_len.max()), int(t_len.max()))
                assert (out[8] >> 32, out[8] & 0xffffffff) == (int(p_len.min()), int(t_len.min()))
                cls = np.searchsorted(limits, np.maximum(p_len, t_len), side="left")
                assert out[9:17].tolist() == np.bincount(cls, minlength=8).tolist()
                for c in range(8):
                    if (cls == c).any():
                        assert out[17 + c] == p_len[cls == c].max() and out[25 + c] == t_len[cls == c].max()
    bad = p_len.copy(); bad[123] = -5
    L.emu_scan_pairs(p_off, bad, t_off, t_len, n, 16, out)
    assert out[0] == 123
    # gather: one pattern against texts spread over a buffer
    seq = rng.integers(65, 91, 200_000).astype(np.uint8)
    m = 3000
    g_toff = np.sort(rng.integers(0, 190_000, m)).astype(np.int64)
    g_tlen = rng.integers(1, 300, m).astype(np.int32)
    g_poff = np.full(m, 199_000, np.int64); g_plen = np.full(m, 77, np.int32)
    dst = np.zeros(int(g_plen.sum() + g_tlen.sum()), np.uint8)
    npo, nto = np.zeros(m, np.int64), np.zeros(m, np.int64)
    L.emu_gather_pairs(seq, g_poff, g_plen, g_toff, g_tlen, m, dst, npo, nto)
    for i in (0, 1, m // 2, m - 1):
        assert np.array_equal(dst[npo[i]:npo[i] + 77], seq[199_000:199_077])
        assert np.array_equal(dst[nto[i]:nto[i] + g_tlen[i]], seq[g_toff[i]:g_toff[i] + g_tlen[i]])
    assert nto[-1] + g_tlen[-1] == len(dst) and np.all(nto == npo + 77)
