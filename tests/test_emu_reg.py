"""CPU check of the register-resident tier: pywfa_b200/csrc/wfa_reg.cuh (the warp code the GPU
runs for short reads) executed on the 32-lane host model tests/emu/lanevec_host.h must reproduce
the oracle bit-exactly -- score, status, CIGAR runs, coordinates and the wavefront cell count --
or report a window overflow (the pair is then retried on a wider tier), never a wrong answer."""
import ctypes as C
import os
import subprocess

import zlib

import numpy as np
import pytest

from pywfa_b200.synth import generate_pairs, pairs_from_strings

HERE = os.path.dirname(os.path.abspath(__file__))
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")


@pytest.fixture(scope="module")
def emu_reg():
    subprocess.run(["make", "-C", os.path.join(HERE, "emu")], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HERE, "emu", "libwfaemu_reg.so"))
    lib.emu_reg_align_batch.argtypes = [C.c_void_p, _u8p, _i64p, _i32p, _i64p, _i32p, C.c_int64, C.c_int, C.c_int,
                                        _i32p, _i32p, _i32p, _i64p, _u32p, C.c_int64, _i32p, _i64p]
    lib.emu_reg_align_batch.restype = C.c_int

    def run(cfg, batch, regs, hrows=None):
        seq, po, pl, to, tl = batch
        n = len(pl)
        out = dict(score=np.zeros(n, np.int32), status=np.zeros(n, np.int32), locs=np.zeros((n, 4), np.int32),
                   cig_off=np.zeros(n + 1, np.int64), ovf=np.zeros(n, np.int32), cells=np.zeros(n, np.int64))
        cap = int(pl.sum() + tl.sum()) + 16
        runs = np.zeros(cap, np.uint32)
        rc = lib.emu_reg_align_batch(C.addressof(cfg), np.ascontiguousarray(seq), po, pl, to, tl, n, regs,
                                     hrows or 32 * regs + 5, out["score"], out["status"], out["locs"],
                                     out["cig_off"], runs, cap, out["ovf"], out["cells"])
        assert rc == 0, rc
        out["runs"] = runs[:out["cig_off"][-1]]
        return out
    return run


def compare(got, want, full):
    """Every pair the tier finished must equal the oracle; returns the number of overflows."""
    done = np.flatnonzero(got["ovf"] == 0)
    for key in ("score", "status", "cells"):
        bad = done[got[key][done] != want[key][done]]
        assert bad.size == 0, f"{key} differs at pairs {bad[:5]}"
    if full:
        for i in done:
            a = got["runs"][got["cig_off"][i]:got["cig_off"][i + 1]]
            b = want["runs"][want["cig_off"][i]:want["cig_off"][i + 1]]
            assert np.array_equal(a, b), f"CIGAR differs at pair {i}"
        assert np.array_equal(got["locs"][done], want["locs"][done])
    return len(got["ovf"]) - len(done)


CASES = [
    # name, config, n, length, divergence, flank, registers, max overflow fraction
    ("cfg1-e2e-full-128", dict(span="end-to-end"), 1500, 150, 0.05, 0, 2, 0.0),
    ("cfg1-e2e-full-64", dict(span="end-to-end"), 1500, 150, 0.05, 0, 1, 0.2),
    ("cfg1-endsfree-default-128", dict(), 1000, 150, 0.05, 0, 2, 0.0),
    ("cfg2-score-256", dict(span="end-to-end", scope="score"), 800, 250, 0.10, 0, 4, 0.0),
    ("cfg2-full-256", dict(span="end-to-end"), 500, 250, 0.10, 0, 4, 0.0),
    ("cfg2-score-192", dict(span="end-to-end", scope="score"), 800, 250, 0.10, 0, 3, 0.15),
    ("cfg2-full-192", dict(span="end-to-end"), 500, 250, 0.10, 0, 3, 0.15),
    ("endsfree-all-four-192", dict(pattern_begin_free=10, pattern_end_free=20, text_begin_free=5, text_end_free=7),
     600, 200, 0.10, 4, 3, 0.1),
    ("cfg2-score-128-overflows", dict(span="end-to-end", scope="score"), 300, 250, 0.10, 0, 2, 1.0),
    ("endsfree-all-four", dict(pattern_begin_free=10, pattern_end_free=20, text_begin_free=5, text_end_free=7),
     1000, 150, 0.10, 4, 4, 0.0),
    ("high-divergence", dict(span="end-to-end"), 800, 100, 0.5, 0, 4, 0.05),
    ("max-steps", dict(span="end-to-end", max_steps=10), 600, 150, 0.1, 0, 2, 0.0),
    ("max-steps-odd", dict(span="end-to-end", max_steps=11, scope="score"), 600, 150, 0.1, 0, 2, 0.0),
    ("long-low-divergence", dict(span="end-to-end"), 60, 2000, 0.01, 0, 4, 0.2),
]


@pytest.mark.parametrize("name,kw,n,length,div,flank,regs,max_ovf", CASES, ids=[c[0] for c in CASES])
def test_register_tier_matches_oracle(emu_reg, oracle, name, kw, n, length, div, flank, regs, max_ovf):
    batch = generate_pairs(n, length, div, seed=zlib.crc32(name.encode()) % 9973, text_flank=flank)
    cfg = oracle.make_config(**kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = emu_reg(cfg, batch, regs)
    novf = compare(got, want, kw.get("scope", "full") == "full")
    assert novf <= max_ovf * n, f"{novf} of {n} pairs overflowed the window"


ZERO_OPEN = [
    # the zero-opening shapes the product instantiates score-only: (x, o + e, e) / gcd = (1, 1, 1) and (2, 1, 1)
    ("edit-like-1-0-1", dict(span="end-to-end", mismatch=1, gap_opening=0, gap_extension=1), 800, 150, 0.08, 0, 2),
    ("indel-like-2-0-1", dict(span="end-to-end", mismatch=2, gap_opening=0, gap_extension=1), 800, 150, 0.08, 0, 2),
    ("linear-like-4-0-2", dict(mismatch=4, gap_opening=0, gap_extension=2, text_begin_free=6, text_end_free=9), 800, 150, 0.08, 6, 3),
    ("edit-like-250bp", dict(span="end-to-end", mismatch=1, gap_opening=0, gap_extension=1), 500, 250, 0.10, 0, 3),
]


@pytest.mark.parametrize("name,kw,n,length,div,flank,regs", ZERO_OPEN, ids=[c[0] for c in ZERO_OPEN])
@pytest.mark.parametrize("scope", ["score", "full"])
def test_register_tier_zero_opening_shapes(emu_reg, oracle, name, kw, n, length, div, flank, regs, scope):
    batch = generate_pairs(n, length, div, seed=zlib.crc32(name.encode()) % 9973, text_flank=flank)
    cfg = oracle.make_config(scope=scope, **kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = emu_reg(cfg, batch, regs)
    assert compare(got, want, scope == "full") <= 0.05 * n


MORE_SHAPES = [
    # (x, o + e, e) / gcd = (4, 7, 1), (1, 2, 1), (1, 3, 1): bwa-like 4/6/1, 1/1/1 (2/2/2), 2/4/2 (1/2/1)
    ("bwa-like-4-6-1", dict(span="end-to-end", gap_extension=1), 800, 150, 0.08, 0, 3, 0.05),
    ("bwa-like-4-6-1-128-overflows", dict(span="end-to-end", gap_extension=1), 400, 150, 0.08, 0, 2, 0.6),
    ("bwa-like-4-6-1-250bp", dict(span="end-to-end", gap_extension=1), 500, 250, 0.08, 0, 4, 0.25),
    ("bwa-like-endsfree", dict(gap_extension=1, pattern_begin_free=10, pattern_end_free=20, text_begin_free=5, text_end_free=7), 600, 200, 0.10, 4, 4, 0.3),
    ("1-1-1", dict(span="end-to-end", mismatch=1, gap_opening=1, gap_extension=1), 800, 150, 0.08, 0, 2, 0.05),
    ("2-2-2", dict(mismatch=2, gap_opening=2, gap_extension=2), 600, 250, 0.10, 0, 4, 0.05),
    ("2-4-2", dict(span="end-to-end", mismatch=2, gap_opening=4, gap_extension=2), 800, 150, 0.08, 0, 2, 0.05),
    ("1-2-1-max-steps", dict(span="end-to-end", mismatch=1, gap_opening=2, gap_extension=1, max_steps=17), 600, 150, 0.1, 0, 3, 0.05),
]


@pytest.mark.parametrize("name,kw,n,length,div,flank,regs,max_ovf", MORE_SHAPES, ids=[c[0] for c in MORE_SHAPES])
@pytest.mark.parametrize("scope", ["score", "full"])
def test_register_tier_more_penalty_shapes(emu_reg, oracle, name, kw, n, length, div, flank, regs, max_ovf, scope):
    batch = generate_pairs(n, length, div, seed=zlib.crc32(name.encode()) % 9973, text_flank=flank)
    cfg = oracle.make_config(scope=scope, **kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = emu_reg(cfg, batch, regs)
    assert compare(got, want, scope == "full") <= max_ovf * n


@pytest.mark.parametrize("kw", [dict(distance="levenshtein", span="end-to-end"), dict(distance="indel", span="end-to-end"),
                                dict(distance="linear", span="end-to-end"), dict(distance="levenshtein"),
                                dict(distance="indel", pattern_end_free=10, text_end_free=10),
                                dict(distance="linear", pattern_begin_free=5, pattern_end_free=8, text_begin_free=6, text_end_free=9),
                                dict(distance="levenshtein", span="end-to-end", max_steps=9),
                                dict(distance="linear", span="end-to-end", max_steps=21)],
                         ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items()))
def test_score_only_metrics_as_zero_opening_affine(emu_reg, oracle, kw):
    """metric_as_affine (wfa_params.h): score-only edit / indel / gap-linear alignments without a cut-off run as
    gap-affine alignments with a zero-cost opening; score and status must equal the M-only recurrence's."""
    rng = np.random.default_rng(41)
    pairs = [("", ""), ("ACGT", ""), ("", "ACGT"), ("A", "A"), ("A", "C")]
    for _ in range(400):
        lp, lt = int(rng.integers(0, 90)), int(rng.integers(0, 90))
        pairs.append(("".join("ACGT"[i] for i in rng.integers(0, 4, lp)), "".join("ACGT"[i] for i in rng.integers(0, 4, lt))))
    if kw.get("span") != "end-to-end":
        pairs = [pt for pt in pairs if min(len(pt[0]), len(pt[1])) >= 10]
    cfg = oracle.make_config(scope="score", **kw)
    for batch in (pairs_from_strings(pairs), generate_pairs(600, 200, 0.1, seed=3, text_flank=6)):
        want = oracle.align_batch(cfg, *batch, kind="port")
        got = emu_reg(cfg, batch, 4)
        done = got["ovf"] == 0
        assert done.sum() > 0.7 * len(done)
        for key in ("score", "status"):
            assert np.array_equal(got[key][done], want[key][done]), (kw, key)


def test_register_tier_ragged_and_empty(emu_reg, oracle):
    rng = np.random.default_rng(5)
    acgt = "ACGT"
    pairs = [("", ""), ("ACGT", ""), ("", "ACGT"), ("A", "A"), ("A", "C"), ("ACGT" * 40, "ACGT" * 40)]
    for _ in range(500):
        lp, lt = int(rng.integers(0, 120)), int(rng.integers(0, 120))
        p = "".join(acgt[i] for i in rng.integers(0, 4, lp))
        if rng.random() < 0.5 and lp:
            cut = int(rng.integers(0, lp))
            t = p[:cut] + "".join(acgt[i] for i in rng.integers(0, 4, int(rng.integers(0, 9)))) + p[cut + int(rng.integers(0, 5)):]
        else:
            t = "".join(acgt[i] for i in rng.integers(0, 4, lt))
        pairs.append((p, t))
    batch = pairs_from_strings(pairs)
    for kw in (dict(span="end-to-end"), dict(), dict(scope="score", span="end-to-end")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind="port")
        got = emu_reg(cfg, batch, 4)
        # unrelated random pairs score beyond what a 256-diagonal window reaches: they overflow
        assert compare(got, want, kw.get("scope", "full") == "full") < len(pairs) // 4


def test_register_tier_origin_rows_overflow(emu_reg, oracle):
    """Too few origin rows for the score: the pair must be handed on, not mis-aligned."""
    batch = generate_pairs(200, 150, 0.10, seed=11)
    cfg = oracle.make_config(span="end-to-end")
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = emu_reg(cfg, batch, 4, hrows=20)
    novf = compare(got, want, True)
    assert 0 < novf < 200


def _pairs_with_symbols(seed, n, lo, hi, extra="N", p_x=0.03, t_x=0.04, odd=0.0):
    """Mutated pairs sprinkled with `extra` symbols (and, with probability `odd` per pair, one byte outside the
    register tier's symbol set)."""
    rng = np.random.default_rng(seed)
    rnd = lambda m, alpha="ACGT": "".join(alpha[i] for i in rng.integers(0, len(alpha), m))
    pairs = []
    for _ in range(n):
        p = list(rnd(int(rng.integers(lo, hi))))
        t = list(p)
        for j in range(len(t)):
            r = rng.random()
            if r < 0.03: t[j] = rnd(1)
            elif r < 0.045: t[j] = ""
            elif r < 0.06: t[j] += rnd(2)
            elif r < 0.06 + t_x: t[j] = rnd(1, extra)
            elif r < 0.06 + 1.1 * t_x: t[j] = t[j].lower()
        for j in range(len(p)):
            if rng.random() < p_x: p[j] = rnd(1, extra)
        if rng.random() < odd and p:
            p[int(rng.integers(0, len(p)))] = "S"
        pairs.append(("".join(p), "".join(t)))
    return pairs


REG_BYTE_KW = [
    dict(span="end-to-end"),                                   # N equals N, differs from every base
    dict(span="end-to-end", wildcard="N"),                     # pywfa's wildcard: N matches everything
    dict(span="end-to-end", wildcard="N", scope="score"),
    dict(wildcard="N", pattern_begin_free=3, text_end_free=5),
    dict(span="end-to-end", wildcard="A"),                     # a base as the wildcard: every pair is a byte pair
    dict(span="end-to-end", wildcard="X"),                     # a wildcard outside the symbol set
    dict(span="end-to-end", wildcard="K"),                     # one of the eight symbols as the wildcard
    dict(span="end-to-end", wildcard="n"),                     # lower case: never equals an (upper-cased) base, like in the reference
    dict(span="end-to-end", gap_extension=1, wildcard="N"),    # shape (4, 7, 1)
    dict(span="end-to-end", mismatch=1, gap_opening=0, gap_extension=1, scope="score", wildcard="N"),   # edit-like, score-only
]


@pytest.mark.parametrize("kw", REG_BYTE_KW, ids=[str(i) for i in range(len(REG_BYTE_KW))])
def test_register_tier_byte_mode(emu_reg, oracle, kw):
    """Byte mode on the register tier (4-bit symbol codes, 8 bases per window word; wfa_reg.cuh CB = 4): pairs with
    N / IUPAC bytes, lower case and pywfa's wildcard= must equal the checker (which is pinned against the reference's
    wavefront_align_lambda path in test_oracle.py); a pair with a byte outside the symbol set is handed on."""
    extra = "NRYKX" if kw.get("wildcard") == "X" else "NRYK"
    pairs = _pairs_with_symbols(7, 500, 20, 240, extra=extra, odd=0.05)
    pairs += [("N", "N"), ("N", "A"), ("NNNN", "ACGT"), ("ACGTNNNNACGT", "ACGTACGT"), ("n" * 40, "N" * 40), ("", "N"), ("N", "")]
    if kw.get("span") != "end-to-end":
        pairs = [pt for pt in pairs if min(len(pt[0]), len(pt[1])) >= 10]
    batch = pairs_from_strings(pairs)
    cfg = oracle.make_config(**kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = emu_reg(cfg, batch, 4)
    novf = compare(got, want, kw.get("scope", "full") == "full")
    # handed on: the ~5 % of the pairs holding an 'S', and window overflows
    odd = sum(("S" in p.upper() or "S" in t.upper()) for p, t in pairs)
    if kw.get("wildcard") == "X":
        odd = len(pairs)       # 'X' itself is fine (the wildcard), but count loosely: X-rich pairs may also overflow
    assert novf <= odd + 0.05 * len(pairs), (novf, odd)
    done = np.flatnonzero(got["ovf"] == 0)
    assert len(done) > 0.8 * len(pairs) or kw.get("wildcard") == "X"
    for i in done:
        assert "S" not in pairs[i][0].upper() and "S" not in pairs[i][1].upper(), "a pair outside the symbol set was aligned"


def test_byte_mode_symbol_table(emu_reg):
    """lv::nib_pack8 (lanevec.cuh): the wildcard is code 0, A C G T N R Y K get eight distinct codes with bit 3 set
    (= "not the wildcard", what the extension's mask tests), every other byte is refused -- for every byte value."""
    lib = C.CDLL(os.path.join(HERE, "emu", "libwfaemu_reg.so"))
    lib.emu_nib_code.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.emu_nib_code.restype = C.c_int
    for wild in (0, ord("N"), ord("A"), ord("X"), ord("n"), 255):
        codes = {}
        for byte in range(256):
            bad = C.c_int(0)
            code = lib.emu_nib_code(byte, wild, C.byref(bad))
            assert code < 0x100, "a valid base leaked into the padding"
            if wild and byte == wild:
                assert (code, bad.value) == (0, 0)
            elif chr(byte) in "ACGTNRYK":
                assert bad.value == 0 and 8 <= code <= 15
                codes[chr(byte)] = code
            else:
                assert bad.value == 1, (byte, wild)
        expect = set("ACGTNRYK") - ({chr(wild)} if wild else set())
        assert set(codes) == expect and len(set(codes.values())) == len(expect)
