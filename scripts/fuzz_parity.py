"""Randomised differential check of the CUDA path against the CPU checker: random penalties, spans,
free ends, cut-offs, lengths and divergences (incl. unequal lengths and N-holding pairs).
    python scripts/fuzz_parity.py [rounds] [seed] [bytes]
(third argument "bytes": every round draws N / IUPAC-holding pairs, half of them with the wildcard -- the byte-mode tiers)
Prints one line per round; exits non-zero on the first mismatch (with the offending configuration)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import assert_same
from oracle import oracle_py
from pywfa_b200 import _ffi
from pywfa_b200.build import build_library
from pywfa_b200.synth import generate_pairs, pairs_from_strings
from test_gpu_parity import _ragged_pairs
from test_emu import _pairs_with_n

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
only_bytes = len(sys.argv) > 3 and sys.argv[3] == "bytes"
rng = np.random.default_rng(seed)
build_library()
oracle_py.build(ref=None if not oracle_py.have_ref() else False)
ctx = _ffi.Context(0)
t0 = time.time()
for r in range(rounds):
    kw = {}
    d = rng.random()
    if d < 0.4:
        kw["distance"] = "affine2p"
    elif d < 0.65:
        kw["distance"] = str(rng.choice(["linear", "levenshtein", "indel"]))
    if rng.random() < 0.5:
        kw["span"] = "end-to-end"
    if rng.random() < 0.3:
        kw["scope"] = "score"
    if rng.random() < 0.6:
        kw["mismatch"] = int(rng.integers(1, 9)); kw["gap_opening"] = int(rng.integers(0, 10)); kw["gap_extension"] = int(rng.integers(1, 5))
        if kw.get("distance") == "affine2p":
            kw["gap_opening2"] = int(rng.integers(kw["gap_opening"], 40)); kw["gap_extension2"] = int(rng.integers(1, 4))
    if rng.random() < 0.15:
        kw["match"] = -int(rng.integers(1, 4))
    h = rng.random()
    if h < 0.25:
        kw.update(heuristic="adaptive", min_wavefront_length=int(rng.integers(2, 30)), max_distance_threshold=int(rng.integers(5, 80)),
                  steps_between_cutoffs=int(rng.integers(1, 5)))
    elif h < 0.45 and kw.get("distance") not in ("levenshtein", "indel"):     # the drops are refused with edit / indel
        kw.update(heuristic="X-drop", xdrop=int(rng.integers(5, 300)), steps_between_cutoffs=int(rng.integers(1, 5)))
    if rng.random() < 0.15:
        kw["max_steps"] = int(rng.integers(5, 400))
    shape = 1.0 if only_bytes else rng.random()
    if shape < 0.4:
        length, div = int(rng.integers(20, 600)), float(rng.choice([0.01, 0.05, 0.1, 0.2, 0.4]))
        batch = generate_pairs(int(rng.integers(200, 3000)), length, div, seed=int(rng.integers(1 << 30)))
        minlen = length // 3
    elif shape < 0.08 + 0.4:
        # long reads: the 16-warp groups, the scalar shared-memory tiers (> 12 kbp with a cut-off) and the
        # several-CTAs-per-pair tier
        length, div = int(rng.integers(5000, 30000)), float(rng.choice([0.02, 0.05, 0.1]))
        batch = generate_pairs(int(rng.integers(2, 7)), length, div, seed=int(rng.integers(1 << 30)))
        minlen = length // 3
    elif shape < 0.6:
        length, div = int(rng.integers(600, 4000)), float(rng.choice([0.01, 0.05, 0.1, 0.2]))
        batch = generate_pairs(int(rng.integers(8, 60)), length, div, seed=int(rng.integers(1 << 30)))
        minlen = length // 3
    elif shape < 0.85:
        lo = int(rng.integers(8, 40))
        batch = pairs_from_strings(_ragged_pairs(int(rng.integers(1 << 30)), int(rng.integers(100, 800)), lo, int(rng.integers(lo + 10, 500))))
        minlen = 0
    else:
        hi = int(rng.integers(60, 400)) if rng.random() < 0.7 else int(rng.integers(400, 3000))
        batch = pairs_from_strings(_pairs_with_n(int(rng.integers(1 << 30)), int(rng.integers(100, 1500)) if hi < 400 else int(rng.integers(20, 200)),
                                                 20, hi, p_n=float(rng.choice([0.0005, 0.01])), t_n=float(rng.choice([0.0005, 0.02]))))
        if rng.random() < 0.6:
            kw["wildcard"] = "N"
        minlen = 0
    if kw.get("span", "ends-free") == "ends-free" and kw.get("match", 0) == 0 and rng.random() < 0.5:
        m = int(min(batch[2].min(), batch[4].min()))
        if m > 0:
            for f in ("pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free"):
                if rng.random() < 0.5:
                    kw[f] = int(rng.integers(0, min(m, 60) + 1))
    cfg = oracle_py.make_config(**kw)
    want = oracle_py.align_batch(cfg, *batch, kind="port")
    got = ctx.align_batch(cfg, *batch)
    try:
        assert_same(got, want, scope_full=kw.get("scope", "full") == "full", what=str(kw))
    except AssertionError as e:
        print(f"round {r}: MISMATCH n={len(batch[1])} maxlen={int(max(batch[2].max(), batch[4].max()))} {kw}\n{e}")
        import os
        for env in ({"WFAGPU_VEC_NW": "1"}, {"WFAGPU_VEC_NW": "8"}, {"WFAGPU_VEC_NW": "16"}, {"WFAGPU_NO_VEC_TIER": "1"},
                    {"WFAGPU_NO_TIER_SKIP": "1"}):
            os.environ.update(env)
            g2 = ctx.align_batch(cfg, *batch)
            bad = np.flatnonzero((g2["score"] != want["score"]) | (g2["status"] != want["status"]))
            print(f"   with {env}: {bad.size} pairs differ", bad[:8].tolist())
            for k in env:
                del os.environ[k]
        bad = np.flatnonzero((got["score"] != want["score"]) | (got["status"] != want["status"]))
        i = int(bad[0])
        seq = batch[0].tobytes()
        print("   first bad pair:", i, "plen", int(batch[2][i]), "tlen", int(batch[4][i]), "want", int(want["score"][i]), int(want["status"][i]),
              "got", int(got["score"][i]), int(got["status"][i]))
        print("   P=", seq[batch[1][i]:batch[1][i] + batch[2][i]].decode())
        print("   T=", seq[batch[3][i]:batch[3][i] + batch[4][i]].decode())
        sys.exit(1)
    print(f"round {r}: ok n={len(batch[1])} maxlen={int(max(batch[2].max(), batch[4].max()))} {kw}", flush=True)
print(f"{rounds} rounds bit-exact in {time.time() - t0:.0f}s")
