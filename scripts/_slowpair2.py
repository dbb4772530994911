import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import pywfa_b200
from pywfa_b200 import _ffi
from pywfa_b200.synth import pairs_from_strings, generate_pairs
p, t = open("scripts/_slowpair.txt").read().split()
ctx = _ffi.Context(0)
cfg = pywfa_b200.WavefrontAligner(span="end-to-end")._cfg
seq, po, pl, to, tl = generate_pairs(131072, 150, 0.05, seed=77)
extra = pairs_from_strings([(p, t)])
def with_extra(k):
    s2 = np.concatenate([seq[:-1], np.tile(extra[0][:-1], k), np.zeros(1, np.uint8)])
    base = len(seq) - 1
    step = len(extra[0]) - 1
    po2 = np.concatenate([po, base + step * np.arange(k) + extra[1][0]]).astype(np.int64)
    to2 = np.concatenate([to, base + step * np.arange(k) + extra[3][0]]).astype(np.int64)
    pl2 = np.concatenate([pl, np.repeat(extra[2], k)]).astype(np.int32)
    tl2 = np.concatenate([tl, np.repeat(extra[4], k)]).astype(np.int32)
    return s2, po2, pl2, to2, tl2
for k in (0, 1, 8):
    batch = with_extra(k)
    b = ctx.prepare(cfg, *batch)
    for _ in range(3): b.run()
    t0 = time.perf_counter()
    for _ in range(10): b.run()
    dt = (time.perf_counter() - t0) / 10
    print("131072 pairs +", k, "outliers: %.3f ms per run" % (dt * 1e3), "retried", b.stats()["retried_pairs"], flush=True)
    b.free()
