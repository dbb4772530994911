import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import pywfa_b200
from pywfa_b200 import _ffi
from pywfa_b200.synth import pairs_from_strings, generate_pairs
p, t = open("scripts/_slowpair.txt").read().split()
ctx = _ffi.Context(0)
cfg = pywfa_b200.WavefrontAligner(span="end-to-end")._cfg
norm = generate_pairs(4, 150, 0.05, seed=1)
for name, batch in (("slow pair", pairs_from_strings([(p, t)])), ("4 normal pairs", norm), ("slow x 64", pairs_from_strings([(p, t)] * 64))):
    b = ctx.prepare(cfg, *batch)
    for _ in range(3): b.run()
    t0 = time.perf_counter()
    for _ in range(20): b.run()
    dt = (time.perf_counter() - t0) / 20
    r = b.fetch()
    print(name, "%.3f ms per run" % (dt * 1e3), r["score"][:2], b.stats()["retried_pairs"])
    b.free()
a = pywfa_b200.WavefrontAligner(span="end-to-end")
for _ in range(3): a(t, p)
t0 = time.perf_counter()
for _ in range(50): a(t, p)
print("single-pair call on the slow pair: %.1f us" % ((time.perf_counter() - t0) / 50 * 1e6), a.score)
