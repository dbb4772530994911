"""Device-resident rate of 1M x 150 bp (5 %) end-to-end alignments for several penalty sets: which ones the
register-resident tier takes (shape (x, o+e, e)/gcd instantiated in wfa_kernels.cu) and what the others cost on the
packed-halfword tier.    python scripts/penalty_shapes.py [pairs]"""
import sys
import time

sys.path.insert(0, ".")
import pywfa_b200
from pywfa_b200 import _ffi
from pywfa_b200.synth import generate_pairs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = _ffi.Context(0)
batch = generate_pairs(n, 150, 0.05, seed=11)
for name, kw in (("default 4/6/2", {}), ("bwa-like 4/6/1", dict(gap_extension=1)), ("1/1/1", dict(mismatch=1, gap_opening=1, gap_extension=1)),
                 ("2/4/2", dict(mismatch=2, gap_opening=4, gap_extension=2)), ("6/10/2", dict(mismatch=6, gap_opening=10, gap_extension=2)),
                 ("match -1, 4/6/2", dict(match=-1)), ("affine2p default", dict(distance="affine2p")),
                 ("levenshtein", dict(distance="levenshtein")), ("linear 4/2", dict(distance="linear"))):
    for scope in ("score", "full"):
        cfg = pywfa_b200.WavefrontAligner(span="end-to-end", scope=scope, **kw)._cfg
        b = ctx.prepare(cfg, *batch)
        for _ in range(2):
            b.run()
        t0 = time.perf_counter()
        for _ in range(3):
            b.run()
        dt = (time.perf_counter() - t0) / 3
        print(f"{name:18s} scope={scope:5s}: {n / dt / 1e6:7.1f} M pairs/s", flush=True)
        b.free()
