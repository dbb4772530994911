"""What a read with an N costs: device-resident rate of 150 bp / 250 bp batches in which a fraction of the pairs holds an
N, with the byte-mode register tier (wfa_reg_bytes.cu) and with those pairs on the scalar tiers (WFAGPU_NO_REG_BYTES=1),
for the default penalties with and without wildcard="N".    python scripts/byte_pairs_rate.py [pairs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import pywfa_b200
from pywfa_b200 import _ffi
from pywfa_b200.synth import generate_pairs

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = _ffi.Context(0)
for length, div, base_kw, fracs in ((150, 0.05, dict(span="end-to-end"), (0.0, 0.02, 0.2, 1.0)), (250, 0.10, dict(span="end-to-end"), (0.0, 0.02, 0.2, 1.0)),
                                    (1000, 0.10, dict(distance="affine2p"), (0.0, 0.05, 1.0)), (1000, 0.10, dict(span="end-to-end"), (0.0, 0.05, 1.0))):
    n = n0 if length < 1000 else max(1000, n0 // 20)
    seq, po, pl, to, tl = generate_pairs(n, length, div, seed=11)
    for frac in fracs:
        s = seq.copy()
        rng = np.random.default_rng(5)
        dirty = np.flatnonzero(rng.random(n) < frac)
        # one N in the pattern and one in the text of every dirty pair
        s[po[dirty] + rng.integers(0, np.maximum(pl[dirty], 1))] = ord("N")
        s[to[dirty] + rng.integers(0, np.maximum(tl[dirty], 1))] = ord("N")
        for kw in (dict(), dict(wildcard="N")):
            for scope in ("score", "full"):
                cfg = pywfa_b200.WavefrontAligner(scope=scope, **base_kw, **kw)._cfg
                rates = []
                for off in (False, True):
                    if off:
                        os.environ["WFAGPU_NO_REG_BYTES"] = "1"
                    else:
                        os.environ.pop("WFAGPU_NO_REG_BYTES", None)
                    b = ctx.prepare(cfg, s, po, pl, to, tl)
                    for _ in range(2):
                        b.run()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        b.run()
                    rates.append(n / ((time.perf_counter() - t0) / 3) / 1e6)
                    b.free()
                print(f"{n} x {length} bp {base_kw}, {100 * frac:5.1f} % of the pairs with N, {str(kw):20s} scope={scope:5s}: "
                      f"{rates[0]:7.1f} M pairs/s with the byte-mode fast tiers, {rates[1]:7.1f} without", flush=True)
os.environ.pop("WFAGPU_NO_REG_BYTES", None)
