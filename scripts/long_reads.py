"""Timing + self-consistency of cfg5-shaped long reads (affine2p, end-to-end, scope=full)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import oracle_py
from pywfa_b200 import _ffi
from pywfa_b200.synth import generate_pairs
from test_gpu_parity import _check_cigar
ctx = _ffi.Context(0)
cfg = oracle_py.make_config(distance="affine2p", span="end-to-end")
for n, length in ((8, 10000), (4, 40000), (int(sys.argv[1]) if len(sys.argv) > 1 else 2, 100000)):
    batch = generate_pairs(n, length, 0.20, seed=7)
    t0 = time.perf_counter()
    b = ctx.prepare(cfg, *batch)
    b.run(); r = b.fetch()
    dt = time.perf_counter() - t0
    st = b.stats()
    seq, po, pl, to, tl = batch
    buf = seq.tobytes().decode()
    ok = True
    for i in range(n):
        runs = r["runs"][r["cig_off"][i]:r["cig_off"][i + 1]]
        cost = _check_cigar(runs, buf[po[i]:po[i] + pl[i]], buf[to[i]:to[i] + tl[i]], 4, 6, 2, 24, 1)
        ok = ok and (-cost == r["score"][i]) and r["status"][i] == 0
    print(f"{n} x {length} bp: {dt:.2f} s ({n / dt:.3f} pairs/s), cells {st['cells']:,} ({st['cells'] / dt / 1e9:.2f} Gcell/s), "
          f"history {st['history_bytes'] / 1e9:.1f} GB, launches {st['kernel_launches']}, scores {r['score'].tolist()}, consistent {ok}", flush=True)
    b.free()
