import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bench, pywfa_b200
from pywfa_b200 import _ffi
ctx = _ffi.Context(0)
cfg = pywfa_b200.WavefrontAligner(span="end-to-end")._cfg
batch = bench.make_batch(1_000_000, 150, 0.05, 0, seed=1234, workers=8, alloc=_ffi.pinned_empty)
seq, po, pl, to, tl = batch
def run(tag):
    for _ in range(3): ctx.align_batch(cfg, *batch)
    os.environ["WFAGPU_TRACE"] = "1"
    print("====", tag, flush=True); sys.stderr.write("==== %s\n" % tag); sys.stderr.flush()
    t0 = time.perf_counter(); ctx.align_batch(cfg, *batch); dt = time.perf_counter() - t0
    del os.environ["WFAGPU_TRACE"]
    print(tag, "%.2f ms" % (dt * 1e3), flush=True)
run("original")
i = 237146
j = 5
seq[po[i]:po[i] + pl[i]] = 65; seq[to[i]:to[i] + tl[i]] = 65      # all-A pair: trivially identical prefix (pl != tl -> few gaps)
run("outlier 237146 replaced by poly-A")
