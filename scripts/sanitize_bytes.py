"""A few small N-holding batches through every byte-mode tier (register, packed-halfword with 1 / 8 / 16 warps per pair),
checked against the CPU checker: the workload of the compute-sanitizer runs of these kernels.
    compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_bytes.py"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import assert_same
from oracle import oracle_py
from pywfa_b200 import _ffi
from pywfa_b200.synth import pairs_from_strings
from test_emu_reg import _pairs_with_symbols

ctx = _ffi.Context(0)
cases = [(dict(span="end-to-end", wildcard="N"), (60, 200), 96, None),
         (dict(span="end-to-end"), (60, 200), 96, None),
         (dict(distance="affine2p", wildcard="N"), (200, 500), 48, "1"),
         (dict(distance="affine2p"), (400, 900), 24, "8"),
         (dict(heuristic="adaptive", span="end-to-end", wildcard="N"), (900, 1500), 12, "16")]
for kw, lens, n, nw in cases:
    if nw:
        os.environ["WFAGPU_VEC_NW"] = nw
    else:
        os.environ.pop("WFAGPU_VEC_NW", None)
    batch = pairs_from_strings(_pairs_with_symbols(3, n, lens[0], lens[1], extra="NRYK", p_x=0.02, t_x=0.02, odd=0.05))
    cfg = oracle_py.make_config(**kw)
    want = oracle_py.align_batch(cfg, *batch, kind="port")
    got = ctx.align_batch(cfg, *batch)
    assert_same(got, want, scope_full=True, what=str(kw))
    print("ok", kw, "vec warps per pair:", nw, flush=True)
ctx.close()
