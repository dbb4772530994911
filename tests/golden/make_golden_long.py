#!/usr/bin/env python
"""Golden vectors for cfg5-shaped long reads (100 kbp, 20 %, gap-affine-2p, end-to-end, full CIGAR)
from the UNMODIFIED reference C library in its low-memory mode (CIGAR-identical to the default
mode, SURVEY.md 0.4, which would need ~200 GB per pair).  One pair takes the reference minutes:

    python tests/golden/make_golden_long.py [n_pairs] [length]

Inputs are regenerated from the seed by pywfa_b200.synth.generate_pairs, so only score, status,
run count and a SHA-256 of the run words are stored (tests/golden/long_reads.json)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402
from pywfa_b200.synth import generate_pairs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
length = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
seed, div = 7, 0.20
oracle_py.build()
assert oracle_py.have_ref(), "needs oracle/_ref (make -C oracle ref)"
kw = dict(distance="affine2p", span="end-to-end")
batch = generate_pairs(n, length, div, seed=seed)
t0 = time.time()
r = oracle_py.align_batch(oracle_py.make_config(**kw), *batch, kind="reference", memory_mode="low")
out = dict(generator=dict(n=n, length=length, div=div, seed=seed), config=kw, memory_mode="low",
           reference_seconds=round(time.time() - t0, 1), pairs=[])
for i in range(n):
    runs = np.ascontiguousarray(r["runs"][r["cig_off"][i]:r["cig_off"][i + 1]], np.uint32)
    out["pairs"].append(dict(score=int(r["score"][i]), status=int(r["status"][i]), nruns=int(len(runs)),
                             runs_sha256=hashlib.sha256(runs.tobytes()).hexdigest()))
path = os.path.join(HERE, f"long_reads_{length // 1000}kbp.json")
json.dump(out, open(path, "w"), indent=1)
print(path, out)
