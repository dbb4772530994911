import sys, numpy as np
sys.path.insert(0,'.')
from oracle import oracle_py as oracle
from pywfa_b200 import _ffi
from pywfa_b200.synth import pairs_from_strings
rng = np.random.default_rng(5)
acgt = "ACGT"
pairs = [("", ""), ("ACGT", ""), ("", "ACGT"), ("A", "A"), ("A", "C"), ("ACGT" * 40, "ACGT" * 40)]
for _ in range(400):
    lp, lt = int(rng.integers(0, 400)), int(rng.integers(0, 400))
    p = "".join(acgt[i] for i in rng.integers(0, 4, lp))
    if rng.random() < 0.5 and lp:
        cut = int(rng.integers(0, lp))
        t = p[:cut] + "".join(acgt[i] for i in rng.integers(0, 4, int(rng.integers(0, 9)))) + p[cut + int(rng.integers(0, 5)):]
    else:
        t = "".join(acgt[i] for i in rng.integers(0, 4, lt))
    pairs.append((p, t))
batch = pairs_from_strings(pairs)
ctx=_ffi.Context(0)
for kw in (dict(span="end-to-end"), dict(), dict(scope="score", span="end-to-end")):
    cfg = oracle.make_config(**kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    got = ctx.align_batch(cfg, *batch)
    for key in ("score","status"):
        bad=np.flatnonzero(got[key]!=want[key])
        print(kw,key,bad[:10],got[key][bad[:10]],want[key][bad[:10]])
        for i in bad[:3]:
            print(' pair',i,'plen',len(pairs[i][0]),'tlen',len(pairs[i][1]), pairs[i])
