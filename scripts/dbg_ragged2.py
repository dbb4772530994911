import sys, os, numpy as np
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
ctx=_ffi.Context(0)
cfg = oracle.make_config(span="end-to-end", scope="score")
def run(ps):
    b=pairs_from_strings(ps)
    return ctx.align_batch(cfg,*b)['score'], oracle.align_batch(cfg,*b,kind="port")['score']
for i in (140,143,194,214,333,402):
    g,w=run([pairs[i]]); print('alone',i,g,w)
g,w=run(pairs); bad=np.flatnonzero(g!=w); print('all',bad)
g,w=run(pairs[100:200]); bad=np.flatnonzero(g!=w); print('100:200',bad+100)
g,w=run(pairs[140:141]*50); bad=np.flatnonzero(g!=w); print('x50',bad)
# synthetic: gaps of given length in the middle of a 300bp read
base="".join(acgt[i] for i in rng.integers(0,4,300))
for L in range(1,9):
    ins="".join(acgt[i] for i in rng.integers(0,4,L))
    ps=[(base, base[:150]+ins+base[150:]), (base[:150]+ins+base[150:], base)]
    g,w=run(ps); print('gap',L,g,w)
