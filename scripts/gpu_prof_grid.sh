mkdir -p gpurun_out/final
cat > /tmp/grid_run.py <<'PY'
import sys
sys.path.insert(0, ".")
from oracle import oracle_py
from pywfa_b200 import _ffi
from pywfa_b200.synth import generate_pairs
ctx = _ffi.Context(0)
cfg = oracle_py.make_config(distance="affine2p", span="end-to-end")
batch = generate_pairs(8, 20000, 0.20, seed=7)
b = ctx.prepare(cfg, *batch)
for _ in range(3):
    b.run()
r = b.fetch()
print(r["score"].tolist(), b.stats())
PY
ncu --set full --clock-control none --import-source on -k regex:wfa_grid_kernel -s 2 -c 1 -f -o gpurun_out/final/prof_grid_20kbp python /tmp/grid_run.py > gpurun_out/final/ncu_g.log 2>&1
tail -2 gpurun_out/final/ncu_g.log
