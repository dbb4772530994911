set -x
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -12
export WFAGPU_TRACE=1
python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json | cut -c1-900
grep -v "^\[wfagpu\]   " gpurun_out/bench_cfg2.err | tail -8
python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json | cut -c1-900
grep -v "^\[wfagpu\]   " gpurun_out/bench_cfg1.err | tail -8
