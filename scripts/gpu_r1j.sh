( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -5
export WFAGPU_TRACE=1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json | cut -c1-300
grep -v "^\[wfagpu\]   " gpurun_out/bench_cfg2.err | tail -7
python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json | cut -c1-300
grep -v "^\[wfagpu\]   " gpurun_out/bench_cfg1.err | tail -9
