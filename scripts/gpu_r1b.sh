set -x
export WFAGPU_TRACE=1
( time python bench.py --steps 3 --warmup 3 ) 2> gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json | cut -c1-2500
grep -v "^\[wfagpu\]   " gpurun_out/bench_cfg2.err | tail -25
python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json | cut -c1-1200
grep -v "^\[wfagpu\]   " gpurun_out/bench_cfg1.err | tail -12
