( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -5
python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json | cut -c1-200
grep -E "device-resident|e2e" gpurun_out/bench_cfg3.err
python bench.py --workload cfg4-adaptive --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg4a.err | tee gpurun_out/bench_cfg4a.json | cut -c1-200
grep -E "device-resident|e2e" gpurun_out/bench_cfg4a.err
python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json | cut -c1-200
grep -E "device-resident|e2e" gpurun_out/bench_cfg1.err
