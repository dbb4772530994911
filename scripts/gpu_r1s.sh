( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -4
python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json | cut -c1-200
grep -E "device-resident|e2e" gpurun_out/bench_cfg3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_r01_cfg3.csv python bench.py --workload cfg3 --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
