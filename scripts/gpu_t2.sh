set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --pairs 1000000 --steps 3 --warmup 3 2> gpurun_out/bench_1m.err | tee gpurun_out/bench_1m.json | cut -c1-400
tail -5 gpurun_out/bench_1m.err
python bench.py --workload cfg1 --steps 3 --warmup 3 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json | cut -c1-400
tail -5 gpurun_out/bench_cfg1.err
