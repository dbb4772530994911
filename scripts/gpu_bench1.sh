set -x
free -g | head -2; nproc; lscpu | grep "Model name"
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --pairs 1000000 --steps 3 --warmup 3 2> gpurun_out/bench_1m.err | tee gpurun_out/bench_1m.json
tail -12 gpurun_out/bench_1m.err
python bench.py --workload cfg1 --steps 3 --warmup 3 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json
tail -8 gpurun_out/bench_cfg1.err
