set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; lscpu | grep "Model name"; free -g | head -2
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8
( time python bench.py --steps 3 --warmup 3 ) 2> gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json | cut -c1-1500
tail -15 gpurun_out/bench_cfg2.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-600
tail -6 gpurun_out/bench_ref.err
python bench.py --workload cfg1 --steps 3 --warmup 3 2> gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json | cut -c1-1500
tail -8 gpurun_out/bench_cfg1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_cfg2.csv python bench.py --pairs 1000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_align -s 6 -c 2 -f -o gpurun_out/prof_r01_cfg2 python bench.py --pairs 1000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
