set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_cfg2.csv python bench.py --pairs 400000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_align -s 2 -c 2 -f -o gpurun_out/prof_r01_cfg2 python bench.py --pairs 400000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_align -s 0 -c 1 -f -o gpurun_out/prof_r01_cfg1 python bench.py --workload cfg1 --pairs 400000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out
