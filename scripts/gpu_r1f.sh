( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_cfg2.err | cut -c1-200
grep "device-resident" gpurun_out/bench_cfg2.err
python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_cfg1.err | cut -c1-200
grep "device-resident" gpurun_out/bench_cfg1.err
ncu --set full --clock-control none --import-source on -k regex:wfa_reg -s 3 -c 1 -f -o gpurun_out/prof_r01_reg_cfg2 python bench.py --pairs 1000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_reg -s 3 -c 1 -f -o gpurun_out/prof_r01_reg_cfg1 python bench.py --workload cfg1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
