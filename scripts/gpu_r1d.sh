python scripts/dbg_ragged.py 2>&1 | tail -30
ncu --set full --clock-control none --import-source on -k regex:wfa_reg -s 3 -c 1 -f -o gpurun_out/prof_r01_reg_cfg2 python bench.py --pairs 1000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_reg -s 3 -c 1 -f -o gpurun_out/prof_r01_reg_cfg1 python bench.py --workload cfg1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_c.log gpurun_out/ncu_d.log
