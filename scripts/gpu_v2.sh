export WFAGPU_TRACE=1
timeout 300 python bench.py --workload cfg3 --pairs 40000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier" | tail -4
timeout 300 python bench.py --workload cfg4-adaptive --pairs 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier" | tail -5
unset WFAGPU_TRACE
ncu --set full --clock-control none --import-source on -k regex:wfa_vec_kernel -s 9 -c 1 -f -o gpurun_out/prof_r01_vec_cfg3 python bench.py --workload cfg3 --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_vec_kernel -s 4 -c 1 -f -o gpurun_out/prof_r01_vec_cfg4a python bench.py --workload cfg4-adaptive --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_v4.log 2>&1
