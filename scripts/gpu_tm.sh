export WFAGPU_LIB=$PWD/pywfa_b200/variants/libwfagpu_tm.so
export WFAGPU_TRACE=1
timeout 300 python bench.py --workload cfg3 --pairs 40000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier 1|cycles per" | tail -3
timeout 300 python bench.py --workload cfg4-adaptive --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier 0|cycles per" | tail -3
