export WFAGPU_LIB=$PWD/pywfa_b200/variants/libwfagpu_tm.so
export WFAGPU_TRACE=1
timeout 300 python bench.py --workload cfg4-adaptive --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|cycles per|scanned-range" | tail -3
