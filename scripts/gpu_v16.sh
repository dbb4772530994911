( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -3
export WFAGPU_TRACE=1
timeout 300 python bench.py --workload cfg3 --pairs 40000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier" | tail -4
timeout 300 python bench.py --workload cfg4-adaptive --pairs 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -1
export WFAGPU_VEC_NW=16
timeout 300 python bench.py --workload cfg3 --pairs 40000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier" | tail -3
