( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -8
export WFAGPU_TRACE=1
timeout 300 python bench.py --workload cfg3 --pairs 40000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident|tier" | tail -4
