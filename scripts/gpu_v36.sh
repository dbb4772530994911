for i in 1 2; do timeout 300 python bench.py --workload cfg4-adaptive --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -1; done
timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -1
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -3
