timeout 600 python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e:" | tail -2
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e:" | tail -2
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -3
