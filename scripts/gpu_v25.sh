timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e:" | tail -3
