timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e:"
timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep device-resident
timeout 300 python bench.py --workload cfg4-adaptive --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep device-resident
