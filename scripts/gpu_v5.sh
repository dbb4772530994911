export WFAGPU_TRACE=1
timeout 300 python bench.py --workload cfg4-adaptive --pairs 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -2
timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e|chunks|tier|single" | tail -24
timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e|chunks" | tail -12
nproc; lscpu | grep -E "Model name|Socket|Thread|Core"
