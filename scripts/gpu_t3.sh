export WFAGPU_TRACE=1
python bench.py --pairs 1000000 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "^{" | tail -12
python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "^{" | tail -12
