for mb in 6 7; do
export WFAGPU_LIB=$PWD/pywfa_b200/libwfagpu_mb$mb.so
echo "== minblocks $mb"
python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "device-resident"
python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "device-resident"
done
unset WFAGPU_LIB
echo "== default"
python bench.py --pairs 2000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "device-resident"
python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "device-resident"
