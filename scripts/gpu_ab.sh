for v in m7 m8 base; do
  if [ $v = base ]; then unset WFAGPU_LIB; else export WFAGPU_LIB=$PWD/pywfa_b200/variants/libwfagpu_$v.so; fi
  echo "== variant $v"
  timeout 300 python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -1
  timeout 300 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -1
done
