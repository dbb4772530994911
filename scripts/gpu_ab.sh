for v in 01 10 00 11; do
  if [ $v = 11 ]; then unset WFAGPU_LIB; else export WFAGPU_LIB=$PWD/pywfa_b200/variants/libwfagpu_$v.so; fi
  echo "== variant simd_nw1/seqw = $v"
  timeout 300 python bench.py --workload cfg4-adaptive --pairs 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "device-resident" | tail -1
done
