mkdir -p gpurun_out/r2j
for v in narrow2 main; do
  if [ $v = main ]; then unset WFAGPU_LIB; else export WFAGPU_LIB=$PWD/pywfa_b200/csrc/build/libwfagpu_$v.so; fi
  timeout 100 python bench.py --workload cfg2 --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 2 2> gpurun_out/r2j/${v}_cfg2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v cfg2', round(d['value']/1e6,2), 'M pairs/s', round(d['ms_per_step'],3), 'ms')" | tee -a gpurun_out/r2j/ab.txt
done
