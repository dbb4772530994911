mkdir -p gpurun_out/r2d
for v in main base; do
  if [ $v = main ]; then unset WFAGPU_LIB; else export WFAGPU_LIB=$PWD/pywfa_b200/csrc/build/libwfagpu_$v.so; fi
  for w in cfg2 cfg1; do
    timeout 200 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 3 2> gpurun_out/r2d/${v}_$w.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $w', round(d['value']/1e6,2), 'M pairs/s', d['ms_per_step'])" | tee -a gpurun_out/r2d/ab.txt
  done
done
unset WFAGPU_LIB
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg1 or cfg2 or ragged or register or shapes or byte" 2>&1 | tail -3 | tee -a gpurun_out/r2d/ab.txt
