#!/bin/bash
# ab_variants.sh OUTDIR VARIANT...: device-resident cfg2 / cfg1 rates of the normal build ("main") and of the
# alternative libraries made by scripts/build_variant.sh
out=$1; shift
mkdir -p $out
for v in "$@"; do
  if [ $v = main ]; then unset WFAGPU_LIB; else export WFAGPU_LIB=$PWD/pywfa_b200/csrc/build/libwfagpu_$v.so; fi
  for w in cfg2 cfg1; do
    timeout 200 python bench.py --workload $w --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 3 2> $out/${v}_$w.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $w', round(d['value']/1e6,2), 'M pairs/s', round(d['ms_per_step'],3), 'ms')" | tee -a $out/ab.txt
  done
done
