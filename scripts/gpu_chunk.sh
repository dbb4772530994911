for c in 0 65536 131072 196608 262144; do
  if [ $c = 0 ]; then unset WFAGPU_CHUNK; else export WFAGPU_CHUNK=$c; fi
  echo "== chunk $c"
  timeout 300 python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "e2e:" | tail -1
done
