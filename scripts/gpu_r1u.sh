for t in 128 256 384 512; do
export WFAGPU_BLOCK_THREADS=$t
echo "== threads $t"
python bench.py --workload cfg3 --pairs 40000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "device-resident"
done
