mkdir -p gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in cfg1 cfg3 cfg4-adaptive cfg4-none; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err
done
timeout 900 python bench.py --workload cfg5 --steps 1 --warmup 3 --no-e2e > gpurun_out/final/bench_cfg5.json 2> gpurun_out/final/bench_cfg5.err
export WFAGPU_VEC_NW=8
ncu --set full --clock-control none --import-source on -k regex:wfa_vec_kernel -s 3 -c 1 -f -o gpurun_out/final/prof_vec_cfg3_nw8 python bench.py --workload cfg3 --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final/ncu_b.log 2>&1
unset WFAGPU_VEC_NW
ncu --set full --clock-control none --import-source on -k regex:wfa_vec_kernel -s 3 -c 1 -f -o gpurun_out/final/prof_vec_cfg4a python bench.py --workload cfg4-adaptive --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final/ncu_c.log 2>&1
