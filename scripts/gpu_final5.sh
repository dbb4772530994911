mkdir -p gpurun_out/final
( time python bench.py ) > gpurun_out/final/bench_cfg2.json 2> gpurun_out/final/bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref_cfg2.json 2> gpurun_out/final/bench_ref_cfg2.err
for w in cfg1 cfg3 cfg4-adaptive cfg4-xdrop cfg4-none; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err
done
timeout 900 python bench.py --workload cfg5 --steps 1 --warmup 2 --no-e2e > gpurun_out/final/bench_cfg5.json 2> gpurun_out/final/bench_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_reg_kernel -s 6 -c 1 -f -o gpurun_out/final/prof_reg_cfg2_10M python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wfa_vec_kernel -s 3 -c 1 -f -o gpurun_out/final/prof_vec_cfg4a python bench.py --workload cfg4-adaptive --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final/ncu_c.log 2>&1
