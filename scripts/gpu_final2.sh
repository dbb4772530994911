mkdir -p gpurun_out/final
export WFAGPU_VEC_NW=8
ncu --set full --clock-control none --import-source on -k regex:wfa_vec_kernel -s 3 -c 1 -f -o gpurun_out/final/prof_vec_cfg3_nw8 python bench.py --workload cfg3 --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/final/ncu_b.log 2>&1
