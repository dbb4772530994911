( python -m pytest tests -x -q -m gpu 2>&1 ) | tail -3
ncu --set full --clock-control none --import-source on -k regex:wfa_align_kernel -s 14 -c 1 -f -o gpurun_out/prof_r01_cfg3_block python bench.py --workload cfg3 --pairs 20000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
