mkdir -p gpurun_out/final
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final/bench_n2.json 2> gpurun_out/final/bench_n2.err
tail -2 gpurun_out/final/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/final/bench_ref_n2.json 2> gpurun_out/final/bench_ref_n2.err
cat gpurun_out/final/bench_ref_n2.json | cut -c1-200
