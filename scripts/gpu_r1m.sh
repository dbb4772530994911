nvidia-smi -L
python -m pytest tests/test_shard.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --pairs 4000000 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | cut -c1-1500
grep -E "rank|e2e|Error|error" gpurun_out/bench_n2.err | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -3 | cut -c1-400
