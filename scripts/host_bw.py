"""What the host side of an 8-GPU box can feed: (a) aggregate pinned H2D bandwidth with all GPUs copying at
once (one process per GPU, like the bench), (b) aggregate DRAM read bandwidth of the CPU cores (what any
host-side packing would be bound by).  Usage: python scripts/host_bw.py NGPUS"""
import multiprocessing as mp
import sys
import time

import numpy as np


def gpu_worker(dev, q, barrier, mb=1024, reps=6):
    import torch
    torch.cuda.set_device(dev)
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
    h.fill_(65)
    d = torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    q.put((dev, reps * (mb << 20) / dt / 1e9))


def cpu_worker(i, q, barrier, mb=512, reps=4):
    a = np.full(mb << 20, 65, np.uint8)
    a.sum()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        a.view(np.uint64).sum()
    dt = time.perf_counter() - t0
    q.put((i, reps * (mb << 20) / dt / 1e9))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    ctx = mp.get_context("spawn")
    for k in sorted({1, n}):
        q, bar = ctx.Queue(), ctx.Barrier(k)
        ps = [ctx.Process(target=gpu_worker, args=(i, q, bar)) for i in range(k)]
        [p.start() for p in ps]; [p.join() for p in ps]
        r = sorted(q.get() for _ in range(k))
        print(f"H2D pinned, {k} GPU(s) at once: " + " ".join(f"{v:.1f}" for _, v in r) + f" GB/s  (sum {sum(v for _, v in r):.1f})")
    import os
    cores = len(os.sched_getaffinity(0))
    for k in sorted({1, cores // 2, cores}):
        q, bar = ctx.Queue(), ctx.Barrier(k)
        ps = [ctx.Process(target=cpu_worker, args=(i, q, bar)) for i in range(k)]
        [p.start() for p in ps]; [p.join() for p in ps]
        r = [q.get()[1] for _ in range(k)]
        print(f"CPU DRAM read, {k} process(es): sum {sum(r):.1f} GB/s (min {min(r):.1f} per core)")
