"""Multi-GPU sharding: independent pairs, contiguous work-balanced shards, host-side
gather, no collective on the data path.  The planning / merging / gathering logic is exercised on
CPU with the oracle standing in for the per-rank engine (world_size-2 gloo processes); on the GPU
box the same code runs with real contexts."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pywfa_b200.shard import align_multi_device, merge_results, plan_shards, slice_batch  # noqa: E402
from pywfa_b200.synth import generate_pairs  # noqa: E402


def test_plan_shards_covers_and_balances():
    rng = np.random.default_rng(3)
    p_len = rng.integers(50, 2000, 5000).astype(np.int32)
    t_len = (p_len + rng.integers(-20, 20, 5000)).astype(np.int32)
    for k in (1, 2, 3, 8):
        sh = plan_shards(p_len, t_len, k)
        assert len(sh) == k and sh[0][0] == 0 and sh[-1][1] == 5000
        assert all(sh[i][1] == sh[i + 1][0] for i in range(k - 1))
        ln = p_len.astype(np.float64) + t_len
        w = ln + ln * ln / 64 + 1
        loads = [w[a:b].sum() for a, b in sh]
        assert max(loads) <= 1.1 * (w.sum() / k) + w.max()
    assert plan_shards(p_len[:1], t_len[:1], 4) == [(0, 0), (0, 0), (0, 0), (0, 1)] or \
        sum(b - a for a, b in plan_shards(p_len[:1], t_len[:1], 4)) == 1
    assert plan_shards(p_len[:0], t_len[:0], 3) == [(0, 0)] * 3


@pytest.mark.parametrize("kw", [dict(span="end-to-end"), dict(scope="score", span="end-to-end"), dict(distance="affine2p")])
def test_merge_equals_single_batch(oracle, kw):
    batch = generate_pairs(600, 120, 0.08, seed=21)
    cfg = oracle.make_config(**kw)
    want = oracle.align_batch(cfg, *batch, kind="port")
    parts = [oracle.align_batch(cfg, *slice_batch(batch, a, b), kind="port") for a, b in plan_shards(batch[2], batch[4], 3)]
    got = merge_results(parts)
    for k in ("score", "status", "locs", "cig_off", "runs"):
        assert np.array_equal(got[k], want[k]), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from oracle import oracle_py
    from pywfa_b200.shard import align_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = generate_pairs(400, 150, 0.06, seed=5)
        for kw in (dict(span="end-to-end"), dict(scope="score", span="end-to-end")):
            cfg = oracle_py.make_config(**kw)

            def engine(c, *b):
                return oracle_py.align_batch(c, *b, kind="port")
            everywhere = align_sharded(engine, cfg, batch)                 # all_gather_object
            at_root = align_sharded(engine, cfg, batch, dst=0)             # gather_object
            want = oracle_py.align_batch(cfg, *batch, kind="port")
            ok = all(np.array_equal(everywhere[k], want[k]) for k in ("score", "status", "locs", "cig_off", "runs"))
            if rank == 0:
                ok = ok and all(np.array_equal(at_root[k], want[k]) for k in ("score", "status", "locs", "cig_off", "runs"))
            else:
                ok = ok and at_root is None
            q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_align_sharded_two_gloo_ranks(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(4)]
    assert all(ok for _, ok in res) and {r for r, _ in res} == {0, 1}


@pytest.mark.gpu
def test_multi_device_threads_on_gpu(oracle):
    """Two contexts (both on device 0 when the box has one GPU): thread-per-device driver."""
    from pywfa_b200 import _ffi
    ndev = _ffi.lib().wfagpu_device_count()
    devices = [0, 1 % max(ndev, 1)]
    batch = generate_pairs(20000, 150, 0.05, seed=13)
    for kw in (dict(span="end-to-end"), dict(scope="score", span="end-to-end")):
        cfg = oracle.make_config(**kw)
        want = oracle.align_batch(cfg, *batch, kind="port")
        got = align_multi_device(cfg, batch, devices)
        for k in ("score", "status", "locs", "cig_off", "runs"):
            assert np.array_equal(got[k], want[k]), k


def _shm_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from oracle import oracle_py
    from pywfa_b200.shard import SharedResults
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_rank = 300
        ok = True
        for kw, full in ((dict(span="end-to-end"), True), (dict(scope="score", span="end-to-end"), False)):
            cfg = oracle_py.make_config(**kw)
            # every rank owns a part of the job's batch (different divergence: different run counts per rank)
            part = generate_pairs(n_rank, 150, 0.04 + 0.06 * rank, seed=50 + rank)
            local = oracle_py.align_batch(cfg, *part, kind="port")
            cap = torch.tensor([len(local["runs"]) + 16], dtype=torch.int64)
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
            shared = SharedResults(world * n_rank, full, rank, world, tag="test-shm", runs_per_rank=int(cap.item()), pin=False)
            out = shared.slices(rank * n_rank, n_rank)
            out["score"][:] = local["score"]; out["status"][:] = local["status"]
            if full:
                out["locs"][:] = local["locs"]; out["cig_off"][:] = local["cig_off"]
                shared.run_slice()[:len(local["runs"])] = local["runs"]
            dist.barrier()
            if rank == 0:
                for r in range(world):
                    want = oracle_py.align_batch(cfg, *generate_pairs(n_rank, 150, 0.04 + 0.06 * r, seed=50 + r), kind="port")
                    got = shared.slices(r * n_rank, n_rank, rank=r)
                    ok = ok and np.array_equal(got["score"], want["score"]) and np.array_equal(got["status"], want["status"])
                    if full:
                        ok = ok and np.array_equal(got["cig_off"], want["cig_off"]) and np.array_equal(got["locs"], want["locs"])
                        ok = ok and np.array_equal(shared.run_slice(r)[:len(want["runs"])], want["runs"])
                shared.touch()
            del out
            shared.close()
            # a rank that disagrees about the layout must fail loudly, not map a short file
            bad = None
            try:
                SharedResults(world * n_rank, True, rank, world, tag="test-shm-bad", runs_per_rank=1000 + 100000 * rank, pin=False)
            except ValueError as e:
                bad = e
            ok = ok and (bad is not None if rank else bad is None)
            dist.barrier()
            if rank == 0:
                try:
                    os.unlink(f"/dev/shm/test-shm-bad-{port}")
                except OSError:
                    pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_shared_results_two_gloo_ranks(oracle):
    """The zero-copy gather: both ranks write their shard's results into one shared-memory segment; rank 0
    reads the whole job's arrays.  Per-rank run counts differ, the layout must still agree."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(2)]
    assert all(ok for _, ok in res) and {r for r, _ in res} == {0, 1}
