"""Sharding of independent read pairs across GPUs.

Pairs are independent alignment problems, so the batch is cut into contiguous, work-balanced
shards -- one per GPU -- that are aligned without any exchange and gathered on the host; there
is no collective on the data path.  Two drivers share the same planning / merging code:

* ``align_sharded``   one process per GPU (``torchrun``): every rank aligns its shard; the
  results are gathered with ``torch.distributed`` object collectives on a host (gloo) group.
* ``align_multi_device``   one process, one host thread and one ``wfagpu_ctx`` per device (the
  C library releases the GIL while it runs).
"""
from __future__ import annotations

import threading

import numpy as np

__all__ = ["plan_shards", "slice_batch", "merge_results", "align_sharded", "align_multi_device"]


def plan_shards(p_len, t_len, n_shards: int):
    """Contiguous ``(start, stop)`` ranges with roughly equal work.

    The wavefront algorithm costs O(n*s + s^2) per pair (W/README.md:7 of the reference) and the
    score ``s`` grows with the length, so a pair weighs ``len + len^2 / 64`` with
    ``len = plen + tlen``; the constant only matters for mixed-length batches.
    """
    n = len(p_len)
    if n_shards <= 0:
        raise ValueError("n_shards must be positive")
    ln = np.asarray(p_len, np.float64) + np.asarray(t_len, np.float64)
    cum = np.concatenate(([0.0], np.cumsum(ln + ln * ln / 64.0 + 1.0)))
    bounds = [0]
    for i in range(1, n_shards):
        cut = int(np.searchsorted(cum, cum[-1] * i / n_shards, side="left"))
        bounds.append(min(max(cut, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(n_shards)]


def slice_batch(batch, start: int, stop: int):
    """Sub-batch ``[start, stop)`` sharing the sequence buffer (offsets stay absolute)."""
    seq, p_off, p_len, t_off, t_len = batch
    return seq, p_off[start:stop], p_len[start:stop], t_off[start:stop], t_len[start:stop]


def merge_results(parts):
    """Concatenate per-shard result dicts (score, status, locs, cig_off, runs) in shard order,
    rebasing the CIGAR offsets."""
    parts = [p for p in parts if p is not None]
    if not parts:
        raise ValueError("no shard results")
    out = {k: np.concatenate([p[k] for p in parts]) for k in ("score", "status")}
    out["locs"] = np.concatenate([np.asarray(p["locs"]).reshape(-1, 4) for p in parts])
    runs = [np.asarray(p["runs"], np.uint32) for p in parts]
    offs, base = [], 0
    for p in parts:
        co = np.asarray(p["cig_off"], np.int64)
        offs.append(co[:-1] + base)
        base += int(co[-1])
    out["cig_off"] = np.concatenate(offs + [np.array([base], np.int64)])
    out["runs"] = np.concatenate(runs) if runs else np.zeros(0, np.uint32)
    return out


def align_sharded(engine, cfg, batch, group=None, dst=None):
    """Align ``batch`` across the ranks of ``torch.distributed``.

    ``engine(cfg, seq, p_off, p_len, t_off, t_len) -> dict`` is the local aligner (for the product:
    ``pywfa_b200._ffi.Context(local_rank).align_batch``).  Every rank passes the same ``batch``
    (or at least the same lengths; only its own shard's bases are read).  The per-rank results
    are gathered as host objects on ``group`` -- pass a gloo group when the default group is NCCL
    -- to rank ``dst`` (``None``: to every rank).  Returns the merged dict (or ``None`` on ranks
    other than ``dst``).
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shards = plan_shards(batch[2], batch[4], world)
    a, b = shards[rank]
    local = engine(cfg, *slice_batch(batch, a, b))
    local = {k: np.ascontiguousarray(local[k]) for k in ("score", "status", "locs", "cig_off", "runs")}
    if dst is None:
        parts = [None] * world
        dist.all_gather_object(parts, local, group=group)
        return merge_results(parts)
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst, group=group)
    return merge_results(parts) if rank == dst else None


def align_multi_device(cfg, batch, devices, contexts=None):
    """Align ``batch`` on several GPUs of this process: one host thread + one context per device."""
    from . import _ffi
    devices = list(devices)
    if not devices:
        raise ValueError("no devices")
    own = contexts is None
    ctxs = contexts if contexts is not None else [_ffi.Context(d) for d in devices]
    shards = plan_shards(batch[2], batch[4], len(devices))
    parts, errors = [None] * len(devices), [None] * len(devices)

    def work(i):
        try:
            a, b = shards[i]
            parts[i] = ctxs[i].align_batch(cfg, *slice_batch(batch, a, b))
        except Exception as e:           # surfaced after the join
            errors[i] = e

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if own:
        for c in ctxs:
            c.close()
    for e in errors:
        if e is not None:
            raise e
    return merge_results(parts)
