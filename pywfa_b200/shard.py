"""Sharding of independent read pairs across GPUs.

Pairs are independent alignment problems, so the batch is cut into contiguous, work-balanced
shards -- one per GPU -- that are aligned without any exchange and gathered on the host; there
is no collective on the data path.  Two drivers share the same planning / merging code:

* ``align_sharded``   one process per GPU (``torchrun``): every rank aligns its shard; the
  results are gathered with ``torch.distributed`` object collectives on a host (gloo) group.
* ``align_multi_device``   one process, one host thread and one ``wfagpu_ctx`` per device (the
  C library releases the GIL while it runs).
* ``SharedResults``   the zero-copy form of the host-side gather for one process per GPU: the
  job's result arrays live in ONE shared-memory segment that every rank maps and pins, so each
  rank's copy engine writes its shard's results straight into the gathered arrays.
"""
from __future__ import annotations

import threading

import numpy as np

__all__ = ["plan_shards", "slice_batch", "merge_results", "align_sharded", "align_multi_device", "SharedResults"]


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


class SharedResults:
    """Result arrays of a whole job (``n_total`` pairs over ``world`` ranks) in one shared-memory file.

    Rank 0 creates ``/dev/shm/<tag>-<MASTER_PORT>`` (score, status and -- ``full`` -- locs, cig_off),
    the other ranks map it after a barrier of ``torch.distributed``; with a visible GPU every rank pins
    the mapping (``wfagpu_host_register``), so ``Context.align_batch(..., out=shared.slices(start, n))``
    downloads by DMA into the gathered arrays and the host-side gather costs no copy.  With
    ``world == 1`` these are plain pinned (or, without a GPU, ordinary) arrays.  CIGAR runs are
    gathered by giving each rank's context its slice of ``runs`` (``Context.set_run_buffer``);
    ``cig_off`` values are relative to the owning rank's slice, whose first word is ``run_base(rank)``.
    """

    def __init__(self, n_total: int, full: bool, rank: int = 0, world: int = 1, tag: str = "wfagpu", runs_per_rank: int = 0,
                 pin: bool = True):
        import os
        self.n, self.full, self.rank, self.world = int(n_total), bool(full), rank, world
        self.runs_per_rank = int(runs_per_rank) if full else 0
        fields = [("score", np.int32, (self.n,)), ("status", np.int32, (self.n,))]
        if full:
            fields += [("cig_off", np.int64, (self.n + world,)), ("locs", np.int32, (self.n, 4))]
            if self.runs_per_rank:
                fields.append(("runs", np.uint32, (self.runs_per_rank * world,)))
        sizes = [int(np.prod(shape)) * np.dtype(dt).itemsize for _, dt, shape in fields]
        offs = np.concatenate(([0], np.cumsum([(sz + 4095) // 4096 * 4096 for sz in sizes])))
        total = int(offs[-1]) or 4096
        self._path = None
        self._registered = False
        self._map = None
        if world > 1:
            import mmap
            import torch.distributed as dist
            self._path = f"/dev/shm/{tag}-{os.environ.get('MASTER_PORT', '0')}"
            if rank == 0:
                with open(self._path, "wb") as fh:
                    fh.truncate(total)
            dist.barrier()
            if os.path.getsize(self._path) != total:
                raise ValueError(f"SharedResults: rank {rank} computed a {total}-byte layout but rank 0 created "
                                 f"{os.path.getsize(self._path)} bytes: n_total / full / runs_per_rank must agree on all ranks")
            fh = open(self._path, "r+b")
            self._map = mmap.mmap(fh.fileno(), total)
            fh.close()
            buf = np.frombuffer(self._map, np.uint8)
            if pin:
                from . import _ffi
                import ctypes as C
                if _ffi.lib().wfagpu_device_count() > 0:
                    self._addr = buf.ctypes.data
                    self._registered = _ffi.lib().wfagpu_host_register(C.c_void_p(self._addr), total) == 0
        else:
            buf = None
            if pin:
                try:
                    from . import _ffi
                    if _ffi.lib().wfagpu_device_count() > 0:
                        buf = _ffi.pinned_empty(total, np.uint8)
                except Exception:
                    buf = None
            if buf is None:
                buf = np.empty(total, np.uint8)
        self._buf = buf
        for (name, dt, shape), off, sz in zip(fields, offs[:-1], sizes):
            setattr(self, name, buf[int(off):int(off) + sz].view(dt).reshape(shape))
        if not hasattr(self, "runs"):
            self.runs = None

    def slices(self, start: int, n: int, rank=None):
        """``out=`` arrays for the shard ``[start, start + n)`` owned by ``rank`` (default: this rank)."""
        r = self.rank if rank is None else rank
        out = {"score": self.score[start:start + n], "status": self.status[start:start + n]}
        if self.full:
            out["locs"] = self.locs[start:start + n]
            out["cig_off"] = self.cig_off[start + r:start + r + n + 1]      # n + 1 entries per shard
        else:
            out["locs"] = np.empty((n, 4), np.int32)
            out["cig_off"] = np.empty(n + 1, np.int64)
        return out

    def run_slice(self, rank=None):
        r = self.rank if rank is None else rank
        return None if self.runs is None else self.runs[r * self.runs_per_rank:(r + 1) * self.runs_per_rank]

    def touch(self) -> int:
        """The consumer's side of the gather: read across all shards of the gathered arrays."""
        return int(self.status[::1024].sum()) + int(self.score[::1024].sum())

    def close(self):
        import os
        for name in ("score", "status", "cig_off", "locs", "runs"):
            if hasattr(self, name):
                setattr(self, name, None)
        if self._registered:
            from . import _ffi
            import ctypes as C
            _ffi.lib().wfagpu_host_unregister(C.c_void_p(self._addr))
            self._registered = False
        self._buf = None
        if self._map is not None:
            try:
                self._map.close()
            except BufferError:
                pass                     # a caller still holds a view; the mapping goes with it
            self._map = None
        if self._path and self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            if self.rank == 0:
                try:
                    os.unlink(self._path)
                except OSError:
                    pass
