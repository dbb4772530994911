#!/usr/bin/env python
"""bench.py -- aligned pairs/s of the batched wavefront aligner on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over the whole synthetic batch.  Workload at N=1 is
BASELINE.json configs[1] ("cfg2": 10 M synthetic 250 bp pairs, 10 % divergence, gap-affine
x=4 o=6 e=2, end-to-end, scope=score); with N>1 every rank aligns its own batch of that size
(independent pairs sharded across GPUs, no collective on the data path: "scaling": "weak").

  value     device-resident: packed batch already in HBM, CUDA events around K runs
  e2e       through the C ABI from HOST buffers: 2-bit pack + H2D + kernels + D2H every step
  roofline  dominant kernel (wfa_reg_kernel for the bench line) against the measured HBM peak, plus the
            integer-issue figure that actually bounds this path
  cpu_baseline  the reference (oracle/_ref: unmodified WFA2-lib + pywfa compiled in the build
            container) on all host cores, bounded sample of the same workload

`--impl reference` times only the reference arm (pywfa's public API on a process pool).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (pairs, length, divergence, flank, config kwargs)
    "cfg1": (1_000_000, 150, 0.05, 0, dict(span="end-to-end", scope="full")),
    "cfg2": (10_000_000, 250, 0.10, 0, dict(span="end-to-end", scope="score")),
    "cfg3": (100_000, 1000, 0.10, 0, dict(distance="affine2p", span="ends-free", scope="full")),
    "cfg4-adaptive": (20_000, 10_000, 0.15, 0, dict(span="end-to-end", scope="full", heuristic="adaptive")),
    "cfg4-xdrop": (20_000, 10_000, 0.15, 0, dict(span="end-to-end", scope="full", heuristic="X-drop", xdrop=20)),
    "cfg4-none": (2_000, 10_000, 0.15, 0, dict(span="end-to-end", scope="full")),
    "cfg5": (16, 100_000, 0.20, 0, dict(distance="affine2p", span="end-to-end", scope="full")),
}
# pywfa pairs/s per host core (survey measurements), used to size the bounded CPU samples
REF_PER_CORE = {"cfg1": 90_000, "cfg2": 25_000, "cfg3": 280, "cfg4-adaptive": 120, "cfg4-xdrop": 20_000, "cfg4-none": 1.5, "cfg5": 0.005}
WORKLOAD_DESC = {
    "cfg1": "1M synthetic 150 bp pairs, 5% divergence, affine (x=4,o=6,e=2), end-to-end, scope=full",
    "cfg2": "10M synthetic 250 bp pairs, 10% divergence, affine (x=4,o=6,e=2), end-to-end, scope=score",
    "cfg3": "100k (of 1M) synthetic 1 kbp pairs, 10% divergence, affine2p, ends-free, scope=full",
    "cfg4-adaptive": "20k (of 100k) synthetic 10 kbp pairs, 15% divergence, affine, heuristic=adaptive, scope=full",
    "cfg4-xdrop": "20k (of 100k) synthetic 10 kbp pairs, 15% divergence, affine, heuristic=X-drop (xdrop=20), scope=full",
    "cfg4-none": "2k (of 100k) synthetic 10 kbp pairs, 15% divergence, affine, no heuristic, scope=full",
    "cfg5": "16 (of 10k) synthetic 100 kbp pairs, 20% divergence, affine2p, end-to-end, scope=full, several CTAs per pair, history in HBM",
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- synthetic data ------------------------------------------------------------------------
def make_batch(n, length, div, flank, seed, chunk=250_000, workers=8):
    """SURVEY.md 8(d) generator, in chunks (seed + chunk index) so that 10 M pairs fit in RAM."""
    from concurrent.futures import ThreadPoolExecutor

    from pywfa_b200.synth import generate_pairs
    nchunks = (n + chunk - 1) // chunk
    sizes = [min(chunk, n - i * chunk) for i in range(nchunks)]

    def one(i):
        return generate_pairs(sizes[i], length, div, seed=seed + i, text_flank=flank)

    with ThreadPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(one, range(nchunks)))
    total = sum(len(p[0]) for p in parts)
    seq = np.empty(total, np.uint8)
    p_off = np.empty(n, np.int64); t_off = np.empty(n, np.int64)
    p_len = np.empty(n, np.int32); t_len = np.empty(n, np.int32)
    pos = 0; row = 0
    for s, po, pl, to, tl in parts:
        seq[pos:pos + len(s)] = s
        m = len(pl)
        p_off[row:row + m] = po + pos; t_off[row:row + m] = to + pos
        p_len[row:row + m] = pl; t_len[row:row + m] = tl
        pos += len(s); row += m
    return seq, p_off, p_len, t_off, t_len


# ---- clocks --------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark(self):
        """Start of the window whose samples count (the timed region)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], 0, set()
        for ts, ln in self.lines:
            if ts < getattr(self, "t_mark", 0.0):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- reference arm / CPU baseline ----------------------------------------------------------
def _pool_worker(args):
    """Align one contiguous shard through pywfa's public API (the reference's own code path)."""
    shard, kw = args
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from pywfa import WavefrontAligner  # the UNMODIFIED reference, compiled under oracle/_ref
    seq, p_off, p_len, t_off, t_len = shard
    buf = seq.tobytes().decode("ascii")
    pats = [buf[o:o + l] for o, l in zip(p_off.tolist(), p_len.tolist())]
    txts = [buf[o:o + l] for o, l in zip(t_off.tolist(), t_len.tolist())]
    a = WavefrontAligner(**kw)
    chk = 0
    t0 = time.perf_counter()
    for p, t in zip(pats, txts):
        r = a(t, p)
        chk += r.score
    return len(pats), time.perf_counter() - t0, chk


def _shards(batch, nshards):
    seq, p_off, p_len, t_off, t_len = batch
    n = len(p_len)
    out = []
    for i in range(nshards):
        a, b = n * i // nshards, n * (i + 1) // nshards
        if a == b:
            continue
        lo = int(min(p_off[a], t_off[a])); hi = int(max(p_off[b - 1] + p_len[b - 1], t_off[b - 1] + t_len[b - 1]))
        out.append((seq[lo:hi].copy(), p_off[a:b] - lo, p_len[a:b].copy(), t_off[a:b] - lo, t_len[a:b].copy()))
    return out


def pywfa_kwargs(kw):
    d = dict(distance="affine", mismatch=4, gap_opening=6, gap_extension=2)
    d.update(kw)
    return d


def reference_pool_rate(batch, kw, cores):
    """pairs/s of pywfa (reference) on a multiprocessing pool over `cores` processes."""
    import multiprocessing as mp
    shards = _shards(batch, cores)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_pool_worker, [(s, pywfa_kwargs(kw)) for s in shards])
        wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return n / slowest, n, slowest, wall


def reference_c_rate(batch, cfg, cores):
    """pairs/s of the reference C library (no Python per-pair overhead), `cores` threads."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle_py
    shards = _shards(batch, cores)

    def run(s):
        t0 = time.perf_counter()
        oracle_py.align_batch(cfg, *s, kind="reference")
        return time.perf_counter() - t0
    with ThreadPoolExecutor(cores) as ex:
        times = list(ex.map(run, shards))
    return len(batch[1]) / max(times)


def have_reference():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libwfa_ref.so")) and \
        os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "pywfa"))


# ---- main ----------------------------------------------------------------------------------
def algorithmic_figures(stats, n, length_p, length_t_mean, two_p, full, runs_total=0):
    """HBM bytes and integer ops of one pass over the batch."""
    cells = stats["cells"]
    ops_cell = 33 if two_p else 19
    extend = 4 * (cells + n * length_p / 16.0)
    int_ops = cells * ops_cell + extend
    hbm = n * (np.ceil(length_p / 4) + np.ceil(length_t_mean / 4) + 32) + 4 * runs_total
    if full:
        hbm += cells               # one origin byte per cell, written once (the backtrace reads only the path)
    return float(hbm), float(int_ops)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="override the number of pairs per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_pairs, length, div, flank, kw = WORKLOADS[args.workload]
    if args.pairs:
        n_pairs = args.pairs
    cores = os.cpu_count() or 1
    two_p = kw.get("distance") == "affine2p"
    full = kw.get("scope", "full") == "full"
    config = {"workload": WORKLOAD_DESC[args.workload] + (f" [pairs per GPU overridden: {n_pairs}]" if args.pairs else ""),
              "pairs_per_gpu": n_pairs, "sharding": f"{world} x independent batches, host-side gather, no collective",
              "l2_policy": "inputs larger than L2 (packed batch > 126 MB)" if n_pairs * length / 2 > 126e6
              else "batch smaller than L2; results are compute-bound, no flush"}

    # ---------------- reference arm ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        if not have_reference():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference build) missing"}))
            return
        # bounded sample per step: about 10 s of wall-clock on the host cores
        per_core = REF_PER_CORE[args.workload]
        sample = int(min(n_pairs, max(cores * (64 if per_core > 10 else 1), per_core * cores * 6)))
        batch = make_batch(sample, length, div, flank, seed=1234)
        rates = []
        for i in range(args.warmup + args.steps):
            r, n, slow, wall = reference_pool_rate(batch, kw, cores)
            log(f"[reference] step {i}: {n} pairs, slowest worker {slow:.2f}s, pool wall {wall:.2f}s -> {r:,.0f} pairs/s")
            if i >= args.warmup:
                rates.append(r)
        v = float(np.mean(rates))
        line = {"impl": "reference", "metric": "aligned pairs/sec", "value": v, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                 "sample": f"{sample} pairs of the workload per step, pywfa a(text, pattern) on multiprocessing.Pool({cores})"},
                "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- native arm ----------------
    import torch
    import torch.distributed as dist

    from oracle import oracle_py  # cpu_baseline leg + config struct helper only
    from pywfa_b200 import _ffi
    from pywfa_b200.build import build_library
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pywfa_b200 has no CPU fallback")
    build_library()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = oracle_py.make_config(**kw)

    t0 = time.perf_counter()
    batch = make_batch(n_pairs, length, div, flank, seed=1234 + 1000 * rank, workers=max(2, cores // max(world, 1)))
    log(f"[rank {rank}] generated {n_pairs} pairs in {time.perf_counter() - t0:.1f}s ({len(batch[0]) / 1e9:.2f} GB ASCII)")

    ctx = _ffi.Context(local_rank)
    t0 = time.perf_counter()
    b = ctx.prepare(cfg, *batch)
    log(f"[rank {rank}] prepare (pack + H2D): {time.perf_counter() - t0:.2f}s")
    stream = torch.cuda.current_stream().cuda_stream

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # nvidia-smi needs ~0.2 s before its first line: start it early
    for _ in range(max(args.warmup, 3)):
        b.run(stream)
    sync_all()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        b.run(stream)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = None
    if rank == 0:
        if not any(ts >= sampler.t_mark for ts, _ in sampler.lines):
            # timed region shorter than one nvidia-smi period: sample an identical untimed replay
            t_end = time.perf_counter() + 0.6
            while time.perf_counter() < t_end:
                b.run(stream)
                torch.cuda.synchronize()
            clocks = sampler.stop()
            clocks["note"] = "timed region < 50 ms: clocks sampled over an identical untimed replay right after it"
        else:
            clocks = sampler.stop()
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        dist.barrier()
    stats = b.stats()
    res = b.fetch(cigars=False)
    status_hist = {int(k): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))}
    ms_per_step = ms / args.steps
    value = world * n_pairs / (ms_per_step / 1e3)
    log(f"[rank {rank}] device-resident: {ms_per_step:.1f} ms/step, {n_pairs / (ms_per_step / 1e3):,.0f} pairs/s/GPU, "
        f"launches/step {stats['kernel_launches']}, retried {stats['retried_pairs']}, status {status_hist}")

    # ---- e2e: host buffers -> results, every step ----
    e2e = None
    if not args.no_e2e:
        # host result arrays are allocated once and refilled every step (host buffers, not pinned)
        outs = dict(score=np.empty(n_pairs, np.int32), status=np.empty(n_pairs, np.int32),
                    locs=np.empty((n_pairs, 4), np.int32), cig_off=np.empty(n_pairs + 1, np.int64))
        for _ in range(2):
            ctx.align_batch(cfg, *batch, copy_runs=False, check=False, out=outs)   # warm staging / buffer pools
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = ctx.align_batch(cfg, *batch, copy_runs=False, check=False, out=outs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h_e2e = 8 * n_pairs + (16 * n_pairs + 8 * (n_pairs + 1) + 4 * int(r["cig_off"][-1]) if full else 0)
        e2e_launches = ctx.last_launches()
        if world > 1:
            tdt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        e2e = {"value": world * n_pairs * args.steps / dt, "unit": "pairs/s",
               "h2d_bytes_per_step": stats["h2d_bytes"], "d2h_bytes_per_step": d2h_e2e, "gpu_launches_per_step": e2e_launches,
               "ms_per_step": 1e3 * dt / args.steps, "timing": "wall clock around wfagpu_align_batch (host pack + H2D + kernels + D2H), max over ranks"}
        log(f"[rank {rank}] e2e: {1e3 * dt / args.steps:.1f} ms/step -> {e2e['value']:,.0f} pairs/s")
    b.free()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    hbm_bytes, int_ops = algorithmic_figures(stats, n_pairs, length, float(np.mean(batch[4])), two_p, full)
    kernel_s = ms_per_step / 1e3            # the register-tier kernels are >99% of the step (see profiles/)
    achieved = hbm_bytes / kernel_s / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    int_peak = 148 * 128 * sm_mhz * 1e6     # INT32 lane-ops/s at the clock seen under load
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch, from the committed
    # `ncu --set full` capture of this exact command (profiles/r01_reg_cfg2_10M_ncu_summary.txt)
    traffic = 1.483855e9 + 87.523584e6 if (args.workload == "cfg2" and n_pairs == 10_000_000) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "note": "integer-DP kernel: HBM is not the binding resource (SURVEY.md 8(d)); see int_issue",
                "int_issue": {"achieved_gops": int_ops / kernel_s / 1e9, "peak_gops": int_peak / 1e9,
                              "frac": int_ops / kernel_s / int_peak, "cells_per_step": stats["cells"],
                              "unit": "INT32 lane-ops/s (19|33 per cell + extend, SURVEY.md 8(d))"}}

    # arithmetic type of the offsets the dominant tier computes in (scores are int32 everywhere)
    dtype = "int32" if length > 12_000 else "int16"
    # ---- CPU baseline: the reference on the host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline and args.workload == "cfg5" and not args.cpu_sample:
        cpu = {"value": None, "unit": "pairs/s", "cores": cores, "kind": "reference",
               "sample": "not timed by default: one 100 kbp pair takes the reference minutes per core (BASELINE.md); pass --cpu-sample N"}
    elif not args.no_cpu_baseline:
        if have_reference():
            per_core = REF_PER_CORE[args.workload]
            sample = args.cpu_sample or int(min(n_pairs, max(cores, per_core * cores * 2)))
            sub = tuple(a[:sample] if i else a for i, a in enumerate(batch))
            r, n, slow, wall = reference_pool_rate(sub, kw, cores)
            c_rate = reference_c_rate(sub, cfg, cores)
            cpu = {"value": r, "unit": "pairs/s", "cores": cores, "kind": "reference",
                   "sample": f"first {sample} pairs of the workload, pywfa a(text, pattern) on multiprocessing.Pool({cores}); "
                             f"slowest worker {slow:.1f}s", "c_library_rate": c_rate}
        else:
            cpu = {"value": None, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"}

    line = {"metric": "aligned pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(stats["kernel_launches"]) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
            "parity": {"status_histogram": status_hist, "retried_pairs": stats["retried_pairs"]}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
