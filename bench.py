#!/usr/bin/env python
"""bench.py -- aligned pairs/s of the batched wavefront aligner on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over the whole synthetic batch.  Workload at N=1 is
BASELINE.json configs[1] ("cfg2": 10 M synthetic 250 bp pairs, 10 % divergence, gap-affine
x=4 o=6 e=2, end-to-end, scope=score); with N>1 the job's batch is N such parts, one per GPU
(independent pairs sharded across GPUs, no collective on the data path: "scaling": "weak"), and the
results are gathered on the host into ONE pair of arrays (shared memory that every rank's copy engine
writes directly: pywfa_b200.shard.SharedResults).

  value     device-resident: packed batch already in HBM, CUDA events around K runs
  e2e       through the C ABI (wfagpu_align_batch) from HOST buffers in pinned memory: H2D of the raw
            ASCII bases + device-side packing + kernels + D2H into the (gathered) result arrays, every step
  roofline  dominant kernel (wfa_reg_kernel for the bench line) against the resource that binds it (warp
            issue / INT pipe), with the HBM figures beside it
  cpu_baseline  the reference (oracle/_ref: unmodified WFA2-lib + pywfa compiled in the build
            container) on all host cores, bounded sample of the same workload
  parity    GPU results against the reference's on the whole cpu_baseline sample (scores; CIGAR runs for
            scope=full), mismatches counted
  secondary the same measurements for BASELINE.json configs[0] ("cfg1": 1 M x 150 bp, scope=full), the
            config the >= 50x target is stated on
  single_pair_latency_us   a(text, pattern) through the pywfa-compatible surface vs pywfa itself

`--impl reference` times only the reference arm (pywfa's public API on a process pool).
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import json
import os
import re
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
    # length buckets: 150 bp + 250 bp + 1 kbp pairs shuffled into one batch (see --workload mixed)
    "mixed": (0, 0, 0.0, 0, dict(span="end-to-end", scope="score")),
    # the M-only metrics: score-only runs on the gap-affine tiers, CIGARs on the scalar tiers
    "edit-score": (4_000_000, 150, 0.05, 0, dict(distance="levenshtein", span="end-to-end", scope="score")),
    "edit-full": (400_000, 150, 0.05, 0, dict(distance="levenshtein", span="end-to-end", scope="full")),
    "linear-score": (4_000_000, 250, 0.10, 0, dict(distance="linear", span="end-to-end", scope="score")),
}
MIXED_PARTS = [(600_000, 150, 0.05), (300_000, 250, 0.10), (20_000, 1000, 0.05)]
# pywfa pairs/s per host core (survey measurements), used to size the bounded CPU samples
REF_PER_CORE = {"cfg1": 90_000, "cfg2": 25_000, "cfg3": 280, "cfg4-adaptive": 120, "cfg4-xdrop": 20_000, "cfg4-none": 1.5,
                "cfg5": 0.005, "mixed": 40_000, "edit-score": 250_000, "edit-full": 150_000, "linear-score": 60_000}
WORKLOAD_DESC = {
    "cfg1": "1M synthetic 150 bp pairs, 5% divergence, affine (x=4,o=6,e=2), end-to-end, scope=full",
    "cfg2": "10M synthetic 250 bp pairs, 10% divergence, affine (x=4,o=6,e=2), end-to-end, scope=score",
    "cfg3": "100k (of 1M) synthetic 1 kbp pairs, 10% divergence, affine2p, ends-free, scope=full",
    "cfg4-adaptive": "20k (of 100k) synthetic 10 kbp pairs, 15% divergence, affine, heuristic=adaptive, scope=full",
    "cfg4-xdrop": "20k (of 100k) synthetic 10 kbp pairs, 15% divergence, affine, heuristic=X-drop (xdrop=20), scope=full",
    "cfg4-none": "2k (of 100k) synthetic 10 kbp pairs, 15% divergence, affine, no heuristic, scope=full",
    "cfg5": "16 (of 10k) synthetic 100 kbp pairs, 20% divergence, affine2p, end-to-end, scope=full, several CTAs per pair, history in HBM",
    "mixed": "600k x 150 bp 5% + 300k x 250 bp 10% + 20k x 1 kbp 5% pairs shuffled into one batch, affine, end-to-end, scope=score (length buckets)",
    "edit-score": "4M synthetic 150 bp pairs, 5% divergence, levenshtein, end-to-end, scope=score",
    "edit-full": "400k synthetic 150 bp pairs, 5% divergence, levenshtein, end-to-end, scope=full",
    "linear-score": "4M synthetic 250 bp pairs, 10% divergence, gap-linear (x=4, indel=2), end-to-end, scope=score",
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- synthetic data ------------------------------------------------------------------------
def make_batch(n, length, div, flank, seed, chunk=250_000, workers=8, alloc=None):
    """SURVEY.md 8(d) generator, in chunks (seed + chunk index) so that 10 M pairs fit in RAM.
    `alloc(shape, dtype)` provides the arrays (pinned host memory for the e2e measurement)."""
    from concurrent.futures import ThreadPoolExecutor

    from pywfa_b200.synth import generate_pairs
    alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))
    nchunks = (n + chunk - 1) // chunk
    sizes = [min(chunk, n - i * chunk) for i in range(nchunks)]

    def one(i):
        return generate_pairs(sizes[i], length, div, seed=seed + i, text_flank=flank)

    with ThreadPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(one, range(nchunks)))
    total = sum(len(p[0]) for p in parts)
    seq = alloc(total + 64, np.uint8)
    p_off = alloc(n, np.int64); t_off = alloc(n, np.int64)
    p_len = alloc(n, np.int32); t_len = alloc(n, np.int32)
    pos = 0; row = 0
    for s, po, pl, to, tl in parts:
        seq[pos:pos + len(s)] = s
        m = len(pl)
        p_off[row:row + m] = po + pos; t_off[row:row + m] = to + pos
        p_len[row:row + m] = pl; t_len[row:row + m] = tl
        pos += len(s); row += m
    seq[pos:] = 0
    return seq, p_off, p_len, t_off, t_len


def make_mixed(seed, alloc=None):
    """150 bp + 250 bp + 1 kbp pairs, shuffled (the caller of a real pipeline does not sort by length)."""
    alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))
    parts = [make_batch(n, length, div, 0, seed + 100 * i) for i, (n, length, div) in enumerate(MIXED_PARTS)]
    n = sum(len(p[1]) for p in parts)
    rng = np.random.default_rng(seed)
    order = rng.permutation(n)
    base = np.cumsum([0] + [len(p[0]) for p in parts])
    seq = alloc(int(base[-1]) + 64, np.uint8)
    for b, p in zip(base, parts):
        seq[b:b + len(p[0])] = p[0]
    cat = lambda k, shift: np.concatenate([p[k] + (b if shift else 0) for b, p in zip(base, parts)])[order]
    out = []
    for k, shift, dt in ((1, True, np.int64), (2, False, np.int32), (3, True, np.int64), (4, False, np.int32)):
        a = alloc(n, dt); a[:] = cat(k, shift); out.append(a)
    return (seq, *out)


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
    """Align one contiguous shard through pywfa's public API (the reference's own code path).
    Returns the per-pair scores and, for scope=full, the CIGAR runs (checked against the GPU's by the
    caller, outside the timed loop)."""
    shard, kw, want_cigars = args
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from pywfa import WavefrontAligner  # the UNMODIFIED reference, compiled under oracle/_ref
    seq, p_off, p_len, t_off, t_len = shard
    buf = seq.tobytes().decode("ascii")
    pats = [buf[o:o + l] for o, l in zip(p_off.tolist(), p_len.tolist())]
    txts = [buf[o:o + l] for o, l in zip(t_off.tolist(), t_len.tolist())]
    a = WavefrontAligner(**kw)
    scores = np.empty(len(pats), np.int32)
    status = np.empty(len(pats), np.int32)
    cigs = []
    t0 = time.perf_counter()
    for i, (p, t) in enumerate(zip(pats, txts)):
        r = a(t, p)
        scores[i] = r.score
        status[i] = r.status
        if want_cigars:
            cigs.append(r.cigartuples)
    dt = time.perf_counter() - t0
    runs = cig_n = None
    if want_cigars:
        cig_n = np.fromiter((len(c) for c in cigs), np.int64, len(cigs))
        runs = np.fromiter(((ln << 4) | op for c in cigs for op, ln in c), np.uint32, int(cig_n.sum()))
    return len(pats), dt, scores, status, cig_n, runs


def _shards(batch, nshards):
    seq, p_off, p_len, t_off, t_len = batch
    n = len(p_len)
    out = []
    for i in range(nshards):
        a, b = n * i // nshards, n * (i + 1) // nshards
        if a == b:
            continue
        lo = int(min(p_off[a:b].min(), t_off[a:b].min())); hi = int(max((p_off[a:b] + p_len[a:b]).max(), (t_off[a:b] + t_len[a:b]).max()))
        out.append((np.array(seq[lo:hi]), p_off[a:b] - lo, np.array(p_len[a:b]), t_off[a:b] - lo, np.array(t_len[a:b])))
    return out


def pywfa_kwargs(kw):
    d = dict(distance="affine", mismatch=4, gap_opening=6, gap_extension=2)
    d.update(kw)
    return d


def reference_pool_rate(batch, kw, cores, want_cigars=False):
    """pairs/s of pywfa (reference) on a multiprocessing pool over `cores` processes, plus its results."""
    import multiprocessing as mp
    shards = _shards(batch, cores)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_pool_worker, [(s, pywfa_kwargs(kw), want_cigars) for s in shards])
        wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    results = {"score": np.concatenate([r[2] for r in res]), "status": np.concatenate([r[3] for r in res])}
    if want_cigars:
        results["cig_n"] = np.concatenate([r[4] for r in res]); results["runs"] = np.concatenate([r[5] for r in res])
    return n / slowest, n, slowest, wall, results


def reference_c_rate(batch, cfg, cores):
    """pairs/s of the reference C library (no Python per-pair overhead), `cores` threads."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle_py
    shards = _shards(batch, cores)
    ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))

    def run(s):
        t0 = time.perf_counter()
        oracle_py.align_batch(ocfg, *s, kind="reference")
        return time.perf_counter() - t0
    with ThreadPoolExecutor(cores) as ex:
        times = list(ex.map(run, shards))
    return len(batch[1]) / max(times)


def have_reference():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libwfa_ref.so")) and \
        os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "pywfa"))


def parity_block(gpu, ref, full):
    """GPU results against the reference's on the same pairs: every score and status, and for
    scope=full every CIGAR run (plus a SHA-256 of the run words of both sides)."""
    n = len(ref["score"])
    bad = (gpu["score"][:n] != ref["score"]) | (gpu["status"][:n] != ref["status"])
    out = {"checked_pairs": int(n), "mismatches": int(bad.sum()), "against": "pywfa (oracle/_ref) on the cpu_baseline sample"}
    if full:
        off = gpu["cig_off"]
        g_n = np.diff(off[:n + 1])
        g_runs = np.ascontiguousarray(gpu["runs"][int(off[0]):int(off[n])], np.uint32)
        same_counts = np.array_equal(g_n, ref["cig_n"])
        cig_bad = 0
        if not same_counts or not np.array_equal(g_runs, ref["runs"]):
            r_off = np.concatenate(([0], np.cumsum(ref["cig_n"])))
            for i in range(n):
                if not np.array_equal(gpu["runs"][off[i]:off[i + 1]], ref["runs"][r_off[i]:r_off[i + 1]]):
                    cig_bad += 1
        out.update({"cigar_mismatches": cig_bad, "cigar_runs_checked": int(len(ref["runs"])),
                    "cigar_sha256_gpu": hashlib.sha256(g_runs.tobytes()).hexdigest(),
                    "cigar_sha256_reference": hashlib.sha256(np.ascontiguousarray(ref["runs"], np.uint32).tobytes()).hexdigest()})
        out["mismatches"] += cig_bad
    return out


# ---- roofline helpers ------------------------------------------------------------------------
def algorithmic_figures(cells, n, length_p, length_t_mean, two_p, full, runs_total=0):
    """HBM bytes and integer ops of one pass over the batch."""
    ops_cell = 33 if two_p else 19
    extend = 4 * (cells + n * length_p / 16.0)
    int_ops = cells * ops_cell + extend
    hbm = n * (np.ceil(length_p / 4) + np.ceil(length_t_mean / 4) + 32) + 4 * runs_total
    if full:
        hbm += cells               # one origin byte per cell, written once (the backtrace reads only the path)
    return float(hbm), float(int_ops)


def measured_traffic(workload, n_pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch, from the newest
    committed `ncu --set full` summary of this workload and size (profiles/rNN_reg_<workload>_<size>_ncu_summary.txt)."""
    size = f"{n_pairs // 1_000_000}M" if n_pairs % 1_000_000 == 0 else str(n_pairs)
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_reg_{workload}_{size}_ncu_summary.txt"))):
        txt = open(path).read()
        vals = {}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(re.escape(key) + r"\s+([0-9.]+)\s+([KMG]?)byte", txt)
            if m:
                vals[key] = float(m.group(1)) * {"": 1, "K": 1e3, "M": 1e6, "G": 1e9}[m.group(2)]
        if len(vals) == 2:
            best = (sum(vals.values()), os.path.relpath(path, ROOT))
    return best or (None, None)


def roofline_block(workload, n_pairs, cells, kernel_s, length, t_mean, two_p, full, clocks, peaks):
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    hbm_bytes, int_ops = algorithmic_figures(cells, n_pairs, length, t_mean, two_p, full)
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    int_peak = 148 * 128 * sm_mhz * 1e6     # INT32 lane-ops/s at the clock seen under load
    traffic, traffic_src = measured_traffic(workload, n_pairs)
    # the path is integer DP: warp issue / the INT pipe binds it, not HBM; report the
    # binding resource at top level and the HBM figures (the contract's byte roofline) beside it
    return {"bound": "int_issue", "achieved": int_ops / kernel_s / 1e9, "peak": int_peak / 1e9, "unit": "Gop/s",
            "frac": int_ops / kernel_s / int_peak,
            "algorithmic_unit": "INT32 lane-ops: 19 (33 for 2p) per wavefront cell + extend (SURVEY.md 8(d))",
            "cells_per_step": int(cells), "peak_source": f"148 SMs x 128 INT32 lanes x {sm_mhz:.0f} MHz (SM clock sampled under load)",
            "traffic": traffic, "traffic_source": traffic_src,
            "hbm": {"bound": "hbm", "achieved": hbm_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_bytes / kernel_s / 1e9 / hbm_peak, "algorithmic_bytes_per_step": hbm_bytes,
                    "traffic": traffic, "peak_source": peak_src}}


# ---- one workload on this rank ------------------------------------------------------------------
class Env:
    pass


def native_config(kw):
    """wfagpu_config_t through the product's own pywfa-compatible constructor."""
    import pywfa_b200
    return pywfa_b200.WavefrontAligner(**pywfa_kwargs(kw))._cfg


def run_workload(env, name, n_pairs, steps, warmup, want_e2e=True, sample_clocks=True):
    """Device-resident rate, e2e rate and the GPU results of one workload on this rank."""
    import torch
    import torch.distributed as dist

    from pywfa_b200 import _ffi
    from pywfa_b200.shard import SharedResults
    _, length, div, flank, kw = WORKLOADS[name]
    full = kw.get("scope", "full") == "full"
    cfg = native_config(kw)
    ctx = env.ctx
    t0 = time.perf_counter()
    if name == "mixed":
        batch = make_mixed(1234 + 1000 * env.rank, alloc=_ffi.pinned_empty)
        n_pairs = len(batch[1])
    else:
        batch = make_batch(n_pairs, length, div, flank, seed=1234 + 1000 * env.rank,
                           workers=max(2, env.cores // max(env.world, 1)), alloc=_ffi.pinned_empty)
    log(f"[rank {env.rank}] {name}: generated {n_pairs} pairs in {time.perf_counter() - t0:.1f}s ({len(batch[0]) / 1e9:.2f} GB ASCII, pinned)")
    t0 = time.perf_counter()
    b = ctx.prepare(cfg, *batch)
    log(f"[rank {env.rank}] {name}: prepare (H2D + device pack): {time.perf_counter() - t0:.2f}s")
    stream = torch.cuda.current_stream().cuda_stream

    def sync_all():
        torch.cuda.synchronize()
        if env.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(env.local_rank) if (sample_clocks and env.rank == 0) else None
    if sampler:
        sampler.start()              # nvidia-smi needs ~0.2 s before its first line: start it early
    for _ in range(max(warmup, 3)):
        b.run(stream)
    sync_all()
    if sampler:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        b.run(stream)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = None
    if sampler:
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
    if env.world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        dist.barrier()
    stats = b.stats()
    res = b.fetch(cigars=full)
    b.free()
    ms_per_step = ms / steps
    out = dict(name=name, n_pairs=n_pairs, kw=kw, full=full, two_p=kw.get("distance") == "affine2p", length=length,
               t_mean=float(np.mean(batch[4])), p_mean=float(np.mean(batch[2])), ms_per_step=ms_per_step,
               value=env.world * n_pairs / (ms_per_step / 1e3),
               stats=stats, clocks=clocks, cfg=cfg,
               status_hist={int(k): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))})
    log(f"[rank {env.rank}] {name}: device-resident {ms_per_step:.2f} ms/step, {n_pairs / (ms_per_step / 1e3):,.0f} pairs/s/GPU, "
        f"launches/step {stats['kernel_launches']}, retried {stats['retried_pairs']}, status {out['status_hist']}")

    # ---- e2e: pinned host buffers -> gathered host results, every step ----
    out["e2e"] = None
    if want_e2e:
        # the job's result arrays: ONE set for all ranks, in shared memory that every rank's copy engine
        # writes its slice of directly (host-side gather without a copy; N = 1: plain pinned arrays)
        runs_cap = int(int(res["cig_off"][-1]) * 1.1) + 4096 if full else 0
        if env.world > 1 and full:
            # every rank maps the same segment: the per-rank slice of the run array is the largest any rank needs
            cap = torch.tensor([runs_cap], device="cuda", dtype=torch.int64)
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
            runs_cap = int(cap.item())
        shared = SharedResults(env.world * n_pairs, full, env.rank, env.world, tag=f"bench-{name}", runs_per_rank=runs_cap)
        outs = shared.slices(env.rank * n_pairs, n_pairs)
        if full:
            ctx.set_run_buffer(shared.run_slice())      # this rank's CIGAR runs land in the gathered array too
        for _ in range(2):
            ctx.align_batch(cfg, *batch, copy_runs=False, check=False, out=outs)   # warm buffer pools
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = ctx.align_batch(cfg, *batch, copy_runs=False, check=False, out=outs)
            if env.world > 1:
                dist.barrier()                 # every shard of the step has landed in the gathered arrays
            if env.rank == 0:
                shared.touch()                 # rank 0 reads the gathered status array (the consumer of the gather)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h_e2e = 8 * n_pairs + (16 * n_pairs + 8 * (n_pairs + 1) + 4 * int(r["cig_off"][-1]) if full else 0)
        e2e_launches = ctx.last_launches()
        if env.world > 1:
            tdt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        same = bool(np.array_equal(outs["score"], res["score"]) and np.array_equal(outs["status"], res["status"]))
        if full:
            same = same and bool(np.array_equal(outs["cig_off"], res["cig_off"]) and np.array_equal(r["runs"], res["runs"]))
            ctx.set_run_buffer(None)
        out["e2e"] = {"value": env.world * n_pairs * steps / dt, "unit": "pairs/s",
                      "h2d_bytes_per_step": int(len(batch[0]) - 64 + 24 * n_pairs), "d2h_bytes_per_step": int(d2h_e2e),
                      "gpu_launches_per_step": int(e2e_launches), "ms_per_step": 1e3 * dt / steps,
                      "equals_device_resident_results": same,
                      "timing": "wall clock around wfagpu_align_batch: raw ASCII in pinned host memory -> H2D -> device-side packing -> "
                                "kernels -> D2H into the job's gathered result arrays" +
                                (" (shared memory written by every rank's copy engine) + barrier" if env.world > 1 else "") + ", max over ranks"}
        log(f"[rank {env.rank}] {name}: e2e {1e3 * dt / steps:.2f} ms/step -> {out['e2e']['value']:,.0f} pairs/s")
        del outs, r
        shared.close()
    # keep only what the CPU baseline / parity leg needs: the first pairs of the batch and their results
    k = min(n_pairs, sample_size(name, n_pairs, env.cores, env.cpu_sample)) if env.rank == 0 else 0
    hi = int(max((batch[1][:k] + batch[2][:k]).max(), (batch[3][:k] + batch[4][:k]).max())) if k else 0
    out["batch"] = (np.array(batch[0][:hi]),) + tuple(np.array(a[:k]) for a in batch[1:])
    keep = {"score": np.array(res["score"][:k]), "status": np.array(res["status"][:k])}
    if full:
        keep["cig_off"] = np.array(res["cig_off"][:k + 1]); keep["runs"] = np.array(res["runs"][:int(res["cig_off"][k])])
    out["results"] = keep
    return out


def sample_size(name, n_pairs, cores, override=0):
    """Pairs in the bounded CPU sample: ~10-20 s of pywfa on all host cores."""
    per_core = REF_PER_CORE[name]
    return override or int(min(n_pairs, max(cores, per_core * cores * 2)))


def mixed_parts(env, steps):
    """Length bucketing: the shuffled mixed batch against its parts aligned one by one (device-resident)."""
    import torch
    kw = WORKLOADS["mixed"][4]
    cfg = native_config(kw)
    stream = torch.cuda.current_stream().cuda_stream
    parts = []
    for i, (n, length, div) in enumerate(MIXED_PARTS):
        batch = make_batch(n, length, div, 0, 1234 + 1000 * env.rank + 100 * i)
        b = env.ctx.prepare(cfg, *batch)
        for _ in range(3):
            b.run(stream)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            b.run(stream)
        ev1.record()
        torch.cuda.synchronize()
        parts.append({"pairs": n, "length": length, "divergence": div, "ms_per_step": ev0.elapsed_time(ev1) / steps})
        b.free()
    return parts


def cpu_and_parity(env, w, c_rate=True):
    """The reference on all host cores over a bounded sample of the workload + parity of the GPU's results on it."""
    name, n_pairs, kw, full = w["name"], w["n_pairs"], w["kw"], w["full"]
    cores = env.cores
    if not have_reference():
        return ({"value": None, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"},
                {"checked_pairs": 0, "mismatches": None, "note": "oracle/_ref missing"})
    sub = w["batch"]
    sample = len(sub[1])
    r, n, slow, wall, ref = reference_pool_rate(sub, kw, cores, want_cigars=full)
    cpu = {"value": r, "unit": "pairs/s", "cores": cores, "kind": "reference",
           "sample": f"first {sample} pairs of the workload, pywfa a(text, pattern) on multiprocessing.Pool({cores}); slowest worker {slow:.1f}s"}
    if c_rate:
        cpu["c_library_rate"] = reference_c_rate(sub, w["cfg"], cores)
    par = parity_block(w["results"], ref, full)
    par["status_histogram"] = w["status_hist"]
    par["retried_pairs"] = w["stats"]["retried_pairs"]
    return cpu, par


def single_pair_latency(n_calls=300):
    """us per a(text, pattern) call: the pywfa-compatible surface of this repo (a batch of one through the
    GPU path) against pywfa itself, same 150 bp pairs."""
    import pywfa_b200
    from pywfa_b200.synth import generate_pairs
    seq, po, pl, to, tl = generate_pairs(n_calls, 150, 0.05, seed=7)
    buf = seq.tobytes().decode()
    pairs = [(buf[a:a + l], buf[b:b + m]) for a, l, b, m in zip(po, pl, to, tl)]
    out = {"pairs": n_calls, "shape": "150 bp, 5% divergence, affine, end-to-end, scope=full"}
    a = pywfa_b200.WavefrontAligner(**pywfa_kwargs(dict(span="end-to-end")))
    for p, t in pairs[:20]:
        a(t, p)
    t0 = time.perf_counter()
    for p, t in pairs:
        a(t, p)
    out["native"] = 1e6 * (time.perf_counter() - t0) / n_calls
    if have_reference():
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        from pywfa import WavefrontAligner as Ref
        ra = Ref(**pywfa_kwargs(dict(span="end-to-end")))
        t0 = time.perf_counter()
        for p, t in pairs:
            ra(t, p)
        out["reference"] = 1e6 * (time.perf_counter() - t0) / n_calls
    return out


# ---- main ----------------------------------------------------------------------------------
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg1 block and the single-pair latency")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_pairs, length, div, flank, kw = WORKLOADS[args.workload]
    if args.workload == "mixed":
        n_pairs = sum(p[0] for p in MIXED_PARTS)
    if args.pairs and args.workload != "mixed":
        n_pairs = args.pairs
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    config = {"workload": WORKLOAD_DESC[args.workload] + (f" [pairs per GPU overridden: {n_pairs}]" if args.pairs else ""),
              "pairs_per_gpu": n_pairs,
              "sharding": f"{world} x independent parts of the job's batch, one per GPU; results gathered on the host into one set of arrays"
                          " (shared memory written by DMA); no collective on the data path",
              "l2_policy": "inputs larger than L2 (packed batch > 126 MB)" if n_pairs * max(length, 150) / 2 > 126e6
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
        if args.workload == "mixed":
            batch = tuple(a[:sample] if i else a for i, a in enumerate(make_mixed(1234)))
        else:
            batch = make_batch(sample, length, div, flank, seed=1234)
        rates = []
        for i in range(args.warmup + args.steps):
            r, n, slow, wall, _ = reference_pool_rate(batch, kw, cores)
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

    from pywfa_b200 import _ffi
    from pywfa_b200.build import build_library
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pywfa_b200 has no CPU fallback")
    build_library()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = Env()
    env.rank, env.world, env.local_rank, env.cores = rank, world, local_rank, cores
    env.ctx = _ffi.Context(local_rank)
    env.cpu_sample = args.cpu_sample

    w = run_workload(env, args.workload, n_pairs, args.steps, args.warmup, want_e2e=not args.no_e2e)
    n_pairs = w["n_pairs"]
    # the config the >= 50x target is stated on, in the same run (every rank takes part: the e2e leg has barriers)
    sec = None
    if not args.no_secondary and args.workload == "cfg2":
        sec = run_workload(env, "cfg1", WORKLOADS["cfg1"][0], max(args.steps, 5), args.warmup, want_e2e=not args.no_e2e, sample_clocks=False)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # the alignment kernels are > 99 % of a device-resident step (profiles/): step time = kernel time
    roofline = roofline_block(args.workload, n_pairs, w["stats"]["cells"], w["ms_per_step"] / 1e3, w["p_mean"], w["t_mean"],
                              w["two_p"], w["full"], w["clocks"], peaks)
    # arithmetic type of the offsets the dominant tier computes in (scores are int32 everywhere)
    dtype = "int32" if length > 12_000 else "int16"

    cpu = par = None
    if not args.no_cpu_baseline and args.workload == "cfg5" and not args.cpu_sample:
        cpu = {"value": None, "unit": "pairs/s", "cores": cores, "kind": "reference",
               "sample": "not timed by default: one 100 kbp pair takes the reference minutes per core (BASELINE.md); pass --cpu-sample N"}
    elif not args.no_cpu_baseline:
        cpu, par = cpu_and_parity(env, w)
    if par is None:
        par = {"checked_pairs": 0, "mismatches": None, "status_histogram": w["status_hist"], "retried_pairs": w["stats"]["retried_pairs"]}

    line = {"metric": "aligned pairs/sec", "value": w["value"], "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": w["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": config,
            "clocks": w["clocks"], "e2e": w["e2e"], "gpu_launches": int(w["stats"]["kernel_launches"]) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "parity": par,
            "host": {"cores": cores, "worker_threads_per_rank": max(1, cores // max(world, 1))}}
    if args.workload == "mixed":
        parts = mixed_parts(env, args.steps)
        total = sum(p["ms_per_step"] for p in parts)
        line["bucketing"] = {"parts": parts, "sum_of_parts_ms": total, "mixed_ms": w["ms_per_step"],
                             "mixed_vs_sum_of_parts": total / w["ms_per_step"],
                             "note": "the same pairs shuffled into one batch (one tier plan per length bucket) against its three "
                                     "length classes aligned as separate batches; >= 0.9 means bucketing costs < 10 %"}
    if sec is not None:
        s_cpu = s_par = None
        if not args.no_cpu_baseline:
            s_cpu, s_par = cpu_and_parity(env, sec, c_rate=False)
        line["secondary"] = {
            "config": {"workload": WORKLOAD_DESC["cfg1"], "pairs_per_gpu": sec["n_pairs"]},
            "value": sec["value"], "unit": "pairs/s", "ms_per_step": sec["ms_per_step"], "e2e": sec["e2e"],
            "cpu_baseline": s_cpu, "parity": s_par,
            "e2e_vs_cpu_baseline": (sec["e2e"]["value"] / s_cpu["value"]) if (sec["e2e"] and s_cpu and s_cpu.get("value")) else None,
            "roofline": roofline_block("cfg1", sec["n_pairs"], sec["stats"]["cells"], sec["ms_per_step"] / 1e3, sec["p_mean"],
                                       sec["t_mean"], False, True, w["clocks"], peaks),
            "gpu_launches_per_step": int(sec["stats"]["kernel_launches"])}
        line["single_pair_latency_us"] = single_pair_latency()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
