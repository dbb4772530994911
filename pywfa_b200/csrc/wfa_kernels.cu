/*
 * wfa_kernels.cu -- sm_100a kernels of the batched wavefront aligner.
 *
 *   wfa_align_kernel<TWO_P, FULL, MODE>   persistent kernel; every thread group pulls read
 *       pairs from a device work queue and runs wfagpu::align_pair (wfa_core.cuh) on them.
 *       MODE 0: warp-per-pair, wavefront ring + packed sequences in shared memory
 *       MODE 1: block-per-pair, ring in shared memory
 *   wfa_vec_kernel<TWO_P, FULL, NW, HEUR>   packed-halfword tier (wfa_vec.cuh): the general case up to 12 kbp
 *   wfa_grid_kernel<TWO_P, FULL>   several CTAs per pair, rings in an L2-resident HBM arena (long reads)
 *   wfa_reg_kernel<P, DX, DOE, FULL>   register-resident tier (wfa_reg.cuh): short gap-affine reads
 *   cigar_count_kernel / cigar_scan_kernel / cigar_gather_kernel   order the per-pair CIGAR
 *       runs into the caller's layout (cig_off[n+1] + contiguous run words).
 *
 * Replaces, for the batch, the reference's per-pair loop wavefront_unialign
 * (W/wavefront/wavefront_unialign.c:241-273) and everything it calls.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <type_traits>

#include "wfa_core.cuh"
#include "wfa_reg.cuh"
#include "wfa_vec.cuh"
#include "wfa_launch.h"
#include "wfa_pack16.cuh"

namespace wfagpu {

constexpr int MAX_RED = 11;

/* ---- thread groups ------------------------------------------------------------------ */
struct WarpGroup {
  static constexpr bool kGrid = false;
  int rank, size, lrank, lsize;
  __device__ __forceinline__ void sync() { __syncwarp(); }
  __device__ __forceinline__ void lsync() { __syncwarp(); }
  template <int N>
  __device__ __forceinline__ void allmin(int (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __reduce_min_sync(0xffffffffu, v[i]);
  }
  __device__ __forceinline__ int bcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }
  __device__ __forceinline__ long long bcastll(long long v) { return __shfl_sync(0xffffffffu, v, 0); }
};

struct BlockGroup {
  static constexpr bool kGrid = false;
  int rank, size, lrank, lsize;
  int* red;        /* 2 * MAX_RED * 32 ints of shared memory */
  int parity;
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void lsync() { __syncthreads(); }
  template <int N>
  __device__ __forceinline__ void allmin(int (&v)[N]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int* buf = red + parity * (MAX_RED * 32);
    parity ^= 1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int r = __reduce_min_sync(0xffffffffu, v[i]);
      if (lane == 0) buf[i * 32 + warp] = r;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int t = (lane < nw) ? buf[i * 32 + lane] : INT_MAX;
      v[i] = __reduce_min_sync(0xffffffffu, t);
    }
  }
  __device__ __forceinline__ int bcast(int v) {
    int* buf = red + parity * (MAX_RED * 32);
    parity ^= 1;
    if (rank == 0) buf[0] = v;
    __syncthreads();
    return buf[0];
  }
  __device__ __forceinline__ long long bcastll(long long v) {
    const int lo = bcast((int)(v & 0xffffffffll));
    const int hi = bcast((int)(v >> 32));
    return ((long long)hi << 32) | (unsigned int)lo;
  }
};

/* ---- the persistent alignment kernel ------------------------------------------------- */
/* shared memory of one group: [metadata int4 x mr*NC][packed sequences][offset rings OffT] */
template <bool TWO_P, bool FULL, int MODE, class OffT>
__global__ void wfa_align_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool BLOCK = (MODE != 0);
  constexpr int NC = TWO_P ? 5 : 3;
  using G = typename std::conditional<BLOCK, BlockGroup, WarpGroup>::type;

  G g;
  int group_id;
  unsigned char* base;
  if (BLOCK) {
    g.rank = threadIdx.x; g.size = blockDim.x; g.lrank = g.rank; g.lsize = g.size;
    group_id = blockIdx.x;
    base = smem_raw;
  } else {
    g.rank = threadIdx.x & 31; g.size = 32; g.lrank = g.rank; g.lsize = 32;
    const int wpb = blockDim.x >> 5;
    group_id = blockIdx.x * wpb + (threadIdx.x >> 5);
    base = smem_raw + (size_t)(threadIdx.x >> 5) * P.group_bytes;
  }
  if constexpr (BLOCK) { g.red = reinterpret_cast<int*>(smem_raw + P.group_bytes); g.parity = 0; }

  GroupMem<OffT> gm;
  gm.meta = reinterpret_cast<int4*>(base);
  uint32_t* const sm_seq = reinterpret_cast<uint32_t*>(base + (size_t)P.mr * NC * 16);
  OffT* ringbase = (MODE == 2) ? reinterpret_cast<OffT*>(P.gring) + (long long)group_id * P.gring_elems
                               : reinterpret_cast<OffT*>(sm_seq + P.seq_words_cap);
  gm.ring[CM] = ringbase;
  gm.ring[CI1] = gm.ring[CM] + P.rm * P.wcap;
  gm.ring[CD1] = gm.ring[CI1] + P.r1 * P.wcap;
  gm.ring[CI2] = gm.ring[CD1] + P.r1 * P.wcap;
  gm.ring[CD2] = gm.ring[CI2] + (TWO_P ? P.r2 * P.wcap : 0);
  if (FULL) {
    gm.h_code = P.hist_code + (long long)group_id * P.hcap;
    gm.ops = P.rops + (long long)group_id * P.ropcap; gm.opcap = P.ropcap;
    gm.hmeta = P.hmeta + (long long)group_id * P.scap;
    gm.runs_stage = P.runs_stage + (long long)group_id * P.runcap;
  } else {
    gm.h_code = nullptr; gm.hmeta = nullptr; gm.runs_stage = nullptr; gm.ops = nullptr; gm.opcap = 0;
  }

  const int n_work = min(*P.n_work, P.work_limit);
  long long cells_acc = 0;
  bool gave_up = false;          /* this group's verdict of tier_gives_up */
  for (;;) {
    int w = 0;
    if (g.rank == 0) w = atomicAdd(P.work_counter, 1);
    w = g.bcast(w);
    if (w >= n_work) break;
    const int pid = P.worklist ? P.worklist[w] : w;
    const PairMeta pm = P.pairs[pid];
    const int plen = pm.plen, tlen = pm.tlen;
    const bool pbytes = P.byte_mode || pm.woff < 0;       /* this pair's sequences are bytes */
    const int wsh = pbytes ? 2 : 4;                       /* bases per word: 4 (bytes) or 16 (2-bit) */
    const int pwn = (plen + (1 << wsh) - 1) >> wsh, twn = (tlen + (1 << wsh) - 1) >> wsh;
    gm.wild = pbytes ? P.wildcard : -1;
    int rc;
    PairResult res;
    if ((P.seq_words_cap > 0 && pwn + twn + 2 > P.seq_words_cap) || tier_gives_up(P, w, gave_up)) {
      rc = PAIR_OVERFLOW;
    } else {
      const uint32_t* gw = pm.woff < 0 ? P.words2 + ~pm.woff : P.words + pm.woff;
      if (P.seq_words_cap > 0) {
        /* stage the 2-bit packed pair in shared memory (coalesced word loads) */
        uint32_t* sp = sm_seq; uint32_t* st = sm_seq + pwn + 1;
        for (int i = g.rank; i < pwn; i += g.size) sp[i] = gw[i];
        for (int i = g.rank; i < twn; i += g.size) st[i] = gw[pwn + i];
        if (g.rank == 0) { sp[pwn] = 0; st[twn] = 0; }
        gm.pw = sp; gm.tw = st;
      } else {
        gm.pw = gw; gm.tw = gw + pwn;      /* device buffer carries one pad word */
      }
      g.sync();
      rc = align_pair<G, OffT, TWO_P, FULL>(g, P, gm, plen, tlen, res);
    }
    if (g.rank == 0 && !gave_up && (pid & 7) == 0) tier_pair_note(P, rc == PAIR_OVERFLOW);
    if (rc == PAIR_OVERFLOW) {
      if (g.rank == 0) { const int idx = atomicAdd(P.retry_count, 1); P.retry_list[idx] = pid; }
    } else {
      cells_acc += res.cells;
      if (FULL) {
        int nr = g.bcast(res.nruns);
        long long rbase = 0;
        int st = res.status;
        if (nr > 0) {
          if (g.rank == 0) rbase = (long long)atomicAdd(P.runs_cursor, (unsigned long long)nr);
          rbase = g.bcastll(rbase);
          if (nr > P.runcap || (unsigned long long)(rbase + nr) > P.runs_tmp_cap) { st = ST_OOM; nr = 0; }
          g.sync();
          for (int i = g.rank; i < nr; i += g.size) P.runs_tmp[rbase + i] = gm.runs_stage[i];
        } else if (nr < 0) { st = ST_OOM; nr = 0; }
        if (g.rank == 0) {
          P.score[pid] = res.score; P.status[pid] = st;
          int4 l = make_int4(res.locs[0], res.locs[1], res.locs[2], res.locs[3]);
          if (nr == 0) l = make_int4(0, 0, 0, 0);
          reinterpret_cast<int4*>(P.locs)[pid] = l;
          P.nruns[pid] = nr; P.runs_base[pid] = rbase;
        }
      } else if (g.rank == 0) {
        P.score[pid] = res.score; P.status[pid] = res.status;
      }
    }
    g.sync();
  }
  if (g.rank == 0 && cells_acc) atomicAdd(P.cells_total, (unsigned long long)cells_acc);
}

/* ---- several CTAs per pair (long reads): the group spans `ncta` co-resident CTAs ------------ */
/*
 * 100 kbp pairs have wavefronts of up to 2*10^5 diagonals and ~10 GB of origin bytes each, so only
 * about a dozen pairs fit in HBM at a time: one CTA per pair would leave 90 % of the SMs idle.
 * Here the diagonals of one wavefront are dealt over all threads of `ncta` CTAs (cooperative
 * launch, so they are co-resident); the offset rings live in an L2-resident HBM arena and are
 * read with ld.global.cg; one software barrier per score (inside the min-reduction that trims
 * the wavefront) orders the ring stores of score s before the loads of score s+1.
 */
struct GridScratch {            /* one per group, in HBM, zero-initialised by the host before the launch */
  unsigned int count;           /* barrier arrivals, monotonic */
  int slot;                     /* broadcast value */
  long long slot_ll;
  int red[3][MAX_RED + 1];      /* three rotating reduction sets */
};

struct GridGroup {
  static constexpr bool kGrid = true;
  int rank, size, lrank, lsize;
  int* red;                     /* 2 * MAX_RED * 32 ints of shared memory (CTA-level reduction) */
  int parity;
  GridScratch* gs;
  unsigned int target;          /* arrivals that complete the next barrier */
  int ncta, phase;
  __device__ __forceinline__ void lsync() { __syncthreads(); }
  __device__ __forceinline__ void sync() {
    /* the pattern of cooperative groups' grid sync: the CTA barrier before makes the CTA's ring stores
     * happen-before thread 0's release fence and arrival; the acquire fence after the poll and the CTA barrier
     * hand what thread 0 saw to the CTA.  fence.acq_rel (MEMBAR.ALL.GPU) instead of __threadfence()'s
     * sequentially-consistent MEMBAR.SC.GPU.  (A red.release / ld.acquire pair in inline PTX made ptxas spill
     * 96 bytes per thread in this 64-register kernel.) */
    __syncthreads();
    if (lrank == 0) {
#ifdef WFA_GRID_SC_FENCE
      __threadfence();
#else
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
      atomicAdd(&gs->count, 1u);
      target += (unsigned int)ncta;
      while ((int)(*reinterpret_cast<volatile unsigned int*>(&gs->count) - target) < 0) { }
#ifdef WFA_GRID_SC_FENCE
      __threadfence();
#else
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
    }
    __syncthreads();
  }
  template <int N>
  __device__ __forceinline__ void allmin(int (&v)[N]) {
    /* CTA level */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int* buf = red + parity * (MAX_RED * 32);
    parity ^= 1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int r = __reduce_min_sync(0xffffffffu, v[i]);
      if (lane == 0) buf[i * 32 + warp] = r;
    }
    __syncthreads();
    /* group level: warp 0 folds the CTA's minima into the group's reduction set */
    int* const gred = gs->red[phase];
    if (warp == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int t = (lane < nw) ? buf[i * 32 + lane] : INT_MAX;
        const int r = __reduce_min_sync(0xffffffffu, t);
        if (lane == 0 && r != INT_MAX) atomicMin(&gred[i], r);
      }
    }
    sync();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = ld_cg(&gred[i]);
    /* the set used two reductions from now was last read before the barrier above */
    {
      int* const nxt = gs->red[phase == 0 ? 2 : phase - 1];
      if (rank < MAX_RED) nxt[rank] = INT_MAX;
      phase = (phase == 2) ? 0 : phase + 1;
    }
  }
  __device__ __forceinline__ int bcast(int v) {
    if (rank == 0) *reinterpret_cast<volatile int*>(&gs->slot) = v;
    sync();
    const int r = ld_cg(&gs->slot);
    sync();
    return r;
  }
  __device__ __forceinline__ long long bcastll(long long v) {
    if (rank == 0) *reinterpret_cast<volatile long long*>(&gs->slot_ll) = v;
    sync();
    const long long r = ld_cg(&gs->slot_ll);
    sync();
    return r;
  }
};

template <bool TWO_P, bool FULL>
__global__ void __launch_bounds__(512, 2) wfa_grid_kernel(const __grid_constant__ KParams P, GridScratch* scratch, int ncta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = TWO_P ? 5 : 3;
  GridGroup g;
  const int group_id = blockIdx.x / ncta, cta = blockIdx.x % ncta;
  g.lrank = threadIdx.x; g.lsize = blockDim.x;
  g.rank = cta * blockDim.x + threadIdx.x; g.size = ncta * blockDim.x;
  g.red = reinterpret_cast<int*>(smem_raw + P.group_bytes); g.parity = 0;
  g.gs = scratch + group_id; g.target = 0; g.ncta = ncta; g.phase = 0;

  GroupMem<int32_t> gm;
  gm.meta = reinterpret_cast<int4*>(smem_raw);
  uint32_t* const sm_seq = reinterpret_cast<uint32_t*>(smem_raw + (size_t)P.mr * NC * 16);
  int32_t* ringbase = reinterpret_cast<int32_t*>(P.gring) + (long long)group_id * P.gring_elems;
  gm.ring[CM] = ringbase;
  gm.ring[CI1] = gm.ring[CM] + P.rm * P.wcap;
  gm.ring[CD1] = gm.ring[CI1] + P.r1 * P.wcap;
  gm.ring[CI2] = gm.ring[CD1] + P.r1 * P.wcap;
  gm.ring[CD2] = gm.ring[CI2] + (TWO_P ? P.r2 * P.wcap : 0);
  if (FULL) {
    gm.h_code = P.hist_code + (long long)group_id * P.hcap;
    gm.ops = P.rops + (long long)group_id * P.ropcap; gm.opcap = P.ropcap;
    gm.hmeta = P.hmeta + (long long)group_id * P.scap;
    gm.runs_stage = P.runs_stage + (long long)group_id * P.runcap;
  } else {
    gm.h_code = nullptr; gm.hmeta = nullptr; gm.runs_stage = nullptr; gm.ops = nullptr; gm.opcap = 0;
  }
  /* the host zeroed the scratch (barrier count); the reduction sets start at +inf */
  if (g.rank < 3 * (MAX_RED + 1)) (&g.gs->red[0][0])[g.rank] = INT_MAX;
  g.sync();
  const int n_work = min(*P.n_work, P.work_limit);
  long long cells_acc = 0;
  bool gave_up = false;          /* this group's verdict of tier_gives_up */
  for (;;) {
    int w = 0;
    if (g.rank == 0) w = atomicAdd(P.work_counter, 1);
    w = g.bcast(w);
    if (w >= n_work) break;
    const int pid = P.worklist ? P.worklist[w] : w;
    const PairMeta pm = P.pairs[pid];
    const int plen = pm.plen, tlen = pm.tlen;
    const bool pbytes = P.byte_mode || pm.woff < 0;
    const int wsh = pbytes ? 2 : 4;
    const int pwn = (plen + (1 << wsh) - 1) >> wsh, twn = (tlen + (1 << wsh) - 1) >> wsh;
    gm.wild = pbytes ? P.wildcard : -1;
    int rc;
    PairResult res;
    if (P.seq_words_cap > 0 && pwn + twn + 2 > P.seq_words_cap) {
      rc = PAIR_OVERFLOW;
    } else {
      const uint32_t* gw = pm.woff < 0 ? P.words2 + ~pm.woff : P.words + pm.woff;
      if (P.seq_words_cap > 0) {
        uint32_t* sp = sm_seq; uint32_t* st = sm_seq + pwn + 1;
        for (int i = g.lrank; i < pwn; i += g.lsize) sp[i] = gw[i];
        for (int i = g.lrank; i < twn; i += g.lsize) st[i] = gw[pwn + i];
        if (g.lrank == 0) { sp[pwn] = 0; st[twn] = 0; }
        gm.pw = sp; gm.tw = st;
      } else {
        gm.pw = gw; gm.tw = gw + pwn;
      }
      g.lsync();
      rc = align_pair<GridGroup, int32_t, TWO_P, FULL>(g, P, gm, plen, tlen, res);
    }
    if (rc == PAIR_OVERFLOW) {
      if (g.rank == 0) { const int idx = atomicAdd(P.retry_count, 1); P.retry_list[idx] = pid; }
    } else {
      cells_acc += res.cells;
      if (FULL) {
        int nr = g.bcast(res.nruns);
        long long rbase = 0;
        int st = res.status;
        if (nr > 0) {
          if (g.rank == 0) rbase = (long long)atomicAdd(P.runs_cursor, (unsigned long long)nr);
          rbase = g.bcastll(rbase);
          if (nr > P.runcap || (unsigned long long)(rbase + nr) > P.runs_tmp_cap) { st = ST_OOM; nr = 0; }
          for (int i = g.rank; i < nr; i += g.size) P.runs_tmp[rbase + i] = ld_cg(gm.runs_stage + i);
        } else if (nr < 0) { st = ST_OOM; nr = 0; }
        if (g.rank == 0) {
          P.score[pid] = res.score; P.status[pid] = st;
          int4 l = make_int4(res.locs[0], res.locs[1], res.locs[2], res.locs[3]);
          if (nr == 0) l = make_int4(0, 0, 0, 0);
          reinterpret_cast<int4*>(P.locs)[pid] = l;
          P.nruns[pid] = nr; P.runs_base[pid] = rbase;
        }
      } else if (g.rank == 0) {
        P.score[pid] = res.score; P.status[pid] = res.status;
      }
    }
    g.sync();
  }
  if (g.rank == 0 && cells_acc) atomicAdd(P.cells_total, (unsigned long long)cells_acc);
}

#ifndef WFA_REG_MINB
#define WFA_REG_MINB 7      /* resident CTAs per SM the register tier is compiled for (register budget; measured r01 with
                               the 192-diagonal window on cfg2 / cfg1: 6 -> 46.1 / 115.2, 7 -> 46.5 / 117.1, 8 -> 40.6 / 111.1
                               M pairs/s; with the 256-diagonal window only: 5 -> 41.6, 6 -> 42.8, 7 -> 37.3) */
#endif
/* r02, after the leaner extension (67 registers at 7 CTAs): 8 CTAs per SM (64 registers, no spills) for the 128- and
 * 192-diagonal windows: cfg1 150.6 -> 165.0, cfg2 68.0 -> 69.8 M pairs/s */
#ifndef WFA_REG_MINB2
#define WFA_REG_MINB2 8                /* ... for the 128-diagonal window */
#endif
#ifndef WFA_REG_MINB3
#define WFA_REG_MINB3 8                /* ... for the 192-diagonal window */
#endif
#ifndef WFA_REG_MINB4
#define WFA_REG_MINB4 6                /* ... for the 256-diagonal window (80 registers: no spills; r01: 6 -> 42.8, 7 -> 37.3 M pairs/s on that window alone) */
#endif
/* (shapes with a deep M ring -- (4, 7, 1): seven slots per packed register -- keep the 72-register budget) */
__host__ __device__ constexpr int reg_min_blocks(int regs, int ring) {
  return regs >= 4 ? WFA_REG_MINB4 : ring > 4 ? WFA_REG_MINB : regs == 2 ? WFA_REG_MINB2 : WFA_REG_MINB3;
}
/* ---- the register-resident tier (wfa_reg.cuh): warp-per-pair, wavefronts in registers ---- */
/* shared memory of one warp: the sequence windows of the pair (seq_words_cap words, one per base + 2) */
template <int P, int DX, int DOE, bool FULL>
__global__ void __launch_bounds__(128, reg_min_blocks(P, DX > DOE ? DX : DOE)) wfa_reg_kernel(const __grid_constant__ KParams K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool HS = reg_hist_in_smem(P, FULL);        /* origin arena, edit-operation stack and run staging in shared memory */
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp_id = blockIdx.x * (blockDim.x >> 5) + wib;
  const RegSmem L = reg_smem_layout(P, FULL, K.seq_words_cap, K.ropcap, K.rhrows);
  unsigned char* const wbase = smem_raw + (size_t)wib * K.group_bytes;
  uint32_t* const sm_seq = reinterpret_cast<uint32_t*>(wbase);
  uint32_t* const sm_pk = reinterpret_cast<uint32_t*>(wbase + L.pk_off());

  RegParams R;
  R.match = K.match; R.g = K.g; R.max_steps = K.max_steps; R.pos_score = K.pos_score;
  R.endsfree = K.endsfree; R.pbf = K.pbf; R.pef = K.pef; R.tbf = K.tbf; R.tef = K.tef;
  R.hrows = K.rhrows; R.opcap = K.ropcap; R.runcap = K.runcap;
  R.kbase = K.reg_kbase; R.c_lo = K.reg_clo; R.c_hi = K.reg_chi;
  uint8_t* const hist_p = !FULL ? nullptr : HS ? wbase + L.hist_off() : K.rhist + (long long)warp_id * K.rhist_bytes;
  const lv::histref hist = lv::make_histref(hist_p, HS);
  uint8_t* const ops = !FULL ? nullptr : HS ? wbase + L.ops_off() : K.rops + (long long)warp_id * K.ropcap;
  /* the run staging re-uses the arena: the backward walk has left it when the replay emits runs */
  uint32_t* const stage = !FULL ? nullptr : HS ? reinterpret_cast<uint32_t*>(hist_p) : K.runs_stage + (long long)warp_id * K.runcap;

  const int n_work = min(*K.n_work, K.work_limit);
  long long cells_acc = 0;
  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(K.work_counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= n_work) break;
    const int pid = K.worklist ? K.worklist[w] : w;
    const PairMeta pm = K.pairs[pid];
    const int plen = pm.plen, tlen = pm.tlen;
    const int pwn = (plen + 15) >> 4;
    int rc = PAIR_OVERFLOW;
    PairResult res;
    /* (woff < 0: byte-mode pair, scalar tiers.  No adaptive tier skipping here: the pairs this tier cannot
     * hold cost it little, and this loop is sensitive to every extra live register) */
    if (pm.woff >= 0 && plen + tlen + 2 <= K.seq_words_cap && plen <= REG_MAX_LEN && tlen <= REG_MAX_LEN) {
      /* packed words in HBM (the batch buffer carries one pad word) -> per-base windows in smem */
      const uint32_t* gp = K.words + pm.woff;
      const uint32_t* gt = gp + pwn;
      uint32_t* sp = sm_seq; uint32_t* st = sm_seq + plen + 1;
      build_windows(gp, plen, sp);
      build_windows(gt, tlen, st);
      if (HS) {
        /* the replay of the backtrace re-extends matches from the packed words: keep them close */
        const int nw = pwn + ((tlen + 15) >> 4) + 1;
        for (int i = lane; i < nw; i += 32) sm_pk[i] = gp[i];
        gp = sm_pk; gt = sm_pk + pwn;
      }
      __syncwarp();
      rc = align_pair_reg<P, DX, DOE, FULL>(R, gp, gt, lv::make_seqref(sp), lv::make_seqref(st), plen, tlen, hist, ops,
                                            stage, lane == 0, res);
    }
    if (rc == PAIR_OVERFLOW) {
      if (lane == 0) { const int idx = atomicAdd(K.retry_count, 1); K.retry_list[idx] = pid; }
    } else {
      cells_acc += res.cells;
      if (FULL) {
        int nr = __shfl_sync(0xffffffffu, res.nruns, 0);
        long long rbase = 0;
        int stt = res.status;
        if (nr > 0) {
          if (lane == 0) rbase = (long long)atomicAdd(K.runs_cursor, (unsigned long long)nr);
          rbase = __shfl_sync(0xffffffffu, rbase, 0);
          if (nr > K.runcap || (unsigned long long)(rbase + nr) > K.runs_tmp_cap) { stt = ST_OOM; nr = 0; }
          __syncwarp();
          for (int i = lane; i < nr; i += 32) K.runs_tmp[rbase + i] = stage[i];
        } else if (nr < 0) { stt = ST_OOM; nr = 0; }
        if (lane == 0) {
          K.score[pid] = res.score; K.status[pid] = stt;
          int4 l = make_int4(res.locs[0], res.locs[1], res.locs[2], res.locs[3]);
          if (nr == 0) l = make_int4(0, 0, 0, 0);
          reinterpret_cast<int4*>(K.locs)[pid] = l;
          K.nruns[pid] = nr; K.runs_base[pid] = rbase;
        }
      } else if (lane == 0) {
        K.score[pid] = res.score; K.status[pid] = res.status;
      }
    }
    __syncwarp();
  }
  if (lane == 0 && cells_acc) atomicAdd(K.cells_total, (unsigned long long)cells_acc);
}

/* ---- one pair, one warp, one launch: the low-latency path of wfagpu_align_pair ------------------ */
/*
 * Replaces ONE wavefront_align call (W/wavefront/wavefront_align.c:212-241) for callers that loop over
 * pairs (pywfa's a(text, pattern)).  Everything a batch needs several kernels and copies for happens in
 * this one launch: the bases are read straight out of the caller-visible mailbox in mapped host memory,
 * packed and turned into windows in shared memory, aligned on the 256-diagonal register window with the
 * origin arena, edit stack and run staging in shared memory, and the results (incl. the CIGAR runs) are
 * written back into the mailbox.  rc = 1 asks the host to take the batch path (non-ACGT bytes, a pair
 * the window cannot hold).
 */
template <bool FULL>
__global__ void __launch_bounds__(32) wfa_pair_kernel(const __grid_constant__ KParams K, PairBox* box) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int P = 4;
  const int lane = threadIdx.x & 31;
  const int plen = box->plen, tlen = box->tlen;
  const int pwn = (plen + 15) >> 4, twn = (tlen + 15) >> 4;
  uint32_t* const sm_pk = reinterpret_cast<uint32_t*>(smem_raw);                     /* packed words: PAIR_PK_WORDS */
  uint32_t* const sm_win = sm_pk + PAIR_PK_WORDS;                                    /* windows: 2 * PAIR_MAX_LEN + 2 */
  uint8_t* const ops = reinterpret_cast<uint8_t*>(sm_win + 2 * PAIR_MAX_LEN + 2);    /* edit stack: PAIR_OPS_BYTES */
  uint8_t* const hist_p = ops + PAIR_OPS_BYTES;                                      /* origin arena / run staging */
  uint32_t* const stage = reinterpret_cast<uint32_t*>(hist_p);
  const uint8_t* const ascii = reinterpret_cast<const uint8_t*>(box) + PAIR_ASCII_OFF;
  bool bad = false;
  for (int w = lane; w < pwn + twn; w += 32) {
    const bool is_t = w >= pwn;
    const int j = is_t ? w - pwn : w;
    const int len = is_t ? tlen : plen;
    sm_pk[w] = pack16<true>(ascii + (is_t ? PAIR_TEXT_OFF : 0) + 16 * j, min(16, len - 16 * j), bad);
  }
  if (lane == 0) sm_pk[pwn + twn] = 0;
  __syncwarp();
  if (__any_sync(0xffffffffu, bad)) { if (lane == 0) box->rc = 1; return; }
  uint32_t* sp = sm_win; uint32_t* st = sm_win + plen + 1;
  build_windows(sm_pk, plen, sp);
  build_windows(sm_pk + pwn, tlen, st);
  __syncwarp();
  RegParams R;
  R.match = K.match; R.g = K.g; R.max_steps = K.max_steps; R.pos_score = K.pos_score;
  R.endsfree = K.endsfree; R.pbf = K.pbf; R.pef = K.pef; R.tbf = K.tbf; R.tef = K.tef;
  R.hrows = PAIR_HIST_ROWS; R.opcap = PAIR_OPS_BYTES; R.runcap = plen + tlen + 2;
  const RegWindow rw = reg_window(P, K.endsfree, K.match, K.pbf, K.tbf);
  R.kbase = rw.kbase; R.c_lo = rw.c_lo; R.c_hi = rw.c_hi;
  PairResult res;
  const lv::histref hist = lv::make_histref(hist_p, true);
  const int rc = align_pair_reg<P, 2, 4, FULL, FULL>(R, sm_pk, sm_pk + pwn, lv::make_seqref(sp), lv::make_seqref(st), plen, tlen,
                                                     hist, ops, stage, lane == 0, res);
  if (rc == PAIR_OVERFLOW) { if (lane == 0) box->rc = 1; return; }
  int nr = 0, stt = res.status;
  if (FULL) {
    nr = __shfl_sync(0xffffffffu, res.nruns, 0);
    if (nr < 0 || nr > R.runcap) { if (lane == 0) box->rc = 1; return; }
    __syncwarp();
    uint32_t* const out = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(box) + PAIR_RUNS_OFF);
    for (int i = lane; i < nr; i += 32) out[i] = stage[i];
  }
  if (lane == 0) {
    box->score = res.score; box->status = stt; box->nruns = nr;
    for (int q = 0; q < 4; ++q) box->locs[q] = (FULL && nr > 0) ? res.locs[q] : 0;
    box->cells = res.cells;
    box->rc = 0;
  }
}

cudaError_t launch_pair(const KParams& P, bool full, PairBox* box, cudaStream_t st) {
  if (full) wfa_pair_kernel<true><<<1, 32, PAIR_SMEM_BYTES, st>>>(P, box);
  else wfa_pair_kernel<false><<<1, 32, PAIR_SMEM_BYTES, st>>>(P, box);
  return cudaGetLastError();
}

/* ---- the packed-halfword tier (wfa_vec.cuh): NW warps per pair, rings in shared memory ---- */
/* shared memory of one group: [metadata int4 x mr*3][flags 256 B][2 step plans 512 B][packed sequences][offset rings] */
template <bool TWO_P, bool FULL, int NW, int HEUR>
__global__ void __launch_bounds__(NW == 1 ? 128 : NW * 32, NW == 1 ? 5 : NW == 8 ? 2 : 1) wfa_vec_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sh_i;
  __shared__ long long sh_ll;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int rank = NW == 1 ? lane : (int)threadIdx.x;
  const int gsize = NW * 32;
  const int group_id = NW == 1 ? (int)(blockIdx.x * (blockDim.x >> 5) + wib) : (int)blockIdx.x;
  unsigned char* const base = smem_raw + (NW == 1 ? (size_t)wib * P.group_bytes : 0);

  vec::VMem vm;
  vm.meta = reinterpret_cast<int4*>(base);
  vm.flags = reinterpret_cast<int*>(base + (size_t)P.mr * 48);
  vm.plan = reinterpret_cast<vec::PlanOut*>(base + (size_t)P.mr * 48 + 256);
  uint32_t* const sm_seq = reinterpret_cast<uint32_t*>(base + (size_t)P.mr * 48 + 768);
  vm.ring = sm_seq + P.seq_words_cap;
  if (FULL) {
    vm.h_code = P.hist_code + (long long)group_id * P.hcap;
    vm.ops = P.rops + (long long)group_id * P.ropcap; vm.opcap = P.ropcap;
    vm.hmeta = P.hmeta + (long long)group_id * P.scap;
    vm.runs_stage = P.runs_stage + (long long)group_id * P.runcap;
  } else {
    vm.h_code = nullptr; vm.hmeta = nullptr; vm.runs_stage = nullptr; vm.ops = nullptr; vm.opcap = 0;
  }
  auto bcast = [&](int v) -> int {
    if (NW == 1) return __shfl_sync(0xffffffffu, v, 0);
    if (rank == 0) sh_i = v;
    __syncthreads();
    const int r = sh_i;
    __syncthreads();
    return r;
  };
  auto bcastll = [&](long long v) -> long long {
    if (NW == 1) return __shfl_sync(0xffffffffu, v, 0);
    if (rank == 0) sh_ll = v;
    __syncthreads();
    const long long r = sh_ll;
    __syncthreads();
    return r;
  };

  const int n_work = min(*P.n_work, P.work_limit);
  long long cells_acc = 0;
  bool gave_up = false;          /* this group's verdict of tier_gives_up */
  for (;;) {
    int w = 0;
    if (rank == 0) w = atomicAdd(P.work_counter, 1);
    w = bcast(w);
    if (w >= n_work) break;
    const int pid = P.worklist ? P.worklist[w] : w;
    const PairMeta pm = P.pairs[pid];
    const int plen = pm.plen, tlen = pm.tlen;
    const int pwn = (plen + 15) >> 4, twn = (tlen + 15) >> 4;
    int rc = PAIR_OVERFLOW;
    PairResult res;
    const int need_words = P.vec_seqw ? plen + tlen + 2 : pwn + twn + 2;
    /* (woff < 0: byte-mode pair, scalar tiers).  Only `gave_up` and `pid` stay live across the alignment:
     * the step loop runs at the register limit */
    if (pm.woff >= 0 && !(NW == 1 && HEUR == 0 && tier_gives_up(P, w, gave_up)) && need_words <= P.seq_words_cap && plen <= VEC_MAX_LEN && tlen <= VEC_MAX_LEN) {
      const uint32_t* gw = P.words + pm.woff;
      vm.bpw = gw; vm.btw = gw + pwn; vm.seqw = P.vec_seqw;
      if (P.vec_seqw) {
        /* per-base windows: word i = the 16 bases from position i, first base in the top bits
         * (the packed batch buffer is readable one word past every sequence) */
        uint32_t* sp = sm_seq; uint32_t* st = sm_seq + plen + 1;
        for (int i = rank; i <= plen; i += gsize) sp[i] = i < plen ? __brev(__funnelshift_r(gw[i >> 4], gw[(i >> 4) + 1], (i & 15) << 1)) : 0u;
        const uint32_t* gt = gw + pwn;
        for (int i = rank; i <= tlen; i += gsize) st[i] = i < tlen ? __brev(__funnelshift_r(gt[i >> 4], gt[(i >> 4) + 1], (i & 15) << 1)) : 0u;
        vm.pw = sp; vm.tw = st;
      } else {
        uint32_t* sp = sm_seq; uint32_t* st = sm_seq + pwn + 1;
        for (int i = rank; i < pwn; i += gsize) sp[i] = gw[i];
        for (int i = rank; i < twn; i += gsize) st[i] = gw[pwn + i];
        if (rank == 0) { sp[pwn] = 0; st[twn] = 0; }
        vm.pw = sp; vm.tw = st;
      }
      vec::gsync<NW>();
      rc = vec::align_pair_vec<TWO_P, FULL, NW, HEUR>(P, vm, plen, tlen, res);
    }
    /* (adaptive tier skipping only where it pays: the warp-per-pair tier without cut-offs is the one whole
     * batches overflow; elsewhere its live state costs the step loop registers: cfg3 -4 %, cfg4-adaptive -6 %) */
    if (NW == 1 && HEUR == 0 && rank == 0 && !gave_up && (pid & 7) == 0 && P.pairs[pid].woff >= 0) tier_pair_note(P, rc == PAIR_OVERFLOW);
    if (rc == PAIR_OVERFLOW) {
      if (rank == 0) { const int idx = atomicAdd(P.retry_count, 1); P.retry_list[idx] = pid; }
    } else {
      cells_acc += res.cells;
      if (FULL) {
        int nr = bcast(res.nruns);
        long long rbase = 0;
        int stt = res.status;
        if (nr > 0) {
          if (rank == 0) rbase = (long long)atomicAdd(P.runs_cursor, (unsigned long long)nr);
          rbase = bcastll(rbase);
          if (nr > P.runcap || (unsigned long long)(rbase + nr) > P.runs_tmp_cap) { stt = ST_OOM; nr = 0; }
          vec::gsync<NW>();
          for (int i = rank; i < nr; i += gsize) P.runs_tmp[rbase + i] = vm.runs_stage[i];
        } else if (nr < 0) { stt = ST_OOM; nr = 0; }
        if (rank == 0) {
          P.score[pid] = res.score; P.status[pid] = stt;
          int4 l = make_int4(res.locs[0], res.locs[1], res.locs[2], res.locs[3]);
          if (nr == 0) l = make_int4(0, 0, 0, 0);
          reinterpret_cast<int4*>(P.locs)[pid] = l;
          P.nruns[pid] = nr; P.runs_base[pid] = rbase;
        }
      } else if (rank == 0) {
        P.score[pid] = res.score; P.status[pid] = res.status;
      }
    }
    vec::gsync<NW>();
  }
  if (rank == 0 && cells_acc) atomicAdd(P.cells_total, (unsigned long long)cells_acc);
}

/* ---- tier probing: hand the untried rest of a work list to the next tier ------------------- */
__global__ void forward_rest_kernel(const int* __restrict__ worklist, const int* __restrict__ n_work, int from,
                                    int* __restrict__ retry_list, int* __restrict__ retry_count) {
  const int n = *n_work;
  const int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pid = worklist ? worklist[i] : i;
  retry_list[atomicAdd(retry_count, 1)] = pid;
}

cudaError_t launch_forward_rest(const int* worklist, const int* n_work, int from, long long n_bound, int* retry_list,
                                int* retry_count, cudaStream_t st) {
  const long long rest = n_bound - from;
  if (rest <= 0) return cudaSuccess;
  forward_rest_kernel<<<(int)((rest + 255) / 256), 256, 0, st>>>(worklist, n_work, from, retry_list, retry_count);
  return cudaGetLastError();
}

/* ---- CIGAR ordering ------------------------------------------------------------------ */
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void cigar_count_kernel(const int* __restrict__ nruns, long long n, long long* __restrict__ tile_sums) {
  __shared__ long long wsum[SCAN_THREADS / 32];
  const long long t0 = (long long)blockIdx.x * SCAN_TILE;
  long long acc = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long long i = t0 + j * SCAN_THREADS + threadIdx.x;
    if (i < n) acc += nruns[i];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int i = 0; i < SCAN_THREADS / 32; ++i) s += wsum[i];
    tile_sums[blockIdx.x] = s;
  }
}

/* single block: exclusive scan of the tile sums in place; total -> tile_sums[ntiles] */
__global__ void cigar_scan_kernel(long long* tile_sums, int ntiles) {
  __shared__ long long carry;
  __shared__ long long wtot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < ntiles; b += blockDim.x) {
    const int i = b + threadIdx.x;
    const long long v = (i < ntiles) ? tile_sums[i] : 0;
    long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = inc;
    __syncthreads();
    long long woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wtot[w];
    const long long excl = carry + woff + inc - v;
    if (i < ntiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}

/* per tile: exclusive offsets of every pair (cig_off); the runs are moved by cigar_copy_kernel */
__global__ void cigar_gather_kernel(const int* __restrict__ nruns, const long long* __restrict__ runs_base,
                                    long long n, const long long* __restrict__ tile_sums,
                                    const uint32_t* __restrict__ runs_tmp,
                                    long long* __restrict__ cig_off, uint32_t* __restrict__ runs_out,
                                    long long cig_base) {
  __shared__ long long wtot[SCAN_THREADS / 32];
  __shared__ long long carry_s;
  const long long t0 = (long long)blockIdx.x * SCAN_TILE;
  if (threadIdx.x == 0) carry_s = tile_sums[blockIdx.x];
  __syncthreads();
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long long i = t0 + j * SCAN_THREADS + threadIdx.x;
    const int v = (i < n) ? nruns[i] : 0;
    long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = inc;
    __syncthreads();
    long long woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wtot[w];
    const long long excl = carry_s + woff + inc - v;
    if (i < n) cig_off[i] = excl + cig_base;
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cig_off[n] = tile_sums[gridDim.x] + cig_base;
}

/* one warp per pair: its runs from the un-ordered pool to their place in the caller's layout, 32 words per
 * step (a lane-per-pair loop moved 8 MB in 83 us: neither load nor store coalesced) */
__global__ void cigar_copy_kernel(const int* __restrict__ nruns, const long long* __restrict__ runs_base, long long n,
                                  const uint32_t* __restrict__ runs_tmp, const long long* __restrict__ cig_off,
                                  uint32_t* __restrict__ runs_out, long long cig_base) {
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const int v = nruns[i];
  if (v <= 0) return;
  const uint32_t* src = runs_tmp + runs_base[i];
  uint32_t* dst = runs_out + (cig_off[i] - cig_base);
  for (int r = lane; r < v; r += 32) dst[r] = src[r];
}

/* ---- launch wrappers (C++ linkage, used by wfagpu_api.cpp) -------------------------- */
template <bool TWO_P, bool FULL, int MODE, class OffT>
static cudaError_t launch_one(const KParams& P, int grid, int block, size_t smem, cudaStream_t st) {
  /* the dynamic shared-memory limit was raised once per device by init_kernels() */
  wfa_align_kernel<TWO_P, FULL, MODE, OffT><<<grid, block, smem, st>>>(P);
  return cudaGetLastError();
}

template <bool TWO_P, bool FULL, int MODE, class OffT>
static cudaError_t init_one(int smem_optin) {
  return cudaFuncSetAttribute(wfa_align_kernel<TWO_P, FULL, MODE, OffT>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
}

template <bool TWO_P, bool FULL, int MODE, class OffT>
static int occupancy_one(int block, size_t smem) {
  auto kern = wfa_align_kernel<TWO_P, FULL, MODE, OffT>;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, block, smem) != cudaSuccess) return 0;
  return nb;
}

/* int16 offset rings exist for the shared-memory tiers (modes 0 and 1) */
#define WFA_DISPATCH(FN, ...)                                                   \
  do {                                                                          \
    const int key = (two_p ? 1 : 0) | (full ? 2 : 0) | (mode << 2);             \
    if (off16 && mode <= 1) {                                                   \
      switch (key) {                                                            \
        case 0: return FN<false, false, 0, int16_t>(__VA_ARGS__);               \
        case 1: return FN<true, false, 0, int16_t>(__VA_ARGS__);                \
        case 2: return FN<false, true, 0, int16_t>(__VA_ARGS__);                \
        case 3: return FN<true, true, 0, int16_t>(__VA_ARGS__);                 \
        case 4: return FN<false, false, 1, int16_t>(__VA_ARGS__);               \
        case 5: return FN<true, false, 1, int16_t>(__VA_ARGS__);                \
        case 6: return FN<false, true, 1, int16_t>(__VA_ARGS__);                \
        default: return FN<true, true, 1, int16_t>(__VA_ARGS__);                \
      }                                                                         \
    }                                                                           \
    switch (key) {                                                              \
      case 0: return FN<false, false, 0, int32_t>(__VA_ARGS__);                 \
      case 1: return FN<true, false, 0, int32_t>(__VA_ARGS__);                  \
      case 2: return FN<false, true, 0, int32_t>(__VA_ARGS__);                  \
      case 3: return FN<true, true, 0, int32_t>(__VA_ARGS__);                   \
      case 4: return FN<false, false, 1, int32_t>(__VA_ARGS__);                 \
      case 5: return FN<true, false, 1, int32_t>(__VA_ARGS__);                  \
      case 6: return FN<false, true, 1, int32_t>(__VA_ARGS__);                  \
      default: return FN<true, true, 1, int32_t>(__VA_ARGS__);                  \
    }                                                                           \
  } while (0)

cudaError_t launch_align(const KParams& P, bool two_p, bool full, int mode, bool off16, int grid, int block,
                         size_t smem, cudaStream_t st) {
  WFA_DISPATCH(launch_one, P, grid, block, smem, st);
}

int align_occupancy(bool two_p, bool full, int mode, bool off16, int block, size_t smem) {
  WFA_DISPATCH(occupancy_one, block, smem);
}

/* register tier: instantiated for the penalty shapes (x, o+e, e)/gcd = (2, 4, 1) -- pywfa's default 4/6/2 --,
 * (4, 7, 1) -- 4/6/1, bwa-like --, (1, 2, 1) -- 1/1/1, 2/2/2 -- and (1, 3, 1) -- 2/4/2, 1/2/1 -- with windows of
 * 128 / 192 / 256 diagonals, and score-only for the zero-opening shapes (1, 1, 1) and (2, 1, 1): edit, indel and
 * pywfa's default gap-linear 4/2 (metric_as_affine).  Other sets run on the packed-halfword tier (3-12x slower
 * on 150 bp, profiles/r02_results.md). */
#define WFA_REG_BOTH(STMT, PP, DX, DOE) do { if (full) { STMT(PP, DX, DOE, true); } else { STMT(PP, DX, DOE, false); } } while (0)
#define WFA_REG_DISPATCH_P(STMT, PP)                            \
  do {                                                          \
    if (shape == 0) WFA_REG_BOTH(STMT, PP, 2, 4);               \
    else if (shape == 1) { STMT(PP, 1, 1, false); }             \
    else if (shape == 2) { STMT(PP, 2, 1, false); }             \
    else if (shape == 3) WFA_REG_BOTH(STMT, PP, 4, 7);          \
    else if (shape == 4) WFA_REG_BOTH(STMT, PP, 1, 2);          \
    else WFA_REG_BOTH(STMT, PP, 1, 3);                          \
  } while (0)
#define WFA_REG_DISPATCH(STMT)                                  \
  do {                                                          \
    if (regs == 2) WFA_REG_DISPATCH_P(STMT, 2);                 \
    else if (regs == 3) WFA_REG_DISPATCH_P(STMT, 3);            \
    else WFA_REG_DISPATCH_P(STMT, 4);                           \
  } while (0)

static int reg_shape(int dx, int doe, int de, bool full) {
  if (de != 1) return -1;
  if (dx == 2 && doe == 4) return 0;
  if (!full && dx == 1 && doe == 1) return 1;
  if (!full && dx == 2 && doe == 1) return 2;
  if (dx == 4 && doe == 7) return 3;
  if (dx == 1 && doe == 2) return 4;
  if (dx == 1 && doe == 3) return 5;
  return -1;
}

bool reg_tier_supported(int dx, int doe, int de, int regs, bool full) {
  return reg_shape(dx, doe, de, full) >= 0 && (regs >= 2 && regs <= 4);
}

cudaError_t launch_reg(const KParams& P, int regs, bool full, int grid, int block, size_t smem, cudaStream_t st) {
  const int shape = reg_shape(P.dx, P.doe1, P.de1, full);
  if (shape < 0) return cudaErrorInvalidValue;
#define WFA_REG_LAUNCH(PP, DX, DOE, FULL) wfa_reg_kernel<PP, DX, DOE, FULL><<<grid, block, smem, st>>>(P)
  WFA_REG_DISPATCH(WFA_REG_LAUNCH);
#undef WFA_REG_LAUNCH
  return cudaGetLastError();
}

int reg_occupancy(const KParams& P, int regs, bool full, int block, size_t smem) {
  int nb = 0;
  const int shape = reg_shape(P.dx, P.doe1, P.de1, full);
  if (shape < 0) return 0;
#define WFA_REG_OCC(PP, DX, DOE, FULL) \
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wfa_reg_kernel<PP, DX, DOE, FULL>, block, smem) != cudaSuccess) nb = 0
  WFA_REG_DISPATCH(WFA_REG_OCC);
#undef WFA_REG_OCC
  return nb;
}

static cudaError_t init_pair(int smem_optin) {
  (void)smem_optin;
  cudaError_t e = cudaFuncSetAttribute(wfa_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wfa_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES);
  return e;
}

static cudaError_t init_reg(int smem_optin) {
  cudaError_t e = cudaSuccess;
  for (int regs = 2; regs <= 4; ++regs)
    for (int v = 0; v < 10; ++v) {
      /* (shape, scope) pairs that exist: shape 0 / 3 / 4 / 5 with both scopes, 1 / 2 score-only */
      static const int shapes[10] = {0, 0, 1, 2, 3, 3, 4, 4, 5, 5};
      static const bool fulls[10] = {false, true, false, false, false, true, false, true, false, true};
      const int shape = shapes[v];
      const bool full = fulls[v];
#define WFA_REG_INIT(PP, DX, DOE, FULL) \
  e = cudaFuncSetAttribute(wfa_reg_kernel<PP, DX, DOE, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)
      WFA_REG_DISPATCH(WFA_REG_INIT);
#undef WFA_REG_INIT
      if (e != cudaSuccess) return e;
    }
  return e;
}

/* several CTAs per pair: cooperative launch of groups * ncta CTAs; scratch = groups zeroed GridScratch */
size_t grid_scratch_bytes(int groups) { return sizeof(GridScratch) * (size_t)groups; }

cudaError_t launch_grid(const KParams& P, bool two_p, bool full, int groups, int ncta, size_t smem, void* scratch, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(scratch, 0, grid_scratch_bytes(groups), st);
  if (e != cudaSuccess) return e;
  KParams Pc = P;
  GridScratch* gs = static_cast<GridScratch*>(scratch);
  void* args[3] = {&Pc, &gs, &ncta};
  const void* fn = two_p ? (full ? (const void*)wfa_grid_kernel<true, true> : (const void*)wfa_grid_kernel<true, false>)
                         : (full ? (const void*)wfa_grid_kernel<false, true> : (const void*)wfa_grid_kernel<false, false>);
  return cudaLaunchCooperativeKernel(fn, dim3(groups * ncta), dim3(512), args, smem, st);
}

int grid_occupancy(bool two_p, bool full, size_t smem) {
  int nb = 0;
  const void* fn = two_p ? (full ? (const void*)wfa_grid_kernel<true, true> : (const void*)wfa_grid_kernel<true, false>)
                         : (full ? (const void*)wfa_grid_kernel<false, true> : (const void*)wfa_grid_kernel<false, false>);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 512, smem) != cudaSuccess) nb = 0;
  return nb;
}

static cudaError_t init_grid(int smem_optin) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(wfa_grid_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(wfa_grid_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(wfa_grid_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(wfa_grid_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
}

/* packed-halfword tier: NW = 1 (warp per pair), 8 or 16 warps per pair; HEUR = 0 none, 1 adaptive, 2 X-drop */
#define WFA_VEC_DISPATCH_H(STMT, TP, FU, NWW)                                            \
  do {                                                                                  \
    if (heur == 0) { STMT(TP, FU, NWW, 0); } else if (heur == 1) { STMT(TP, FU, NWW, 1); } else { STMT(TP, FU, NWW, 2); } \
  } while (0)
#define WFA_VEC_DISPATCH_K(STMT, NWW)                                                   \
  do {                                                                                  \
    const int key = (two_p ? 1 : 0) | (full ? 2 : 0);                                   \
    switch (key) {                                                                      \
      case 0: WFA_VEC_DISPATCH_H(STMT, false, false, NWW); break;                       \
      case 1: WFA_VEC_DISPATCH_H(STMT, true, false, NWW); break;                        \
      case 2: WFA_VEC_DISPATCH_H(STMT, false, true, NWW); break;                        \
      default: WFA_VEC_DISPATCH_H(STMT, true, true, NWW); break;                        \
    }                                                                                   \
  } while (0)
#define WFA_VEC_DISPATCH(STMT)                                                          \
  do {                                                                                  \
    if (nw == 1) WFA_VEC_DISPATCH_K(STMT, 1);                                           \
    else if (nw == 8) WFA_VEC_DISPATCH_K(STMT, 8);                                      \
    else WFA_VEC_DISPATCH_K(STMT, 16);                                                  \
  } while (0)

cudaError_t launch_vec(const KParams& P, bool two_p, bool full, int nw, int heur, int grid, int block, size_t smem, cudaStream_t st) {
#define WFA_VEC_LAUNCH(TP, FU, NWW, HH) wfa_vec_kernel<TP, FU, NWW, HH><<<grid, block, smem, st>>>(P)
  WFA_VEC_DISPATCH(WFA_VEC_LAUNCH);
#undef WFA_VEC_LAUNCH
  return cudaGetLastError();
}

int vec_occupancy(bool two_p, bool full, int nw, int heur, int block, size_t smem) {
  int nb = 0;
#define WFA_VEC_OCC(TP, FU, NWW, HH) \
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wfa_vec_kernel<TP, FU, NWW, HH>, block, smem) != cudaSuccess) nb = 0
  WFA_VEC_DISPATCH(WFA_VEC_OCC);
#undef WFA_VEC_OCC
  return nb;
}

static cudaError_t init_vec(int smem_optin) {
  cudaError_t e = cudaSuccess;
  for (int nw : {1, 8, 16})
    for (int heur = 0; heur < 3; ++heur)
      for (int k = 0; k < 4; ++k) {
        const bool two_p = k & 1, full = k & 2;
#define WFA_VEC_INIT(TP, FU, NWW, HH) \
  e = cudaFuncSetAttribute(wfa_vec_kernel<TP, FU, NWW, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 256)   /* the kernel has a few bytes of static shared memory */
        WFA_VEC_DISPATCH(WFA_VEC_INIT);
#undef WFA_VEC_INIT
        if (e != cudaSuccess) return e;
      }
  return e;
}

/* Raise the dynamic shared-memory limit of every instantiation on the CURRENT device, once, at
 * context creation: the attribute is per function and device, and changing it per launch would
 * race between the packing thread (occupancy queries) and the launching thread. */
static cudaError_t init_dispatch(bool two_p, bool full, int mode, bool off16, int smem_optin) {
  WFA_DISPATCH(init_one, smem_optin);
}
cudaError_t init_kernels(int smem_optin) {
  for (int two_p = 0; two_p < 2; ++two_p)
    for (int full = 0; full < 2; ++full)
      for (int mode = 0; mode < 2; ++mode)         /* (the HBM-ring mode runs as wfa_grid_kernel) */
        for (int off16 = 0; off16 < 2; ++off16) {
          const cudaError_t e = init_dispatch(two_p, full, mode, off16, smem_optin);
          if (e != cudaSuccess) return e;
        }
  const cudaError_t e = init_reg(smem_optin);
  if (e != cudaSuccess) return e;
  const cudaError_t eb = init_regb(smem_optin);
  if (eb != cudaSuccess) return eb;
  const cudaError_t e2 = init_vec(smem_optin);
  if (e2 != cudaSuccess) return e2;
  const cudaError_t e2b = init_vecb(smem_optin);
  if (e2b != cudaSuccess) return e2b;
  const cudaError_t e3 = init_pair(smem_optin);
  if (e3 != cudaSuccess) return e3;
  return init_grid(smem_optin);
}

size_t block_reduce_smem_bytes() { return 2 * MAX_RED * 32 * sizeof(int); }

cudaError_t launch_cigar_order(const int* nruns, const long long* runs_base, long long n,
                               long long* tile_sums, const uint32_t* runs_tmp, long long* cig_off,
                               uint32_t* runs_out, long long cig_base, cudaStream_t st) {
  const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  if (runs_out == nullptr) {   /* phase 1: counts + scan (total lands in tile_sums[ntiles]) */
    cigar_count_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(nruns, n, tile_sums);
    cigar_scan_kernel<<<1, 1024, 0, st>>>(tile_sums, ntiles);
  } else {
    cigar_gather_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(nruns, runs_base, n, tile_sums, runs_tmp, cig_off, runs_out, cig_base);
    if (n > 0) cigar_copy_kernel<<<(int)((n + 7) / 8), 256, 0, st>>>(nruns, runs_base, n, runs_tmp, cig_off, runs_out, cig_base);
  }
  return cudaGetLastError();
}

int cigar_order_tiles(long long n) { return (int)((n + SCAN_TILE - 1) / SCAN_TILE); }

}  // namespace wfagpu
