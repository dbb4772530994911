/*
 * wfa_vec_bytes.cu -- byte mode on the packed-halfword tier (sm_100a).
 *
 *   wfa_vecb_kernel<TWO_P, FULL, NW, HEUR>   wfa_vec.cuh compiled with WFA_VEC_BYTES: NW warps per pair
 *
 * Pairs that hold bytes other than ACGT, or are aligned with pywfa's wildcard= (pywfa/align.pyx:297-304,438-442;
 * the reference extends base by base through wavefront_extend_matches_custom,
 * W/wavefront/wavefront_extend_kernels.c:167-203), beyond what the byte-mode register tier (wfa_reg_bytes.cu) takes:
 * gap-affine-2p, cut-offs, reads up to VEC_MAX_LEN.  The pair's upper-cased bytes become 4-bit symbol codes
 * (lv::nib_pack8: wildcard 0, A C G T N R Y K 8..15) and per-base windows of 8 codes in shared memory; the
 * extension is LDS, LDS, XOR, AND (a wildcard position never differs), CLZ per 8 bases.  Recurrence, cut-offs,
 * origin bytes and backtrace are those of the 2-bit tier (wfa_vec.cuh); the replay compares the bytes themselves.
 * A pair holding any other byte, or beyond the tier's capacity, is handed on and ends on the scalar tiers.
 *
 * Its own translation unit: the kernels of wfa_kernels.cu are not recompiled differently because of it.
 */
#define WFA_VEC_BYTES 1
#include <cuda_runtime.h>
#include <stdint.h>

#include "wfa_core.cuh"
#include "wfa_vec.cuh"
#include "wfa_launch.h"

namespace wfagpu {

/* shared memory of one group: [metadata int4 x mr*3][flags 256 B][2 step plans 512 B][windows plen + 1, tlen + 1 | symbol codes][offset rings] */
template <bool TWO_P, bool FULL, int NW, int HEUR>
__global__ void __launch_bounds__(NW == 1 ? 128 : NW * 32, NW == 1 ? 5 : NW == 8 ? 2 : 1) wfa_vecb_kernel(const __grid_constant__ KParams P) {
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
  for (;;) {
    int w = 0;
    if (rank == 0) w = atomicAdd(P.work_counter, 1);
    w = bcast(w);
    if (w >= n_work) break;
    const int pid = P.worklist ? P.worklist[w] : w;
    const PairMeta pm = P.pairs[pid];
    const int plen = pm.plen, tlen = pm.tlen;
    int rc = PAIR_OVERFLOW;
    PairResult res;
    const bool bytes = P.byte_mode || pm.woff < 0;
    const int np_w = (plen >> 3) + 2, nt_w = (tlen >> 3) + 2;      /* words of symbol codes the windows are built from */
    /* (group-uniform condition: every thread takes the same branch and meets the same barriers) */
    if (bytes && plen + tlen + 2 + np_w + nt_w <= P.seq_words_cap && plen <= VEC_MAX_LEN && tlen <= VEC_MAX_LEN) {
      const uint32_t* gp = pm.woff < 0 ? P.words2 + ~pm.woff : P.words + pm.woff;
      const int pbn = (plen + 3) >> 2, tbn = (tlen + 3) >> 2;
      const uint32_t* gt = gp + pbn;
      uint32_t* sp = sm_seq; uint32_t* st = sp + plen + 1;
      uint32_t* cp = st + tlen + 1; uint32_t* ct = cp + np_w;
      bool bad = false;
      for (int j = rank; j < np_w; j += gsize)
        cp[j] = lv::nib_pack8(2 * j < pbn ? gp[2 * j] : 0u, 2 * j + 1 < pbn ? gp[2 * j + 1] : 0u, plen - 8 * j, (uint32_t)P.wildcard, bad);
      for (int j = rank; j < nt_w; j += gsize)
        ct[j] = lv::nib_pack8(2 * j < tbn ? gt[2 * j] : 0u, 2 * j + 1 < tbn ? gt[2 * j + 1] : 0u, tlen - 8 * j, (uint32_t)P.wildcard, bad);
      const bool any_bad = NW == 1 ? (__any_sync(0xffffffffu, bad) != 0) : (__syncthreads_or(bad) != 0);
      vec::gsync<NW>();
      if (!any_bad) {
        /* per-base windows: word i = the 8 symbol codes from position i on, first base in the top bits */
        for (int i = rank; i <= plen; i += gsize) sp[i] = i < plen ? __brev(__funnelshift_r(cp[i >> 3], cp[(i >> 3) + 1], (i & 7) << 2)) : 0u;
        for (int i = rank; i <= tlen; i += gsize) st[i] = i < tlen ? __brev(__funnelshift_r(ct[i >> 3], ct[(i >> 3) + 1], (i & 7) << 2)) : 0u;
        vm.bpw = gp; vm.btw = gt; vm.seqw = 1;
        vm.pw = sp; vm.tw = st;
        vec::gsync<NW>();
        rc = vec::align_pair_vec<TWO_P, FULL, NW, HEUR>(P, vm, plen, tlen, res);
      }
    }
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

/* NW = 1 (warp per pair), 8 or 16 warps per pair; HEUR = 0 none, 1 adaptive, 2 X-drop */
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

cudaError_t launch_vecb(const KParams& P, bool two_p, bool full, int nw, int heur, int grid, int block, size_t smem, cudaStream_t st) {
#define WFA_VEC_LAUNCH(TP, FU, NWW, HH) wfa_vecb_kernel<TP, FU, NWW, HH><<<grid, block, smem, st>>>(P)
  WFA_VEC_DISPATCH(WFA_VEC_LAUNCH);
#undef WFA_VEC_LAUNCH
  return cudaGetLastError();
}

int vecb_occupancy(bool two_p, bool full, int nw, int heur, int block, size_t smem) {
  int nb = 0;
#define WFA_VEC_OCC(TP, FU, NWW, HH) \
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wfa_vecb_kernel<TP, FU, NWW, HH>, block, smem) != cudaSuccess) nb = 0
  WFA_VEC_DISPATCH(WFA_VEC_OCC);
#undef WFA_VEC_OCC
  return nb;
}

cudaError_t init_vecb(int smem_optin) {
  cudaError_t e = cudaSuccess;
  for (int nw : {1, 8, 16})
    for (int heur = 0; heur < 3; ++heur)
      for (int k = 0; k < 4; ++k) {
        const bool two_p = k & 1, full = k & 2;
#define WFA_VEC_INIT(TP, FU, NWW, HH) \
  e = cudaFuncSetAttribute(wfa_vecb_kernel<TP, FU, NWW, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 256)   /* the kernel has a few bytes of static shared memory */
        WFA_VEC_DISPATCH(WFA_VEC_INIT);
#undef WFA_VEC_INIT
        if (e != cudaSuccess) return e;
      }
  return e;
}

}  // namespace wfagpu
