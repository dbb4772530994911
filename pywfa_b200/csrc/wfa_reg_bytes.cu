/*
 * wfa_reg_bytes.cu -- byte mode on the register-resident tier (sm_100a).
 *
 *   wfa_regb_kernel<DX, DOE, FULL>   one warp per pair, 256-diagonal window (wfa_reg.cuh with CB = 4)
 *
 * Read pairs that hold bytes other than ACGT (N, IUPAC codes) or are aligned with pywfa's wildcard=
 * (pywfa/align.pyx:297-304,438-442: the reference switches to wavefront_align_lambda and extends base by base
 * through wavefront_extend_matches_custom, W/wavefront/wavefront_extend_kernels.c:167-203) used to run on the scalar
 * tiers only.  Here the pair's upper-cased bytes (the side buffer of wfa_pack.cu, or the whole batch when the
 * wildcard is itself one of ACGT) are turned into 4-bit symbol codes -- wildcard 0, A C G T N R Y K 8..15
 * (lv::nib_pack8) -- and into per-base windows of 8 codes, and the register tier's extension finds the first
 * differing base of 8 with LDS, LDS, XOR, AND (wildcard positions never differ), CLZ.  The recurrence, the origin
 * codes and the backtrace are those of the 2-bit tier; the replay of the backtrace compares the bytes themselves
 * (extend_offset in byte mode).  A pair holding any other byte, or one the window cannot hold, is handed on
 * (retry list) and ends on the scalar tiers; a 2-bit pair that reaches this tier is handed on untried.
 *
 * Its own translation unit: the kernels of wfa_kernels.cu are not recompiled differently because of it.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "wfa_core.cuh"
#include "wfa_reg.cuh"
#include "wfa_launch.h"

namespace wfagpu {

constexpr int REGB_P = 4;     /* packed registers per wavefront: the 256-diagonal window */

/* shared memory of one warp: [pattern windows plen + 1][text windows tlen + 1][pattern codes (plen >> 3) + 2][text codes (tlen >> 3) + 2] */
template <int DX, int DOE, bool FULL>
__global__ void __launch_bounds__(128, 6) wfa_regb_kernel(const __grid_constant__ KParams K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int P = REGB_P;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp_id = blockIdx.x * (blockDim.x >> 5) + wib;
  uint32_t* const sm_seq = reinterpret_cast<uint32_t*>(smem_raw + (size_t)wib * K.group_bytes);

  RegParams R;
  R.match = K.match; R.g = K.g; R.max_steps = K.max_steps; R.pos_score = K.pos_score;
  R.endsfree = K.endsfree; R.pbf = K.pbf; R.pef = K.pef; R.tbf = K.tbf; R.tef = K.tef;
  R.hrows = K.rhrows; R.opcap = K.ropcap; R.runcap = K.runcap;
  R.kbase = K.reg_kbase; R.c_lo = K.reg_clo; R.c_hi = K.reg_chi;
  uint8_t* const hist_p = FULL ? K.rhist + (long long)warp_id * K.rhist_bytes : nullptr;
  const lv::histref hist = lv::make_histref(hist_p, false);
  uint8_t* const ops = FULL ? K.rops + (long long)warp_id * K.ropcap : nullptr;
  uint32_t* const stage = FULL ? K.runs_stage + (long long)warp_id * K.runcap : nullptr;

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
    int rc = PAIR_OVERFLOW;
    PairResult res;
    const bool bytes = K.byte_mode || pm.woff < 0;
    const int need = plen + tlen + 2 + (plen >> 3) + (tlen >> 3) + 4;
    if (bytes && need <= K.seq_words_cap && plen <= REG_MAX_LEN && tlen <= REG_MAX_LEN) {
      const uint32_t* gp = pm.woff < 0 ? K.words2 + ~pm.woff : K.words + pm.woff;
      const uint32_t* gt = gp + ((plen + 3) >> 2);
      uint32_t* sp = sm_seq; uint32_t* st = sp + plen + 1;
      uint32_t* np = st + tlen + 1; uint32_t* nt = np + (plen >> 3) + 2;
      const bool okp = nibble_words(gp, plen, np, K.wildcard);
      const bool okt = nibble_words(gt, tlen, nt, K.wildcard);
      if (okp && okt) {
        __syncwarp();
        build_windows<4>(np, plen, sp);
        build_windows<4>(nt, tlen, st);
        __syncwarp();
        rc = align_pair_reg<P, DX, DOE, FULL, false, 4>(R, gp, gt, lv::make_seqref(sp), lv::make_seqref(st), plen, tlen, hist, ops,
                                                        stage, lane == 0, res, K.wildcard);
      }
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

/* the penalty shapes of the 2-bit register tier (wfa_kernels.cu: reg_shape) */
static int regb_shape(int dx, int doe, int de, bool full) {
  if (de != 1) return -1;
  if (dx == 2 && doe == 4) return 0;
  if (!full && dx == 1 && doe == 1) return 1;
  if (!full && dx == 2 && doe == 1) return 2;
  if (dx == 4 && doe == 7) return 3;
  if (dx == 1 && doe == 2) return 4;
  if (dx == 1 && doe == 3) return 5;
  return -1;
}

#define WFA_REGB_BOTH(STMT, DX, DOE) do { if (full) { STMT(DX, DOE, true); } else { STMT(DX, DOE, false); } } while (0)
#define WFA_REGB_DISPATCH(STMT)                             \
  do {                                                      \
    if (shape == 0) WFA_REGB_BOTH(STMT, 2, 4);              \
    else if (shape == 1) { STMT(1, 1, false); }             \
    else if (shape == 2) { STMT(2, 1, false); }             \
    else if (shape == 3) WFA_REGB_BOTH(STMT, 4, 7);         \
    else if (shape == 4) WFA_REGB_BOTH(STMT, 1, 2);         \
    else WFA_REGB_BOTH(STMT, 1, 3);                         \
  } while (0)

bool regb_tier_supported(int dx, int doe, int de, bool full) { return regb_shape(dx, doe, de, full) >= 0; }
int regb_regs() { return REGB_P; }

cudaError_t launch_regb(const KParams& P, bool full, int grid, int block, size_t smem, cudaStream_t st) {
  const int shape = regb_shape(P.dx, P.doe1, P.de1, full);
  if (shape < 0) return cudaErrorInvalidValue;
#define WFA_REGB_LAUNCH(DX, DOE, FULL) wfa_regb_kernel<DX, DOE, FULL><<<grid, block, smem, st>>>(P)
  WFA_REGB_DISPATCH(WFA_REGB_LAUNCH);
#undef WFA_REGB_LAUNCH
  return cudaGetLastError();
}

int regb_occupancy(const KParams& P, bool full, int block, size_t smem) {
  int nb = 0;
  const int shape = regb_shape(P.dx, P.doe1, P.de1, full);
  if (shape < 0) return 0;
#define WFA_REGB_OCC(DX, DOE, FULL) \
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wfa_regb_kernel<DX, DOE, FULL>, block, smem) != cudaSuccess) nb = 0
  WFA_REGB_DISPATCH(WFA_REGB_OCC);
#undef WFA_REGB_OCC
  return nb;
}

cudaError_t init_regb(int smem_optin) {
  cudaError_t e = cudaSuccess;
  static const int shapes[10] = {0, 0, 1, 2, 3, 3, 4, 4, 5, 5};
  static const bool fulls[10] = {false, true, false, false, false, true, false, true, false, true};
  for (int v = 0; v < 10; ++v) {
    const int shape = shapes[v];
    const bool full = fulls[v];
#define WFA_REGB_INIT(DX, DOE, FULL) \
  e = cudaFuncSetAttribute(wfa_regb_kernel<DX, DOE, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)
    WFA_REGB_DISPATCH(WFA_REGB_INIT);
#undef WFA_REGB_INIT
    if (e != cudaSuccess) return e;
  }
  return e;
}

}  // namespace wfagpu
