/*
 * wfagpu_api.cpp -- the C ABI of include/wfagpu.h: configuration checks, per-device context,
 * batch staging (raw upload + device-side packing), length buckets, the tiered kernel schedule and
 * result download.
 *
 * Host-side counterpart of what pywfa/align.pyx does around wavefront_align
 * (pywfa/align.pyx:309-443) and of WFA2-lib's aligner lifecycle
 * (W/wavefront/wavefront_aligner.c:387-463), restructured for batches: one configuration POD,
 * one upload of the caller's bases as they are (DMA straight out of pinned memory when the caller
 * provides it), packing and bucketing on the device, kernels, one download.  No host core touches a
 * base, and there is no CPU alignment path in this library.
 */
#include <cuda_runtime.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/wfagpu.h"
#include "pack.h"
#include "wfa_core.cuh"
#include "wfa_launch.h"
#include "wfa_params.h"

using namespace wfagpu;

namespace {

std::atomic<long long> g_dev_reallocs{0};      /* cudaFree + cudaMalloc of a buffer that was too small (trace) */
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); g_dev_reallocs.fetch_add(1, std::memory_order_relaxed); }
    p = nullptr; cap = 0;
    const size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = std::max<size_t>((bytes * 5 / 4 + 4095) & ~(size_t)4095, 4096);
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int MAX_LAUNCH = 160;     /* tier launches of one batch run, over all length buckets (up to ~15 tiers per bucket when byte pairs exist) */
struct DevCounters {       /* one per batch, in HBM */
  int work[MAX_LAUNCH];
  int retry[MAX_LAUNCH];
  int done[MAX_LAUNCH];       /* pairs a tier tried so far / of those, beyond its capacity (adaptive tier skipping) */
  int ovf[MAX_LAUNCH];
  int bucket_n[MAX_BUCKETS];  /* pairs per length bucket: the first tier's work count */
  unsigned long long runs_cursor;
  unsigned long long cells_total;
  unsigned long long dbg[16];   /* WFA_VEC_TIMING builds: cycle counters of the packed-halfword tier */
};

/* what kind of memory a caller's pointer is */
enum MemKind { MEM_PAGEABLE = 0,   /* plain host memory: staged through pinned pieces by the worker threads */
               MEM_DMA = 1,        /* pinned / registered host memory (or another device): the copy engine reads or writes it directly */
               MEM_LOCAL = 2 };    /* memory of this context's device (or managed): used in place */

/* pinned pieces through which pageable input reaches the device */
struct StageRing {
  static constexpr int kSlots = 3;
  static constexpr size_t kPiece = 32u << 20;
  PinBuf buf[kSlots];
  cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr};
  bool busy[kSlots] = {false, false, false};
  int next = 0;
};

/* test / tuning switches, read from the environment once per call (WFAGPU_* variables) */
struct DebugKnobs {
  bool trace = false, no_reg = false, no_regb = false, no_vec = false, no_tier_skip = false, no_buckets = false, host_stage = false, no_metric_map = false;
  int vec_nw = 0, block_threads = 0;
  long long chunk = 0;
};

struct Tier {
  int regs = 0;            /* > 0: register-resident tier (wfa_reg.cuh) with a window of 64*regs diagonals */
  bool bytes = false;      /* ... in byte mode (wfa_reg_bytes.cu): takes the pairs with non-ACGT bytes / the wildcard */
  int vec_nw = 0;          /* > 0: packed-halfword tier (wfa_vec.cuh) with vec_nw warps per pair */
  bool vec_seqw = false;   /* ... with the sequences staged as per-base windows */
  int max_groups = 0;      /* > 0: at most this many pairs in flight (history arena per pair grows accordingly) */
  int mode = 0;            /* 0 warp/smem, 1 block/smem, 2 block/HBM ring */
  int threads = 128;
  int groups_per_block = 4;
  int wcap = 64;
  bool off16 = false;      /* int16 offset rings + history */
  int seq_words_cap = 0;
  int group_bytes = 0;
  size_t smem = 0;
  int blocks_per_sm = 1;
  long long hcap = 0;
  int scap = 0;
};

/* pairs of one length class: planned and run on their own */
struct Bucket {
  long long n = 0;
  int max_len = 0;         /* class limit of max(plen, tlen) */
  int maxp = 0, maxt = 0;  /* longest pattern / text in the bucket */
  int list_base = 0;       /* first slot of its pair ids in the batch's bucket list */
  std::vector<Tier> tiers;
};

}  // namespace

/* chunks of one call in flight: staging (uploads + packing) runs up to kShells - 1 chunks ahead of the
 * alignment kernels, so the copy engine never waits for a kernel to finish */
constexpr int kShells = 4;

struct wfagpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   /* uploads of the next chunks overlap the kernels of this one */
  cudaStream_t pack_stream = nullptr;   /* layout + packing kernels of a chunk, behind its upload; not in the copy stream:
                                           they wait for SM room while the next upload must go on */
  cudaEvent_t copied[4] = {nullptr, nullptr, nullptr, nullptr};   /* the raw bytes of the chunk in slot c % kShells are in HBM */
  cudaEvent_t uploaded[kShells] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t d2h_stream = nullptr;    /* result downloads of chunk c overlap the kernels of chunk c+1 */
  cudaEvent_t d2h_done[kShells] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t run_done = nullptr;       /* all kernels of a chunk (incl. CIGAR ordering) finished */
  PinBuf out_stage[kShells];            /* pinned landing zone of one chunk's result arrays */
  int sms = 0;
  int smem_optin = 0;
  std::string err;
  std::mutex err_mu;                    /* fail() is called from the staging, GPU and drain threads */
  std::mutex call_mu;                   /* one call at a time per context: aligners on several threads may share it */
  DebugKnobs knobs;
  StageRing ring;                       /* pageable input -> pinned pieces -> device */
  DevBuf whole_seq;                     /* the caller's whole byte range, when its pairs are shuffled (every chunk spans it) */
  cudaEvent_t whole_done = nullptr;
  PinBuf gather_seq, gather_off;        /* scattered input gathered back to back (and the offsets it has there) */
  cudaEvent_t gather_done = nullptr;
  bool gather_busy = false;
  PinBuf pin_pack[kShells];             /* PackCounters of the chunk staged in slot c % kShells */
  PinBuf pin_runs, pin_small;
  unsigned char* pair_box = nullptr;    /* mailbox of the single-pair path: mapped pinned memory (wfa_launch.h: PairBox) */
  uint32_t* user_runs = nullptr;        /* caller-provided destination of the CIGAR runs (wfagpu_set_run_buffer) */
  size_t user_runs_cap = 0;             /* ... and its capacity in words */
  std::vector<wfagpu_batch*> spare;
  int64_t last_launches = 0;
  /* per-run scratch shared by all batches of this context (grow-only) */
  DevBuf hist_code, hmeta, runs_stage, gring, rhist, rops, gscratch;
};

struct wfagpu_batch {
  wfagpu_config_t cfg;
  int64_t n = 0;
  int32_t maxp = 0, maxt = 0;
  int64_t total_words = 0;
  bool two_p = false, full = false;
  bool byte_mode = false;       /* every pair is uploaded as bytes (the wildcard is one of ACGT): scalar tiers only */
  int64_t n_side = 0;           /* pairs holding a non-ACGT byte: uploaded as bytes in a side buffer, scalar tiers */
  int64_t total_words2 = 0;
  KParams kp;
  std::vector<Bucket> buckets;
  PackArgs pack;                /* device-side packing arguments of this batch (the side-buffer pass runs later) */
  DevBuf ascii, offs, lens, lay, blist, pcount;      /* raw bases, p_off|t_off, p_len|t_len, layout scan tiles, bucket lists, PackCounters + bucket cursors */
  DevBuf pairs, words, words2, score, status, locs, nruns, runs_base, runs_tmp, retry_a, retry_b, counters,
      cig_off, tile_sums, runs_out;
  unsigned long long runs_tmp_cap = 0, runs_bound = 0;
  int64_t seq_bytes = 0;
  long long total_runs = 0;
  long long cig_base = 0;       /* added to this batch's cig_off values (chunked calls) */
  bool ran = false;
  wfagpu_batch_stats_t stats;
};

namespace {

void batch_release(wfagpu_batch* b);

int fail(wfagpu_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = buf; }
  return code;
}
void set_err(char* err, size_t errlen, const char* fmt, ...) {
  if (!err || !errlen) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, errlen, fmt, ap);
  va_end(ap);
}

#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? WFAGPU_ENOMEM : WFAGPU_ECUDA,      \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

long long score_bound(const KParams& k, long long maxp, long long maxt) {
  /* Bound of the optimum of EVERY pair with plen <= maxp, tlen <= maxt (the batch maxima come from
   * different pairs: a 305 x 61 pair in a batch whose longest text has 320 bases needs 244 gap
   * extensions -- found by the fuzzer).  Any alignment bounds the optimum; both candidates below are
   * monotone in the lengths: "mismatch the shorter sequence, one gap for the rest" at its worst
   * (shorter = min of the maxima, rest = the longer maximum) and "delete everything, insert everything". */
  const long long one_gap = (long long)k.x * std::min(maxp, maxt) + k.o1 + (long long)k.e1 * std::max(maxp, maxt);
  const long long gaps = (maxp ? k.o1 + (long long)k.e1 * maxp : 0) + (maxt ? k.o1 + (long long)k.e1 * maxt : 0);
  return k.no_mis ? gaps : std::min(one_gap, gaps);      /* indel: no mismatches to pay with */
}

int pow2_floor(long long v) { int p = 1; while (2ll * p <= v) p <<= 1; return p; }
int pow2_ceil(long long v) { int p = 1; while (p < v) p <<= 1; return p; }

size_t group_bytes_of(const KParams& k, bool two_p, int seqw, int wcap, int elem) {
  const int ns = k.rm + 2 * k.r1 + (two_p ? 2 * k.r2 : 0);
  const int nc = two_p ? 5 : 3;
  const size_t bytes = (size_t)k.mr * nc * 16 + 4ull * seqw + (size_t)ns * wcap * elem;
  return (bytes + 15) & ~(size_t)15;
}

void plan_tiers(wfagpu_ctx* ctx, wfagpu_batch* b, Bucket& bk) {
  const KParams& k = b->kp;
  const DebugKnobs& dbg = ctx->knobs;
  bk.tiers.clear();
  struct { int maxp, maxt; bool byte_mode, two_p, full; int64_t n_side; std::vector<Tier>& tiers; } B{bk.maxp, bk.maxt, b->byte_mode, b->two_p, b->full, b->n_side, bk.tiers};
  /* words of one pair's sequences: 2-bit codes for the fast tiers, bytes where a scalar tier may meet a byte-mode pair */
  const bool any_bytes = B.byte_mode || B.n_side > 0;
  const int seqw2 = (B.maxp + 15) / 16 + (B.maxt + 15) / 16 + 2;
  const int seqw = any_bytes ? (B.maxp + 3) / 4 + (B.maxt + 3) / 4 + 2 : seqw2;
  /* widest computed wavefront: the DP matrix has plen + tlen + 1 diagonals and a step's range reaches one
   * diagonal beyond either side before it is trimmed (compute.c:40-86 on trimmed sources) */
  const long long wmax = (long long)B.maxp + B.maxt + 3;
  const int wmax2 = pow2_ceil(std::max<long long>(wmax, 32));
  /* capacity of the score tables: the optimum is bounded by any alignment; with a cut-off the path is
   * whatever survives the pruning (fuzzing found X-drop alignments at twice the optimum's bound), so
   * only the trivial bound "every column pays the dearest operation" holds */
  const long long dearest = std::max<long long>(k.x, std::max<long long>(k.o1 + k.e1, B.two_p ? k.o2 + k.e2 : 0));
  const long long sb_any = ((long long)B.maxp + B.maxt) * dearest + dearest;
  const long long sb = std::min<long long>(k.heuristic ? sb_any : score_bound(k, B.maxp, B.maxt), k.max_steps);
  const long long scap_bound = std::min<long long>(sb / k.g + k.rm + 4, INT_MAX / 4);
  const long long cells_bound = std::min<long long>(scap_bound * wmax, (long long)4e18 / 8);
  const int smem_max = ctx->smem_optin;
  /* int16 rings hold offsets up to ~tlen + width and nulls that drift by one per step */
  const bool short_reads = 2ll * ((long long)B.maxp + B.maxt) + 4096 < 30000;
  /* register-resident tiers first: gap-affine, no heuristic, instantiated penalty shape, short reads */
  const bool no_reg = dbg.no_reg;
  const int winw = B.maxp + B.maxt + 2;       /* sequence windows: one word per base */
  if (!no_reg && !B.byte_mode && !B.two_p && !k.m_only && k.heuristic == 0 && std::max(B.maxp, B.maxt) <= REG_MAX_LEN && 4 * winw <= 8192) {
    const int maxlen = std::max(B.maxp, B.maxt);
    /* window the typical pair of this length needs: tuned for the default shape (x/gcd = 2); a wavefront
     * spreads one diagonal per score unit either side, and a mismatch costs x/gcd units, so other shapes scale */
    int first = maxlen <= 192 ? 2 : maxlen <= 320 ? 3 : 4;
    if (k.dx != 2) first = std::max(2, std::min(4, (int)((0.32 * maxlen * k.dx + 63.0) / 64.0)));
    for (int regs = first; regs <= 4; ++regs) {
      if (regs == 3 && first == 2) continue;                         /* 128 -> 256 directly: few pairs get that far */
      if (!reg_tier_supported(k.dx, k.doe1, k.de1, regs, B.full)) continue;
      Tier t;
      t.regs = regs; t.mode = 0; t.threads = 128; t.groups_per_block = 4; t.wcap = 64 * regs;
      t.scap = 32 * regs + k.doe1 + 1;          /* origin rows: scores the window can hold */
      /* per warp: sequence windows; scope=full with the 128-diagonal window: packed sequences, edit-operation
       * stack and origin arena as well (wfa_reg.cuh: RegSmem) */
      const int ropcap = (int)(((long long)B.maxp + B.maxt + 8 + 15) & ~15ll);
      const RegSmem L = reg_smem_layout(regs, B.full, winw, ropcap, t.scap);
      if (reg_hist_in_smem(regs, B.full) && 4 * (B.maxp + B.maxt + 2) > L.hist_bytes) continue;   /* run staging must fit in the arena */
      t.seq_words_cap = winw; t.group_bytes = L.total(); t.smem = (size_t)t.group_bytes * 4;
      B.tiers.push_back(t);
    }
  }
  /* ... and the same tier in byte mode for the pairs the 2-bit tiers hand on untried (non-ACGT bytes, the wildcard):
   * 4-bit symbol codes, 8 bases per window word; what it cannot take (other bytes than ACGTNRYK and the wildcard,
   * wavefronts beyond 256 diagonals) goes on to the scalar tiers */
  if (!no_reg && !dbg.no_regb && any_bytes && !B.two_p && !k.m_only && k.heuristic == 0 && std::max(B.maxp, B.maxt) <= REG_MAX_LEN &&
      4 * winw <= 8192 && regb_tier_supported(k.dx, k.doe1, k.de1, B.full)) {
    Tier t;
    t.regs = regb_regs(); t.bytes = true; t.mode = 0; t.threads = 128; t.groups_per_block = 4; t.wcap = 64 * t.regs;
    t.scap = 32 * t.regs + k.doe1 + 1;
    t.seq_words_cap = winw + (B.maxp >> 3) + (B.maxt >> 3) + 4;       /* windows + the symbol codes they are built from */
    t.group_bytes = (4 * t.seq_words_cap + 15) & ~15; t.smem = (size_t)t.group_bytes * 4;
    B.tiers.push_back(t);
  }
  /* packed-halfword tiers (wfa_vec.cuh): everything the register tier does not take, reads up to
   * VEC_MAX_LEN; warp per pair first, then 8 and 16 warps per pair with the widest rings that
   * leave two / one CTA per SM */
  const bool no_vec = dbg.no_vec;
  bool vec_covers_smem = false;
  const bool use_vec = !no_vec && !B.byte_mode && !k.m_only && std::max(B.maxp, B.maxt) <= VEC_MAX_LEN;   /* byte mode, gap-linear / edit / indel: scalar tiers */
  const int nslots = k.rm + 2 * k.r1 + (B.two_p ? 2 * k.r2 : 0) + 1;      /* + the all-null slot */
  const long long nblk_max = (wmax + 63) / 64 + 1;
  long long last_nblk = 0;
  /* bytes: the byte-mode kernels (wfa_vec_bytes.cu): always windows (8 symbol codes per word), plus the codes they are built from */
  auto add_vec = [&](int nw, long long budget, long long hcap, long long scap, bool bytes) {
      /* per-base sequence windows (one word per base: 16 bases = LDS, LDS, XOR, CLZ) where they are small
       * next to the rings, 2-bit packed words otherwise */
      const long long winw = (long long)B.maxp + B.maxt + 2 + (bytes ? (B.maxp >> 3) + (B.maxt >> 3) + 4 : 0);
      const bool seqw_t = bytes || (nw > 1 && 4 * winw <= 16384);          /* (the one-warp kernel is compiled without the window variant) */
      const long long fixed = (long long)k.mr * 48 + 256 + 4 * (seqw_t ? winw : seqw2) + 512;   /* + two step plans */
      long long nblk = (budget - fixed) / ((long long)nslots * 128);
      nblk = std::min(nblk, nblk_max);
      if (nblk < 2 || nblk <= last_nblk) return;
      Tier t;
      t.vec_nw = nw; t.mode = nw == 1 ? 0 : 1; t.bytes = bytes;
      t.threads = nw == 1 ? 128 : nw * 32; t.groups_per_block = nw == 1 ? 4 : 1;
      t.wcap = (int)(64 * nblk); t.seq_words_cap = (int)(seqw_t ? winw : seqw2); t.vec_seqw = seqw_t;
      t.group_bytes = (int)((fixed + nblk * nslots * 128 + 15) & ~15ll);
      t.smem = (size_t)t.group_bytes * t.groups_per_block;
      t.scap = (int)std::min<long long>(scap, scap_bound);
      t.hcap = std::min<long long>(hcap, (long long)t.scap * (t.wcap + 64));
      t.hcap = (t.hcap + 63) & ~63ll;
      B.tiers.push_back(t);
      last_nblk = nblk;
      if (nblk == nblk_max && !bytes) vec_covers_smem = true;
  };
  auto add_vecs = [&](bool bytes) {
    last_nblk = 0;
    const int only = dbg.vec_nw;                          /* tests: push everything through one group size */
    if (!only || only == 1) {
      add_vec(1, 11264 + 512, 2ll << 20, 16384, bytes);
      if (last_nblk == 0) add_vec(1, 22528 + 512, 2ll << 20, 16384, bytes);
    }
    if (!only || only == 8) add_vec(8, (smem_max - 2048) / 2, 16ll << 20, 32768, bytes);
    if (!only || only == 16) add_vec(16, smem_max - 1024, 64ll << 20, 65536, bytes);
  };
  if (use_vec) add_vecs(false);
  /* ... and in byte mode, for the pairs the 2-bit tiers hand on untried; what these cannot take (other bytes than
   * ACGTNRYK and the wildcard, wavefronts beyond the rings) ends on the scalar tiers below */
  if (!no_vec && !dbg.no_regb && any_bytes && !k.m_only && std::max(B.maxp, B.maxt) <= VEC_MAX_LEN) add_vecs(true);
  int last_wcap = 0;
  auto add_warp = [&](int wcap, long long hcap, int scap) {
    if ((vec_covers_smem || use_vec) && !any_bytes) return;   /* the vec tiers replace the scalar shared-memory tiers (byte-mode pairs still need them) */
    wcap = std::min(wcap, wmax2);
    if (wcap <= last_wcap) return;
    Tier t;
    t.mode = 0; t.threads = 128; t.groups_per_block = 4; t.wcap = wcap; t.seq_words_cap = seqw;
    t.scap = (int)std::min<long long>(scap, scap_bound);
    t.off16 = short_reads && t.scap <= 16384 && wcap <= 2048;
    t.group_bytes = (int)group_bytes_of(k, B.two_p, seqw, wcap, t.off16 ? 2 : 4);
    if (t.group_bytes > 48 * 1024) return;
    t.smem = (size_t)t.group_bytes * 4;
    t.hcap = std::min(hcap, cells_bound);
    B.tiers.push_back(t);
    last_wcap = wcap;
  };
  /* first warp tier: the widest power of two that still leaves >= 32 resident warps per SM */
  int w0 = 32;
  while (w0 < 1024 && group_bytes_of(k, B.two_p, seqw, w0 * 2, short_reads ? 2 : 4) <= 6912) w0 *= 2;
  add_warp(w0, 16384, 1024);
  add_warp(w0 * 4, 131072, 8192);
  {
    const long long avail = (long long)smem_max - 1024 - (long long)block_reduce_smem_bytes() -
                            (long long)group_bytes_of(k, B.two_p, seqw, 0, 4);
    const int ns = k.rm + 2 * k.r1 + (B.two_p ? 2 * k.r2 : 0);
    /* int16 rings double the width that fits in shared memory (1 kbp gap-affine-2p pairs need
     * ~1100 diagonals x 36 ring slots) */
    const long long elem = short_reads ? 2 : 4;
    int wcap = avail > elem * ns * 32 ? pow2_floor(avail / (elem * ns)) : 0;
    wcap = std::min(wcap, wmax2);
    const bool vec_has = use_vec && !any_bytes;
    if (!vec_has && (wcap > last_wcap || (B.full && wcap >= 32 && wcap == wmax2))) {
      Tier t;
      t.mode = 1; t.threads = wcap > 1024 ? 512 : 256; t.groups_per_block = 1; t.wcap = wcap;
      if (dbg.block_threads) t.threads = dbg.block_threads;       /* tuning experiments */
      t.seq_words_cap = seqw;
      t.off16 = short_reads;
      t.group_bytes = (int)group_bytes_of(k, B.two_p, seqw, wcap, (int)elem);
      t.smem = (size_t)t.group_bytes + block_reduce_smem_bytes();
      t.hcap = std::min<long long>(8ll << 20, cells_bound);
      t.scap = (int)std::min<long long>(t.off16 ? 16384 : 1 << 16, scap_bound);
      B.tiers.push_back(t);
      last_wcap = wcap;
    }
  }
  {
    /* widest tier: ring in HBM, history sized at run time from free memory */
    Tier t;
    t.mode = 2; t.threads = 512; t.groups_per_block = 1; t.wcap = wmax2;
    t.seq_words_cap = (4ll * seqw <= 160 * 1024) ? seqw : 0;
    t.group_bytes = (int)group_bytes_of(k, B.two_p, t.seq_words_cap, 0, 4);
    t.smem = (size_t)t.group_bytes + block_reduce_smem_bytes();
    t.hcap = cells_bound;
    t.scap = (int)std::min<long long>(2 * scap_bound, INT_MAX / 4);
    B.tiers.push_back(t);
    if (B.full) {
      /* scope=full: the origin bytes of a 100 kbp pair need ~10 GB; pairs whose history outgrows
       * an even split of the free HBM over the pairs in flight are redone with fewer neighbours */
      t.max_groups = 4; B.tiers.push_back(t);
      t.max_groups = 1; B.tiers.push_back(t);
    }
  }
  for (auto& t : B.tiers) {
    int bps = t.regs && t.bytes ? regb_occupancy(k, B.full, t.threads, t.smem)
              : t.regs ? reg_occupancy(k, t.regs, B.full, t.threads, t.smem)
              : t.vec_nw && t.bytes ? vecb_occupancy(B.two_p, B.full, t.vec_nw, k.heuristic, t.threads, t.smem)
              : t.mode == 2 ? grid_occupancy(B.two_p, B.full, t.smem)
              : t.vec_nw ? vec_occupancy(B.two_p, B.full, t.vec_nw, k.heuristic, t.threads, t.smem)
                         : align_occupancy(B.two_p, B.full, t.mode, t.off16, t.threads, t.smem);
    t.blocks_per_sm = std::max(1, bps);
  }
}

}  // namespace

/* ---- configuration ------------------------------------------------------------------- */
extern "C" void wfagpu_config_default(wfagpu_config_t* cfg) {
  /* pywfa/align.pyx:309-334 */
  memset(cfg, 0, sizeof *cfg);
  cfg->distance = WFAGPU_DISTANCE_AFFINE;
  cfg->scope = WFAGPU_SCOPE_FULL;
  cfg->span = WFAGPU_SPAN_ENDSFREE;
  cfg->heuristic = WFAGPU_HEURISTIC_NONE;
  cfg->min_wavefront_length = 10;
  cfg->max_distance_threshold = 50;
  cfg->steps_between_cutoffs = 1;
  cfg->xdrop = 20;
  cfg->match = 0; cfg->mismatch = 4;
  cfg->gap_opening1 = 6; cfg->gap_extension1 = 2;
  cfg->gap_opening2 = 24; cfg->gap_extension2 = 1;
  cfg->max_steps = 0;
}

extern "C" int wfagpu_config_check(const wfagpu_config_t* c, int64_t plen, int64_t tlen, char* err, size_t errlen) {
  if (!c) { set_err(err, errlen, "null configuration"); return WFAGPU_EINVAL; }
  if (c->distance < WFAGPU_DISTANCE_AFFINE || c->distance > WFAGPU_DISTANCE_INDEL) {
    set_err(err, errlen, "bad distance %d", c->distance);
    return WFAGPU_EINVAL;
  }
  const bool edit_like = c->distance == WFAGPU_DISTANCE_EDIT || c->distance == WFAGPU_DISTANCE_INDEL;
  if (c->scope != WFAGPU_SCOPE_SCORE && c->scope != WFAGPU_SCOPE_FULL) { set_err(err, errlen, "bad scope %d", c->scope); return WFAGPU_EINVAL; }
  if (c->span != WFAGPU_SPAN_END2END && c->span != WFAGPU_SPAN_ENDSFREE) { set_err(err, errlen, "bad span %d", c->span); return WFAGPU_EINVAL; }
  if (c->heuristic < WFAGPU_HEURISTIC_NONE || c->heuristic > WFAGPU_HEURISTIC_XDROP) { set_err(err, errlen, "bad heuristic %d", c->heuristic); return WFAGPU_EINVAL; }
  /* wavefront_penalties_set_affine/_affine2p, W/wavefront/wavefront_penalties.c:95-173 */
  if (c->wildcard < 0 || c->wildcard > 255) { set_err(err, errlen, "wildcard must be 0 (none) or one byte"); return WFAGPU_EINVAL; }
  if (edit_like) {
    /* wavefront_penalties_set_edit / _indel (penalties.c:38-61) ignore every penalty field;
     * the drop heuristics are refused with these metrics (wavefront_align_presets__checks, W/wavefront/wavefront_align.c:82-89) */
    if (c->heuristic == WFAGPU_HEURISTIC_XDROP) {
      set_err(err, errlen, "[WFA] Heuristics drops are not compatible with 'edit'/'indel' distance metrics");
      return WFAGPU_EINVAL;
    }
  } else {
    if (c->match > 0) { set_err(err, errlen, "[WFA::Penalties] Match score must be negative or zero (M=%d)", c->match); return WFAGPU_EINVAL; }
    if (c->distance == WFAGPU_DISTANCE_LINEAR) {
      /* wavefront_penalties_set_linear, penalties.c:62-93; the indel penalty travels in gap_extension1 (align.pyx:355) */
      if (c->mismatch <= 0 || c->gap_extension1 <= 0) {
        set_err(err, errlen, "[WFA::Penalties] Penalties (X=%d,D=%d,I=%d) must be (X>0,D>0,I>0)", c->mismatch, c->gap_extension1, c->gap_extension1);
        return WFAGPU_EINVAL;
      }
    } else if (c->mismatch <= 0 || c->gap_opening1 < 0 || c->gap_extension1 <= 0) {
      set_err(err, errlen, "[WFA::Penalties] Penalties (X=%d,O=%d,E=%d) must be (X>0,O>=0,E>0)", c->mismatch, c->gap_opening1, c->gap_extension1);
      return WFAGPU_EINVAL;
    }
  }
  if (c->distance == WFAGPU_DISTANCE_AFFINE2P && (c->gap_opening2 < 0 || c->gap_extension2 <= 0)) {
    set_err(err, errlen, "[WFA::Penalties] Penalties (X=%d,O1=%d,E1=%d,O2=%d,E2=%d) must be (X>0,O1>=0,E1>0,O1>=0,E1>0)",
            c->mismatch, c->gap_opening1, c->gap_extension1, c->gap_opening2, c->gap_extension2);
    return WFAGPU_EINVAL;
  }
  if (!edit_like && c->match < 0 && c->span == WFAGPU_SPAN_ENDSFREE && (c->pattern_begin_free > 0 || c->text_begin_free > 0)) {
    set_err(err, errlen, "match < 0 with begin-free ends is outside the accelerated path");
    return WFAGPU_EUNSUPPORTED;
  }
  if (c->span == WFAGPU_SPAN_ENDSFREE) {
    if (c->pattern_begin_free < 0 || c->pattern_end_free < 0 || c->text_begin_free < 0 || c->text_end_free < 0) {
      set_err(err, errlen, "ends-free parameters must be non-negative"); return WFAGPU_EINVAL;
    }
    /* wavefront_align_presets__checks, W/wavefront/wavefront_align.c:89-100 */
    if (plen >= 0 && tlen >= 0 &&
        (c->pattern_begin_free > plen || c->pattern_end_free > plen || c->text_begin_free > tlen || c->text_end_free > tlen)) {
      set_err(err, errlen, "[WFA] Ends-free parameters must be not larger than the sequences (P0=%d,Pf=%d,T0=%d,Tf=%d). "
              "Must be (P0<=|P|,Pf<=|P|,T0<=|T|,Tf<=|T|) where (|P|,|T|)=(%lld,%lld)",
              c->pattern_begin_free, c->pattern_end_free, c->text_begin_free, c->text_end_free, (long long)plen, (long long)tlen);
      return WFAGPU_EINVAL;
    }
  }
  /* bounded so that penalties * ring slots stay in int range */
  if (c->mismatch > (1 << 20) || c->gap_opening1 > (1 << 20) || c->gap_extension1 > (1 << 20) ||
      c->gap_opening2 > (1 << 20) || c->gap_extension2 > (1 << 20) || c->match < -(1 << 20)) {
    set_err(err, errlen, "penalties above 2^20 are not supported"); return WFAGPU_EUNSUPPORTED;
  }
  return WFAGPU_OK;
}

extern "C" const char* wfagpu_strerror(int code) {
  switch (code) {
    case WFAGPU_OK: return "ok";
    case WFAGPU_EINVAL: return "invalid argument or configuration";
    case WFAGPU_ECUDA: return "CUDA runtime error";
    case WFAGPU_ENOMEM: return "out of memory";
    case WFAGPU_ENODEVICE: return "no CUDA device (this library has no CPU fallback)";
    case WFAGPU_EUNSUPPORTED: return "input outside the accelerated path";
    default: return "unknown error";
  }
}

/* ---- context ------------------------------------------------------------------------- */
extern "C" int wfagpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int wfagpu_create(wfagpu_ctx** out, int device, char* err, size_t errlen) {
  if (!out) return WFAGPU_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_err(err, errlen, "no CUDA device visible (%s); wfagpu has no CPU fallback",
            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return WFAGPU_ENODEVICE;
  }
  if (device < 0 || device >= n) { set_err(err, errlen, "device %d out of range (0..%d)", device, n - 1); return WFAGPU_EINVAL; }
  wfagpu_ctx* ctx = new wfagpu_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  int prio_lo = 0, prio_hi = 0;
  /* every failure leaves through wfagpu_destroy, which releases whatever exists by then */
  auto bail = [&](int code) { wfagpu_destroy(ctx); return code; };
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    set_err(err, errlen, "CUDA init failed: %s", cudaGetErrorString(e));
    return bail(WFAGPU_ECUDA);
  }
  if (prop.major < 10) {
    set_err(err, errlen, "device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor);
    return bail(WFAGPU_ENODEVICE);
  }
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  /* the staging stream outranks the alignment stream: the packing kernels of chunk c+1 take the SM
   * slots that the persistent alignment kernel of chunk c frees in its tail */
  if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_lo)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&ctx->pack_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&ctx->d2h_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->run_done, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->gather_done, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->whole_done, cudaEventDisableTiming)) != cudaSuccess) {
    set_err(err, errlen, "CUDA init failed: %s", cudaGetErrorString(e));
    return bail(WFAGPU_ECUDA);
  }
  for (int i = 0; i < kShells; ++i)
    if ((e = cudaEventCreateWithFlags(&ctx->copied[i], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->uploaded[i], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->d2h_done[i], cudaEventDisableTiming)) != cudaSuccess) {
      set_err(err, errlen, "CUDA init failed: %s", cudaGetErrorString(e));
      return bail(WFAGPU_ECUDA);
    }
  for (int i = 0; i < StageRing::kSlots; ++i)
    if ((e = cudaEventCreateWithFlags(&ctx->ring.ev[i], cudaEventDisableTiming)) != cudaSuccess) {
      set_err(err, errlen, "CUDA init failed: %s", cudaGetErrorString(e));
      return bail(WFAGPU_ECUDA);
    }
  ctx->sms = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((e = init_kernels(ctx->smem_optin)) != cudaSuccess) {
    set_err(err, errlen, "kernel setup failed: %s", cudaGetErrorString(e));
    return bail(WFAGPU_ECUDA);
  }
  *out = ctx;
  return WFAGPU_OK;
}

extern "C" void wfagpu_destroy(wfagpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& pb : ctx->ring.buf) pb.release();
  if (ctx->pair_box) cudaFreeHost(ctx->pair_box);
  ctx->gather_seq.release(); ctx->gather_off.release();
  for (auto& pb : ctx->pin_pack) pb.release();
  ctx->pin_runs.release(); ctx->pin_small.release();
  for (wfagpu_batch* b : ctx->spare) batch_release(b);
  ctx->spare.clear();
  ctx->hist_code.release(); ctx->hmeta.release(); ctx->runs_stage.release(); ctx->gring.release();
  ctx->rhist.release(); ctx->rops.release(); ctx->gscratch.release();
  for (cudaStream_t s : {ctx->stream, ctx->copy_stream, ctx->pack_stream, ctx->d2h_stream}) if (s) cudaStreamDestroy(s);
  ctx->whole_seq.release();
  for (cudaEvent_t ev : {ctx->run_done, ctx->gather_done, ctx->whole_done, ctx->ring.ev[0], ctx->ring.ev[1], ctx->ring.ev[2]})
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < kShells; ++i) {
    if (ctx->uploaded[i]) cudaEventDestroy(ctx->uploaded[i]);
    if (ctx->copied[i]) cudaEventDestroy(ctx->copied[i]);
    if (ctx->d2h_done[i]) cudaEventDestroy(ctx->d2h_done[i]);
    ctx->out_stage[i].release();
  }
  cudaGetLastError();
  delete ctx;
}

extern "C" const char* wfagpu_last_error(const wfagpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

/* ---- caller-side pinned memory ---------------------------------------------------------- */
extern "C" void* wfagpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void wfagpu_host_free(void* p) { if (p) { cudaFreeHost(p); cudaGetLastError(); } }
extern "C" int wfagpu_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return WFAGPU_EINVAL;
  if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return WFAGPU_ECUDA; }
  return WFAGPU_OK;
}
extern "C" int wfagpu_host_unregister(void* p) {
  if (!p) return WFAGPU_EINVAL;
  if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return WFAGPU_ECUDA; }
  return WFAGPU_OK;
}

/* host copy of a result array out of the pinned landing zone: a few threads for large arrays (one
 * core moves ~8 GB/s; the last chunk's copy is not overlapped by anything) */
static void par_memcpy(void* dst, const void* src, size_t bytes) {
  const size_t kMin = 1u << 20;
  if (bytes < 2 * kMin) { memcpy(dst, src, bytes); return; }
  const int parts = (int)std::min<size_t>(std::min(4, host_threads()), bytes / kMin);
  if (parts <= 1) { memcpy(dst, src, bytes); return; }
  const size_t step = ((bytes / parts) + 63) & ~(size_t)63;
  std::vector<std::thread> th;
  for (int i = 1; i < parts; ++i) {
    const size_t off = step * i, len = std::min(step, bytes - std::min(bytes, off));
    if (len) th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
  }
  memcpy(dst, src, std::min(step, bytes));
  for (auto& t : th) t.join();
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* ---- batches ------------------------------------------------------------------------- */
namespace {

void read_knobs(wfagpu_ctx* ctx) {
  DebugKnobs k;
  auto flag = [](const char* name) { return getenv(name) != nullptr; };
  auto num = [](const char* name) { const char* e = getenv(name); return e ? atoll(e) : 0ll; };
  k.trace = flag("WFAGPU_TRACE");
  k.no_reg = flag("WFAGPU_NO_REG_TIER");
  k.no_regb = flag("WFAGPU_NO_REG_BYTES");
  k.no_vec = flag("WFAGPU_NO_VEC_TIER");
  k.no_tier_skip = flag("WFAGPU_NO_TIER_SKIP");
  k.no_buckets = flag("WFAGPU_NO_BUCKETS");
  k.no_metric_map = flag("WFAGPU_NO_METRIC_MAP");
  k.vec_nw = (int)num("WFAGPU_VEC_NW");
  k.block_threads = (int)num("WFAGPU_BLOCK_THREADS");
  k.chunk = num("WFAGPU_CHUNK");
  ctx->knobs = k;
}

MemKind mem_kind(const wfagpu_ctx* ctx, const void* p) {
  if (!p) return MEM_PAGEABLE;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return MEM_PAGEABLE; }
  if (a.type == cudaMemoryTypeHost) return MEM_DMA;
  if (a.type == cudaMemoryTypeManaged) return MEM_LOCAL;
  if (a.type == cudaMemoryTypeDevice) return a.device == ctx->device ? MEM_LOCAL : MEM_DMA;
  return MEM_PAGEABLE;
}

void batch_release(wfagpu_batch* b) {
  for (DevBuf* d : {&b->ascii, &b->offs, &b->lens, &b->lay, &b->blist, &b->pcount,
                    &b->pairs, &b->words, &b->words2, &b->score, &b->status, &b->locs, &b->nruns, &b->runs_base, &b->runs_tmp,
                    &b->retry_a, &b->retry_b, &b->counters, &b->cig_off, &b->tile_sums, &b->runs_out})
    d->release();
  delete b;
}

/* batch shells keep their device buffers (grow-only) and are recycled through the context */
wfagpu_batch* batch_acquire(wfagpu_ctx* ctx) {
  wfagpu_batch* b;
  if (!ctx->spare.empty()) { b = ctx->spare.back(); ctx->spare.pop_back(); }
  else b = new wfagpu_batch();
  b->buckets.clear();
  b->ran = false; b->total_runs = 0; b->runs_tmp_cap = 0;
  memset(&b->stats, 0, sizeof b->stats);
  memset(&b->kp, 0, sizeof b->kp);
  return b;
}
void batch_recycle(wfagpu_ctx* ctx, wfagpu_batch* b) {
  if (ctx && ctx->spare.size() < kShells + 1) ctx->spare.push_back(b);
  else batch_release(b);
}

/* Host -> device copy of `bytes` on stream st.  Pinned (or peer) memory goes by DMA; pageable memory
 * is copied by the worker threads into a ring of pinned pieces, each sent off as soon as it is full. */
int upload(wfagpu_ctx* ctx, void* dst, const void* src, size_t bytes, MemKind kind, cudaStream_t st) {
  if (!bytes) return WFAGPU_OK;
  if (kind != MEM_PAGEABLE) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return WFAGPU_OK;
  }
  StageRing& r = ctx->ring;
  for (size_t off = 0; off < bytes; off += StageRing::kPiece) {
    const size_t len = std::min(StageRing::kPiece, bytes - off);
    const int slot = r.next;
    r.next = (r.next + 1) % StageRing::kSlots;
    if (r.busy[slot]) { CK(cudaEventSynchronize(r.ev[slot])); r.busy[slot] = false; }
    CK(r.buf[slot].ensure(std::min(StageRing::kPiece, std::max<size_t>(len, 1u << 16))));
    parallel_copy(r.buf[slot].p, (const char*)src + off, len);
    CK(cudaMemcpyAsync((char*)dst + off, r.buf[slot].p, len, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(r.ev[slot], st));
    r.busy[slot] = true;
  }
  return WFAGPU_OK;
}

/* the caller's arrays of one call and what kind of memory each lives in */
struct Inputs {
  const uint8_t* seq; const int64_t* p_off; const int32_t* p_len; const int64_t* t_off; const int32_t* t_len;
  MemKind k_seq, k_poff, k_plen, k_toff, k_tlen;
};

int classify_inputs(wfagpu_ctx* ctx, Inputs& in) {
  in.k_seq = mem_kind(ctx, in.seq);
  in.k_poff = mem_kind(ctx, in.p_off); in.k_plen = mem_kind(ctx, in.p_len);
  in.k_toff = mem_kind(ctx, in.t_off); in.k_tlen = mem_kind(ctx, in.t_len);
  /* the planner reads the offset / length arrays on the host */
  if (in.k_poff == MEM_LOCAL || in.k_plen == MEM_LOCAL || in.k_toff == MEM_LOCAL || in.k_tlen == MEM_LOCAL)
    return fail(ctx, WFAGPU_EINVAL, "offset / length arrays must be host memory (only the sequence buffer may live on the device)");
  return WFAGPU_OK;
}

/*
 * Stage one batch (or chunk): host pass over the offset / length arrays, upload of the raw bases
 * and arrays on stream st, layout scan + packing kernels behind them, and the D2H of the packer's
 * counters into `counts` (pinned).  Pairs [first, first + n) of the caller's arrays.  No stream
 * synchronisation: batch_finish_stage completes the batch once the stream got there.
 */
int batch_stage(wfagpu_ctx* ctx, wfagpu_batch* b, const wfagpu_config_t* cfg, const Inputs& in, int64_t first, int64_t n,
                cudaStream_t st, cudaStream_t pk, cudaEvent_t copied, PackCounters* counts) {
  b->cfg = *cfg; b->n = n;
  b->two_p = cfg->distance == WFAGPU_DISTANCE_AFFINE2P;
  b->full = cfg->scope == WFAGPU_SCOPE_FULL;
  b->stats.n_pairs = n;
  b->n_side = 0; b->total_words2 = 0;
  const int64_t* p_off = in.p_off + first; const int64_t* t_off = in.t_off + first;
  const int32_t* p_len = in.p_len + first; const int32_t* t_len = in.t_len + first;
  /* 2-bit codes unless the wildcard is itself one of ACGT (then every pair is staged as bytes).
   * Pairs that hold any other byte are found by the packing kernel, staged a second time as bytes in
   * a small side buffer and routed to the scalar tiers (byte mode: 4 bases per word, the extension
   * honours the wildcard); the rest of the batch stays on the fast path. */
  const int wc = cfg->wildcard & 0xff;
  b->byte_mode = wc == 'A' || wc == 'C' || wc == 'G' || wc == 'T';
  const int bpw = b->byte_mode ? 4 : 16;
  PairScan hs;
  scan_pairs(p_off, p_len, t_off, t_len, n, bpw, &hs);
  if (hs.first_negative >= 0) return fail(ctx, WFAGPU_EINVAL, "negative length at pair %lld", (long long)(first + hs.first_negative));
  if (n && cfg->span == WFAGPU_SPAN_ENDSFREE &&
      (cfg->pattern_begin_free > hs.minp || cfg->pattern_end_free > hs.minp || cfg->text_begin_free > hs.mint || cfg->text_end_free > hs.mint)) {
    /* wavefront_align_presets__checks, W/wavefront/wavefront_align.c:89-100: name the first offender */
    char msg[400];
    for (int64_t i = 0; i < n; ++i) {
      const int rc = wfagpu_config_check(cfg, p_len[i], t_len[i], msg, sizeof msg);
      if (rc != WFAGPU_OK) return fail(ctx, rc, "pair %lld: %s", (long long)(first + i), msg);
    }
  }
  if ((long long)hs.maxp + hs.maxt > (1ll << 27)) return fail(ctx, WFAGPU_EUNSUPPORTED, "sequences longer than 2^27 bases");
  b->maxp = hs.maxp; b->maxt = hs.maxt; b->seq_bytes = hs.seq_bytes; b->total_words = hs.total_words;

  /* length buckets: the fixed classes, small ones folded into the next larger one that has pairs
   * (a handful of pairs is not worth a launch ladder of its own; a bigger plan is always correct) */
  b->buckets.clear();
  {
    const long long min_pairs = ctx->knobs.no_buckets ? LLONG_MAX : std::max<long long>(2048, n / 64);
    Bucket cur;
    for (int c = 0; c < MAX_LEN_CLASSES; ++c) {
      if (!hs.cls_n[c]) continue;
      cur.n += hs.cls_n[c];
      cur.maxp = std::max(cur.maxp, hs.cls_maxp[c]); cur.maxt = std::max(cur.maxt, hs.cls_maxt[c]);
      cur.max_len = length_class_limit(c);
      bool later = false;
      for (int d = c + 1; d < MAX_LEN_CLASSES; ++d) later |= hs.cls_n[d] != 0;
      if (cur.n >= min_pairs || !later) { b->buckets.push_back(cur); cur = Bucket(); }
    }
    if (b->buckets.empty()) b->buckets.push_back(Bucket());
    int base = 0;
    for (Bucket& q : b->buckets) { q.list_base = base; base += (int)q.n; }
  }
  const int nb = (int)b->buckets.size();

  const size_t n1 = (size_t)std::max<int64_t>(n, 1);
  CK(b->pairs.ensure(sizeof(PairMeta) * n1));
  CK(b->words.ensure(4 * (size_t)(b->total_words + 1)));
  CK(b->offs.ensure(16 * n1));
  CK(b->lens.ensure(8 * n1));
  CK(b->lay.ensure(16 * (size_t)(layout_tiles(n) + 2)));
  CK(b->pcount.ensure(sizeof(PackCounters) + 4 * MAX_BUCKETS));
  if (nb > 1) CK(b->blist.ensure(4 * n1));
  int64_t* d_poff = b->offs.as<int64_t>(); int64_t* d_toff = d_poff + n1;
  int32_t* d_plen = b->lens.as<int32_t>(); int32_t* d_tlen = d_plen + n1;
  int64_t h2d = 0;
  if (n) {
    int rc = upload(ctx, d_plen, p_len, 4 * (size_t)n, in.k_plen, st);
    if (rc == WFAGPU_OK) rc = upload(ctx, d_tlen, t_len, 4 * (size_t)n, in.k_tlen, st);
    if (rc != WFAGPU_OK) return rc;
    h2d += 8 * n;
  }
  PackArgs& A = b->pack;
  A.n = n; A.pairs = b->pairs.as<PairMeta>(); A.words = b->words.as<uint32_t>(); A.words2 = nullptr;
  A.counters = b->pcount.as<PackCounters>();
  A.p_off = d_poff; A.t_off = d_toff;
  const int64_t span = hs.hi - hs.lo;
  const bool dense = span <= 2 * hs.seq_bytes + 65536;
  bool gen_off = false;
  if (n && in.k_seq != MEM_LOCAL && !dense) {
    /* scattered pairs (e.g. one pattern against texts spread over a genome): gather them back to back
     * on the host; the offsets they have there go up instead of the caller's */
    if (ctx->gather_busy) { CK(cudaEventSynchronize(ctx->gather_done)); ctx->gather_busy = false; }
    CK(ctx->gather_seq.ensure((size_t)hs.seq_bytes + 16));
    CK(ctx->gather_off.ensure(16 * n1));
    int64_t* g_poff = ctx->gather_off.as<int64_t>(); int64_t* g_toff = g_poff + n1;
    gather_pairs(in.seq, p_off, p_len, t_off, t_len, n, ctx->gather_seq.as<uint8_t>(), g_poff, g_toff);
    CK(b->ascii.ensure((size_t)hs.seq_bytes + 16));
    CK(cudaMemcpyAsync(b->ascii.p, ctx->gather_seq.p, (size_t)hs.seq_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_poff, g_poff, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_toff, g_toff, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->gather_done, st));
    ctx->gather_busy = true;
    A.ascii = b->ascii.as<uint8_t>(); A.base = 0;
    h2d += hs.seq_bytes + 16 * n;
  } else if (n) {
    int rc = WFAGPU_OK;
    gen_off = hs.back_to_back && hs.first_negative < 0;     /* pairs back to back: the layout scan writes their offsets */
    if (!gen_off) {
      rc = upload(ctx, d_poff, p_off, 8 * (size_t)n, in.k_poff, st);
      if (rc == WFAGPU_OK) rc = upload(ctx, d_toff, t_off, 8 * (size_t)n, in.k_toff, st);
      if (rc != WFAGPU_OK) return rc;
      h2d += 16 * n;
    }
    if (in.k_seq == MEM_LOCAL) {
      A.ascii = in.seq; A.base = 0;          /* the bases already live on this device: packed in place */
    } else {
      CK(b->ascii.ensure((size_t)span + 16));
      rc = upload(ctx, b->ascii.p, in.seq + hs.lo, (size_t)span, in.k_seq, st);
      if (rc != WFAGPU_OK) return rc;
      A.ascii = b->ascii.as<uint8_t>(); A.base = hs.lo;
      h2d += span;
    }
  }
  BucketArgs B;
  memset(&B, 0, sizeof B);
  B.nbuckets = nb;
  for (int q = 0; q < nb; ++q) { B.max_len[q] = b->buckets[q].max_len; B.list_base[q] = b->buckets[q].list_base; }
  B.cursor = reinterpret_cast<int*>(b->pcount.as<unsigned char>() + sizeof(PackCounters));
  B.list = b->blist.as<int>();
  if (pk != st) {                      /* the kernels run behind the copies, in a stream of their own */
    CK(cudaEventRecord(copied, st));
    CK(cudaStreamWaitEvent(pk, copied, 0));
  }
  CK(launch_layout(d_plen, d_tlen, n, bpw, b->lay.as<long long>(), A.pairs, B, b->pcount.as<uint32_t>(),
                   (int)((sizeof(PackCounters) + 4 * MAX_BUCKETS) / 4), b->words.as<uint32_t>() + b->total_words,
                   gen_off ? (long long*)d_poff : nullptr, gen_off ? (long long*)d_toff : nullptr, n ? (long long)p_off[0] : 0, pk));
  CK(launch_pack(A, b->byte_mode, std::max(b->maxp, b->maxt), ctx->sms, pk));
  CK(cudaMemcpyAsync(counts, b->pcount.p, sizeof(PackCounters), cudaMemcpyDeviceToHost, pk));
  b->stats.kernel_launches = n ? 4 : 1;
  b->stats.packed_bytes = 4 * b->total_words;
  b->stats.h2d_bytes = h2d;
  return WFAGPU_OK;
}

/* Second half of the staging, after stream `staged_on` reached the end of batch_stage (the caller
 * waited for it): the byte side buffer when the packer met other bytes than ACGT, result buffers,
 * kernel parameters, one tier plan per length bucket.  Kernels go to stream st. */
int batch_finish_stage(wfagpu_ctx* ctx, wfagpu_batch* b, const PackCounters* counts, cudaStream_t st) {
  const int64_t n = b->n;
  const size_t n1 = (size_t)std::max<int64_t>(n, 1);
  b->n_side = b->byte_mode ? 0 : (int64_t)counts->n_side;
  b->total_words2 = b->byte_mode ? 0 : (int64_t)counts->side_words;
  if (b->n_side) {
    CK(b->words2.ensure(4 * (size_t)(b->total_words2 + 1)));
    b->pack.words2 = b->words2.as<uint32_t>();
    CK(cudaMemsetAsync(b->words2.as<uint32_t>() + b->total_words2, 0, 4, st));
    CK(launch_pack_side(b->pack, ctx->sms, st));
    b->stats.kernel_launches++;
  }
  CK(b->score.ensure(4 * n1));
  CK(b->status.ensure(4 * n1));
  CK(b->retry_a.ensure(4 * n1));
  CK(b->retry_b.ensure(4 * n1));
  CK(b->counters.ensure(sizeof(DevCounters)));
  if (b->full) {
    CK(b->locs.ensure(16 * n1));
    CK(b->nruns.ensure(4 * n1));
    CK(b->runs_base.ensure(8 * n1));
    CK(b->cig_off.ensure(8 * (n1 + 1)));
    CK(b->tile_sums.ensure(8 * (size_t)(cigar_order_tiles(n) + 2)));
    /* staging for the un-ordered runs: sized for typical CIGARs, regrown (and the batch re-run)
     * in batch_run if a batch needs more; the hard bound is one run per base */
    const unsigned long long bound = (unsigned long long)b->seq_bytes + 2ull * (unsigned long long)n;
    unsigned long long cap = std::max<unsigned long long>(b->runs_tmp.cap / 4, std::min<unsigned long long>(bound, 40ull * n + 4096));
    cap = std::min(cap, bound);
    CK(b->runs_tmp.ensure(4 * (size_t)std::max<unsigned long long>(cap, 1)));
    b->runs_tmp_cap = cap;
    b->runs_bound = bound;
  }
  KParams& k = b->kp;
  fill_kparams(b->cfg, k);
  if (!ctx->knobs.no_metric_map) metric_as_affine(b->cfg, k);      /* score-only linear / edit / indel: the gap-affine tiers */
  k.byte_mode = b->byte_mode ? 1 : 0;
  k.wildcard = b->cfg.wildcard & 0xff;
  k.pairs = b->pairs.as<PairMeta>();
  k.words = b->words.as<uint32_t>();
  k.words2 = b->n_side ? b->words2.as<uint32_t>() : nullptr;
  k.score = b->score.as<int>(); k.status = b->status.as<int>();
  k.locs = b->locs.as<int>(); k.nruns = b->nruns.as<int>(); k.runs_base = b->runs_base.as<long long>();
  k.runs_tmp = b->runs_tmp.as<uint32_t>(); k.runs_tmp_cap = b->runs_tmp_cap;
  DevCounters* dc = b->counters.as<DevCounters>();
  k.runs_cursor = &dc->runs_cursor;
  k.cells_total = &dc->cells_total;
  k.dbg = dc->dbg;
  for (Bucket& q : b->buckets) plan_tiers(ctx, b, q);
  return WFAGPU_OK;
}

/* kernels of one batch on stream st; blocks the calling host thread between tiers */
int batch_run(wfagpu_ctx* ctx, wfagpu_batch* b, cudaStream_t st, DevCounters* hc) {
  DevCounters* dc = b->counters.as<DevCounters>();
  const bool trace = ctx->knobs.trace;
  const int64_t staging_launches = b->ran ? 0 : b->stats.kernel_launches;   /* the packing kernels belong to the first run */
  b->stats.kernel_launches = staging_launches;
  b->stats.retried_pairs = 0;
  b->stats.history_bytes = 0;
  b->total_runs = 0;
  if (b->n == 0) { b->ran = true; return WFAGPU_OK; }
  const int nb = (int)b->buckets.size();
  for (int attempt = 0;; ++attempt) {
    /* counters: zero, and the pairs per bucket.  Written by a kernel: a host-to-device copy on this stream would
     * queue behind the uploads of the chunks ahead (measured: 2.3 ms of an 12 ms call, once per call) */
    memset(hc, 0, sizeof *hc);
    SmallInts bn;
    for (int q = 0; q < MAX_BUCKETS; ++q) bn.v[q] = q < nb ? (int)b->buckets[q].n : 0;
    static_assert(sizeof(DevCounters) % 4 == 0 && offsetof(DevCounters, bucket_n) % 4 == 0, "DevCounters is an array of words");
    CK(launch_init_words(reinterpret_cast<uint32_t*>(dc), (int)(sizeof(DevCounters) / 4), (int)(offsetof(DevCounters, bucket_n) / 4), bn, st));
    int li = 0;                      /* launch index: every tier launch of every bucket has its own counters */
    for (int q = 0; q < nb; ++q) {
      Bucket& bk = b->buckets[q];
      long long nwork = bk.n;
      int* lists[2] = {b->retry_a.as<int>(), b->retry_b.as<int>()};
      const int* cur_list = nb > 1 ? b->blist.as<int>() + bk.list_base : nullptr;
      int last_li = -1;
      /* Warp-per-pair tiers that need no arena sized from the work count are launched back to back: each
       * reads its work count (the retries of the tier before it) from device memory, so the host only looks
       * at the counters at the end of such a chain -- one round trip instead of one per tier.  Everything
       * else (one CTA or several CTAs per pair, history arenas) is launched tier by tier. */
      auto chainable = [&](const Tier& t) { return t.mode == 0 && (t.regs > 0 || !b->full); };
      int first_li = -1;
      double chain_t0 = 0;
      for (size_t ti = 0; ti < bk.tiers.size() && nwork > 0; ++ti, ++li) {
        if (li >= MAX_LAUNCH) return fail(ctx, WFAGPU_EINVAL, "tier schedule longer than %d launches", MAX_LAUNCH);
        Tier t = bk.tiers[ti];
        const bool chain_on = chainable(t) && ti + 1 < bk.tiers.size() && chainable(bk.tiers[ti + 1]);   /* the next launch follows without a look */
        KParams k = b->kp;
        k.runcap = (int)std::min<long long>((long long)bk.maxp + bk.maxt + 2, INT_MAX / 2);
        long long groups;
        int blocks;
        int grid_ctas = 1;
        if (t.mode == 0) {
          blocks = (int)std::min<long long>((nwork + t.groups_per_block - 1) / t.groups_per_block, (long long)ctx->sms * t.blocks_per_sm);
          groups = (long long)blocks * t.groups_per_block;
        } else {
          blocks = (int)std::min<long long>(nwork, (long long)ctx->sms * t.blocks_per_sm);
          if (t.max_groups) blocks = std::min(blocks, t.max_groups);
          groups = blocks;
          if (t.mode == 2) {
            /* several CTAs per pair: all SMs work even when only a few pairs fit in HBM */
            /* CTAs per pair: between ~16 and ~2 diagonals per thread and score (a wavefront is about half
             * as wide as the sequences are long).  Few pairs: one CTA per SM and as many CTAs per pair as
             * that allows (two CTAs sharing an SM lengthen every score's critical path); many pairs: two
             * CTAs per SM and the fewest CTAs per pair, which keeps the per-score barrier short.
             * Measured r01: 8 x 10 kbp 68 ms (18 CTAs/pair) vs 83 (8) vs 116 (37, 2/SM); 16 x 100 kbp
             * 3.9 s (24 CTAs/pair, 2/SM) vs 5.4 s (49) vs 9.1 s (9, 1/SM). */
            const long long L = (long long)bk.maxp + bk.maxt;
            const int resident2 = ctx->sms * std::max(1, t.blocks_per_sm);     /* co-resident CTAs (cooperative launch) */
            const int ncta_min = (int)std::min<long long>(resident2, std::max<long long>(1, L / 8192));
            const int ncta_max = (int)std::min<long long>(resident2, std::max<long long>(1, L / 1024));
            groups = std::min<long long>(groups, std::max(1, resident2 / ncta_min));
            const int resident = (groups * ncta_min <= ctx->sms) ? ctx->sms : resident2;
            grid_ctas = std::max(1, std::min(std::max(resident / (int)groups, ncta_min), ncta_max));
            if ((long long)grid_ctas * groups > resident2) grid_ctas = std::max(1, resident2 / (int)groups);
            blocks = (int)groups;
          }
        }
        k.wcap = t.wcap; k.seq_words_cap = t.seq_words_cap; k.group_bytes = t.group_bytes; k.vec_seqw = t.vec_seqw ? 1 : 0;
        k.hcap = t.hcap; k.scap = t.scap;
        if (t.regs) {
          const RegWindow rw = reg_window(t.regs, k.endsfree, k.match, k.pbf, k.tbf);
          k.reg_kbase = rw.kbase; k.reg_clo = rw.c_lo; k.reg_chi = rw.c_hi;
          if (b->full) {
            k.rhrows = t.scap; k.rhist_bytes = (long long)t.scap * 32 * t.regs;      /* one byte per lane and packed register */
            k.ropcap = (int)(((long long)bk.maxp + bk.maxt + 8 + 15) & ~15ll);
            if (!reg_hist_in_smem(t.regs, true)) {
              CK(ctx->rhist.ensure((size_t)k.rhist_bytes * (size_t)groups));
              CK(ctx->rops.ensure((size_t)k.ropcap * (size_t)groups));
              CK(ctx->runs_stage.ensure(4ull * (size_t)k.runcap * (size_t)groups));
              k.rhist = ctx->rhist.as<uint8_t>(); k.rops = ctx->rops.as<uint8_t>();
              k.runs_stage = ctx->runs_stage.as<uint32_t>();
              b->stats.history_bytes = std::max<int64_t>(b->stats.history_bytes, (int64_t)k.rhist_bytes * groups);
            }
          }
        } else if (t.mode == 2) {
          const int ns = k.rm + 2 * k.r1 + (b->two_p ? 2 * k.r2 : 0);
          k.gring_elems = (long long)ns * t.wcap;
          CK(ctx->gring.ensure(4ull * (size_t)k.gring_elems * (size_t)groups));
          k.gring = ctx->gring.as<int>();
        }
        if (b->full && !t.regs) {
          if (t.mode == 2) {
            /* history arena of the widest tier: what is free now, split over the groups */
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            free_b += ctx->hist_code.cap;
            const long long per_group = (long long)((double)free_b * 0.85 / (double)groups);     /* 1 byte per cell */
            k.hcap = std::max<long long>(1024, std::min<long long>(t.hcap, per_group));
          }
          k.ropcap = (int)(((long long)bk.maxp + bk.maxt + 8 + 15) & ~15ll);
          CK(ctx->hist_code.ensure((size_t)k.hcap * (size_t)groups));
          CK(ctx->hmeta.ensure(sizeof(HistRow) * (size_t)k.scap * (size_t)groups));
          CK(ctx->rops.ensure((size_t)k.ropcap * (size_t)groups));
          CK(ctx->runs_stage.ensure(4ull * (size_t)k.runcap * (size_t)groups));
          k.hist_code = ctx->hist_code.as<uint8_t>(); k.rops = ctx->rops.as<uint8_t>();
          k.hmeta = ctx->hmeta.as<HistRow>(); k.runs_stage = ctx->runs_stage.as<uint32_t>();
          b->stats.history_bytes = std::max<int64_t>(b->stats.history_bytes, (int64_t)((size_t)k.hcap * (size_t)groups));
        }
        k.worklist = cur_list;
        k.n_work = (last_li < 0) ? &dc->bucket_n[q] : &dc->retry[last_li];
        k.work_counter = &dc->work[li];
        k.retry_list = lists[li & 1];
        k.skip_groups = (ti + 1 < bk.tiers.size() && !ctx->knobs.no_tier_skip) ? (int)std::min<long long>(groups, INT_MAX / 2) : 0;   /* never on the last tier */
        k.retry_count = &dc->retry[li];
        k.done_count = &dc->done[li]; k.ovf_count = &dc->ovf[li];
        if (first_li < 0) { first_li = li; chain_t0 = trace ? now_ms() : 0; }
        /* One CTA (or several) per pair and many more pairs than CTAs in flight: probe the tier with one
         * wave of pairs first.  If three quarters of them exceed its capacity the rest of the queue goes to
         * the next tier untried (a 10 kbp batch without cut-off otherwise spends a quarter of its time in a
         * tier every pair overflows); else the same launch index continues with the whole queue. */
        const bool probing = t.mode != 0 && ti + 1 < bk.tiers.size() && !ctx->knobs.no_tier_skip && nwork >= 3 * groups;
        k.work_limit = probing ? (int)groups : INT_MAX;
        auto launch_tier = [&]() -> int {
          if (t.regs && t.bytes) CK(launch_regb(k, b->full, blocks, t.threads, t.smem, st));
          else if (t.regs) CK(launch_reg(k, t.regs, b->full, blocks, t.threads, t.smem, st));
          else if (t.vec_nw && t.bytes) CK(launch_vecb(k, b->two_p, b->full, t.vec_nw, k.heuristic, blocks, t.threads, t.smem, st));
          else if (t.vec_nw) CK(launch_vec(k, b->two_p, b->full, t.vec_nw, k.heuristic, blocks, t.threads, t.smem, st));
          else if (t.mode == 2) {
            CK(ctx->gscratch.ensure(grid_scratch_bytes((int)groups)));
            CK(launch_grid(k, b->two_p, b->full, (int)groups, grid_ctas, t.smem, ctx->gscratch.p, st));
          }
          else CK(launch_align(k, b->two_p, b->full, t.mode, t.off16, blocks, t.threads, t.smem, st));
          b->stats.kernel_launches++;
          return WFAGPU_OK;
        };
        { const int r = launch_tier(); if (r != WFAGPU_OK) return r; }
        if (probing) {
          CK(cudaMemcpyAsync(hc, dc, sizeof *hc, cudaMemcpyDeviceToHost, st));
          CK(cudaStreamSynchronize(st));
          if (4ll * hc->retry[li] >= 3ll * groups) {
            CK(launch_forward_rest(cur_list, k.n_work, (int)groups, nwork, k.retry_list, k.retry_count, st));
            b->stats.kernel_launches++;
            if (trace) fprintf(stderr, "[wfagpu]   bucket %d tier %zu probed with %lld pairs: %d overflowed, the rest is forwarded untried\n",
                               q, ti, groups, hc->retry[li]);
          } else {
            /* go on with the whole queue from where the probe stopped (every group's last, failed fetch
             * moved the work counter past the probe) */
            SmallInts w;
            memset(&w, 0, sizeof w);
            w.v[0] = (int)groups;
            CK(launch_init_words(reinterpret_cast<uint32_t*>(&dc->work[li]), 1, 0, w, st));       /* (no copy on this stream: see above) */
            k.work_limit = INT_MAX;
            const int r = launch_tier(); if (r != WFAGPU_OK) return r;
          }
        }
        if (trace)
          fprintf(stderr, "[wfagpu]   bucket %d (<= %d bp) tier %zu (%s nw=%d regs=%d mode=%d wcap=%d smem=%zu B x %d CTA/SM, grid %d x %d)%s\n",
                  q, bk.max_len, ti, t.vec_nw ? (t.bytes ? "vec-bytes" : "vec") : t.bytes ? "reg-bytes" : t.regs ? "reg" : "scalar", t.vec_nw, t.regs, t.mode, t.wcap, t.smem, t.blocks_per_sm, blocks * grid_ctas, t.threads,
                  chain_on ? " +" : "");
        cur_list = lists[li & 1];
        last_li = li;
        if (chain_on) continue;           /* nwork stays the bound of what the next tier can receive */
        CK(cudaMemcpyAsync(hc, dc, sizeof *hc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (trace) {
          fprintf(stderr, "[wfagpu]     %lld pairs in ->", nwork);
          for (int x = first_li; x <= li; ++x) fprintf(stderr, " %d", hc->retry[x]);
          fprintf(stderr, " overflowed, %.2f ms\n", now_ms() - chain_t0);
        }
        if (trace && hc->dbg[0])
          fprintf(stderr, "[wfagpu]     cycles per warp-step: overhead %.0f, blocks %.0f, planner %.0f, barrier wait %.0f, after-barrier %.0f (warp-steps %llu)\n",
                  (double)hc->dbg[1] / hc->dbg[0], (double)hc->dbg[2] / hc->dbg[0], (double)hc->dbg[3] / hc->dbg[0],
                  (double)hc->dbg[4] / hc->dbg[0], (double)hc->dbg[5] / hc->dbg[0], hc->dbg[0]);
        if (trace && hc->dbg[0])
          fprintf(stderr, "[wfagpu]     scanned-range mode entered through: matrix edge %llu, cut-off end cell M %llu I1 %llu D1 %llu I2 %llu D2 %llu\n",
                  hc->dbg[8], hc->dbg[9], hc->dbg[10], hc->dbg[11], hc->dbg[12], hc->dbg[13]);
        nwork = hc->retry[li];
        if (first_li == li - (int)ti) b->stats.retried_pairs += hc->retry[first_li];     /* the chain held the bucket's first tier */
        first_li = -1;
      }
      if (nwork > 0) {
        /* capacity exhausted even on the widest tier: WF_STATUS_OOM (W/wavefront/wfa.h:50) */
        std::vector<int> ids((size_t)nwork);
        CK(cudaMemcpyAsync(ids.data(), cur_list, 4 * (size_t)nwork, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const int oom = WFAGPU_STATUS_OOM, sc = INT_MIN, zero = 0;
        for (int id : ids) {
          CK(cudaMemcpyAsync(b->score.as<int>() + id, &sc, 4, cudaMemcpyHostToDevice, st));
          CK(cudaMemcpyAsync(b->status.as<int>() + id, &oom, 4, cudaMemcpyHostToDevice, st));
          if (b->full) {
            CK(cudaMemcpyAsync(b->nruns.as<int>() + id, &zero, 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(b->locs.as<int>() + 4 * (size_t)id, 0, 16, st));
          }
        }
        CK(cudaStreamSynchronize(st));
      }
    }
    b->stats.cells = (int64_t)hc->cells_total;
    if (b->full && hc->runs_cursor > b->runs_tmp_cap) {
      /* CIGARs longer than the staging estimate: regrow to what this batch asked for and redo it.  The
       * cursor counts every run any pair wanted, and one run per base bounds it, so this ends. */
      if (b->runs_tmp_cap >= b->runs_bound || attempt >= 3)
        return fail(ctx, WFAGPU_ENOMEM, "CIGAR staging overflow: %llu runs wanted, bound %llu", hc->runs_cursor, b->runs_bound);
      const unsigned long long cap = std::min<unsigned long long>(b->runs_bound, hc->runs_cursor + hc->runs_cursor / 8 + 4096);
      CK(b->runs_tmp.ensure(4 * (size_t)cap));
      b->runs_tmp_cap = cap;
      b->kp.runs_tmp = b->runs_tmp.as<uint32_t>(); b->kp.runs_tmp_cap = cap;
      b->stats.kernel_launches = staging_launches;
      b->stats.retried_pairs = 0;
      continue;
    }
    break;
  }
  if (b->full) {
    b->total_runs = (long long)std::min<unsigned long long>(hc->runs_cursor, b->runs_tmp_cap);
    CK(b->runs_out.ensure(4 * (size_t)std::max<long long>(b->total_runs, 1)));
    CK(launch_cigar_order(b->nruns.as<int>(), b->runs_base.as<long long>(), b->n, b->tile_sums.as<long long>(),
                          b->runs_tmp.as<uint32_t>(), b->cig_off.as<long long>(), nullptr, 0, st));
    CK(launch_cigar_order(b->nruns.as<int>(), b->runs_base.as<long long>(), b->n, b->tile_sums.as<long long>(),
                          b->runs_tmp.as<uint32_t>(), b->cig_off.as<long long>(), b->runs_out.as<uint32_t>(), b->cig_base, st));
    b->stats.kernel_launches += 4;
  }
  b->ran = true;
  return WFAGPU_OK;
}

/* download into host arrays; runs go to runs_dst (pinned, library-owned) */
int batch_download(wfagpu_ctx* ctx, wfagpu_batch* b, cudaStream_t st, int32_t* score, int32_t* status, int32_t* locs,
                   int64_t* cig_off, bool last_offset, uint32_t* runs_dst) {
  const size_t n = (size_t)b->n;
  int64_t d2h = 0;
  if (n) {
    if (score) { CK(cudaMemcpyAsync(score, b->score.p, 4 * n, cudaMemcpyDeviceToHost, st)); d2h += 4 * n; }
    if (status) { CK(cudaMemcpyAsync(status, b->status.p, 4 * n, cudaMemcpyDeviceToHost, st)); d2h += 4 * n; }
  }
  if (b->full && n) {
    if (locs) { CK(cudaMemcpyAsync(locs, b->locs.p, 16 * n, cudaMemcpyDeviceToHost, st)); d2h += 16 * n; }
    if (cig_off) {
      const size_t cnt = n + (last_offset ? 1 : 0);
      CK(cudaMemcpyAsync(cig_off, b->cig_off.p, 8 * cnt, cudaMemcpyDeviceToHost, st)); d2h += 8 * cnt;
    }
    if (runs_dst && b->total_runs) {
      CK(cudaMemcpyAsync(runs_dst, b->runs_out.p, 4 * (size_t)b->total_runs, cudaMemcpyDeviceToHost, st));
      d2h += 4 * b->total_runs;
    }
  } else {
    if (locs && n) memset(locs, 0, 16 * n);
    if (cig_off) for (size_t i = 0; i < n + (last_offset ? 1 : 0); ++i) cig_off[i] = b->cig_base;
  }
  CK(cudaStreamSynchronize(st));
  b->stats.d2h_bytes = d2h;
  return WFAGPU_OK;
}

int check_args(wfagpu_ctx* ctx, const wfagpu_config_t* cfg, const uint8_t* seq, const int64_t* p_off,
               const int32_t* p_len, const int64_t* t_off, const int32_t* t_len, int64_t n) {
  if (!ctx || !cfg || n < 0 || (n > 0 && (!seq || !p_off || !p_len || !t_off || !t_len)))
    return fail(ctx, WFAGPU_EINVAL, "bad arguments");
  if (n > INT_MAX / 2) return fail(ctx, WFAGPU_EINVAL, "at most %d pairs per batch", INT_MAX / 2);
  char msg[400];
  const int rc = wfagpu_config_check(cfg, -1, -1, msg, sizeof msg);
  if (rc != WFAGPU_OK) return fail(ctx, rc, "%s", msg);
  return WFAGPU_OK;
}

/* after a failed call: nothing of it may still be in flight when its buffers are reused */
void quiesce(wfagpu_ctx* ctx) {
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->pack_stream);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->d2h_stream);
  cudaGetLastError();
  for (bool& f : ctx->ring.busy) f = false;
  ctx->gather_busy = false;
}

}  // namespace

extern "C" void wfagpu_batch_free(wfagpu_ctx* ctx, wfagpu_batch* b) {
  if (!b) return;
  if (ctx) {
    std::lock_guard<std::mutex> call(ctx->call_mu);
    cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
    batch_recycle(ctx, b);
  } else {
    batch_release(b);
  }
}

extern "C" int wfagpu_batch_prepare(wfagpu_ctx* ctx, const wfagpu_config_t* cfg, const uint8_t* seq,
                                    const int64_t* p_off, const int32_t* p_len, const int64_t* t_off,
                                    const int32_t* t_len, int64_t n, wfagpu_batch** out) {
  if (!out) return fail(ctx, WFAGPU_EINVAL, "bad arguments");
  *out = nullptr;
  int rc = check_args(ctx, cfg, seq, p_off, p_len, t_off, t_len, n);
  if (rc != WFAGPU_OK) return rc;
  std::lock_guard<std::mutex> call(ctx->call_mu);
  CK(cudaSetDevice(ctx->device));
  read_knobs(ctx);
  Inputs in{seq, p_off, p_len, t_off, t_len};
  if ((rc = classify_inputs(ctx, in)) != WFAGPU_OK) return rc;
  CK(ctx->pin_pack[0].ensure(sizeof(PackCounters)));
  wfagpu_batch* b = batch_acquire(ctx);
  const double t0 = now_ms();
  rc = batch_stage(ctx, b, cfg, in, 0, n, ctx->stream, ctx->stream, nullptr, ctx->pin_pack[0].as<PackCounters>());
  if (rc == WFAGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, WFAGPU_ECUDA, "staging failed: %s", cudaGetErrorString(cudaGetLastError()));
  const double t1 = now_ms();
  if (rc == WFAGPU_OK) rc = batch_finish_stage(ctx, b, ctx->pin_pack[0].as<PackCounters>(), ctx->stream);
  if (rc == WFAGPU_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, WFAGPU_ECUDA, "staging failed: %s", cudaGetErrorString(cudaGetLastError()));
  /* a resident batch keeps only what the kernels read: the raw bases and the caller's arrays go */
  b->ascii.release(); b->offs.release(); b->lens.release();
  if (ctx->knobs.trace) fprintf(stderr, "[wfagpu]   prepare n=%lld: upload + pack %.2f ms, plan %.2f ms\n", (long long)n, t1 - t0, now_ms() - t1);
  if (rc != WFAGPU_OK) { quiesce(ctx); batch_recycle(ctx, b); return rc; }
  *out = b;
  return WFAGPU_OK;
}

extern "C" int wfagpu_batch_run(wfagpu_ctx* ctx, wfagpu_batch* b, void* stream) {
  if (!ctx || !b) return fail(ctx, WFAGPU_EINVAL, "bad arguments to wfagpu_batch_run");
  std::lock_guard<std::mutex> call(ctx->call_mu);
  CK(cudaSetDevice(ctx->device));
  read_knobs(ctx);
  CK(ctx->pin_small.ensure(sizeof(DevCounters) + 64));
  b->cig_base = 0;
  return batch_run(ctx, b, stream ? (cudaStream_t)stream : ctx->stream, ctx->pin_small.as<DevCounters>());
}

extern "C" int wfagpu_batch_fetch(wfagpu_ctx* ctx, wfagpu_batch* b, int32_t* score, int32_t* status, int32_t* locs,
                                  int64_t* cig_off, const uint32_t** cig_runs) {
  if (!ctx || !b) return fail(ctx, WFAGPU_EINVAL, "bad arguments to wfagpu_batch_fetch");
  if (!b->ran) return fail(ctx, WFAGPU_EINVAL, "wfagpu_batch_fetch before wfagpu_batch_run");
  std::lock_guard<std::mutex> call(ctx->call_mu);
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  uint32_t* runs_dst = nullptr;
  if (cig_runs) {
    CK(ctx->pin_runs.ensure(4 * (size_t)std::max<long long>(b->total_runs, 1)));
    runs_dst = ctx->pin_runs.as<uint32_t>();
    *cig_runs = runs_dst;
  }
  return batch_download(ctx, b, ctx->stream, score, status, locs, cig_off, true, runs_dst);
}

extern "C" int wfagpu_batch_get_stats(const wfagpu_batch* b, wfagpu_batch_stats_t* out) {
  if (!b || !out) return WFAGPU_EINVAL;
  *out = b->stats;
  return WFAGPU_OK;
}

/*
 * The one-call path.  Large batches are cut into chunks and pipelined over three host threads and
 * three streams: the calling thread stages chunk c+1 (host pass over the arrays, upload of the raw
 * bases, packing kernels) while the GPU thread runs the alignment kernels of chunk c and queues its
 * result download, and the drain thread hands finished downloads to the caller's arrays.  Pinned
 * caller memory is read and written by DMA directly, so for it the host threads move no data.
 */
extern "C" int wfagpu_align_batch(wfagpu_ctx* ctx, const wfagpu_config_t* cfg, const uint8_t* seq,
                                  const int64_t* p_off, const int32_t* p_len, const int64_t* t_off,
                                  const int32_t* t_len, int64_t n, int32_t* score, int32_t* status,
                                  int32_t* locs, int64_t* cig_off, const uint32_t** cig_runs) {
  int rc = check_args(ctx, cfg, seq, p_off, p_len, t_off, t_len, n);
  if (rc != WFAGPU_OK) return rc;
  std::lock_guard<std::mutex> call(ctx->call_mu);
  CK(cudaSetDevice(ctx->device));
  read_knobs(ctx);
  const bool trace = ctx->knobs.trace;
  Inputs in{seq, p_off, p_len, t_off, t_len};
  if ((rc = classify_inputs(ctx, in)) != WFAGPU_OK) return rc;
  /* result arrays in pinned memory are written by the copy engine; others through the landing zone */
  const bool dma_score = mem_kind(ctx, score) == MEM_DMA, dma_status = mem_kind(ctx, status) == MEM_DMA;
  const bool dma_locs = mem_kind(ctx, locs) == MEM_DMA, dma_cig = mem_kind(ctx, cig_off) == MEM_DMA;
  CK(ctx->pin_small.ensure(sizeof(DevCounters) + 64));
  for (auto& pb : ctx->pin_pack) CK(pb.ensure(sizeof(PackCounters)));
  if (cig_runs) { CK(ctx->pin_runs.ensure(4)); *cig_runs = ctx->user_runs ? ctx->user_runs : ctx->pin_runs.as<uint32_t>(); }
  const double t_start = now_ms();
  /* chunking: the first chunks are small (the first upload is the only one nothing overlaps), the others are a
   * twelfth of the batch: when the kernels bind, a chunk costs a fixed ~0.3 ms (launches, one host round
   * trip, the tail of its persistent kernels); when the upload binds (several GPUs sharing the host's
   * memory bandwidth), what is not overlapped is the kernel time of the LAST chunk, so chunks must not be
   * large either.  Chunks are also bounded by bytes (kShells shells of raw bases live in HBM). */
  std::vector<int64_t> starts;           /* chunk c covers pairs [starts[c], starts[c+1]) */
  {
    const int64_t want = ctx->knobs.chunk;
    int64_t chunk = n, first = n;
    if (want > 0) chunk = first = want;
    else if (n >= 262144) {
      double mean = 0;
      const int64_t probe = std::min<int64_t>(n, 1024);
      for (int64_t i = 0; i < probe; ++i) mean += (double)std::max(p_len[i], 0) + (double)std::max(t_len[i], 0);
      mean = std::max(mean / (double)probe, 1.0);
      const int64_t by_bytes = std::max<int64_t>(65536, (int64_t)(1.0e9 / mean));
      chunk = std::min(by_bytes, std::max<int64_t>(131072, (n + 11) / 12));
      first = std::min<int64_t>(chunk, 32768);
    }
    /* ramp: 32 k pairs, then doubling up to the chunk size -- a chunk's kernels hide the upload of the next,
     * which the copy engine moves about twice as fast as the short-read kernels consume it */
    starts.push_back(0);
    for (int64_t off = std::min(first, n), sz = first; off < n; off += sz) {
      starts.push_back(off);
      sz = std::min(chunk, 2 * sz);
    }
    starts.push_back(n);
    if (n == 0) starts.assign({0, 0});
  }
  int64_t nchunks = (int64_t)starts.size() - 1;
  if (nchunks > 1 && in.k_seq != MEM_LOCAL) {
    /* Pairs in shuffled order: every chunk touches the caller's whole byte range although the range as
     * a whole is dense.  Upload it once and pack every chunk in place from that copy, instead of
     * gathering each chunk's pairs on the host. */
    PairScan c0, all;
    scan_pairs(p_off, p_len, t_off, t_len, starts[1], 16, &c0);
    if (c0.first_negative < 0 && c0.hi - c0.lo > 2 * c0.seq_bytes + 65536) {
      scan_pairs(p_off, p_len, t_off, t_len, n, 16, &all);
      if (all.first_negative < 0 && all.hi - all.lo <= 2 * all.seq_bytes + 65536) {
        const size_t span = (size_t)(all.hi - all.lo);
        CK(ctx->whole_seq.ensure(span + 16));
        rc = upload(ctx, ctx->whole_seq.p, seq + all.lo, span, in.k_seq, ctx->copy_stream);
        if (rc != WFAGPU_OK) return rc;
        CK(cudaEventRecord(ctx->whole_done, ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->pack_stream, ctx->whole_done, 0));
        in.seq = ctx->whole_seq.as<uint8_t>() - all.lo;      /* device address of the caller's byte 0 */
        in.k_seq = MEM_LOCAL;
        /* nothing can start before the whole range has arrived, so small first chunks buy nothing, and a
         * mixed-length chunk pays its fixed costs once per length bucket: three chunks (pack / align / download
         * still overlap).  Measured on the 920 k-pair mixed workload: 9 chunks 29.7 ms, 3 chunks see profiles/. */
        if (ctx->knobs.chunk <= 0) {
          const int64_t third = std::max<int64_t>(131072, (n + 2) / 3);
          starts.clear();
          for (int64_t off = 0; off < n; off += third) starts.push_back(off);
          starts.push_back(n);
        }
      }
    }
  }
  nchunks = (int64_t)starts.size() - 1;
  wfagpu_batch* shells[kShells];
  for (int i = 0; i < kShells; ++i) shells[i] = i < nchunks ? batch_acquire(ctx) : nullptr;
  std::mutex mu;
  std::condition_variable cv;
  int64_t staged = 0, consumed = 0;     /* chunks staged by the caller thread / finished by the GPU thread */
  int err = WFAGPU_OK;
  long long run_base = 0;
  int64_t launches = 0;
  double gpu_busy = 0, t_run = 0, t_down = 0;

  auto stage_side = [&](int64_t c) -> int {     /* runs on the calling thread */
    const int r = batch_stage(ctx, shells[c % kShells], cfg, in, starts[c], starts[c + 1] - starts[c], ctx->copy_stream,
                              ctx->pack_stream, ctx->copied[c % kShells], ctx->pin_pack[c % kShells].as<PackCounters>());
    if (r != WFAGPU_OK) return r;
    CK(cudaEventRecord(ctx->uploaded[c % kShells], ctx->pack_stream));
    return WFAGPU_OK;
  };
  auto finish_stage = [&](int64_t c) -> int {   /* GPU thread: the packer's verdict is in, plan the chunk */
    CK(cudaEventSynchronize(ctx->uploaded[c % kShells]));
    return batch_finish_stage(ctx, shells[c % kShells], ctx->pin_pack[c % kShells].as<PackCounters>(), ctx->stream);
  };

  if (nchunks == 1) {
    wfagpu_batch* b = shells[0];
    rc = stage_side(0);
    const double t1 = now_ms();
    if (rc == WFAGPU_OK) rc = finish_stage(0);
    b->cig_base = 0;
    if (rc == WFAGPU_OK) rc = batch_run(ctx, b, ctx->stream, ctx->pin_small.as<DevCounters>());
    uint32_t* runs_dst = nullptr;
    if (rc == WFAGPU_OK && cig_runs && b->full) {
      rc = [&]() -> int {
        if (ctx->user_runs) {
          if ((size_t)b->total_runs > ctx->user_runs_cap)
            return fail(ctx, WFAGPU_ENOMEM, "run buffer too small: %lld runs, capacity %zu", b->total_runs, ctx->user_runs_cap);
          return WFAGPU_OK;
        }
        CK(ctx->pin_runs.ensure(4 * (size_t)std::max<long long>(b->total_runs, 1)));
        return WFAGPU_OK;
      }();
      runs_dst = ctx->user_runs ? ctx->user_runs : ctx->pin_runs.as<uint32_t>();
    }
    if (rc == WFAGPU_OK) rc = batch_download(ctx, b, ctx->stream, score, status, locs, cig_off, true, runs_dst);
    launches = b->stats.kernel_launches;
    if (trace) fprintf(stderr, "[wfagpu] n=%lld single chunk: stage %.2f ms, gpu side %.2f ms\n", (long long)n, t1 - t_start, now_ms() - t1);
  } else {
    struct Drain { int64_t off, m; bool last; long long run_base; bool full; };
    Drain drains[kShells];
    int64_t queued = 0, drained = 0;     /* chunks whose download was queued / handed to the caller */
    auto out_bytes = [&](int64_t m, bool full) { return (size_t)(full ? 8 * (m + 1) + 24 * m : 8 * m) + 64; };
    auto gpu_side_async = [&](int64_t c) -> int {
      wfagpu_batch* b = shells[c % kShells];
      const int64_t off = starts[c];
      const double g0 = now_ms();
      CK(cudaEventSynchronize(ctx->uploaded[c % kShells]));
      const double g0a = now_ms();
      int r = batch_finish_stage(ctx, shells[c % kShells], ctx->pin_pack[c % kShells].as<PackCounters>(), ctx->stream);
      if (r != WFAGPU_OK) return r;
      CK(cudaStreamWaitEvent(ctx->stream, ctx->d2h_done[c % kShells], 0));   /* chunk c - kShells left this shell's result buffers */
      b->cig_base = run_base;
      const double g0b = now_ms();
      r = batch_run(ctx, b, ctx->stream, ctx->pin_small.as<DevCounters>());
      if (r != WFAGPU_OK) return r;
      const double g1 = now_ms();
      if (trace) fprintf(stderr, "[wfagpu]   chunk %lld: waited %.2f ms for its staging, planned in %.2f ms, kernels %.2f ms\n",
                         (long long)c, g0a - g0, g0b - g0a, g1 - g0b);
      {
        std::unique_lock<std::mutex> lk(mu);                          /* landing zone c&1 free again? */
        cv.wait(lk, [&] { return drained + kShells > c || err != WFAGPU_OK; });
        if (err != WFAGPU_OK) return err;
      }
      const size_t m = (size_t)b->n;
      const bool last = c == nchunks - 1;
      PinBuf& stg = ctx->out_stage[c % kShells];
      CK(stg.ensure(out_bytes((int64_t)m, b->full)));
      unsigned char* base = stg.as<unsigned char>();
      cudaStream_t ds = ctx->d2h_stream;
      CK(cudaEventRecord(ctx->run_done, ctx->stream));       /* the CIGAR ordering kernels are still in flight */
      CK(cudaStreamWaitEvent(ds, ctx->run_done, 0));
      /* landing zone layout: cig_off[m+1] | locs[4m] | score[m] | status[m]; pinned caller arrays are written directly */
      if (b->full) {
        if (cig_off) CK(cudaMemcpyAsync(dma_cig ? (void*)(cig_off + off) : (void*)base, b->cig_off.p, 8 * (m + (last || !dma_cig ? 1 : 0)), cudaMemcpyDeviceToHost, ds));
        if (locs) CK(cudaMemcpyAsync(dma_locs ? (void*)(locs + 4 * off) : (void*)(base + 8 * (m + 1)), b->locs.p, 16 * m, cudaMemcpyDeviceToHost, ds));
        if (cig_runs && b->total_runs && ctx->user_runs) {
          if ((size_t)(run_base + b->total_runs) > ctx->user_runs_cap)
            return fail(ctx, WFAGPU_ENOMEM, "run buffer too small: %lld runs so far, capacity %zu", run_base + b->total_runs, ctx->user_runs_cap);
          CK(cudaMemcpyAsync(ctx->user_runs + run_base, b->runs_out.p, 4 * (size_t)b->total_runs, cudaMemcpyDeviceToHost, ds));
        } else if (cig_runs && b->total_runs) {
          const size_t need = 4 * (size_t)(run_base + b->total_runs);
          if (need > ctx->pin_runs.cap) {
            /* grow the library-owned run buffer (sized from this chunk for all remaining ones),
             * keeping what earlier chunks wrote; their downloads must have landed first */
            CK(cudaStreamSynchronize(ds));
            PinBuf bigger;
            CK(bigger.ensure(std::max(need, 4 * (size_t)run_base + 4 * (size_t)(b->total_runs + b->total_runs / 8) * (size_t)(nchunks - c))));
            if (run_base) memcpy(bigger.p, ctx->pin_runs.p, 4 * (size_t)run_base);
            ctx->pin_runs.release();
            ctx->pin_runs = bigger;
          }
          CK(cudaMemcpyAsync(ctx->pin_runs.as<uint32_t>() + run_base, b->runs_out.p, 4 * (size_t)b->total_runs,
                             cudaMemcpyDeviceToHost, ds));
        }
      }
      const size_t so = b->full ? 8 * (m + 1) + 16 * m : 0;
      if (score) CK(cudaMemcpyAsync(dma_score ? (void*)(score + off) : (void*)(base + so), b->score.p, 4 * m, cudaMemcpyDeviceToHost, ds));
      if (status) CK(cudaMemcpyAsync(dma_status ? (void*)(status + off) : (void*)(base + so + 4 * m), b->status.p, 4 * m, cudaMemcpyDeviceToHost, ds));
      CK(cudaEventRecord(ctx->d2h_done[c % kShells], ds));
      {
        std::lock_guard<std::mutex> lk(mu);
        drains[c % kShells] = Drain{off, (int64_t)m, last, run_base, b->full};
        queued = c + 1;
        cv.notify_all();
      }
      if (b->full) run_base += b->total_runs;
      launches += b->stats.kernel_launches;
      t_run += g1 - g0; t_down += now_ms() - g1;
      return WFAGPU_OK;
    };
    std::thread drain_thread([&] {
      cudaSetDevice(ctx->device);
      for (int64_t c = 0; c < nchunks; ++c) {
        Drain d;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return queued > c || err != WFAGPU_OK; });
          if (queued <= c) return;
          d = drains[c % kShells];
        }
        if (cudaEventSynchronize(ctx->d2h_done[c % kShells]) != cudaSuccess) {
          std::lock_guard<std::mutex> lk(mu);
          if (err == WFAGPU_OK) err = fail(ctx, WFAGPU_ECUDA, "result download failed");
          cv.notify_all();
          return;
        }
        const unsigned char* base = ctx->out_stage[c % kShells].as<unsigned char>();
        const size_t m = (size_t)d.m;
        const size_t so = d.full ? 8 * (m + 1) + 16 * m : 0;
        if (d.full) {
          if (cig_off && !dma_cig) par_memcpy(cig_off + d.off, base, 8 * (m + (d.last ? 1 : 0)));
          if (locs && !dma_locs) par_memcpy(locs + 4 * d.off, base + 8 * (m + 1), 16 * m);
        } else {
          if (locs && m) memset(locs + 4 * d.off, 0, 16 * m);
          if (cig_off) for (size_t i = 0; i < m + (d.last ? 1 : 0); ++i) cig_off[d.off + i] = d.run_base;
        }
        if (score && !dma_score) par_memcpy(score + d.off, base + so, 4 * m);
        if (status && !dma_status) par_memcpy(status + d.off, base + so + 4 * m, 4 * m);
        std::lock_guard<std::mutex> lk(mu);
        drained = c + 1;
        cv.notify_all();
      }
    });
    std::thread gpu_thread([&] {
      cudaSetDevice(ctx->device);
      for (int64_t c = 0; c < nchunks; ++c) {
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return staged > c || err != WFAGPU_OK; });
          if (err != WFAGPU_OK) return;
        }
        const double t0 = now_ms();
        const int r = gpu_side_async(c);
        gpu_busy += now_ms() - t0;
        std::lock_guard<std::mutex> lk(mu);
        if (r != WFAGPU_OK && err == WFAGPU_OK) err = r;
        consumed = c + 1;
        cv.notify_all();
        if (r != WFAGPU_OK) return;
      }
    });
    double stage_ms = 0;
    for (int64_t c = 0; c < nchunks; ++c) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return consumed + kShells > c || err != WFAGPU_OK; });   /* shell c % kShells is free */
        if (err != WFAGPU_OK) break;
      }
      const double t0 = now_ms();
      const int r = stage_side(c);
      stage_ms += now_ms() - t0;
      std::lock_guard<std::mutex> lk(mu);
      if (r != WFAGPU_OK) { if (err == WFAGPU_OK) err = r; cv.notify_all(); break; }
      staged = c + 1;
      cv.notify_all();
    }
    gpu_thread.join();
    { std::lock_guard<std::mutex> lk(mu); cv.notify_all(); }
    drain_thread.join();
    cudaStreamSynchronize(ctx->d2h_stream);
    rc = err;
    if (trace)
      fprintf(stderr, "[wfagpu] n=%lld in %lld chunks: total %.2f ms (staging %.2f ms on the caller; gpu side %.2f ms = plan + kernels %.2f + queueing downloads %.2f; device buffers re-allocated so far: %lld)\n",
              (long long)n, (long long)nchunks, now_ms() - t_start, stage_ms, gpu_busy, t_run, t_down, g_dev_reallocs.load());
  }
  if (cig_runs) *cig_runs = ctx->user_runs ? ctx->user_runs : ctx->pin_runs.as<uint32_t>();
  ctx->last_launches = launches;
  if (rc != WFAGPU_OK) quiesce(ctx);
  for (int i = kShells - 1; i >= 0; --i) if (shells[i]) batch_recycle(ctx, shells[i]);
  return rc;
}

extern "C" int64_t wfagpu_last_launches(const wfagpu_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

/*
 * One pair, low latency.  Short gap-affine pairs without cut-offs whose penalties the register tier is
 * instantiated for go through ONE kernel launch that reads the bases from a mapped mailbox and writes the
 * results back into it (wfa_pair_kernel); everything else is a batch of one.
 */
extern "C" int wfagpu_align_pair(wfagpu_ctx* ctx, const wfagpu_config_t* cfg, const char* pattern, int32_t plen,
                                 const char* text, int32_t tlen, int32_t* score, int32_t* status, int32_t* locs,
                                 const uint32_t** cig_runs, int32_t* n_runs) {
  if (!ctx || !cfg || plen < 0 || tlen < 0 || (plen && !pattern) || (tlen && !text) || !score || !status)
    return fail(ctx, WFAGPU_EINVAL, "bad arguments");
  char msg[400];
  int rc = wfagpu_config_check(cfg, plen, tlen, msg, sizeof msg);
  if (rc != WFAGPU_OK) return fail(ctx, rc, "%s", msg);
  const bool full = cfg->scope == WFAGPU_SCOPE_FULL;
  KParams k;
  memset(&k, 0, sizeof k);
  fill_kparams(*cfg, k);
  /* a wildcard other than A, C, G, T changes nothing for a pair of pure ACGT reads, and the kernel hands every other
   * pair back (rc = 1): loops like pywfa's a(text, pattern) with wildcard="N" keep the one-launch path */
  const int wc = cfg->wildcard & 0xff;
  const bool wc_is_base = wc == 'A' || wc == 'C' || wc == 'G' || wc == 'T';
  const bool fast = cfg->distance == WFAGPU_DISTANCE_AFFINE && cfg->heuristic == WFAGPU_HEURISTIC_NONE && !wc_is_base &&
                    plen <= PAIR_MAX_LEN && tlen <= PAIR_MAX_LEN && k.dx == 2 && k.doe1 == 4 && k.de1 == 1;   /* the shape wfa_pair_kernel is compiled for */
  if (fast) {
    std::lock_guard<std::mutex> call(ctx->call_mu);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->pair_box) CK(cudaHostAlloc((void**)&ctx->pair_box, PAIR_BOX_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
    PairBox* box = reinterpret_cast<PairBox*>(ctx->pair_box);
    box->plen = plen; box->tlen = tlen; box->rc = -1;
    memcpy(ctx->pair_box + PAIR_ASCII_OFF, pattern, (size_t)plen);
    memcpy(ctx->pair_box + PAIR_ASCII_OFF + PAIR_TEXT_OFF, text, (size_t)tlen);
    CK(launch_pair(k, full, box, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->last_launches = 1;
    if (box->rc == 0) {
      *score = box->score; *status = box->status;
      if (locs) memcpy(locs, box->locs, 16);
      if (n_runs) *n_runs = box->nruns;
      if (cig_runs) *cig_runs = reinterpret_cast<const uint32_t*>(ctx->pair_box + PAIR_RUNS_OFF);
      return WFAGPU_OK;
    }
    if (box->rc != 1) return fail(ctx, WFAGPU_ECUDA, "single-pair kernel left no result");
    /* non-ACGT bytes or a wavefront wider than the window: the batch path takes the pair */
  }
  std::string seq;
  seq.reserve((size_t)plen + tlen + 1);
  seq.append(pattern ? pattern : "", (size_t)plen).append(text ? text : "", (size_t)tlen).push_back('\0');
  const int64_t p_off = 0, t_off = plen;
  int64_t cig_off[2] = {0, 0};
  int32_t l4[4] = {0, 0, 0, 0};
  const uint32_t* runs = nullptr;
  rc = wfagpu_align_batch(ctx, cfg, reinterpret_cast<const uint8_t*>(seq.data()), &p_off, &plen, &t_off, &tlen, 1, score, status,
                          l4, cig_off, &runs);
  if (rc != WFAGPU_OK) return rc;
  if (locs) memcpy(locs, l4, 16);
  if (n_runs) *n_runs = (int32_t)(cig_off[1] - cig_off[0]);
  if (cig_runs) *cig_runs = runs + cig_off[0];
  return WFAGPU_OK;
}

extern "C" int wfagpu_set_run_buffer(wfagpu_ctx* ctx, uint32_t* buf, int64_t capacity_words) {
  if (!ctx || capacity_words < 0 || (!buf && capacity_words)) return WFAGPU_EINVAL;
  std::lock_guard<std::mutex> call(ctx->call_mu);
  ctx->user_runs = buf;
  ctx->user_runs_cap = buf ? (size_t)capacity_words : 0;
  return WFAGPU_OK;
}
