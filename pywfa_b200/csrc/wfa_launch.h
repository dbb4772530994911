/* wfa_launch.h -- host-callable launch wrappers of wfa_kernels.cu (internal, C++ linkage). */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wfagpu {

struct KParams;
struct PairMeta;

/* ---- batch staging on the device (wfa_pack.cu) ---- */
constexpr int MAX_BUCKETS = 8;
struct PackCounters {                 /* in HBM, read back by the host after the pack kernels */
  unsigned long long n_side;          /* pairs holding a byte other than ACGT/acgt */
  unsigned long long side_words;      /* words of the byte side buffer they need */
};
struct BucketArgs {                   /* work lists by max(plen, tlen); nbuckets <= 1: none */
  int nbuckets;
  int max_len[MAX_BUCKETS];           /* bucket q takes pairs with max(plen, tlen) <= max_len[q] (the last one: the rest) */
  int list_base[MAX_BUCKETS];         /* first slot of bucket q in `list` (host-side histogram) */
  int* cursor;                        /* [MAX_BUCKETS] zeroed */
  int* list;                          /* n pair ids */
};
struct PackArgs {
  const uint8_t* ascii;               /* device address of byte `base` of the caller's sequence buffer */
  long long base;
  const int64_t* p_off; const int64_t* t_off;     /* offsets into the caller's buffer (device copies) */
  long long n;
  PairMeta* pairs;
  uint32_t* words; uint32_t* words2;
  PackCounters* counters;
};
int layout_tiles(long long n);
/* exclusive scan of the per-pair word counts -> pairs[i] = {woff, plen, tlen}; tile_sums: 2 * (layout_tiles(n) + 1)
 * entries.  Also zeroes zero_a[0 .. n_zero_a) (the packer's counters and bucket cursors, n_zero_a <= 1024) and
 * *zero_b.  gen_poff / gen_toff != nullptr: the pairs lie back to back from byte off_base on; their offsets are
 * generated here instead of uploaded. */
cudaError_t launch_layout(const int32_t* p_len, const int32_t* t_len, long long n, int bases_per_word,
                          long long* tile_sums, PairMeta* pairs, const BucketArgs& B, uint32_t* zero_a, int n_zero_a,
                          uint32_t* zero_b, long long* gen_poff, long long* gen_toff, long long off_base, cudaStream_t st);
/* ASCII -> 2-bit words (or bytes when byte_mode); flags pairs with other bytes (PackCounters) */
cudaError_t launch_pack(const PackArgs& A, bool byte_mode, int max_len, int sms, cudaStream_t st);
/* bytes of the flagged pairs -> words2 */
cudaError_t launch_pack_side(const PackArgs& A, int sms, cudaStream_t st);
/* dst[0 .. nwords) = 0, then dst[at + i] = vals.v[i] (i < MAX_BUCKETS): the per-batch counters, set by a kernel
 * because a host-to-device copy on the compute stream queues behind the uploads of the chunks ahead */
struct SmallInts { int v[MAX_BUCKETS]; };
cudaError_t launch_init_words(uint32_t* dst, int nwords, int at, const SmallInts& vals, cudaStream_t st);

/* mode 0: warp-per-pair (smem ring), 1: block-per-pair (smem ring), 2: block-per-pair (HBM ring) */
cudaError_t launch_align(const KParams& P, bool two_p, bool full, int mode, bool off16, int grid, int block,
                         size_t smem, cudaStream_t st);
cudaError_t init_kernels(int smem_optin);   /* once per device, at context creation */
int align_occupancy(bool two_p, bool full, int mode, bool off16, int block, size_t smem);
size_t block_reduce_smem_bytes();

/* register-resident tier (wfa_reg.cuh): regs = packed registers per wavefront (window 64*regs) */
bool reg_tier_supported(int dx, int doe, int de, int regs, bool full);
cudaError_t launch_reg(const KParams& P, int regs, bool full, int grid, int block, size_t smem, cudaStream_t st);
int reg_occupancy(const KParams& P, int regs, bool full, int block, size_t smem);

/* byte mode on the register tier (wfa_reg_bytes.cu): pairs with non-ACGT bytes / the wildcard, 256-diagonal window */
bool regb_tier_supported(int dx, int doe, int de, bool full);
int regb_regs();
cudaError_t launch_regb(const KParams& P, bool full, int grid, int block, size_t smem, cudaStream_t st);
int regb_occupancy(const KParams& P, bool full, int block, size_t smem);
cudaError_t init_regb(int smem_optin);

/* packed-halfword tier (wfa_vec.cuh): nw = warps per pair (1, 8 or 16) */
cudaError_t launch_vec(const KParams& P, bool two_p, bool full, int nw, int heur, int grid, int block, size_t smem, cudaStream_t st);
int vec_occupancy(bool two_p, bool full, int nw, int heur, int block, size_t smem);

/* ... in byte mode (wfa_vec_bytes.cu): pairs with non-ACGT bytes / the wildcard */
cudaError_t launch_vecb(const KParams& P, bool two_p, bool full, int nw, int heur, int grid, int block, size_t smem, cudaStream_t st);
int vecb_occupancy(bool two_p, bool full, int nw, int heur, int block, size_t smem);
cudaError_t init_vecb(int smem_optin);

/* several CTAs per pair (long reads): groups * ncta co-resident CTAs of 512 threads; `scratch` holds
 * grid_scratch_bytes(groups) bytes of device memory */
size_t grid_scratch_bytes(int groups);
int grid_occupancy(bool two_p, bool full, size_t smem);   /* resident CTAs per SM */
cudaError_t launch_grid(const KParams& P, bool two_p, bool full, int groups, int ncta, size_t smem, void* scratch, cudaStream_t st);

/* append work items [from, *n_work) of a tier's work list to its retry list untried (the tier was probed and given up on) */
cudaError_t launch_forward_rest(const int* worklist, const int* n_work, int from, long long n_bound, int* retry_list,
                                int* retry_count, cudaStream_t st);

/* ---- single pair (wfa_pair_kernel) ---- */
constexpr int PAIR_MAX_LEN = 1000;                       /* longest sequence the one-launch path takes */
constexpr int PAIR_TEXT_OFF = 1024;                      /* text bytes start here in the mailbox's ASCII region */
constexpr int PAIR_ASCII_OFF = 64;                       /* mailbox layout: header | ASCII (2 x 1024) | runs */
constexpr int PAIR_RUNS_OFF = PAIR_ASCII_OFF + 2 * 1024;
constexpr int PAIR_BOX_BYTES = PAIR_RUNS_OFF + 4 * (2 * PAIR_MAX_LEN + 2);
constexpr int PAIR_PK_WORDS = 2 * ((PAIR_MAX_LEN + 15) / 16) + 2;
constexpr int PAIR_OPS_BYTES = 2 * PAIR_MAX_LEN + 16;
constexpr int PAIR_HIST_ROWS = 32 * 4 + 4 + 1;           /* scores the 256-diagonal window can hold */
constexpr int PAIR_HIST_BYTES = PAIR_HIST_ROWS * 32 * 4 > 4 * (2 * PAIR_MAX_LEN + 2) ? PAIR_HIST_ROWS * 32 * 4 : 4 * (2 * PAIR_MAX_LEN + 2);
constexpr int PAIR_SMEM_BYTES = 4 * PAIR_PK_WORDS + 4 * (2 * PAIR_MAX_LEN + 2) + PAIR_OPS_BYTES + PAIR_HIST_BYTES + 32;
struct PairBox {                                          /* header of the mailbox (mapped pinned host memory) */
  int32_t plen, tlen;
  int32_t score, status, locs[4], nruns;
  int32_t rc;                                             /* 0: done; 1: take the batch path */
  long long cells;
};
cudaError_t launch_pair(const KParams& P, bool full, PairBox* box, cudaStream_t st);

/* runs_out == nullptr: count + scan (tile_sums needs cigar_order_tiles(n)+1 entries, total in the
 * last one); otherwise gather into cig_off[n+1] (values offset by cig_base) / runs_out. */
cudaError_t launch_cigar_order(const int* nruns, const long long* runs_base, long long n,
                               long long* tile_sums, const uint32_t* runs_tmp, long long* cig_off,
                               uint32_t* runs_out, long long cig_base, cudaStream_t st);
int cigar_order_tiles(long long n);

}  // namespace wfagpu
