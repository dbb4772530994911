/*
 * wfa_pack.cu -- device side of the batch staging: raw ASCII bases in HBM -> the batch layout the
 * alignment kernels read (PairMeta + 2-bit packed words), plus the per-length-bucket work lists.
 *
 * Replaces wavefront_sequences_init_ascii (W/wavefront/wavefront_sequences.c:141-170: one
 * sentinel-padded byte copy of both sequences per alignment, on the CPU) for the whole batch.
 * The caller's bases are uploaded as they are (straight from pinned memory when the caller
 * provides it) and packed here, so no host core touches a base:
 *
 *   layout_count / layout_scan / layout_apply   exclusive scan of the per-pair word counts
 *        -> PairMeta{woff, plen, tlen}; optional bucket work lists (pair ids by max(plen, tlen))
 *   pack2_kernel        16 bases per 32-bit word, base j in bits 2j..2j+1, code = (c >> 1) & 3
 *        (A=0 C=1 T=2 G=3, either case); pairs holding any other byte get a slot in the byte side
 *        buffer (PairMeta::woff = ~offset) and are counted
 *   pack_bytes_kernel   upper-cased bytes, 4 per word (byte mode / the side buffer)
 *
 * All kernels are HBM streaming kernels (read 1 B, write 0.25 B per base); the bound that matters
 * for them is the PCIe link feeding the bytes, two orders of magnitude below HBM.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "wfa_core.cuh"
#include "wfa_launch.h"
#include "wfa_pack16.cuh"

namespace wfagpu {

namespace {

constexpr int LAY_THREADS = 256;
constexpr int LAY_ITEMS = 16;
constexpr int LAY_TILE = LAY_THREADS * LAY_ITEMS;

__host__ __device__ __forceinline__ long long words_of(int plen, int tlen, int bpw) {
  return (long long)((plen + bpw - 1) / bpw) + (long long)((tlen + bpw - 1) / bpw);
}

/* the two running sums of the layout: packed words and raw bases */
struct Sum2 { long long w, b; };
__device__ __forceinline__ Sum2 operator+(const Sum2& x, const Sum2& y) { return Sum2{x.w + y.w, x.b + y.b}; }
__device__ __forceinline__ Sum2 shfl_up2(const Sum2& x, int o) {
  return Sum2{__shfl_up_sync(0xffffffffu, x.w, o), __shfl_up_sync(0xffffffffu, x.b, o)};
}
__device__ __forceinline__ Sum2 shfl_down2(const Sum2& x, int o) {
  return Sum2{__shfl_down_sync(0xffffffffu, x.w, o), __shfl_down_sync(0xffffffffu, x.b, o)};
}
__device__ __forceinline__ Sum2 item_of(int plen, int tlen, int bpw) { return Sum2{words_of(plen, tlen, bpw), (long long)plen + tlen}; }

__global__ void layout_count_kernel(const int32_t* __restrict__ p_len, const int32_t* __restrict__ t_len, long long n,
                                    int bpw, Sum2* __restrict__ tile_sums) {
  __shared__ Sum2 wsum[LAY_THREADS / 32];
  const long long t0 = (long long)blockIdx.x * LAY_TILE;
  Sum2 acc{0, 0};
  for (int j = 0; j < LAY_ITEMS; ++j) {
    const long long i = t0 + j * LAY_THREADS + threadIdx.x;
    if (i < n) acc = acc + item_of(p_len[i], t_len[i], bpw);
  }
  for (int o = 16; o > 0; o >>= 1) acc = acc + shfl_down2(acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    Sum2 s{0, 0};
    for (int i = 0; i < LAY_THREADS / 32; ++i) s = s + wsum[i];
    tile_sums[blockIdx.x] = s;
  }
}

/* single block: exclusive scan of the tile sums in place; total -> tile_sums[ntiles].  Also clears the
 * packer's counters and the pad word behind the packed sequences (no memset in the copy stream: a
 * memset is a kernel, and kernels queue behind the persistent alignment kernel of the chunk before) */
__global__ void layout_scan_kernel(Sum2* tile_sums, int ntiles, uint32_t* zero_a, int n_zero_a, uint32_t* zero_b) {
  __shared__ Sum2 carry;
  __shared__ Sum2 wtot[32];
  if ((int)threadIdx.x < n_zero_a) zero_a[threadIdx.x] = 0;
  if (threadIdx.x == 0 && zero_b) *zero_b = 0;
  if (threadIdx.x == 0) carry = Sum2{0, 0};
  __syncthreads();
  for (int b = 0; b < ntiles; b += blockDim.x) {
    const int i = b + threadIdx.x;
    const Sum2 v = (i < ntiles) ? tile_sums[i] : Sum2{0, 0};
    Sum2 inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const Sum2 t = shfl_up2(inc, o);
      if ((threadIdx.x & 31) >= o) inc = inc + t;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = inc;
    __syncthreads();
    Sum2 woff{0, 0};
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff = woff + wtot[w];
    const Sum2 excl{carry.w + woff.w + inc.w - v.w, carry.b + woff.b + inc.b - v.b};
    if (i < ntiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}

/* per tile: word offset of every pair -> PairMeta; bucket work lists when nbuckets > 1; when the caller's
 * pairs lie back to back (pattern, text, pattern, ...: gen_off != nullptr) their byte offsets are written
 * here instead of being uploaded: pair i starts off_base + (bases before it) */
__global__ void layout_apply_kernel(const int32_t* __restrict__ p_len, const int32_t* __restrict__ t_len, long long n,
                                    int bpw, const Sum2* __restrict__ tile_sums, PairMeta* __restrict__ pairs,
                                    BucketArgs B, long long* __restrict__ gen_poff, long long* __restrict__ gen_toff,
                                    long long off_base) {
  __shared__ Sum2 wtot[LAY_THREADS / 32];
  __shared__ Sum2 carry_s;
  const long long t0 = (long long)blockIdx.x * LAY_TILE;
  if (threadIdx.x == 0) carry_s = tile_sums[blockIdx.x];
  __syncthreads();
  for (int j = 0; j < LAY_ITEMS; ++j) {
    const long long i = t0 + j * LAY_THREADS + threadIdx.x;
    const int pl = (i < n) ? p_len[i] : 0, tl = (i < n) ? t_len[i] : 0;
    const Sum2 v = item_of(pl, tl, bpw);
    Sum2 inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const Sum2 t = shfl_up2(inc, o);
      if ((threadIdx.x & 31) >= o) inc = inc + t;
    }
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = inc;
    __syncthreads();
    Sum2 woff{0, 0};
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff = woff + wtot[w];
    const Sum2 excl{carry_s.w + woff.w + inc.w - v.w, carry_s.b + woff.b + inc.b - v.b};
    if (i < n) {
      PairMeta m;
      m.woff = excl.w; m.plen = pl; m.tlen = tl;
      pairs[i] = m;
      if (gen_poff) { gen_poff[i] = off_base + excl.b; gen_toff[i] = off_base + excl.b + pl; }
    }
    if (B.nbuckets > 1) {
      /* one atomic per bucket and warp: the lanes of a bucket take consecutive list slots */
      const int L = pl > tl ? pl : tl;
      int bk = B.nbuckets - 1;
      for (int q = B.nbuckets - 2; q >= 0; --q) if (L <= B.max_len[q]) bk = q;
      if (i >= n) bk = -1;
      const unsigned peers = __match_any_sync(0xffffffffu, bk);
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (lane == leader && bk >= 0) base = atomicAdd(&B.cursor[bk], __popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (bk >= 0) B.list[B.list_base[bk] + base + __popc(peers & ((1u << lane) - 1u))] = (int)i;
    }
    __syncthreads();
    if (threadIdx.x == LAY_THREADS - 1) carry_s = excl + v;
    __syncthreads();
  }
}

/* One warp (BLOCK = false) or one CTA (long reads) per pair. */
template <bool BLOCK>
__global__ void __launch_bounds__(256) pack2_kernel(PackArgs A) {
  const int lane = threadIdx.x & 31;
  const int gsize = BLOCK ? (int)blockDim.x : 32;
  const int grank = BLOCK ? (int)threadIdx.x : lane;
  const long long gid = BLOCK ? (long long)blockIdx.x : (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ngroups = BLOCK ? (long long)gridDim.x : (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = gid; i < A.n; i += ngroups) {
    const PairMeta m = A.pairs[i];
    const uint8_t* ps = A.ascii + (A.p_off[i] - A.base);
    const uint8_t* ts = A.ascii + (A.t_off[i] - A.base);
    const int pw = (m.plen + 15) >> 4, tw = (m.tlen + 15) >> 4;
    uint32_t* out = A.words + m.woff;
    bool bad = false;
    for (int w = grank; w < pw + tw; w += gsize) {
      const bool is_t = w >= pw;
      const int j = is_t ? w - pw : w;
      const int len = is_t ? m.tlen : m.plen;
      out[w] = pack16((is_t ? ts : ps) + 16 * j, min(16, len - 16 * j), bad);
    }
    const bool any_bad = BLOCK ? (__syncthreads_or(bad) != 0) : __any_sync(0xffffffffu, bad);
    if (any_bad && grank == 0) {
      /* byte side buffer slot: the scalar tiers read this pair as bytes (4 per word) */
      const long long bw = words_of(m.plen, m.tlen, 4);
      const unsigned long long off = atomicAdd(&A.counters->side_words, (unsigned long long)bw);
      atomicAdd(&A.counters->n_side, 1ull);
      A.pairs[i].woff = ~(long long)off;
    }
  }
}

__device__ __forceinline__ uint32_t upper4(const uint8_t* s, int nb) {
  uint32_t w = 0;
  for (int b = 0; b < 4; ++b) {
    if (b < nb) {
      uint32_t c = s[b];
      if (c >= 'a' && c <= 'z') c -= 32;     /* pywfa upper-cases before the C call (pywfa/align.pyx:431-435) */
      w |= c << (8 * b);
    }
  }
  return w;
}

/* Upper-cased bytes, 4 per word.  side = true: only the pairs flagged by pack2_kernel, into words2
 * at ~woff; side = false (byte mode): every pair, into words at woff. */
__global__ void __launch_bounds__(256) pack_bytes_kernel(PackArgs A, bool side) {
  const int lane = threadIdx.x & 31;
  const long long gid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ngroups = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = gid; i < A.n; i += ngroups) {
    const PairMeta m = A.pairs[i];
    if (side && m.woff >= 0) continue;
    const uint8_t* ps = A.ascii + (A.p_off[i] - A.base);
    const uint8_t* ts = A.ascii + (A.t_off[i] - A.base);
    const int pw = (m.plen + 3) >> 2, tw = (m.tlen + 3) >> 2;
    uint32_t* out = side ? A.words2 + ~m.woff : A.words + m.woff;
    for (int w = lane; w < pw + tw; w += 32) {
      const bool is_t = w >= pw;
      const int j = is_t ? w - pw : w;
      const int len = is_t ? m.tlen : m.plen;
      out[w] = upper4((is_t ? ts : ps) + 4 * j, min(4, len - 4 * j));
    }
  }
}

}  // namespace

int layout_tiles(long long n) { return (int)((n + LAY_TILE - 1) / LAY_TILE); }

cudaError_t launch_layout(const int32_t* p_len, const int32_t* t_len, long long n, int bases_per_word,
                          long long* tile_sums, PairMeta* pairs, const BucketArgs& B, uint32_t* zero_a, int n_zero_a,
                          uint32_t* zero_b, long long* gen_poff, long long* gen_toff, long long off_base, cudaStream_t st) {
  const int tiles = layout_tiles(n);
  Sum2* sums = reinterpret_cast<Sum2*>(tile_sums);
  if (n > 0) layout_count_kernel<<<tiles, LAY_THREADS, 0, st>>>(p_len, t_len, n, bases_per_word, sums);
  layout_scan_kernel<<<1, 1024, 0, st>>>(sums, tiles, zero_a, n_zero_a, zero_b);
  if (n <= 0) return cudaGetLastError();
  layout_apply_kernel<<<tiles, LAY_THREADS, 0, st>>>(p_len, t_len, n, bases_per_word, sums, pairs, B, gen_poff, gen_toff, off_base);
  return cudaGetLastError();
}

cudaError_t launch_pack(const PackArgs& A, bool byte_mode, int max_len, int sms, cudaStream_t st) {
  if (A.n <= 0) return cudaSuccess;
  if (byte_mode) {
    const long long blocks = std::min<long long>((A.n + 7) / 8, (long long)sms * 8);
    pack_bytes_kernel<<<(int)blocks, 256, 0, st>>>(A, false);
  } else if (max_len > 4096) {
    const long long blocks = std::min<long long>(A.n, (long long)sms * 8);
    pack2_kernel<true><<<(int)blocks, 256, 0, st>>>(A);
  } else {
    const long long blocks = std::min<long long>((A.n + 7) / 8, (long long)sms * 8);
    pack2_kernel<false><<<(int)blocks, 256, 0, st>>>(A);
  }
  return cudaGetLastError();
}

__global__ void init_words_kernel(uint32_t* dst, int nwords, int at, SmallInts vals) {
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
    const int j = i - at;
    dst[i] = (j >= 0 && j < MAX_BUCKETS) ? (uint32_t)vals.v[j] : 0u;
  }
}
cudaError_t launch_init_words(uint32_t* dst, int nwords, int at, const SmallInts& vals, cudaStream_t st) {
  init_words_kernel<<<1, 256, 0, st>>>(dst, nwords, at, vals);
  return cudaGetLastError();
}

cudaError_t launch_pack_side(const PackArgs& A, int sms, cudaStream_t st) {
  if (A.n <= 0) return cudaSuccess;
  const long long blocks = std::min<long long>((A.n + 7) / 8, (long long)sms * 8);
  pack_bytes_kernel<<<(int)blocks, 256, 0, st>>>(A, true);
  return cudaGetLastError();
}

}  // namespace wfagpu
