/* wfa_pack16.cuh -- 16 ASCII bases -> one 2-bit packed word (device): shared by the batch packer
 * (wfa_pack.cu) and the single-pair kernel (wfa_kernels.cu).  Layout as pack.cpp's pack_sequence:
 * base j of a word in bits 2j..2j+1, code = (c >> 1) & 3 (A=0 C=1 T=2 G=3, either case). */
#pragma once
#include <stdint.h>

namespace wfagpu {

/* 16 bases -> one word.  `s` may have any alignment: two aligned 16-byte loads (only chunks that
 * hold at least one of the wanted bytes are touched, so nothing outside the caller's bytes' own
 * 16-byte lines is ever read) and a funnel shift.  Returns the packed word; `bad` is raised when a
 * byte other than ACGT/acgt was seen. */
/* (FRESH: the bytes live in host memory another call rewrote -- fetch them anew instead of through the read-only path) */
template <bool FRESH = false>
__device__ __forceinline__ uint32_t pack16(const uint8_t* s, int nb, bool& bad) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
  const uint4* al = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
  const int sh = (int)(a & 15);
  const uint4 lo = FRESH ? __ldcv(al) : __ldg(al);
  uint4 hi = make_uint4(0, 0, 0, 0);
  if (sh + nb > 16) hi = FRESH ? __ldcv(al + 1) : __ldg(al + 1);
  const int r = (sh & 3) * 8;
  uint32_t w0, w1, w2, w3;
  switch (sh >> 2) {
    case 0: w0 = __funnelshift_r(lo.x, lo.y, r); w1 = __funnelshift_r(lo.y, lo.z, r); w2 = __funnelshift_r(lo.z, lo.w, r); w3 = __funnelshift_r(lo.w, hi.x, r); break;
    case 1: w0 = __funnelshift_r(lo.y, lo.z, r); w1 = __funnelshift_r(lo.z, lo.w, r); w2 = __funnelshift_r(lo.w, hi.x, r); w3 = __funnelshift_r(hi.x, hi.y, r); break;
    case 2: w0 = __funnelshift_r(lo.z, lo.w, r); w1 = __funnelshift_r(lo.w, hi.x, r); w2 = __funnelshift_r(hi.x, hi.y, r); w3 = __funnelshift_r(hi.y, hi.z, r); break;
    default: w0 = __funnelshift_r(lo.w, hi.x, r); w1 = __funnelshift_r(hi.x, hi.y, r); w2 = __funnelshift_r(hi.y, hi.z, r); w3 = __funnelshift_r(hi.z, hi.w, r); break;
  }
  uint32_t w[4] = {w0, w1, w2, w3};
  uint32_t out = 0, ok = 0xffffffffu;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    /* bytes of this word that exist: nb - 4q of them (clamped to 0..4) */
    const int live = min(max(nb - 4 * q, 0), 4);
    const uint32_t keep = live >= 4 ? 0xffffffffu : ((1u << (8 * live)) - 1u);
    const uint32_t x = w[q] & keep;
    const uint32_t u = x & 0xDFDFDFDFu;
    const uint32_t good = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    ok &= good | ~keep;
    uint32_t c = (x >> 1) & 0x03030303u;
    c |= c >> 6;
    c = (c | (c >> 12)) & 0xffu;
    out |= c << (8 * q);
  }
  bad |= ok != 0xffffffffu;
  return out;
}


}  // namespace wfagpu
