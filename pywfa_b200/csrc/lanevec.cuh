/*
 * lanevec.cuh -- the per-lane value types and warp primitives the register-resident wavefront
 * tier (wfa_reg.cuh) is written against.
 *
 * On the device a `vi`/`vu`/`vb` is simply the calling thread's int / uint32_t / bool and every
 * primitive is one sm_100a instruction: VIMNMX.S16x2 / VIMNMX3.S16x2 / VIADD.16x2 (the DPX
 * packed-halfword integer ops), PRMT, SHFL, VOTE, LDS.  tests/emu/ supplies a 32-lane host
 * model of the same names (test infrastructure; the warp code is then executed lane-vector by
 * lane-vector on the CPU by the CPU test-suite), which is why wfa_reg.cuh contains no
 * per-lane control flow: every branch is warp-uniform, lanes differ only through selects.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(WFA_LANEVEC_HOST)

namespace wfagpu {
namespace lv {

typedef int vi;
typedef uint32_t vu;
typedef bool vb;

__device__ __forceinline__ vi lane_id() { return (int)(threadIdx.x & 31); }

/* ---- packed s16x2 (DPX) ---- */
__device__ __forceinline__ vu vimax2(vu a, vu b) { return __vmaxs2(a, b); }
__device__ __forceinline__ vu vimax3(vu a, vu b, vu c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ vu vadd2(vu a, vu b) { return __vadd2(a, b); }
__device__ __forceinline__ vu vsub2(vu a, vu b) { return __vsub2(a, b); }
/* max with "a >= b" predicates per half */
__device__ __forceinline__ vu vimax2p(vu a, vu b, vb& hi, vb& lo) { return __vibmax_s16x2(a, b, &hi, &lo); }
/* 0xFFFF in every half whose sign bit is set (PRMT with sign replication) */
/* (inline PTX: __byte_perm masks the selector to 3 bits per nibble and loses the replication flag) */
__device__ __forceinline__ vu signmask2(vu a) {
  vu d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(0xbb99u));
  return d;
}
__device__ __forceinline__ vu bitsel(vu mask, vu a, vu b) { return (a & mask) | (b & ~mask); }
__device__ __forceinline__ vu prmt(vu a, vu b, vu sel) { return __byte_perm(a, b, sel); }
__device__ __forceinline__ vi sx_lo(vu a) { return (int)(short)(a & 0xffffu); }
__device__ __forceinline__ vi sx_hi(vu a) { return ((int)a) >> 16; }
__device__ __forceinline__ vu pack2(vi lo, vi hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
/* replace one half of a packed register by the low 16 bits of v */
__device__ __forceinline__ vu put_lo(vu a, vi v) { return __byte_perm(a, (uint32_t)v, 0x3254); }
__device__ __forceinline__ vu put_hi(vu a, vi v) { return __byte_perm(a, (uint32_t)v, 0x5410); }

/* ---- lane exchange ---- */
__device__ __forceinline__ vu from_prev_lane(vu a) { return __shfl_sync(0xffffffffu, a, (threadIdx.x + 31) & 31); }
__device__ __forceinline__ vu from_next_lane(vu a) { return __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31); }
/* the value of lane `src` (SHFL.IDX takes the index modulo 32): rotations with a source index kept in a register */
__device__ __forceinline__ vu from_lane(vu a, vi src) { return __shfl_sync(0xffffffffu, a, src); }
__device__ __forceinline__ int lane_value(vi a, int lane) { return __shfl_sync(0xffffffffu, a, lane); }
__device__ __forceinline__ uint32_t ballot(vb p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ bool any(vb p) { return __any_sync(0xffffffffu, p); }

/* ---- per-lane integer helpers ---- */
__device__ __forceinline__ vi vsel(vb p, vi a, vi b) { return p ? a : b; }
__device__ __forceinline__ vu vselu(vb p, vu a, vu b) { return p ? a : b; }
__device__ __forceinline__ vi vmin(vi a, vi b) { return a < b ? a : b; }
__device__ __forceinline__ vi vmax(vi a, vi b) { return a > b ? a : b; }
__device__ __forceinline__ vi vffs0(vu x) { return __ffs((int)x) - 1; }
__device__ __forceinline__ vu vfunnel_r(vu lo, vu hi, vi sh) { return __funnelshift_r(lo, hi, sh); }
__device__ __forceinline__ vi as_vi(vu a) { return (int)a; }
__device__ __forceinline__ vu as_vu(vi a) { return (uint32_t)a; }
__device__ __forceinline__ vb vnot(vb a) { return !a; }
__device__ __forceinline__ vb vfalse() { return false; }
__device__ __forceinline__ vu splat(uint32_t x) { return x; }
__device__ __forceinline__ vi splati(int x) { return x; }

__device__ __forceinline__ vi vclz(vu x) { return __clz((int)x); }
/* 4-bit symbol codes, bit 0 of a nibble = "not the wildcard": 0xF in every nibble in which neither operand is the wildcard */
__device__ __forceinline__ vu nib_both(vu a, vu b) { return (a & b & 0x11111111u) * 15u; }
__device__ __forceinline__ vu vbrev(vu x) { return __brev(x); }

/* ---- memory ---- */
/* a window array in shared memory: its 32-bit shared-window address */
typedef uint32_t seqref;
__device__ __forceinline__ seqref make_seqref(const uint32_t* smem_ptr) { return (uint32_t)__cvta_generic_to_shared(smem_ptr); }
/* predicated LDS: lanes with p == false read nothing and get 0 (no branch, no reconvergence) */
__device__ __forceinline__ vu load_win(seqref base, vi idx, vb p) {
  vu r;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}"
               : "=r"(r) : "r"(base + 4u * (uint32_t)idx), "r"((uint32_t)p) : "memory");
  return r;
}
/* a per-lane position in a window array: the shared-window address of element idx0; element
 * idx0 + idx + IMM is then one address add away, IMM folded into the LDS immediate */
typedef uint32_t lanead;
__device__ __forceinline__ lanead lane_addr(seqref base, vi idx0) { return base + 4u * (uint32_t)idx0; }
/* predicated LDS of element idx + IMM; lanes with p == false read nothing and get `dflt` */
template <int IMM>
__device__ __forceinline__ vu load_win_at(lanead a, vi idx, vb p, uint32_t dflt) {
  vu r;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, %4;\n\t@q ld.shared.u32 %0, [%1+%3];\n\t}"
               : "=r"(r) : "r"(a + 4u * (uint32_t)idx), "r"((uint32_t)p), "n"(4 * IMM), "r"(dflt) : "memory");
  return r;
}
/* idx >= 0 ? a[idx + IA] ^ b[idx + IB] : 0x80000000 -- the first compare of an extension: one SETP on the offset itself,
 * two predicated LDS and a predicated XOR (a null cell loads nothing and sees "first base differs") */
template <int IA, int IB>
__device__ __forceinline__ vu diff_win_nonneg(lanead a, lanead b, vi idx) {
  vu r;
  asm volatile("{\n\t.reg .pred q;\n\t.reg .u32 x, y;\n\tsetp.ge.s32 q, %3, 0;\n\tmov.u32 %0, 0x80000000;\n\t"
               "@q ld.shared.u32 x, [%1+%4];\n\t@q ld.shared.u32 y, [%2+%5];\n\t@q xor.b32 %0, x, y;\n\t}"
               : "=r"(r) : "r"(a + 4u * (uint32_t)idx), "r"(b + 4u * (uint32_t)idx), "r"(idx), "n"(4 * IA), "n"(4 * IB) : "memory");
  return r;
}
/* min(a + b, c): VIADDMNMX */
__device__ __forceinline__ vi vaddmin(vi a, vi b, vi c) { return __viaddmin_s32(a, b, c); }
/* orders the warp's shared-memory accesses: what other lanes stored before is visible to the lane that reads after */
__device__ __forceinline__ void fence_warp() { __syncwarp(); }
/* pin a loop-invariant value in its register: the compiler may neither re-derive it nor fold it away */
__device__ __forceinline__ void keep(uint32_t& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void keep(int& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void scatter_u32(uint32_t* base, vi idx, vu val, vb p) { if (p) base[idx] = val; }
/* word gather from (shared) memory; lanes with p == false read nothing and get 0 */
__device__ __forceinline__ vu gather_u32(const uint32_t* base, vi idx, vb p) { return p ? base[idx] : 0u; }
__device__ __forceinline__ void scatter_u8(uint8_t* base, vi idx, vi val, vb p) { if (p) base[idx] = (uint8_t)val; }
/* a byte arena of one warp, in shared (SH accessors: STS.U8 / LDS.U8) or global memory */
struct histref { uint8_t* p; uint32_t s; };
__device__ __forceinline__ histref make_histref(uint8_t* p, bool shared) {
  histref h; h.p = p; h.s = shared ? (uint32_t)__cvta_generic_to_shared(p) : 0u; return h;
}
template <bool SH>
__device__ __forceinline__ void hist_store(const histref& h, int off, vi idx, vi val) {
  if (SH) asm volatile("st.shared.u8 [%0], %1;" :: "r"(h.s + (uint32_t)(off + idx)), "r"(val) : "memory");
  else h.p[off + idx] = (uint8_t)val;
}
template <bool SH>
__device__ __forceinline__ int hist_load(const histref& h, int off) {
  if (SH) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(h.s + (uint32_t)off) : "memory"); return (int)v; }
  return h.p[off];
}

}  // namespace lv
}  // namespace wfagpu

#endif

/* ---- scalar helpers shared by the device code and the host model ---- */
#ifdef __CUDACC__
#define WFA_LV_HD __host__ __device__ __forceinline__
#else
#define WFA_LV_HD inline
#endif
namespace wfagpu {
namespace lv {

/*
 * Byte mode on the register tier: 4-bit symbol codes, 8 bases per word (base b of a word in bits 4b..4b+3).
 * The wildcard (pywfa's wildcard= byte, pywfa/align.pyx:297-304) is code 0; A C G T N R Y K are 8..15 (bit 3 =
 * "not the wildcard"); any other byte makes the pair ineligible (`bad`; it runs on the scalar tiers).
 * w0 / w1: upper-cased bytes 0-3 / 4-7 of the word, nvalid: how many of them lie inside the sequence.
 */
WFA_LV_HD uint32_t nib_pack8(uint32_t w0, uint32_t w1, int nvalid, uint32_t wild, bool& bad) {
  const unsigned long long LO = (8ull << 0) | (9ull << 8) | (10ull << 24) | (15ull << 40) | (12ull << 52);   /* A C G K N */
  const unsigned long long HI = (13ull << 4) | (11ull << 12) | (14ull << 32);                                /* R T Y */
  uint32_t out = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const uint32_t c = ((b < 4 ? w0 : w1) >> (8 * (b & 3))) & 0xffu;
    const uint32_t idx = c - 'A';
    const unsigned long long tb = (idx & 16u) ? HI : LO;
    uint32_t code = idx < 26u ? (uint32_t)(tb >> ((idx & 15u) * 4u)) & 15u : 0u;
    const bool is_wild = wild != 0u && c == wild;
    if (b < nvalid) { if (is_wild) code = 0u; else if (code == 0u) bad = true; }
    else code = 0u;
    out |= code << (4 * b);
  }
  return out;
}

/* per lane: the 8 bases of one nibble word; `bad` collects "a byte outside the symbol set" */
#if defined(__CUDACC__) && !defined(WFA_LANEVEC_HOST)
__device__ __forceinline__ vu nib8(vu w0, vu w1, vi nvalid, uint32_t wild, vb& bad) {
  bool b = false;
  const vu r = nib_pack8(w0, w1, nvalid, wild, b);
  bad = bad || b;
  return r;
}
#elif defined(WFA_LANEVEC_HOST)
inline vu nib8(const vu& w0, const vu& w1, const vi& nvalid, uint32_t wild, vb& bad) {
  vu r;
  for (int l = 0; l < 32; ++l) {
    bool b = false;
    r.v[l] = nib_pack8(w0.v[l], w1.v[l], nvalid.v[l], wild, b);
    if (b) bad.m |= 1u << l;
  }
  return r;
}
#endif

}  // namespace lv
}  // namespace wfagpu
