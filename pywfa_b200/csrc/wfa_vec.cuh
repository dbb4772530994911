/*
 * wfa_vec.cuh -- packed-halfword wavefront tier for the general case: gap-affine and
 * gap-affine-2p, any penalties, end-to-end / ends-free, WF-adaptive and X-drop cut-offs, score
 * only or full CIGAR, reads up to VEC_MAX_LEN bases.  One warp (NW = 1) or one CTA of NW warps
 * aligns one read pair.
 *
 * Same reference path as wfa_core.cuh (W/ = pywfa/WFA2_lib/): wavefront_compute_affine_idm
 * (W/wavefront/wavefront_compute_affine.c:44-86), wavefront_compute_affine2p_idm
 * (W/wavefront/wavefront_compute_affine2p.c:45-106), the extension
 * (W/wavefront/wavefront_extend_kernels.c:64-163), trim_ends
 * (W/wavefront/wavefront_compute.c:571-605), termination
 * (W/wavefront/wavefront_termination.c:37-162), the cut-offs
 * (W/wavefront/wavefront_heuristic.c:257-383,509-567) and the backtrace
 * (W/wavefront/wavefront_backtrace.c:320-529).
 *
 * Layout.  Offsets are signed 16-bit, two ADJACENT diagonals per 32-bit word: word w of a
 * wavefront holds diagonals 2w and 2w+1 (biased so that w >= 0).  A block is 32 words = 64
 * diagonals = one warp instruction.  Every score's M / I1 / D1 (/ I2 / D2) wavefront lives in a
 * ring slot of `wcap` diagonals in shared memory, indexed circularly by word, so a step is:
 *   LDS own / left / right words of the 4 (7) source wavefronts, PRMT to the k-1 / k+1 views,
 *   VIMNMX.S16x2 / VIADD.16x2 recurrence on 64 cells per instruction, out-of-matrix and negative
 *   M offsets nulled with one packed add + sign-replicating PRMT, extension of the two M cells
 *   of the lane (16 bases per XOR of two funnel-shifted words of the 2-bit packed sequences),
 *   STS of the new words, one origin byte per cell (scope=full) as a 2-byte coalesced store.
 * Blocks whose neighbourhood lies inside every source's written range take a path without any
 * range checks.
 *
 * Ranges without reductions.  While no offset has touched the border of the DP matrix, every
 * offset computed from a valid source is valid, so the [lo,hi] of every component that the
 * reference obtains by trimming (compute.c:571-605) follows from the source ranges alone
 * (every thread derives them, nobody communicates) and a step needs ONE group barrier.  Once an
 * extended M offset reaches a sequence end (a sticky, group-wide flag) the step switches to the
 * exact variant: first / last in-matrix cell per component by ballots and shared-memory
 * atomics, and I/D offsets beyond the matrix that end up outside the trimmed range ("poison"
 * the reference drops with the range) are nulled in place.  The cut-offs work the same way:
 * cells they drop are nulled in place, so "outside [lo,hi] reads as NULL"
 * (compute.c:490-567) needs no per-cell range check.
 *
 * Device-only file (no host model): parity is pinned by the GPU suite against the CPU checker.
 */
#pragma once
#include <stdint.h>
#include <limits.h>

#include "lanevec.cuh"
#include "wfa_core.cuh"

#if defined(__CUDACC__)

namespace wfagpu {
namespace vec {

constexpr int BIAS = 1 << 20;              /* diagonal bias: multiple of 64, k + BIAS >= 0 */
constexpr uint32_t NULL2 = 0xC000C000u;    /* two int16 nulls */
constexpr int NULL16 = -16384;
constexpr int UB_MIN = -8192;              /* floor of ub[k] left of the matrix */
constexpr uint32_t ONE2 = 0x00010001u;
#ifndef WFA_VEC_PAD
#define WFA_VEC_PAD 1
#endif
/* WFA_VEC_BYTES (defined by wfa_vec_bytes.cu only): byte mode.  The sequence windows hold 8 four-bit symbol codes
 * per word (lv::nib_pack8: wildcard 0, A C G T N R Y K 8..15) instead of 16 two-bit codes, every group size uses the
 * windows, and the replay of the backtrace compares the pair's bytes (bpw / btw: 4 per word). */
#ifndef WFA_VEC_BYTES
#define WFA_VEC_BYTES 0
#endif
#ifndef WFA_VEC_EXT2
#define WFA_VEC_EXT2 0      /* interleaving the two extensions of a lane measured slower (r01: -6 % on cfg3) */
#endif
#define VEC_WIN(NW) ((NW) > 1 || WFA_VEC_BYTES != 0)   /* the window variant of the extension is compiled in */
constexpr int RENORM_MASK = 4095;          /* I/D nulls drift by one per step: re-based every 4096 scores */

/* per-step reduction cells in shared memory (three rotating sets) */
enum { F_TERM = 0, F_EDGE = 1, F_POISON = 2, F_LO = 3, F_HI = 8, F_MINA = 13, F_MINB = 14, F_MAXA = 15, F_MAXB = 16, NFLAG = 20 };
__device__ __forceinline__ int flag_init(int i) {
  if (i == F_EDGE || i == F_POISON || i > F_MAXB) return 0;
  if ((i >= F_HI && i < F_HI + 5) || i == F_MAXA || i == F_MAXB) return INT_MIN;
  return INT_MAX;
}

struct VMem {
  int4* meta;            /* [mr][3]: {mlo, mhi, wblo, wbhi}, {i1lo, i1hi, d1lo, d1hi}, {i2lo, i2hi, d2lo, d2hi} */
  int* flags;            /* [3][NFLAG] */
  const uint32_t* pw; const uint32_t* tw;     /* shared memory: 2-bit packed sequences (+1 pad word), or per-base windows */
  const uint32_t* bpw; const uint32_t* btw;   /* 2-bit packed sequences for the backtrace (any memory) */
  int seqw;                                   /* 1: pw / tw are windows (16 bases from every position, first base in the top bits) */
  uint32_t* ring;        /* slots x capw words */
  struct PlanOut* plan;  /* [2], NW > 1 only */
  uint8_t* h_code; HistRow* hmeta; uint32_t* runs_stage; uint8_t* ops; int opcap;
};

struct VSrc { int off; int w0; unsigned span; };     /* ring word offset of the slot; written word range [w0, w0 + span] */

template <bool CHECK>
__device__ __forceinline__ uint32_t vld(const uint32_t* ring, const VSrc& s, int pos, int wi) {
  if (CHECK) return ((unsigned)(wi - s.w0) <= s.span) ? ring[s.off + pos] : NULL2;
  return ring[s.off + pos];
}

/* everything a step needs that is uniform across the group */
struct __align__(16) PlanOut {
  VSrc mx, mo1, i1e, d1e, mo2, i2e, d2e;     /* sources */
  int oM, oI1, oD1, oI2, oD2;                /* output slots as ring word offsets (-1: component does not exist) */
  uint8_t* hrow;                             /* scope=full: origin bytes of the score, indexed by biased diagonal */
  long long cell_off;                        /* history cursor after this score */
  int nblo, nbhi;                            /* block range written by this score (null step: empty, nblo = tracked block) */
  int tp;                                    /* ring position (words) of block nblo */
  int fl, fh;                                /* blocks in [fl, fh] need no range checks */
  int kind;                                  /* 0 compute, 1 null step, 2 capacity exceeded */
  int exact;                                 /* ranges must be scanned from this score on */
  int pad;                                   /* 1: one all-null block is written either side of [nblo, nbhi] */
  int lo[5], hi[5];                          /* ranges of the valid cells (derived; replaced by the scan when exact) */
};
static_assert(sizeof(PlanOut) == 192, "PlanOut is copied through shared memory as 12 int4");

/* constants of the pair */
struct VCtx {
  uint32_t* ring;
  const uint32_t *pw, *tw;
  int plen, tlen, capw, ak;
  int endsfree, pef, tef;
  int seqw;
};

template <int NW> __device__ __forceinline__ void gsync() {
  if (NW == 1) __syncwarp(); else __syncthreads();
}

/* first / last set cell of a 64-cell block given the ballots of its even and odd cells */
__device__ __forceinline__ void report_range(int* F, int c, uint32_t b0, uint32_t b1, int kblock, int lane) {
  if ((b0 | b1) == 0) return;
  int first = 99, last = -1;
  if (b0) { first = 2 * first_set(b0); last = 2 * last_set(b0); }
  if (b1) { first = imin(first, 2 * first_set(b1) + 1); last = imax(last, 2 * last_set(b1) + 1); }
  if (lane == 0) { atomicMin(&F[F_LO + c], kblock + first); atomicMax(&F[F_HI + c], kblock + last); }
}

/* extension of one valid M offset; rem = bases left on the diagonal */
template <bool WIN>
__device__ __forceinline__ int vext(const uint32_t* pw, const uint32_t* tw, int seqw, int v, int h, int rem) {
  int n = 0;
  if (WIN && seqw) {
#if WFA_VEC_BYTES
    /* 8 symbol codes per window; a position where either side is the wildcard (code 0) never differs */
    while (n < rem) {
      const uint32_t a = pw[v + n], b = tw[h + n];
      const int c = __clz((int)((a ^ b) & ((a & b & 0x11111111u) * 15u))) >> 2;
      n += c;
      if (c < 8) break;
    }
#else
    /* per-base windows: 16 bases = LDS, LDS, XOR, CLZ */
    while (n < rem) {
      const int a = __clz((int)(pw[v + n] ^ tw[h + n])) >> 1;
      n += a;
      if (a < 16) break;
    }
#endif
    return imin(n, rem);
  }
  while (n < rem) {
    const uint32_t x = fetch16(pw, v + n) ^ fetch16(tw, h + n);
    if (x) { n += first_set(x) >> 1; break; }
    n += 16;
  }
  return imin(n, rem);
}

/* extension of the lane's two M offsets, interleaved so that their loads overlap;
 * rem = bases left on the diagonal (0 for a null offset) */
__device__ __forceinline__ void vext2(const uint32_t* pw, const uint32_t* tw, int v0, int& o0, int rem0, int v1, int& o1, int rem1) {
  int n0 = 0, n1 = 0;
  bool m0 = rem0 > 0, m1 = rem1 > 0;
  while (m0 || m1) {
    uint32_t x0 = 0, x1 = 0;
    if (m0) x0 = fetch16(pw, v0 + n0) ^ fetch16(tw, o0 + n0);
    if (m1) x1 = fetch16(pw, v1 + n1) ^ fetch16(tw, o1 + n1);
    if (m0) { const int a = x0 ? (first_set(x0) >> 1) : 16; n0 += a; m0 = (a == 16) && n0 < rem0; }
    if (m1) { const int a = x1 ? (first_set(x1) >> 1) : 16; n1 += a; m1 = (a == 16) && n1 < rem1; }
  }
  o0 += imin(n0, rem0); o1 += imin(n1, rem1);
}

/*
 * Extend the two M cells of the lane, detect matrix-edge contact / termination, store the word.
 * k0 = diagonal of the low half, u0 / u1 = ub of the two diagonals.
 */
template <bool WIN>
__device__ __forceinline__ void finish_m(const VCtx& c, uint32_t* oM, int* F, bool exact, uint32_t Mn, int pos, int k0, int u0, int u1, int kblock, int lane) {
  using namespace lv;
  int o0 = sx_lo(Mn), o1 = sx_hi(Mn);
  const bool v0 = o0 >= 0, v1 = o1 >= 0;
#if WFA_VEC_EXT2
  vext2(c.pw, c.tw, o0 - k0, o0, v0 ? u0 - o0 : 0, o1 - k0 - 1, o1, v1 ? u1 - o1 : 0);
#else
  o0 += vext<WIN>(c.pw, c.tw, c.seqw, o0 - k0, o0, v0 ? u0 - o0 : 0);
  o1 += vext<WIN>(c.pw, c.tw, c.seqw, o1 - k0 - 1, o1, v1 ? u1 - o1 : 0);
#endif
  oM[pos] = pack2(o0, o1);
  const bool e0 = v0 && o0 == u0, e1 = v1 && o1 == u1;
  if (__any_sync(0xffffffffu, e0 || e1)) {
    if (e0 || e1) F[F_EDGE] = 1;
    if (c.endsfree) {                              /* termination.c:115-162: lowest terminating diagonal */
      int tk = KNONE;
      if (e1) { const int vv = o1 - k0 - 1; if ((o1 >= c.tlen && c.plen - vv <= c.pef) || (vv >= c.plen && c.tlen - o1 <= c.tef)) tk = k0 + 1; }
      if (e0) { const int vv = o0 - k0; if ((o0 >= c.tlen && c.plen - vv <= c.pef) || (vv >= c.plen && c.tlen - o0 <= c.tef)) tk = k0; }
      if (tk != KNONE) atomicMin(&F[F_TERM], tk);
    } else {                                       /* termination.c:37-61 */
      if (k0 == c.ak && e0 && o0 >= c.tlen) F[F_TERM] = c.ak;
      if (k0 + 1 == c.ak && e1 && o1 >= c.tlen) F[F_TERM] = c.ak;
    }
  }
  if (exact) report_range(F, CM, __ballot_sync(0xffffffffu, o0 >= 0), __ballot_sync(0xffffffffu, o1 >= 0), kblock, lane);
}

/* exact mode: in-matrix ballots of an I/D word; flags offsets beyond the matrix */
__device__ __forceinline__ void scan_gap(int* F, int comp, uint32_t x, uint32_t nub, int kblock, int lane) {
  using namespace lv;
  const uint32_t t = vadd2(x, nub);                /* sign set: x <= ub */
  const uint32_t ok = t & ~x, bad = ~t & ~x;       /* sign bits: valid / non-negative but beyond the matrix */
  report_range(F, comp, __ballot_sync(0xffffffffu, (ok & 0x8000u) != 0), __ballot_sync(0xffffffffu, (ok & 0x80000000u) != 0), kblock, lane);
  if (bad & 0x80008000u) F[F_POISON] = 1;
}

/* one block of 64 diagonals of the recurrence */
template <bool TWO_P, bool FULL, bool CHECK, bool WIN>
__device__ __forceinline__ void vec_block(const VCtx& c, const PlanOut& pl, int* F, bool exact, int b, int posb, int lane) {
  const uint32_t* const ring = c.ring;
  using namespace lv;
  const int wi = 32 * b + lane;
  const int pos = posb + lane;
  int posl = pos - 1; if (posl < 0) posl += c.capw;
  int posr = pos + 1; if (posr >= c.capw) posr -= c.capw;
  const int kblock = 64 * b - BIAS;
  const int k0 = kblock + 2 * lane;
  const int u0 = imax(imin(c.tlen, c.plen + k0), UB_MIN), u1 = imax(imin(c.tlen, c.plen + k0 + 1), UB_MIN);
  const uint32_t nub = ~pack2(u0, u1);

  const uint32_t mx = vld<CHECK>(ring, pl.mx, pos, wi);
  const uint32_t mo1 = vld<CHECK>(ring, pl.mo1, pos, wi);
  const uint32_t MoL = prmt(vld<CHECK>(ring, pl.mo1, posl, wi - 1), mo1, 0x5432u);
  const uint32_t MoR = prmt(mo1, vld<CHECK>(ring, pl.mo1, posr, wi + 1), 0x5432u);
  const uint32_t IeL = prmt(vld<CHECK>(ring, pl.i1e, posl, wi - 1), vld<CHECK>(ring, pl.i1e, pos, wi), 0x5432u);
  const uint32_t DeR = prmt(vld<CHECK>(ring, pl.d1e, pos, wi), vld<CHECK>(ring, pl.d1e, posr, wi + 1), 0x5432u);
  const uint32_t mis = vadd2(mx, ONE2);
  uint32_t ins1, del1, ins2 = NULL2, del2 = NULL2, m;
  bool x1h = false, x1l = false, y1h = false, y1l = false, x2h = false, x2l = false, y2h = false, y2l = false;
  if (FULL) {
    ins1 = vadd2(vimax2p(IeL, MoL, x1h, x1l), ONE2);       /* ext >= open: extension (backtrace.c:49-59) */
    del1 = vimax2p(DeR, MoR, y1h, y1l);
  } else {
    ins1 = vadd2(vimax2(IeL, MoL), ONE2);
    del1 = vimax2(DeR, MoR);
  }
  if (TWO_P) {
    const uint32_t mo2 = vld<CHECK>(ring, pl.mo2, pos, wi);
    const uint32_t Mo2L = prmt(vld<CHECK>(ring, pl.mo2, posl, wi - 1), mo2, 0x5432u);
    const uint32_t Mo2R = prmt(mo2, vld<CHECK>(ring, pl.mo2, posr, wi + 1), 0x5432u);
    const uint32_t I2L = prmt(vld<CHECK>(ring, pl.i2e, posl, wi - 1), vld<CHECK>(ring, pl.i2e, pos, wi), 0x5432u);
    const uint32_t D2R = prmt(vld<CHECK>(ring, pl.d2e, pos, wi), vld<CHECK>(ring, pl.d2e, posr, wi + 1), 0x5432u);
    if (FULL) {
      ins2 = vadd2(vimax2p(I2L, Mo2L, x2h, x2l), ONE2);
      del2 = vimax2p(D2R, Mo2R, y2h, y2l);
    } else {
      ins2 = vadd2(vimax2(I2L, Mo2L), ONE2);
      del2 = vimax2(D2R, Mo2R);
    }
  }
  if (FULL) {
    /* winner of the reference's max over (offset << 4 | type): on ties M > D2 > D1 > I2 > I1 */
    bool p1h = false, p1l = false, p2h, p2l, p3h = false, p3l = false, p4h, p4l;
    uint32_t r = ins1;
    if (TWO_P) r = vimax2p(ins2, r, p1h, p1l);
    r = vimax2p(del1, r, p2h, p2l);
    if (TWO_P) r = vimax2p(del2, r, p3h, p3l);
    m = vimax2p(mis, r, p4h, p4l);
    /* origin byte = the eight predicates as they are (the backtrace decodes the priority) */
    const uint32_t cl = (p1l ? 1u : 0u) | (p2l ? 2u : 0u) | (p3l ? 4u : 0u) | (p4l ? 8u : 0u) |
                        (x1l ? 0x10u : 0u) | (y1l ? 0x20u : 0u) | (x2l ? 0x40u : 0u) | (y2l ? 0x80u : 0u);
    const uint32_t ch = (p1h ? 1u : 0u) | (p2h ? 2u : 0u) | (p3h ? 4u : 0u) | (p4h ? 8u : 0u) |
                        (x1h ? 0x10u : 0u) | (y1h ? 0x20u : 0u) | (x2h ? 0x40u : 0u) | (y2h ? 0x80u : 0u);
    *reinterpret_cast<uint16_t*>(pl.hrow + (64 * b + 2 * lane)) = (uint16_t)(cl | (ch << 8));
  } else if (TWO_P) {
    m = vimax3(vimax3(mis, ins1, ins2), del1, del2);
  } else {
    m = vimax3(mis, ins1, del1);
  }
  /* keep 0 <= M <= ub, everything else becomes the null (compute_affine.c:80-84) */
  const uint32_t keep = signmask2(vadd2(m, nub) & ~m);
  const uint32_t Mn = (m & keep) | (NULL2 & ~keep);
  if (pl.oI1 >= 0) c.ring[pl.oI1 + pos] = ins1;
  if (pl.oD1 >= 0) c.ring[pl.oD1 + pos] = del1;
  if (TWO_P) {
    if (pl.oI2 >= 0) c.ring[pl.oI2 + pos] = ins2;
    if (pl.oD2 >= 0) c.ring[pl.oD2 + pos] = del2;
  }
  if (exact) {
    if (pl.oI1 >= 0) scan_gap(F, CI1, ins1, nub, kblock, lane);
    if (pl.oD1 >= 0) scan_gap(F, CD1, del1, nub, kblock, lane);
    if (TWO_P) {
      if (pl.oI2 >= 0) scan_gap(F, CI2, ins2, nub, kblock, lane);
      if (pl.oD2 >= 0) scan_gap(F, CD2, del2, nub, kblock, lane);
    }
  }
  finish_m<WIN>(c, c.ring + pl.oM, F, exact, Mn, pos, k0, u0, u1, kblock, lane);
}

/* null the cells of one ring word that lie outside [lo, hi] */
__device__ __forceinline__ void clip_word(uint32_t* slot, int pos, int k0, int lo, int hi) {
  const uint32_t mask = ((k0 >= lo && k0 <= hi) ? 0x0000ffffu : 0u) | ((k0 + 1 >= lo && k0 + 1 <= hi) ? 0xffff0000u : 0u);
  if (mask != 0xffffffffu) { const uint32_t x = slot[pos]; slot[pos] = (x & mask) | (NULL2 & ~mask); }
}

/* Backtrace over the origin bytes of this tier (one thread): bits 0-3 = "this candidate >= the best of
 * the lower-priority ones" for I2, D1, D2, mismatch (the winner of M is the highest bit set, I1 if none),
 * bits 4-7 "extension >= opening" of I1, D1, I2, D2 at the cell. */
__device__ inline int backtrace_vcodes(const KParams& P, const uint8_t* h_code, const HistRow* hmeta, int a_score, int a_k,
                                       int plen, int tlen, const uint32_t* pw, const uint32_t* tw, uint8_t* ops, int opcap,
                                       FwdEmitter& em) {
  int mt = CM, score = a_score, k = a_k, nops = 0;
  while (score > 0) {
    const HistRow hm = hmeta[score];
    const int code = h_code[hm.off + (k - hm.lo)];
    int comp = mt;
    if (mt == CM) {
      /* winner of the max chain I1 < I2 < D1 < D2 < mismatch (later wins ties): highest predicate set */
      if (code & 8) { if (nops < opcap) ops[nops] = EOP_X; ++nops; score -= P.dx; continue; }
      comp = (code & 4) ? CD2 : (code & 2) ? CD1 : (code & 1) ? CI2 : CI1;
    }
    const bool ext = (code >> (comp == CI1 ? 4 : comp == CD1 ? 5 : comp == CI2 ? 6 : 7)) & 1;
    const bool is_ins = comp == CI1 || comp == CI2;
    const bool two = comp == CI2 || comp == CD2;
    if (nops < opcap) ops[nops] = (uint8_t)(is_ins ? (ext ? EOP_I_EXT : EOP_I_OPEN) : (ext ? EOP_D_EXT : EOP_D_OPEN));
    ++nops;
    if (ext) { score -= two ? P.de2 : P.de1; mt = comp; } else { score -= two ? P.doe2 : P.doe1; mt = CM; }
    if (is_ins) --k; else ++k;
  }
  if (nops > opcap) return -1;
  replay_ops(ops, nops, k, plen, tlen, pw, tw, em, WFA_VEC_BYTES ? P.wildcard : -1);
  return em.n;
}

/*
 * Plan score s: fetch_input (compute.c:298-344), limits (compute.c:40-86), allocate_output
 * (compute.c:401-486) and, while no offset has touched the matrix border, the trimmed ranges
 * (compute.c:571-605) derived from the source ranges.  A pure function of the metadata ring and
 * the scalars passed in, so that one warp can run it for score s+1 while the others still work
 * on score s.  `writer`: this thread publishes the metadata / history row of the score.
 */
template <bool TWO_P, bool FULL, bool PAD>
__device__ __forceinline__ void plan_step(const KParams& P, const VMem& vm, int plen, int tlen, int s, int cm, int c1, int c2,
                                          int tp, int tb, long long cell_off, bool exact, bool writer, PlanOut& o) {
  const int capw = P.wcap >> 1, nblk = P.wcap >> 6, mmask = P.mr - 1;
  const int4* const meta = vm.meta;
  const int rI1 = P.rm * capw, rD1 = rI1 + P.r1 * capw, rI2 = rD1 + P.r1 * capw, rD2 = rI2 + (TWO_P ? P.r2 * capw : 0);   /* ring word offsets */
  const int rNull = (P.rm + 2 * P.r1 + (TWO_P ? 2 * P.r2 : 0)) * capw;       /* the all-null slot follows the last ring slot */
  const int4 aMx = meta[((s - P.dx) & mmask) * 3];
  const int4 aMo1 = meta[((s - P.doe1) & mmask) * 3];
  const int4* const rowe1 = meta + ((s - P.de1) & mmask) * 3;
  const int4 aE1 = rowe1[0], bE1 = rowe1[1];
  int4 aMo2 = make_int4(1, -1, 1, -1), aE2 = aMo2, cE2 = aMo2;
  if (TWO_P) {
    aMo2 = meta[((s - P.doe2) & mmask) * 3];
    const int4* const rowe2 = meta + ((s - P.de2) & mmask) * 3;
    aE2 = rowe2[0]; cE2 = rowe2[2];
  }
  const bool n_mx = aMx.x > aMx.y, n_mo1 = aMo1.x > aMo1.y, n_i1 = bE1.x > bE1.y, n_d1 = bE1.z > bE1.w;
  const bool n_mo2 = TWO_P ? aMo2.x > aMo2.y : true;
  const bool n_i2 = TWO_P ? cE2.x > cE2.y : true;
  const bool n_d2 = TWO_P ? cE2.z > cE2.w : true;
  int4* const mrow = vm.meta + (s & mmask) * 3;
  o.cell_off = cell_off; o.tp = tp; o.nblo = tb; o.nbhi = tb - 1; o.exact = exact; o.fl = INT_MAX; o.fh = INT_MIN; o.pad = 0;
  for (int c = 0; c < 5; ++c) { o.lo[c] = 1; o.hi[c] = -1; }
  if (n_mx && n_mo1 && n_i1 && n_d1 && n_mo2 && n_i2 && n_d2) {
    /* null step: allocate_output_null, compute.c:374-400 */
    o.kind = 1;
    if (writer) { mrow[0] = make_int4(1, -1, 1, -1); mrow[1] = make_int4(1, -1, 1, -1); mrow[2] = make_int4(1, -1, 1, -1); }
    return;
  }
  int lo = INT_MAX, hi = INT_MIN, li1 = INT_MAX, hi1 = INT_MIN, ld1 = INT_MAX, hd1 = INT_MIN;
  int li2 = INT_MAX, hi2 = INT_MIN, ld2 = INT_MAX, hd2 = INT_MIN;
  int fl = INT_MIN, fh = INT_MAX;
  auto use = [&](const int4& a, bool null_) {
    if (null_) return;                       /* reads of a null source go to the all-null slot */
    fl = imax(fl, a.z + 1); fh = imin(fh, a.w - 1);
  };
  if (!n_mx) { lo = aMx.x; hi = aMx.y; }
  if (!n_mo1) { li1 = aMo1.x + 1; hi1 = aMo1.y + 1; ld1 = aMo1.x - 1; hd1 = aMo1.y - 1; }
  if (!n_i1) { li1 = imin(li1, bE1.x + 1); hi1 = imax(hi1, bE1.y + 1); }
  if (!n_d1) { ld1 = imin(ld1, bE1.z - 1); hd1 = imax(hd1, bE1.w - 1); }
  lo = imin(lo, imin(li1, ld1)); hi = imax(hi, imax(hi1, hd1));
  use(aMx, n_mx); use(aMo1, n_mo1); use(aE1, n_i1); use(aE1, n_d1);
  if (TWO_P) {
    if (!n_mo2) { li2 = aMo2.x + 1; hi2 = aMo2.y + 1; ld2 = aMo2.x - 1; hd2 = aMo2.y - 1; }
    if (!n_i2) { li2 = imin(li2, cE2.x + 1); hi2 = imax(hi2, cE2.y + 1); }
    if (!n_d2) { ld2 = imin(ld2, cE2.z - 1); hd2 = imax(hd2, cE2.w - 1); }
    lo = imin(lo, imin(li2, ld2)); hi = imax(hi, imax(hi2, hd2));
    use(aMo2, n_mo2); use(aE2, n_i2); use(aE2, n_d2);
  }
  if (lo < -plen || hi > tlen) exact = true;
  const int nblo = (lo + BIAS) >> 6, nbhi = (hi + BIAS) >> 6;
  const int rowlen = (nbhi - nblo + 1) << 6;
  if (nbhi - nblo + 1 > nblk || (FULL && (s >= P.scap || cell_off + rowlen > P.hcap))) { o.kind = 2; return; }
  /* one all-null block either side, ring capacity permitting: later scores, a few diagonals wider,
   * then find their whole neighbourhood inside the written range and skip the range checks */
  const int pad = (PAD && nbhi - nblo + 3 <= nblk) ? 1 : 0;
  tp += (nblo - tb) << 5;
  while (tp >= capw) tp -= capw;
  while (tp < 0) tp += capw;
  /* allocate_output, compute.c:401-486 */
  const bool has_i1 = !n_mo1 || !n_i1, has_d1 = !n_mo1 || !n_d1;
  const bool has_i2 = TWO_P && (!n_mo2 || !n_i2), has_d2 = TWO_P && (!n_mo2 || !n_d2);
  auto src = [&](VSrc& v, int off, const int4& a, bool null_) {
    if (null_) { v.off = rNull; v.w0 = INT_MAX; v.span = 0; }
    else { v.off = off; v.w0 = a.z << 5; v.span = (unsigned)(((a.w - a.z) << 5) + 31); }
  };
  auto mslot = [&](int d) { int sl = cm - d; if (sl < 0) sl += P.rm; return sl * capw; };
  const int e1s = (c1 + 1 == P.r1) ? 0 : c1 + 1;           /* slot of score s - e1 */
  src(o.mx, mslot(P.dx), aMx, n_mx);
  src(o.mo1, mslot(P.doe1), aMo1, n_mo1);
  src(o.i1e, rI1 + e1s * capw, aE1, n_i1);
  src(o.d1e, rD1 + e1s * capw, aE1, n_d1);
  if (TWO_P) {
    const int e2s = (c2 + 1 == P.r2) ? 0 : c2 + 1;
    src(o.mo2, mslot(P.doe2), aMo2, n_mo2);
    src(o.i2e, rI2 + e2s * capw, aE2, n_i2);
    src(o.d2e, rD2 + e2s * capw, aE2, n_d2);
  } else {
    o.mo2 = o.mo1; o.i2e = o.i1e; o.d2e = o.d1e;
  }
  o.oM = cm * capw;
  o.oI1 = has_i1 ? rI1 + c1 * capw : -1;
  o.oD1 = has_d1 ? rD1 + c1 * capw : -1;
  o.oI2 = has_i2 ? rI2 + c2 * capw : -1;
  o.oD2 = has_d2 ? rD2 + c2 * capw : -1;
  o.hrow = nullptr;
  if (FULL) {
    o.hrow = vm.h_code + cell_off - 64ll * nblo;
    if (writer) { HistRow hr; hr.off = cell_off; hr.lo = 64 * nblo - BIAS; hr.pad = 0; vm.hmeta[s] = hr; }
    o.cell_off = cell_off + rowlen;
  }
  o.kind = 0; o.exact = exact; o.nblo = nblo; o.nbhi = nbhi; o.tp = tp; o.fl = fl; o.fh = fh; o.pad = pad;
  o.lo[CM] = lo; o.hi[CM] = hi;
  if (has_i1) { o.lo[CI1] = li1; o.hi[CI1] = hi1; }
  if (has_d1) { o.lo[CD1] = ld1; o.hi[CD1] = hd1; }
  if (has_i2) { o.lo[CI2] = li2; o.hi[CI2] = hi2; }
  if (has_d2) { o.lo[CD2] = ld2; o.hi[CD2] = hd2; }
  if (writer && !exact) {
    /* derived ranges are final: publish them now (the scan publishes them otherwise) */
    mrow[0] = make_int4(lo, hi, nblo - pad, nbhi + pad);
    mrow[1] = make_int4(o.lo[CI1], o.hi[CI1], o.lo[CD1], o.hi[CD1]);
    mrow[2] = make_int4(o.lo[CI2], o.hi[CI2], o.lo[CD2], o.hi[CD2]);
  }
}

/*
 * plan_step spread over the lanes of one warp: lane i < 7 owns source i (Mx, Mo1, I1e, D1e, Mo2, I2e,
 * D2e), the ranges are warp min-reductions, and the plan goes straight to shared memory (lane i
 * stores its source, lane 0 the rest).  ~4x shorter dependent chain than the one-thread version,
 * which is what bounds a step when one warp plans for the whole group.  Must produce exactly what
 * plan_step produces (the forced-group-size parity tests run both).
 */
template <bool TWO_P, bool FULL, bool PAD>
__device__ __forceinline__ void plan_step_simd(const KParams& P, const VMem& vm, int plen, int tlen, int s, int cm, int c1, int c2,
                                               int tp, int tb, long long cell_off, bool exact, PlanOut* out) {
  const int lane = (int)(threadIdx.x & 31);
  const int capw = P.wcap >> 1, nblk = P.wcap >> 6, mmask = P.mr - 1;
  const int4* const meta = vm.meta;
  const int rI1 = P.rm * capw, rD1 = rI1 + P.r1 * capw, rI2 = rD1 + P.r1 * capw, rD2 = rI2 + (TWO_P ? P.r2 * capw : 0);
  const int rNull = (P.rm + 2 * P.r1 + (TWO_P ? 2 * P.r2 : 0)) * capw;
  /* lane -> source: look-back, metadata row / half holding its range, ring base and current slot */
  const int nsrc = TWO_P ? 7 : 4;
  const bool is_m = lane == 0 || lane == 1 || lane == 4;
  const int d = lane == 0 ? P.dx : lane == 1 ? P.doe1 : lane <= 3 ? P.de1 : lane == 4 ? P.doe2 : P.de2;
  const int rowsel = is_m ? 0 : (lane <= 3 ? 1 : 2);
  const bool second_half = lane == 3 || lane == 6;
  int slot_off;
  if (is_m) { int sl = cm - d; if (sl < 0) sl += P.rm; slot_off = sl * capw; }
  else if (lane <= 3) { const int e1s = (c1 + 1 == P.r1) ? 0 : c1 + 1; slot_off = (lane == 2 ? rI1 : rD1) + e1s * capw; }
  else { const int e2s = (c2 + 1 == P.r2) ? 0 : c2 + 1; slot_off = (lane == 5 ? rI2 : rD2) + e2s * capw; }
  const int row = ((s - d) & mmask) * 3;
  const int4 a = meta[row];
  const int4 r = meta[row + rowsel];
  const int lo_c = second_half ? r.z : r.x, hi_c = second_half ? r.w : r.y;
  const bool null_ = lane >= nsrc || lo_c > hi_c;
  /* contributions to the ranges */
  const bool to_m = lane == 0, to_i = lane == 1 || lane == 2 || lane == 4 || lane == 5, to_d = lane == 1 || lane == 3 || lane == 4 || lane == 6;
  const bool p2 = lane >= 4;
  const int BIG = INT_MAX;
  auto rmin = [](int v) { return __reduce_min_sync(0xffffffffu, v); };
  auto rmax = [](int v) { return __reduce_max_sync(0xffffffffu, v); };
  const bool ok = !null_;
  int lo = rmin(ok && to_m ? lo_c : BIG), hi = rmax(ok && to_m ? hi_c : INT_MIN);
  const int li1 = rmin(ok && to_i && !p2 ? lo_c + 1 : BIG), hi1 = rmax(ok && to_i && !p2 ? hi_c + 1 : INT_MIN);
  const int ld1 = rmin(ok && to_d && !p2 ? lo_c - 1 : BIG), hd1 = rmax(ok && to_d && !p2 ? hi_c - 1 : INT_MIN);
  int li2 = BIG, hi2 = INT_MIN, ld2 = BIG, hd2 = INT_MIN;
  if (TWO_P) {
    li2 = rmin(ok && to_i && p2 ? lo_c + 1 : BIG); hi2 = rmax(ok && to_i && p2 ? hi_c + 1 : INT_MIN);
    ld2 = rmin(ok && to_d && p2 ? lo_c - 1 : BIG); hd2 = rmax(ok && to_d && p2 ? hi_c - 1 : INT_MIN);
  }
  const int fl = rmax(ok ? a.z + 1 : INT_MIN), fh = rmin(ok ? a.w - 1 : BIG);
  const bool any_src = __any_sync(0xffffffffu, ok);
  int4* const mrow = vm.meta + (s & mmask) * 3;
  /* this lane's source */
  {
    VSrc v;
    if (null_) { v.off = rNull; v.w0 = INT_MAX; v.span = 0; }
    else { v.off = slot_off; v.w0 = a.z << 5; v.span = (unsigned)(((a.w - a.z) << 5) + 31); }
    if (lane < 7) (&out->mx)[lane] = v;
  }
  if (!any_src) {
    /* null step: allocate_output_null, compute.c:374-400 */
    if (lane == 0) {
      out->cell_off = cell_off; out->tp = tp; out->nblo = tb; out->nbhi = tb - 1; out->exact = exact;
      out->fl = INT_MAX; out->fh = INT_MIN; out->pad = 0; out->kind = 1;
      for (int c = 0; c < 5; ++c) { out->lo[c] = 1; out->hi[c] = -1; }
      mrow[0] = make_int4(1, -1, 1, -1); mrow[1] = make_int4(1, -1, 1, -1); mrow[2] = make_int4(1, -1, 1, -1);
    }
    return;
  }
  lo = imin(lo, imin(imin(li1, ld1), imin(li2, ld2)));
  hi = imax(hi, imax(imax(hi1, hd1), imax(hi2, hd2)));
  if (lo < -plen || hi > tlen) exact = true;
  const int nblo = (lo + BIAS) >> 6, nbhi = (hi + BIAS) >> 6;
  const int rowlen = (nbhi - nblo + 1) << 6;
  if (nbhi - nblo + 1 > nblk || (FULL && (s >= P.scap || cell_off + rowlen > P.hcap))) {
    if (lane == 0) out->kind = 2;
    return;
  }
  const int pad = (PAD && nbhi - nblo + 3 <= nblk) ? 1 : 0;
  tp += (nblo - tb) << 5;
  while (tp >= capw) tp -= capw;
  while (tp < 0) tp += capw;
  const bool has_i1 = li1 != BIG, has_d1 = ld1 != BIG, has_i2 = TWO_P && li2 != BIG, has_d2 = TWO_P && ld2 != BIG;
  if (lane == 0) {
    out->oM = cm * capw;
    out->oI1 = has_i1 ? rI1 + c1 * capw : -1;
    out->oD1 = has_d1 ? rD1 + c1 * capw : -1;
    out->oI2 = has_i2 ? rI2 + c2 * capw : -1;
    out->oD2 = has_d2 ? rD2 + c2 * capw : -1;
    out->hrow = nullptr; out->cell_off = cell_off;
    if (FULL) {
      out->hrow = vm.h_code + cell_off - 64ll * nblo;
      HistRow hr; hr.off = cell_off; hr.lo = 64 * nblo - BIAS; hr.pad = 0; vm.hmeta[s] = hr;
      out->cell_off = cell_off + rowlen;
    }
    out->kind = 0; out->exact = exact; out->nblo = nblo; out->nbhi = nbhi; out->tp = tp; out->fl = fl; out->fh = fh; out->pad = pad;
    const int l1 = has_i1 ? li1 : 1, h1 = has_i1 ? hi1 : -1, l2 = has_d1 ? ld1 : 1, h2 = has_d1 ? hd1 : -1;
    const int l3 = has_i2 ? li2 : 1, h3 = has_i2 ? hi2 : -1, l4 = has_d2 ? ld2 : 1, h4 = has_d2 ? hd2 : -1;
    out->lo[CM] = lo; out->hi[CM] = hi;
    out->lo[CI1] = l1; out->hi[CI1] = h1; out->lo[CD1] = l2; out->hi[CD1] = h2;
    out->lo[CI2] = l3; out->hi[CI2] = h3; out->lo[CD2] = l4; out->hi[CD2] = h4;
    if (!exact) {
      mrow[0] = make_int4(lo, hi, nblo - pad, nbhi + pad);
      mrow[1] = make_int4(l1, h1, l2, h2);
      mrow[2] = make_int4(l3, h3, l4, h4);
    }
  }
}

/*
 * Align one pair with a group of NW warps; HEUR = 0 none, 1 WF-adaptive, 2 X-drop (compiled apart: the
 * step loop has to stay small enough for the instruction cache).  Returns PAIR_DONE (res filled; scope=full: runs in
 * vm.runs_stage, res.nruns / res.locs valid on rank 0 only) or PAIR_OVERFLOW.
 */
template <bool TWO_P, bool FULL, int NW, int HEUR>
__device__ int align_pair_vec(const KParams& P, const VMem& vm, int plen, int tlen, PairResult& res) {
  constexpr int GS = NW * 32;
  constexpr int NC = TWO_P ? 5 : 3;
  constexpr int CW = NW > 1 ? NW - 1 : 1;       /* warps that take blocks; with NW > 1 the last warp only plans */
  const int lane = (int)(threadIdx.x & 31);
  const int warp = NW == 1 ? 0 : (int)(threadIdx.x >> 5);
  const int wofs = warp < CW ? warp : (1 << 28);     /* block offset of this warp (the planner warp never matches a range) */
  const int rank = NW == 1 ? lane : (int)threadIdx.x;
  const int capw = P.wcap >> 1, nblk = P.wcap >> 6, mmask = P.mr - 1;
  int4* const meta = vm.meta;
  uint32_t* const rM = vm.ring;
  uint32_t* const rI1 = rM + P.rm * capw;
  uint32_t* const rD1 = rI1 + P.r1 * capw;
  uint32_t* const rI2 = rD1 + P.r1 * capw;
  uint32_t* const rD2 = rI2 + (TWO_P ? P.r2 * capw : 0);
  const int ak = tlen - plen;

  VCtx cx;
  cx.ring = vm.ring; cx.pw = vm.pw; cx.tw = vm.tw; cx.plen = plen; cx.tlen = tlen; cx.capw = capw; cx.ak = ak;
  cx.endsfree = P.endsfree; cx.pef = P.pef; cx.tef = P.tef; cx.seqw = vm.seqw;

  int s = 0, cm = 0, c1 = 0, c2 = 0, fb = 0;
  int s_exist = 0, steps_wait = P.steps_between, max_sw = 0;
  bool sw_init = false, exact = false, cur_exists = true;
  bool sticky = false;                          /* scanned ranges for good: an offset touched the matrix border */
  int exact_until = -1;                         /* scanned ranges up to this score: a cut-off left a null end cell in a wavefront still in the ring */
  long long cells = 0, cell_off = 0;
  int clo[5], chi[5];
  int end_k = KNONE, end_off = OFFNULL, end_score = 0, status = 0;
  int tb, tp = 0;                               /* ring position tp (words) of block tb */
  int blo, bhi;                                 /* written block range of the current score */
  int cur_pad = 0;                              /* padding blocks of the current score */
  bool have_plan = false;                       /* vm.plan[(s+1)&1] holds the plan of the next score */
  const bool is_writer = NW == 1 ? lane == 0 : (warp == NW - 1 && lane == 0);   /* publishes metadata (the planner's lane 0) */

  /* ---- score 0: wavefront_aligner_init_wf_m, W/wavefront/wavefront_aligner.c:251-310 ---- */
  {
    const bool ef = P.endsfree && P.match == 0;
    const int lo0 = ef ? -P.pbf : 0, hi0 = ef ? P.tbf : 0;
    blo = (lo0 + BIAS) >> 6; bhi = (hi0 + BIAS) >> 6;
    if (bhi - blo + 1 > nblk) return PAIR_OVERFLOW;
    for (int i = rank; i < P.mr * 3; i += GS) meta[i] = make_int4(1, -1, 1, -1);
    for (int i = rank; i < 3 * NFLAG; i += GS) vm.flags[i] = flag_init(i % NFLAG);
    {
      uint32_t* const nullslot = rD2 + (TWO_P ? P.r2 * capw : 0);      /* read in place of null sources */
      for (int i = rank; i < capw; i += GS) nullslot[i] = NULL2;
    }
    gsync<NW>();
    tb = blo;
    for (int b = blo + wofs; b <= bhi; b += CW) {
      const int pos = ((b - blo) << 5) + lane;
      const int kblock = 64 * b - BIAS, k0 = kblock + 2 * lane;
      const int u0 = imax(imin(tlen, plen + k0), UB_MIN), u1 = imax(imin(tlen, plen + k0 + 1), UB_MIN);
      const int s0 = (k0 >= lo0 && k0 <= hi0) ? imax(k0, 0) : NULL16;
      const int s1 = (k0 + 1 >= lo0 && k0 + 1 <= hi0) ? imax(k0 + 1, 0) : NULL16;
      finish_m<VEC_WIN(NW)>(cx, rM, vm.flags, false, lv::pack2(s0, s1), pos, k0, u0, u1, kblock, lane);
    }
    for (int c = 0; c < 5; ++c) { clo[c] = 1; chi[c] = -1; }
    clo[CM] = lo0; chi[CM] = hi0;
    if (rank == 0) meta[0] = make_int4(lo0, hi0, blo, bhi);
    gsync<NW>();
  }

#ifdef WFA_VEC_TIMING   /* debugging build: where the cycles of a step go (per warp; see wfagpu_api.cpp trace) */
  long long tm_prev = clock64(), tm_acc[6] = {0, 0, 0, 0, 0, 0};
#define WFA_TM(i) { const long long t_ = clock64(); tm_acc[i] += t_ - tm_prev; tm_prev = t_; }
#else
#define WFA_TM(i)
#endif
  for (;;) {
    /* ---- after-extend step of score s (extend.c:90-125 / :263-297) ---- */
    WFA_TM(5)
    int* const F = vm.flags + fb * NFLAG;
    if (cur_exists) {
      const int term_k = F[F_TERM];
#ifdef WFA_VEC_TIMING
      if (F[F_EDGE] && !exact && lane == 0 && P.dbg) atomicAdd(P.dbg + 8, 1ull);
#endif
      if (F[F_EDGE]) { exact = true; sticky = true; }
      if (term_k != KNONE) {
        end_k = term_k;
        const int wi = (term_k + BIAS) >> 1;
        int pos = tp + (((wi >> 5) - tb) << 5); if (pos >= capw) pos -= capw;
        const uint32_t w = rM[cm * capw + pos + (wi & 31)];
        end_off = ((term_k + BIAS) & 1) ? lv::sx_hi(w) : lv::sx_lo(w);
        status = 1; end_score = s * P.g;
        cells += imax(0, chi[CM] - clo[CM] + 1);
        break;
      }
      if (HEUR != 0 && clo[CM] <= chi[CM]) {
        /* wavefront_heuristic_cufoff, heuristic.c:509-567 */
        --steps_wait;
        const int lo_base = clo[CM], hi_base = chi[CM];
        const uint32_t* const mslot = rM + cm * capw;
        if (steps_wait <= 0) {
          const int hb_lo = (lo_base + BIAS) >> 6, hb_hi = (hi_base + BIAS) >> 6;
          if (HEUR == 1) {
            /* wavefront_heuristic_wfadaptive, heuristic.c:257-293 */
            if (hi_base - lo_base + 1 >= P.min_wf_len) {
              int dm = INT_MAX;
              for (int b = hb_lo + wofs; b <= hb_hi; b += CW) {
                int pos = tp + ((b - tb) << 5); if (pos >= capw) pos -= capw;
                const uint32_t w = mslot[pos + lane];
                const int k0 = 64 * b - BIAS + 2 * lane;
                const int f0 = lv::sx_lo(w), f1 = lv::sx_hi(w);
                if (k0 >= lo_base && k0 <= hi_base) dm = imin(dm, f0 >= 0 ? imax(plen - (f0 - k0), tlen - f0) : (1 << 30));
                if (k0 + 1 >= lo_base && k0 + 1 <= hi_base) dm = imin(dm, f1 >= 0 ? imax(plen - (f1 - k0 - 1), tlen - f1) : (1 << 30));
              }
              dm = __reduce_min_sync(0xffffffffu, dm);
              if (lane == 0 && dm != INT_MAX) atomicMin(&F[F_MINA], dm);
              gsync<NW>();
              const int min_d = imin(imax(plen, tlen), F[F_MINA]);
              int kf = INT_MAX, kl = INT_MIN;
              for (int b = hb_lo + wofs; b <= hb_hi; b += CW) {
                int pos = tp + ((b - tb) << 5); if (pos >= capw) pos -= capw;
                const uint32_t w = mslot[pos + lane];
                const int k0 = 64 * b - BIAS + 2 * lane;
                const int f0 = lv::sx_lo(w), f1 = lv::sx_hi(w);
                if (k0 >= lo_base && k0 <= hi_base) {
                  const int d = f0 >= 0 ? imax(plen - (f0 - k0), tlen - f0) : (1 << 30);
                  if (d - min_d <= P.max_dist_thr) { kf = imin(kf, k0); kl = imax(kl, k0); }
                }
                if (k0 + 1 >= lo_base && k0 + 1 <= hi_base) {
                  const int d = f1 >= 0 ? imax(plen - (f1 - k0 - 1), tlen - f1) : (1 << 30);
                  if (d - min_d <= P.max_dist_thr) { kf = imin(kf, k0 + 1); kl = imax(kl, k0 + 1); }
                }
              }
              kf = __reduce_min_sync(0xffffffffu, kf); kl = __reduce_max_sync(0xffffffffu, kl);
              if (lane == 0) { if (kf != INT_MAX) atomicMin(&F[F_MINB], kf); if (kl != INT_MIN) atomicMax(&F[F_MAXA], kl); }
              gsync<NW>();
              kf = F[F_MINB]; kl = F[F_MAXA];
              const int top_limit = imin(ak, hi_base);
              const int nlo = (kf < top_limit) ? kf : imax(lo_base, top_limit);
              const int bottom = imax(ak, nlo);
              const int nhi = (kl > bottom) ? kl : imin(hi_base, bottom);
              clo[CM] = nlo; chi[CM] = nhi;
              steps_wait = P.steps_between;
            }
          } else {
            /* wavefront_heuristic_xdrop, heuristic.c:329-383 (+ sw scores :297-328) */
            const int swg = (P.match != 0) ? -P.match : -1;
            const int so = s * P.g;
            int cmax = INT_MIN, kf = INT_MAX, kl = INT_MIN;
            for (int b = hb_lo + wofs; b <= hb_hi; b += CW) {
              int pos = tp + ((b - tb) << 5); if (pos >= capw) pos -= capw;
              const uint32_t w = mslot[pos + lane];
              const int k0 = 64 * b - BIAS + 2 * lane;
              const int f0 = lv::sx_lo(w), f1 = lv::sx_hi(w);
              if (k0 >= lo_base && k0 <= hi_base && f0 >= 0) {
                const int sw = (swg * (2 * f0 - k0) - so) / 2;
                cmax = imax(cmax, sw);
                if (sw_init && max_sw - sw < P.xdrop) { kf = imin(kf, k0); kl = imax(kl, k0); }
              }
              if (k0 + 1 >= lo_base && k0 + 1 <= hi_base && f1 >= 0) {
                const int sw = (swg * (2 * f1 - k0 - 1) - so) / 2;
                cmax = imax(cmax, sw);
                if (sw_init && max_sw - sw < P.xdrop) { kf = imin(kf, k0 + 1); kl = imax(kl, k0 + 1); }
              }
            }
            cmax = __reduce_max_sync(0xffffffffu, cmax);
            kf = __reduce_min_sync(0xffffffffu, kf); kl = __reduce_max_sync(0xffffffffu, kl);
            if (lane == 0) {
              if (cmax != INT_MIN) atomicMax(&F[F_MAXB], cmax);
              if (kf != INT_MAX) atomicMin(&F[F_MINB], kf);
              if (kl != INT_MIN) atomicMax(&F[F_MAXA], kl);
            }
            gsync<NW>();
            cmax = F[F_MAXB]; kf = F[F_MINB]; kl = F[F_MAXA];
            if (sw_init) {
              const int nlo = (kf == INT_MAX) ? hi_base + 1 : kf;
              const int nhi = (kl >= nlo) ? kl : imin(nlo - 1, hi_base);
              clo[CM] = nlo; chi[CM] = nhi;
              if (cmax > max_sw) max_sw = cmax;
            } else { max_sw = cmax; sw_init = true; }
            steps_wait = P.steps_between;
          }
        }
        if (clo[CM] != lo_base || chi[CM] != hi_base) {
          /* wf_heuristic_equate, heuristic.c:161-172; dropped cells are nulled in place */
          bool ex[5];
          ex[CM] = true;
          for (int c = 1; c < 5; ++c) {
            ex[c] = clo[c] <= chi[c];
            if (!ex[c]) continue;
            if (clo[CM] > clo[c]) clo[c] = clo[CM];
            if (chi[CM] < chi[c]) chi[c] = chi[CM];
          }
          uint32_t* const slots[5] = {rM + cm * capw, rI1 + c1 * capw, rD1 + c1 * capw, rI2 + c2 * capw, rD2 + c2 * capw};
          if (!sticky) {
            /* the derived ranges of later scores assume valid end cells: otherwise scan while this
             * wavefront can still be a source (max look-back = ring depth) */
            for (int c = 0; c < NC; ++c) {
              if (clo[c] > chi[c]) continue;
              for (int e = 0; e < 2; ++e) {
                const int k = e ? chi[c] : clo[c];
                const int wi = (k + BIAS) >> 1;
                int pos = tp + (((wi >> 5) - tb) << 5); if (pos >= capw) pos -= capw;
                const uint32_t w = slots[c][pos + (wi & 31)];
                if ((((k + BIAS) & 1) ? lv::sx_hi(w) : lv::sx_lo(w)) < 0) {
#ifdef WFA_VEC_TIMING
                  if (!exact && lane == 0 && P.dbg) atomicAdd(P.dbg + 9 + c, 1ull);
#endif
                  exact_until = s + P.rm;
                }
              }
            }
          }
          for (int b = blo + wofs; b <= bhi; b += CW) {
            int pos = tp + ((b - tb) << 5); if (pos >= capw) pos -= capw;
            pos += lane;
            const int k0 = 64 * b - BIAS + 2 * lane;
            for (int c = 0; c < NC; ++c)
              if (ex[c]) clip_word(slots[c], pos, k0, clo[c], chi[c]);
          }
          if (rank == 0) {
            int4* const mrow = meta + (s & mmask) * 3;
            const bool n0 = clo[CM] > chi[CM];
            mrow[0] = make_int4(n0 ? 1 : clo[CM], n0 ? -1 : chi[CM], blo - cur_pad, bhi + cur_pad);
            const bool n1 = clo[CI1] > chi[CI1], n2 = clo[CD1] > chi[CD1], n3 = clo[CI2] > chi[CI2], n4 = clo[CD2] > chi[CD2];
            mrow[1] = make_int4(n1 ? 1 : clo[CI1], n1 ? -1 : chi[CI1], n2 ? 1 : clo[CD1], n2 ? -1 : chi[CD1]);
            mrow[2] = make_int4(n3 ? 1 : clo[CI2], n3 ? -1 : chi[CI2], n4 ? 1 : clo[CD2], n4 ? -1 : chi[CD2]);
          }
          gsync<NW>();
        }
      }
#ifdef WFA_VEC_DEBUG_EXACT
      cells += exact ? 1 : 0;          /* debugging build: count the scores spent in the scanned-range variant */
#else
      cells += imax(0, chi[CM] - clo[CM] + 1);
#endif
    }

    /* ---- compute score s+1 (compute_affine.c:229-260 / compute_affine2p.c:334-368) ---- */
    ++s;
    if (++cm == P.rm) cm = 0;
    if (++c1 == P.r1) c1 = 0;
    if (TWO_P) { if (++c2 == P.r2) c2 = 0; }
    if (++fb == 3) fb = 0;
    int* const Fn = vm.flags + fb * NFLAG;                 /* cells of this step */
    {
      /* reset the set the NEXT step will use: its last readers finished before the previous barrier */
      int* const Fr = vm.flags + (fb == 2 ? 0 : fb + 1) * NFLAG;
      if (rank < NFLAG) Fr[rank] = flag_init(rank);
    }
    if ((s & RENORM_MASK) == 0) {
      /* re-base drifting I/D nulls (every negative offset is a null) */
      const int nw = (2 * P.r1 + (TWO_P ? 2 * P.r2 : 0)) * capw;
      for (int i = rank; i < nw; i += GS) {
        const uint32_t x = rI1[i];
        const uint32_t neg = lv::signmask2(x);
        rI1[i] = (x & ~neg) | (NULL2 & neg);
      }
      gsync<NW>();
    }
    if (HEUR != 0) exact = sticky || s <= exact_until;
    {
      /* plan of this score: precomputed by the planner warp during the previous step, or by everybody now */
      PlanOut pl;
      if (NW > 1 && have_plan) {
        pl = vm.plan[s & 1];
        if (exact) pl.exact = 1;      /* (multi-warp groups never run cut-offs pipelined: exact here means sticky) */
      } else {
        if (NW == 1) {
          plan_step<TWO_P, FULL, false>(P, vm, plen, tlen, s, cm, c1, c2, tp, tb, cell_off, exact, is_writer, pl);
        } else {
          /* start-up and scanned-range steps: every warp runs the lane-parallel planner (same values,
           * same destination), then the plan is read back like a precomputed one -- the kernel carries
           * ONE planner, which keeps the step loop small */
          plan_step_simd<TWO_P, FULL, (NW > 1) && WFA_VEC_PAD>(P, vm, plen, tlen, s, cm, c1, c2, tp, tb, cell_off, exact, &vm.plan[s & 1]);
          gsync<NW>();
          pl = vm.plan[s & 1];
        }
      }
      if (pl.kind == 2) return PAIR_OVERFLOW;
      if (pl.exact != 0 && !exact) sticky = true;        /* the planned range left the matrix */
      exact = pl.exact != 0;
      tp = pl.tp; tb = pl.nblo; cell_off = pl.cell_off;
      blo = pl.nblo; bhi = pl.nbhi; cur_pad = pl.pad;
      for (int c = 0; c < 5; ++c) { clo[c] = pl.lo[c]; chi[c] = pl.hi[c]; }
      WFA_TM(1)
      /* the planner may run one score ahead while nothing can invalidate derived ranges */
      const bool pipelined = NW > 1 && HEUR == 0 && !exact;
      if (pl.kind == 1) {
        cur_exists = false;
      } else {
        cur_exists = true;
        s_exist = s * P.g;
        for (int b = blo - pl.pad + wofs; b <= bhi + pl.pad; b += CW) {
          int posb = tp + ((b - blo) << 5);
          if (posb >= capw) posb -= capw;
          if (posb < 0) posb += capw;
          if (b < blo || b > bhi) {
            /* padding block: nulls in every component this score owns */
            vm.ring[pl.oM + posb + lane] = NULL2;
            if (pl.oI1 >= 0) vm.ring[pl.oI1 + posb + lane] = NULL2;
            if (pl.oD1 >= 0) vm.ring[pl.oD1 + posb + lane] = NULL2;
            if (TWO_P) {
              if (pl.oI2 >= 0) vm.ring[pl.oI2 + posb + lane] = NULL2;
              if (pl.oD2 >= 0) vm.ring[pl.oD2 + posb + lane] = NULL2;
            }
          } else if (b >= pl.fl && b <= pl.fh) vec_block<TWO_P, FULL, false, VEC_WIN(NW)>(cx, pl, Fn, exact, b, posb, lane);
          else vec_block<TWO_P, FULL, true, VEC_WIN(NW)>(cx, pl, Fn, exact, b, posb, lane);
        }
      }
      WFA_TM(2)
      if (NW > 1) {
        if (pipelined && warp == NW - 1) {
          const int ncm = (cm + 1 == P.rm) ? 0 : cm + 1, nc1 = (c1 + 1 == P.r1) ? 0 : c1 + 1;
          const int nc2 = TWO_P ? ((c2 + 1 == P.r2) ? 0 : c2 + 1) : 0;
          __syncwarp();                                    /* lane 0's metadata of score s is visible to the warp */
          plan_step_simd<TWO_P, FULL, (NW > 1) && WFA_VEC_PAD>(P, vm, plen, tlen, s + 1, ncm, nc1, nc2, tp, tb, cell_off, false, &vm.plan[(s + 1) & 1]);
        }
        have_plan = pipelined;
      }
      WFA_TM(3)
      if (exact && pl.kind == 0) {
        gsync<NW>();
        /* trim_ends, compute.c:571-605: [first in-matrix cell, last in-matrix cell] per component */
        const bool has[5] = {true, pl.oI1 >= 0, pl.oD1 >= 0, TWO_P && pl.oI2 >= 0, TWO_P && pl.oD2 >= 0};
        for (int c = 0; c < 5; ++c) {
          const int l = Fn[F_LO + c], h = Fn[F_HI + c];
          if (has[c] && l != INT_MAX) { clo[c] = l; chi[c] = h; } else { clo[c] = 1; chi[c] = -1; }
        }
        if (Fn[F_POISON]) {
          /* offsets beyond the matrix that the trimmed range no longer covers read as NULL */
          for (int b = blo + wofs; b <= bhi; b += CW) {
            int pos = tp + ((b - blo) << 5); if (pos >= capw) pos -= capw;
            pos += lane;
            const int k0 = 64 * b - BIAS + 2 * lane;
            if (has[CI1]) clip_word(vm.ring + pl.oI1, pos, k0, clo[CI1], chi[CI1]);
            if (has[CD1]) clip_word(vm.ring + pl.oD1, pos, k0, clo[CD1], chi[CD1]);
            if (TWO_P) {
              if (has[CI2]) clip_word(vm.ring + pl.oI2, pos, k0, clo[CI2], chi[CI2]);
              if (has[CD2]) clip_word(vm.ring + pl.oD2, pos, k0, clo[CD2], chi[CD2]);
            }
          }
        }
        if (is_writer) {
          int4* const mrow = meta + (s & mmask) * 3;
          const bool n0 = clo[CM] > chi[CM];
          mrow[0] = make_int4(n0 ? 1 : clo[CM], n0 ? -1 : chi[CM], blo - pl.pad, bhi + pl.pad);
          mrow[1] = make_int4(clo[CI1], chi[CI1], clo[CD1], chi[CD1]);
          mrow[2] = make_int4(clo[CI2], chi[CI2], clo[CD2], chi[CD2]);
        }
      }
      gsync<NW>();
      WFA_TM(4)
#ifdef WFA_VEC_TIMING
      ++tm_acc[0];
#endif
    }
    /* unreachable (extend.c:99-106) and the step limit (unialign.c:98-109), in ORIGINAL score
     * units: between two multiples of g every score is a null step of the reference */
    {
      const int so = s * P.g;
      if (!cur_exists) {
        const int su = s_exist + P.max_scope + 1;
        if (su <= so && su < P.max_steps) { status = 2; end_score = su; break; }
      }
      if (so >= P.max_steps) {
        status = 3;
        if (so == P.max_steps) cells += imax(0, chi[CM] - clo[CM] + 1);
        break;
      }
    }
  }

#ifdef WFA_VEC_TIMING
  if (lane == 0 && P.dbg) for (int i = 0; i < 6; ++i) atomicAdd(P.dbg + i, (unsigned long long)tm_acc[i]);
#endif
  /* ---- wavefront_unialign_terminate, unialign.c:147-237 ---- */
  res.cells = cells;
  res.nruns = 0;
  res.locs[0] = res.locs[1] = res.locs[2] = res.locs[3] = 0;
  if (status == 3) {
    res.score = -P.max_steps; res.status = ST_MAX_STEPS;
  } else if (!FULL) {
    if (status == 1) { res.score = classic_score(P.match, plen, tlen, end_score, P.pos_score); res.status = ST_COMPLETED; }
    else {
      const int32_t end_v = (int32_t)((uint32_t)OFFNULL - (uint32_t)INT_MAX);
      res.score = classic_score(P.match, end_v, OFFNULL, end_score, P.pos_score); res.status = ST_PARTIAL;
    }
  } else {
    if (status == 1) {
      if (rank == 0) {
        FwdEmitter em; em.init(vm.runs_stage, P.runcap);
        const int n = backtrace_vcodes(P, vm.h_code, vm.hmeta, s, end_k, plen, tlen, vm.bpw, vm.btw, vm.ops, vm.opcap, em);
        res.nruns = n;
        if (n >= 0) locations_from_runs(vm.runs_stage, imin(n, P.runcap), plen, tlen, res.locs);
      }
      res.score = classic_score(P.match, end_off - end_k, end_off, end_score, P.pos_score);
      res.status = ST_COMPLETED;
    } else {
      res.score = INT32_MIN; res.status = ST_PARTIAL;
    }
  }
  return PAIR_DONE;
}

}  // namespace vec
}  // namespace wfagpu

#endif  /* __CUDACC__ */
