/*
 * wfa_reg.cuh -- register-resident wavefront tier: one warp aligns one read pair with the
 * gap-affine wavefronts held entirely in registers.
 *
 * Replaces, for short reads / low scores, the same reference path as wfa_core.cuh
 * (W/ = pywfa/WFA2_lib/): wavefront_compute_affine_idm (W/wavefront/wavefront_compute_affine.c:44-86)
 * fused with wavefront_extend_matches_packed_end2end / _endsfree
 * (W/wavefront/wavefront_extend_kernels.c:96-163), trim_ends
 * (W/wavefront/wavefront_compute.c:571-605), the end-to-end / ends-free termination tests
 * (W/wavefront/wavefront_termination.c:37-162), the step limit
 * (W/wavefront/wavefront_unialign.c:98-109) and the backtrace
 * (W/wavefront/wavefront_backtrace.c:320-529).
 *
 * Layout.  A window of 64*P consecutive diagonals around the score-0 seeds is mapped onto the
 * warp: packed register p of lane j holds, as two signed 16-bit offsets, the diagonals
 * kbase + 64p + j (low half) and kbase + 64p + 32 + j (high half).  M wavefronts of the last
 * max(x, o+e) scores (in units of the penalties' gcd) and the current I / D wavefronts stay in
 * registers (the ring is rotated by register moves, so every slot is a fixed register and the
 * step body exists once and stays resident in the instruction cache).  The recurrence runs on the DPX packed-halfword pipe (VIMNMX.S16x2, VIMNMX3.S16x2,
 * VIADD.16x2): 64 cells per instruction.  Neighbouring diagonals (k-1, k+1) come from one lane
 * rotation (SHFL) per direction and register plus a PRMT that mends the block seams.
 *
 * Without heuristics the wavefront of score s covers exactly the diagonals within
 * reach(s) = s - (o+e) + 1 of the seeds (clipped to the DP matrix), so the active blocks are
 * known without any reduction; blocks outside are skipped by warp-uniform branches.
 *
 * Nulls and out-of-matrix cells.  Offsets are int16: everything negative is null
 * (W/wavefront/wavefront_offset.h:44; the reference's nulls drift upwards by one per step, so
 * do these).  ub[k] = min(tlen, plen + k) is the largest in-matrix offset of diagonal k; an M
 * offset above it is nulled (compute_affine.c:80-84).  I/D offsets above it stay as they are
 * ("poison": they win every max on their diagonal and null the M cell, exactly as in the
 * reference) unless trim_ends removes them; that exact trimming runs only in the rare steps
 * in which a packed compare finds a new I/D offset above ub.
 *
 * scope=full.  One origin code (4 bits) per cell is written to a per-warp arena (row = score, the two
 * cells of a lane's packed register share a byte; shared memory for the 128-diagonal window, HBM / L2
 * otherwise): bits 0-1 winner of M (ties resolve M > D > I as W/wavefront/wavefront_backtrace.c:49-59),
 * bit 2 "I[s][k+1] is opened" and bit 3 "D[s][k-1] is opened" (ext >= open extends), stored at the
 * source diagonal.  The backtrace walks these
 * bytes from the end cell to score 0 collecting the edit operations, then replays them forwards
 * re-extending the matches from the sequences, which yields the run-length encoded CIGAR in
 * order without ever storing offsets.
 *
 * The file contains no per-lane control flow (see lanevec.cuh); tests/emu/ runs it on the CPU.
 */
#pragma once
#include <stdint.h>
#include <limits.h>

#include <utility>

#include "lanevec.cuh"
#include "wfa_core.cuh"

namespace wfagpu {

#ifndef WFA_REG_LEAN_EXT
#define WFA_REG_LEAN_EXT 1     /* extension of a block: first compare with the predicate taken from the offset, one vote for the common visit */
#endif
#ifndef WFA_REG_NARROW
#define WFA_REG_NARROW 0       /* 1: recurrence and ring rotation on the middle register(s) only while the wavefront stays inside them.
                                  Measured r02 with the lean extension: cfg2 68.0 M pairs/s with it, 72.0 without (the two push variants
                                  cost more register moves in the step's tail than the narrow recurrence saves).
                                  2: narrow recurrence, but one rotation of the whole ring: measured last in r02, cfg2 73.8 against 72.1 M
                                  pairs/s (+2.4 %); not the default because the round's GPU budget did not cover a second pass of the GPU
                                  suite and a fresh ncu capture on it (CPU suite green with it) */
#endif
constexpr uint32_t REG_NULL2 = 0xC000C000u;     /* two int16 nulls (-16384) */
constexpr int REG_NULL16 = -16384;
constexpr int REG_UB_MIN = -8192;              /* floor of ub[k] for diagonals left of the matrix */
constexpr uint32_t REG_ONE2 = 0x00010001u;

struct RegParams {
  int match, g, max_steps, pos_score;
  int endsfree, pbf, pef, tbf, tef;
  int kbase, c_lo, c_hi;   /* reg_window() of the launch */
  int hrows;          /* scope=full: rows of the origin arena (scores 0..hrows-1) */
  int opcap;          /* edit-operation stack bytes */
  int runcap;         /* CIGAR run staging words */
};

/*
 * Origin arena layout.  One byte per lane and packed register: low nibble = code of the register's
 * low half (window column 64p + lane), high nibble = code of its high half (column 64p + 32 + lane);
 * row = score, 32 * P bytes per row.  Code bits: 0 "deletion beats mismatch", 1 "insertion beats
 * both", 2 "I of this diagonal is opened (not extended)", 3 "D of this diagonal is opened".
 */
template <bool SH>
WFA_DEV int origin_code(const lv::histref& h, int row_bytes, int s, int d) {
  const int b = lv::hist_load<SH>(h, s * row_bytes + ((d >> 6) << 5) + (d & 31));
  return (d & 32) ? (b >> 4) : (b & 15);
}

/*
 * Backtrace over origin codes (one thread): backward walk, then forward replay.
 * Returns the number of runs (may exceed em.cap, then the CIGAR did not fit) or -1 if the
 * operation stack overflowed.  `em` may write over the arena: the walk is over by then.
 */
template <bool SH>
WFA_DEV int backtrace_origin(const lv::histref& hist, int row_bytes, int kbase, int dx, int doe, int de,
                             int s_end, int k_end, int plen, int tlen, const uint32_t* pw, const uint32_t* tw,
                             uint8_t* ops, int opcap, FwdEmitter& em, int wild = -1) {
  int s = s_end, d = k_end - kbase, nops = 0;
  int mt = CM;
  while (s > 0) {
    int enter = mt;
    if (mt == CM) {
      const int c = origin_code<SH>(hist, row_bytes, s, d);
      if (!(c & 3)) { if (nops < opcap) ops[nops] = EOP_X; ++nops; s -= dx; continue; }
      enter = (c & 2) ? CI1 : CD1;
    }
    if (enter == CI1) {
      const int ext = !((origin_code<SH>(hist, row_bytes, s, d - 1) >> 2) & 1);
      if (nops < opcap) ops[nops] = ext ? EOP_I_EXT : EOP_I_OPEN;
      ++nops; --d;
      if (ext) { s -= de; mt = CI1; } else { s -= doe; mt = CM; }
    } else {
      const int ext = !((origin_code<SH>(hist, row_bytes, s, d + 1) >> 3) & 1);
      if (nops < opcap) ops[nops] = ext ? EOP_D_EXT : EOP_D_OPEN;
      ++nops; ++d;
      if (ext) { s -= de; mt = CD1; } else { s -= doe; mt = CM; }
    }
  }
  if (nops > opcap) return -1;
  replay_ops(ops, nops, kbase + d, plen, tlen, pw, tw, em, wild);
  return em.n;
}

/* ------------------------------------------------------------------------------------ */
/*
 * Sequence windows.  For the extension every base position i of a sequence gets one 32-bit
 * word in shared memory holding the 16 bases i .. i+15, first base in the top bits, so that
 * comparing 16 bases of pattern and text is LDS, LDS, XOR, CLZ with no funnel shift and no
 * word/bit index arithmetic.  Built once per pair from the 2-bit packed words (readable one
 * word past the end); bases beyond the sequence end are garbage and are clamped away by the
 * caller.  `len + 1` windows are written.
 */
/* CB = bits per base of `words`: 2 (ACGT codes, 16 bases per window) or 4 (the symbol codes of byte mode,
 * nibble_words below: 8 bases per window; the bit reversal also reverses every code, which equality does not mind) */
template <int CB = 2>
WFA_DEV void build_windows(const uint32_t* words, int len, uint32_t* win) {
  using namespace lv;
  constexpr int LG = CB == 2 ? 4 : 3;              /* log2(bases per word) */
  const vi lane = lane_id();
  for (int i0 = 0; i0 <= len; i0 += 32) {
    const vi i = lane + i0;
    const vb in = i <= len;
    const vi j = i >> LG;
    const vu w = vfunnel_r(gather_u32(words, j, in), gather_u32(words, j + 1, in), (i & ((1 << LG) - 1)) << (5 - LG));
    scatter_u32(win, i, vbrev(w), in);
  }
}

/*
 * Byte mode (non-ACGT input / the wildcard; the reference compares bytes:
 * W/wavefront/wavefront_extend_kernels.c:167-203 through wildcard_match_fun, pywfa/align.pyx:297-304):
 * the pair's upper-cased bytes (4 per word, wfa_pack.cu: pack_bytes_kernel) -> 4-bit symbol codes, 8 per word
 * (lv::nib_pack8), (len >> 3) + 2 words written (the windows read one word past the last base).
 * Returns false if the sequence holds a byte outside the symbol set (the pair then takes the scalar tiers).
 */
WFA_DEV bool nibble_words(const uint32_t* bytes, int len, uint32_t* out, int wild) {
  using namespace lv;
  const vi lane = lane_id();
  const int nout = (len >> 3) + 2, nin = (len + 3) >> 2;
  vb bad = vfalse();
  for (int j0 = 0; j0 < nout; j0 += 32) {
    const vi j = lane + j0;
    const vb in = j < nout;
    const vi b0 = j + j;
    const vu w0 = gather_u32(bytes, b0, in & (b0 < nin)), w1 = gather_u32(bytes, b0 + 1, in & (b0 + 1 < nin));
    const vi left = splati(len) - (j << 3);         /* bases of the sequence from this word on (<= 0: padding) */
    scatter_u32(out, j, nib8(w0, w1, left, (uint32_t)wild, bad), in);
  }
  return !any(bad);
}

template <int P, int DX, int DOE, bool FULL, bool HS_ = reg_hist_in_smem(P, FULL), int CB = 2>
struct RegAligner {
  static constexpr int RM = DX > DOE ? DX : DOE;   /* M ring: M[r] = wavefront of score s-1-r when score s is computed */
  static constexpr int DE = 1;
  static constexpr int NB = 2 * P;                 /* 32-diagonal blocks */
  static constexpr int WIN = 64 * P;
  /* warp-uniform flags: bit r < RM "M[r] holds a valid offset"; then I, D; then the sticky "wavefront
   * extents must be scanned, not derived from reach(s)" */
  static constexpr uint32_t F_I = 1u << 8, F_D = 1u << 9, F_EXACT = 1u << 10;
  static constexpr uint32_t F_MRING = (1u << RM) - 1u;
  static constexpr uint32_t F_SOURCES = (1u << (DX - 1)) | (1u << (DOE - 1)) | F_I | F_D;

  /* wavefront registers */
  lv::vu M[RM][P], I[P], D[P];
  lv::vu ub2[P];                                   /* packed ub[k] per register */
  /* per-lane constants of the pair, pinned in registers (lv::keep) so that the step loop never re-derives them */
  lv::vi lane, lprev, lnext;                       /* lane, and the lanes holding diagonals k - 1 / k + 1 */
  lv::vu selL, selR;                               /* PRMT selectors mending the block seams */
  lv::lanead pa, ta;                               /* window of pattern position (offset - diagonal of block 0) / text position offset */
  /* warp-uniform state */
  uint32_t flags;
  bool endsfree;
  int pef, tef;
  static constexpr bool HS = HS_;                  /* origin arena in shared memory */
  lv::histref hist;
  int hrows;
  int plen, tlen, kbase;
  int c_lo, c_hi;                                  /* window columns of the score-0 seeds [lo0, hi0] */
  int m_lo, m_hi;                                  /* window columns of the outermost diagonals of the DP matrix (-plen, tlen) */
  int s, s_limit, s_limit_exact;                   /* step limit in units of g (unialign.c:98-109) */
  int s_event;                                     /* first score at which the derived range needs a second look: it leaves the
                                                      DP matrix (clip + F_EXACT), the window or the origin arena (overflow) */
  int wprev;                                       /* cells of the newest M wavefront (0: it holds no valid offset) */
  int cells;
  int term_d, term_off;                            /* window column and offset of the end cell (term_d < 0: none) */
  int dak;                                         /* window column of the end-to-end target tlen - plen */
  int status;                                      /* 0 running, 1 end reached, 3 max steps, 4 overflow */
  int cur_lo, cur_hi;                              /* window range [first, last] of the newest M wavefront */

  /* pattern XOR text from the cells' positions on: the first set bit marks the first differing base.  Byte mode
   * (CB = 4): 8 symbol codes per word, and a position where either side is the wildcard (code 0) never differs */
  static constexpr int LGC = CB == 2 ? 1 : 2;      /* log2(bits per base) */
  template <int B>
  WFA_DEV lv::vu diff_at(const lv::vi& off, const lv::vb& p) {
    using namespace lv;
    if constexpr (CB == 2) {
      return load_win_at<-32 * B>(pa, off, p, 0u) ^ load_win_at<0>(ta, off, p, 0x80000000u);
    } else {
      /* idle lanes see the codes 8 and 9 (bit-reversed in the window): "first base differs" */
      const vu a = load_win_at<-32 * B>(pa, off, p, 0x10000000u), b = load_win_at<0>(ta, off, p, 0x90000000u);
      return (a ^ b) & nib_both(a, b);
    }
  }

  /* ---- extension of one block of 32 diagonals (extend_kernels.c:64-110), in place in its half of
   * the packed register `mn`, followed by the edge / termination bookkeeping (termination.c:37-162) ---- */
  template <int B>
  WFA_DEV void extend_block(lv::vu& mn) {
    using namespace lv;
    constexpr int p = B >> 1;
    constexpr bool HI = (B & 1) != 0;
    const vi off0 = HI ? sx_hi(mn) : sx_lo(mn);
    const vi ubk = HI ? sx_hi(ub2[p]) : sx_lo(ub2[p]);
    const vb valid = off0 >= 0;
    /* 16 bases per XOR; a null cell loads nothing and sees "first base differs", so it never moves;
     * min(offset + matches, ub) clamps the run at the end of the diagonal (VIADDMNMX) */
#if WFA_REG_LEAN_EXT
    vi off;
    if constexpr (CB == 2) {
      /* the common visit -- every run ends within 16 bases and no cell touches the edge of the matrix -- costs one
       * vote: "a cell needs a second look" = 16 bases matched and there is room left, or the offset reached ub */
      const vu x = diff_win_nonneg<-32 * B, 0>(pa, ta, off0);
      off = vaddmin(off0, vclz(x) >> 1, ubk);
      vb more = (x == 0u) & (off < ubk);
      if (!any(more | (off >= ubk))) { mn = HI ? put_hi(mn, off) : put_lo(mn, off); return; }
      while (any(more)) {
        const vu y = diff_at<B>(off, more);
        off = vaddmin(off, vclz(y) >> 1, ubk);
        more = more & (y == 0u) & (off < ubk);
      }
    } else {
      const vu x = diff_at<B>(off0, valid);
      off = vaddmin(off0, vclz(x) >> LGC, ubk);
      vb more = (x == 0u) & (off < ubk);
      while (any(more)) {
        const vu y = diff_at<B>(off, more);
        off = vaddmin(off, vclz(y) >> LGC, ubk);
        more = more & (y == 0u) & (off < ubk);
      }
    }
#else
    const vu x = diff_at<B>(off0, valid);
    vi off = vaddmin(off0, vclz(x) >> LGC, ubk);
    vb more = (x == 0u) & (off < ubk);
    while (any(more)) {
      const vu y = diff_at<B>(off, more);
      off = vaddmin(off, vclz(y) >> LGC, ubk);
      more = more & (y == 0u) & (off < ubk);
    }
#endif
    mn = HI ? put_hi(mn, off) : put_lo(mn, off);
    /* a cell on the edge of the matrix (null offsets are far below every ub) */
    const vb edge = off >= ubk;
    if (!any(edge)) return;
    const uint32_t eb = ballot(edge);
    flags |= F_EXACT;
    if (term_d >= 0) return;
    if (endsfree) {                                   /* termination.c:115-162 */
      const vi vv = off - (lane + (kbase + 32 * B));
      const vb t = valid & (((off >= tlen) & (vv >= plen - pef)) | ((vv >= plen) & (off >= tlen - tef)));
      const uint32_t tb = ballot(t);
      if (tb) { term_d = 32 * B + first_set(tb); term_off = lane_value(off, first_set(tb)); }
    } else {                                          /* termination.c:37-61 */
      if ((dak >> 5) == B && ((eb >> (dak & 31)) & 1u)) { term_d = dak; term_off = lane_value(off, dak & 31); }
    }
  }

  /* blocks 2p and 2p + 1 (the halves of packed register p), then the registers after it */
  template <int p>
  WFA_DEV void extend_blocks(lv::vu (&Mn)[P], uint32_t bmask) {
    if constexpr (p < P) {
      if ((bmask >> (2 * p)) & 3u) {
        if ((bmask >> (2 * p)) & 1u) extend_block<2 * p>(Mn[p]);
        if ((bmask >> (2 * p + 1)) & 1u) extend_block<2 * p + 1>(Mn[p]);
      }
      extend_blocks<p + 1>(Mn, bmask);
    }
  }

  /* first / last window diagonal holding a valid offset of the newest M wavefront */
  WFA_DEV void scan_valid_range(const lv::vu (&Mn)[P]) {
    using namespace lv;
    cur_lo = 1; cur_hi = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t b0 = ballot(sx_lo(Mn[p]) >= 0), b1 = ballot(sx_hi(Mn[p]) >= 0);
      if (b0) { if (cur_lo > cur_hi) cur_lo = 64 * p + first_set(b0); cur_hi = 64 * p + last_set(b0); }
      if (b1) { if (cur_lo > cur_hi) cur_lo = 64 * p + 32 + first_set(b1); cur_hi = 64 * p + 32 + last_set(b1); }
    }
  }

  /* exact trim_ends of an I or D wavefront (compute.c:571-605); returns whether anything is left */
  WFA_DEV bool trim_component(lv::vu (&X)[P]) {
    using namespace lv;
    int first = -1, last = -1;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const vi a = sx_lo(X[p]), b = sx_hi(X[p]);
      const uint32_t b0 = ballot((a >= 0) & (a <= sx_lo(ub2[p]))), b1 = ballot((b >= 0) & (b <= sx_hi(ub2[p])));
      if (b0) { if (first < 0) first = 64 * p + first_set(b0); last = 64 * p + last_set(b0); }
      if (b1) { if (first < 0) first = 64 * p + 32 + first_set(b1); last = 64 * p + 32 + last_set(b1); }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const vi dl = lane + 64 * p, dh = lane + (64 * p + 32);
      const vb kl = (dl >= first) & (dl <= last), kh = (dh >= first) & (dh <= last);
      const vu mask = vselu(kl, splat(0x0000ffffu), splat(0u)) | vselu(kh, splat(0xffff0000u), splat(0u));
      X[p] = bitsel(mask, X[p], splat(REG_NULL2));
    }
    return first >= 0;
  }

  /* reach(s) = s - DOE + 1 diagonals either side of the seeds (0 before the first gap can open); the first
   * score whose reach is r >= 1 is r + DOE - 1 */
  WFA_DEV int next_event() const {
    int e = INT_MAX;
    if (!(flags & F_EXACT)) {                       /* the range leaves the matrix: clip, and scan from then on */
      e = imin(e, imin(c_lo - m_lo, m_hi - c_hi) + DOE);
    }
    if (m_lo < 0) e = imin(e, c_lo + DOE);          /* ... leaves the window on the left / right: overflow */
    if (m_hi >= WIN) e = imin(e, WIN - c_hi + DOE - 1);
    if (FULL) e = imin(e, hrows);                   /* ... or the origin arena */
    return e;
  }

  /* ---- per-pair set-up; score 0 = wavefront_aligner_init_wf_m (wavefront_aligner.c:251-310) -- */
  WFA_DEV void init(const RegParams& R, lv::seqref pwin_, lv::seqref twin_, int plen_, int tlen_, const lv::histref& hist_) {
    using namespace lv;
    plen = plen_; tlen = tlen_; hist = hist_; hrows = R.hrows;
    endsfree = R.endsfree != 0; pef = R.pef; tef = R.tef;
    lane = lane_id();
    kbase = R.kbase; c_lo = R.c_lo; c_hi = R.c_hi;
    m_lo = -plen - kbase; m_hi = tlen - kbase;
    dak = tlen - plen - kbase;
    /* so >= max_steps  <=>  s >= ceil(max_steps / g);  so == max_steps  <=>  s == max_steps / g exactly */
    s_limit = R.max_steps / R.g + (R.max_steps % R.g != 0);
    s_limit_exact = (R.max_steps % R.g == 0) ? R.max_steps / R.g : -1;
    s = 0; cells = 0; term_d = -1; term_off = 0; wprev = 0;
    status = (c_lo < 0 || c_hi >= WIN) ? 4 : 0;
    flags = 0;
    cur_lo = 1; cur_hi = 0;
    s_event = next_event();
    selL = vselu(lane == 0, splat(0x5432u), splat(0x7654u));
    selR = vselu(lane == 31, splat(0x5432u), splat(0x3210u));
    /* pattern position of a cell of block 0 = offset - (kbase + lane); block B is 32 B diagonals further */
    pa = lane_addr(pwin_, splati(-kbase) - lane);
    ta = lane_addr(twin_, splati(0));
    lprev = lane + 31; lnext = lane + 1;
    keep(lane); keep(lprev); keep(lnext); keep(selL); keep(selR); keep(pa); keep(ta);
#pragma unroll
    for (int r = 0; r < RM; ++r) {
#pragma unroll
      for (int p = 0; p < P; ++p) M[r][p] = splat(REG_NULL2);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      I[p] = splat(REG_NULL2); D[p] = splat(REG_NULL2);
      const vi kl = lane + (kbase + 64 * p), kh = kl + 32;
      ub2[p] = pack2(vmax(vmin(splati(tlen), kl + plen), splati(REG_UB_MIN)), vmax(vmin(splati(tlen), kh + plen), splati(REG_UB_MIN)));
      keep(ub2[p]);
    }
  }

  /* age the M ring by one score: `Mn` becomes M[0] (registers PLO..PHI; the others hold nulls throughout) */
  template <int PLO, int PHI>
  WFA_DEV void push(const lv::vu (&Mn)[P], bool exists) {
#pragma unroll
    for (int r = RM - 1; r > 0; --r) {
#pragma unroll
      for (int p = PLO; p <= PHI; ++p) M[r][p] = M[r - 1][p];
    }
#pragma unroll
    for (int p = PLO; p <= PHI; ++p) M[0][p] = Mn[p];
    flags = (flags & ~F_MRING) | ((flags << 1) & F_MRING) | (exists ? 1u : 0u);
  }

  /* the middle register(s): columns 64 * NLO .. 64 * NHI + 63 hold the score-0 seeds of an end-to-end alignment */
  static constexpr int NLO = (P - 1) / 2, NHI = P / 2;

  /* ---- phases A + B of a score step on packed registers PLO..PHI (the others are all null) ---- */
  template <int PLO, int PHI>
  WFA_DEV void recurrence(lv::vu (&Mn)[P]) {
    using namespace lv;
    /* phase A: per source diagonal max(open, extend), rotated to the consuming lane */
    vu rl[P], rr[P];
#pragma unroll
    for (int p = PLO; p <= PHI; ++p) {
      rl[p] = from_lane(vimax2(M[DOE - 1][p], I[p]), lprev);
      rr[p] = from_lane(vimax2(M[DOE - 1][p], D[p]), lnext);
    }
    /* phase B: the recurrence (compute_affine.c:44-86), 64 diagonals per instruction */
#pragma unroll
    for (int p = PLO; p <= PHI; ++p) {
      const vu L = prmt(p > PLO ? rl[p > PLO ? p - 1 : PLO] : splat(REG_NULL2), rl[p], selL);
      const vu Rr = prmt(rr[p], p < PHI ? rr[p < PHI ? p + 1 : p] : splat(REG_NULL2), selR);
      const vu ins = vadd2(L, splat(REG_ONE2));
      const vu del = Rr;
      const vu mis = vadd2(M[DX - 1][p], splat(REG_ONE2));
      vu m;
      if (FULL) {
        /* origin code of both cells of the register, from the sign bits of packed differences (all
         * operands lie within +-16384 + drift, so no difference wraps): bit 0 "deletion beats mismatch"
         * (mismatch wins ties), bit 1 "insertion beats both" (it loses ties), bit 2 / 3 "I[s][k+1] /
         * D[s][k-1] is opened, not extended" (ext >= open extends) -- wavefront_backtrace.c:49-59 order */
        const vu m1 = vimax2(mis, del);
        m = vimax2(m1, ins);
        const vu t1 = vsub2(mis, del), t2 = vsub2(m1, ins);
        const vu t3 = vsub2(I[p], M[DOE - 1][p]), t4 = vsub2(D[p], M[DOE - 1][p]);
        const vu n = ((t1 >> 15) & 0x00010001u) | ((t2 >> 14) & 0x00020002u) | ((t3 >> 13) & 0x00040004u) | ((t4 >> 12) & 0x00080008u);
        hist_store<HS>(hist, s * (32 * P) + 32 * p, lane, as_vi((n | (n >> 12)) & 0xffu));
      } else {
        m = vimax3(mis, ins, del);
      }
      /* offsets beyond the matrix are nulled: M > ub  <=>  M + ~ub >= 0 */
      Mn[p] = bitsel(signmask2(vadd2(m, ~ub2[p])), m, splat(REG_NULL2));
      I[p] = ins; D[p] = del;
    }
  }

  /* ---- one score step: score 0 seeds the wavefront, every later score computes it -------- */
  WFA_DEV bool step(bool seeding) {
    using namespace lv;
    vu Mn[P];
    int dlo, dhi;                                     /* window range the new wavefront can occupy */
    bool narrow = false;                              /* this step touched the middle register(s) only */
    if (seeding) {
      dlo = c_lo; dhi = c_hi;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const vi dl = lane + 64 * p, dh = dl + 32;
        Mn[p] = pack2(vsel((dl >= c_lo) & (dl <= c_hi), vmax(dl + kbase, splati(0)), splati(REG_NULL16)),
                      vsel((dh >= c_lo) & (dh <= c_hi), vmax(dh + kbase, splati(0)), splati(REG_NULL16)));
      }
    } else {
      /* the previous wavefront did not end the alignment: count it (unialign.c:241-273) */
      cells += wprev;
      ++s;
      if (!(flags & F_SOURCES)) {
        /* null step (allocate_output_null, compute.c:374-400) */
#pragma unroll
        for (int p = 0; p < P; ++p) Mn[p] = splat(REG_NULL2);
        push<0, P - 1>(Mn, false);
        wprev = 0;
        if (s >= s_limit) { status = 3; return true; }
        return false;
      }
      /* active window: diagonals within reach of the seeds, clipped to the DP matrix */
      const int reach = imax(s - (DOE - 1), 0);
      dlo = imax(c_lo - reach, m_lo); dhi = imin(c_hi + reach, m_hi);
      if (s >= s_event) {
        if (c_lo - reach < m_lo || c_hi + reach > m_hi) flags |= F_EXACT;
        if (dlo < 0 || dhi >= WIN) { status = 4; return true; }
        if (FULL) { if (s >= hrows) { status = 4; return true; } }
        s_event = next_event();
      }

      /* phases A + B on the packed registers the wavefront can occupy: while it stays inside the middle
       * register(s) (about half the scores of a typical pair on the 192- and 256-diagonal windows) the
       * outer ones hold nothing but nulls in every ring slot and are neither computed nor rotated */
      narrow = WFA_REG_NARROW && NLO > 0 && dlo >= 64 * NLO && dhi < 64 * (NHI + 1);
      if (narrow) {
        recurrence<NLO, NHI>(Mn);
#pragma unroll
        for (int p = 0; p < P; ++p) if (p < NLO || p > NHI) Mn[p] = splat(REG_NULL2);
      } else {
        recurrence<0, P - 1>(Mn);
      }
      const bool ex_o = (flags >> (DOE - 1)) & 1u;
      bool exIn = ex_o || (flags & F_I), exDn = ex_o || (flags & F_D);
      if (flags & F_EXACT) {
        /* An I/D offset can only leave the matrix after some offset has touched its edge, and the
         * first offset to do so is an M offset (M >= I, D on every diagonal), which raised F_EXACT
         * in extend_block.  From then on look for such offsets (offset - ub - 1 >= 0, per half);
         * trim_ends (compute.c:571-605) decides which of them survive. */
        vu over = splat(0x80008000u);
#pragma unroll
        for (int p = 0; p < P; ++p) over = vimax3(over, vadd2(I[p], ~ub2[p]), vadd2(D[p], ~ub2[p]));
        if (any((over & 0x80008000u) != 0x80008000u)) {
          exIn = trim_component(I);
          exDn = trim_component(D);
        }
      }
      flags = (flags & ~(F_I | F_D)) | (exIn ? F_I : 0u) | (exDn ? F_D : 0u);
    }

    /* phase C: extend the new M offsets (extend.c:90-125 / :263-297), active blocks only */
    const uint32_t bmask = (2u << (dhi >> 5)) - (1u << (dlo >> 5));
    extend_blocks<0>(Mn, bmask);
    wprev = dhi - dlo + 1;
    if (flags & F_EXACT) {
      scan_valid_range(Mn);
      wprev = cur_hi - cur_lo + 1;                    /* 0 when nothing is valid (cur_lo = cur_hi + 1) */
    }
    /* (far from the matrix edges the outermost diagonals are reached by one gap of `reach` bases and are valid) */
    if (narrow && WFA_REG_NARROW == 1) push<NLO, NHI>(Mn, wprev > 0);      /* (2: narrow recurrence, one rotation of the whole ring) */
    else push<0, P - 1>(Mn, wprev > 0);
    /* step limit first, then termination (unialign.c:241-273 order); score 0 has no limit check */
    if (!seeding && s >= s_limit) {
      status = 3;
      if (s == s_limit_exact) cells += wprev;
      return true;
    }
    if (term_d >= 0) { status = 1; cells += wprev; return true; }
    return false;
  }

  /* Run the alignment.  Returns PAIR_DONE / PAIR_OVERFLOW; end cell in (end_k, end_off). */
  WFA_DEV int run(int& end_k, int& end_off) {
    if (status == 0) {
      bool seeding = true;
#pragma unroll 1
      for (;;) { if (step(seeding)) break; seeding = false; }
    }
    if (status == 4) return PAIR_OVERFLOW;
    if (status == 1) { end_k = kbase + term_d; end_off = term_off; }
    return PAIR_DONE;
  }
};

/*
 * Align one pair on the register tier.  pw / tw: 2-bit packed words (any memory, readable one
 * word past the end; used by the backtrace) -- byte mode (CB = 4): the pair's bytes, 4 per word, and
 * wild = the wildcard byte or 0; pwin / twin: the sequence windows of build_windows<CB> in shared memory.  ops / runs_stage are per-warp scratch (scope=full).
 * is_leader: exactly one lane of the warp (device) or true (host model) -- it runs the
 * backtrace.  Returns PAIR_DONE (res filled by every lane except nruns / locs, which only the
 * leader knows and must be broadcast by the caller) or PAIR_OVERFLOW.
 */
template <int P, int DX, int DOE, bool FULL, bool HS = reg_hist_in_smem(P, FULL), int CB = 2>
WFA_DEV int align_pair_reg(const RegParams& R, const uint32_t* pw, const uint32_t* tw, lv::seqref pwin, lv::seqref twin,
                           int plen, int tlen, const lv::histref& hist, uint8_t* ops, uint32_t* runs_stage, bool is_leader,
                           PairResult& res, int wild = -1) {
  RegAligner<P, DX, DOE, FULL, HS, CB> A;
  A.init(R, pwin, twin, plen, tlen, hist);
  int end_k = 0, end_off = 0;
  if (A.run(end_k, end_off) == PAIR_OVERFLOW) return PAIR_OVERFLOW;
  res.cells = A.cells;
  res.nruns = 0;
  res.locs[0] = res.locs[1] = res.locs[2] = res.locs[3] = 0;
  const int end_score = (int)((long long)A.s * R.g);
  if (A.status == 3) {
    res.score = -R.max_steps; res.status = ST_MAX_STEPS;
  } else if (!FULL) {
    res.score = classic_score(R.match, plen, tlen, end_score, R.pos_score); res.status = ST_COMPLETED;
  } else {
    res.score = classic_score(R.match, end_off - end_k, end_off, end_score, R.pos_score);
    res.status = ST_COMPLETED;
    /* the leader reads origin codes every lane stored, and later overwrites the arena with the runs
     * (compute-sanitizer racecheck flagged the missing barrier as an intra-warp hazard) */
    lv::fence_warp();
    if (is_leader) {
      FwdEmitter em; em.init(runs_stage, R.runcap);
      const int n = backtrace_origin<A.HS>(hist, 32 * P, A.kbase, DX, DOE, 1, A.s, end_k, plen, tlen, pw, tw, ops, R.opcap, em, wild);
      res.nruns = n;
      if (n >= 0) locations_from_runs(runs_stage, imin(n, R.runcap), plen, tlen, res.locs);
    }
  }
  return PAIR_DONE;
}

}  // namespace wfagpu
