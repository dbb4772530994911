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
 * scope=full.  One origin byte per cell is written to a per-warp arena (row = score, column =
 * window diagonal): bits 0-1 winner of M (1 mismatch, 2 insertion, 3 deletion; ties resolve
 * M > D > I as W/wavefront/wavefront_backtrace.c:49-59), bit 2 "I[s][k+1] extends" and bit 3
 * "D[s][k-1] extends" (ext >= open), stored at the source diagonal.  The backtrace walks these
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

constexpr uint32_t REG_NULL2 = 0xC000C000u;     /* two int16 nulls (-16384) */
constexpr int REG_NULL16 = -16384;
constexpr int REG_UB_MIN = -8192;              /* floor of ub[k] for diagonals left of the matrix */
constexpr uint32_t REG_ONE2 = 0x00010001u;

struct RegParams {
  int match, g, max_steps;
  int endsfree, pbf, pef, tbf, tef;
  int hrows;          /* scope=full: rows of the origin arena (scores 0..hrows-1) */
  int opcap;          /* edit-operation stack bytes */
  int runcap;         /* CIGAR run staging words */
};

/*
 * Backtrace over origin bytes (one thread): backward walk, then forward replay.
 * hist row stride = win; column = diagonal - kbase.  Returns the number of runs (may exceed
 * em.cap, then the CIGAR did not fit) or -1 if the operation stack overflowed.
 */
WFA_DEV int backtrace_origin(const uint8_t* hist, int win, int kbase, int dx, int doe, int de,
                             int s_end, int k_end, int plen, int tlen, const uint32_t* pw, const uint32_t* tw,
                             uint8_t* ops, int opcap, FwdEmitter& em) {
  int s = s_end, d = k_end - kbase, nops = 0;
  int mt = CM;
  while (s > 0) {
    const uint8_t* row = hist + (long long)s * win;
    int enter = mt;
    if (mt == CM) {
      const int w = row[d] & 3;
      if (w == 1) { if (nops < opcap) ops[nops] = EOP_X; ++nops; s -= dx; continue; }
      enter = (w == 2) ? CI1 : CD1;
    }
    if (enter == CI1) {
      const int ext = (row[d - 1] >> 2) & 1;
      if (nops < opcap) ops[nops] = ext ? EOP_I_EXT : EOP_I_OPEN;
      ++nops; --d;
      if (ext) { s -= de; mt = CI1; } else { s -= doe; mt = CM; }
    } else {
      const int ext = (row[d + 1] >> 3) & 1;
      if (nops < opcap) ops[nops] = ext ? EOP_D_EXT : EOP_D_OPEN;
      ++nops; ++d;
      if (ext) { s -= de; mt = CD1; } else { s -= doe; mt = CM; }
    }
  }
  if (nops > opcap) return -1;
  replay_ops(ops, nops, kbase + d, plen, tlen, pw, tw, em);
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
WFA_DEV void build_windows(const uint32_t* words, int len, uint32_t* win) {
  using namespace lv;
  const vi lane = lane_id();
  for (int i0 = 0; i0 <= len; i0 += 32) {
    const vi i = lane + i0;
    const vb in = i <= len;
    const vi j = i >> 4;
    const vu w = vfunnel_r(gather_u32(words, j, in), gather_u32(words, j + 1, in), (i & 15) << 1);
    scatter_u32(win, i, vbrev(w), in);
  }
}

template <int P, int DX, int DOE, bool FULL>
struct RegAligner {
  static constexpr int RM = DX > DOE ? DX : DOE;   /* M ring: M[r] = wavefront of score s-r */
  static constexpr int DE = 1;
  static constexpr int NB = 2 * P;                 /* 32-diagonal blocks */
  static constexpr int WIN = 64 * P;

  /* wavefront registers */
  lv::vu M[RM][P], I[P], D[P];
  lv::vu ub2[P];                                   /* packed ub[k] per register */
  /* warp-uniform state */
  bool exM[RM], exI, exD;
  bool exact;                                      /* sticky: wavefront extents must be scanned, not derived from reach(s) */
  bool endsfree;
  int pef, tef;
  lv::seqref pwin, twin;                           /* sequence windows (shared memory) */
  uint8_t* hist;
  int hrows;
  int plen, tlen, kbase, lo0, hi0;
  int s, s_limit, s_limit_exact;                   /* step limit in units of g (unialign.c:98-109) */
  int cells;
  int term_d, term_off;                            /* window diagonal and offset of the end cell (term_d < 0: none) */
  int dak;                                         /* window diagonal of the end-to-end target tlen - plen */
  int status;                                      /* 0 running, 1 end reached, 3 max steps, 4 overflow */
  int cur_lo, cur_hi;                              /* window range [first, last] of the current M wavefront */
  lv::vi lane;
  lv::vu selL, selR;                               /* PRMT selectors mending the block seams */

  /* ---- extension of one block of 32 diagonals (extend_kernels.c:64-110) -------------- */
  WFA_DEV lv::vi extend_block(lv::vi off, lv::vi ubk, lv::vi k, lv::vb valid) {
    using namespace lv;
    const vi rem = ubk - off;                       /* bases left on the diagonal: min(plen - v, tlen - h) */
    const vi v = off - k;
    const vu x = load_win(pwin, v, valid) ^ load_win(twin, off, valid);
    vi n = vmin(vclz(x) >> 1, rem);                 /* x == 0: 16 bases agree */
    vb more = valid & (n == 16) & (n < rem);
    while (any(more)) {
      const vu y = load_win(pwin, v + n, more) ^ load_win(twin, off + n, more);
      const vi n2 = vmin(n + (vclz(y) >> 1), rem);
      const vb cont = more & (n2 == n + 16) & (n2 < rem);
      n = vsel(more, n2, n);
      more = cont;
    }
    return vsel(valid, off + n, off);
  }

  /* ---- after a block was extended: edge / termination bookkeeping -------------------- */
  WFA_DEV void after_extend(int b, lv::vi off, lv::vi ubk, lv::vi k, lv::vb valid) {
    using namespace lv;
    const uint32_t eb = ballot(valid & (off == ubk));
    if (eb == 0) return;
    exact = true;                                     /* a cell touches the matrix edge */
    if (term_d >= 0) return;
    if (endsfree) {                                   /* termination.c:115-162 */
      const vi vv = off - k;
      const vb t = valid & (((off >= tlen) & (vv >= plen - pef)) | ((vv >= plen) & (off >= tlen - tef)));
      const uint32_t tb = ballot(t);
      if (tb) { term_d = 32 * b + first_set(tb); term_off = lane_value(off, first_set(tb)); }
    } else {                                          /* termination.c:37-61 */
      if ((dak >> 5) == b && ((eb >> (dak & 31)) & 1u)) { term_d = dak; term_off = lane_value(off, dak & 31); }
    }
  }

  /* first / last window diagonal holding a valid offset of the newest M wavefront */
  WFA_DEV void scan_valid_range() {
    using namespace lv;
    cur_lo = 1; cur_hi = -1;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t b0 = ballot(sx_lo(M[0][p]) >= 0), b1 = ballot(sx_hi(M[0][p]) >= 0);
      if (b0) { if (cur_lo > cur_hi) cur_lo = 64 * p + first_set(b0); cur_hi = 64 * p + last_set(b0); }
      if (b1) { if (cur_lo > cur_hi) cur_lo = 64 * p + 32 + first_set(b1); cur_hi = 64 * p + 32 + last_set(b1); }
    }
  }

  /* exact trim_ends of an I or D wavefront (compute.c:571-605) */
  WFA_DEV void trim_component(lv::vu (&X)[P], bool& exists) {
    using namespace lv;
    int first = -1, last = -1;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const vi a = sx_lo(X[p]), b = sx_hi(X[p]);
      const uint32_t b0 = ballot((a >= 0) & (a <= sx_lo(ub2[p]))), b1 = ballot((b >= 0) & (b <= sx_hi(ub2[p])));
      if (b0) { if (first < 0) first = 64 * p + first_set(b0); last = 64 * p + last_set(b0); }
      if (b1) { if (first < 0) first = 64 * p + 32 + first_set(b1); last = 64 * p + 32 + last_set(b1); }
    }
    exists = first >= 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const vi dl = lane + 64 * p, dh = lane + (64 * p + 32);
      const vb kl = (dl >= first) & (dl <= last), kh = (dh >= first) & (dh <= last);
      const vu mask = vselu(kl, splat(0x0000ffffu), splat(0u)) | vselu(kh, splat(0xffff0000u), splat(0u));
      X[p] = bitsel(mask, X[p], splat(REG_NULL2));
    }
  }

  /* ---- per-pair set-up; score 0 = wavefront_aligner_init_wf_m (wavefront_aligner.c:251-310) -- */
  WFA_DEV void init(const RegParams& R, lv::seqref pwin_, lv::seqref twin_, int plen_, int tlen_, uint8_t* hist_) {
    using namespace lv;
    pwin = pwin_; twin = twin_; plen = plen_; tlen = tlen_; hist = hist_; hrows = R.hrows;
    endsfree = R.endsfree != 0; pef = R.pef; tef = R.tef;
    lane = lane_id();
    const bool ef = R.endsfree && R.match == 0;
    lo0 = ef ? -R.pbf : 0; hi0 = ef ? R.tbf : 0;
    kbase = ((lo0 + hi0) >> 1) - WIN / 2;
    dak = tlen - plen - kbase;
    /* so >= max_steps  <=>  s >= ceil(max_steps / g);  so == max_steps  <=>  s == max_steps / g exactly */
    s_limit = R.max_steps / R.g + (R.max_steps % R.g != 0);
    s_limit_exact = (R.max_steps % R.g == 0) ? R.max_steps / R.g : -1;
    s = 0; cells = 0; term_d = -1; term_off = 0;
    status = (lo0 < kbase || hi0 >= kbase + WIN) ? 4 : 0;
    exI = exD = false; exact = false;
    selL = vselu(lane == 0, splat(0x5432u), splat(0x7654u));
    selR = vselu(lane == 31, splat(0x5432u), splat(0x3210u));
#pragma unroll
    for (int r = 0; r < RM; ++r) {
      exM[r] = false;
#pragma unroll
      for (int p = 0; p < P; ++p) M[r][p] = splat(REG_NULL2);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      I[p] = splat(REG_NULL2); D[p] = splat(REG_NULL2);
      const vi kl = lane + (kbase + 64 * p), kh = kl + 32;
      ub2[p] = pack2(vmax(vmin(splati(tlen), kl + plen), splati(REG_UB_MIN)), vmax(vmin(splati(tlen), kh + plen), splati(REG_UB_MIN)));
    }
  }

  /* ---- one score step: score 0 seeds the wavefront, every later score computes it -------- */
  WFA_DEV bool step(bool seeding) {
    using namespace lv;
    vu Mn[P];
    int wlo, whi;
    if (seeding) {
      wlo = lo0; whi = hi0;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const vi kl = lane + (kbase + 64 * p), kh = kl + 32;
        Mn[p] = pack2(vsel((kl >= lo0) & (kl <= hi0), vmax(kl, splati(0)), splati(REG_NULL16)),
                      vsel((kh >= lo0) & (kh <= hi0), vmax(kh, splati(0)), splati(REG_NULL16)));
      }
    } else {
      /* the previous wavefront did not end the alignment: count it (unialign.c:241-273) */
      if (exM[0]) cells += cur_hi - cur_lo + 1;
      ++s;
      const bool ex_x = exM[DX - 1], ex_o = exM[DOE - 1];
      if (!(ex_x | ex_o | exI | exD)) {
        /* null step (allocate_output_null, compute.c:374-400) */
        rotate();
        if (s >= s_limit) { status = 3; return true; }
        return false;
      }
      /* active window: diagonals within reach of the seeds, clipped to the DP matrix */
      const int reach = s >= DOE ? s - DOE + 1 : 0;
      wlo = lo0 - reach; whi = hi0 + reach;
      if (wlo < -plen) { wlo = -plen; exact = true; }
      if (whi > tlen) { whi = tlen; exact = true; }
      if (wlo < kbase || whi >= kbase + WIN) { status = 4; return true; }
      if (FULL) { if (s >= hrows) { status = 4; return true; } }

      /* phase A: per source diagonal max(open, extend), rotated to the consuming lane */
      vu rl[P], rr[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        rl[p] = from_prev_lane(vimax2(M[DOE - 1][p], I[p]));
        rr[p] = from_next_lane(vimax2(M[DOE - 1][p], D[p]));
      }
      /* phase B: the recurrence (compute_affine.c:44-86), 64 diagonals per instruction */
      uint8_t* const hrow = FULL ? hist + s * WIN : nullptr;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const vu L = prmt(p > 0 ? rl[p > 0 ? p - 1 : 0] : splat(REG_NULL2), rl[p], selL);
        const vu Rr = prmt(rr[p], p < P - 1 ? rr[p < P - 1 ? p + 1 : p] : splat(REG_NULL2), selR);
        const vu ins = vadd2(L, splat(REG_ONE2));
        const vu del = Rr;
        const vu mis = vadd2(M[DX - 1][p], splat(REG_ONE2));
        vu m;
        if (FULL) {
          vb xh, xl, yh, yl, ah, al, bh, bl;
          (void)vimax2p(I[p], M[DOE - 1][p], xh, xl);    /* I[s][k+1] extends: ext >= open */
          (void)vimax2p(D[p], M[DOE - 1][p], yh, yl);    /* D[s][k-1] extends */
          const vu m1 = vimax2p(mis, del, ah, al);       /* mismatch beats deletion on ties */
          m = vimax2p(m1, ins, bh, bl);                  /* both beat insertion on ties */
          const vi cl = vsel(bl, vsel(al, splati(1), splati(3)), splati(2)) | vsel(xl, splati(4), splati(0)) | vsel(yl, splati(8), splati(0));
          const vi ch = vsel(bh, vsel(ah, splati(1), splati(3)), splati(2)) | vsel(xh, splati(4), splati(0)) | vsel(yh, splati(8), splati(0));
          scatter_u8(hrow, lane + 64 * p, cl, lane >= 0);
          scatter_u8(hrow, lane + (64 * p + 32), ch, lane >= 0);
        } else {
          m = vimax3(mis, ins, del);
        }
        /* offsets beyond the matrix are nulled: M > ub  <=>  M + ~ub >= 0 */
        Mn[p] = bitsel(signmask2(vadd2(m, ~ub2[p])), m, splat(REG_NULL2));
        I[p] = ins; D[p] = del;
      }
      bool exIn = ex_o | exI, exDn = ex_o | exD;
      if (exact) {
        /* An I/D offset can only leave the matrix after some offset has touched its edge, and the
         * first offset to do so is an M offset (M >= I, D on every diagonal), which raised `exact`
         * in after_extend.  From then on look for such offsets (offset - ub - 1 >= 0, per half);
         * trim_ends (compute.c:571-605) decides which of them survive. */
        vu over = splat(0x80008000u);
#pragma unroll
        for (int p = 0; p < P; ++p) over = vimax3(over, vadd2(I[p], ~ub2[p]), vadd2(D[p], ~ub2[p]));
        if (any((over & 0x80008000u) != 0x80008000u)) {
          trim_component(I, exIn);
          trim_component(D, exDn);
        }
      }
      exI = exIn; exD = exDn;
      rotate();
    }

    /* phase C: extend the new M offsets (extend.c:90-125 / :263-297), active blocks only */
    const int blo = (wlo - kbase) >> 5, bhi = (whi - kbase) >> 5;
    const uint32_t bmask = (2u << bhi) - (1u << blo);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (((bmask >> (2 * p)) & 3u) == 0) continue;
      vi o[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int b = 2 * p + hh;
        vi off = hh ? sx_hi(Mn[p]) : sx_lo(Mn[p]);
        if ((bmask >> b) & 1u) {
          const vi k = lane + (kbase + 32 * b);
          const vi ubk = hh ? sx_hi(ub2[p]) : sx_lo(ub2[p]);
          const vb valid = off >= 0;
          off = extend_block(off, ubk, k, valid);
          after_extend(b, off, ubk, k, valid);
        }
        o[hh] = off;
      }
      Mn[p] = pack2(o[0], o[1]);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) M[0][p] = Mn[p];
    if (exact) {
      scan_valid_range();
      exM[0] = cur_lo <= cur_hi;
    } else {
      /* far from the matrix edges the outermost diagonals are reached by one gap of `reach`
       * bases and are valid */
      exM[0] = true; cur_lo = wlo - kbase; cur_hi = whi - kbase;
    }
    /* step limit first, then termination (unialign.c:241-273 order); score 0 has no limit check */
    if (!seeding && s >= s_limit) {
      status = 3;
      if (s == s_limit_exact && exM[0]) cells += cur_hi - cur_lo + 1;
      return true;
    }
    if (term_d >= 0) { status = 1; if (exM[0]) cells += cur_hi - cur_lo + 1; return true; }
    return false;
  }

  /* age the M ring by one score; M[0] becomes null (and is then overwritten by the new wavefront) */
  WFA_DEV void rotate() {
    using namespace lv;
#pragma unroll
    for (int r = RM - 1; r > 0; --r) {
      exM[r] = exM[r - 1];
#pragma unroll
      for (int p = 0; p < P; ++p) M[r][p] = M[r - 1][p];
    }
    exM[0] = false;
#pragma unroll
    for (int p = 0; p < P; ++p) M[0][p] = splat(REG_NULL2);
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
 * word past the end; used by the backtrace); pwin / twin: the sequence windows of
 * build_windows in shared memory.  ops / runs_stage are per-warp scratch (scope=full).
 * is_leader: exactly one lane of the warp (device) or true (host model) -- it runs the
 * backtrace.  Returns PAIR_DONE (res filled by every lane except nruns / locs, which only the
 * leader knows and must be broadcast by the caller) or PAIR_OVERFLOW.
 */
template <int P, int DX, int DOE, bool FULL>
WFA_DEV int align_pair_reg(const RegParams& R, const uint32_t* pw, const uint32_t* tw, lv::seqref pwin, lv::seqref twin,
                           int plen, int tlen, uint8_t* hist, uint8_t* ops, uint32_t* runs_stage, bool is_leader,
                           PairResult& res) {
  RegAligner<P, DX, DOE, FULL> A;
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
    res.score = classic_score(R.match, plen, tlen, end_score); res.status = ST_COMPLETED;
  } else {
    res.score = classic_score(R.match, end_off - end_k, end_off, end_score);
    res.status = ST_COMPLETED;
    if (is_leader) {
      FwdEmitter em; em.init(runs_stage, R.runcap);
      const int n = backtrace_origin(hist, A.WIN, A.kbase, DX, DOE, 1, A.s, end_k, plen, tlen, pw, tw, ops, R.opcap, em);
      res.nruns = n;
      if (n >= 0) locations_from_runs(runs_stage, imin(n, R.runcap), plen, tlen, res.locs);
    }
  }
  return PAIR_DONE;
}

}  // namespace wfagpu
