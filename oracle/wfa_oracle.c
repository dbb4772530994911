/*
 * wfa_oracle.c -- TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by,
 * or called from the product path (pywfa_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
 *
 * A plain-C CPU restatement of the reference algorithm on the hot path: the gap-affine and
 * gap-affine-2p wavefront alignment of WFA2-lib v2.3 as driven by pywfa 0.5.1
 * (`WavefrontAligner.wavefront_align`).  It is NOT a copy of the reference sources: the
 * reference's slab allocator, null/victim wavefronts and init-bounds bookkeeping are replaced
 * by one semantic rule -- "a wavefront component is a range [lo,hi] of offsets; everything
 * outside reads as NULL" -- which is what that machinery implements
 * (W/wavefront/wavefront_compute.c:490-567 init_ends, :571-605 trim_ends).
 * Each function cites the reference file:line it restates (W/ = pywfa/WFA2_lib/).
 *
 * PARITY PINNED: tests/test_oracle.py checks this file against (a) every known-answer vector
 * of the reference's own test-suite (pywfa/tests/test.py) and README, committed under
 * tests/golden/, and (b) the unmodified reference compiled here (oracle/_ref/libwfa_ref.so)
 * on seeded random batches for every supported configuration.
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "wfagpu.h"

#define OFFSET_NULL (INT32_MIN / 2)          /* W/wavefront/wavefront_offset.h:44 */
#define DIAGONAL_NULL INT_MAX                /* W/wavefront/wavefront_offset.h:54 */
#define MAXI(a, b) ((a) > (b) ? (a) : (b))
#define MINI(a, b) ((a) < (b) ? (a) : (b))

/* internal status, W/wavefront/wfa.h:52-55 */
#define ST_OK -1
#define ST_END_REACHED -2
#define ST_END_UNREACHABLE -3

enum { CM = 0, CI1 = 1, CD1 = 2, CI2 = 3, CD2 = 4, NCOMP = 5 };

/* backtrace_type priorities, W/wavefront/wavefront_backtrace.c:49-59 */
enum { BT_I1_OPEN = 1, BT_I1_EXT = 2, BT_I2_OPEN = 3, BT_I2_EXT = 4, BT_D1_OPEN = 5,
       BT_D1_EXT = 6, BT_D2_OPEN = 7, BT_D2_EXT = 8, BT_M = 9 };

/* One score's wavefront set.  off[c] is indexed by (k - clo); lo[c] > hi[c] means the
 * component is absent/null (wavefront_t.null or a NULL pointer in the reference). */
typedef struct {
  int exists;            /* mwavefronts[s] != NULL */
  int clo, chi;          /* computed (allocated) diagonal range */
  int32_t* off[NCOMP];
  uint8_t* code;         /* forward-recorded backtrace codes (bt_mode 1 only) */
  int lo[NCOMP], hi[NCOMP];
} wfset_t;

typedef struct {
  /* normalised penalties, W/wavefront/wavefront_penalties.c:95-173 */
  int match, x, o1, e1, o2, e2;
  int affine2p;
  int m_only;            /* gap-linear / edit / indel: M wavefronts only (compute_linear.c, compute_edit.c) */
  int no_mis;            /* indel: no mismatch source */
  int edit_like;         /* edit / indel: compute_edit.c's driver (no null steps, positive scores) */
  int max_scope;         /* W/wavefront/wavefront_components.c:81-124 */
  int plen, tlen;
  const char* p;
  const char* t;
  wfagpu_config_t cfg;
  int pbf, pef, tbf, tef;
  wfset_t* wf;           /* indexed by score */
  int wf_cap;
  /* status, W/wavefront/wfa.h:64-74 */
  int status, num_null_steps;
  int end_score, end_k, end_off;
  /* heuristic state, W/wavefront/wavefront_heuristic.h:50-69 */
  int steps_wait, max_sw_score, max_sw_score_k;
  int64_t cells;
  int bt_mode;
} oracle_t;

typedef struct {
  char* ops;             /* filled right-to-left like cigar->operations */
  int cap, begin, end;
  int score;
} cigar_buf_t;

/* ---- wavefront storage ------------------------------------------------------------- */
static wfset_t* wf_at(oracle_t* o, int s) {
  if (s >= o->wf_cap) {
    int ncap = o->wf_cap ? o->wf_cap : 64;
    while (ncap <= s) ncap *= 2;
    o->wf = (wfset_t*)realloc(o->wf, sizeof(wfset_t) * (size_t)ncap);
    memset(o->wf + o->wf_cap, 0, sizeof(wfset_t) * (size_t)(ncap - o->wf_cap));
    for (int i = o->wf_cap; i < ncap; ++i)
      for (int c = 0; c < NCOMP; ++c) { o->wf[i].lo[c] = 1; o->wf[i].hi[c] = -1; }
    o->wf_cap = ncap;
  }
  return &o->wf[s];
}
static void wf_release(wfset_t* w) {
  for (int c = 0; c < NCOMP; ++c) { free(w->off[c]); w->off[c] = NULL; w->lo[c] = 1; w->hi[c] = -1; }
  free(w->code); w->code = NULL;
  w->exists = 0;
}
/* Read with the "outside [lo,hi] is NULL" rule (compute.c:258-297 null wavefront,
 * :490-567 init_ends).  s < 0 or absent component -> NULL. */
static inline int32_t rd(const oracle_t* o, int s, int c, int k) {
  if (s < 0 || s >= o->wf_cap) return OFFSET_NULL;
  const wfset_t* w = &o->wf[s];
  if (k < w->lo[c] || k > w->hi[c]) return OFFSET_NULL;
  return w->off[c][k - w->clo];
}
static inline int comp_null(const oracle_t* o, int s, int c) {
  if (s < 0 || s >= o->wf_cap) return 1;
  return o->wf[s].lo[c] > o->wf[s].hi[c];
}

/* ---- penalties: wavefront_penalties_set_affine / _affine2p (penalties.c:95-173) ------ */
static void set_penalties(oracle_t* o) {
  const wfagpu_config_t* c = &o->cfg;
  o->affine2p = (c->distance == WFAGPU_DISTANCE_AFFINE2P);
  o->m_only = o->no_mis = o->edit_like = 0;
  if (c->distance == WFAGPU_DISTANCE_LINEAR) {
    /* wavefront_penalties_set_linear, penalties.c:62-93: the indel penalty sits in gap_opening1; here it
     * is the "extension" of a zero-cost opening, so that s - o1 - e1 addresses the open source */
    o->m_only = 1;
    o->o1 = 0; o->o2 = 0; o->e2 = 1;
    if (c->match < 0) { o->match = c->match; o->x = 2 * c->mismatch - 2 * c->match; o->e1 = 2 * c->gap_extension1 - c->match; }
    else { o->match = 0; o->x = c->mismatch; o->e1 = c->gap_extension1; }
    o->max_scope = MAXI(o->x, o->e1) + 1;                /* components.c:60-66 */
    return;
  }
  if (c->distance == WFAGPU_DISTANCE_EDIT || c->distance == WFAGPU_DISTANCE_INDEL) {
    /* wavefront_penalties_set_edit / _indel, penalties.c:38-61 */
    o->m_only = 1; o->edit_like = 1; o->no_mis = (c->distance == WFAGPU_DISTANCE_INDEL);
    o->match = 0; o->x = 1; o->o1 = 0; o->e1 = 1; o->o2 = 0; o->e2 = 1;
    o->max_scope = 2;                                    /* components_dimensions_edit, components.c:44-58 */
    return;
  }
  if (c->match < 0) {
    o->match = c->match;
    o->x = 2 * c->mismatch - 2 * c->match;
    o->o1 = 2 * c->gap_opening1;
    o->e1 = 2 * c->gap_extension1 - c->match;
    o->o2 = 2 * c->gap_opening2;
    o->e2 = 2 * c->gap_extension2 - c->match;
  } else {
    o->match = 0;
    o->x = c->mismatch; o->o1 = c->gap_opening1; o->e1 = c->gap_extension1;
    o->o2 = c->gap_opening2; o->e2 = c->gap_extension2;
  }
  /* components.c:89-90 (affine), :109-112 (affine2p) */
  int scope_indel = o->o1 + o->e1;
  if (o->affine2p) scope_indel = MAXI(scope_indel, o->o2 + o->e2);
  o->max_scope = MAXI(scope_indel, o->x) + 1;
}

/* wavefront_compute_classic_score, compute.c:108-120 */
static int classic_score(const oracle_t* o, int plen, int tlen, int wf_score) {
  const int swg_match = -o->match;
  if (o->edit_like) return wf_score;                     /* distance_metric <= edit, compute.c:117 */
  if (swg_match == 0) return -wf_score;
  /* WF_SCORE_TO_SW_SCORE, penalties.h:73 (int32 wrap-around, C truncating division) */
  const int32_t sum = (int32_t)((uint32_t)plen + (uint32_t)tlen);
  const int32_t prod = (int32_t)((uint32_t)swg_match * (uint32_t)sum);
  return (int32_t)((uint32_t)prod - (uint32_t)wf_score) / 2;
}

/* ---- ends-free seeding for match<0: compute.c:124-254 -------------------------------- */
static int endsfree_required(const oracle_t* o, int score) {      /* compute.c:124-138 */
  if (o->match == 0) return 0;
  if (o->cfg.span != WFAGPU_SPAN_ENDSFREE) return 0;
  if (o->tbf == 0 && o->pbf == 0) return 0;
  if (score % (-o->match) != 0) return 0;
  return 1;
}

/* ---- initial wavefront: wavefront_aligner_init_wf_m, aligner.c:251-310 --------------- */
static void init_wf0(oracle_t* o) {
  wfset_t* w = wf_at(o, 0);
  const int endsfree = (o->cfg.span == WFAGPU_SPAN_ENDSFREE) && (o->match == 0);
  /* aligner.c:260-261 widen [lo,hi] whenever match==0; for span=end-to-end with non-zero
   * begin-free values the reference then reads uninitialised slab memory (undefined
   * behaviour) -- we treat that case as lo=hi=0. */
  const int hi = endsfree ? o->tbf : 0;
  const int lo = endsfree ? -o->pbf : 0;
  w->exists = 1;
  w->clo = lo; w->chi = hi;
  w->off[CM] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(hi - lo + 1));
  if (o->bt_mode) w->code = (uint8_t*)calloc((size_t)(hi - lo + 1), 1);
  for (int k = lo; k <= hi; ++k) w->off[CM][k - lo] = (k > 0) ? k : 0;  /* aligner.c:278-303 */
  w->lo[CM] = lo; w->hi[CM] = hi;
}

/* ---- extend: extend_kernels.c:64-88 (the sentinels clamp at either sequence end) ----- */
static inline int32_t extend_cell(const oracle_t* o, int k, int32_t off) {
  int v = off - k, h = off;
  /* with a wildcard: wildcard_match_fun of pywfa/align.pyx:302-304 through wavefront_extend_matches_custom
   * (extend_kernels.c:167-203) and wavefront_sequences_cmp (wavefront_sequences.c:226-252) */
  const int w = o->cfg.wildcard;
  while (v < o->plen && h < o->tlen && (o->p[v] == o->t[h] || (w && ((uint8_t)o->p[v] == w || (uint8_t)o->t[h] == w)))) { ++v; ++h; ++off; }
  return off;
}
/* wavefront_termination_endsfree, termination.c:115-162 */
static int term_endsfree(oracle_t* o, int score, int k, int32_t off) {
  const int h = off, v = off - k;
  if (h >= o->tlen && o->plen - v <= o->pef) goto done;
  if (v >= o->plen && o->tlen - h <= o->tef) goto done;
  return 0;
done:
  o->end_score = score; o->end_k = k; o->end_off = off;
  return 1;
}

/* ---- heuristics: heuristic.c ---------------------------------------------------------- */
static void heur_adaptive(oracle_t* o, wfset_t* w) {               /* heuristic.c:257-293 */
  if (o->steps_wait > 0) return;
  const int base_lo = w->lo[CM], base_hi = w->hi[CM];
  if (base_hi - base_lo + 1 < o->cfg.min_wavefront_length) return;
  const int32_t* off = w->off[CM] - w->clo;
  /* wf_compute_distance_end2end, :176-192; wf_distance_end2end, :125-133 */
  int min_d = MAXI(o->plen, o->tlen);
  int* dist = (int*)malloc(sizeof(int) * (size_t)(base_hi - base_lo + 1));
  for (int k = base_lo; k <= base_hi; ++k) {
    const int32_t f = off[k];
    const int left_v = o->plen - (f - k), left_h = o->tlen - f;
    const int d = (f >= 0) ? MAXI(left_v, left_h) : -OFFSET_NULL;
    dist[k - base_lo] = d;
    min_d = MINI(min_d, d);
  }
  /* wf_heuristic_wfadaptive_reduce, :232-256 with min_k = max_k = alignment_k */
  const int thr = o->cfg.max_distance_threshold;
  const int ak = o->tlen - o->plen;
  const int top_limit = MINI(ak, w->hi[CM]);
  int lo_red = w->lo[CM];
  for (int k = w->lo[CM]; k < top_limit; ++k) {
    if (dist[k - base_lo] - min_d <= thr) break;
    ++lo_red;
  }
  w->lo[CM] = lo_red;
  const int bottom_limit = MAXI(ak, w->lo[CM]);
  int hi_red = w->hi[CM];
  for (int k = w->hi[CM]; k > bottom_limit; --k) {
    if (dist[k - base_lo] - min_d <= thr) break;
    --hi_red;
  }
  w->hi[CM] = hi_red;
  free(dist);
  o->steps_wait = o->cfg.steps_between_cutoffs;
}
static void heur_xdrop(oracle_t* o, wfset_t* w, int score) {       /* heuristic.c:329-383 */
  if (o->steps_wait > 0) return;
  const int32_t* off = w->off[CM] - w->clo;
  /* wf_heuristic_compute_sw_scores, :297-328 */
  const int swg_match = (o->match != 0) ? -o->match : -1;
  int cmax = INT_MIN, cmax_k = 0;
  const int base_lo = w->lo[CM], base_hi = w->hi[CM];
  int* sw = (int*)calloc((size_t)(base_hi - base_lo + 1), sizeof(int));
  for (int k = base_lo; k <= base_hi; ++k) {
    const int32_t f = off[k];
    if (f < 0) continue;
    const int v = f - k, h = f;
    const int s = (swg_match * (v + h) - score) / 2;
    sw[k - base_lo] = s;
    if (cmax < s) { cmax = s; cmax_k = k; }
  }
  if (o->max_sw_score_k != DIAGONAL_NULL) {
    const int xdrop = o->cfg.xdrop, max_prev = o->max_sw_score;
    int k;
    for (k = w->lo[CM]; k <= w->hi[CM]; ++k) {
      if (off[k] < 0) continue;
      if (max_prev - sw[k - base_lo] < xdrop) break;
    }
    w->lo[CM] = k;
    for (k = w->hi[CM]; k >= w->lo[CM]; --k) {
      if (off[k] < 0) continue;
      if (max_prev - sw[k - base_lo] < xdrop) break;
    }
    w->hi[CM] = k;
    if (cmax > o->max_sw_score) { o->max_sw_score = cmax; o->max_sw_score_k = cmax_k; }
  } else {
    o->max_sw_score = cmax; o->max_sw_score_k = cmax_k;
  }
  free(sw);
  o->steps_wait = o->cfg.steps_between_cutoffs;
}
/* wavefront_heuristic_cufoff, heuristic.c:509-567 (+ wf_heuristic_equate :161-172) */
static void heur_cutoff(oracle_t* o, int score) {
  wfset_t* w = &o->wf[score];
  if (!w->exists || w->lo[CM] > w->hi[CM]) return;
  --o->steps_wait;
  const int lo_base = w->lo[CM], hi_base = w->hi[CM];
  if (o->cfg.heuristic == WFAGPU_HEURISTIC_ADAPTIVE) heur_adaptive(o, w);
  else if (o->cfg.heuristic == WFAGPU_HEURISTIC_XDROP) heur_xdrop(o, w, score);
  if (lo_base == w->lo[CM] && hi_base == w->hi[CM]) return;
  for (int c = CI1; c < NCOMP; ++c) {
    if (w->lo[c] > w->hi[c]) continue;
    if (w->lo[CM] > w->lo[c]) w->lo[c] = w->lo[CM];
    if (w->hi[CM] < w->hi[c]) w->hi[c] = w->hi[CM];
  }
}

/* ---- extend step: wavefront_extend_end2end / _endsfree, extend.c:90-125, :263-297 ---- */
static int extend_step(oracle_t* o, int score) {
  wfset_t* w = wf_at(o, score);
  if (!w->exists) {                                                /* extend.c:99-106 */
    if (o->num_null_steps > o->max_scope) {
      o->status = ST_END_UNREACHABLE;
      o->end_score = score;          /* align_status.score; end_pos stays at its init value */
      return 1;
    }
    return 0;
  }
  int32_t* off = w->off[CM] - w->clo;
  const int lo = w->lo[CM], hi = w->hi[CM];
  if (o->cfg.span == WFAGPU_SPAN_END2END) {
    for (int k = lo; k <= hi; ++k) {                               /* extend_kernels.c:96-110 */
      if (off[k] == OFFSET_NULL) continue;
      off[k] = extend_cell(o, k, off[k]);
    }
    /* wavefront_termination_end2end, termination.c:37-61 */
    const int ak = o->tlen - o->plen;
    if (lo <= ak && ak <= hi && off[ak] >= o->tlen) {
      o->end_score = score; o->end_k = ak; o->end_off = o->tlen;
      o->status = ST_END_REACHED;
      o->cells += (hi >= lo) ? (hi - lo + 1) : 0;
      return 1;
    }
  } else {
    for (int k = lo; k <= hi; ++k) {                               /* extend_kernels.c:131-163 */
      if (off[k] == OFFSET_NULL) continue;
      off[k] = extend_cell(o, k, off[k]);
      if (term_endsfree(o, score, k, off[k])) {
        o->status = ST_END_REACHED;
        o->cells += (hi >= lo) ? (hi - lo + 1) : 0;
        return 1;
      }
    }
  }
  if (o->cfg.heuristic != WFAGPU_HEURISTIC_NONE) heur_cutoff(o, score);
  /* SURVEY.md 8(d) "C": final (post cut-off) M-wavefront widths, as ref_harness.c counts */
  o->cells += (w->hi[CM] >= w->lo[CM]) ? (w->hi[CM] - w->lo[CM] + 1) : 0;
  return 0;
}

/* ---- compute step: wavefront_compute_affine / _affine2p ------------------------------ */
static inline int inbounds(const oracle_t* o, int k, int32_t off) {
  const uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
  return h <= (uint32_t)o->tlen && v <= (uint32_t)o->plen;
}

/* wavefront_compute_edit_exact_prune, compute_edit.c:198-275: on wavefronts of >= 1000 diagonals drop the ends
 * whose best case |remaining v - remaining h| is worse than the best worst case max(remaining v, remaining h) */
static void edit_exact_prune(const oracle_t* o, wfset_t* w) {
  const int lo = w->lo[CM], hi = w->hi[CM];
  if (hi - lo + 1 < 1000) return;
  const int32_t* off = w->off[CM] - w->clo;
  const int ak = o->tlen - o->plen;
#define EBEST(k) ((k) >= ak ? (k) - ak : ak - (k))
#define EWORST(k) MAXI(o->plen - (off[k] - (k)), o->tlen - off[k])
  const int sample_k = lo + (hi - lo) / 2;
  if (off[sample_k] < 0) return;
  const int smax = EWORST(sample_k);
  if (EBEST(lo) <= smax && EBEST(hi) <= smax) return;
  int min_worst = INT_MAX;
  for (int k = lo; k <= hi; ++k) if (off[k] >= 0) min_worst = MINI(min_worst, EWORST(k));
  int nlo = lo, nhi = hi;
  for (int k = lo; k <= hi; ++k) { if (EBEST(k) <= min_worst) break; ++nlo; }
  for (int k = hi; k > nlo; --k) { if (EBEST(k) <= min_worst) break; --nhi; }
#undef EBEST
#undef EWORST
  w->lo[CM] = nlo; w->hi[CM] = nhi;
}
/* wavefront_compute_trim_ends, compute.c:571-605 */
static void trim(const oracle_t* o, wfset_t* w, int c) {
  const int32_t* off = w->off[c] - w->clo;
  int k;
  for (k = w->hi[c]; k >= w->lo[c]; --k) if (inbounds(o, k, off[k])) break;
  w->hi[c] = k;
  for (k = w->lo[c]; k <= w->hi[c]; ++k) if (inbounds(o, k, off[k])) break;
  w->lo[c] = k;
}
static void compute_step(oracle_t* o, int s) {
  const int two = o->affine2p;
  const int sx = s - o->x, so1 = s - o->o1 - o->e1, se1 = s - o->e1;
  const int so2 = s - o->o2 - o->e2, se2 = s - o->e2;
  wfset_t* w = wf_at(o, s);          /* may realloc o->wf: take it before any other pointer */
  /* fetch_input + null tests, compute.c:298-344, compute_affine.c:235-244, affine2p.c:340-353 */
  const int n_mx = o->no_mis ? 1 : comp_null(o, sx, CM), n_mo1 = comp_null(o, so1, CM);
  const int n_i1 = o->m_only ? 1 : comp_null(o, se1, CI1), n_d1 = o->m_only ? 1 : comp_null(o, se1, CD1);
  const int n_mo2 = two ? comp_null(o, so2, CM) : 1;
  const int n_i2 = two ? comp_null(o, se2, CI2) : 1, n_d2 = two ? comp_null(o, se2, CD2) : 1;
  const int ef_req = endsfree_required(o, s);
  const int efk = ef_req ? s / (-o->match) : 0;
  const int ef_t = ef_req && (o->tbf >= efk), ef_p = ef_req && (o->pbf >= efk);
  if (o->edit_like && n_mo1) {
    /* wavefront_compute_edit (compute_edit.c:329-374) has no null step: a null predecessor yields a null
     * wavefront and num_null_steps = INT_MAX, i.e. "unreachable" at the next extend */
    o->num_null_steps = INT_MAX;
    w->exists = 1; w->clo = 0; w->chi = 0;
    w->off[CM] = (int32_t*)malloc(sizeof(int32_t));
    w->off[CM][0] = OFFSET_NULL;
    if (o->bt_mode) w->code = (uint8_t*)calloc(1, 1);
    return;
  }
  if (n_mx && n_mo1 && n_i1 && n_d1 && n_mo2 && n_i2 && n_d2) {
    o->num_null_steps++;
    /* allocate_output_null, compute.c:374-400; endsfree_allocate_null :208-254.
     * (The reference creates the M wavefront whenever endsfree_required, even if neither
     * seed applies: lo=hi=0 with no valid offset; we keep that as an existing, empty M.) */
    if (ef_req) {
      int lo = 0, hi = 0;
      if (ef_t && ef_p) { lo = -efk; hi = efk; }
      else if (ef_t) { lo = efk; hi = efk; }
      else if (ef_p) { lo = -efk; hi = -efk; }
      w->exists = 1; w->clo = lo; w->chi = hi;
      w->off[CM] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(hi - lo + 1));
      if (o->bt_mode) w->code = (uint8_t*)calloc((size_t)(hi - lo + 1), 1);
      for (int k = lo; k <= hi; ++k) w->off[CM][k - lo] = OFFSET_NULL;
      if (ef_t) w->off[CM][efk - lo] = efk;
      if (ef_p) w->off[CM][-efk - lo] = 0;
      if (ef_t || ef_p) { w->lo[CM] = lo; w->hi[CM] = hi; }
    }
    return;
  }
  o->num_null_steps = 0;
  /* wavefront_compute_limits_input, compute.c:40-86; a null input contributes lo=1,hi=-1 */
#define LO_OF(isnull, sc, c) ((isnull) ? 1 : o->wf[sc].lo[c])
#define HI_OF(isnull, sc, c) ((isnull) ? -1 : o->wf[sc].hi[c])
  int lo = LO_OF(n_mx, sx, CM), hi = HI_OF(n_mx, sx, CM);
  lo = MINI(lo, LO_OF(n_mo1, so1, CM) - 1); hi = MAXI(hi, HI_OF(n_mo1, so1, CM) + 1);
  if (o->edit_like) { lo = o->wf[so1].lo[CM] - 1; hi = o->wf[so1].hi[CM] + 1; }   /* compute_edit.c:346-347 */
  if (!o->m_only) {                                                              /* compute.c:52-58: gap-linear stops here */
    lo = MINI(lo, LO_OF(n_i1, se1, CI1) + 1); hi = MAXI(hi, HI_OF(n_i1, se1, CI1) + 1);
    lo = MINI(lo, LO_OF(n_d1, se1, CD1) - 1); hi = MAXI(hi, HI_OF(n_d1, se1, CD1) - 1);
  }
  if (two) {
    lo = MINI(lo, LO_OF(n_mo2, so2, CM) - 1); hi = MAXI(hi, HI_OF(n_mo2, so2, CM) + 1);
    lo = MINI(lo, LO_OF(n_i2, se2, CI2) + 1); hi = MAXI(hi, HI_OF(n_i2, se2, CI2) + 1);
    lo = MINI(lo, LO_OF(n_d2, se2, CD2) - 1); hi = MAXI(hi, HI_OF(n_d2, se2, CD2) - 1);
  }
  /* allocate_output, compute.c:401-486: M always; I/D only if open or own ext is non-null */
  int clo = lo, chi = hi;
  if (ef_t) { chi = MAXI(chi, efk); clo = MINI(clo, efk); }
  if (ef_p) { clo = MINI(clo, -efk); chi = MAXI(chi, -efk); }
  const size_t width = (size_t)(chi - clo + 1);
  const int has[NCOMP] = {1, !o->m_only && (!n_mo1 || !n_i1), !o->m_only && (!n_mo1 || !n_d1),
                          two && (!n_mo2 || !n_i2), two && (!n_mo2 || !n_d2)};
  w->exists = 1; w->clo = clo; w->chi = chi;
  for (int c = 0; c < NCOMP; ++c) {
    if (!has[c]) continue;
    w->off[c] = (int32_t*)malloc(sizeof(int32_t) * width);
    for (size_t i = 0; i < width; ++i) w->off[c][i] = OFFSET_NULL;
    w->lo[c] = lo; w->hi[c] = hi;
  }
  if (o->bt_mode) w->code = (uint8_t*)calloc(width, 1);
  /* the recurrence: compute_affine.c:44-86, compute_affine2p.c:45-106 (falls back to the
   * affine kernel when every piece-2 input is null, :286-308 -- same values either way) */
  for (int k = lo; k <= hi; ++k) {
    const int32_t i1o = rd(o, so1, CM, k - 1), i1e = rd(o, se1, CI1, k - 1);
    const int32_t ins1 = MAXI(i1o, i1e) + 1;
    const int32_t d1o = rd(o, so1, CM, k + 1), d1e = rd(o, se1, CD1, k + 1);
    const int32_t del1 = MAXI(d1o, d1e);
    const int32_t misms = o->no_mis ? OFFSET_NULL : rd(o, sx, CM, k) + 1;
    int32_t ins = ins1, del = del1;
    int32_t ins2 = OFFSET_NULL, del2 = OFFSET_NULL;
    int32_t i2o = OFFSET_NULL, i2e = OFFSET_NULL, d2o = OFFSET_NULL, d2e = OFFSET_NULL;
    if (two && !(n_mo2 && n_i2 && n_d2)) {
      i2o = rd(o, so2, CM, k - 1); i2e = rd(o, se2, CI2, k - 1);
      ins2 = MAXI(i2o, i2e) + 1;
      d2o = rd(o, so2, CM, k + 1); d2e = rd(o, se2, CD2, k + 1);
      del2 = MAXI(d2o, d2e);
      ins = MAXI(ins1, ins2); del = MAXI(del1, del2);
    }
    int32_t mx = MAXI(del, MAXI(misms, ins));
    if (!inbounds(o, k, mx)) mx = OFFSET_NULL;
    const int i = k - clo;
    w->off[CM][i] = mx;
    if (has[CI1]) w->off[CI1][i] = ins1;
    if (has[CD1]) w->off[CD1][i] = del1;
    if (has[CI2]) w->off[CI2][i] = ins2;
    if (has[CD2]) w->off[CD2][i] = del2;
    if (o->bt_mode) {
      /* Forward-recorded origin code (the B200 kernels' scheme, cross-checked here on the
       * CPU): winner of max over (offset<<4 | type), backtrace.c:366-389, plus one
       * "came from extend" bit per I/D component (ext beats open on ties: 2>1, 6>5, ...). */
      int64_t best = -1; int bt = 0;
#define CAND(val, type) do { int64_t c_ = ((int64_t)(val) * 16) | (type); \
                             if ((val) >= 0 && c_ > best) { best = c_; bt = (type); } } while (0)
      CAND(misms, BT_M);
      CAND(i1o + 1, BT_I1_OPEN); CAND(i1e + 1, BT_I1_EXT);
      CAND(d1o, BT_D1_OPEN); CAND(d1e, BT_D1_EXT);
      if (two) {
        CAND(i2o + 1, BT_I2_OPEN); CAND(i2e + 1, BT_I2_EXT);
        CAND(d2o, BT_D2_OPEN); CAND(d2e, BT_D2_EXT);
      }
#undef CAND
      uint8_t code = (uint8_t)bt;
      if (i1e >= i1o) code |= 0x10;
      if (d1e >= d1o) code |= 0x20;
      if (i2e >= i2o) code |= 0x40;
      if (d2e >= d2o) code |= 0x80;
      w->code[i] = code;
    }
  }
  /* process_ends, compute.c:606-624: ends-free seeds (match<0), then trim every component */
  if (ef_req) {                                          /* endsfree_init, compute.c:163-207 */
    int32_t* off = w->off[CM] - clo;
    if (ef_t) {
      if (w->hi[CM] >= efk) { if (off[efk] <= efk) off[efk] = efk; }
      else { off[efk] = efk; w->hi[CM] = efk; }          /* gap cells already NULL-filled */
    }
    if (ef_p) {
      if (w->lo[CM] <= -efk) { if (off[-efk] <= 0) off[-efk] = 0; }
      else { off[-efk] = 0; w->lo[CM] = -efk; }
    }
  }
  for (int c = 0; c < NCOMP; ++c) if (has[c]) trim(o, w, c);
  if (o->edit_like) {
    /* compute_edit.c:367-373 */
    if (w->lo[CM] > w->hi[CM]) o->num_null_steps = INT_MAX;
    else if (!o->no_mis && o->cfg.span == WFAGPU_SPAN_END2END) edit_exact_prune(o, w);
  }
}

/* ---- backtrace ------------------------------------------------------------------------ */
static inline void cg_push(cigar_buf_t* cg, char op, int n) {
  while (n-- > 0) cg->ops[cg->begin--] = op;
}
/* candidate value with the type piggybacked in the low 4 bits, backtrace.c:64-219 */
static inline int64_t bt_cand(const oracle_t* o, int s, int c, int k, int add, int type) {
  if (s < 0 || s >= o->wf_cap) return OFFSET_NULL;
  const wfset_t* w = &o->wf[s];
  if (k < w->lo[c] || k > w->hi[c]) return OFFSET_NULL;
  return (((int64_t)(w->off[c][k - w->clo] + add)) * 16) | type;
}
/* wavefront_backtrace_affine, backtrace.c:320-529 (component_begin = component_end = M) */
static void backtrace(oracle_t* o, cigar_buf_t* cg, int a_score, int a_k, int32_t a_off) {
  const int two = o->affine2p;
  cg->end = cg->cap - 1;
  cg->begin = cg->cap - 2;
  int mt = CM, score = a_score, k = a_k;
  int h = a_off, v = a_off - a_k;
  int32_t off = a_off;
  if (v < o->plen) cg_push(cg, 'D', o->plen - v);
  if (h < o->tlen) cg_push(cg, 'I', o->tlen - h);
  while (v > 0 && h > 0 && score > 0) {
    const int sx = score - o->x, so1 = score - o->o1 - o->e1, se1 = score - o->e1;
    const int so2 = score - o->o2 - o->e2, se2 = score - o->e2;
    int64_t best;
    if (o->bt_mode) {
      /* follow the forward-recorded codes instead of re-deriving the max */
      const wfset_t* w = &o->wf[score];
      const uint8_t code = w->code[k - w->clo];
      int type;
      switch (mt) {
        case CM: type = code & 15; break;
        case CI1: type = (code & 0x10) ? BT_I1_EXT : BT_I1_OPEN; break;
        case CD1: type = (code & 0x20) ? BT_D1_EXT : BT_D1_OPEN; break;
        case CI2: type = (code & 0x40) ? BT_I2_EXT : BT_I2_OPEN; break;
        default:  type = (code & 0x80) ? BT_D2_EXT : BT_D2_OPEN; break;
      }
      switch (type) {
        case BT_M: best = bt_cand(o, sx, CM, k, 1, type); break;
        case BT_I1_OPEN: best = bt_cand(o, so1, CM, k - 1, 1, type); break;
        case BT_I1_EXT: best = bt_cand(o, se1, CI1, k - 1, 1, type); break;
        case BT_I2_OPEN: best = bt_cand(o, so2, CM, k - 1, 1, type); break;
        case BT_I2_EXT: best = bt_cand(o, se2, CI2, k - 1, 1, type); break;
        case BT_D1_OPEN: best = bt_cand(o, so1, CM, k + 1, 0, type); break;
        case BT_D1_EXT: best = bt_cand(o, se1, CD1, k + 1, 0, type); break;
        case BT_D2_OPEN: best = bt_cand(o, so2, CM, k + 1, 0, type); break;
        case BT_D2_EXT: best = bt_cand(o, se2, CD2, k + 1, 0, type); break;
        default: best = -1; break;
      }
    } else {
      const int64_t i1 = MAXI(bt_cand(o, so1, CM, k - 1, 1, BT_I1_OPEN),
                              bt_cand(o, se1, CI1, k - 1, 1, BT_I1_EXT));
      const int64_t d1 = MAXI(bt_cand(o, so1, CM, k + 1, 0, BT_D1_OPEN),
                              bt_cand(o, se1, CD1, k + 1, 0, BT_D1_EXT));
      const int64_t i2 = two ? MAXI(bt_cand(o, so2, CM, k - 1, 1, BT_I2_OPEN),
                                    bt_cand(o, se2, CI2, k - 1, 1, BT_I2_EXT)) : OFFSET_NULL;
      const int64_t d2 = two ? MAXI(bt_cand(o, so2, CM, k + 1, 0, BT_D2_OPEN),
                                    bt_cand(o, se2, CD2, k + 1, 0, BT_D2_EXT)) : OFFSET_NULL;
      switch (mt) {
        case CM: {
          const int64_t ms = o->no_mis ? OFFSET_NULL : bt_cand(o, sx, CM, k, 1, BT_M);   /* backtrace.c:261-263 */
          best = two ? MAXI(ms, MAXI(MAXI(i1, i2), MAXI(d1, d2))) : MAXI(ms, MAXI(i1, d1));
          break;
        }
        case CI1: best = i1; break;
        case CI2: best = i2; break;
        case CD1: best = d1; break;
        default: best = d2; break;
      }
    }
    if (best < 0) break;
    if (mt == CM) {
      const int max_off = (int)(best >> 4);
      cg_push(cg, 'M', off - max_off);
      off = max_off;
      v = off - k; h = off;
      if (v <= 0 || h <= 0) break;
    }
    const int type = (int)(best & 15);
    switch (type) {
      case BT_M: score = sx; mt = CM; break;
      case BT_I1_OPEN: score = so1; mt = CM; break;
      case BT_I1_EXT: score = se1; mt = CI1; break;
      case BT_I2_OPEN: score = so2; mt = CM; break;
      case BT_I2_EXT: score = se2; mt = CI2; break;
      case BT_D1_OPEN: score = so1; mt = CM; break;
      case BT_D1_EXT: score = se1; mt = CD1; break;
      case BT_D2_OPEN: score = so2; mt = CM; break;
      default: score = se2; mt = CD2; break;
    }
    if (type == BT_M) { cg_push(cg, 'X', 1); --off; }
    else if (type <= BT_I2_EXT) { cg_push(cg, 'I', 1); --k; --off; }
    else { cg_push(cg, 'D', 1); ++k; }
    v = off - k; h = off;
  }
  if (mt == CM) {
    if (v > 0 && h > 0) {
      const int n = MINI(v, h);
      cg_push(cg, 'M', n);
      v -= n; h -= n;
    }
    cg_push(cg, 'D', v);
    cg_push(cg, 'I', h);
  }
  ++cg->begin;
  cg->score = a_score;
}

/* cigar_maxtrim_gap_affine, W/alignment/cigar.c:473-528 (user penalties, match==0 -> -1) */
static int maxtrim_affine(const oracle_t* o, cigar_buf_t* cg) {
  const wfagpu_config_t* c = &o->cfg;
  const int match_score = (c->match != 0) ? c->match : -1;
  int max_score = 0, max_off = cg->begin;
  char last = '\0';
  int score = 0;
  for (int i = cg->begin; i < cg->end; ++i) {
    switch (cg->ops[i]) {
      case 'M': score -= match_score; break;
      case 'X': score -= c->mismatch; break;
      case 'I': score -= c->gap_extension1 + ((last == 'I') ? 0 : c->gap_opening1); break;
      case 'D': score -= c->gap_extension1 + ((last == 'D') ? 0 : c->gap_opening1); break;
    }
    last = cg->ops[i];
    if (max_score < score) { max_score = score; max_off = i; }
  }
  const int trimmed = (max_off != cg->end - 1);
  if (max_score == 0) { cg->begin = cg->end = 0; cg->score = INT32_MIN; }
  else { cg->end = max_off + 1; cg->score = max_score; }
  return trimmed;
}
/* cigar_maxtrim_gap_linear, W/alignment/cigar.c:419-472 (user penalties; the indel penalty travels in gap_extension1) */
static int maxtrim_linear(const oracle_t* o, cigar_buf_t* cg) {
  const wfagpu_config_t* c = &o->cfg;
  const int match_score = (c->match != 0) ? c->match : -1;
  int max_score = 0, max_off = cg->begin, score = 0;
  for (int i = cg->begin; i < cg->end; ++i) {
    switch (cg->ops[i]) {
      case 'M': score -= match_score; break;
      case 'X': score -= c->mismatch; break;
      default: score -= c->gap_extension1; break;
    }
    if (max_score < score) { max_score = score; max_off = i; }
  }
  const int trimmed = (max_off != cg->end - 1);
  if (max_score == 0) { cg->begin = cg->end = 0; cg->score = INT32_MIN; }
  else { cg->end = max_off + 1; cg->score = max_score; }
  return trimmed;
}
/* cigar_maxtrim_gap_affine2p, W/alignment/cigar.c:529-616 */
static int score_op_2p(const wfagpu_config_t* c, char op, int len) {
  switch (op) {
    case 'M': return ((c->match != 0) ? c->match : -1) * len;
    case 'X': return c->mismatch * len;
    default: {
      const int s1 = c->gap_opening1 + c->gap_extension1 * len;
      const int s2 = c->gap_opening2 + c->gap_extension2 * len;
      return MINI(s1, s2);
    }
  }
}
static int maxtrim_affine2p(const oracle_t* o, cigar_buf_t* cg) {
  const wfagpu_config_t* c = &o->cfg;
  if (cg->begin >= cg->end) return 0;
  int max_score = 0, max_off = cg->begin;
  char last = '\0';
  int score = 0, op_len = 0;
  for (int i = cg->begin; i < cg->end; ++i) {
    const char op = cg->ops[i];
    if (op != last && last != '\0') {
      score -= score_op_2p(c, last, op_len);
      op_len = 0;
      if (max_score < score) { max_score = score; max_off = i - 1; }
    }
    last = op;
    ++op_len;
  }
  score -= score_op_2p(c, last, op_len);
  if (max_score < score) { max_score = score; max_off = cg->end - 1; }
  const int trimmed = (max_off != cg->end - 1);
  if (max_score == 0) { cg->begin = cg->end = 0; cg->score = INT32_MIN; }
  else { cg->end = max_off + 1; cg->score = max_score; }
  return trimmed;
}

/* ---- driver: wavefront_unialign + _terminate, unialign.c:147-273 ---------------------- */
/*
 * Align one pair.  ops (may be NULL) receives the operation characters left-to-right;
 * returns their count or -(needed) if ops_cap is too small.  bt_mode: 0 = the reference's
 * backtrace (re-derive the max from stored offsets), 1 = forward-recorded origin codes.
 */
int oracle_align(const wfagpu_config_t* cfg, int bt_mode,
                 const char* pattern, int plen, const char* text, int tlen,
                 int32_t* out_score, int32_t* out_status,
                 char* ops, int ops_cap, int64_t* out_cells) {
  oracle_t o;
  memset(&o, 0, sizeof(o));
  o.cfg = *cfg;
  o.bt_mode = bt_mode;
  o.plen = plen; o.tlen = tlen; o.p = pattern; o.t = text;
  o.pbf = cfg->pattern_begin_free; o.pef = cfg->pattern_end_free;
  o.tbf = cfg->text_begin_free; o.tef = cfg->text_end_free;
  set_penalties(&o);
  const int full = (cfg->scope == WFAGPU_SCOPE_FULL);
  const int max_steps = (cfg->max_steps <= 0) ? INT_MAX : cfg->max_steps;
  /* wavefront_aligner_init, aligner.c:387-417 */
  o.status = ST_OK;
  o.num_null_steps = 0;
  o.end_score = -1; o.end_k = DIAGONAL_NULL; o.end_off = OFFSET_NULL;
  o.steps_wait = cfg->steps_between_cutoffs;                       /* heuristic.c:114-121 */
  o.max_sw_score = 0; o.max_sw_score_k = DIAGONAL_NULL;
  cigar_buf_t cg;
  cg.cap = full ? 2 * (plen + tlen) + 2 : 2;
  cg.ops = (char*)malloc((size_t)cg.cap);
  cg.begin = cg.end = 0;
  cg.score = INT32_MIN;                                            /* cigar_clear */
  init_wf0(&o);
  int score = 0, status;
  for (;;) {                                                       /* unialign.c:241-273 */
    if (extend_step(&o, score)) break;
    ++score;
    compute_step(&o, score);
    if (score >= max_steps) {                                      /* unialign.c:98-109 */
      cg.score = -max_steps;
      o.status = WFAGPU_STATUS_MAX_STEPS;
      if (o.wf[score].hi[CM] >= o.wf[score].lo[CM])   /* computed, never extended */
        o.cells += o.wf[score].hi[CM] - o.wf[score].lo[CM] + 1;
      break;
    }
    /* scope=score keeps only max_scope wavefronts (modular memory, components.c:92-93) */
    if (!full && score >= o.max_scope) wf_release(&o.wf[score - o.max_scope]);
  }
  status = o.status;
  if (status == ST_END_REACHED || status == ST_END_UNREACHABLE) {  /* unialign.c:147-237 */
    const int unreachable = (status == ST_END_UNREACHABLE);
    if (!full) {
      if (!unreachable) {
        cg.score = classic_score(&o, plen, tlen, score);
        status = WFAGPU_STATUS_COMPLETED;
      } else {
        /* end_v/end_h from the (never assigned) end position: int32 wrap-around */
        const int32_t end_v = (int32_t)((uint32_t)o.end_off - (uint32_t)o.end_k);
        cg.score = classic_score(&o, end_v, o.end_off, score);
        status = WFAGPU_STATUS_PARTIAL;
      }
    } else {
      if (o.end_off != OFFSET_NULL) backtrace(&o, &cg, score, o.end_k, o.end_off);
      if (unreachable) {
        /* wavefront_aligner_maxtrim_cigar, aligner.c:663-675: does not apply to edit / indel */
        if (o.edit_like) { /* the CIGAR and the score the backtrace left stay as they are */ }
        else if (o.m_only) (void)maxtrim_linear(&o, &cg);
        else if (o.affine2p) (void)maxtrim_affine2p(&o, &cg);
        else (void)maxtrim_affine(&o, &cg);
        status = WFAGPU_STATUS_PARTIAL;
      } else {
        cg.score = classic_score(&o, o.end_off - o.end_k, o.end_off, score);
        status = WFAGPU_STATUS_COMPLETED;
      }
    }
  }
  if (out_score) *out_score = cg.score;
  if (out_status) *out_status = status;
  if (out_cells) *out_cells = o.cells;
  int n = cg.end - cg.begin;
  if (n < 0) n = 0;
  int ret = n;
  if (ops) {
    if (n > ops_cap) ret = -n;
    else memcpy(ops, cg.ops + cg.begin, (size_t)n);
  }
  for (int s = 0; s < o.wf_cap; ++s) wf_release(&o.wf[s]);
  free(o.wf);
  free(cg.ops);
  return ret;
}

/* Batch driver, same layout as ref_align_batch (oracle/ref_harness.c). */
int oracle_align_batch(const wfagpu_config_t* cfg, int bt_mode, const uint8_t* seq,
                       const int64_t* p_off, const int32_t* p_len,
                       const int64_t* t_off, const int32_t* t_len, int64_t n,
                       int32_t* score, int32_t* status,
                       int64_t* ops_off, char* ops, int64_t ops_cap, int64_t* cells) {
  int64_t used = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (ops_off) ops_off[i] = used;
    int32_t sc, st;
    int64_t c;
    int64_t room = ops ? (ops_cap - used) : 0;
    if (room > INT_MAX) room = INT_MAX;
    const int r = oracle_align(cfg, bt_mode, (const char*)seq + p_off[i], p_len[i],
                               (const char*)seq + t_off[i], t_len[i], &sc, &st,
                               ops ? ops + used : NULL, (int)room, &c);
    if (r < 0) return -1;
    if (ops) used += r;
    if (score) score[i] = sc;
    if (status) status[i] = st;
    if (cells) cells[i] = c;
  }
  if (ops_off) ops_off[n] = used;
  return 0;
}
