/* wfa_params.h -- configuration POD -> kernel parameters (internal; shared with tests/emu). */
#pragma once
#include <limits.h>

#include <algorithm>

#include "../../include/wfagpu.h"
#include "wfa_core.cuh"

namespace wfagpu {

/* Normalised penalties (W/wavefront/wavefront_penalties.c:95-173: Eizenga's transform when
 * match < 0), ring geometry (max_score_scope, W/wavefront/wavefront_components.c:81-124) and
 * the alignment form / heuristic fields of wavefront_aligner_attr_t. */
inline void fill_kparams(const wfagpu_config_t& c, KParams& k) {
  const bool two_p = c.distance == WFAGPU_DISTANCE_AFFINE2P;
  if (c.match < 0) {
    k.match = c.match;
    k.x = 2 * c.mismatch - 2 * c.match;
    k.o1 = 2 * c.gap_opening1; k.e1 = 2 * c.gap_extension1 - c.match;
    k.o2 = 2 * c.gap_opening2; k.e2 = 2 * c.gap_extension2 - c.match;
  } else {
    k.match = 0; k.x = c.mismatch;
    k.o1 = c.gap_opening1; k.e1 = c.gap_extension1;
    k.o2 = c.gap_opening2; k.e2 = c.gap_extension2;
  }
  k.m_only = k.no_mis = k.edit_like = k.edit_prune = k.pos_score = 0;
  if (c.distance == WFAGPU_DISTANCE_LINEAR) {
    /* wavefront_penalties_set_linear, penalties.c:62-93: one indel penalty (carried in gap_extension1 as pywfa does,
     * align.pyx:351-355), here the extension of a zero-cost opening: s - o1 - e1 addresses the gap source */
    k.m_only = 1; k.o1 = 0;
  } else if (c.distance == WFAGPU_DISTANCE_EDIT || c.distance == WFAGPU_DISTANCE_INDEL) {
    /* wavefront_penalties_set_edit / _indel, penalties.c:38-61 */
    k.m_only = k.edit_like = k.pos_score = 1; k.no_mis = c.distance == WFAGPU_DISTANCE_INDEL;
    k.edit_prune = c.distance == WFAGPU_DISTANCE_EDIT && c.span == WFAGPU_SPAN_END2END;
    k.match = 0; k.x = 1; k.o1 = 0; k.e1 = 1;
  }
  int scope_indel = k.o1 + k.e1;
  if (two_p) scope_indel = std::max(scope_indel, k.o2 + k.e2);
  k.max_scope = k.edit_like ? 2 : std::max(scope_indel, k.x) + 1;     /* components.c:44-66 / :81-124 */
  if (!two_p) { k.o2 = 0; k.e2 = 1; }
  /* every reachable score is a sum of x, o+e and e terms: step in units of their gcd */
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  int g = gcd(gcd(k.x, k.o1 + k.e1), k.e1);
  if (two_p) g = gcd(gcd(g, k.o2 + k.e2), k.e2);
  k.g = g;
  k.dx = k.x / g; k.doe1 = (k.o1 + k.e1) / g; k.de1 = k.e1 / g;
  k.doe2 = two_p ? (k.o2 + k.e2) / g : 1; k.de2 = two_p ? k.e2 / g : 1;
  k.rm = std::max(k.dx, std::max(k.doe1, two_p ? k.doe2 : 0)) + 1;
  k.r1 = k.de1 + 1;
  k.r2 = two_p ? k.de2 + 1 : 1;
  k.mr = 1;
  while (k.mr < k.rm + 1) k.mr <<= 1;     /* one spare entry: the planner warp of wfa_vec.cuh publishes score s+1 while score s-rm+1 may still be read */
  k.endsfree = c.span == WFAGPU_SPAN_ENDSFREE;
  k.pbf = c.pattern_begin_free; k.pef = c.pattern_end_free;
  k.tbf = c.text_begin_free; k.tef = c.text_end_free;
  k.heuristic = c.heuristic;
  k.min_wf_len = c.min_wavefront_length; k.max_dist_thr = c.max_distance_threshold;
  k.steps_between = c.steps_between_cutoffs; k.xdrop = c.xdrop;
  k.max_steps = c.max_steps <= 0 ? INT_MAX : c.max_steps;
}

/* Score-only alignments of the M-only metrics without a cut-off are gap-affine alignments with a zero-cost
 * opening: gap-linear (x, e) = affine (x, 0, e); edit = affine (1, 0, 1); indel = affine (2, 0, 1), a mismatch
 * costing what an insertion plus a deletion cost.  Score and status are the optimum's either way (WFA is exact;
 * the two recurrences differ in wavefront ranges and CIGAR tie-breaks only), so these run on the fast gap-affine
 * tiers.  Returns whether `k` (filled by fill_kparams) was rewritten. */
inline bool metric_as_affine(const wfagpu_config_t& c, KParams& k) {
  if (!k.m_only || c.scope != WFAGPU_SCOPE_SCORE || c.heuristic != WFAGPU_HEURISTIC_NONE) return false;
  wfagpu_config_t a = c;
  a.distance = WFAGPU_DISTANCE_AFFINE;
  a.gap_opening1 = 0;
  if (k.edit_like) { a.match = 0; a.mismatch = k.no_mis ? 2 : 1; a.gap_extension1 = 1; }
  const int pos = k.pos_score, scope = k.max_scope;
  fill_kparams(a, k);
  k.pos_score = pos; k.max_scope = scope;
  return true;
}

}  // namespace wfagpu
