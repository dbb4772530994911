/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * A thin batch driver around the UNMODIFIED reference WFA2-lib, compiled from
 * the sources where they lie under /root/reference (see oracle/Makefile; the
 * objects land in oracle/_ref/ which is git-ignored).  It only calls the
 * reference's public API the same way pywfa/align.pyx does:
 *   wavefront_aligner_new   (align.pyx:419)
 *   wavefront_align         (align.pyx:439)
 *   reads of aligner->cigar / ->align_status (align.pyx:443,463,731-786)
 *   wavefront_aligner_delete (align.pyx:883)
 * Nothing here is shipped or measured as the product; tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs are the only callers.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#include "wavefront/wavefront_align.h"
#include "wfagpu.h"

typedef struct {
  wavefront_aligner_t* aligner;
  wfagpu_config_t cfg;
} ref_handle_t;

/* kwargs -> wavefront_aligner_attr_t, exactly as pywfa/align.pyx:343-417 */
void* ref_new(const wfagpu_config_t* cfg, int memory_mode) {
  wavefront_aligner_attr_t attr = wavefront_aligner_attr_default;
  if (cfg->distance == WFAGPU_DISTANCE_INDEL) {
    attr.distance_metric = indel;                     /* align.pyx:347-348 */
  } else if (cfg->distance == WFAGPU_DISTANCE_EDIT) {
    attr.distance_metric = edit;                      /* align.pyx:349-350 */
  } else if (cfg->distance == WFAGPU_DISTANCE_LINEAR) {
    attr.distance_metric = gap_linear;                /* align.pyx:351-355 */
    attr.linear_penalties.match = cfg->match;
    attr.linear_penalties.mismatch = cfg->mismatch;
    attr.linear_penalties.indel = cfg->gap_extension1;
  } else if (cfg->distance == WFAGPU_DISTANCE_AFFINE) {
    attr.distance_metric = gap_affine;
    attr.affine_penalties.match = cfg->match;
    attr.affine_penalties.mismatch = cfg->mismatch;
    attr.affine_penalties.gap_opening = cfg->gap_opening1;
    attr.affine_penalties.gap_extension = cfg->gap_extension1;
  } else {
    attr.distance_metric = gap_affine_2p;
    attr.affine2p_penalties.match = cfg->match;
    attr.affine2p_penalties.mismatch = cfg->mismatch;
    attr.affine2p_penalties.gap_opening1 = cfg->gap_opening1;
    attr.affine2p_penalties.gap_extension1 = cfg->gap_extension1;
    attr.affine2p_penalties.gap_opening2 = cfg->gap_opening2;
    attr.affine2p_penalties.gap_extension2 = cfg->gap_extension2;
  }
  attr.alignment_scope = (cfg->scope == WFAGPU_SCOPE_FULL) ? compute_alignment : compute_score;
  attr.memory_mode = (wavefront_memory_t)memory_mode;
  attr.alignment_form.pattern_begin_free = cfg->pattern_begin_free;
  attr.alignment_form.pattern_end_free = cfg->pattern_end_free;
  attr.alignment_form.text_begin_free = cfg->text_begin_free;
  attr.alignment_form.text_end_free = cfg->text_end_free;
  attr.alignment_form.span =
      (cfg->span == WFAGPU_SPAN_ENDSFREE) ? alignment_endsfree : alignment_end2end;
  if (cfg->heuristic == WFAGPU_HEURISTIC_NONE) {
    attr.heuristic.strategy = wf_heuristic_none;
  } else if (cfg->heuristic == WFAGPU_HEURISTIC_ADAPTIVE) {
    attr.heuristic.strategy = wf_heuristic_wfadaptive;
    attr.heuristic.min_wavefront_length = cfg->min_wavefront_length;
    attr.heuristic.max_distance_threshold = cfg->max_distance_threshold;
    attr.heuristic.steps_between_cutoffs = cfg->steps_between_cutoffs;
  } else {
    attr.heuristic.strategy = wf_heuristic_xdrop;
    attr.heuristic.xdrop = cfg->xdrop;
    attr.heuristic.steps_between_cutoffs = cfg->steps_between_cutoffs;
  }
  attr.system.max_alignment_steps = (cfg->max_steps <= 0) ? INT_MAX : cfg->max_steps;
  ref_handle_t* h = (ref_handle_t*)calloc(1, sizeof(ref_handle_t));
  h->aligner = wavefront_aligner_new(&attr);
  h->cfg = *cfg;
  return h;
}

void ref_delete(void* handle) {
  ref_handle_t* h = (ref_handle_t*)handle;
  if (!h) return;
  wavefront_aligner_delete(h->aligner);
  free(h);
}

/* Sum of M-wavefront widths over all computed scores (SURVEY.md 8(d) "C").
 * Only meaningful in the non-modular (scope=full, memory high) layout. */
static int64_t ref_count_cells(wavefront_aligner_t* a) {
  wavefront_components_t* c = &a->wf_components;
  if (c->memory_modular || a->bialigner != NULL) return -1;
  const int last = a->align_status.score;
  int64_t cells = 0;
  for (int s = 0; s <= last && s < c->num_wavefronts; ++s) {
    wavefront_t* m = c->mwavefronts[s];
    if (m != NULL && m->hi >= m->lo) cells += (int64_t)(m->hi - m->lo + 1);
  }
  return cells;
}

/* wildcard_fun_args / wildcard_match_fun of pywfa/align.pyx:297-304 */
typedef struct { const char* pattern; const char* query; char wildcard; } ref_wildcard_args_t;
static int ref_wildcard_match(int pattern_pos, int query_pos, void* argsptr) {
  const ref_wildcard_args_t* a = (const ref_wildcard_args_t*)argsptr;
  return a->pattern[pattern_pos] == a->wildcard || a->query[query_pos] == a->wildcard ||
         a->pattern[pattern_pos] == a->query[query_pos];
}

/*
 * Align one pair.  ops receives the raw operation characters
 * cigar->operations[begin_offset, end_offset) (no terminator); returns the
 * number of operations, or -(needed) if ops_cap is too small.
 */
int ref_align(void* handle, const char* pattern, int plen, const char* text, int tlen,
              int32_t* score, int32_t* status, char* ops, int ops_cap, int64_t* cells) {
  ref_handle_t* h = (ref_handle_t*)handle;
  if (h->cfg.wildcard) {
    /* pywfa/align.pyx:438-442: the wildcard goes through wavefront_align_lambda */
    ref_wildcard_args_t args = {pattern, text, (char)h->cfg.wildcard};
    wavefront_align_lambda(h->aligner, ref_wildcard_match, &args, plen, tlen);
  } else {
    wavefront_align(h->aligner, pattern, plen, text, tlen);
  }
  cigar_t* cg = h->aligner->cigar;
  if (score) *score = cg->score;
  if (status) *status = h->aligner->align_status.status;
  if (cells) *cells = ref_count_cells(h->aligner);
  int n = cg->end_offset - cg->begin_offset;
  if (n < 0) n = 0;
  if (ops) {
    if (n > ops_cap) return -n;
    memcpy(ops, cg->operations + cg->begin_offset, (size_t)n);
  }
  return n;
}

/*
 * Batch driver with the same input layout as wfagpu_align_batch.  ops_off has
 * n+1 entries; operations of pair i are ops[ops_off[i], ops_off[i+1]).
 * Returns 0, or -1 if ops_cap was exhausted (outputs valid up to that pair).
 */
int ref_align_batch(void* handle, const uint8_t* seq,
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
    int r = ref_align(handle, (const char*)seq + p_off[i], p_len[i],
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
