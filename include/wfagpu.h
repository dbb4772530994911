/*
 * wfagpu.h -- C ABI of the B200-native batched wavefront aligner.
 *
 * This is the drop-in boundary for the pywfa hot path: the entry points below
 * replace what pywfa/align.pyx binds from WFA2-lib through pywfa/WFA_wrap.pxd.
 * Every declaration cites the reference interface it stands in for
 * (paths relative to the reference checkout; W/ = pywfa/WFA2_lib/).
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  All entry points
 * return 0 (WFAGPU_OK) or a negative WFAGPU_E* code -- never exit().
 */
#ifndef WFAGPU_H_
#define WFAGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (values are ABI) -------------------------------------------- */
/* distance_metric_t, W/wavefront/wavefront_penalties.h:41-47.  The gap-affine
 * metrics have the fast tiers; gap-linear, edit (levenshtein) and indel are
 * M-only recurrences (W/wavefront/wavefront_compute_linear.c:44-75,
 * wavefront_compute_edit.c:43-97) and run on the scalar tiers.  For LINEAR the
 * indel penalty travels in gap_extension1, as in pywfa (align.pyx:352-355). */
#define WFAGPU_DISTANCE_AFFINE    0
#define WFAGPU_DISTANCE_AFFINE2P  1
#define WFAGPU_DISTANCE_LINEAR    2
#define WFAGPU_DISTANCE_EDIT      3
#define WFAGPU_DISTANCE_INDEL     4
/* alignment_scope_t, W/wavefront/wavefront_attributes.h:50-53 */
#define WFAGPU_SCOPE_SCORE        0
#define WFAGPU_SCOPE_FULL         1
/* alignment_span_t, W/wavefront/wavefront_attributes.h:54-57 */
#define WFAGPU_SPAN_END2END       0
#define WFAGPU_SPAN_ENDSFREE      1
/* wf_heuristic_strategy, W/wavefront/wavefront_heuristic.h (the two strategies
 * reachable from pywfa/align.pyx:401-413). */
#define WFAGPU_HEURISTIC_NONE     0
#define WFAGPU_HEURISTIC_ADAPTIVE 1
#define WFAGPU_HEURISTIC_XDROP    2

/* Per-alignment status, identical to W/wavefront/wfa.h:46-51. */
#define WFAGPU_STATUS_COMPLETED      0
#define WFAGPU_STATUS_PARTIAL        1
#define WFAGPU_STATUS_MAX_STEPS   -100
#define WFAGPU_STATUS_OOM         -200

/* Library error codes (return values). */
#define WFAGPU_OK              0
#define WFAGPU_EINVAL         -1   /* bad argument / configuration the reference would exit(1) on */
#define WFAGPU_ECUDA          -2   /* CUDA runtime error (see wfagpu_last_error) */
#define WFAGPU_ENOMEM         -3   /* host or device allocation failed */
#define WFAGPU_ENODEVICE      -4   /* no CUDA device: there is no CPU fallback */
#define WFAGPU_EUNSUPPORTED   -5   /* input outside the accelerated path (e.g. non-ACGT bases) */

/* SAM operation codes used in CIGAR runs (pywfa/align.pyx:11-14 `codes` LUT). */
#define WFAGPU_OP_M 0
#define WFAGPU_OP_I 1
#define WFAGPU_OP_D 2
#define WFAGPU_OP_X 8

/*
 * Alignment configuration: the POD that replaces wavefront_aligner_attr_t
 * (W/wavefront/wavefront_attributes.h:108-128) for this path.  Field meaning
 * and defaults follow the kwargs of pywfa/align.pyx:309-334.
 */
typedef struct wfagpu_config {
  int32_t distance;                /* WFAGPU_DISTANCE_*                      */
  int32_t scope;                   /* WFAGPU_SCOPE_*                         */
  int32_t span;                    /* WFAGPU_SPAN_*                          */
  int32_t pattern_begin_free;      /* alignment_form_t, attributes.h:58-68   */
  int32_t pattern_end_free;
  int32_t text_begin_free;
  int32_t text_end_free;
  int32_t heuristic;               /* WFAGPU_HEURISTIC_*                     */
  int32_t min_wavefront_length;    /* wf-adaptive, heuristic.h:50-69         */
  int32_t max_distance_threshold;
  int32_t steps_between_cutoffs;
  int32_t xdrop;
  int32_t match;                   /* <= 0; user penalties, NOT normalised   */
  int32_t mismatch;                /* (EDIT / INDEL ignore all six penalties) */
  int32_t gap_opening1;            /* unused by LINEAR                        */
  int32_t gap_extension1;          /* LINEAR: the indel penalty               */
  int32_t gap_opening2;
  int32_t gap_extension2;
  int32_t max_steps;               /* <= 0 means unlimited (align.pyx:415)   */
  int32_t wildcard;                /* 0: none; else the (upper-case) byte that matches every base --
                                      pywfa's wildcard= kwarg, wavefront_align_lambda with
                                      wildcard_match_fun (pywfa/align.pyx:297-304,438-442).  Pairs
                                      whose bytes are among ACGTNRYK and this byte stay on the fast
                                      kernels (4-bit symbol codes); any other byte: scalar kernels   */
} wfagpu_config_t;

typedef struct wfagpu_ctx wfagpu_ctx;       /* one per CUDA device            */
typedef struct wfagpu_batch wfagpu_batch;   /* a packed, device-resident batch */

/* Fill *cfg with pywfa's constructor defaults (pywfa/align.pyx:309-334). */
void wfagpu_config_default(wfagpu_config_t* cfg);

/* Validate what wavefront_penalties_set_linear/affine/affine2p (W/wavefront/
 * wavefront_penalties.c:62-173) and wavefront_align_presets__checks
 * (W/wavefront/wavefront_align.c:48-103, incl. "drop heuristics with edit /
 * indel") exit(1) on.  plen/tlen < 0 skips the
 * per-pair ends-free bound check.  Returns WFAGPU_OK or WFAGPU_EINVAL (message
 * in err). */
int wfagpu_config_check(const wfagpu_config_t* cfg, int64_t plen, int64_t tlen,
                        char* err, size_t errlen);

int wfagpu_device_count(void);

/* Replaces wavefront_aligner_new (W/wavefront/wavefront_aligner.c:421-463;
 * bound at pywfa/WFA_wrap.pxd:1215, called at pywfa/align.pyx:419). */
int wfagpu_create(wfagpu_ctx** out, int device, char* err, size_t errlen);
/* Replaces wavefront_aligner_delete (pywfa/align.pyx:883). */
void wfagpu_destroy(wfagpu_ctx* ctx);
const char* wfagpu_last_error(const wfagpu_ctx* ctx);
const char* wfagpu_strerror(int code);

/*
 * The batched hot path.  Replaces n calls of wavefront_align
 * (W/wavefront/wavefront_align.c:212-241; bound at pywfa/WFA_wrap.pxd:1281,
 * called at pywfa/align.pyx:439) plus the result read-back pywfa does from
 * aligner->cigar / ->align_status (pywfa/align.pyx:443,463,731-833).
 *
 * Inputs: `seq` holds ASCII bases (any case); pair i is
 * pattern = seq[p_off[i] .. +p_len[i]), text = seq[t_off[i] .. +t_len[i]).
 * `seq` may be plain host memory, pinned host memory (wfagpu_host_alloc /
 * wfagpu_host_register: the copy engine then reads it directly and no host
 * core touches a base) or memory of the context's device (packed in place,
 * no upload).  The offset / length arrays are host memory, pinned or not.
 * Outputs (host, caller-allocated, n entries each; any may be NULL):
 *   score[i]   cigar->score            (align.pyx:443)
 *   status[i]  align_status.status     (align.pyx:463)
 *   locs[4i..] pattern_start, pattern_end, text_start, text_end
 *                                      (align.pyx:788-833 `locations`)
 * CIGAR (scope=full): run-length encoded like `cigartuples`
 * (align.pyx:759-786): *cig_runs points at library-owned host memory holding
 * (length<<4 | op) words; pair i owns runs [cig_off[i], cig_off[i+1]).
 * cig_off must have n+1 entries.  The memory stays valid until the next
 * align call on this ctx or wfagpu_destroy.
 *
 * Result arrays in pinned memory are written by the copy engine directly.
 *
 * The bases are uploaded as they are and packed (2 bits per base) and
 * bucketed by length on the device; this replaces the per-alignment copy of
 * wavefront_sequences_init_ascii (W/wavefront/wavefront_sequences.c:141-170).
 * Large batches are processed in chunks: the calling thread stages chunk c+1
 * while a second host thread drives the kernels of chunk c and a third hands
 * finished downloads to the caller.  One call at a time per context (calls
 * from several threads are serialised inside the library).
 */
int wfagpu_align_batch(wfagpu_ctx* ctx, const wfagpu_config_t* cfg,
                       const uint8_t* seq,
                       const int64_t* p_off, const int32_t* p_len,
                       const int64_t* t_off, const int32_t* t_len,
                       int64_t n,
                       int32_t* score, int32_t* status, int32_t* locs,
                       int64_t* cig_off, const uint32_t** cig_runs);

/*
 * One pair: the literal replacement of a single wavefront_align(aligner, pattern,
 * plen, text, tlen) call (W/wavefront/wavefront_align.c:212-241; bound at
 * pywfa/WFA_wrap.pxd:1281, called at pywfa/align.pyx:439) followed by the reads
 * of aligner->cigar / ->align_status (pywfa/align.pyx:443,463,731-833).  pattern /
 * text are raw ASCII (any case, no terminator needed).  locs[4] as above;
 * *cig_runs points at *n_runs run words (length << 4 | op) in library-owned
 * memory, valid until the next call on this context.  Short gap-affine pairs
 * without cut-offs run through one kernel launch on a mapped mailbox (no staging
 * copies); anything else is a batch of one.  A loop over single pairs is the
 * slowest way to use a GPU -- prefer wfagpu_align_batch -- but it is what
 * pywfa's a(text, pattern) does, so it is kept as cheap as a launch allows.
 */
int wfagpu_align_pair(wfagpu_ctx* ctx, const wfagpu_config_t* cfg,
                      const char* pattern, int32_t plen, const char* text, int32_t tlen,
                      int32_t* score, int32_t* status, int32_t* locs,
                      const uint32_t** cig_runs, int32_t* n_runs);

/*
 * Staged form of the same path, for callers that keep batches resident in HBM
 * (bench.py's device-resident `value`, pipelined streaming).
 *   prepare : H2D of the raw bases, 2-bit packing and length bucketing on the device
 *   run     : kernels only, on `stream` (a cudaStream_t; NULL = ctx stream)
 *   fetch   : D2H of results into the caller's arrays (same layout as above)
 */
int wfagpu_batch_prepare(wfagpu_ctx* ctx, const wfagpu_config_t* cfg,
                         const uint8_t* seq,
                         const int64_t* p_off, const int32_t* p_len,
                         const int64_t* t_off, const int32_t* t_len,
                         int64_t n, wfagpu_batch** out);
int wfagpu_batch_run(wfagpu_ctx* ctx, wfagpu_batch* b, void* stream);
int wfagpu_batch_fetch(wfagpu_ctx* ctx, wfagpu_batch* b,
                       int32_t* score, int32_t* status, int32_t* locs,
                       int64_t* cig_off, const uint32_t** cig_runs);
void wfagpu_batch_free(wfagpu_ctx* ctx, wfagpu_batch* b);

/*
 * Pinned host memory for callers.  Batches built in it (and result arrays
 * allocated from it) move by DMA without any host-side copy; this is what
 * the reference's "inputs are copied" step (W/wavefront/wavefront_sequences.c:97)
 * becomes when the consumer is a GPU.  wfagpu_host_register pins memory the
 * caller already owns (e.g. a numpy array) in place.
 */
void* wfagpu_host_alloc(size_t bytes);                 /* NULL on failure */
void wfagpu_host_free(void* p);
int wfagpu_host_register(void* p, size_t bytes);
int wfagpu_host_unregister(void* p);

/* Counters of the last run on this batch (for bench.py / roofline maths). */
typedef struct wfagpu_batch_stats {
  int64_t n_pairs;
  int64_t kernel_launches;     /* launches of OUR kernels in the last run (the first run counts the packing kernels) */
  int64_t packed_bytes;        /* 2-bit input bytes resident in HBM           */
  int64_t h2d_bytes;           /* bytes copied host->device by the staging    */
  int64_t d2h_bytes;           /* bytes copied device->host by the last fetch */
  int64_t cells;               /* sum of computed wavefront cells (device counter) */
  int64_t history_bytes;       /* scope=full backtrace history written to HBM */
  int64_t retried_pairs;       /* pairs re-run on a larger tier               */
} wfagpu_batch_stats_t;
int wfagpu_batch_get_stats(const wfagpu_batch* b, wfagpu_batch_stats_t* out);
/* Kernel launches issued by the last wfagpu_align_batch call on this context. */
int64_t wfagpu_last_launches(const wfagpu_ctx* ctx);

/*
 * Destination of the CIGAR runs of wfagpu_align_batch.  By default *cig_runs
 * points at library-owned pinned memory (the analogue of aligner->cigar, which
 * WFA2-lib owns and overwrites on the next call, pywfa/align.pyx:737-756).  A
 * caller that gathers the results of several contexts into one array (one
 * context per GPU) hands each context its slice of that array
 * instead (pinned memory: written by DMA); a call whose runs do not fit fails
 * with WFAGPU_ENOMEM.  buf = NULL restores the default.
 */
int wfagpu_set_run_buffer(wfagpu_ctx* ctx, uint32_t* buf, int64_t capacity_words);

#ifdef __cplusplus
}
#endif
#endif /* WFAGPU_H_ */
