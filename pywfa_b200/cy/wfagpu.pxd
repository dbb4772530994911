# wfagpu.pxd -- Cython declarations of include/wfagpu.h: what a pywfa maintainer adds next to
# pywfa/WFA_wrap.pxd to bind the B200 library instead of (or beside) WFA2-lib.
from libc.stdint cimport int32_t, int64_t, uint8_t, uint32_t

cdef extern from "wfagpu.h" nogil:
    ctypedef struct wfagpu_config_t:
        int32_t distance, scope, span
        int32_t pattern_begin_free, pattern_end_free, text_begin_free, text_end_free
        int32_t heuristic, min_wavefront_length, max_distance_threshold, steps_between_cutoffs, xdrop
        int32_t match, mismatch, gap_opening1, gap_extension1, gap_opening2, gap_extension2
        int32_t max_steps, wildcard
    ctypedef struct wfagpu_ctx
    void wfagpu_config_default(wfagpu_config_t* cfg)
    int wfagpu_config_check(const wfagpu_config_t* cfg, int64_t plen, int64_t tlen, char* err, size_t errlen)
    int wfagpu_device_count()
    int wfagpu_create(wfagpu_ctx** out, int device, char* err, size_t errlen)
    void wfagpu_destroy(wfagpu_ctx* ctx)
    const char* wfagpu_last_error(const wfagpu_ctx* ctx)
    int wfagpu_align_batch(wfagpu_ctx* ctx, const wfagpu_config_t* cfg, const uint8_t* seq,
                           const int64_t* p_off, const int32_t* p_len,
                           const int64_t* t_off, const int32_t* t_len, int64_t n,
                           int32_t* score, int32_t* status, int32_t* locs,
                           int64_t* cig_off, const uint32_t** cig_runs)
    int wfagpu_align_pair(wfagpu_ctx* ctx, const wfagpu_config_t* cfg, const char* pattern, int32_t plen,
                          const char* text, int32_t tlen, int32_t* score, int32_t* status, int32_t* locs,
                          const uint32_t** cig_runs, int32_t* n_runs)
    void* wfagpu_host_alloc(size_t nbytes)
    void wfagpu_host_free(void* p)
