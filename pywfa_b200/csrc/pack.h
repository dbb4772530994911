/* pack.h -- host side of the batch staging (internal, C++ linkage): see pack.cpp. */
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace wfagpu {

/* Pack one sequence (ASCII, any case) into 2-bit words; false if a non-ACGT byte was seen.
 * The host statement of the layout the device packer (wfa_pack.cu) produces. */
bool pack_sequence(const uint8_t* s, int len, uint32_t* out);

constexpr int MAX_LEN_CLASSES = 8;
int32_t length_class_limit(int c);      /* class c holds pairs with max(plen, tlen) <= limit (ascending; last = INT32_MAX) */

/* What the planner needs to know about a chunk of pairs, from one parallel pass over the arrays. */
struct PairScan {
  int64_t first_negative = -1;          /* first pair with a negative length, or -1 */
  bool back_to_back = true;             /* pattern_0 text_0 pattern_1 text_1 ... without gaps: offsets follow from the lengths */
  int64_t seq_bytes = 0;                /* sum of plen + tlen */
  int64_t total_words = 0;              /* sum of ceil(plen / bpw) + ceil(tlen / bpw) */
  int64_t lo = INT64_MAX, hi = INT64_MIN;   /* byte range [lo, hi) of the sequence buffer the pairs touch */
  int32_t maxp = 0, maxt = 0, minp = INT32_MAX, mint = INT32_MAX;
  int64_t cls_n[MAX_LEN_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0};
  int32_t cls_maxp[MAX_LEN_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0}, cls_maxt[MAX_LEN_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0};
};
void scan_pairs(const int64_t* p_off, const int32_t* p_len, const int64_t* t_off, const int32_t* t_len, int64_t n,
                int bases_per_word, PairScan* out);

/* Scattered input: copy every pair's bases back to back (pattern, text) into dst and write the
 * offsets they have there. */
void gather_pairs(const uint8_t* seq, const int64_t* p_off, const int32_t* p_len, const int64_t* t_off,
                  const int32_t* t_len, int64_t n, uint8_t* dst, int64_t* new_p_off, int64_t* new_t_off);

void parallel_copy(void* dst, const void* src, size_t bytes);
int host_threads();                     /* worker threads this process may use (WFAGPU_THREADS, affinity, ranks per node) */
int pack_threads(int64_t n_items, int64_t bytes);

}  // namespace wfagpu
