/* pack.h -- host-side batch layout + 2-bit packing (internal, C++ linkage). */
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace wfagpu {

struct PairMetaHost {   /* must match wfagpu::PairMeta (wfa_core.cuh) */
  int64_t woff;
  int32_t plen, tlen;
};

/* Pack one sequence (ASCII, any case) into 2-bit words; false if a non-ACGT byte was seen. */
bool pack_sequence(const uint8_t* s, int len, uint32_t* out);

/* Word offsets of every pair (pattern words, then text words); returns the total word count.
 * bases_per_word: 16 (2-bit codes) or 4 (bytes). */
int64_t layout_pairs(const int32_t* p_len, const int32_t* t_len, int64_t n, PairMetaHost* meta,
                     int32_t* max_plen, int32_t* max_tlen, int bases_per_word = 16);

/* Byte mode (non-ACGT input / wildcard): copy all pairs as upper-cased bytes, 4 per word. */
void pack_pairs_bytes(const uint8_t* seq, const int64_t* p_off, const int64_t* t_off,
                      const PairMetaHost* meta, int64_t n, uint32_t* words, int64_t seq_bytes_hint);

/* Pack all pairs (multi-threaded).  Returns -1, or the index of the first pair holding a
 * byte outside ACGT/acgt; `bad` (optional) receives every such pair, ascending. */
int64_t pack_pairs(const uint8_t* seq, const int64_t* p_off, const int64_t* t_off,
                   const PairMetaHost* meta, int64_t n, uint32_t* words, int64_t seq_bytes_hint,
                   std::vector<int64_t>* bad = nullptr);

/* Byte-pack the pairs listed in `ids` into a side buffer (layout: pattern words, text words, 4 bases
 * per word, pairs back to back) and point their metadata at it: meta[i].woff = ~offset. */
int64_t layout_side_pairs(const std::vector<int64_t>& ids, PairMetaHost* meta);
void pack_side_pairs(const uint8_t* seq, const int64_t* p_off, const int64_t* t_off, const PairMetaHost* meta,
                     const std::vector<int64_t>& ids, uint32_t* words2);

void parallel_copy(void* dst, const void* src, size_t bytes);
int pack_threads(int64_t n_items, int64_t bytes);

}  // namespace wfagpu
