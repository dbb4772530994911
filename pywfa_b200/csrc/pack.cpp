/*
 * pack.cpp -- host side of the batch layout: ASCII bases -> 2-bit words in pinned memory.
 *
 * Replaces wavefront_sequences_init_ascii (W/wavefront/wavefront_sequences.c:141-170), which
 * copies both sequences into one sentinel-padded byte buffer per alignment.  Here every pair is
 * packed once, 16 bases per 32-bit word (base j of a word in bits 2j..2j+1, code = (c>>1)&3 so
 * upper and lower case map alike: A=0 C=1 T=2 G=3), pattern words first, text words after.
 * Bytes other than ACGT/acgt are reported (the 2-bit path cannot represent them).
 */
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "pack.h"

namespace wfagpu {

static inline uint32_t code_of(uint8_t c) { return (c >> 1) & 3u; }
static inline bool is_acgt(uint8_t c) {
  const uint8_t u = c & 0xDF;
  return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

/* returns true when every byte was A/C/G/T (any case) */
bool pack_sequence(const uint8_t* s, int len, uint32_t* out) {
  int i = 0, w = 0;
  bool ok = true;
#if defined(__AVX2__) && defined(__BMI2__)
  const __m256i up = _mm256_set1_epi8((char)0xDF);
  const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C');
  const __m256i cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
  __m256i good = _mm256_set1_epi8((char)0xFF);
  const uint64_t m = 0x0606060606060606ull;
  for (; i + 32 <= len; i += 32, w += 2) {
    const __m256i x = _mm256_loadu_si256((const __m256i*)(s + i));
    const __m256i u = _mm256_and_si256(x, up);
    const __m256i e = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                      _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
    good = _mm256_and_si256(good, e);
    const uint64_t q0 = (uint64_t)_mm256_extract_epi64(x, 0), q1 = (uint64_t)_mm256_extract_epi64(x, 1);
    const uint64_t q2 = (uint64_t)_mm256_extract_epi64(x, 2), q3 = (uint64_t)_mm256_extract_epi64(x, 3);
    out[w] = (uint32_t)(_pext_u64(q0, m) | (_pext_u64(q1, m) << 16));
    out[w + 1] = (uint32_t)(_pext_u64(q2, m) | (_pext_u64(q3, m) << 16));
  }
  ok = (_mm256_movemask_epi8(good) == -1);
#endif
  /* 8 bases at a time: SWAR validity test + one pext; `cur` collects up to two 16-bit halves */
  uint32_t cur = 0;
  int half = 0;
#if defined(__BMI2__)
  for (; i + 8 <= len; i += 8) {
    uint64_t q;
    memcpy(&q, s + i, 8);
    const uint64_t u = q & 0xDFDFDFDFDFDFDFDFull;
    /* per byte: zero iff equal to the letter; OR the four "is non-zero" masks together */
    auto nz = [](uint64_t v) { return ((v & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | v; };
    const uint64_t miss = nz(u ^ 0x4141414141414141ull) & nz(u ^ 0x4343434343434343ull) &
                          nz(u ^ 0x4747474747474747ull) & nz(u ^ 0x5454545454545454ull);
    ok &= (miss & 0x8080808080808080ull) == 0;
    cur |= (uint32_t)_pext_u64(q, 0x0606060606060606ull) << (16 * half);
    if (++half == 2) { out[w++] = cur; cur = 0; half = 0; }
  }
#endif
  uint32_t acc = cur;
  int nb = 16 * half;       /* bits already in acc */
#if defined(__BMI2__)
  if (i < len && len >= 8) {
    /* last 1..7 bases: re-read the final 8 bytes of the sequence and shift the tail down */
    const int rem = len - i;
    uint64_t q;
    memcpy(&q, s + len - 8, 8);
    q >>= 8 * (8 - rem);
    const uint64_t u = q & 0xDFDFDFDFDFDFDFDFull;
    auto nz = [](uint64_t v) { return ((v & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | v; };
    const uint64_t miss = nz(u ^ 0x4141414141414141ull) & nz(u ^ 0x4343434343434343ull) &
                          nz(u ^ 0x4747474747474747ull) & nz(u ^ 0x5454545454545454ull);
    const uint64_t live = (~0ull) >> (8 * (8 - rem));
    ok &= (miss & live & 0x8080808080808080ull) == 0;
    acc |= (uint32_t)_pext_u64(q, 0x0606060606060606ull) << nb;
    nb += 2 * rem;           /* nb counts bits here */
    out[w++] = acc;
    return ok;
  }
#endif
  nb /= 2;                   /* bases */
  for (; i < len; ++i) {
    const uint8_t c = s[i];
    ok &= is_acgt(c);
    acc |= code_of(c) << (2 * nb);
    if (++nb == 16) { out[w++] = acc; acc = 0; nb = 0; }
  }
  if (nb) out[w++] = acc;
  return ok;
}

int pack_threads(int64_t n_items, int64_t bytes) {
  static int hw = [] {
    const char* e = getenv("WFAGPU_THREADS");
    int t = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(t, 256));
  }();
  if (bytes < (1 << 20) || n_items < 64) return 1;
  return (int)std::min<int64_t>(hw, std::max<int64_t>(1, bytes >> 19));
}

/* A small persistent pool: chunked batches call parallel_for dozens of times per second, so
 * threads are created once per process instead of once per call. */
namespace {
class Pool {
 public:
  static Pool& get() { static Pool p; return p; }
  void run(int nthreads, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> guard(call_mu_);       /* one parallel region at a time */
    ensure(nthreads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn; want_ = nthreads - 1; next_ = 0; done_ = 0; ++epoch_;
    }
    cv_.notify_all();
    fn(nthreads - 1);                                  /* the caller works too */
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return done_ == want_; });
    fn_ = nullptr;
  }

 private:
  void ensure(int n) {
    while ((int)th_.size() < n) th_.emplace_back([this] { loop(); });
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      int idx;
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && next_ < want_); });
        if (stop_) return;
        idx = next_++;
        if (next_ >= want_) seen = epoch_;
        fn = fn_;
      }
      (*fn)(idx);
      std::lock_guard<std::mutex> lk(mu_);
      if (++done_ == want_) cv_done_.notify_all();
    }
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  std::vector<std::thread> th_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, cv_done_;
  const std::function<void(int)>* fn_ = nullptr;
  int want_ = 0, next_ = 0, done_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};
}  // namespace

template <class F>
static void parallel_for(int nthreads, int64_t n, F&& fn) {
  if (nthreads <= 1) { fn(0, (int64_t)0, n); return; }
  const std::function<void(int)> body = [&](int t) {
    const int64_t a = n * t / nthreads, b = n * (t + 1) / nthreads;
    fn(t, a, b);
  };
  Pool::get().run(nthreads, body);
}

int64_t layout_pairs(const int32_t* p_len, const int32_t* t_len, int64_t n, PairMetaHost* meta,
                     int32_t* max_plen, int32_t* max_tlen, int bases_per_word) {
  const int64_t bpw = bases_per_word;
  const int nt = pack_threads(n, n * 16);
  std::vector<int64_t> part(nt + 1, 0);
  std::vector<int32_t> mp(nt, 0), mt(nt, 0);
  parallel_for(nt, n, [&](int t, int64_t a, int64_t b) {
    int64_t s = 0; int32_t xp = 0, xt = 0;
    for (int64_t i = a; i < b; ++i) {
      s += ((int64_t)p_len[i] + bpw - 1) / bpw + ((int64_t)t_len[i] + bpw - 1) / bpw;
      xp = std::max(xp, p_len[i]); xt = std::max(xt, t_len[i]);
    }
    part[t + 1] = s; mp[t] = xp; mt[t] = xt;
  });
  for (int t = 0; t < nt; ++t) part[t + 1] += part[t];
  parallel_for(nt, n, [&](int t, int64_t a, int64_t b) {
    int64_t off = part[t];
    for (int64_t i = a; i < b; ++i) {
      meta[i].woff = off; meta[i].plen = p_len[i]; meta[i].tlen = t_len[i];
      off += ((int64_t)p_len[i] + bpw - 1) / bpw + ((int64_t)t_len[i] + bpw - 1) / bpw;
    }
  });
  *max_plen = *std::max_element(mp.begin(), mp.end());
  *max_tlen = *std::max_element(mt.begin(), mt.end());
  return part[nt];
}

int64_t pack_pairs(const uint8_t* seq, const int64_t* p_off, const int64_t* t_off,
                   const PairMetaHost* meta, int64_t n, uint32_t* words, int64_t seq_bytes_hint,
                   std::vector<int64_t>* bad) {
  const int nt = pack_threads(n, seq_bytes_hint);
  std::atomic<int64_t> first_bad(INT64_MAX);
  std::vector<std::vector<int64_t>> bad_t(nt);
  parallel_for(nt, n, [&](int tid, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      uint32_t* w = words + meta[i].woff;
      const int pw = (meta[i].plen + 15) / 16;
      bool ok = pack_sequence(seq + p_off[i], meta[i].plen, w);
      ok &= pack_sequence(seq + t_off[i], meta[i].tlen, w + pw);
      if (!ok) {
        int64_t cur = first_bad.load();
        while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
        if (bad) bad_t[tid].push_back(i);
      }
    }
  });
  if (bad) {
    bad->clear();
    for (auto& v : bad_t) bad->insert(bad->end(), v.begin(), v.end());     /* threads own ascending ranges */
    std::sort(bad->begin(), bad->end());
  }
  const int64_t fb = first_bad.load();
  return fb == INT64_MAX ? -1 : fb;
}

static void put_bytes(const uint8_t* s, int len, uint32_t* out) {
  /* upper-case a-z like pywfa does before the C call (pywfa/align.pyx:431-435); byte j of a word in bits 8j.. */
  uint8_t* o = reinterpret_cast<uint8_t*>(out);
  for (int i = 0; i < len; ++i) { const uint8_t c = s[i]; o[i] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
  for (int i = len; i < ((len + 3) & ~3); ++i) o[i] = 0;
}

int64_t layout_side_pairs(const std::vector<int64_t>& ids, PairMetaHost* meta) {
  int64_t off = 0;
  for (int64_t i : ids) {
    meta[i].woff = ~off;
    off += ((int64_t)meta[i].plen + 3) / 4 + ((int64_t)meta[i].tlen + 3) / 4;
  }
  return off;
}

void pack_side_pairs(const uint8_t* seq, const int64_t* p_off, const int64_t* t_off, const PairMetaHost* meta,
                     const std::vector<int64_t>& ids, uint32_t* words2) {
  for (int64_t i : ids) {
    uint32_t* w = words2 + ~meta[i].woff;
    put_bytes(seq + p_off[i], meta[i].plen, w);
    put_bytes(seq + t_off[i], meta[i].tlen, w + (meta[i].plen + 3) / 4);
  }
}

void pack_pairs_bytes(const uint8_t* seq, const int64_t* p_off, const int64_t* t_off,
                      const PairMetaHost* meta, int64_t n, uint32_t* words, int64_t seq_bytes_hint) {
  const int nt = pack_threads(n, seq_bytes_hint);
  parallel_for(nt, n, [&](int, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      uint32_t* w = words + meta[i].woff;
      put_bytes(seq + p_off[i], meta[i].plen, w);
      put_bytes(seq + t_off[i], meta[i].tlen, w + (meta[i].plen + 3) / 4);
    }
  });
}

void parallel_copy(void* dst, const void* src, size_t bytes) {
  const int nt = pack_threads(1 << 20, (int64_t)bytes);
  if (nt <= 1) { memcpy(dst, src, bytes); return; }
  parallel_for(nt, (int64_t)bytes, [&](int, int64_t a, int64_t b) {
    memcpy((char*)dst + a, (const char*)src + a, (size_t)(b - a));
  });
}

}  // namespace wfagpu
