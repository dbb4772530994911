/*
 * pack.cpp -- host side of the batch staging.
 *
 * The bases themselves are packed on the device (wfa_pack.cu); the host only (1) reduces the
 * caller's offset / length arrays of a chunk to the few numbers the planner needs (scan_pairs:
 * byte range, longest sequences, word count, length-bucket histogram), (2) moves bytes from
 * pageable memory into pinned staging (parallel_copy, gather_pairs) when the caller did not hand
 * over pinned or device memory.  All of it runs on a small persistent thread pool whose size is
 * capped per process (host_threads).
 *
 * pack_sequence is the host statement of the 2-bit layout (16 bases per 32-bit word, base j in bits
 * 2j..2j+1, code = (c>>1)&3 so upper and lower case map alike: A=0 C=1 T=2 G=3); the CPU test-suite
 * feeds the host-compiled kernel sources with it and the GPU tests pin the device packer to it.
 * Together they replace wavefront_sequences_init_ascii (W/wavefront/wavefront_sequences.c:141-170).
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

#include <sched.h>

int host_threads() {
  static int hw = [] {
    /* WFAGPU_THREADS wins; else the cores this process may run on, shared evenly between the ranks
     * torchrun placed on this node (8 ranks must not each start one thread per core) */
    if (const char* e = getenv("WFAGPU_THREADS")) return std::max(1, std::min(atoi(e), 256));
    int t = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) t = std::min(t > 0 ? t : 1 << 20, CPU_COUNT(&set));
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    return std::max(1, std::min(t / ranks, 64));
  }();
  return hw;
}

int pack_threads(int64_t n_items, int64_t bytes) {
  if (bytes < (1 << 20) || n_items < 64) return 1;
  return (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, bytes >> 19));
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


/* fixed length classes (max(plen, tlen) <= limit): the capacity steps of the alignment tiers */
static const int32_t kClassLimit[MAX_LEN_CLASSES] = {192, 320, 512, 1024, 2048, 4096, 12000, INT32_MAX};
int32_t length_class_limit(int c) { return kClassLimit[c]; }

void scan_pairs(const int64_t* p_off, const int32_t* p_len, const int64_t* t_off, const int32_t* t_len, int64_t n,
                int bases_per_word, PairScan* out) {
  const int64_t bpw = bases_per_word;
  const int nt = pack_threads(n, n * 24);
  std::vector<PairScan> part(nt);
  parallel_for(nt, n, [&](int t, int64_t a, int64_t b) {
    PairScan r;
    for (int64_t i = a; i < b; ++i) {
      const int32_t pl = p_len[i], tl = t_len[i];
      if ((pl | tl) < 0) { if (r.first_negative < 0) r.first_negative = i; continue; }
      if (t_off[i] != p_off[i] + pl || (i + 1 < n && p_off[i + 1] != t_off[i] + tl)) r.back_to_back = false;
      r.seq_bytes += (int64_t)pl + tl;
      r.total_words += ((int64_t)pl + bpw - 1) / bpw + ((int64_t)tl + bpw - 1) / bpw;
      if (pl) { r.lo = std::min(r.lo, p_off[i]); r.hi = std::max(r.hi, p_off[i] + pl); }
      if (tl) { r.lo = std::min(r.lo, t_off[i]); r.hi = std::max(r.hi, t_off[i] + tl); }
      r.maxp = std::max(r.maxp, pl); r.maxt = std::max(r.maxt, tl);
      r.minp = std::min(r.minp, pl); r.mint = std::min(r.mint, tl);
      const int32_t L = std::max(pl, tl);
      int c = 0;
      while (L > kClassLimit[c]) ++c;
      r.cls_n[c]++; r.cls_maxp[c] = std::max(r.cls_maxp[c], pl); r.cls_maxt[c] = std::max(r.cls_maxt[c], tl);
    }
    part[t] = r;
  });
  PairScan r;
  for (const PairScan& q : part) {
    if (q.first_negative >= 0 && (r.first_negative < 0 || q.first_negative < r.first_negative)) r.first_negative = q.first_negative;
    r.back_to_back = r.back_to_back && q.back_to_back;
    r.seq_bytes += q.seq_bytes; r.total_words += q.total_words;
    r.lo = std::min(r.lo, q.lo); r.hi = std::max(r.hi, q.hi);
    r.maxp = std::max(r.maxp, q.maxp); r.maxt = std::max(r.maxt, q.maxt);
    r.minp = std::min(r.minp, q.minp); r.mint = std::min(r.mint, q.mint);
    for (int c = 0; c < MAX_LEN_CLASSES; ++c) {
      r.cls_n[c] += q.cls_n[c];
      r.cls_maxp[c] = std::max(r.cls_maxp[c], q.cls_maxp[c]); r.cls_maxt[c] = std::max(r.cls_maxt[c], q.cls_maxt[c]);
    }
  }
  if (r.lo > r.hi) r.lo = r.hi = 0;
  *out = r;
}

void gather_pairs(const uint8_t* seq, const int64_t* p_off, const int32_t* p_len, const int64_t* t_off,
                  const int32_t* t_len, int64_t n, uint8_t* dst, int64_t* new_p_off, int64_t* new_t_off) {
  const int nt = pack_threads(n, n * 64);
  std::vector<int64_t> part(nt + 1, 0);
  parallel_for(nt, n, [&](int t, int64_t a, int64_t b) {
    int64_t s = 0;
    for (int64_t i = a; i < b; ++i) s += (int64_t)p_len[i] + t_len[i];
    part[t + 1] = s;
  });
  for (int t = 0; t < nt; ++t) part[t + 1] += part[t];
  parallel_for(nt, n, [&](int t, int64_t a, int64_t b) {
    int64_t off = part[t];
    for (int64_t i = a; i < b; ++i) {
      new_p_off[i] = off;
      memcpy(dst + off, seq + p_off[i], (size_t)p_len[i]);
      off += p_len[i];
      new_t_off[i] = off;
      memcpy(dst + off, seq + t_off[i], (size_t)t_len[i]);
      off += t_len[i];
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
