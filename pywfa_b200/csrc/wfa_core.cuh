/*
 * wfa_core.cuh -- the wavefront-alignment hot path for one read pair, written for a
 * cooperating thread group (a warp, or a whole CTA) on sm_100a.
 *
 * This is a from-scratch B200 design of the path that pywfa drives in WFA2-lib
 * (reference citations: W/ = pywfa/WFA2_lib/ in the reference checkout):
 *
 *   - the gap-affine / gap-affine-2p recurrence (W/wavefront/wavefront_compute_affine.c:44-86,
 *     wavefront_compute_affine2p.c:45-106) and the exact-match extension
 *     (W/wavefront/wavefront_extend_kernels.c:64-163) are FUSED: the lane that produces
 *     M[s][k] extends it at once, 16 bases per step, by XOR-ing two funnel-shifted 32-bit
 *     words of the 2-bit packed sequences and counting the agreeing low bit pairs;
 *   - scores advance in units of g = gcd of the penalties (every reachable score is a
 *     multiple of g), so x=4,o=6,e=2 needs half the steps and a 5-slot ring;
 *   - the offset wavefronts live in rings of max_score_scope/g slots (M) and e/g+1 slots
 *     (I1/D1/I2/D2) in shared memory -- int16 offsets for short reads, int32 otherwise -- or
 *     in an L2-resident HBM arena for very wide wavefronts; a slot is indexed circularly by
 *     (k & (wcap-1)), and "outside [lo,hi] reads as NULL"
 *     (W/wavefront/wavefront_compute.c:490-567) is a range check at load time;
 *   - trim_ends (W/wavefront/wavefront_compute.c:571-605), end-to-end / ends-free termination
 *     (W/wavefront/wavefront_termination.c:37-162) and the WF-adaptive / X-drop cut-offs
 *     (W/wavefront/wavefront_heuristic.c:257-383,509-567) are min-reductions over the group;
 *   - scope=full spills ONE origin byte per cell to a per-group HBM arena with coalesced
 *     stores: the winner of the reference's (offset<<4 | type) max
 *     (W/wavefront/wavefront_backtrace.c:366-389) and one ext/open bit per gap component.  The
 *     backtrace (W/wavefront/wavefront_backtrace.c:320-529) walks these bytes from the end cell
 *     back to score 0 collecting the edit operations, then replays them forwards re-extending
 *     the matches from the sequences, which emits the run-length encoded CIGAR in order; no
 *     offsets are stored (1 byte per cell instead of the reference's 4 x ncomp).
 *
 * The file is also compiled as plain host C++ with a one-thread group by tests/emu/ (test
 * infrastructure for the CPU-only CI; never part of the product library).
 */
#pragma once
#include <stdint.h>
#include <limits.h>

#ifdef __CUDACC__
#define WFA_DEV __device__ __forceinline__
#else
#define WFA_DEV inline
#endif

namespace wfagpu {

constexpr int OFFNULL = INT32_MIN / 2;      /* W/wavefront/wavefront_offset.h:44 */
constexpr int KNONE = INT_MAX;
constexpr int VEC_MAX_LEN = 12000;           /* packed-halfword tier (wfa_vec.cuh): int16 offsets incl. out-of-matrix I offsets */
constexpr int REG_MAX_LEN = 8000;            /* register tier (wfa_reg.cuh): longest sequence its int16 offsets are sized for */

enum { CM = 0, CI1 = 1, CD1 = 2, CI2 = 3, CD2 = 4 };
/* backtrace_type priorities, W/wavefront/wavefront_backtrace.c:49-59 */
enum { BT_NONE = 0, BT_I1_OPEN = 1, BT_I1_EXT = 2, BT_I2_OPEN = 3, BT_I2_EXT = 4, BT_D1_OPEN = 5,
       BT_D1_EXT = 6, BT_D2_OPEN = 7, BT_D2_EXT = 8, BT_M = 9 };

/* per-pair outcome of one tier */
enum { PAIR_DONE = 0, PAIR_OVERFLOW = 1 };

/* status codes, W/wavefront/wfa.h:46-55 */
constexpr int ST_COMPLETED = 0, ST_PARTIAL = 1, ST_MAX_STEPS = -100, ST_OOM = -200;

/* SAM op codes (pywfa/align.pyx:11-14) */
constexpr uint32_t OP_M = 0, OP_I = 1, OP_D = 2, OP_X = 8;

/* Offsets as stored in the rings / history.  Everything negative is "null": the reference's
 * drifting nulls (NULL+1, NULL+2, ... in I/D cells, compute_affine.c:69-75) never take part in
 * a comparison whose outcome matters, so a narrow type only has to keep negatives negative. */
template <class T> struct OffTraits;
template <> struct OffTraits<int32_t> { static constexpr int kNull = OFFNULL; };
template <> struct OffTraits<int16_t> { static constexpr int kNull = -30000; };
template <class T> WFA_DEV T off_store(int v) {
  const int n = OffTraits<T>::kNull;
  return (T)(v > n ? v : n);
}

struct HistRow {       /* scope=full: where the origin bytes of one score live */
  long long off;       /* first cell of the score in the group's arena */
  int lo;              /* diagonal of that cell */
  int pad;
};

struct PairMeta {      /* 16 bytes, one per pair, in HBM */
  int64_t woff;        /* first 32-bit word of the packed pattern; text words follow it.  Negative: the pair holds
                          a non-ACGT byte, its sequences are BYTES at words2 + ~woff (scalar tiers only) */
  int32_t plen, tlen;
};

struct KParams {
  /* normalised penalties (W/wavefront/wavefront_penalties.c:95-173) */
  int x, o1, e1, o2, e2, match;
  int max_scope;                 /* W/wavefront/wavefront_components.c:81-124 (original units) */
  /* gap-linear / edit / indel (W/wavefront/wavefront_compute_linear.c, wavefront_compute_edit.c): M wavefronts
   * only (the gap sources are the "open" sources of a zero-cost opening); indel has no mismatch source;
   * edit / indel run compute_edit.c's driver (no null steps, range = previous range +-1, positive scores) */
  int m_only, no_mis, edit_like;
  int pos_score;                 /* edit / indel report the distance itself (compute.c:117) */
  int edit_prune;                /* edit, end-to-end: wavefront_compute_edit_exact_prune (compute_edit.c:219-275) */
  /* score unit g = gcd(x, o1+e1, e1[, o2+e2, e2]) and the penalties in that unit */
  int g, dx, doe1, de1, doe2, de2;
  int rm, r1, r2;                /* ring slots: M, I1/D1, I2/D2 (scaled look-back + 1) */
  int mr;                        /* metadata ring entries: power of two >= rm */
  int endsfree;                  /* span */
  int pbf, pef, tbf, tef;
  int heuristic, min_wf_len, max_dist_thr, steps_between, xdrop;
  int max_steps;                 /* INT_MAX = unlimited */
  int byte_mode;                 /* sequences are bytes (4 per word) instead of 2-bit codes: non-ACGT input / wildcard */
  int wildcard;                  /* byte mode: 0 or the byte that matches everything */
  /* tier capacities */
  int wcap;                      /* wavefront width capacity per ring slot: power of two */
  int seq_words_cap;             /* words of smem for both packed sequences (0: read HBM) */
  int vec_seqw;                  /* packed-halfword tier: the sequences are staged as per-base windows */
  int group_bytes;               /* shared memory of one group (multiple of 16) */
  long long hcap;                /* history cells per group */
  int scap;                      /* history score-table entries per group */
  int runcap;                    /* CIGAR run staging words per group */
  /* batch */
  const PairMeta* pairs;
  const uint32_t* words;
  const uint32_t* words2;        /* byte-packed sequences of the pairs whose PairMeta::woff is negative (= ~offset into words2) */
  const int* worklist;           /* pair ids, or nullptr = identity */
  const int* n_work;             /* device pointer to the number of work items */
  int* work_counter;
  int* retry_list; int* retry_count;
  int* done_count; int* ovf_count;   /* pairs this tier tried so far / of those, beyond its capacity (tier_gives_up) */
  int skip_groups;               /* > 0: groups in flight; a tier that overflows most of its first pairs forwards the rest */
  int work_limit;                /* this launch takes work items below this index only (host-side probing of a tier) */
  /* results (SoA) */
  int* score; int* status; int* locs; int* nruns; long long* runs_base;
  /* scope=full scratch */
  uint8_t* hist_code; HistRow* hmeta; uint32_t* runs_stage;
  uint32_t* runs_tmp; unsigned long long* runs_cursor; unsigned long long runs_tmp_cap;
  /* global ring arena for the widest tier (elements per group = gring_elems) */
  int* gring; long long gring_elems;
  unsigned long long* cells_total;
  unsigned long long* dbg;       /* WFA_VEC_TIMING builds only */
  /* register-resident tier (wfa_reg.cuh): per-warp origin-byte arena and edit-operation stack */
  uint8_t* rhist; long long rhist_bytes; int rhrows;
  uint8_t* rops; int ropcap;
  int reg_kbase, reg_clo, reg_chi;   /* window of the launch: diagonal of column 0, columns of the score-0 seeds (reg_window) */
};

/* The register tier's window of 64 * regs diagonals is centred on the score-0 seeds [lo0, hi0]
 * (wavefront_aligner.c:251-310): a constant of the launch, computed once on the host. */
struct RegWindow { int kbase, c_lo, c_hi; };
WFA_DEV RegWindow reg_window(int regs, int endsfree, int match, int pbf, int tbf) {
  const bool ef = endsfree && match == 0;
  const int lo0 = ef ? -pbf : 0, hi0 = ef ? tbf : 0;
  RegWindow w;
  w.kbase = ((lo0 + hi0) >> 1) - 32 * regs;
  w.c_lo = lo0 - w.kbase; w.c_hi = hi0 - w.kbase;
  return w;
}

/* pointers a group works with for the current pair */
template <class OffT>
struct GroupMem {
  const uint32_t* pw; const uint32_t* tw;   /* packed sequences, readable one word past the end */
  int wild;                                 /* -1: 2-bit codes; >= 0: bytes (4 per word), value = wildcard byte or 0 */
  OffT* ring[5];
  int4* meta;                               /* [mr][NC] : lo, hi, ring slot, exists */
  uint8_t* h_code; HistRow* hmeta; uint32_t* runs_stage;
  uint8_t* ops; int opcap;                  /* edit-operation stack of the backtrace */
};

struct PairResult {
  int score, status;
  int locs[4];
  int nruns;
  long long cells;
};

/* Shared memory of one warp of the register tier: [sequence windows][scope=full with the arena in
 * shared memory (REG_SMEM_HIST): packed sequences for the replay | edit-operation stack | origin arena,
 * which the CIGAR run staging re-uses once the backward walk is over]. */
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr bool reg_hist_in_smem(int regs, bool full) { return full && regs == 2; }
struct RegSmem {
  int win_words, pk_words, ops_bytes, hist_bytes;
  WFA_DEV int pk_off() const { return 4 * win_words; }
  WFA_DEV int ops_off() const { return pk_off() + 4 * pk_words; }
  WFA_DEV int hist_off() const { return (ops_off() + ops_bytes + 15) & ~15; }
  WFA_DEV int total() const { return (hist_off() + hist_bytes + 15) & ~15; }
};
WFA_DEV RegSmem reg_smem_layout(int regs, bool full, int seq_words_cap, int ropcap, int hrows) {
  RegSmem L;
  L.win_words = seq_words_cap;
  const bool hs = reg_hist_in_smem(regs, full);
  L.pk_words = hs ? (seq_words_cap + 15) / 16 + 4 : 0;
  L.ops_bytes = hs ? ropcap : 0;
  L.hist_bytes = hs ? hrows * 32 * regs : 0;
  return L;
}

/* ------------------------------------------------------------------------------------ */
#ifdef __CUDACC__
WFA_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, sh); }
WFA_DEV int first_set(uint32_t x) { return __ffs((int)x) - 1; }
WFA_DEV int last_set(uint32_t x) { return 31 - __clz((int)x); }
#else
WFA_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (sh & 31));
}
WFA_DEV int first_set(uint32_t x) { return __builtin_ctz(x); }
WFA_DEV int last_set(uint32_t x) { return 31 - __builtin_clz(x); }
#endif
/* load that bypasses the (non-coherent) L1: data another CTA of the same pair may have written */
template <class T> WFA_DEV T ld_cg(const T* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);
#else
  return *p;
#endif
}
WFA_DEV int imax(int a, int b) { return a > b ? a : b; }
WFA_DEV int imin(int a, int b) { return a < b ? a : b; }

/* 16 bases starting at base index i (base j of a word sits in bits 2j..2j+1) */
WFA_DEV uint32_t fetch16(const uint32_t* w, int i) {
  const int j = i >> 4;
  return funnel_r(w[j], w[j + 1], (i & 15) << 1);
}

/*
 * Exact-match extension of one diagonal (what wavefront_extend_matches_kernel_blockwise,
 * W/wavefront/wavefront_extend_kernels.c:64-88, does 8 bytes at a time with sentinels):
 * here 16 bases per XOR, clamped to the sequence ends instead of sentinels.
 */
/* Byte mode (any ASCII, optional wildcard): 4 bases per word, byte j of a word in bits 8j..8j+7.
 * `wild` = 0 or the wildcard byte: a position matches if the bytes are equal or either one is the
 * wildcard -- wildcard_match_fun of pywfa/align.pyx:302-304, reached in the reference through
 * wavefront_extend_matches_custom (W/wavefront/wavefront_extend_kernels.c:167-203). */
WFA_DEV uint32_t fetch4(const uint32_t* w, int i) {
  const int j = i >> 2;
  return funnel_r(w[j], w[j + 1], (i & 3) << 3);
}
/* 0x80 in every byte of x that is zero (exact, no carries between bytes) */
WFA_DEV uint32_t zero_bytes(uint32_t x) {
  return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}
WFA_DEV int extend_offset(const uint32_t* pw, const uint32_t* tw, int plen, int tlen, int k, int off, int wild = -1) {
  const int v = off - k, h = off;
  const int rem = imin(plen - v, tlen - h);
  int n = 0;
  if (wild >= 0) {
    const uint32_t w4 = (uint32_t)wild * 0x01010101u;
    while (n < rem) {
      const uint32_t a = fetch4(pw, v + n), b = fetch4(tw, h + n);
      uint32_t eq = zero_bytes(a ^ b);
      if (wild) eq |= zero_bytes(a ^ w4) | zero_bytes(b ^ w4);
      const uint32_t ne = ~eq & 0x80808080u;
      if (ne) { n += first_set(ne) >> 3; break; }
      n += 4;
    }
    return off + imin(n, rem);
  }
  while (n < rem) {
    const uint32_t x = fetch16(pw, v + n) ^ fetch16(tw, h + n);
    if (x) { n += first_set(x) >> 1; break; }
    n += 16;
  }
  return off + imin(n, rem);
}

WFA_DEV bool in_bounds(int k, int off, int plen, int tlen) {
  return (uint32_t)off <= (uint32_t)tlen && (uint32_t)(off - k) <= (uint32_t)plen;
}

/* wavefront_compute_classic_score, W/wavefront/wavefront_compute.c:108-120 with
 * WF_SCORE_TO_SW_SCORE (wavefront_penalties.h:73): int32 wrap-around, C truncating division */
WFA_DEV int classic_score(int match, int plen, int tlen, int wf_score, int pos_score = 0) {
  const int swg_match = -match;
  if (pos_score) return wf_score;               /* distance_metric <= edit, compute.c:117 */
  if (swg_match == 0) return -wf_score;
  const int32_t sum = (int32_t)((uint32_t)plen + (uint32_t)tlen);
  const int32_t prod = (int32_t)((uint32_t)swg_match * (uint32_t)sum);
  return (int32_t)((uint32_t)prod - (uint32_t)wf_score) / 2;
}

/*
 * Adaptive tier skipping: once a sample of at least 64 pairs this launch tried has left the tier and three
 * quarters of it exceeded the tier's capacity, the remaining pairs are forwarded to the next tier
 * unprocessed (their partial work would be thrown away).  w = index of the work item just fetched.
 */
WFA_DEV bool tier_gives_up(const KParams& P, int w, bool& gave_up) {
  /* `gave_up` is the group's cached verdict: final once true, re-examined at every 4th work item otherwise */
  if (gave_up || P.skip_groups <= 0 || w < 512 || (w & 3) != 0) return gave_up;
  const int done = ld_cg(P.done_count);               /* sampled: every 8th work item reports (tier_pair_note) */
  if (done >= 64 && 4ll * ld_cg(P.ovf_count) > 3ll * done) gave_up = true;
  return gave_up;
}
/* one pair this tier actually tried left it (one thread of the group); pairs that were forwarded
 * untried -- given up on, or byte-mode pairs on a 2-bit tier -- do not count: they finish at once
 * and would dominate the first samples.  Callers report every 8th pair, so that the two counters are
 * not an atomic hot spot next to the work queue. */
WFA_DEV void tier_pair_note(const KParams& P, bool overflowed) {
#ifdef __CUDA_ARCH__
  if (P.skip_groups > 0) { atomicAdd(P.done_count, 1); if (overflowed) atomicAdd(P.ovf_count, 1); }
#endif
}

/* one source wavefront component as the recurrence reads it */
template <class OffT>
struct Src {
  const OffT* slot;  /* ring slot base (circular: element of diagonal k is slot[k & wmask]) */
  int lo, hi;        /* valid range; lo > hi = null */
};
template <bool CG, class OffT>
WFA_DEV int rd(const Src<OffT>& s, int k, int km) {
  if (!(k >= s.lo && k <= s.hi)) return OFFNULL;
  return CG ? (int)ld_cg(s.slot + km) : (int)s.slot[km];
}

/* ---- CIGAR run emitter (one thread); runs are produced in CIGAR order ----------------- */
struct FwdEmitter {
  uint32_t* stage; int cap; int n; uint32_t op; int len;
  WFA_DEV void init(uint32_t* s, int c) { stage = s; cap = c; n = 0; op = 0xffu; len = 0; }
  WFA_DEV void flush() { if (len > 0) { if (n < cap) stage[n] = ((uint32_t)len << 4) | op; ++n; } len = 0; }
  WFA_DEV void push(uint32_t o, int cnt) {
    if (cnt <= 0) return;
    if (o == op) { len += cnt; return; }
    flush(); op = o; len = cnt;
  }
};

/* edit operations collected by the backward walk (one byte each) */
enum { EOP_X = 0, EOP_I_OPEN = 1, EOP_I_EXT = 2, EOP_D_OPEN = 3, EOP_D_EXT = 4 };

/*
 * Forward replay of the edit operations (ops[nops-1] is the first one) from the score-0 seed of
 * diagonal k: every arrival in the M matrix re-extends the matches from the sequences, exactly
 * the extension the forward pass made there.  Emits what wavefront_backtrace_affine
 * (W/wavefront/wavefront_backtrace.c:320-529) produces right-to-left, in CIGAR order: free
 * prefix, leading matches, operations with their match runs, free suffix.
 */
WFA_DEV void replay_ops(const uint8_t* ops, int nops, int k, int plen, int tlen, const uint32_t* pw, const uint32_t* tw,
                        FwdEmitter& em, int wild = -1) {
  int off = k > 0 ? k : 0;
  em.push(OP_I, k > 0 ? k : 0);            /* free text prefix (ends-free seeds) */
  em.push(OP_D, k < 0 ? -k : 0);           /* free pattern prefix */
  {
    const int e = extend_offset(pw, tw, plen, tlen, k, off, wild);
    em.push(OP_M, e - off); off = e;
  }
  for (int i = nops - 1; i >= 0; --i) {
    const int op = ops[i];
    bool at_m = true;
    if (op == EOP_X) { em.push(OP_X, 1); ++off; }
    else if (op == EOP_I_OPEN || op == EOP_I_EXT) {
      em.push(OP_I, 1); ++k; ++off;
      at_m = !(i > 0 && ops[i - 1] == EOP_I_EXT);
    } else {
      em.push(OP_D, 1); --k;
      at_m = !(i > 0 && ops[i - 1] == EOP_D_EXT);
    }
    if (at_m) {
      const int e = extend_offset(pw, tw, plen, tlen, k, off, wild);
      em.push(OP_M, e - off); off = e;
    }
  }
  em.push(OP_I, tlen - off);                /* free text suffix */
  em.push(OP_D, plen - (off - k));          /* free pattern suffix */
  em.flush();
}

/*
 * Backtrace over the recorded origin bytes (one thread).  Backward walk from the end cell: at an
 * M cell the byte names the winning source (mismatch / open / extend of which gap), inside a gap
 * only the ext/open bit of that component is needed; scores are in units of g.  Returns the
 * number of runs (may exceed em.cap: the CIGAR did not fit) or -1 if the operation stack is full.
 */
WFA_DEV int backtrace_codes(const KParams& P, const uint8_t* h_code, const HistRow* hmeta, int a_score, int a_k,
                            int plen, int tlen, const uint32_t* pw, const uint32_t* tw, uint8_t* ops, int opcap,
                            FwdEmitter& em, int wild = -1) {
  int mt = CM, score = a_score, k = a_k, nops = 0;
  while (score > 0) {
    HistRow hm;
    hm.off = ld_cg(&hmeta[score].off); hm.lo = ld_cg(&hmeta[score].lo);
    const int code = ld_cg(h_code + hm.off + (k - hm.lo));
    int type;
    if (mt == CM) type = code & 15;
    else if (mt == CI1) type = (code & 0x10) ? BT_I1_EXT : BT_I1_OPEN;
    else if (mt == CD1) type = (code & 0x20) ? BT_D1_EXT : BT_D1_OPEN;
    else if (mt == CI2) type = (code & 0x40) ? BT_I2_EXT : BT_I2_OPEN;
    else type = (code & 0x80) ? BT_D2_EXT : BT_D2_OPEN;
    if (type == BT_NONE) break;
    int op;
    switch (type) {
      case BT_M: score -= P.dx; mt = CM; op = EOP_X; break;
      case BT_I1_OPEN: score -= P.doe1; mt = CM; op = EOP_I_OPEN; break;
      case BT_I1_EXT: score -= P.de1; mt = CI1; op = EOP_I_EXT; break;
      case BT_I2_OPEN: score -= P.doe2; mt = CM; op = EOP_I_OPEN; break;
      case BT_I2_EXT: score -= P.de2; mt = CI2; op = EOP_I_EXT; break;
      case BT_D1_OPEN: score -= P.doe1; mt = CM; op = EOP_D_OPEN; break;
      case BT_D1_EXT: score -= P.de1; mt = CD1; op = EOP_D_EXT; break;
      case BT_D2_OPEN: score -= P.doe2; mt = CM; op = EOP_D_OPEN; break;
      default: score -= P.de2; mt = CD2; op = EOP_D_EXT; break;
    }
    if (nops < opcap) ops[nops] = (uint8_t)op;
    ++nops;
    if (op == EOP_I_OPEN || op == EOP_I_EXT) --k;
    else if (op == EOP_D_OPEN || op == EOP_D_EXT) ++k;
  }
  if (nops > opcap) return -1;
  replay_ops(ops, nops, k, plen, tlen, pw, tw, em, wild);
  return em.n;
}

/* `locations` of pywfa (pywfa/align.pyx:788-833) from runs in CIGAR order */
WFA_DEV void locations_from_runs(const uint32_t* runs, int n, int plen, int tlen, int* locs) {
  locs[0] = locs[1] = locs[2] = locs[3] = 0;
  if (n == 0 || plen == 0 || tlen == 0) return;
  int ps = 0, ts = 0;
  for (int i = 0; i < n; ++i) {
    const uint32_t w = runs[i]; const uint32_t op = w & 15; const int ln = (int)(w >> 4);
    if (op == OP_M) break;
    if (op == OP_D) ps += ln; else if (op == OP_X) { ps += ln; ts += ln; } else ts += ln;
  }
  int pe = plen, te = tlen;
  for (int i = n - 1; i >= 0; --i) {
    const uint32_t w = runs[i]; const uint32_t op = w & 15; const int ln = (int)(w >> 4);
    if (op == OP_M) break;
    if (op == OP_D) pe -= ln; else if (op == OP_X) { pe -= ln; te -= ln; } else te -= ln;
  }
  locs[0] = ps; locs[1] = pe; locs[2] = ts; locs[3] = te;
}

WFA_DEV bool term_cell(const KParams& P, int plen, int tlen, int ak, int k, int off) {
  if (P.endsfree) {                            /* termination.c:115-162 */
    const int hh = off, vv = off - k;
    return (hh >= tlen && plen - vv <= P.pef) || (vv >= plen && tlen - hh <= P.tef);
  }
  return k == ak && off >= tlen;               /* termination.c:37-61 */
}

/* ------------------------------------------------------------------------------------ */
/*
 * Align one pair with the thread group `g`.  G provides: rank, size (over the whole group),
 * lrank, lsize (within the CTA: the metadata ring is per CTA), sync() (group), lsync() (CTA),
 * template<int N> allmin(int (&v)[N]) (group-wide, implies sync) and kGrid (group spans CTAs:
 * ring data comes from other SMs and must bypass L1).
 * Returns PAIR_DONE (res filled; for scope=full the runs are in gm.runs_stage, in CIGAR order)
 * or PAIR_OVERFLOW (a tier capacity was exceeded; retry on a larger tier).
 */
template <class G, class OffT, bool TWO_P, bool FULL>
WFA_DEV int align_pair(G& g, const KParams& P, const GroupMem<OffT>& gm, int plen, int tlen, PairResult& res) {
  constexpr int NC = TWO_P ? 5 : 3;
  const int wcap = P.wcap, wmask = P.wcap - 1, mmask = P.mr - 1;
  int4* const meta = gm.meta;
  const int ak = tlen - plen;
  const int wild = gm.wild;                   /* -1: 2-bit packed sequences; else byte mode for this pair */

  int s = 0;                                  /* score in units of g */
  int cm = 0, c1 = 0, c2 = 0;                 /* ring slots of the current score */
  int s_exist = 0;                            /* last score (original units) whose wavefront exists */
  int steps_wait = P.steps_between;           /* W/wavefront/wavefront_heuristic.c:114-121 */
  int max_sw = 0; bool sw_init = false;
  long long cells = 0;
  long long cell_off = 0;
  /* state of the current score's wavefront, uniform across the group */
  bool cur_exists = true;
  int clo[5], chi[5];
  int term_k;
  int end_k = KNONE, end_off = OFFNULL;
  int end_score = 0;                          /* original units */
  int status;                                 /* 1 end reached, 2 unreachable, 3 max steps */

  /* ---- score 0: wavefront_aligner_init_wf_m, W/wavefront/wavefront_aligner.c:251-310 ---- */
  {
    const bool ef = P.endsfree && P.match == 0;
    const int lo = ef ? -P.pbf : 0, hi = ef ? P.tbf : 0;
    if (hi - lo + 1 > wcap) return PAIR_OVERFLOW;
    /* every metadata slot starts null: scores < 0 read as the null wavefront */
    for (int i = g.lrank; i < P.mr * NC; i += g.lsize) meta[i] = make_int4(1, -1, 0, 0);
    g.sync();
    int t = KNONE;
    OffT* const mslot = gm.ring[CM];
    for (int k = lo + g.rank; k <= hi; k += g.size) {
      const int off0 = k > 0 ? k : 0;
      const int off = extend_offset(gm.pw, gm.tw, plen, tlen, k, off0, wild);
      mslot[k & wmask] = (OffT)off;
      if (term_cell(P, plen, tlen, ak, k, off)) t = imin(t, k);
    }
    int r[1] = {t};
    g.template allmin<1>(r);
    term_k = r[0];
    for (int c = 0; c < 5; ++c) { clo[c] = 1; chi[c] = -1; }
    clo[CM] = lo; chi[CM] = hi;
    if (g.lrank == 0) meta[CM] = make_int4(lo, hi, 0, 1);
    g.lsync();
  }

#if defined(WFA_GRID_TIMING) && defined(__CUDA_ARCH__)   /* debugging build: where the cycles of a step go (one thread per CTA) */
  long long tg_prev = clock64(), tg_acc[6] = {0, 0, 0, 0, 0, 0};
#define WFA_TG(i) { if (G::kGrid) { const long long t_ = clock64(); tg_acc[i] += t_ - tg_prev; tg_prev = t_; } }
#else
#define WFA_TG(i)
#endif
  for (;;) {
    /* ---- after-extend step of score s (extend.c:90-125 / :263-297) ---- */
    WFA_TG(5)
    if (cur_exists) {
      if (term_k != KNONE) {
        end_k = term_k;
        end_off = G::kGrid ? (int)ld_cg(gm.ring[CM] + cm * wcap + (term_k & wmask)) : (int)gm.ring[CM][cm * wcap + (term_k & wmask)];   /* (shared-memory rings: plain load) */
        status = 1; end_score = s * P.g;
        cells += imax(0, chi[CM] - clo[CM] + 1);
        break;
      }
      if (P.heuristic != 0 && clo[CM] <= chi[CM]) {
        /* wavefront_heuristic_cufoff, heuristic.c:509-567 */
        --steps_wait;
        const int lo_base = clo[CM], hi_base = chi[CM];
        const OffT* const mslot = gm.ring[CM] + cm * wcap;
        if (steps_wait <= 0) {
          if (P.heuristic == 1) {
            /* wavefront_heuristic_wfadaptive, heuristic.c:257-293 */
            if (hi_base - lo_base + 1 >= P.min_wf_len) {
              int dm = INT_MAX;
              for (int k = lo_base + g.rank; k <= hi_base; k += g.size) {
                const int f = G::kGrid ? (int)ld_cg(mslot + (k & wmask)) : (int)mslot[k & wmask];
                const int d = (f >= 0) ? imax(plen - (f - k), tlen - f) : (1 << 30);
                dm = imin(dm, d);
              }
              int r1[1] = {dm};
              g.template allmin<1>(r1);
              const int min_d = imin(imax(plen, tlen), r1[0]);
              int kf = INT_MAX, kl = INT_MIN;
              for (int k = lo_base + g.rank; k <= hi_base; k += g.size) {
                const int f = G::kGrid ? (int)ld_cg(mslot + (k & wmask)) : (int)mslot[k & wmask];
                const int d = (f >= 0) ? imax(plen - (f - k), tlen - f) : (1 << 30);
                if (d - min_d <= P.max_dist_thr) { kf = imin(kf, k); kl = imax(kl, k); }
              }
              int r2[2] = {kf, kl == INT_MIN ? INT_MAX : -kl};
              g.template allmin<2>(r2);
              kf = r2[0]; kl = (r2[1] == INT_MAX) ? INT_MIN : -r2[1];
              const int top_limit = imin(ak, hi_base);
              const int nlo = (kf < top_limit) ? kf : imax(lo_base, top_limit);
              const int bottom = imax(ak, nlo);
              const int nhi = (kl > bottom) ? kl : imin(hi_base, bottom);
              clo[CM] = nlo; chi[CM] = nhi;
              steps_wait = P.steps_between;
            }
          } else {
            /* wavefront_heuristic_xdrop, heuristic.c:329-383 (+ sw scores :297-328) */
            const int swg = (P.match != 0) ? -P.match : -1;
            const int so = s * P.g;
            int cmax = INT_MIN, kf = INT_MAX, kl = INT_MIN;
            for (int k = lo_base + g.rank; k <= hi_base; k += g.size) {
              const int f = G::kGrid ? (int)ld_cg(mslot + (k & wmask)) : (int)mslot[k & wmask];
              if (f < 0) continue;
              const int sw = (swg * (2 * f - k) - so) / 2;
              cmax = imax(cmax, sw);
              if (sw_init && max_sw - sw < P.xdrop) { kf = imin(kf, k); kl = imax(kl, k); }
            }
            int r3[3] = {cmax == INT_MIN ? INT_MAX : -cmax, kf, kl == INT_MIN ? INT_MAX : -kl};
            g.template allmin<3>(r3);
            cmax = (r3[0] == INT_MAX) ? INT_MIN : -r3[0];
            kf = r3[1]; kl = (r3[2] == INT_MAX) ? INT_MIN : -r3[2];
            if (sw_init) {
              const int nlo = (kf == INT_MAX) ? hi_base + 1 : kf;
              const int nhi = (kl >= nlo) ? kl : imin(nlo - 1, hi_base);
              clo[CM] = nlo; chi[CM] = nhi;
              if (cmax > max_sw) max_sw = cmax;
            } else { max_sw = cmax; sw_init = true; }
            steps_wait = P.steps_between;
          }
        }
        if (clo[CM] != lo_base || chi[CM] != hi_base) {
          /* wf_heuristic_equate, heuristic.c:161-172 */
          for (int c = 1; c < NC; ++c) {
            if (clo[c] > chi[c]) continue;
            if (clo[CM] > clo[c]) clo[c] = clo[CM];
            if (chi[CM] < chi[c]) chi[c] = chi[CM];
          }
          if (g.lrank == 0) {
            int4* const mrow = meta + (s & mmask) * NC;
            for (int c = 0; c < NC; ++c) {
              const bool nn = clo[c] <= chi[c];
              int4 m = mrow[c];
              m.x = nn ? clo[c] : 1; m.y = nn ? chi[c] : -1;
              mrow[c] = m;
            }
          }
          g.sync();
        }
      }
      cells += imax(0, chi[CM] - clo[CM] + 1);
    }

    /* ---- compute score s+1 (compute_affine.c:229-260 / compute_affine2p.c:334-368) ---- */
    ++s;
    if (++cm == P.rm) cm = 0;
    if (++c1 == P.r1) c1 = 0;
    if (TWO_P) { if (++c2 == P.r2) c2 = 0; }
    {
      /* fetch_input, compute.c:298-344: one 16-byte metadata read per source component */
      /* gap-linear / edit / indel are never two-piece: the flags are dead code in those instantiations */
      const bool m_only = !TWO_P && P.m_only != 0, no_mis = !TWO_P && P.no_mis != 0, edit_like = !TWO_P && P.edit_like != 0;
      int4 aMo2 = make_int4(1, -1, 0, 0), aI2 = aMo2, aD2 = aMo2;
      const int4 aMx = no_mis ? aMo2 : meta[((s - P.dx) & mmask) * NC + CM];
      const int4 aMo1 = meta[((s - P.doe1) & mmask) * NC + CM];
      const int4* const rowe1 = meta + ((s - P.de1) & mmask) * NC;
      const int4 aI1 = m_only ? aMo2 : rowe1[CI1], aD1 = m_only ? aMo2 : rowe1[CD1];
      if (TWO_P) {
        aMo2 = meta[((s - P.doe2) & mmask) * NC + CM];
        const int4* const rowe2 = meta + ((s - P.de2) & mmask) * NC;
        aI2 = rowe2[CI2]; aD2 = rowe2[CD2];
      }
      const bool n_mx = aMx.x > aMx.y, n_mo1 = aMo1.x > aMo1.y, n_i1 = aI1.x > aI1.y, n_d1 = aD1.x > aD1.y;
      const bool n_mo2 = TWO_P ? aMo2.x > aMo2.y : true;
      const bool n_i2 = TWO_P ? aI2.x > aI2.y : true;
      const bool n_d2 = TWO_P ? aD2.x > aD2.y : true;
      int4* const mrow = meta + (s & mmask) * NC;
      if ((n_mx && n_mo1 && n_i1 && n_d1 && n_mo2 && n_i2 && n_d2) || (edit_like && n_mo1)) {
        /* null step: allocate_output_null, compute.c:374-400.  wavefront_compute_edit (compute_edit.c:329-374)
         * has no null steps: a null predecessor sets num_null_steps = INT_MAX, "unreachable" at once (below) */
        cur_exists = false;
        for (int c = 0; c < 5; ++c) { clo[c] = 1; chi[c] = -1; }
        term_k = KNONE;
        for (int c = g.lrank; c < NC; c += g.lsize) mrow[c] = make_int4(1, -1, 0, 0);
        g.lsync();
      } else {
        s_exist = s * P.g;
        Src<OffT> sMx, sMo1, sI1, sD1, sMo2, sI2, sD2;
        sMx.slot = gm.ring[CM] + aMx.z * wcap; sMx.lo = aMx.x; sMx.hi = aMx.y;
        sMo1.slot = gm.ring[CM] + aMo1.z * wcap; sMo1.lo = aMo1.x; sMo1.hi = aMo1.y;
        sI1.slot = gm.ring[CI1] + aI1.z * wcap; sI1.lo = aI1.x; sI1.hi = aI1.y;
        sD1.slot = gm.ring[CD1] + aD1.z * wcap; sD1.lo = aD1.x; sD1.hi = aD1.y;
        if (TWO_P) {
          sMo2.slot = gm.ring[CM] + aMo2.z * wcap; sMo2.lo = aMo2.x; sMo2.hi = aMo2.y;
          sI2.slot = gm.ring[CI2] + aI2.z * wcap; sI2.lo = aI2.x; sI2.hi = aI2.y;
          sD2.slot = gm.ring[CD2] + aD2.z * wcap; sD2.lo = aD2.x; sD2.hi = aD2.y;
        }
        /* wavefront_compute_limits_input, compute.c:40-86 (null inputs carry lo=1, hi=-1) */
        int lo = sMx.lo, hi = sMx.hi;
        lo = imin(lo, sMo1.lo - 1); hi = imax(hi, sMo1.hi + 1);
        if (edit_like) { lo = sMo1.lo - 1; hi = sMo1.hi + 1; }       /* compute_edit.c:346-347 */
        if (!m_only) {                                             /* compute.c:52-58: gap-linear stops at the M sources */
          lo = imin(lo, sI1.lo + 1); hi = imax(hi, sI1.hi + 1);
          lo = imin(lo, sD1.lo - 1); hi = imax(hi, sD1.hi - 1);
        }
        if (TWO_P) {
          lo = imin(lo, sMo2.lo - 1); hi = imax(hi, sMo2.hi + 1);
          lo = imin(lo, sI2.lo + 1); hi = imax(hi, sI2.hi + 1);
          lo = imin(lo, sD2.lo - 1); hi = imax(hi, sD2.hi - 1);
        }
        const int width = hi - lo + 1;
        if (width > wcap) return PAIR_OVERFLOW;
        if (FULL) { if (s >= P.scap || cell_off + width > P.hcap) return PAIR_OVERFLOW; }
        /* allocate_output, compute.c:401-486 */
        const bool has_i1 = !m_only && (!n_mo1 || !n_i1), has_d1 = !m_only && (!n_mo1 || !n_d1);
        const bool has_i2 = TWO_P && (!n_mo2 || !n_i2), has_d2 = TWO_P && (!n_mo2 || !n_d2);
        OffT* const oM = gm.ring[CM] + cm * wcap;
        OffT* const oI1 = gm.ring[CI1] + c1 * wcap;
        OffT* const oD1 = gm.ring[CD1] + c1 * wcap;
        OffT* const oI2 = TWO_P ? gm.ring[CI2] + c2 * wcap : nullptr;
        OffT* const oD2 = TWO_P ? gm.ring[CD2] + c2 * wcap : nullptr;
        /* reductions: [2c] = first in-bounds k, [2c+1] = -(last in-bounds k), [10] = term */
        int red[2 * 5 + 1];
        for (int i = 0; i < 2 * 5 + 1; ++i) red[i] = INT_MAX;
        WFA_TG(1)
        /* (r02 experiment, not kept: four diagonals per thread and pass -- all source loads first, then the first
         * sequence fetches -- for the several-CTAs-per-pair group, whose rings sit behind L2.  16 x 100 kbp got 7 %
         * slower: ncu shows issue slots 51 % busy, DRAM 14 %, L2 20 %, i.e. the step is bound by its ~400
         * warp-instructions per cell under the 64-register cap and by the per-score barrier, not by latency.) */
        for (int k = lo + g.rank; k <= hi; k += g.size) {
          const int km = k & wmask, kl = (k - 1) & wmask, kr = (k + 1) & wmask;
          const int i1o = rd<G::kGrid>(sMo1, k - 1, kl), i1e = rd<G::kGrid>(sI1, k - 1, kl);
          const int d1o = rd<G::kGrid>(sMo1, k + 1, kr), d1e = rd<G::kGrid>(sD1, k + 1, kr);
          const int mis = rd<G::kGrid>(sMx, k, km) + 1;
          const int ins1 = imax(i1o, i1e) + 1, del1 = imax(d1o, d1e);
          int ins = ins1, del = del1;
          int i2o = OFFNULL, i2e = OFFNULL, d2o = OFFNULL, d2e = OFFNULL, ins2 = OFFNULL, del2 = OFFNULL;
          if (TWO_P) {
            i2o = rd<G::kGrid>(sMo2, k - 1, kl); i2e = rd<G::kGrid>(sI2, k - 1, kl);
            d2o = rd<G::kGrid>(sMo2, k + 1, kr); d2e = rd<G::kGrid>(sD2, k + 1, kr);
            ins2 = imax(i2o, i2e) + 1; del2 = imax(d2o, d2e);
            ins = imax(ins1, ins2); del = imax(del1, del2);
          }
          int mx = imax(del, imax(mis, ins));
          const bool m_in = in_bounds(k, mx, plen, tlen);
          if (!m_in) mx = OFFNULL;
          if (has_i1) { oI1[km] = off_store<OffT>(ins1); if (in_bounds(k, ins1, plen, tlen)) { red[2] = imin(red[2], k); red[3] = imin(red[3], -k); } }
          if (has_d1) { oD1[km] = off_store<OffT>(del1); if (in_bounds(k, del1, plen, tlen)) { red[4] = imin(red[4], k); red[5] = imin(red[5], -k); } }
          if (TWO_P) {
            if (has_i2) { oI2[km] = off_store<OffT>(ins2); if (in_bounds(k, ins2, plen, tlen)) { red[6] = imin(red[6], k); red[7] = imin(red[7], -k); } }
            if (has_d2) { oD2[km] = off_store<OffT>(del2); if (in_bounds(k, del2, plen, tlen)) { red[8] = imin(red[8], k); red[9] = imin(red[9], -k); } }
          }
          if (FULL) {
            /* origin code: winner of max over (offset<<4 | type), backtrace.c:366-389 */
            const int x1 = (i1e >= i1o) ? 1 : 0, y1 = (d1e >= d1o) ? 1 : 0;
            int best = (imax(mis, -1) << 4) | BT_M;
            best = imax(best, (imax(ins1, -1) << 4) | (BT_I1_OPEN + x1));
            best = imax(best, (imax(del1, -1) << 4) | (BT_D1_OPEN + y1));
            int code = (x1 << 4) | (y1 << 5);
            if (TWO_P) {
              const int x2 = (i2e >= i2o) ? 1 : 0, y2 = (d2e >= d2o) ? 1 : 0;
              best = imax(best, (imax(ins2, -1) << 4) | (BT_I2_OPEN + x2));
              best = imax(best, (imax(del2, -1) << 4) | (BT_D2_OPEN + y2));
              code |= (x2 << 6) | (y2 << 7);
            }
            code |= (best >= 0) ? (best & 15) : 0;
            gm.h_code[cell_off + (k - lo)] = (uint8_t)code;
          }
          if (m_in) {
            red[0] = imin(red[0], k); red[1] = imin(red[1], -k);
            mx = extend_offset(gm.pw, gm.tw, plen, tlen, k, mx, wild);
            if (term_cell(P, plen, tlen, ak, k, mx)) red[10] = imin(red[10], k);
          }
          oM[km] = off_store<OffT>(mx);
        }
        WFA_TG(2)
        if (TWO_P) g.template allmin<11>(red);
        else {
          int r7[7] = {red[0], red[1], red[2], red[3], red[4], red[5], red[10]};
          g.template allmin<7>(r7);
          red[0] = r7[0]; red[1] = r7[1]; red[2] = r7[2]; red[3] = r7[3]; red[4] = r7[4]; red[5] = r7[5];
          red[10] = r7[6];
        }
        WFA_TG(3)
        term_k = red[10];
        /* trim_ends, compute.c:571-605: [first in-bounds, last in-bounds], else null */
        const bool has[5] = {true, has_i1, has_d1, has_i2, has_d2};
        for (int c = 0; c < 5; ++c) {
          if (c < NC && has[c] && red[2 * c] != INT_MAX) { clo[c] = red[2 * c]; chi[c] = -red[2 * c + 1]; }
          else { clo[c] = 1; chi[c] = -1; }
        }
        if (!TWO_P && P.edit_prune && chi[CM] - clo[CM] + 1 >= 1000) {
          /* exact pruning of the ends whose best case |k - ak| is worse than the best worst case
           * max(remaining v, remaining h); the reference looks at the offsets BEFORE their extension,
           * which are recomputed here from the source (edit: M[s-1] for all three moves; its descriptor is
           * fetched again so that nothing stays live across the loop above) */
          const int lo_t = clo[CM], hi_t = chi[CM];
          const int4 aQ = meta[((s - 1) & mmask) * NC + CM];
          Src<OffT> sQ;
          sQ.slot = gm.ring[CM] + aQ.z * wcap; sQ.lo = aQ.x; sQ.hi = aQ.y;
          auto pre = [&](int k) {
            const int km = k & wmask, kl = (k - 1) & wmask, kr = (k + 1) & wmask;
            const int v = imax(rd<G::kGrid>(sQ, k + 1, kr), imax(rd<G::kGrid>(sQ, k, km), rd<G::kGrid>(sQ, k - 1, kl)) + 1);
            return in_bounds(k, v, plen, tlen) ? v : OFFNULL;
          };
          auto best = [&](int k) { return k >= ak ? k - ak : ak - k; };
          const int sample_k = lo_t + (hi_t - lo_t) / 2;
          const int so = pre(sample_k);
          const int smax = imax(plen - (so - sample_k), tlen - so);
          if (so >= 0 && !(best(lo_t) <= smax && best(hi_t) <= smax)) {
            int mw = INT_MAX;
            for (int k = lo_t + g.rank; k <= hi_t; k += g.size) {
              const int f = pre(k);
              if (f >= 0) mw = imin(mw, imax(plen - (f - k), tlen - f));
            }
            int r1[1] = {mw};
            g.template allmin<1>(r1);
            mw = r1[0];
            /* best(k) is |k - ak|: the surviving range is [ak - mw, ak + mw] clipped, scanned as the reference does */
            int nlo = lo_t, nhi = hi_t;
            while (nlo <= hi_t && best(nlo) > mw) ++nlo;
            while (nhi > nlo && best(nhi) > mw) --nhi;
            clo[CM] = nlo; chi[CM] = nhi;
          }
        }
        cur_exists = true;
        for (int c = g.lrank; c < NC; c += g.lsize) {
          /* lane c publishes component c */
          int l = clo[0], h = chi[0], z = cm;
          if (c == 1) { l = clo[1]; h = chi[1]; z = c1; }
          if (c == 2) { l = clo[2]; h = chi[2]; z = c1; }
          if (TWO_P) {
            if (c == 3) { l = clo[3]; h = chi[3]; z = c2; }
            if (c == 4) { l = clo[4]; h = chi[4]; z = c2; }
          }
          mrow[c] = make_int4(l, h, z, 1);
        }
        if (FULL) {
          if (g.rank == 0) { HistRow hr; hr.off = cell_off; hr.lo = lo; hr.pad = 0; gm.hmeta[s] = hr; }
          cell_off += width;
        }
        g.lsync();          /* ring stores were made visible by the barrier inside allmin */
        WFA_TG(4)
#if defined(WFA_GRID_TIMING) && defined(__CUDA_ARCH__)
        ++tg_acc[0];
#endif
      }
    }
    /* unreachable (extend.c:99-106: M[s] missing and num_null_steps > max_score_scope) and the
     * step limit (unialign.c:98-109), decided in ORIGINAL score units: between two multiples
     * of g every score is a null step of the reference. */
    {
      const int so = s * P.g;
      if (!cur_exists) {
        const int su = (!TWO_P && P.edit_like) ? so : s_exist + P.max_scope + 1;
        if (su <= so && su < P.max_steps) { status = 2; end_score = su; break; }
      }
      if (so >= P.max_steps) {
        status = 3;
        if (so == P.max_steps) cells += imax(0, chi[CM] - clo[CM] + 1);
        break;
      }
    }
  }

#if defined(WFA_GRID_TIMING) && defined(__CUDA_ARCH__)
  if (G::kGrid && g.lrank == 0 && P.dbg) for (int i = 0; i < 6; ++i) atomicAdd(P.dbg + i, (unsigned long long)tg_acc[i]);
#endif
  /* ---- wavefront_unialign_terminate, unialign.c:147-237 ---- */
  res.cells = cells;
  res.nruns = 0;
  res.locs[0] = res.locs[1] = res.locs[2] = res.locs[3] = 0;
  if (status == 3) {
    res.score = -P.max_steps; res.status = ST_MAX_STEPS;
  } else if (!FULL) {
    if (status == 1) { res.score = classic_score(P.match, plen, tlen, end_score, P.pos_score); res.status = ST_COMPLETED; }
    else {
      /* end position was never assigned: end_v = NULL - DIAGONAL_NULL with int32 wrap */
      const int32_t end_v = (int32_t)((uint32_t)OFFNULL - (uint32_t)INT_MAX);
      res.score = classic_score(P.match, end_v, OFFNULL, end_score, P.pos_score); res.status = ST_PARTIAL;
    }
  } else {
    if (status == 1) {
      if (g.rank == 0) {
        FwdEmitter em; em.init(gm.runs_stage, P.runcap);
        const int n = backtrace_codes(P, gm.h_code, gm.hmeta, s, end_k, plen, tlen, gm.pw, gm.tw, gm.ops, gm.opcap, em, wild);
        res.nruns = n;
        if (n >= 0) locations_from_runs(gm.runs_stage, imin(n, P.runcap), plen, tlen, res.locs);
      }
      res.score = classic_score(P.match, end_off - end_k, end_off, end_score, P.pos_score);
      res.status = ST_COMPLETED;
    } else {
      /* dropped: no end position -> empty CIGAR; maxtrim on an empty CIGAR clears the
       * score (W/alignment/cigar.c:473-528) */
      res.score = INT32_MIN; res.status = ST_PARTIAL;
    }
  }
  return PAIR_DONE;
}

}  // namespace wfagpu
