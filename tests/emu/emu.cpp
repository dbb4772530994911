/*
 * emu.cpp -- TEST INFRASTRUCTURE ONLY (tests/).  Compiles the device source
 * pywfa_b200/csrc/wfa_core.cuh as plain host C++ with a one-thread "group", so that the
 * CPU-only CI (`pytest -m "not gpu"`) can check the kernel's wavefront logic, capacity
 * handling and the host packer against the oracle without a GPU.  It is never linked into,
 * loaded by or reachable from the product package: libwfagpu.so has no CPU alignment path.
 */
#include <cuda_runtime.h>   /* int2 / make_int2 only; nothing is linked */
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/wfagpu.h"
#include "../../pywfa_b200/csrc/pack.h"
#include "../../pywfa_b200/csrc/wfa_core.cuh"
#include "../../pywfa_b200/csrc/wfa_params.h"

using namespace wfagpu;

struct EmuGroup {
  static constexpr bool kGrid = false;
  int rank = 0, size = 1, lrank = 0, lsize = 1;
  void sync() {}
  void lsync() {}
  template <int N> void allmin(int (&)[N]) {}
};

/* the several-CTAs-per-pair group's code path (G::kGrid), one thread */
struct EmuGridGroup : EmuGroup {
  static constexpr bool kGrid = true;
};

struct EmuBuffers {      /* allocated once per batch */
  std::vector<unsigned char> ring;
  std::vector<int4> meta;
  std::vector<uint8_t> h_code;
  std::vector<HistRow> hmeta;
};

template <class OffT, bool TWO_P, bool FULL, class Group = EmuGroup>
static int run_one(KParams& P, EmuBuffers& B, const uint32_t* pw, const uint32_t* tw, int plen, int tlen,
                   std::vector<uint32_t>& stage, PairResult& res) {
  std::vector<uint8_t> ops((size_t)plen + tlen + 8);
  const int wcap = P.wcap;
  GroupMem<OffT> gm;
  gm.pw = pw; gm.tw = tw;
  gm.wild = P.byte_mode ? P.wildcard : -1;
  gm.ring[CM] = reinterpret_cast<OffT*>(B.ring.data());
  gm.ring[CI1] = gm.ring[CM] + P.rm * wcap;
  gm.ring[CD1] = gm.ring[CI1] + P.r1 * wcap;
  gm.ring[CI2] = gm.ring[CD1] + P.r1 * wcap;
  gm.ring[CD2] = gm.ring[CI2] + (TWO_P ? P.r2 * wcap : 0);
  gm.meta = B.meta.data();
  gm.h_code = B.h_code.data(); gm.hmeta = B.hmeta.data();
  gm.runs_stage = stage.data();
  gm.ops = ops.data(); gm.opcap = (int)ops.size();
  Group g;
  return align_pair<Group, OffT, TWO_P, FULL>(g, P, gm, plen, tlen, res);
}

/* off16 != 0 selects the int16 offset rings of the short-read tiers */
extern "C" int emu_align_batch(const wfagpu_config_t* cfg, const uint8_t* seq, const int64_t* p_off,
                               const int32_t* p_len, const int64_t* t_off, const int32_t* t_len, int64_t n,
                               int wcap, long long hcap, int scap, int off16, int32_t* score, int32_t* status,
                               int32_t* locs, int64_t* cig_off, uint32_t* runs, int64_t runs_cap,
                               int32_t* overflow, int64_t* cells) {
  KParams P;
  memset(&P, 0, sizeof P);
  fill_kparams(*cfg, P);
  const bool two_p = cfg->distance == WFAGPU_DISTANCE_AFFINE2P;
  const bool full = cfg->scope == WFAGPU_SCOPE_FULL;
  if (wcap & (wcap - 1)) return -3;            /* power of two */
  P.wcap = wcap; P.hcap = hcap; P.scap = scap;
  EmuBuffers B;
  {
    const int ns = P.rm + 2 * P.r1 + (two_p ? 2 * P.r2 : 0);
    B.ring.resize((size_t)ns * wcap * 4);
    B.meta.resize((size_t)P.mr * 5);
    B.h_code.resize(full ? (size_t)hcap : 1);
    B.hmeta.resize(full ? (size_t)scap : 1);
  }
  int64_t used = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int plen = p_len[i], tlen = t_len[i];
    std::vector<uint32_t> pw((plen + 15) / 16 + 1, 0), tw((tlen + 15) / 16 + 1, 0);
    const int wc = cfg->wildcard & 0xff;
    bool bytes = wc == 'A' || wc == 'C' || wc == 'G' || wc == 'T';
    if (!bytes) bytes = !pack_sequence(seq + p_off[i], plen, pw.data()) || !pack_sequence(seq + t_off[i], tlen, tw.data());
    if (bytes) {
      /* byte mode of the library (wfagpu_api.cpp batch_pack): upper-cased bytes, 4 per word */
      pw.assign((plen + 3) / 4 + 1, 0); tw.assign((tlen + 3) / 4 + 1, 0);
      auto put = [](const uint8_t* s8, int len, uint32_t* out) {
        for (int j = 0; j < len; ++j) { uint8_t c = s8[j]; if (c >= 'a' && c <= 'z') c -= 32; out[j >> 2] |= (uint32_t)c << (8 * (j & 3)); }
      };
      put(seq + p_off[i], plen, pw.data()); put(seq + t_off[i], tlen, tw.data());
    }
    P.byte_mode = bytes ? 1 : 0; P.wildcard = wc;
    std::vector<uint32_t> stage((size_t)plen + tlen + 2);
    P.runcap = (int)stage.size();
    PairResult res;
    memset(&res, 0, sizeof res);
    int rc;
#define EMU_RUN(T, G)                                                                                          \
    (two_p ? (full ? run_one<T, true, true, G>(P, B, pw.data(), tw.data(), plen, tlen, stage, res)   \
                   : run_one<T, true, false, G>(P, B, pw.data(), tw.data(), plen, tlen, stage, res)) \
           : (full ? run_one<T, false, true, G>(P, B, pw.data(), tw.data(), plen, tlen, stage, res)  \
                   : run_one<T, false, false, G>(P, B, pw.data(), tw.data(), plen, tlen, stage, res)))
    if (off16 == 2) rc = EMU_RUN(int32_t, EmuGridGroup);          /* off16 == 2: the grid group's code path */
    else if (off16) rc = EMU_RUN(int16_t, EmuGroup);
    else rc = EMU_RUN(int32_t, EmuGroup);
#undef EMU_RUN
    cig_off[i] = used;
    overflow[i] = (rc == PAIR_OVERFLOW);
    if (rc == PAIR_OVERFLOW) { score[i] = 0; status[i] = 0; cells[i] = 0; memset(locs + 4 * i, 0, 16); continue; }
    score[i] = res.score; status[i] = res.status; cells[i] = res.cells;
    memcpy(locs + 4 * i, res.locs, 16);
    if (used + res.nruns > runs_cap) return -1;
    if (res.nruns < 0) return -1;
    for (int r = 0; r < res.nruns; ++r) runs[used + r] = stage[r];
    used += res.nruns;
  }
  cig_off[n] = used;
  return 0;
}

/* expose the host packer for tests */
extern "C" int emu_pack(const uint8_t* s, int len, uint32_t* out) { return pack_sequence(s, len, out) ? 1 : 0; }

/* host side of the batch staging (pywfa_b200/csrc/pack.cpp), exposed to the CPU test-suite */
extern "C" void emu_scan_pairs(const int64_t* p_off, const int32_t* p_len, const int64_t* t_off, const int32_t* t_len, int64_t n,
                               int bpw, int64_t* out /* 9 + 3 * MAX_LEN_CLASSES */) {
  PairScan r;
  scan_pairs(p_off, p_len, t_off, t_len, n, bpw, &r);
  out[0] = r.first_negative; out[1] = r.back_to_back ? 1 : 0; out[2] = r.seq_bytes; out[3] = r.total_words;
  out[4] = r.lo; out[5] = r.hi; out[6] = r.maxp; out[7] = r.maxt; out[8] = ((int64_t)r.minp << 32) | (uint32_t)r.mint;
  for (int c = 0; c < MAX_LEN_CLASSES; ++c) {
    out[9 + c] = r.cls_n[c]; out[9 + MAX_LEN_CLASSES + c] = r.cls_maxp[c]; out[9 + 2 * MAX_LEN_CLASSES + c] = r.cls_maxt[c];
  }
}
extern "C" void emu_gather_pairs(const uint8_t* seq, const int64_t* p_off, const int32_t* p_len, const int64_t* t_off,
                                 const int32_t* t_len, int64_t n, uint8_t* dst, int64_t* np_off, int64_t* nt_off) {
  gather_pairs(seq, p_off, p_len, t_off, t_len, n, dst, np_off, nt_off);
}
extern "C" int emu_length_class_limit(int c) { return length_class_limit(c); }
