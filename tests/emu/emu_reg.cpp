/*
 * emu_reg.cpp -- TEST INFRASTRUCTURE ONLY (tests/).  Runs the warp code of the register-resident
 * tier (pywfa_b200/csrc/wfa_reg.cuh) on the CPU through the 32-lane host model lanevec_host.h,
 * so that `pytest -m "not gpu"` can compare it with the oracle.  Never linked into the product.
 */
#include "lanevec_host.h"

#include <cuda_runtime.h>   /* int2 / int4 types only; nothing is linked */
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/wfagpu.h"
#include "../../pywfa_b200/csrc/pack.h"
#include "../../pywfa_b200/csrc/wfa_params.h"
#include "../../pywfa_b200/csrc/wfa_reg.cuh"

using namespace wfagpu;

template <int P, int DX, int DOE>
static int run_pair(bool full, const RegParams& R, const uint32_t* pw, const uint32_t* tw, int plen, int tlen,
                    uint8_t* hist, uint8_t* ops, uint32_t* stage, PairResult& res) {
  std::vector<uint32_t> pwin((size_t)plen + 1), twin((size_t)tlen + 1);
  build_windows(pw, plen, pwin.data());
  build_windows(tw, tlen, twin.data());
  const lv::histref h = lv::make_histref(hist, false);
  if (full) return align_pair_reg<P, DX, DOE, true>(R, pw, tw, pwin.data(), twin.data(), plen, tlen, h, ops, stage, true, res);
  return align_pair_reg<P, DX, DOE, false>(R, pw, tw, pwin.data(), twin.data(), plen, tlen, h, ops, stage, true, res);
}

/* byte mode (what wfa_regb_kernel does per pair): bytes -> 4-bit symbol codes -> windows of 8 bases; a pair holding
 * a byte outside the symbol set is handed on like an overflow */
template <int P, int DX, int DOE>
static int run_pair_bytes(bool full, const RegParams& R, const uint32_t* pb, const uint32_t* tb, int plen, int tlen, int wild,
                          uint8_t* hist, uint8_t* ops, uint32_t* stage, PairResult& res) {
  std::vector<uint32_t> pn((size_t)(plen >> 3) + 2), tn((size_t)(tlen >> 3) + 2);
  const bool okp = nibble_words(pb, plen, pn.data(), wild), okt = nibble_words(tb, tlen, tn.data(), wild);
  if (!okp || !okt) return PAIR_OVERFLOW;
  std::vector<uint32_t> pwin((size_t)plen + 1), twin((size_t)tlen + 1);
  build_windows<4>(pn.data(), plen, pwin.data());
  build_windows<4>(tn.data(), tlen, twin.data());
  const lv::histref h = lv::make_histref(hist, false);
  if (full) return align_pair_reg<P, DX, DOE, true, false, 4>(R, pb, tb, pwin.data(), twin.data(), plen, tlen, h, ops, stage, true, res, wild);
  return align_pair_reg<P, DX, DOE, false, false, 4>(R, pb, tb, pwin.data(), twin.data(), plen, tlen, h, ops, stage, true, res, wild);
}

/* the symbol table of byte mode: code of one byte (in the low byte of a word, 1 valid base); *bad = outside the set */
extern "C" int emu_nib_code(int byte, int wild, int* bad) {
  bool b = false;
  const uint32_t w = lv::nib_pack8((uint32_t)byte & 0xffu, 0u, 1, (uint32_t)wild, b);
  *bad = b ? 1 : 0;
  return (int)(w & 15u) | (int)((w >> 4) != 0u ? 0x100 : 0);      /* (nothing may leak into the padding codes) */
}

/* regs = packed registers per wavefront (window = 64 * regs diagonals); hrows = origin rows */
extern "C" int emu_reg_align_batch(const wfagpu_config_t* cfg, const uint8_t* seq, const int64_t* p_off,
                                   const int32_t* p_len, const int64_t* t_off, const int32_t* t_len, int64_t n,
                                   int regs, int hrows, int32_t* score, int32_t* status, int32_t* locs,
                                   int64_t* cig_off, uint32_t* runs, int64_t runs_cap, int32_t* overflow,
                                   int64_t* cells) {
  KParams K;
  memset(&K, 0, sizeof K);
  fill_kparams(*cfg, K);
  const bool mapped = metric_as_affine(*cfg, K);       /* score-only linear / edit / indel as zero-opening gap-affine */
  if ((cfg->distance != WFAGPU_DISTANCE_AFFINE && !mapped) || cfg->heuristic != WFAGPU_HEURISTIC_NONE) return -4;
  /* the penalty shapes the product instantiates (wfa_kernels.cu: reg_shape) */
  const int shape = K.de1 != 1 ? -1 : (K.dx == 2 && K.doe1 == 4) ? 0 : (K.dx == 1 && K.doe1 == 1) ? 1 : (K.dx == 2 && K.doe1 == 1) ? 2
                    : (K.dx == 4 && K.doe1 == 7) ? 3 : (K.dx == 1 && K.doe1 == 2) ? 4 : (K.dx == 1 && K.doe1 == 3) ? 5 : -1;
  if (shape < 0) return -4;
  const bool full = cfg->scope == WFAGPU_SCOPE_FULL;
  RegParams R;
  R.match = K.match; R.g = K.g; R.max_steps = K.max_steps; R.pos_score = K.pos_score;
  R.endsfree = K.endsfree; R.pbf = K.pbf; R.pef = K.pef; R.tbf = K.tbf; R.tef = K.tef;
  R.hrows = hrows;
  const RegWindow rw = reg_window(regs, K.endsfree, K.match, K.pbf, K.tbf);
  R.kbase = rw.kbase; R.c_lo = rw.c_lo; R.c_hi = rw.c_hi;
  std::vector<uint8_t> hist((size_t)hrows * 64 * regs + 64);
  int64_t used = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int plen = p_len[i], tlen = t_len[i];
    std::vector<uint32_t> pw((plen + 15) / 16 + 1, 0), tw((tlen + 15) / 16 + 1, 0);
    const int wc = cfg->wildcard & 0xff;
    bool bytes = wc == 'A' || wc == 'C' || wc == 'G' || wc == 'T';
    if (!bytes) bytes = !pack_sequence(seq + p_off[i], plen, pw.data()) || !pack_sequence(seq + t_off[i], tlen, tw.data());
    if (bytes) {
      /* byte mode of the library (wfa_pack.cu pack_bytes_kernel): upper-cased bytes, 4 per word */
      pw.assign((plen + 3) / 4 + 1, 0); tw.assign((tlen + 3) / 4 + 1, 0);
      auto put = [](const uint8_t* s8, int len, uint32_t* out) {
        for (int j = 0; j < len; ++j) { uint8_t c = s8[j]; if (c >= 'a' && c <= 'z') c -= 32; out[j >> 2] |= (uint32_t)c << (8 * (j & 3)); }
      };
      put(seq + p_off[i], plen, pw.data()); put(seq + t_off[i], tlen, tw.data());
    }
    std::vector<uint32_t> stage((size_t)plen + tlen + 2);
    std::vector<uint8_t> ops((size_t)plen + tlen + 8);
    R.runcap = (int)stage.size(); R.opcap = (int)ops.size();
    PairResult res;
    memset(&res, 0, sizeof res);
    int rc;
    if (plen > REG_MAX_LEN || tlen > REG_MAX_LEN) rc = PAIR_OVERFLOW;
#define RUN(PP, DX, DOE) rc = bytes ? run_pair_bytes<PP, DX, DOE>(full, R, pw.data(), tw.data(), plen, tlen, wc, hist.data(), ops.data(), stage.data(), res) \
                                   : run_pair<PP, DX, DOE>(full, R, pw.data(), tw.data(), plen, tlen, hist.data(), ops.data(), stage.data(), res)
#define RUN_SHAPE(PP) do { if (shape == 0) RUN(PP, 2, 4); else if (shape == 1) RUN(PP, 1, 1); else if (shape == 2) RUN(PP, 2, 1); \
                           else if (shape == 3) RUN(PP, 4, 7); else if (shape == 4) RUN(PP, 1, 2); else RUN(PP, 1, 3); } while (0)
    else if (regs == 1) RUN_SHAPE(1);
    else if (regs == 2) RUN_SHAPE(2);
    else if (regs == 3) RUN_SHAPE(3);
    else if (regs == 4) RUN_SHAPE(4);
    else return -3;
#undef RUN_SHAPE
#undef RUN
    cig_off[i] = used;
    overflow[i] = (rc == PAIR_OVERFLOW);
    if (rc == PAIR_OVERFLOW) { score[i] = 0; status[i] = 0; cells[i] = 0; memset(locs + 4 * i, 0, 16); continue; }
    score[i] = res.score; status[i] = res.status; cells[i] = res.cells;
    memcpy(locs + 4 * i, res.locs, 16);
    if (res.nruns < 0 || used + res.nruns > runs_cap) return -1;
    for (int r = 0; r < res.nruns; ++r) runs[used + r] = stage[r];
    used += res.nruns;
  }
  cig_off[n] = used;
  return 0;
}
