/*
 * lanevec_host.h -- TEST INFRASTRUCTURE ONLY.  A 32-lane host model of the names in
 * pywfa_b200/csrc/lanevec.cuh, so that the warp code of the register-resident tier
 * (pywfa_b200/csrc/wfa_reg.cuh) can be executed on the CPU and compared with the oracle.
 * Every primitive is modelled on the documented behaviour of the sm_100a instruction it stands
 * for (VIMNMX.S16x2, VIADD.16x2, PRMT, SHFL, VOTE).
 */
#pragma once
#define WFA_LANEVEC_HOST 1
#include <stdint.h>

namespace wfagpu {
namespace lv {

struct vb { uint32_t m; };
struct vi { int32_t v[32]; };
struct vu { uint32_t v[32]; };

#define LV_FOR for (int l = 0; l < 32; ++l)
inline vi lane_id() { vi r; LV_FOR r.v[l] = l; return r; }
inline vi splati(int x) { vi r; LV_FOR r.v[l] = x; return r; }
inline vu splat(uint32_t x) { vu r; LV_FOR r.v[l] = x; return r; }
inline vb vfalse() { return vb{0}; }

/* vi arithmetic */
#define LV_BIN_VI(OP)                                                                         \
  inline vi operator OP(const vi& a, const vi& b) { vi r; LV_FOR r.v[l] = a.v[l] OP b.v[l]; return r; } \
  inline vi operator OP(const vi& a, int b) { vi r; LV_FOR r.v[l] = a.v[l] OP b; return r; }           \
  inline vi operator OP(int a, const vi& b) { vi r; LV_FOR r.v[l] = a OP b.v[l]; return r; }
LV_BIN_VI(+) LV_BIN_VI(-) LV_BIN_VI(&) LV_BIN_VI(|)
inline vi operator>>(const vi& a, int b) { vi r; LV_FOR r.v[l] = a.v[l] >> b; return r; }
inline vi operator<<(const vi& a, int b) { vi r; LV_FOR r.v[l] = (int32_t)((uint32_t)a.v[l] << b); return r; }
#define LV_CMP_VI(OP)                                                                          \
  inline vb operator OP(const vi& a, const vi& b) { vb r{0}; LV_FOR if (a.v[l] OP b.v[l]) r.m |= 1u << l; return r; } \
  inline vb operator OP(const vi& a, int b) { vb r{0}; LV_FOR if (a.v[l] OP b) r.m |= 1u << l; return r; }
LV_CMP_VI(>=) LV_CMP_VI(<=) LV_CMP_VI(>) LV_CMP_VI(<) LV_CMP_VI(==) LV_CMP_VI(!=)
inline vb operator&(const vb& a, const vb& b) { return vb{a.m & b.m}; }
inline vb operator|(const vb& a, const vb& b) { return vb{a.m | b.m}; }
inline vb vnot(const vb& a) { return vb{~a.m}; }

/* vu bit operations */
inline vu operator^(const vu& a, const vu& b) { vu r; LV_FOR r.v[l] = a.v[l] ^ b.v[l]; return r; }
inline vu operator|(const vu& a, const vu& b) { vu r; LV_FOR r.v[l] = a.v[l] | b.v[l]; return r; }
inline vu operator&(const vu& a, uint32_t b) { vu r; LV_FOR r.v[l] = a.v[l] & b; return r; }
inline vu operator~(const vu& a) { vu r; LV_FOR r.v[l] = ~a.v[l]; return r; }
inline vb operator!=(const vu& a, uint32_t b) { vb r{0}; LV_FOR if (a.v[l] != b) r.m |= 1u << l; return r; }

/* packed s16x2 */
inline int16_t lo16(uint32_t x) { return (int16_t)(x & 0xffffu); }
inline int16_t hi16(uint32_t x) { return (int16_t)(x >> 16); }
inline uint32_t mk2(int lo, int hi) { return ((uint32_t)(uint16_t)lo) | ((uint32_t)(uint16_t)hi << 16); }
inline vu vimax2(const vu& a, const vu& b) {
  vu r; LV_FOR r.v[l] = mk2(lo16(a.v[l]) > lo16(b.v[l]) ? lo16(a.v[l]) : lo16(b.v[l]),
                            hi16(a.v[l]) > hi16(b.v[l]) ? hi16(a.v[l]) : hi16(b.v[l]));
  return r;
}
inline vu vimax3(const vu& a, const vu& b, const vu& c) { return vimax2(vimax2(a, b), c); }
inline vu vadd2(const vu& a, const vu& b) {     /* wraps per half, like VIADD.16x2 */
  vu r; LV_FOR r.v[l] = mk2((int16_t)(uint16_t)((uint16_t)a.v[l] + (uint16_t)b.v[l]),
                            (int16_t)(uint16_t)((uint16_t)(a.v[l] >> 16) + (uint16_t)(b.v[l] >> 16)));
  return r;
}
inline vu vsub2(const vu& a, const vu& b) {     /* wraps per half */
  vu r; LV_FOR r.v[l] = mk2((int16_t)(uint16_t)((uint16_t)a.v[l] - (uint16_t)b.v[l]),
                            (int16_t)(uint16_t)((uint16_t)(a.v[l] >> 16) - (uint16_t)(b.v[l] >> 16)));
  return r;
}
inline vu operator>>(const vu& a, int b) { vu r; LV_FOR r.v[l] = a.v[l] >> b; return r; }
inline vi as_vi(const vu& a) { vi r; LV_FOR r.v[l] = (int32_t)a.v[l]; return r; }
inline vu vimax2p(const vu& a, const vu& b, vb& hi, vb& lo) {   /* predicates: a >= b */
  hi.m = lo.m = 0;
  LV_FOR {
    if (lo16(a.v[l]) >= lo16(b.v[l])) lo.m |= 1u << l;
    if (hi16(a.v[l]) >= hi16(b.v[l])) hi.m |= 1u << l;
  }
  return vimax2(a, b);
}
inline uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 0xf;
    uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
    if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;   /* sign replication */
    r |= byte << (8 * i);
  }
  return r;
}
inline vu prmt(const vu& a, const vu& b, const vu& sel) { vu r; LV_FOR r.v[l] = byte_perm(a.v[l], b.v[l], sel.v[l]); return r; }
inline vu signmask2(const vu& a) { vu r; LV_FOR r.v[l] = byte_perm(a.v[l], 0, 0xbb99); return r; }
inline vu bitsel(const vu& mask, const vu& a, const vu& b) { vu r; LV_FOR r.v[l] = (a.v[l] & mask.v[l]) | (b.v[l] & ~mask.v[l]); return r; }
inline vi sx_lo(const vu& a) { vi r; LV_FOR r.v[l] = lo16(a.v[l]); return r; }
inline vi sx_hi(const vu& a) { vi r; LV_FOR r.v[l] = hi16(a.v[l]); return r; }
inline vu pack2(const vi& lo, const vi& hi) { vu r; LV_FOR r.v[l] = mk2(lo.v[l], hi.v[l]); return r; }
inline vu put_lo(const vu& a, const vi& v) { vu r; LV_FOR r.v[l] = byte_perm(a.v[l], (uint32_t)v.v[l], 0x3254); return r; }
inline vu put_hi(const vu& a, const vi& v) { vu r; LV_FOR r.v[l] = byte_perm(a.v[l], (uint32_t)v.v[l], 0x5410); return r; }

/* lane exchange */
inline vu from_prev_lane(const vu& a) { vu r; LV_FOR r.v[l] = a.v[(l + 31) & 31]; return r; }
inline vu from_next_lane(const vu& a) { vu r; LV_FOR r.v[l] = a.v[(l + 1) & 31]; return r; }
inline vu from_lane(const vu& a, const vi& src) { vu r; LV_FOR r.v[l] = a.v[src.v[l] & 31]; return r; }
inline int lane_value(const vi& a, int lane) { return a.v[lane & 31]; }
inline uint32_t ballot(const vb& p) { return p.m; }
inline bool any(const vb& p) { return p.m != 0; }

/* per-lane helpers */
inline vi vsel(const vb& p, const vi& a, const vi& b) { vi r; LV_FOR r.v[l] = (p.m >> l & 1) ? a.v[l] : b.v[l]; return r; }
inline vu vselu(const vb& p, const vu& a, const vu& b) { vu r; LV_FOR r.v[l] = (p.m >> l & 1) ? a.v[l] : b.v[l]; return r; }
inline vi vmin(const vi& a, const vi& b) { vi r; LV_FOR r.v[l] = a.v[l] < b.v[l] ? a.v[l] : b.v[l]; return r; }
inline vi vmax(const vi& a, const vi& b) { vi r; LV_FOR r.v[l] = a.v[l] > b.v[l] ? a.v[l] : b.v[l]; return r; }
inline vi vffs0(const vu& x) { vi r; LV_FOR r.v[l] = x.v[l] ? __builtin_ctz(x.v[l]) : -1; return r; }
inline vu vfunnel_r(const vu& lo, const vu& hi, const vi& sh) {
  vu r; LV_FOR { const uint64_t v = ((uint64_t)hi.v[l] << 32) | lo.v[l]; r.v[l] = (uint32_t)(v >> (sh.v[l] & 31)); }
  return r;
}
inline vu nib_both(const vu& a, const vu& b) { vu r; LV_FOR r.v[l] = (a.v[l] & b.v[l] & 0x11111111u) * 15u; return r; }
inline vu operator&(const vu& a, const vu& b) { vu r; LV_FOR r.v[l] = a.v[l] & b.v[l]; return r; }
inline vi vclz(const vu& x) { vi r; LV_FOR r.v[l] = x.v[l] ? __builtin_clz(x.v[l]) : 32; return r; }
inline vu vbrev(const vu& x) {
  vu r;
  LV_FOR { uint32_t v = x.v[l], o = 0; for (int i = 0; i < 32; ++i) o |= ((v >> i) & 1u) << (31 - i); r.v[l] = o; }
  return r;
}
typedef const uint32_t* seqref;
inline seqref make_seqref(const uint32_t* p) { return p; }
inline vu load_win(seqref base, const vi& idx, const vb& p) { vu r; LV_FOR r.v[l] = (p.m >> l & 1) ? base[idx.v[l]] : 0u; return r; }
struct lanead { const uint32_t* base; vi idx0; };
inline lanead lane_addr(seqref base, const vi& idx0) { return lanead{base, idx0}; }
template <int IMM>
inline vu load_win_at(const lanead& a, const vi& idx, const vb& p, uint32_t dflt) {
  vu r; LV_FOR r.v[l] = (p.m >> l & 1) ? a.base[a.idx0.v[l] + idx.v[l] + IMM] : dflt; return r;
}
template <int IA, int IB>
inline vu diff_win_nonneg(const lanead& a, const lanead& b, const vi& idx) {
  vu r;
  LV_FOR r.v[l] = idx.v[l] >= 0 ? (a.base[a.idx0.v[l] + idx.v[l] + IA] ^ b.base[b.idx0.v[l] + idx.v[l] + IB]) : 0x80000000u;
  return r;
}
inline vi vaddmin(const vi& a, const vi& b, const vi& c) { vi r; LV_FOR { const int t = a.v[l] + b.v[l]; r.v[l] = t < c.v[l] ? t : c.v[l]; } return r; }
inline vb operator==(const vu& a, uint32_t b) { vb r{0}; LV_FOR if (a.v[l] == b) r.m |= 1u << l; return r; }
inline void fence_warp() {}          /* one host thread models the whole warp: nothing to order */
template <class T> inline void keep(T&) {}
inline void scatter_u32(uint32_t* base, const vi& idx, const vu& val, const vb& p) { LV_FOR if (p.m >> l & 1) base[idx.v[l]] = val.v[l]; }
inline vu gather_u32(const uint32_t* base, const vi& idx, const vb& p) { vu r; LV_FOR r.v[l] = (p.m >> l & 1) ? base[idx.v[l]] : 0u; return r; }
inline void scatter_u8(uint8_t* base, const vi& idx, const vi& val, const vb& p) { LV_FOR if (p.m >> l & 1) base[idx.v[l]] = (uint8_t)val.v[l]; }
struct histref { uint8_t* p; };
inline histref make_histref(uint8_t* p, bool) { return histref{p}; }
template <bool SH> inline void hist_store(const histref& h, int off, const vi& idx, const vi& val) { LV_FOR h.p[off + idx.v[l]] = (uint8_t)val.v[l]; }
template <bool SH> inline int hist_load(const histref& h, int off) { return h.p[off]; }
#undef LV_FOR

}  // namespace lv
}  // namespace wfagpu
