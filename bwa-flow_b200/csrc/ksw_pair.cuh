// Two-jobs-per-lane banded extension engine: exact ksw_extend2 semantics (bwa/ksw.c:380-479) with the DP cell
// computed in packed 16-bit halves, so that every DPX instruction (VIMNMX3.S16x2, VIMNMX.S16x2, VIMNMX.U16x2)
// advances one cell of TWO independent extension jobs: slot 0 in the low half of each word, slot 1 in the high half.
// As in ksw_lane.cuh a lane advances ONE ROW per call for each of its slots (a row must be complete before the next
// row's [beg,end) window is known, bwa/ksw.c:466-469), and all lanes of a warp reconverge at row boundaries.
//
// Row state in shared memory, per (column j, lane): uint2 { H , E } with
//     H half s = H_s(i-1, j-1) + BIAS     (eh[j].h of the reference for slot s)
//     E half s = E_s(i,   j)   + BIAS     (eh[j].e)
// and the query of a slot as 2-bit codes, 8 columns per 16-bit half of one word per (column group, lane).
// Everything is kept BIASED by a constant (>= mismatch + gap open + gap extend) so that all packed intermediates are
// non-negative: plain 32-bit multiply-adds (FMA pipe) then act as two independent 16-bit adds without borrows, and only
// min/max run on the integer ALU / DPX pipe, which is the one that saturates.
//
// Restrictions of the packed engine (jobs outside go to the one-job-per-lane int32 engine, never to a CPU):
//   * the 4x4 part of the scoring matrix has the bwa_fill_scmat form (bwa/bwa.c:109-119): a on the diagonal, -b elsewhere;
//   * no ambiguous base (code 4) in the query or the target window of the job;
//   * (h0 + qlen*a)*(1+a) + BIAS fits a signed 16-bit half.
//
// Cell recurrence (ksw.c:431-447), per half, biased:
//   x  = Hd + (q==t ? a : -b)                          one VIMNMX.U16x2 (0/1 mismatch flag) + multiply-adds
//   M  = min(x, Hd*(1+a) - BIAS*a)                     = "Hd ? Hd+s : 0" up to values <= 0, which all later max() discard
//   h  = max(M, E, F)                                  VIMNMX3.S16x2
//   E' = max(M - oe_del, E - e_del, 0)                 VIMNMX3.S16x2 on multiply-add results
//   F' = max(M - oe_ins, F - e_ins, 0)                 VIMNMX3.S16x2
//   row max + last column attaining it                 VIMNMX.S16x2 with its two predicate outputs + 2 selects
#pragma once
#include "internal.cuh"
#include "ksw_lane.cuh"

namespace bwaflow {

struct PairConsts {
  int ok;                      // scoring matrix has the packed form
  int a, b, bias;
  uint32_t B2, A2, NAB, K, HKC, OED2, ED2, OEI2, EI2;
};

inline void fill_pair_consts(const bwaflow_opt_t &o, const DevScore &sc, PairConsts &c)
{
  c.ok = 1;
  const int a = o.mat[0], b = -o.mat[1];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (o.mat[i * 5 + j] != (i == j ? a : -b)) c.ok = 0;
  if (a <= 0 || b <= 0 || a > 60 || b > 60) c.ok = 0;
  int oe = sc.o_del + sc.e_del > sc.o_ins + sc.e_ins ? sc.o_del + sc.e_del : sc.o_ins + sc.e_ins;
  int ex = sc.e_del > sc.e_ins ? sc.e_del : sc.e_ins;
  if (sc.o_del < 0 || sc.o_ins < 0 || sc.e_del <= 0 || sc.e_ins <= 0 || oe > 2000) c.ok = 0;
  c.a = a; c.b = b;
  c.bias = b + oe + ex;
  const uint32_t one = 0x00010001u;
  c.B2 = (uint32_t)c.bias * one;
  c.A2 = (uint32_t)a * one;
  c.NAB = (uint32_t)(-(a + b));
  c.K = (uint32_t)(1 + a);
  c.HKC = (uint32_t)0 - (uint32_t)(c.bias * a) * one;
  c.OED2 = (uint32_t)(sc.o_del + sc.e_del) * one; c.ED2 = (uint32_t)sc.e_del * one;
  c.OEI2 = (uint32_t)(sc.o_ins + sc.e_ins) * one; c.EI2 = (uint32_t)sc.e_ins * one;
}

__device__ __forceinline__ uint32_t mad_fma(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t sub_fma(uint32_t a, uint32_t b) { return mad_fma(b, 0xffffffffu, a); }   // a - b on the FMA pipe

// Scalar state of one slot (one job).
struct KswSlot {
  int qlen, tlen, w, h0;
  int i, lo, hi;
  int best, best_i, best_j, end_i, gscore, max_off;
  unsigned cells;
};

enum { PAIR_OK = 0, PAIR_SLOW = 1, PAIR_BAD = 2 };

// Row -1 of slot S (ksw.c:395-397), its clamped band (ksw.c:399-407) and its packed query.  st = this lane's column 0,
// qa = this lane's group 0.  Returns PAIR_SLOW when the job must run on the int32 engine, PAIR_BAD when it cannot run at all.
template <class QF>
__device__ __forceinline__ int pair_begin(KswSlot &k, int S, uint2 *st, uint32_t *qa, int stride, const DevScore &sc,
                                          const PairConsts &c, int qlen_, int tlen_, int w_, int h0_, int end_bonus, QF qf)
{
  k.qlen = qlen_; k.tlen = tlen_; k.h0 = h0_;
  int lim = gap_limit(qlen_ * sc.smax + end_bonus - sc.o_ins, sc.e_ins);
  k.w = w_ < lim ? w_ : lim;
  lim = gap_limit(qlen_ * sc.smax + end_bonus - sc.o_del, sc.e_del);
  k.w = k.w < lim ? k.w : lim;
  k.i = 0; k.lo = 0; k.hi = qlen_;
  k.best = h0_; k.best_i = k.best_j = -1; k.end_i = -1; k.gscore = -1; k.max_off = 0; k.cells = 0;
  if (h0_ + qlen_ * sc.smax > KSW_MAX_SCORE || h0_ <= 0) return PAIR_BAD;
  if ((h0_ + qlen_ * c.a) * (1 + c.a) + c.bias > 32767) return PAIR_SLOW;
  const int gi = sc.o_ins + sc.e_ins;
  uint16_t *h16 = (uint16_t *)st + S, *e16 = (uint16_t *)st + 2 + S, *q16 = (uint16_t *)qa + S;
  int hv = h0_;                                          // H(-1,-1)
  uint32_t qw = 0, bad = 0;
  for (int j = 0; j < qlen_; ++j) {
    h16[(size_t)j * stride * 4] = (uint16_t)(hv + c.bias);
    e16[(size_t)j * stride * 4] = (uint16_t)c.bias;
    const uint32_t q = qf(j);
    bad |= q;
    qw |= (q & 3u) << ((j & 7) * 2);
    if ((j & 7) == 7) { q16[(size_t)(j >> 3) * stride * 2] = (uint16_t)qw; qw = 0; }
    hv = h0_ - gi - j * sc.e_ins;                        // H(-1, j)
    hv = hv > 0 ? hv : 0;
  }
  h16[(size_t)qlen_ * stride * 4] = (uint16_t)(hv + c.bias);
  e16[(size_t)qlen_ * stride * 4] = (uint16_t)c.bias;
  if (qlen_ & 7) q16[(size_t)(qlen_ >> 3) * stride * 2] = (uint16_t)qw;
  return bad > 3 ? PAIR_SLOW : PAIR_OK;
}

// Puts the halves of slot S back to the all-zero (= BIAS) state for columns 0..qlen.  That state is absorbing under the cell
// recurrence (H = 0 gives M <= 0, so h, E', F' stay 0), which is what makes it safe for an idle slot to shadow its
// partner's window: the values it computes there can never grow and spill into the partner's half.
__device__ __forceinline__ void pair_reset(int S, uint2 *st, int stride, const PairConsts &c, int qlen)
{
  uint16_t *h16 = (uint16_t *)st + S, *e16 = (uint16_t *)st + 2 + S;
  for (int j = 0; j <= qlen; ++j) {
    h16[(size_t)j * stride * 4] = (uint16_t)c.bias;
    e16[(size_t)j * stride * 4] = (uint16_t)c.bias;
  }
}

struct PairRow {                 // packed loop-carried state of one row of both slots
  uint32_t f, left, m, tt;
  int mj0, mj1;
};

// One cell of both slots.  qsh = query codes of the two slots already shifted to bits 14-15 / 30-31 (other bits arbitrary).
template <bool MASKED>
__device__ __forceinline__ void pair_cell(uint2 *p, uint32_t qsh, int j, PairRow &r, const PairConsts &c, int sel)
{
  const uint2 v = *p;
  const uint32_t d = (qsh ^ r.tt) & 0xC000C000u;
  const uint32_t nm = __vminu2(d, 0x00010001u);                      // 1 = mismatch
  const uint32_t x = mad_fma(nm, c.NAB, mad_fma(v.x, 1u, c.A2));
  const uint32_t hk = mad_fma(v.x, c.K, c.HKC);
  const uint32_t M = __vmins2(x, hk);
  const uint32_t h = __vimax3_s16x2(M, v.y, r.f);
  const uint32_t e2 = __vimax3_s16x2(sub_fma(M, c.OED2), sub_fma(v.y, c.ED2), c.B2);
  r.f = __vimax3_s16x2(sub_fma(M, c.OEI2), sub_fma(r.f, c.EI2), c.B2);
  bool ph, pl;
  r.m = __vibmax_s16x2(h, r.m, &ph, &pl);                            // predicates: h >= running max (last column wins ties)
  r.mj0 = pl ? j : r.mj0;
  r.mj1 = ph ? j : r.mj1;
  if (!MASKED) *p = make_uint2(r.left, e2);
  else {                                                             // only slot `sel` is inside its window here
    uint16_t *p16 = (uint16_t *)p + sel;
    const int sh = sel * 16;
    p16[0] = (uint16_t)(r.left >> sh);
    p16[2] = (uint16_t)(e2 >> sh);
  }
  r.left = h;
}

template <bool MASKED>
__device__ __forceinline__ void pair_cols(uint2 *st, const uint32_t *qa, int stride, int j0, int j1, PairRow &r,
                                          const PairConsts &c, int sel)
{
  int j = j0;
  uint2 *p = st + (size_t)j * stride;
  for (; j < j1 && (j & 7); ++j, p += stride) pair_cell<MASKED>(p, qa[(j >> 3) * stride] << (14 - 2 * (j & 7)), j, r, c, sel);
  for (; j + 8 <= j1; j += 8, p += 8 * stride) {
    const uint32_t qw = qa[(j >> 3) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) pair_cell<MASKED>(p + u * stride, mad_fma(qw, 1u << (14 - 2 * u), 0u), j + u, r, c, sel);
  }
  for (; j < j1; ++j, p += stride) pair_cell<MASKED>(p, qa[(j >> 3) * stride] << (14 - 2 * (j & 7)), j, r, c, sel);
}

__device__ __forceinline__ void slot_row_begin(KswSlot &k, const DevScore &sc, int &left)
{
  if (k.lo < k.i - k.w) k.lo = k.i - k.w;
  if (k.hi > k.i + k.w + 1) k.hi = k.i + k.w + 1;
  if (k.hi > k.qlen) k.hi = k.qlen;
  left = 0;
  if (k.lo == 0) { left = k.h0 - (sc.o_del + sc.e_del * (k.i + 1)); left = left > 0 ? left : 0; }
}

// Everything of ksw.c:449-469 for one slot after its row's cells.  left/m are unbiased.  Returns true when the job is finished.
__device__ __forceinline__ bool slot_row_end(KswSlot &k, int S, uint2 *st, int stride, const DevScore &sc, int bias, int left, int m, int mj)
{
  uint16_t *h16 = (uint16_t *)st + S, *e16 = (uint16_t *)st + 2 + S;
  const size_t cs = (size_t)stride * 4;                // u16 elements per column
  const int jend = k.lo < k.hi ? k.hi : k.lo;
  k.cells += (unsigned)(jend - k.lo);
  h16[k.hi * cs] = (uint16_t)(left + bias);            // eh[end] = {h1, 0}   ksw.c:449
  e16[k.hi * cs] = (uint16_t)bias;
  if (jend == k.qlen) {                                // ksw.c:450-453
    if (left >= k.gscore) k.end_i = k.i;
    k.gscore = left > k.gscore ? left : k.gscore;
  }
  if (m <= 0 || k.lo >= k.hi) return true;             // m == 0 (or empty row)   ksw.c:454
  if (m > k.best) {                                    // ksw.c:455-457
    k.best = m; k.best_i = k.i; k.best_j = mj;
    const int d = mj > k.i ? mj - k.i : k.i - mj;
    k.max_off = d > k.max_off ? d : k.max_off;
  } else if (sc.zdrop > 0) {                           // ksw.c:458-464
    const int di = k.i - k.best_i, dj = mj - k.best_j;
    if (di > dj) { if (k.best - m - (di - dj) * sc.e_del > sc.zdrop) return true; }
    else if (k.best - m - (dj - di) * sc.e_ins > sc.zdrop) return true;
  }
  const uint16_t z = (uint16_t)bias;                   // ksw.c:466-469
  int j = k.lo;
  while (j < k.hi && h16[j * cs] == z && e16[j * cs] == z) ++j;
  k.lo = j;
  j = k.hi;
  while (j >= k.lo && h16[j * cs] == z && e16[j * cs] == z) --j;
  k.hi = j + 2 < k.qlen ? j + 2 : k.qlen;
  return ++k.i >= k.tlen;
}

// One row of each active slot of this lane.  t0/t1 = target bases of the slots' current rows (0..3).
// done0/done1 are set when the slot's job finished with this row.
__device__ __forceinline__ void pair_step(uint2 *st, const uint32_t *qa, int stride, const DevScore &sc, const PairConsts &c,
                                          KswSlot &k0, KswSlot &k1, bool act0, bool act1, int t0, int t1, bool &done0, bool &done1)
{
  int l0 = 0, l1 = 0;
  if (act0) slot_row_begin(k0, sc, l0);
  if (act1) slot_row_begin(k1, sc, l1);
  // windows; an idle slot shadows the other one (it computes on valid-format stale halves that nobody reads)
  int a0 = k0.lo, a1 = k0.hi, b0 = k1.lo, b1 = k1.hi;
  if (!act0 || a0 >= a1) { a0 = 0; a1 = 0; }
  if (!act1 || b0 >= b1) { b0 = 0; b1 = 0; }
  if (!act0) { a0 = b0; a1 = b1; }
  if (!act1) { b0 = a0; b1 = a1; }
  PairRow r;
  r.f = c.B2; r.m = c.B2;
  r.left = (uint32_t)(l0 + c.bias) | (uint32_t)(l1 + c.bias) << 16;
  r.tt = (uint32_t)t0 << 14 | (uint32_t)t1 << 30;
  r.mj0 = -1; r.mj1 = -1;
  // the four window ends in order: [v0,v1) at most one slot, [v1,v2) both (when they overlap), [v2,v3) at most one slot
  const int v0 = min(a0, b0), v3 = max(a1, b1);
  const int x1 = max(a0, b0), x2 = min(a1, b1);
  const int v1 = min(x1, x2), v2 = max(x1, x2);
  {                                                    // phase 1
    const int in0 = (v0 >= a0 && v1 <= a1), in1 = (v0 >= b0 && v1 <= b1);
    if (v0 < v1 && (in0 || in1)) {
      const int sel = in0 ? 0 : 1;
      const uint32_t keep = sel ? 0x0000FFFFu : 0xFFFF0000u;
      const PairRow s = r;
      pair_cols<true>(st, qa, stride, v0, v1, r, c, sel);
      r.f = (r.f & ~keep) | (s.f & keep); r.left = (r.left & ~keep) | (s.left & keep); r.m = (r.m & ~keep) | (s.m & keep);
      if (sel) r.mj0 = s.mj0; else r.mj1 = s.mj1;
    }
  }
  if (v1 < v2 && v1 >= a0 && v2 <= a1 && v1 >= b0 && v2 <= b1)   // phase 2: both slots inside their windows
    pair_cols<false>(st, qa, stride, v1, v2, r, c, 0);
  else if (v1 < v2) {                                  // disjoint windows: the gap belongs to nobody, unless one window is empty
    const int in0 = (v1 >= a0 && v2 <= a1), in1 = (v1 >= b0 && v2 <= b1);
    if (in0 || in1) {
      const int sel = in0 ? 0 : 1;
      const uint32_t keep = sel ? 0x0000FFFFu : 0xFFFF0000u;
      const PairRow s = r;
      pair_cols<true>(st, qa, stride, v1, v2, r, c, sel);
      r.f = (r.f & ~keep) | (s.f & keep); r.left = (r.left & ~keep) | (s.left & keep); r.m = (r.m & ~keep) | (s.m & keep);
      if (sel) r.mj0 = s.mj0; else r.mj1 = s.mj1;
    }
  }
  {                                                    // phase 3
    const int in0 = (v2 >= a0 && v3 <= a1), in1 = (v2 >= b0 && v3 <= b1);
    if (v2 < v3 && (in0 || in1)) {
      const int sel = in0 ? 0 : 1;
      const uint32_t keep = sel ? 0x0000FFFFu : 0xFFFF0000u;
      const PairRow s = r;
      pair_cols<true>(st, qa, stride, v2, v3, r, c, sel);
      r.f = (r.f & ~keep) | (s.f & keep); r.left = (r.left & ~keep) | (s.left & keep); r.m = (r.m & ~keep) | (s.m & keep);
      if (sel) r.mj0 = s.mj0; else r.mj1 = s.mj1;
    }
  }
  done0 = done1 = false;
  if (act0) done0 = slot_row_end(k0, 0, st, stride, sc, c.bias, (int)(r.left & 0xFFFFu) - c.bias, (int)(r.m & 0xFFFFu) - c.bias, r.mj0);
  if (act1) done1 = slot_row_end(k1, 1, st, stride, sc, c.bias, (int)(r.left >> 16) - c.bias, (int)(r.m >> 16) - c.bias, r.mj1);
}

} // namespace bwaflow
