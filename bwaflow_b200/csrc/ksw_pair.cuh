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
  uint32_t B2, A2, NAB, K, HKC, OED2, ED2, OEI2, EI2, NEI2;
  uint32_t one, neg1, sh16, kmul;   // 1, -1, 65536 and the key multiplier as run-time values: multiply-adds by them stay IMADs (FMA pipe)
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
  c.NEI2 = (uint32_t)(uint16_t)(-sc.e_ins) * one;      // per-half -e_ins, added inside the DPX instruction
  c.one = 1u; c.neg1 = 0xFFFFFFFFu; c.sh16 = 65536u; c.kmul = 128u;
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

// Retires the job of slot S: its query becomes all-T (code 3) while an idle slot's target base is A (code 0), so the values the idle
// half goes on computing while it shadows its partner's window see a mismatch in every cell and can only decay.  That bounds
// them by the finished job's own (valid) values: they can never overflow 16 bits and spill into the partner's half.
__device__ __forceinline__ void pair_retire(int S, uint32_t *qa, int stride, int qlen)
{
  uint16_t *q16 = (uint16_t *)qa + S;
  for (int g = 0; g <= (qlen >> 3); ++g) q16[(size_t)g * stride * 2] = 0xFFFFu;
}

struct PairRow {                 // packed loop-carried state of one row of both slots
  uint32_t f, left, m, tt;
  int mj0, mj1;
};

// One cell of both slots.  qsh = query codes of the two slots already shifted to bits 14-15 / 30-31 (other bits arbitrary).
__device__ __forceinline__ void pair_cell(uint2 *p, uint32_t qsh, int j, PairRow &r, const PairConsts &c)
{
  const uint2 v = *p;
  const uint32_t d = (qsh ^ r.tt) & 0xC000C000u;
  const uint32_t nm = __vminu2(d, 0x00010001u);                      // 1 = mismatch
  const uint32_t x = mad_fma(nm, c.NAB, mad_fma(v.x, 1u, c.A2));
  const uint32_t hk = mad_fma(v.x, c.K, c.HKC);
  const uint32_t M = __vmins2(x, hk);
  const uint32_t h = __vimax3_s16x2(M, v.y, r.f);
  const uint32_t e2 = __vimax3_s16x2(sub_fma(M, c.OED2), sub_fma(v.y, c.ED2), c.B2);
  r.f = __viaddmax_s16x2(r.f, c.NEI2, __vmaxs2(sub_fma(M, c.OEI2), c.B2));   // one dependent op per cell on the F chain
  bool ph, pl;
  r.m = __vibmax_s16x2(h, r.m, &ph, &pl);                            // predicates: h >= running max (last column wins ties)
  r.mj0 = pl ? j : r.mj0;
  r.mj1 = ph ? j : r.mj1;
  *p = make_uint2(r.left, e2);
  r.left = h;
}

// A cell of a column group that is only partly inside the two windows: computed unconditionally (the columns exist and hold
// valid halves), every effect gated per half by gm (0xFFFF per half whose slot has this column inside its window).
// A gated-out half contributes h = 0 < BIAS to the running maximum, so neither the maximum nor its column can change.
__device__ __forceinline__ void pair_cell_gated(uint2 *p, uint32_t qsh, int j, PairRow &r, const PairConsts &c, uint32_t gm)
{
  const uint2 v = *p;
  const uint32_t d = (qsh ^ r.tt) & 0xC000C000u;
  const uint32_t nm = __vminu2(d, 0x00010001u);
  const uint32_t x = mad_fma(nm, c.NAB, mad_fma(v.x, 1u, c.A2));
  const uint32_t hk = mad_fma(v.x, c.K, c.HKC);
  const uint32_t M = __vmins2(x, hk);
  const uint32_t hg = __vimax3_s16x2(M, v.y, r.f) & gm;
  const uint32_t e2 = __vimax3_s16x2(sub_fma(M, c.OED2), sub_fma(v.y, c.ED2), c.B2);
  const uint32_t f2 = __viaddmax_s16x2(r.f, c.NEI2, __vmaxs2(sub_fma(M, c.OEI2), c.B2));
  r.f = (f2 & gm) | (r.f & ~gm);
  bool ph, pl;
  r.m = __vibmax_s16x2(hg, r.m, &ph, &pl);
  r.mj0 = pl ? j : r.mj0;
  r.mj1 = ph ? j : r.mj1;
  *p = make_uint2((r.left & gm) | (v.x & ~gm), (e2 & gm) | (v.y & ~gm));
  r.left = hg | (r.left & ~gm);
}

// Columns of one row for both slots of every lane of the warp: slot 0 in [a0,a1), slot 1 in [b0,b1) (empty = [0,0)).
// The warp walks aligned 4-column subgroups from the leftmost window start to the rightmost window end of ANY lane, and
// decides per subgroup, by vote, between the plain cell (every lane has the subgroup inside both of its windows) and the
// gated cell (per-lane, per-half masks; a lane with nothing there rewrites what it read).  Control flow is therefore
// warp-uniform: no divergence, no reconvergence stacks, and the cells of a subgroup overlap in the pipeline.
// Must be called by all 32 lanes.
__device__ __forceinline__ void pair_cols(uint2 *st, const uint32_t *qa, int stride, int a0, int a1, int b0, int b1, PairRow &r,
                                          const PairConsts &c)
{
  const unsigned full = 0xffffffffu;
  const int in0 = max(a0, b0), in1 = min(a1, b1);                  // columns inside both windows of this lane
  const int lo = a0 < a1 ? (b0 < b1 ? min(a0, b0) : a0) : (b0 < b1 ? b0 : 0x7fffffff);
  const int hi = max(a1, b1);
  const bool nowin = hi == 0;                                      // a lane without any window goes along with the plain cell: what it
                                                                   // computes on its retired halves can only decay (pair_retire)
  int jb = __reduce_min_sync(full, lo) & ~3;
  const int jend = __reduce_max_sync(full, hi);
  uint2 *p = st + (size_t)jb * stride;
#pragma unroll 1
  for (; jb < jend; jb += 4, p += 4 * stride) {
    uint32_t qw = qa[(jb >> 3) * stride];
    if (jb & 4) qw >>= 8;
    if (__all_sync(full, nowin || (jb >= in0 && jb + 4 <= in1))) {
#pragma unroll
      for (int u = 0; u < 4; ++u) pair_cell(p + u * stride, mad_fma(qw, 1u << (14 - 2 * u), 0u), jb + u, r, c);
    } else {
      // bit u (slot 0) / bit 16+u (slot 1) = column jb+u inside the slot's window
      const int sa = min(max(a0 - jb, 0), 4), ea = min(max(a1 - jb, 0), 4), sb = min(max(b0 - jb, 0), 4), eb = min(max(b1 - jb, 0), 4);
      const uint32_t mm = (((1u << ea) - 1u) & ~((1u << sa) - 1u)) | (((1u << eb) - 1u) & ~((1u << sb) - 1u)) << 16;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        pair_cell_gated(p + u * stride, mad_fma(qw, 1u << (14 - 2 * u), 0u), jb + u, r, c, ((mm >> u) & 0x00010001u) * 0xFFFFu);
    }
  }
}

// Band of the row (ksw.c:416-423).  k.lo is a lower bound of the reference's `beg` (exact when EXACT): columns below `beg`
// hold the all-zero state, which the recurrence reproduces (H = E = F = 0 gives h = E' = F' = 0 and passes h1 = 0 on), so starting a
// row anywhere in [max(0, i-w), beg] gives the same row.  c0 = the column this row starts at.
template <bool EXACT>
__device__ __forceinline__ void slot_row_begin(KswSlot &k, const DevScore &sc, int &left, int &c0)
{
  if (k.lo < k.i - k.w) k.lo = k.i - k.w;
  if (k.hi > k.i + k.w + 1) k.hi = k.i + k.w + 1;
  if (k.hi > k.qlen) k.hi = k.qlen;
  c0 = k.lo;
  if (!EXACT) { c0 &= ~3; if (c0 < k.i - k.w) c0 = k.i - k.w; }     // start on a subgroup boundary when the band allows
  left = 0;
  if (c0 == 0) { left = k.h0 - (sc.o_del + sc.e_del * (k.i + 1)); left = left > 0 ? left : 0; }
}

// ksw.c:449-464 for one slot after its row's cells (left/m unbiased).  Returns true when the job is finished.
template <bool EXACT>
__device__ __forceinline__ bool slot_row_score(KswSlot &k, int S, uint2 *st, int stride, const DevScore &sc, int bias, int left, int m, int mj)
{
  uint16_t *h16 = (uint16_t *)st + S, *e16 = (uint16_t *)st + 2 + S;
  const size_t cs = (size_t)stride * 4;                // u16 elements per column
  const int jend = k.lo < k.hi ? k.hi : k.lo;
  if (EXACT) k.cells += (unsigned)(jend - k.lo);
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
  return false;
}

struct TrimProbe { uint32_t a, b, d; };                // (H,E) of columns lo, lo+1 and hi-1, each as H | E << 16

__device__ __forceinline__ void slot_trim_probe(const KswSlot &k, int S, const uint2 *st, int stride, TrimProbe &t)
{
  const uint16_t *h16 = (const uint16_t *)st + S, *e16 = (const uint16_t *)st + 2 + S;
  const size_t cs = (size_t)stride * 4;
  t.a = h16[k.lo * cs] | (uint32_t)e16[k.lo * cs] << 16;
  t.b = h16[(k.lo + 1) * cs] | (uint32_t)e16[(k.lo + 1) * cs] << 16;
  t.d = h16[(k.hi - 1) * cs] | (uint32_t)e16[(k.hi - 1) * cs] << 16;
}

// ksw.c:466-469: next row's window.  `left` = the value just stored at eh[hi].h (its e is 0).  Returns true when all rows are done.
// The right end is always exact (columns beyond it keep stale values that later rows may read, so none may be touched).  The left
// end is exact when EXACT; otherwise it advances by at most two columns per row, which keeps it a lower bound (a row with
// m > 0 has a non-zero column, so neither scan of the reference can run past the other end).
template <bool EXACT>
__device__ __forceinline__ bool slot_trim(KswSlot &k, int S, const uint2 *st, int stride, uint32_t z2, const TrimProbe &t, int left)
{
  const uint16_t *h16 = (const uint16_t *)st + S, *e16 = (const uint16_t *)st + 2 + S;
  const size_t cs = (size_t)stride * 4;
  const uint16_t z = (uint16_t)z2;
  int j = k.lo;                                        // here lo < hi
  if (t.a == z2) {
    ++j;
    if (j < k.hi && t.b == z2) {
      ++j;
      if (EXACT) while (j < k.hi && h16[j * cs] == z && e16[j * cs] == z) ++j;
    }
  }
  k.lo = j;
  j = k.hi;
  if (left == 0) {                                     // eh[hi] is all zero
    --j;
    if (j >= k.lo && t.d == z2) {
      --j;
      while (j >= k.lo && h16[j * cs] == z && e16[j * cs] == z) --j;
    }
  }
  k.hi = j + 2 < k.qlen ? j + 2 : k.qlen;
  return ++k.i >= k.tlen;
}

// One row of each active slot of this lane.  t0/t1 = target bases of the slots' current rows (0..3).
// done0/done1 are set when the slot's job finished with this row.  EXACT: keep the reference's exact left band end and count
// the DP cells the reference executes (work-unit counter); the default trades that for register-only row bookkeeping.
template <bool EXACT>
__device__ __forceinline__ void pair_step(uint2 *st, const uint32_t *qa, int stride, const DevScore &sc, const PairConsts &c,
                                          KswSlot &k0, KswSlot &k1, bool act0, bool act1, int t0, int t1, bool &done0, bool &done1)
{
  int l0 = 0, l1 = 0, a0 = 0, b0 = 0;
  if (act0) slot_row_begin<EXACT>(k0, sc, l0, a0);
  if (act1) slot_row_begin<EXACT>(k1, sc, l1, b0);
  // windows; an idle slot shadows the other one (it computes on all-zero halves that nobody reads, see pair_reset)
  int a1 = k0.hi, b1 = k1.hi;
  if (!act0 || k0.lo >= a1) { a0 = 0; a1 = 0; }
  if (!act1 || k1.lo >= b1) { b0 = 0; b1 = 0; }
  if (!act0) { a0 = b0; a1 = b1; }
  if (!act1) { b0 = a0; b1 = a1; }
  PairRow r;
  r.f = c.B2; r.m = c.B2;
  r.left = (uint32_t)(l0 + c.bias) | (uint32_t)(l1 + c.bias) << 16;
  r.tt = (uint32_t)t0 << 14 | (uint32_t)t1 << 30;
  r.mj0 = -1; r.mj1 = -1;
  const PairRow r0 = r;
  pair_cols(st, qa, stride, a0, a1, b0, b1, r, c);
  if (a1 == 0 && b1 == 0) r = r0;                       // no window: whatever the lane computed along with the warp is discarded
  const int e0 = (int)(r.left & 0xFFFFu) - c.bias, e1 = (int)(r.left >> 16) - c.bias;
  done0 = act0 && slot_row_score<EXACT>(k0, 0, st, stride, sc, c.bias, e0, (int)(r.m & 0xFFFFu) - c.bias, r.mj0);
  done1 = act1 && slot_row_score<EXACT>(k1, 1, st, stride, sc, c.bias, e1, (int)(r.m >> 16) - c.bias, r.mj1);
  const bool tr0 = act0 && !done0, tr1 = act1 && !done1;
  TrimProbe p0, p1;
  if (tr0) slot_trim_probe(k0, 0, st, stride, p0);
  if (tr1) slot_trim_probe(k1, 1, st, stride, p1);
  if (tr0) done0 = slot_trim<EXACT>(k0, 0, st, stride, c.B2, p0, e0);
  if (tr1) done1 = slot_trim<EXACT>(k1, 1, st, stride, c.B2, p1, e1);
}

} // namespace bwaflow
