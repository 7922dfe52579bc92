// One-job-per-lane banded extension engine with TWO ADJACENT COLUMNS per packed word: exact ksw_extend2 semantics
// (bwa/ksw.c:380-479), every DPX instruction (VIMNMX3.S16x2, VIADDMNMX.S16x2, VIMNMX.S16x2/U16x2) advancing columns 2p and
// 2p+1 of one row of one job.  Compared with ksw_pair.cuh (two JOBS per lane) a lane owns one job, so
//   * every lane walks its OWN window [lo,hi) from its own start: no warp-wide union of windows, no gated cells;
//   * the row state is half as large per lane (4 bytes per column), which doubles the resident warps per SM;
//   * a row takes half as many dependent steps.
// The price is the F chain (f(j+1) depends on f(j) inside the word): three ALU-pipe instructions per pair instead of two.
//
// Row state in shared memory, per (pair p, lane): uint2 { Hw, Ew }, biased 16-bit halves (low = column 2p, high = 2p+1):
//     Hw half = H(i-1, j-1) + BIAS   (eh[j].h of the reference)        Ew half = E(i, j) + BIAS   (eh[j].e)
// plus one byte per pair with the two query codes (code(2p) | code(2p+1) << 4).  Same biased arithmetic and the same
// restrictions as ksw_pair.cuh (bwa_fill_scmat-form matrix, no N, 16-bit range) and additionally h0 + qlen*a + BIAS < 512:
// the row maximum and the last column attaining it (ksw.c:437-438) are one unsigned max over keys (H << 7 | pair index).
//
// What makes the per-lane walk legal: a column outside the reference's window [beg,end] holds the all-zero
// state, never a stale non-zero value -- the right scan of ksw.c:468 stops at the last non-zero column j and the next row ends
// at j+2, so everything it leaves outside was zero when it was left, and the band clamp only drops columns on the left, which
// are never read again.  Hence (i) a row may start one column early on an even boundary (the zero state reproduces itself), the
// one case where that column is not zero -- it was dropped by the band, not trimmed -- being zeroed explicitly; (ii) the right
// end is exact: nothing beyond the sentinel eh[end] is touched.
#pragma once
#include "ksw_pair.cuh"

namespace bwaflow {

struct ColJob {
  int qlen, tlen, w, h0;
  int i, lo, hi;
  int best, best_i, best_j, end_i, gscore, max_off;
};

constexpr int COLS_KEY_BITS = 7;            // pair index inside a row (rows of at most 2 * 127 columns)
constexpr int COLS_MAX_H = 511;             // biased scores must fit the remaining 9 bits of a key half

__host__ __device__ inline size_t cols_smem(int qmax) { return (size_t)32 * ((qmax >> 1) + 1) * (sizeof(uint2) + 1); }

// Row -1 (ksw.c:395-397), the clamped band (ksw.c:399-407) and the query bytes of the lane's job.
template <class QF>
__device__ __forceinline__ int cols_begin(ColJob &k, uint2 *st, uint8_t *qb, const DevScore &sc, const PairConsts &c,
                                          int qlen_, int tlen_, int w_, int h0_, int end_bonus, QF qf)
{
  k.qlen = qlen_; k.tlen = tlen_; k.h0 = h0_;
  int lim = gap_limit(qlen_ * sc.smax + end_bonus - sc.o_ins, sc.e_ins);
  k.w = w_ < lim ? w_ : lim;
  lim = gap_limit(qlen_ * sc.smax + end_bonus - sc.o_del, sc.e_del);
  k.w = k.w < lim ? k.w : lim;
  k.i = 0; k.lo = 0; k.hi = qlen_;
  k.best = h0_; k.best_i = k.best_j = -1; k.end_i = -1; k.gscore = -1; k.max_off = 0;
  if (h0_ + qlen_ * sc.smax > KSW_MAX_SCORE || h0_ <= 0) return PAIR_BAD;
  if ((h0_ + qlen_ * c.a) * (1 + c.a) + c.bias > 32767) return PAIR_SLOW;
  if (h0_ + qlen_ * c.a + c.bias > COLS_MAX_H || qlen_ > 2 * ((1 << COLS_KEY_BITS) - 1)) return PAIR_SLOW;
  const int gi = sc.o_ins + sc.e_ins;
  const int np = (qlen_ >> 1) + 1;                       // pairs holding columns 0..qlen
  uint32_t bad = 0;
  for (int p = 0; p < np; ++p) {
    const int j0 = 2 * p, j1 = j0 + 1;
    int v0 = j0 == 0 ? h0_ : h0_ - gi - (j0 - 1) * sc.e_ins;          // column j holds H(-1, j-1)
    int v1 = h0_ - gi - (j1 - 1) * sc.e_ins;
    v0 = v0 > 0 ? v0 : 0; v1 = v1 > 0 && j1 <= qlen_ ? v1 : 0;
    st[(size_t)p * 32] = make_uint2((uint32_t)(v0 + c.bias) | (uint32_t)(v1 + c.bias) << 16, c.B2);
    const uint32_t q0 = j0 < qlen_ ? qf(j0) : 0u, q1 = j1 < qlen_ ? qf(j1) : 0u;
    bad |= q0 | q1;
    qb[(size_t)p * 32] = (uint8_t)((q0 & 3u) | (q1 & 3u) << 4);
  }
  return bad > 3 ? PAIR_SLOW : PAIR_OK;
}

// The cell's constants, pinned in registers for the life of the kernel (read back from shared memory once): the compiler would otherwise re-load
// them from the constant bank inside the cell loop -- a multiply-add takes only one such operand -- at 3-4 issue slots per pair.
struct ColsK {
  uint32_t B2, A2, NAB, K, HKC, OED2, ED2, OEI2, NEI2, one, neg1, sh16, kmul, bias;
  // staged through shared memory (`stage` = 16 words, all lanes of the warp call this): a value read back from there is one
  // ptxas keeps in a register instead of re-loading it from the constant bank
  __device__ __forceinline__ void load(const PairConsts &c, volatile uint32_t *stage)
  {
    if ((threadIdx.x & 31) == 0) {
      stage[0] = c.B2; stage[1] = c.A2; stage[2] = c.NAB; stage[3] = c.K; stage[4] = c.HKC; stage[5] = c.OED2; stage[6] = c.ED2;
      stage[7] = c.OEI2; stage[8] = c.NEI2; stage[9] = c.one; stage[10] = c.neg1; stage[11] = c.sh16; stage[12] = c.kmul;
      stage[13] = (uint32_t)c.bias;
    }
    __syncwarp();
    B2 = stage[0]; A2 = stage[1]; NAB = stage[2]; K = stage[3]; HKC = stage[4]; OED2 = stage[5]; ED2 = stage[6]; OEI2 = stage[7];
    NEI2 = stage[8]; one = stage[9]; neg1 = stage[10]; sh16 = stage[11]; kmul = stage[12]; bias = stage[13];
    __syncwarp();
  }
};
// adds and shifts as multiply-adds by run-time constants: ptxas cannot turn them back into ALU-pipe VIADD / LEA
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, const ColsK &c) { return mad_fma(a, c.one, b); }
__device__ __forceinline__ uint32_t fsub(uint32_t a, uint32_t b, const ColsK &c) { return mad_fma(b, c.neg1, a); }

struct ColsRow {                 // loop-carried state of one row
  uint32_t prev;                 // h of the previous pair (high half = H(i, j-1) of the next pair's low column)
  uint32_t C;                    // F entering the next pair, in the high half (low half: any valid biased value)
  uint32_t kmax, trep;           // running maximum of (h << 7 | pair index) per half; pair index in both halves
};

// Columns 2p, 2p+1 of the row.  tt = target base at bits 14-15 and 30-31.
__device__ __forceinline__ void cols_cell(uint2 *pp, const uint8_t *qq, uint32_t tt, ColsRow &r, const ColsK &c, uint32_t &h_out,
                                          uint2 &w_out)
{
  const uint2 v = *pp;
  const uint32_t qv = *qq;
  const uint32_t d = (mad_fma(qv, 0x04004000u, 0u) ^ tt) & 0xC000C000u;     // code(2p) -> bits 14-15, code(2p+1) -> bits 30-31
  const uint32_t nm = __vminu2(d, 0x00010001u);                             // 1 = mismatch
  const uint32_t x = mad_fma(nm, c.NAB, fadd(v.x, c.A2, c));
  const uint32_t hk = mad_fma(v.x, c.K, c.HKC);
  const uint32_t M = __vmins2(x, hk);                                       // Hd ? Hd + s : (something <= 0)
  const uint32_t e2 = __vimax3_s16x2(fsub(M, c.OED2, c), fsub(v.y, c.ED2, c), c.B2);
  const uint32_t T0 = __vmaxs2(fsub(M, c.OEI2, c), c.B2);                   // max(M - oe_ins, 0), the sources of F
  const uint32_t tsh = mad_fma(T0, c.sh16, c.bias);               // (0, T0 of the low column)
  const uint32_t G = __viaddmax_s16x2(r.C, c.NEI2, tsh);                    // high half: F of the high column
  const uint32_t Fw = __byte_perm(r.C, G, 0x7632);                          // (F low column, F high column)
  r.C = __viaddmax_s16x2(G, c.NEI2, T0);                                    // high half: F entering the next pair
  const uint32_t h = __vimax3_s16x2(M, v.y, Fw);
  w_out = make_uint2(__byte_perm(r.prev, h, 0x5432), e2);                   // eh[j].h = H(i, j-1)
  r.prev = h;
  h_out = h;
}

// One row of the lane's job.  tb = target base of the row (0..3).  Returns true when the job finished with this row.
// Must be called by all 32 lanes (act = this lane has a job).
__device__ __forceinline__ bool cols_row(ColJob &k, uint2 *st, const uint8_t *qb, const DevScore &sc, const ColsK &c, bool act, int tb)
{
  int lo = k.lo, hi = k.hi, c0 = 0, npair = 0, h1 = 0;
  if (act) {
    const int bl = k.i - k.w;
    if (lo < bl) {                                       // ksw.c:416: the band drops columns on the left
      lo = bl;
      if (lo & 1) { uint16_t *z = (uint16_t *)(st + (size_t)(lo >> 1) * 32); z[0] = (uint16_t)c.bias; z[2] = (uint16_t)c.bias; }
    }
    if (hi > k.i + k.w + 1) hi = k.i + k.w + 1;
    if (hi > k.qlen) hi = k.qlen;
    c0 = lo & ~1;
    if (lo == 0) { h1 = k.h0 - (sc.o_del + sc.e_del * (k.i + 1)); h1 = h1 > 0 ? h1 : 0; }   // ksw.c:420-423
    npair = hi > c0 ? (hi - c0) >> 1 : 0;
  }
  ColsRow r;
  r.prev = ((uint32_t)h1 + c.bias) << 16 | c.bias;
  r.C = c.B2; r.kmax = 0; r.trep = 0;
  const uint32_t tt = (uint32_t)tb << 14 | (uint32_t)tb << 30;
  uint2 *pp = st + (size_t)(c0 >> 1) * 32;
  const uint8_t *qq = qb + (size_t)(c0 >> 1) * 32;
#pragma unroll 2
  for (int t = 0; t < npair; ++t, pp += 32, qq += 32) {
    uint32_t h; uint2 wv;
    cols_cell(pp, qq, tt, r, c, h, wv);
    r.kmax = __vmaxu2(r.kmax, mad_fma(h, c.kmul, r.trep));
    r.trep = fadd(r.trep, 0x00010001u, c);
    *pp = wv;
  }
  __syncwarp();
  if (!act) return false;
  int left = h1;                                          // H(i, hi-1), what the sentinel eh[hi].h receives
  if (lo < hi) {
    if ((hi - c0) & 1) {                                  // pair (hi-1, hi): the low column is the row's last, the high one the sentinel
      uint32_t h; uint2 wv;
      cols_cell(pp, qq, tt, r, c, h, wv);
      r.kmax = __vmaxu2(r.kmax, mad_fma(h & 0xFFFFu, c.kmul, r.trep & 0xFFFFu));
      wv.y = (wv.y & 0xFFFFu) | c.bias << 16;   // eh[hi].e = 0   ksw.c:449
      *pp = wv;
      left = (int)(h & 0xFFFFu) - (int)c.bias;
    } else {                                              // the sentinel is the low column of the next pair
      uint16_t *z = (uint16_t *)pp;
      left = (int)(r.prev >> 16) - (int)c.bias;
      z[0] = (uint16_t)(r.prev >> 16); z[2] = (uint16_t)c.bias;
    }
  }
  // ksw.c:450-464
  const int jend = lo < hi ? hi : lo;
  if (jend == k.qlen) {
    if (left >= k.gscore) k.end_i = k.i;
    k.gscore = left > k.gscore ? left : k.gscore;
  }
  if (lo >= hi) return true;
  const uint32_t klo = r.kmax & 0xFFFFu, khi = r.kmax >> 16;
  const bool take_hi = khi >= klo;                        // larger score, or the same score in the later column
  const uint32_t kk = take_hi ? khi : klo;
  const int m = (int)(kk >> COLS_KEY_BITS) - (int)c.bias;
  const int mj = c0 + 2 * (int)(kk & ((1u << COLS_KEY_BITS) - 1)) + (take_hi ? 1 : 0);
  if (m <= 0) return true;                                // ksw.c:454
  if (m > k.best) {
    k.best = m; k.best_i = k.i; k.best_j = mj;
    const int dd = mj > k.i ? mj - k.i : k.i - mj;
    k.max_off = dd > k.max_off ? dd : k.max_off;
  } else if (sc.zdrop > 0) {
    const int di = k.i - k.best_i, dj = mj - k.best_j;
    if (di > dj) { if (k.best - m - (di - dj) * sc.e_del > sc.zdrop) return true; }
    else if (k.best - m - (dj - di) * sc.e_ins > sc.zdrop) return true;
  }
  // ksw.c:466-469.  Left end: a lower bound of the reference's beg, advanced by whole pairs; right end: exact.
  {
    const uint2 w0 = st[(size_t)(c0 >> 1) * 32];
    if (w0.x == c.B2 && w0.y == c.B2) lo = c0 + 2 < hi ? c0 + 2 : hi;
  }
  int j = hi;
  if (left == 0) {                                        // eh[hi] is all zero
    const uint16_t z = (uint16_t)c.bias;
    --j;
    while (j >= lo) {
      const uint16_t *q = (const uint16_t *)(st + (size_t)(j >> 1) * 32) + (j & 1);
      if (q[0] != z || q[2] != z) break;
      --j;
    }
  }
  k.lo = lo;
  k.hi = j + 2 < k.qlen ? j + 2 : k.qlen;
  return ++k.i >= k.tlen;
}

} // namespace bwaflow
