// One-lane banded extension engine: exact ksw_extend2 semantics (bwa/ksw.c:380-479), one DP problem
// per thread, advanced ONE ROW per call so that all lanes of a warp stay convergent at row boundaries
// while each walks its own [lo,hi) column window.
//
// Row state lives in shared memory, one 32-bit word per (column, thread), laid out [column][thread] so
// that a warp's accesses hit 32 distinct banks whatever column each lane is at:
//     bits  0..13  H(i-1, j-1)      (eh[j].h of the reference)
//     bits 14..16  query base j     (0..3, 4 = N)   -- folded in so a cell needs a single LDS
//     bits 17..30  E(i, j)          (eh[j].e)
// Scores therefore must stay below 2^14; begin() reports jobs that could exceed it (h0 + qlen*max_mat).
//
// Arithmetic is int32 on the DPX paths: VIMNMX3 for h = max(M,e,f) and VIADDMNMX.RELU for
// e' = max(e-e_del, M-oe_del, 0), f' likewise.  The row maximum and "last column attaining it"
// (ksw.c:437-438) are tracked as one key (h<<16 | j) under a plain max.  The substitution score comes from a
// sign-extending byte permute of the target row of the matrix held in two registers.
#pragma once
#include "internal.cuh"

namespace bwaflow {

constexpr uint32_t KSW_HMASK = (1u << 14) - 1;
constexpr uint32_t KSW_QMASK = 7u << 14;
constexpr uint32_t KSW_HEMASK = ~KSW_QMASK & 0x7FFFFFFFu;      // H and E fields
constexpr int KSW_ESHIFT = 17;
constexpr int KSW_MAX_SCORE = (1 << 14) - 1;

// byte q of {hi:lo} sign-extended to 32 bits: selector nibbles q, q|8, q|8, q|8 (msb of a nibble = replicate the sign)
__device__ __forceinline__ int prmt_sx(uint32_t lo, uint32_t hi, uint32_t sel)
{
  int r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
}
// The integer ALU pipe (LOP3/SHF/PRMT and the DPX min/max forms) is what saturates in the cell loop, so shifts and adds
// are expressed as multiply-adds, which issue on the FMA pipe.
__device__ __forceinline__ uint32_t shr_fma(uint32_t v, int sh) { return __umulhi(v, 1u << (32 - sh)); }   // v >> sh
__device__ __forceinline__ int add_fma(int a, int b)
{
  int r;
  asm("mad.lo.s32 %0, %1, 1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// (int)((double)x / e + 1.) floored at 1, for e > 0 -- integer-exact form of ksw.c:402-406 / bwamem.c:632-635
__device__ __forceinline__ int gap_limit(int x, int e) { return x >= e ? x / e + 1 : 1; }

struct KswLane {
  int qlen, tlen, w, h0;
  int i, lo, hi;
  int best, best_i, best_j, end_i, gscore, max_off;
  unsigned cells;              // DP cells executed (ksw.c:424-448 iterations) since begin()

  // Sets up row -1 (ksw.c:395-397) and the clamped band (ksw.c:399-407).  qf(j) yields query base j.
  // Returns false when the job cannot be represented in the packed row state.
  template <class QF>
  __device__ __forceinline__ bool begin(uint32_t *col, int stride, const DevScore &sc, int qlen_, int tlen_, int w_,
                                        int h0_, int end_bonus, QF qf)
  {
    qlen = qlen_; tlen = tlen_; h0 = h0_;
    int lim = gap_limit(qlen * sc.smax + end_bonus - sc.o_ins, sc.e_ins);
    w = w_ < lim ? w_ : lim;
    lim = gap_limit(qlen * sc.smax + end_bonus - sc.o_del, sc.e_del);
    w = w < lim ? w : lim;
    i = 0; lo = 0; hi = qlen;
    best = h0; best_i = best_j = -1; end_i = -1; gscore = -1; max_off = 0; cells = 0;
    if (h0 + qlen * sc.smax > KSW_MAX_SCORE || h0 <= 0) return false;
    const int gi = sc.o_ins + sc.e_ins;
    uint32_t *p = col;
    int hv = h0;                                         // H(-1,-1)
    for (int j = 0; j < qlen; ++j, p += stride) {
      *p = (uint32_t)qf(j) << 14 | (uint32_t)hv;
      hv = h0 - gi - j * sc.e_ins;                       // H(-1, j)
      hv = hv > 0 ? hv : 0;
    }
    *p = (uint32_t)hv;
    return true;
  }

  // One DP row with target base t.  Returns true when the extension is finished.
  __device__ __forceinline__ bool row(uint32_t *col, int stride, const DevScore &sc, int t)
  {
    const int gd = sc.o_del + sc.e_del, gi = sc.o_ins + sc.e_ins;
    if (lo < i - w) lo = i - w;
    if (hi > i + w + 1) hi = i + w + 1;
    if (hi > qlen) hi = qlen;
    int left = 0;
    if (lo == 0) { left = h0 - (sc.o_del + sc.e_del * (i + 1)); left = left > 0 ? left : 0; }
    const uint32_t rlo = sc.row_lo[t], rhi = sc.row_hi[t];
    int f = 0, mkey = -1;
    uint32_t *p = col + lo * stride;
    for (int j = lo; j < hi; ++j, p += stride) {
      const uint32_t wv = *p;
      const int s = prmt_sx(rlo, rhi, (shr_fma(wv, 14) & 7u) * 0x1111u + 0x8880u);   // score of (target row, query base)
      const int hd = (int)(wv & KSW_HMASK);
      const int e = (int)shr_fma(wv, KSW_ESHIFT);
      // M = hd ? hd + s : 0 (ksw.c:433).  Any M <= 0 is equivalent to 0 below (e, f >= 0 and the relu), and
      // min(hd + s, 128*hd) is <= 0 when hd == 0 and equals hd + s otherwise (|s| <= 127): one DPX op, no compare/select
      const int M = __viaddmin_s32(hd, s, hd * 128);
      const int h = __vimax3_s32(M, e, f);
      const int e2 = __viaddmax_s32_relu(M, -gd, add_fma(e, -sc.e_del));
      f = __viaddmax_s32_relu(M, -gi, add_fma(f, -sc.e_ins));
      mkey = max(mkey, h * 65536 + j);
      *p = ((wv & KSW_QMASK) | (uint32_t)left) + (uint32_t)e2 * (1u << KSW_ESHIFT);
      left = h;
    }
    const int jend = lo < hi ? hi : lo;
    cells += (unsigned)(jend - lo);
    p = col + hi * stride;
    *p = (*p & KSW_QMASK) | (uint32_t)left;              // eh[end] = {h1, 0}   ksw.c:449
    if (jend == qlen) {                                  // ksw.c:450-453
      if (left >= gscore) end_i = i;
      gscore = left > gscore ? left : gscore;
    }
    if (mkey < 0x10000) return true;                     // m == 0 (or empty row)   ksw.c:454
    const int m = mkey >> 16, mj = mkey & 0xffff;
    if (m > best) {                                      // ksw.c:455-457
      best = m; best_i = i; best_j = mj;
      const int d = mj > i ? mj - i : i - mj;
      max_off = d > max_off ? d : max_off;
    } else if (sc.zdrop > 0) {                           // ksw.c:458-464
      const int di = i - best_i, dj = mj - best_j;
      if (di > dj) { if (best - m - (di - dj) * sc.e_del > sc.zdrop) return true; }
      else if (best - m - (dj - di) * sc.e_ins > sc.zdrop) return true;
    }
    int j = lo;                                          // ksw.c:466-469
    p = col + lo * stride;
    while (j < hi && (*p & KSW_HEMASK) == 0) { ++j; p += stride; }
    lo = j;
    j = hi; p = col + hi * stride;
    while (j >= lo && (*p & KSW_HEMASK) == 0) { --j; p -= stride; }
    hi = j + 2 < qlen ? j + 2 : qlen;
    return ++i >= tlen;
  }
};

} // namespace bwaflow
