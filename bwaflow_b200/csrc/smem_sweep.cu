// Stage 2, the production seeding kernel: passes 1 and 2 of mem_collect_intv_new (src/bwa_wrapper.cpp:207-226) for reads
// shorter than 256 bases, one read per lane.
//
// What bwt_smem1a_new (src/bwa_wrapper.cpp:325-450) computes.  Its forward sweep from x records the ends e_0 < e_1 < .. at which
// the interval of q[x, e) changes size.  Its backward loop then visits these forward matches shortest first and, for
// each, finds how far the match extends to the left with an interval of at least min_intv rows -- either by extending
// it base by base ("back-enlarge", :371-386) or by re-extending the saved leftward trail of an earlier match to the right
// ("forward-enlarge", :403-417, which tries the trail entries longest first and takes the first that survives).  Interval
// sizes only shrink when a pattern grows on either side, so both routes return the same thing: the leftmost begin b with
// |I(q[b, e_i))| >= min_intv (not crossing an N or the read start), together with the interval of q[b, e_i) -- which is a
// property of the pattern, not of the route.  Which route the reference takes (its `stop - start >= 3` rule) changes its
// cost, never its output.  The candidate/emit logic (:392-400, :418-427, :431-443) consumes only (b, e_i, interval).
//
// This kernel computes exactly that function per forward match, by the cheapest exact means:
//   T  patterns of at most K bases: one 16-byte read of the level tables (ctx.cu; intervals of all 1..K-mers);
//   X  longer, non-unique patterns: bwt_extend on the occ32 index (fmindex.cuh);
//   A  unique intervals with min_intv == 1 (template TEXT, needs the full suffix array): anchored in the 2-bit text, extended
//      by comparing up to 32 bases per step; emitted as {text position, TEXT_MARK, 1} (the hot path needs only positions).
// A short forward match [x, e) starts from the table entry of the K bases ending at e (its begin cannot lie to the left of the
// previous match's begin, another consequence of monotonicity, so the window is clipped there) and is extended to the left on the
// index; if that window is already too rare, the begin is found by walking the table levels down.  A long forward match
// starts from its own stored interval.  Per forward match this is ~1 table read and 2-3 index steps instead of the
// reference's ~15-30 bwt_extend calls; the reference's call count is still what the roofline is stated in, and is
// produced by the exact step-by-step instances (smem_fast.cu, COUNT) which replay the reference's route.
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "batches.cuh"
#include "fmindex.cuh"
#include "smem_args.cuh"
#include "smem_rows.cuh"

namespace bwaflow {

enum { SK_T = 0, SK_X = 1, SK_A = 2 };
enum { SM_F = 0, SM_W = 1, SM_S = 2, SM_B = 3, SM_C = 4, SM_CF = 5, SM_J = 6 };   // forward sweep | window lookup | level walk | leftward extension | continue the previous anchored match | (that failed) | forward sweep: all remaining table levels in one read
enum { S_NONE = 0, S_FETCH, S_NEXT, S_ENTRY, S_ENTRY_DONE, S_CALL_DONE, S_FINISH, S_EXIT, S_DEFER };

#ifndef SWEEP_MIN_BLOCKS
#define SWEEP_MIN_BLOCKS 6
#endif

template <int MAXL, bool TEXT>
__global__ void __launch_bounds__(128, SWEEP_MIN_BLOCKS) smem_sweep_kernel(const SmemArgs a)
{
  using R = Row<MAXL>;
  extern __shared__ uint64_t s_rows[];
  LaneRead<MAXL> q;
  q.bind(s_rows + (size_t)threadIdx.x * (R::ROW_BYTES / 8));
  uint8_t *fend = (uint8_t *)q.q2 + R::PACK_BYTES;           // ends of the forward matches (first 64; later ones are stored entries)
  uint4 fwd_l[MAXL];                                         // intervals of forward matches longer than K
  const DevIndex &ix = a.ix;
  const unsigned full = 0xffffffffu;
  const int msl = a.min_seed_len, K = ix.ktab ? ix.ktab_k : 0;

  int ts = S_FETCH, kind = SK_T, mode = SM_F;
  int64_t r = -1; bwaflow_intv_t *mem = nullptr; int n_mem = 0;
  int pass = 1, x = 0, k2 = 0, old_n = 0;
  int sx = 0, mi = 1, nb = 0, fwd_n = 0, fwd_ns = 0, fi = 0, f_end = 0, lb = 0, sL = 0;
  int pb = 0, pe = 0;
  Intv ik = {0, 0, 0}, cand = {0, 0, 0}; int cand_beg = 0, cand_end = 0;

  // The forward sweep walks the table levels one base at a time only up to length JL; from there the level-K entry of the K bases
  // from sx gives the interval at length K, and its 11 history bits say at which of the levels in between the interval changed
  // size -- which is all the forward list keeps of them (ends only: patterns of at most K bases are re-read from the tables).
  const int JL = K >= 2 ? max(1, K - 11) : 0;
  int jm = 0;
  auto tab = [&](int b, int L) -> Intv { int d; return unpack_entry(__ldg(ix.ktab + ktab_base(L) + q.code(b, L)), &d); };
  auto get_end = [&](int k) -> int { return k < 64 ? (int)fend[k] : entry_pos(fwd_l[k - fwd_ns]); };
  auto push_fwd = [&](const Intv &v, int end) {
    if (fwd_n < 64) fend[fwd_n] = (uint8_t)end;
    if (end - sx > K) fwd_l[fwd_n - fwd_ns] = pack_entry(v, end); else ++fwd_ns;
    ++fwd_n;
  };
  auto push_mem = [&](const Intv &v, int beg, int end) {
    if (n_mem < a.cap) {
      bwaflow_intv_t o; o.x[0] = v.x0; o.x[1] = v.x1; o.x[2] = v.x2; o.info = (uint64_t)(uint32_t)beg << 32 | (uint32_t)end;
      mem[n_mem] = o;
    }
    ++n_mem;
  };
  // end of a forward sweep: on to its matches -- unless there are so many that the read is one for smem_heavy_kernel (S_DEFER)
  auto fwd_done = [&]() { fi = 0; lb = nb; ts = fwd_n >= a.heavy_min ? S_DEFER : S_ENTRY; };
  // kind of the next step for the pattern [pb, pe) growing by one base
  auto set_kind = [&]() { kind = (TEXT && is_text(ik)) ? SK_A : (pe - pb + 1 <= K ? SK_T : SK_X); ts = S_NONE; };
  // next step of a forward sweep standing at q[pb, pe): the jump if the K bases from pb are all there
  auto set_kind_fwd = [&]() { set_kind(); if (JL && pe - pb == JL && kind == SK_T && q.next_n(pb) >= pb + K) mode = SM_J; };

  for (;;) {
    // ---- the entry boundary (every ~4 steps of a lane): one straight-line block so that the lanes at a boundary run it together
    while (ts == S_ENTRY || ts == S_ENTRY_DONE) {
      if (ts == S_ENTRY_DONE) {                              // candidate logic (:392-400, :418-427) and the skip rule (:431-443)
        if (fi > 0 && pb > cand_beg && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
        cand = ik; cand_beg = pb; cand_end = f_end; lb = pb;
        ++fi;
        if (fi < fwd_n) {
          int stop = 0, max_len = cand_beg + get_end(fi);
          while (max_len < msl && fi < fwd_n) {
            ++fi;
            if (fi < fwd_n) stop = get_end(fi);
            max_len = cand_beg + stop;
          }
        }
        if (fi >= fwd_n && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
      }
      // forward match fi = q[sx, f_end): how far does it reach to the left?
      if (fi >= fwd_n) { ts = S_CALL_DONE; break; }
      f_end = get_end(fi);
      if (TEXT && fi > 0 && mode != SM_CF && is_text(cand) && f_end - cand_end <= 32) {
        // the previous match is one occurrence in the text and begins can only move right: if the read keeps matching the
        // text up to f_end, this match is the same occurrence with the same begin
        ik = cand; pb = cand_beg; pe = cand_end; mode = SM_C; kind = SK_A; ts = S_NONE;
        break;
      }
      pe = f_end;
      if (f_end - sx > K) {                                  // from its own interval
        int d;
        ik = unpack_entry(fwd_l[fi - fwd_ns], &d); pb = sx; mode = SM_B;
        if (pb > lb) set_kind(); else ts = S_ENTRY_DONE;
      } else {                                               // from the table entry of the (at most K) bases ending at f_end
        sL = min(K, f_end - lb); mode = SM_W; kind = SK_T; ts = S_NONE;
      }
    }
    // ---- everything else (once per call / per read)
    while (ts != S_NONE && ts != S_EXIT && ts != S_ENTRY) {
      switch (ts) {
      case S_FETCH: {
        r = (int64_t)atomicAdd(a.ctl, 1);
        if (r >= a.n_reads) { ts = S_EXIT; break; }
        const int len = (int)(a.read_off[r + 1] - a.read_off[r]);
        mem = a.slots + r * a.cap; n_mem = 0; pass = 1; x = 0;
        if (len >= MAXL) { atomicExch(a.ctl + 2, 1); a.counts[r] = 0; break; }
        q.load(a.packed, r, len);
        ts = len < msl ? S_FINISH : S_NEXT;                  // src/bwa_wrapper.cpp:150
        break;
      }
      case S_NEXT: {                                         // the call loops of mem_collect_intv_new (:214-226)
        bool call = false;
        if (pass == 1) {
          while (x < q.len && q.isN(x)) ++x;
          if (x < q.len) { sx = x; mi = 1; call = true; }
          else { pass = 2; old_n = n_mem < a.cap ? n_mem : a.cap; k2 = 0; }
        }
        if (pass == 2 && !call) {
          while (k2 < old_n) {
            const bwaflow_intv_t p = mem[k2++];
            const int s = (int)(p.info >> 32), e = (int)(uint32_t)p.info;
            if (e - s < a.split_len || p.x[2] > (uint64_t)a.split_width) continue;
            if (q.isN((s + e) >> 1)) continue;               // bwt_smem1a_new :338
            sx = (s + e) >> 1; mi = (int)p.x[2] + 1; call = true;
            break;
          }
          if (!call) pass = 3;
        }
        if (call) {                                          // forward sweep from sx (:338-363)
          fwd_n = 0; fwd_ns = 0; ik = set_intv(ix, q.base(sx)); pb = sx; pe = sx + 1; mode = SM_F;
          nb = q.prev_n(sx) + 1;
          if (pe < q.len && !q.isN(pe)) set_kind_fwd();
          else { push_fwd(ik, pe); fwd_done(); }
          break;
        }
        ts = S_FINISH;                                       // pass 3 has its own kernel
        break;
      }
      case S_DEFER:
        // A call with this many forward matches is a read inside a repeat: its intervals stay large along the whole read and the
        // leftward searches cost (matches x read length) index steps, one after the other on this lane -- tens of milliseconds
        // for a single read while the rest of the grid has long finished.  Such a read is handed, whole, to smem_heavy_kernel,
        // which spreads the searches of a call over a warp.
        a.defer[2 + atomicAdd(a.defer, 1)] = (int)r;
        a.counts[r] = 0;
        ts = S_FETCH;
        break;
      case S_CALL_DONE:
        if (pass == 1) x = get_end(fwd_n - 1);
        ts = S_NEXT;
        break;
      case S_FINISH:
        a.counts[r] = n_mem;
        if (n_mem > a.cap) atomicExch(a.ctl + 1, 1);
        ts = S_FETCH;
        break;
      }
    }
    __syncwarp();                                            // lanes leave the transition loop at different times: rejoin them here
    if (__ballot_sync(full, ts != S_EXIT) == 0) break;
    Intv o = {0, 0, 0}; bool have = false;
    // ---- T and A share ONE load site: two 8-byte words, of a level-table entry (T) or of the 2-bit text (A).  Lanes of either
    // kind issue their loads together instead of one mode after the other, each waiting out its own latency.
    const bool isT = ts == S_NONE && kind == SK_T, isA = TEXT && ts == S_NONE && kind == SK_A;
    uint64_t w0 = 0, w1 = 0, tp = 0;
    {
      const uint64_t *lp = nullptr;
      if (isT) {
        int tb, tL;
        if (mode == SM_J) { tb = pb; tL = K; }
        else if (mode == SM_F) { tb = pb; tL = pe - pb + 1; }
        else if (mode == SM_B) { tb = pb - 1; tL = pe - pb + 1; }
        else { tb = pe - sL; tL = sL; }                      // window / level walk: the sL bases ending at pe
        lp = (const uint64_t *)(ix.ktab + ktab_base(tL) + q.code(tb, tL));
      } else if (isA) {
        tp = mode == SM_B ? (ik.x0 >= 32 ? ik.x0 - 32 : 0) : ik.x0 + (uint64_t)(pe - pb);   // first text symbol of the window
        lp = ix.text + (tp >> 5);
      }
      if (lp) { w0 = __ldg(lp); w1 = __ldg(lp + 1); }
    }
    if (isT) {
      int d;
      o = unpack_entry(make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32)), &d);
      if (mode == SM_J) jm = d;
      have = true;
    }
    // ---- X: bwt_extend on the index once enough lanes want one
    const unsigned wx = __ballot_sync(full, ts == S_NONE && kind == SK_X);
    const unsigned wo = __ballot_sync(full, ts == S_NONE && kind != SK_X);
    if (wx && (__popc(wx) >= a.x_min || !wo)) {
      if (ts == S_NONE && kind == SK_X) {
        unsigned long long nbk = 0;
        const int c = mode == SM_B ? q.base(pb - 1) : 3 - q.base(pe);
        o = extend1(ix, ik, mode == SM_B, c, &nbk);
        have = true;
      }
    }
    if (have) {
      const bool small = o.x2 < (uint64_t)mi;
      if (TEXT && mi == 1 && o.x2 == 1) o = text_intv(ix.sa[o.x0]);       // unique from here on: anchor it in the text
      if (mode == SM_J) {
        mode = SM_F;
        if (small) set_kind();                               // some prefix is too rare: take the levels one by one
        else {
          for (int L = pe - pb; L < K; ++L)                  // ends at which the interval changed size (level L -> L + 1)
            if (jm >> (K - 1 - L) & 1) { fend[fwd_n++] = (uint8_t)(pb + L); ++fwd_ns; }
          ik = o; pe = pb + K;
          if (pe < q.len && !q.isN(pe)) set_kind(); else { push_fwd(ik, pe); fwd_done(); }
        }
      } else if (mode == SM_F) {                             // :350-363
        const bool changed = o.x2 != ik.x2;
        if (changed) push_fwd(ik, pe);
        if (changed && small) { fwd_done(); }
        else {
          ik = o; ++pe;
          if (pe < q.len && !q.isN(pe)) set_kind_fwd(); else { push_fwd(ik, pe); fwd_done(); }
        }
      } else if (mode == SM_B) {                             // one more base on the left (:374-386)
        if (small) ts = S_ENTRY_DONE;
        else {
          ik = o; --pb;
          if (pb > lb) set_kind(); else ts = S_ENTRY_DONE;
        }
      } else {                                               // SM_W / SM_S: is the window q[pe - sL, pe) frequent enough?
        if (!small || sL == pe - sx) {                       // (the forward match itself is taken as it is, :371)
          ik = o; pb = pe - sL;
          if (mode == SM_W && pb > lb) { mode = SM_B; set_kind(); } else ts = S_ENTRY_DONE;
        } else { --sL; mode = SM_S; }                        // walk the levels down: everything longer is too rare
      }
    }
    // ---- A: anchored interval: compare up to 32 bases of the read with the text (words loaded above)
    if (isA) {
      const int sh = (int)(tp & 31) << 1;
      const uint64_t twin = sh ? (w0 << sh | w1 >> (64 - sh)) : w0;                       // 32 text symbols from tp, MSB first
      if (mode == SM_B) {
        int cap = pb - lb;                                   // read bases available on the left (>= 1)
        if ((uint64_t)cap > ik.x0) cap = (int)ik.x0;         // ... and text symbols
        const int want = min(cap, 32);
        int n = 0;
        if (want > 0) {
          // the 32 symbols before x0, the last one at bit 0 (fewer than 32: the top of word 0 shifted down)
          const uint64_t tend = ik.x0 >= 32 ? twin : (ik.x0 > 0 ? w0 >> (64 - 2 * (int)ik.x0) : 0ull);
          const uint64_t diff = q.win_end(pb) ^ tend;
          n = min(diff ? (__ffsll((long long)diff) - 1) >> 1 : 32, want);
        }
        pb -= n; ik.x0 -= (uint64_t)n;
        if (!(n == 32 && cap > 32)) ts = S_ENTRY_DONE;
      } else if (mode == SM_C) {                             // does the previous occurrence also cover [cand_end, f_end)?
        const int want = f_end - pe;
        const uint64_t diff = q.win(pe) ^ twin;
        const int n = diff ? __clzll((long long)diff) >> 1 : 32;
        if (n >= want && tp + (uint64_t)want <= ix.seq_len) { pe = f_end; ts = S_ENTRY_DONE; }
        else { mode = SM_CF; ts = S_ENTRY; }
      } else {                                               // forward sweep on a unique interval: runs to the first mismatch
        const int lim = q.next_n(pe);
        const int want = lim - pe;
        const int64_t tcap = (int64_t)ix.seq_len - (int64_t)tp;
        const int w = (int)min((int64_t)min(want, 32), max(tcap, (int64_t)0));
        int n = 0;
        if (w > 0) {
          const uint64_t diff = q.win(pe) ^ twin;
          n = min(diff ? __clzll((long long)diff) >> 1 : 32, w);
        }
        pe += n;
        if (!(n == 32 && want > 32 && tcap > 32)) { push_fwd(ik, pe); fwd_done(); }
      }
    }
  }
}

// Passes 1 and 2 of mem_collect_intv_new for the reads the sweep kernel handed over, ONE WARP PER READ.  Per call of
// bwt_smem1a_new (src/bwa_wrapper.cpp:325-450): the forward sweep (:338-363) step by step on the index, every lane the same; then
// the leftmost begin of every forward match -- the interval of q[b, e_i) has at least min_intv rows, b does not cross an N or the
// read start -- 32 matches at a time, one per lane, each by plain leftward bwt_extend steps from the match's own interval (these
// searches are independent: the bound the sweep kernel takes from the previous match only saves steps, interval sizes being
// monotone in the pattern); then the candidate / emit logic (:392-400, :418-427, :431-443) over the 32 results in order.  No table
// or text shortcuts: the reads that come here spend their time on large intervals.  Intervals leave as BWT rows.
template <int MAXL>
__global__ void __launch_bounds__(128) smem_heavy_kernel(const SmemArgs a)
{
  using R = Row<MAXL>;
  constexpr int WROW = R::ROW_BYTES / 8;
  extern __shared__ uint64_t s_heavy[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  constexpr int WSTRIDE = (WROW + 1) + 2 * MAXL + 2 * 32;      // row (padded to 16 bytes) + MAXL forward matches + 32 results, in 8-byte words
  uint64_t *wbase = s_heavy + (size_t)wib * WSTRIDE;
  LaneRead<MAXL> q;
  q.bind(wbase);
  uint4 *fwd = (uint4 *)(wbase + WROW + 1);                    // forward matches of the call: packed interval + end
  uint4 *res = fwd + MAXL;                                     // results of the 32 matches in flight: packed interval + begin
  const DevIndex &ix = a.ix;
  const int msl = a.min_seed_len;
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(a.defer + 1, 1);
    slot = __shfl_sync(full, slot, 0);
    if (slot >= a.defer[0]) break;
    const int64_t r = a.defer[2 + slot];
    const int len = (int)(a.read_off[r + 1] - a.read_off[r]);
    bwaflow_intv_t *mem = a.slots + r * a.cap;
    int n_mem = 0;
    __syncwarp();
    if (lane == 0) q.load(a.packed, r, len);
    __syncwarp();
    q.len = len;
    { uint32_t any = 0; for (int t = 0; t < R::W; ++t) any |= q.nm[t]; q.anyN = any != 0; }
    auto push_mem = [&](const Intv &v, int beg, int end) {
      if (lane == 0 && n_mem < a.cap) {
        bwaflow_intv_t o; o.x[0] = v.x0; o.x[1] = v.x1; o.x[2] = v.x2; o.info = (uint64_t)(uint32_t)beg << 32 | (uint32_t)end;
        mem[n_mem] = o;
      }
      ++n_mem;
    };
    // one call of bwt_smem1a_new from sx with min_intv mi; returns the end of the longest forward match
    auto call = [&](int sx, int mi) -> int {
      unsigned long long nbk = 0;
      int n = 0, pe = sx + 1;
      Intv ik = set_intv(ix, q.base(sx));
      for (;;) {                                               // forward sweep, :338-363
        if (pe >= len || q.isN(pe)) { if (lane == 0) fwd[n] = pack_entry(ik, pe); ++n; break; }
        const Intv o = extend1(ix, ik, 0, 3 - q.base(pe), &nbk);
        const bool changed = o.x2 != ik.x2;
        if (changed) { if (lane == 0) fwd[n] = pack_entry(ik, pe); ++n; }
        if (changed && o.x2 < (uint64_t)mi) break;
        ik = o; ++pe;
      }
      __syncwarp();
      const int nb = q.prev_n(sx) + 1;
      int fi = 0, cand_beg = 0, cand_end = 0, last_end = 0;
      Intv cand = {0, 0, 0};
      for (int base = 0; base < n; base += 32) {
        const int my = base + lane;
        int pb = sx, e_my = 0;
        Intv im = {0, 0, 0};
        bool go = false;
        if (my < n) { im = unpack_entry(fwd[my], &e_my); go = pb > nb; }
        while (__any_sync(full, go)) {                         // leftward extension of 32 matches, one index step each per round
          if (go) {
            const Intv o = extend1(ix, im, 1, q.base(pb - 1), &nbk);
            if (o.x2 < (uint64_t)mi) go = false;
            else { im = o; --pb; go = pb > nb; }
          }
        }
        if (my < n) res[lane] = pack_entry(im, pb);
        __syncwarp();
        const int lim = n - base < 32 ? n - base : 32;
        for (int u = 0; u < lim; ++u) {                        // the matches in order, every lane alike
          const int j = base + u;
          if (j < fi) continue;                                // passed over by the skip rule below
          int pbj, f_end;
          const Intv ij = unpack_entry(res[u], &pbj);
          f_end = entry_pos(fwd[j]);
          if (fi > 0 && pbj > cand_beg && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
          cand = ij; cand_beg = pbj; cand_end = f_end;
          ++fi;
          if (fi < n) {
            int stop = 0, max_len = cand_beg + entry_pos(fwd[fi]);
            while (max_len < msl && fi < n) {
              ++fi;
              if (fi < n) stop = entry_pos(fwd[fi]);
              max_len = cand_beg + stop;
            }
          }
          if (fi >= n && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
        }
        __syncwarp();
      }
      last_end = entry_pos(fwd[n - 1]);
      __syncwarp();
      return last_end;
    };
    if (len >= msl && len < MAXL) {
      int x = 0;                                               // pass 1, :214-218
      while (x < len) {
        while (x < len && q.isN(x)) ++x;
        if (x >= len) break;
        x = call(x, 1);
      }
      const int old_n = n_mem < a.cap ? n_mem : a.cap;         // pass 2, :220-226
      __syncwarp();
      for (int k2 = 0; k2 < old_n; ++k2) {
        const bwaflow_intv_t p = mem[k2];
        const int s = (int)(p.info >> 32), e = (int)(uint32_t)p.info;
        if (e - s < a.split_len || p.x[2] > (uint64_t)a.split_width) continue;
        if (q.isN((s + e) >> 1)) continue;
        call((s + e) >> 1, (int)p.x[2] + 1);
      }
    }
    if (lane == 0) {
      a.counts[r] = n_mem;
      if (n_mem > a.cap) atomicExch(a.ctl + 1, 1);
    }
  }
}

template <int MAXL>
static int sweep_launch_t(const SmemArgs &a, int grid, int heavy_grid, bool text, cudaStream_t st)
{
  const size_t sh = 128 * (size_t)Row<MAXL>::ROW_BYTES;
  if (text) smem_sweep_kernel<MAXL, true><<<grid, 128, sh, st>>>(a);
  else smem_sweep_kernel<MAXL, false><<<grid, 128, sh, st>>>(a);
  // the reads handed over (none on most batches: the warps find an empty queue and leave)
  const size_t shh = 4 * (size_t)(Row<MAXL>::ROW_BYTES / 8 + 1 + 2 * MAXL + 2 * 32) * 8;
  BF_CUDA(cudaFuncSetAttribute(smem_heavy_kernel<MAXL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shh));
  smem_heavy_kernel<MAXL><<<heavy_grid, 128, shh, st>>>(a);
  BF_CUDA(cudaGetLastError());
  return 0;
}

// passes 1-2 for a batch of packed reads shorter than 256 bases (pack and pass 3 are launched by smem_fast_launch)
int smem_sweep_launch(bwaflow_smem_batch *b, const SmemArgs &a, bool text)
{
  const int64_t want = (b->n_reads + 127) / 128;
  const int64_t cap = (int64_t)b->ctx->n_sm * b->sweep_blocks_per_sm;
  const int grid = (int)(want < cap ? want : cap);
  const int hg = b->ctx->n_sm * 4;
  return b->max_len < 160 ? sweep_launch_t<160>(a, grid, hg, text, b->stream) : sweep_launch_t<256>(a, grid, hg, text, b->stream);
}

int smem_sweep_setup(bwaflow_smem_batch *b)
{
  int per_sm = 0;
  BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smem_sweep_kernel<160, true>, 128, 128 * Row<160>::ROW_BYTES));
  if (per_sm < 1) per_sm = 1;
  if (const char *s = getenv("BWAFLOW_SWEEP_BLOCKS_PER_SM")) per_sm = atoi(s) > 0 ? atoi(s) : per_sm;
  b->sweep_blocks_per_sm = per_sm;
  return 0;
}

} // namespace bwaflow
