// Stage 2, fast kernels: mem_collect_intv_new (src/bwa_wrapper.cpp:207-246) for reads shorter than 256 bases with the same
// per-lane flow machine as smem.cu, but an interval is obtained by the cheapest of three exact means:
//   T  pattern length <= K: ONE 16-byte read of the level tables (interval of every 1..K-mer, ctx.cu); the lookups of a
//      sweep do not depend on each other and the levels up to ~11 stay in L2, so the ~30 % of bwt_extend calls
//      bwt_smem1a_new (src/bwa_wrapper.cpp:325-450) makes on short patterns cost no index traffic;
//   X  longer, non-unique pattern: bwt_extend on the occ32 index as before (fmindex.cuh);
//   A  unique interval while min_intv == 1 (template TEXT): the occurrence is anchored in the 2-bit text through the
//      full suffix array, and extending it is a word compare of read against text, up to 32 bases per step (another
//      ~35 % of the calls).  bwt_extend on a one-row interval yields one row iff the neighbouring text symbol equals
//      the base, which is what the compare decides.  Anchored intervals are emitted as {text position, TEXT_MARK, 1}:
//      the fused hot path needs only the position (no suffix-array read later); the interval API (rows wanted) runs
//      the TEXT = false instance.  Every back-trail entry of an anchored run is the same occurrence, so one failed
//      forward-enlarge compare retires the whole run (the reference repeats the identical failure per entry).
// The read is staged as 2-bit codes + an N mask (pack_reads_kernel), the forward list keeps one byte per entry for
// patterns the tables can re-derive, and only the few longer entries (typically 3-4) touch local memory.
// COUNT instances (TEXT off) add up the reference's occ-block touches exactly as oracle.c does.
#include <cub/device/device_scan.cuh>
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "batches.cuh"
#include "fmindex.cuh"
#include "smem_args.cuh"
#include "smem_rows.cuh"

namespace bwaflow {

// Level L of the tables from level L - 1: one thread per child, entry(P.c) = bwt_extend(entry(P)) forwards by c.
__global__ void ktab_level_kernel(DevIndex ix, int L, uint4 *tab)
{
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >> (2 * L)) return;
  Intv o;
  int hist = 0;
  if (L == 1) o = set_intv(ix, (int)v);
  else {
    const Intv p = unpack_entry(tab[ktab_base(L - 1) + (v >> 2)], &hist);
    unsigned long long nb = 0;
    if (p.x2 == 0) { o.x0 = o.x1 = o.x2 = 0; }
    else o = extend1(ix, p, 0, 3 - (int)(v & 3), &nb);
    // the entry's spare 11 bits: did the interval change size at this level and at the ten before it (bit j = level L - j)?
    // Lets the sweep kernel rebuild the forward list of the levels it jumps over (smem_sweep.cu, SM_J).
    hist = (hist << 1 | (o.x2 != p.x2)) & 0x7ff;
  }
  tab[ktab_base(L) + v] = pack_entry(o, hist);
}

int build_level_tables(bwaflow_ctx *c, int K, void **d_tab)
{
  const uint64_t n = ktab_base(K + 1);
  BF_CUDA(cudaMalloc(d_tab, n * 16));
  DevIndex ix = c->ix;
  ix.occ_cs &= ~2;
  for (int L = 1; L <= K; ++L) {
    const uint64_t m = 1ull << (2 * L);
    ktab_level_kernel<<<(unsigned)((m + 255) / 256), 256>>>(ix, L, (uint4 *)*d_tab);
    BF_CUDA(cudaGetLastError());
  }
  BF_CUDA(cudaDeviceSynchronize());
  return 0;
}

enum { KT = 0, KX = 1, KA = 2 };
enum { F_NONE = 0, F_FETCH, F_NEXT, F_F_DONE, F_B_SELECT, F_BE_DONE, F_FE_PICK, F_FE_FAIL, F_FE_REACHED, F_B_NEXT, F_CALL_DONE,
       F_FINISH, F_EXIT };
enum { P_F = 0, P_BE = 1, P_FE = 2 };

#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS 6
#endif

template <int MAXL, bool TEXT, bool COUNT>
__global__ void __launch_bounds__(128, FAST_MIN_BLOCKS) smem_fast_kernel(const SmemArgs a)
{
  static_assert(!(TEXT && COUNT), "block touches cannot be counted for text-anchored steps");
  using R = Row<MAXL>;
  extern __shared__ uint64_t s_rows[];
  uint64_t *q2 = s_rows + (size_t)threadIdx.x * (R::ROW_BYTES / 8);
  const uint32_t *nm = (const uint32_t *)(q2 + R::QW);
  uint8_t *fend = (uint8_t *)q2 + R::PACK_BYTES;
  uint4 fwd_l[MAXL], back_l[MAXL];                 // entries the tables cannot re-derive (pattern longer than K)
  const DevIndex &ix = a.ix;
  const unsigned full = 0xffffffffu;
  const int msl = a.min_seed_len, K = ix.ktab ? ix.ktab_k : 0;
  constexpr int BIG = 1 << 30;
  unsigned long long blocks = 0;

  int ts = F_FETCH, kind = KT;
  int64_t r = -1; int len = 0; bool anyN = false; bwaflow_intv_t *mem = nullptr; int n_mem = 0;
  int pass = 1, x = 0, k2 = 0, old_n = 0;
  int sx = 0, mi = 1, fwd_n = 0, fwd_ns = 0, fi = 0, start = 0, stop = 0, ret = 0, bk = 0, f_end = 0;
  int back_n = 0, back_s0 = 0, bt0 = BIG; uint64_t bta = 0;
  int phase = P_F, pb = 0, pe = 0;
  Intv ik = {0, 0, 0}, cand = {0, 0, 0}; int cand_beg = 0, cand_end = 0;

  auto base = [&](int i) -> int { return (int)(q2[i >> 5] >> ((~i & 31) << 1)) & 3; };
  auto isN = [&](int i) -> bool { return anyN && ((nm[i >> 5] >> (i & 31)) & 1u); };
  auto qwin = [&](int i) -> uint64_t {                        // 32 read bases from i, MSB first
    const int s = (i & 31) << 1; const uint64_t w0 = q2[i >> 5], w1 = q2[(i >> 5) + 1];
    return s ? (w0 << s | w1 >> (64 - s)) : w0;
  };
  auto qwin_end = [&](int i) -> uint64_t { return i >= 32 ? qwin(i - 32) : (i > 0 ? q2[0] >> (64 - 2 * i) : 0ull); };   // bases [i-32, i), last at bit 0
  auto twin = [&](uint64_t p) -> uint64_t {
    const int s = (int)(p & 31) << 1; const uint64_t w0 = __ldg(ix.text + (p >> 5)), w1 = __ldg(ix.text + (p >> 5) + 1);
    return s ? (w0 << s | w1 >> (64 - s)) : w0;
  };
  auto twin_end = [&](uint64_t p) -> uint64_t { return p >= 32 ? twin(p - 32) : (p > 0 ? __ldg(ix.text) >> (64 - 2 * (int)p) : 0ull); };
  auto next_n = [&](int i) -> int {                           // first N at or after i, else len
    for (int w = i >> 5; w < R::W; ++w) {
      uint32_t m = nm[w];
      if (w == (i >> 5)) m &= ~0u << (i & 31);
      if (m) return min(len, w * 32 + __ffs(m) - 1);
    }
    return len;
  };
  auto prev_n = [&](int i) -> int {                           // last N before i (i >= 1), else -1
    for (int w = (i - 1) >> 5; w >= 0; --w) {
      uint32_t m = nm[w];
      if (w == ((i - 1) >> 5)) m &= 0xffffffffu >> (31 - ((i - 1) & 31));
      if (m) return w * 32 + 31 - __clz(m);
    }
    return -1;
  };
  auto tab = [&](int b, int L) -> Intv {                      // interval of q[b, b+L), L <= K
    const uint64_t code = qwin(b) >> (64 - 2 * L);
    int d;
    return unpack_entry(__ldg(ix.ktab + ktab_base(L) + code), &d);
  };
  auto cross = [&](const Intv &v, int is_back) -> int {       // occ blocks bwt_extend touches from this interval (bwa/bwt.c:194)
    const uint64_t xf = is_back ? v.x0 : v.x1, k = xf - 1, l = k + v.x2;
    const uint64_t _k = k - (k >= ix.primary), _l = l - (l >= ix.primary);
    return (_k >> 7) != (_l >> 7) ? 2 : 1;
  };
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
  auto take_cand = [&](int beg, bool first) {                // src/bwa_wrapper.cpp:392-400 / 418-427
    if (!first && beg > cand_beg && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
    cand = ik; cand_beg = beg; cand_end = f_end;
  };
  auto set_kind = [&]() { kind = (TEXT && is_text(ik)) ? KA : (pe - pb + 1 <= K ? KT : KX); ts = F_NONE; };

  for (;;) {
    // ---- transitions: until this lane has a step pending (or no reads are left)
    while (ts != F_NONE && ts != F_EXIT) {
      switch (ts) {
      case F_FETCH: {
        r = (int64_t)atomicAdd(a.ctl, 1);
        if (r >= a.n_reads) { ts = F_EXIT; break; }
        len = (int)(a.read_off[r + 1] - a.read_off[r]);
        mem = a.slots + r * a.cap; n_mem = 0; pass = 1; x = 0;
        if (len >= MAXL) { atomicExch(a.ctl + 2, 1); a.counts[r] = 0; break; }
        const uint64_t *src = a.packed + r * (R::PACK_BYTES / 8);
#pragma unroll
        for (int t = 0; t < R::PACK_BYTES / 8; ++t) q2[t] = __ldg(src + t);
        uint32_t any = 0;
#pragma unroll
        for (int t = 0; t < R::W; ++t) any |= nm[t];
        anyN = any != 0;
        ts = len < msl ? F_FINISH : F_NEXT;                  // src/bwa_wrapper.cpp:150
        break;
      }
      case F_NEXT: {                                         // the loops of mem_collect_intv_new (:214-226); pass 3 has its own kernel
        bool call = false;
        if (pass == 1) {
          while (x < len && isN(x)) ++x;
          if (x < len) { sx = x; mi = 1; call = true; }
          else { pass = 2; old_n = n_mem < a.cap ? n_mem : a.cap; k2 = 0; }
        }
        if (pass == 2 && !call) {
          while (k2 < old_n) {
            const bwaflow_intv_t p = mem[k2++];
            const int s = (int)(p.info >> 32), e = (int)(uint32_t)p.info;
            if (e - s < a.split_len || p.x[2] > (uint64_t)a.split_width) continue;
            if (isN((s + e) >> 1)) continue;                 // bwt_smem1a_new :338
            sx = (s + e) >> 1; mi = (int)p.x[2] + 1; call = true;
            break;
          }
          if (!call) pass = 3;
        }
        if (call) {                                          // bwt_smem1a_new entry (:338-346)
          fwd_n = 0; fwd_ns = 0; back_n = 0; ik = set_intv(ix, base(sx)); pb = sx; pe = sx + 1; phase = P_F;
          if (pe < len && !isN(pe)) set_kind();
          else { push_fwd(ik, pe); ts = F_F_DONE; }
          break;
        }
        ts = F_FINISH;
        break;
      }
      case F_F_DONE:                                         // :364-366
        ret = get_end(fwd_n - 1); start = sx; stop = sx; fi = 0;
        ts = F_B_SELECT;
        break;
      case F_B_SELECT: {                                     // :367-371
        if (fi >= fwd_n) { ts = F_CALL_DONE; break; }
        f_end = get_end(fi);
        if (back_n == 0 || stop - start >= 3) {              // back-enlarge from the forward match [sx, f_end)
          int d;
          ik = fi < fwd_ns ? tab(sx, f_end - sx) : unpack_entry(fwd_l[fi - fwd_ns], &d);
          pb = sx; pe = f_end; phase = P_BE;
          back_s0 = max(0, K - (f_end - sx) + 1); back_n = 1; bt0 = BIG;
          if (TEXT && is_text(ik)) { bt0 = 0; bta = ik.x0; }
          else if (back_s0 == 0) back_l[0] = pack_entry(ik, 0);
          if (pb > 0 && !isN(pb - 1)) set_kind(); else ts = F_BE_DONE;
        } else { stop = f_end; bk = back_n - 1; ts = F_FE_PICK; }   // forward-enlarge the saved trail
        break;
      }
      case F_BE_DONE:                                        // :387-400
        start = f_end; stop = fi == fwd_n - 1 ? len : get_end(fi + 1);
        take_cand(pb, fi == 0);
        ts = F_B_NEXT;
        break;
      case F_FE_FAIL:
        --bk;
        // fall through
      case F_FE_PICK: {                                      // :405-408: trail entry bk = q[sx-bk, start)
        if (bk < 0) { ts = F_B_NEXT; break; }
        pb = sx - bk; pe = start; phase = P_FE;
        if (TEXT && bk >= bt0) { ik = text_intv(bta - (uint64_t)(bk - bt0)); kind = KA; ts = F_NONE; }
        else if (bk >= back_s0) { int d; ik = unpack_entry(back_l[bk - back_s0], &d); set_kind(); }
        else if (COUNT || pe - pb + 1 > K) { ik = tab(pb, pe - pb); set_kind(); }
        else { ik.x1 = 0; kind = KT; ts = F_NONE; }          // the table step needs no source interval
        break;
      }
      case F_FE_REACHED:                                     // :413-427
        take_cand(sx - bk, false);
        ts = F_B_NEXT;
        break;
      case F_B_NEXT:                                         // :431-443
        ++fi;
        if (fi < fwd_n) {
          int max_len = cand_beg + get_end(fi);
          while (max_len < msl && fi < fwd_n) {
            ++fi;
            if (fi < fwd_n) stop = get_end(fi);
            max_len = cand_beg + stop;
          }
        }
        if (fi >= fwd_n && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
        ts = F_B_SELECT;
        break;
      case F_CALL_DONE:
        if (pass == 1) x = ret;
        ts = F_NEXT;
        break;
      case F_FINISH:                                         // passes 1 and 2 done: pass 3, the sort and the final count follow
        a.counts[r] = n_mem;
        if (n_mem > a.cap) atomicExch(a.ctl + 1, 1);
        ts = F_FETCH;
        break;
      }
    }
    if (__ballot_sync(full, ts == F_NONE) == 0) break;
    Intv o = {0, 0, 0}; bool have = false;
    // ---- T: the new pattern is at most K long: one table read
    if (ts == F_NONE && kind == KT) {
      if (COUNT) blocks += cross(ik, phase == P_BE);
      o = tab(phase == P_BE ? pb - 1 : pb, pe - pb + 1);
      have = true;
    }
    // ---- X: bwt_extend on the index, for every lane that wants one once enough of them do
    const unsigned wx = __ballot_sync(full, ts == F_NONE && kind == KX);
    const unsigned wo = __ballot_sync(full, ts == F_NONE && kind != KX);
    if (wx && (__popc(wx) >= a.x_min || !wo)) {
      if (ts == F_NONE && kind == KX) {
        const int c = phase == P_BE ? base(pb - 1) : 3 - base(pe);
        o = extend1(ix, ik, phase == P_BE, c, &blocks);
        have = true;
      }
    }
    if (have) {
      const bool small = o.x2 < (uint64_t)mi, changed = o.x2 != ik.x2;
      if (TEXT && mi == 1 && o.x2 == 1) o = text_intv(ix.sa[o.x0]);       // unique from here on: anchor it in the text
      if (phase == P_F) {                                    // :350-363
        if (changed) push_fwd(ik, pe);
        if (changed && small) ts = F_F_DONE;
        else {
          ik = o; ++pe;
          if (pe < len && !isN(pe)) set_kind(); else { push_fwd(ik, pe); ts = F_F_DONE; }
        }
      } else if (phase == P_BE) {                            // :374-386
        if (small) ts = F_BE_DONE;
        else {
          ik = o; --pb;
          const int j = back_n++;
          if (TEXT && is_text(ik)) { if (bt0 > j) { bt0 = j; bta = ik.x0; } }
          else if (j >= back_s0) back_l[j - back_s0] = pack_entry(ik, 0);
          if (pb > 0 && !isN(pb - 1)) set_kind(); else ts = F_BE_DONE;
        }
      } else {                                               // :407-417
        if (small) ts = F_FE_FAIL;
        else {
          ik = o; ++pe;
          if (pe == stop) ts = F_FE_REACHED; else set_kind();
        }
      }
    }
    // ---- A: anchored interval: compare up to 32 bases of the read with the text
    if (TEXT && ts == F_NONE && kind == KA) {
      if (phase == P_BE) {
        int cap = pb - (anyN ? prev_n(pb) + 1 : 0);          // read bases available on the left (>= 1)
        if ((uint64_t)cap > ik.x0) cap = (int)ik.x0;         // ... and text symbols
        const int want = min(cap, 32);
        int n = 0;
        if (want > 0) {
          const uint64_t diff = qwin_end(pb) ^ twin_end(ik.x0);
          n = min(diff ? (__ffsll((long long)diff) - 1) >> 1 : 32, want);
        }
        pb -= n; ik.x0 -= (uint64_t)n; back_n += n;
        if (!(n == 32 && cap > 32)) ts = F_BE_DONE;
      } else {
        const int lim = phase == P_F ? (anyN ? next_n(pe) : len) : stop;
        const int want = lim - pe;
        const uint64_t tp = ik.x0 + (uint64_t)(pe - pb);
        const int64_t tcap = (int64_t)ix.seq_len - (int64_t)tp;
        const int w = (int)min((int64_t)min(want, 32), max(tcap, (int64_t)0));
        int n = 0;
        if (w > 0) {
          const uint64_t diff = qwin(pe) ^ twin(tp);
          n = min(diff ? __clzll((long long)diff) >> 1 : 32, w);
        }
        pe += n;
        const bool more = n == 32 && want > 32 && tcap > 32;
        if (phase == P_F) { if (!more) { push_fwd(ik, pe); ts = F_F_DONE; } }
        else if (pe == stop) ts = F_FE_REACHED;
        else if (!more) { if (bk >= bt0) bk = bt0; ts = F_FE_FAIL; }   // every anchored trail entry is this occurrence
      }
    }
  }
  if (COUNT && blocks) atomicAdd(a.blocks, blocks);
}

// Pass 3 (src/bwa_wrapper.cpp:228-243): bwt_seed_strategy1 (bwa/bwt.c:358-379) from every position the previous seed ended
// at, on the packed rows; each seed starts from the level-K table (no seed shorter than min_seed_len + 1 > K is ever
// emitted, so the first K-1 extensions only matter through their result).  Appends to the read's slot, sorts it by info
// (:245) and writes the final count.  COUNT walks every extension on the index and counts its block touches.
template <int MAXL, bool COUNT>
__global__ void __launch_bounds__(128) smem_fast_p3_kernel(const SmemArgs a)
{
  using R = Row<MAXL>;
  extern __shared__ uint64_t s_rows[];
  uint64_t *q2 = s_rows + (size_t)threadIdx.x * (R::PACK_BYTES / 8 + ((R::PACK_BYTES / 8) % 2 == 0));
  const uint32_t *nm = (const uint32_t *)(q2 + R::QW);
  const DevIndex &ix = a.ix;
  const unsigned full = 0xffffffffu;
  const int msl = a.min_seed_len;
  const int K = (!COUNT && ix.ktab && ix.ktab_k <= msl) ? ix.ktab_k : 0;
  unsigned long long blocks = 0;
  int64_t r = -1; int len = 0, x = 0, px = 0, cur = 0, n_mem = 0; bool anyN = false; bwaflow_intv_t *mem = nullptr;
  Intv ik = {0, 0, 0};
  bool pending = false, live = true;
  auto base = [&](int i) -> int { return (int)(q2[i >> 5] >> ((~i & 31) << 1)) & 3; };
  auto isN = [&](int i) -> bool { return anyN && ((nm[i >> 5] >> (i & 31)) & 1u); };
  auto qwin = [&](int i) -> uint64_t {
    const int s = (i & 31) << 1; const uint64_t w0 = q2[i >> 5], w1 = q2[(i >> 5) + 1];
    return s ? (w0 << s | w1 >> (64 - s)) : w0;
  };
  for (;;) {
    while (live && !pending) {                               // find the next extension of this lane
      if (r < 0) {
        r = (int64_t)atomicAdd(a.ctl + 3, 1);
        if (r >= a.n_reads) { live = false; break; }
        len = (int)(a.read_off[r + 1] - a.read_off[r]);
        mem = a.slots + r * a.cap; n_mem = a.counts[r]; x = 0;
        if (len >= MAXL || len < msl || a.max_mem_intv <= 0) x = len;   // nothing to add (too-long reads are flagged by pass 1)
        else {
          const uint64_t *src = a.packed + r * (R::PACK_BYTES / 8);
#pragma unroll
          for (int t = 0; t < R::PACK_BYTES / 8; ++t) q2[t] = __ldg(src + t);
          uint32_t any = 0;
#pragma unroll
          for (int t = 0; t < R::W; ++t) any |= nm[t];
          anyN = any != 0;
        }
      }
      while (x < len && isN(x)) ++x;
      if (x >= len) {                                        // read finished: sort by info, final count
        const int n = n_mem < a.cap ? n_mem : a.cap;
        for (int u = 1; u < n; ++u) {
          const bwaflow_intv_t v = mem[u];
          int t = u - 1;
          while (t >= 0 && mem[t].info > v.info) { mem[t + 1] = mem[t]; --t; }
          mem[t + 1] = v;
        }
        a.counts[r] = n_mem;
        if (n_mem > a.cap) atomicExch(a.ctl + 1, 1);
        r = -1;
        continue;
      }
      px = x; ik = set_intv(ix, base(x)); cur = x + 1;       // bwa/bwt.c:364-366
      if (K >= 2 && x + K <= len) {
        bool acgt = true;
        if (anyN) for (int t = 1; t < K; ++t) acgt &= !isN(x + t);
        if (acgt) {
          int d;
          ik = unpack_entry(__ldg(ix.ktab + ktab_base(K) + (qwin(x) >> (64 - 2 * K))), &d);
          cur = x + K;
        }
      }
      if (cur >= len) { x = len; continue; }
      if (isN(cur)) { x = cur + 1; continue; }
      pending = true;
    }
    if (__ballot_sync(full, pending) == 0) break;
    if (pending) {
      Intv o = {0, 0, 0};
      if (COUNT || ik.x2) o = extend1(ix, ik, 0, 3 - base(cur), &blocks);   // bwa/bwt.c:367-376 (an empty interval stays empty; the reference still looks it up)
      if (o.x2 < (uint64_t)a.max_mem_intv && cur - px >= msl) {
        if (o.x2 > 0) {
          if (n_mem < a.cap) {
            bwaflow_intv_t w; w.x[0] = o.x0; w.x[1] = o.x1; w.x[2] = o.x2; w.info = (uint64_t)(uint32_t)px << 32 | (uint32_t)(cur + 1);
            mem[n_mem] = w;
          }
          ++n_mem;
        }
        x = cur + 1; pending = false;
      } else {
        ik = o; ++cur;
        if (cur >= len) { x = len; pending = false; }
        else if (isN(cur)) { x = cur + 1; pending = false; }
        else pending = true;
      }
    }
  }
  if (COUNT && blocks) atomicAdd(a.blocks, blocks);
}

template <int MAXL>
static int fast_launch_t(bwaflow_smem_batch *b, const SmemArgs &a, int grid, int grid3, bool text, bool count)
{
  using R = Row<MAXL>;
  cudaStream_t st = b->stream;
  const int64_t nr = b->n_reads;
  pack_reads_kernel<MAXL><<<(unsigned)((nr * 32 + 255) / 256), 256, 0, st>>>(a.bases, a.read_off, nr, (uint64_t *)a.packed);
  const size_t sh = 128 * (size_t)R::ROW_BYTES;
  const size_t sh3 = 128 * (size_t)(R::PACK_BYTES / 8 + ((R::PACK_BYTES / 8) % 2 == 0)) * 8;
  if (count) {                                  // the reference's route step by step, counting its occ-block touches
    smem_fast_kernel<MAXL, false, true><<<grid, 128, sh, st>>>(a);
    smem_fast_p3_kernel<MAXL, true><<<grid3, 128, sh3, st>>>(a);
  } else {
    int r = smem_sweep_launch(b, a, text);
    if (r) return r;
    smem_fast_p3_kernel<MAXL, false><<<grid3, 128, sh3, st>>>(a);
  }
  BF_CUDA(cudaGetLastError());
  return 0;
}

// Launches pack + passes 1-2 + pass 3 for a batch whose reads are all shorter than 256 bases.
int smem_fast_launch(bwaflow_smem_batch *b, const SmemArgs &a, bool text, bool count)
{
  const int64_t want = (b->n_reads + 127) / 128;
  const int n_sm = b->ctx->n_sm;
  const int grid = (int)(want < (int64_t)n_sm * b->fast_blocks_per_sm ? want : (int64_t)n_sm * b->fast_blocks_per_sm);
  const int grid3 = (int)(want < (int64_t)n_sm * 10 ? want : (int64_t)n_sm * 10);
  if (b->max_len < 160) return fast_launch_t<160>(b, a, grid, grid3, text, count);
  return fast_launch_t<256>(b, a, grid, grid3, text, count);
}

int smem_fast_setup(bwaflow_smem_batch *b)
{
  int per_sm = 0;
  BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smem_fast_kernel<160, false, true>, 128, 128 * Row<160>::ROW_BYTES));
  if (per_sm < 1) per_sm = 1;
  if (const char *s = getenv("BWAFLOW_SMEM_BLOCKS_PER_SM")) per_sm = atoi(s) > 0 ? atoi(s) : per_sm;
  b->fast_blocks_per_sm = per_sm;
  b->x_min = 16;                                 // measured: 18.9 ms at 1, 17.3 ms at 16 and 24 (3.1 Gbp, 533k pairs)
  if (const char *s = getenv("BWAFLOW_SMEM_XMIN")) b->x_min = atoi(s);
  BF_CUDA(cudaMalloc(&b->d_packed, (size_t)b->max_reads * Row<256>::PACK_BYTES + 64));
  return smem_sweep_setup(b);
}

} // namespace bwaflow
