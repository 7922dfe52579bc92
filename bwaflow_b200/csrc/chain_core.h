// Chaining core shared by the device kernel (chain_dev.cu) and the host entry point (chain_host.cpp):
// seeds of one read (intervals + their SA positions) -> filtered chains, exactly as the reference does between
// its two accelerated stages:
//   mem_chain_postprocess   src/bwa_wrapper.cpp:248-323  (frac_rep, B-tree chaining with test_and_merge, in-order dump)
//   mem_chain_flt           bwa/bwamem.c:336-394         (weights, klib introsort by weight, overlap filter)
// klib behaviour that shows in the output is reproduced exactly: the B-tree of bwa/kbtree.h (t = 5: 40-byte keys in
// 512-byte nodes; equal chain positions make kb_intervalp's answer and the traversal order layout-dependent) and the
// introsort of bwa/ksort.h:176-232 (order of weight ties).  Everything works on caller-provided scratch arrays sized
// by the read's seed count P, so one thread (host or device) handles one read with no allocation.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/bwaflow_cuda.h"
#include "klib_sort.h"

#ifdef __CUDACC__
#define BF_HD __host__ __device__ __forceinline__
#define BF_UNROLL _Pragma("unroll")
#else
#define BF_HD inline
#define BF_UNROLL
#endif

namespace bwaflow {
namespace chain {

struct SeedNode { bwaflow_seed_t s; int next; int _pad; };      // 32 B
struct Chain {
  int64_t pos;
  int n, rid, head, tail;                                       // seeds: singly linked list in the read's pool
  int w, kept, first, is_alt;
  int beg, end;                                                 // query span, filled by chain_flt (chn_beg / chn_end, bwa/bwamem.c:40-41)
};                                                              // 48 B
constexpr int BT_T = 5, BT_MAXK = 2 * BT_T - 1;                 // kb_init(chn, 512): t = ((512-4-8)/(8+40)+1)>>1
struct Node { int internal, n; int key[BT_MAXK]; int ptr[BT_MAXK + 1]; int64_t kpos[BT_MAXK]; };   // kpos[i] = chains[key[i]].pos, kept beside the key: one load per probe instead of two dependent ones

struct Scratch {
  SeedNode *seeds; Chain *chains; Node *nodes; int *order; Chain *sorted; int *kept;
  int n_seeds, n_chains, n_nodes, root;
};
// scratch entries needed for a read with P seed positions
BF_HD int64_t node_cap(int64_t P) { return P / 4 + 4; }

struct Ctg { int n; const int64_t *off; const int32_t *len; const int32_t *alt; int64_t l_pac; };

// bns_pos2rid, bwa/bntseq.c:349-363
BF_HD int pos2rid(const Ctg &c, int64_t pos_f)
{
  if (pos_f >= c.l_pac) return -1;
  int left = 0, mid = 0, right = c.n;
  while (left < right) {
    mid = (left + right) >> 1;
    if (pos_f >= c.off[mid]) {
      if (mid == c.n - 1) break;
      if (pos_f < c.off[mid + 1]) break;
      left = mid + 1;
    } else right = mid;
  }
  return mid;
}
BF_HD int64_t depos(const Ctg &c, int64_t pos) { return pos >= c.l_pac ? (c.l_pac << 1) - 1 - pos : pos; }
// bns_intv2rid, bwa/bntseq.c:365-373
BF_HD int intv2rid(const Ctg &c, int64_t rb, int64_t re)
{
  if (rb < c.l_pac && re > c.l_pac) return -2;
  const int rid_b = pos2rid(c, depos(c, rb));
  const int rid_e = rb < re ? pos2rid(c, depos(c, re - 1)) : rid_b;
  return rid_b == rid_e ? rid_b : -1;
}

// ---- bwa/kbtree.h over chain indices -----------------------------------------------------------------------
// __kb_getp_aux: index of the last key <= k in the node (first of a run of equal keys); *r = sign of the comparison.
// klib's binary search is a lower bound over the node's sorted keys, i.e. the number of keys below k: all (at most 9) key
// positions are loaded at once and counted, one memory latency instead of a chain of three or four dependent probes -- the
// chaining kernel is bound by exactly that latency, and a read inside a repeat walks trees several levels deep per seed.
BF_HD int bt_find(const Scratch &S, const Node &x, int64_t k, int *r)
{
  const int n = x.n;
  int64_t kp[BT_MAXK];
BF_UNROLL
  for (int m = 0; m < BT_MAXK; ++m) kp[m] = x.kpos[m];          // entries at and beyond n are never looked at below
  if (n == 0) return -1;
  int lt = 0, le = 0;
BF_UNROLL
  for (int m = 0; m < BT_MAXK; ++m) { lt += (m < n) & (kp[m] < k); le += (m < n) & (kp[m] <= k); }
  if (lt == n) { *r = 1; return n - 1; }
  if (le > lt) { *r = 0; return lt; }                            // kpos[lt] == k
  *r = -1;
  return lt - 1;
}
// kb_intervalp (only `lower` is used by the caller)
BF_HD int bt_lower(const Scratch &S, int64_t k)
{
  int lower = -1, x = S.root, r = 0;
  while (x >= 0) {
    const Node &nd = S.nodes[x];
    const int i = bt_find(S, nd, k, &r);
    if (i >= 0 && r == 0) return nd.key[i];
    if (i >= 0) lower = nd.key[i];
    if (!nd.internal) return lower;
    x = nd.ptr[i + 1];
  }
  return lower;
}
BF_HD int bt_new_node(Scratch &S, int internal)
{
  Node &nd = S.nodes[S.n_nodes];
  nd.internal = internal; nd.n = 0;
  for (int i = 0; i < BT_MAXK; ++i) nd.key[i] = 0;
  for (int i = 0; i <= BT_MAXK; ++i) nd.ptr[i] = -1;
  return S.n_nodes++;
}
// __kb_split: child y = x.ptr[i] is full; its upper half moves to a new right sibling, the median moves up into x
BF_HD void bt_split(Scratch &S, int xi, int i, int yi)
{
  const int zi = bt_new_node(S, S.nodes[yi].internal);
  Node &x = S.nodes[xi], &y = S.nodes[yi], &z = S.nodes[zi];
  z.n = BT_T - 1;
  for (int j = 0; j < BT_T - 1; ++j) { z.key[j] = y.key[BT_T + j]; z.kpos[j] = y.kpos[BT_T + j]; }
  if (y.internal) for (int j = 0; j < BT_T; ++j) z.ptr[j] = y.ptr[BT_T + j];
  y.n = BT_T - 1;
  for (int j = x.n; j > i; --j) x.ptr[j + 1] = x.ptr[j];
  x.ptr[i + 1] = zi;
  for (int j = x.n - 1; j >= i; --j) { x.key[j + 1] = x.key[j]; x.kpos[j + 1] = x.kpos[j]; }
  x.key[i] = y.key[BT_T - 1]; x.kpos[i] = y.kpos[BT_T - 1];
  ++x.n;
}
// kb_putp
BF_HD void bt_put(Scratch &S, int c)
{
  int xi = S.root;
  if (S.nodes[xi].n == BT_MAXK) {
    const int s = bt_new_node(S, 1);
    S.nodes[s].ptr[0] = xi;
    S.root = s;
    bt_split(S, s, 0, xi);
    xi = s;
  }
  const int64_t k = S.chains[c].pos;
  int r;
  for (;;) {
    Node &x = S.nodes[xi];
    if (!x.internal) {
      const int i = bt_find(S, x, k, &r);
      for (int j = x.n - 1; j > i; --j) { x.key[j + 1] = x.key[j]; x.kpos[j + 1] = x.kpos[j]; }
      x.key[i + 1] = c; x.kpos[i + 1] = k;
      ++x.n;
      return;
    }
    int i = bt_find(S, x, k, &r) + 1;
    if (S.nodes[x.ptr[i]].n == BT_MAXK) {
      bt_split(S, xi, i, x.ptr[i]);
      if (k > S.nodes[xi].kpos[i]) ++i;
    }
    xi = S.nodes[xi].ptr[i];
  }
}
// __kb_traverse: in-order into S.order; returns the count
BF_HD int bt_inorder(const Scratch &S)
{
  int sx[24], si[24], sp = 0, n = 0;
  sx[0] = S.root; si[0] = 0;
  for (;;) {
    while (sx[sp] >= 0 && si[sp] <= S.nodes[sx[sp]].n) {
      const Node &nd = S.nodes[sx[sp]];
      sx[sp + 1] = nd.internal ? nd.ptr[si[sp]] : -1;
      si[sp + 1] = 0;
      ++sp;
    }
    --sp;
    if (sp < 0) break;
    if (sx[sp] >= 0 && si[sp] < S.nodes[sx[sp]].n) S.order[n++] = S.nodes[sx[sp]].key[si[sp]];
    ++si[sp];
  }
  return n;
}

// test_and_merge, bwa/bwamem.c:199-220
BF_HD int test_and_merge(Scratch &S, const bwaflow_chain_opt_t &o, int64_t l_pac, int ci, const bwaflow_seed_t &p, int seed_rid)
{
  Chain &c = S.chains[ci];
  const bwaflow_seed_t first = S.seeds[c.head].s, last = S.seeds[c.tail].s;
  const int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
  if (seed_rid != c.rid) return 0;
  if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return 1;
  if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return 0;
  const int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
  if (y >= 0 && x - y <= o.w && y - x <= o.w && x - last.len < o.max_chain_gap && y - last.len < o.max_chain_gap) {
    SeedNode &sn = S.seeds[S.n_seeds];
    sn.s = p; sn.next = -1;
    S.seeds[c.tail].next = S.n_seeds;
    c.tail = S.n_seeds++;
    ++c.n;
    return 1;
  }
  return 0;
}

// mem_chain_weight, bwa/bwamem.c:222-243
BF_HD int chain_weight(const Scratch &S, const Chain &c)
{
  int64_t end = 0; int w = 0;
  for (int j = c.head; j >= 0; j = S.seeds[j].next) {
    const bwaflow_seed_t &s = S.seeds[j].s;
    if (s.qbeg >= end) w += s.len; else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
    end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
  }
  const int tmp = w; w = 0; end = 0;
  for (int j = c.head; j >= 0; j = S.seeds[j].next) {
    const bwaflow_seed_t &s = S.seeds[j].s;
    if (s.rbeg >= end) w += s.len; else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
    end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
  }
  w = w < tmp ? w : tmp;
  return w < 1 << 30 ? w : (1 << 30) - 1;
}

// ---- bwa/ksort.h:140-232 for KSORT_INIT(mem_flt, mem_chain_t, flt_lt), flt_lt(a,b) = a.w > b.w ---------------------
struct FltLt { BF_HD bool operator()(const Chain &a, const Chain &b) const { return a.w > b.w; } };   // flt_lt, bwa/bwamem.c:334

BF_HD int chn_beg(const Scratch &S, const Chain &c) { return S.seeds[c.head].s.qbeg; }
BF_HD int chn_end(const Scratch &S, const Chain &c) { return S.seeds[c.tail].s.qbeg + S.seeds[c.tail].s.len; }

// mem_chain_flt, bwa/bwamem.c:336-394 on S.sorted[0..n_chn), in three steps so that the device can run the middle one -- the overlap
// filter, quadratic in the number of chains for a read inside a repeat -- with a whole warp (hotpath.cu).
struct KeptEnt { int beg, end, w, alt; };                      // what the filter reads of a kept chain
BF_HD KeptEnt *flt_kept_ents(const Scratch &S) { return (KeptEnt *)S.chains; }            // S.chains is free once the chains are in S.sorted
BF_HD int *flt_kept_first(const Scratch &S, int n_chn) { return (int *)(flt_kept_ents(S) + n_chn); }   // 16 + 4 bytes per entry <= sizeof(Chain)

// (1) weights, query spans, minimum weight, klib introsort by weight (:341-354); returns the number of chains left
BF_HD int flt_prepare(Scratch &S, const bwaflow_chain_opt_t &o, int n_chn)
{
  Chain *a = S.sorted;
  int k = 0;
  for (int i = 0; i < n_chn; ++i) {
    Chain &c = a[i];
    c.first = -1; c.kept = 0;
    c.w = chain_weight(S, c);
    c.beg = chn_beg(S, c); c.end = chn_end(S, c);
    if (c.w >= o.min_chain_weight) a[k++] = c;
  }
  if (k < 32) { if (k) klib_introsort(k, a, FltLt()); return k; }
  // Many chains (a read inside a repeat): klib's introsort runs on (weight, index) pairs instead of the 48-byte chains -- it only
  // ever looks at weights, so it makes the same comparisons and the same swaps and leaves the same order among equal weights --
  // and the chains are then moved once, in place, along the cycles of that permutation.
  struct WK { int w, idx; };
  struct WKLt { BF_HD bool operator()(const WK &x, const WK &y) const { return x.w > y.w; } };
  WK *wk = (WK *)S.chains;                                       // free once the chains are in S.sorted
  for (int i = 0; i < k; ++i) { wk[i].w = a[i].w; wk[i].idx = i; }
  klib_introsort(k, wk, WKLt());
  for (int i = 0; i < k; ++i) {
    if (wk[i].idx == i) continue;
    const Chain tmp = a[i];
    int j = i;
    for (;;) {
      const int src = wk[j].idx;
      wk[j].idx = j;
      if (src == i) { a[j] = tmp; break; }
      a[j] = a[src];
      j = src;
    }
  }
  return k;
}

// one kept chain against chain i (:359-371): bit 0 = significant overlap, bit 1 = chain i is dropped here
BF_HD int flt_test(const bwaflow_chain_opt_t &o, const KeptEnt &e, int bi, int ei, int wi, int alt_i)
{
  const int b_max = e.beg > bi ? e.beg : bi, e_min = e.end < ei ? e.end : ei;
  if (e_min > b_max && (!e.alt || alt_i)) {
    const int li = ei - bi, lj = e.end - e.beg, min_l = li < lj ? li : lj;
    if (e_min - b_max >= min_l * o.mask_level && min_l < o.max_chain_gap)
      return 1 | ((wi < e.w * o.drop_ratio && e.w - wi >= o.min_seed_len << 1) ? 2 : 0);
  }
  return 0;
}

// (2) the overlap filter (:355-380), one thread: kept chains in a compact array in kept order, four entries loaded ahead of the
// tests instead of one dependent chain of loads (kept index -> chain -> fields) per entry.  Returns the number of kept chains.
BF_HD int flt_filter(Scratch &S, const bwaflow_chain_opt_t &o, int n_chn)
{
  Chain *a = S.sorted;
  KeptEnt *K = flt_kept_ents(S);
  int *F = flt_kept_first(S, n_chn);
  int n_kept = 0;
  a[0].kept = 3; S.kept[n_kept] = 0; K[n_kept] = KeptEnt{a[0].beg, a[0].end, a[0].w, a[0].is_alt}; F[n_kept] = -1; ++n_kept;
  for (int i = 1; i < n_chn; ++i) {
    int large_ovlp = 0, kk = 0;
    const int bi = a[i].beg, ei = a[i].end, wi = a[i].w, alt_i = a[i].is_alt;
    bool dropped = false;
    while (kk < n_kept && !dropped) {
      KeptEnt e[4];
      const int nb = n_kept - kk < 4 ? n_kept - kk : 4;
      BF_UNROLL
      for (int u = 0; u < 4; ++u) e[u] = K[kk + (u < nb ? u : 0)];
      BF_UNROLL
      for (int u = 0; u < 4; ++u) {
        if (u >= nb || dropped) continue;
        const int t = flt_test(o, e[u], bi, ei, wi, alt_i);
        if (t & 1) { large_ovlp = 1; if (F[kk + u] < 0) F[kk + u] = i; }
        if (t & 2) dropped = true;
      }
      kk += nb;
    }
    if (!dropped) {
      S.kept[n_kept] = i; K[n_kept] = KeptEnt{bi, ei, wi, alt_i}; F[n_kept] = -1; ++n_kept;
      a[i].kept = large_ovlp ? 2 : 3;
    }
  }
  return n_kept;
}

// (3) shadowed hits, max_chain_extend, compaction (:381-393); returns the number of surviving chains
BF_HD int flt_finish(Scratch &S, const bwaflow_chain_opt_t &o, int n_chn, int n_kept)
{
  Chain *a = S.sorted;
  const int *F = flt_kept_first(S, n_chn);
  int i, k;
  for (i = 0; i < n_kept; ++i) if (F[i] >= 0) a[F[i]].kept = 1;
  for (i = k = 0; i < n_chn; ++i) {
    if (a[i].kept == 0 || a[i].kept == 3) continue;
    if (++k >= o.max_chain_extend) break;
  }
  for (; i < n_chn; ++i) if (a[i].kept < 3) a[i].kept = 0;
  for (i = k = 0; i < n_chn; ++i) if (a[i].kept != 0) a[k++] = a[i];
  return k;
}

// Would mem_flt_chained_seeds (bwa/bwamem.c:607-624) rescore the seeds of a read this long?  (Then we refuse it.)
BF_HD bool needs_seed_rescoring(const bwaflow_chain_opt_t &o, int len)
{
  const double min_l = o.min_chain_weight ? 1.1f * o.min_chain_weight : 5.5f * log((double)len);
  return !(min_l > 0.05f * len);
}

// One read.  intv[0..n_intv) sorted by info, positions of interval i at pos[pos_off[i]-pos_base .. ).  Chains land in
// S.sorted[0..ret) (traversal order filtered by mem_chain_flt), their seeds are the linked lists in S.seeds.
// *frac_rep receives (float)l_rep/len (src/bwa_wrapper.cpp:273-280,318).
// chain_seeds = mem_chain_postprocess and the first step of mem_chain_flt (weights, sort): S.sorted[0..ret) by weight.
BF_HD int chain_seeds(Scratch &S, const bwaflow_chain_opt_t &o, const Ctg &ctg, int len, const bwaflow_intv_t *intv, int64_t n_intv,
                     const int64_t *pos, const int64_t *pos_off, float *frac_rep)
{
  S.n_seeds = S.n_chains = S.n_nodes = 0;
  *frac_rep = 0.f;
  if (len < o.min_seed_len) return 0;                                   // src/bwa_wrapper.cpp:270
  S.root = bt_new_node(S, 0);
  int b = 0, e = 0, l_rep = 0;
  for (int64_t i = 0; i < n_intv; ++i) {
    const int sb = (int)(intv[i].info >> 32), se = (int)(uint32_t)intv[i].info;
    if (intv[i].x[2] <= (uint64_t)o.max_occ) continue;
    if (sb > e) { l_rep += e - b; b = sb; e = se; } else e = e > se ? e : se;
  }
  l_rep += e - b;
  for (int64_t i = 0; i < n_intv; ++i) {                                // :281-309
    const uint64_t info = intv[i].info;
    const int slen = (int)((uint32_t)info - (uint32_t)(info >> 32));
    for (int64_t t = pos_off[i]; t < pos_off[i + 1]; ++t) {
      bwaflow_seed_t s; s.rbeg = pos[t]; s.qbeg = (int)(info >> 32); s.len = s.score = slen; s._pad = 0;
      const int rid = intv2rid(ctg, s.rbeg, s.rbeg + s.len);
      if (rid < 0) continue;
      bool to_add = true;
      if (S.n_chains) {
        const int lower = bt_lower(S, s.rbeg);
        if (lower >= 0 && test_and_merge(S, o, ctg.l_pac, lower, s, rid)) to_add = false;
      }
      if (to_add) {
        SeedNode &sn = S.seeds[S.n_seeds];
        sn.s = s; sn.next = -1;
        Chain &c = S.chains[S.n_chains];
        c.pos = s.rbeg; c.n = 1; c.rid = rid; c.head = c.tail = S.n_seeds++;
        c.w = 0; c.kept = 0; c.first = -1; c.is_alt = ctg.alt ? !!ctg.alt[rid] : 0;
        bt_put(S, S.n_chains++);
      }
    }
  }
  const int n = bt_inorder(S);
  for (int i = 0; i < n; ++i) S.sorted[i] = S.chains[S.order[i]];
  *frac_rep = (float)l_rep / len;
  return flt_prepare(S, o, n);
}

// mem_chain_flt's remaining steps on what chain_seeds left in S.sorted[0..n)
BF_HD int chain_filter(Scratch &S, const bwaflow_chain_opt_t &o, int n)
{
  if (n == 0) return 0;
  return flt_finish(S, o, n, flt_filter(S, o, n));
}

// One read, everything on the calling thread
BF_HD int chain_read(Scratch &S, const bwaflow_chain_opt_t &o, const Ctg &ctg, int len, const bwaflow_intv_t *intv, int64_t n_intv,
                     const int64_t *pos, const int64_t *pos_off, float *frac_rep)
{
  return chain_filter(S, o, chain_seeds(S, o, ctg, len, intv, n_intv, pos, pos_off, frac_rep));
}

} // namespace chain
} // namespace bwaflow
