// Host side of the accelerator return path: intervals (+ SA positions computed on the device) -> chains.
// This is what ChainsPipe::compute (src/Pipeline.cpp:408-438) does between the two device stages:
//   mem_chain_postprocess   src/bwa_wrapper.cpp:248-323  (frac_rep, seeds, B-tree chaining, in-order traversal)
//   mem_chain_flt           bwa/bwamem.c:336-394         (weights, unstable introsort by weight, overlap filter)
//   mem_flt_chained_seeds   bwa/bwamem.c:607-624         (a no-op for reads this short; longer reads are refused)
// restated so that a backend can run reads -> regions without bouncing through the reference's own host code.
// Two library behaviours are observable in the output and are therefore reproduced exactly, not approximated:
//   * klib's B-tree (bwa/kbtree.h, t = 5 for 40-byte keys in 512-byte nodes): with equal chain positions the
//     chain found by kb_intervalp and the traversal order depend on the node layout;
//   * klib's introsort (bwa/ksort.h:176-232): ties in chain weight come out in an order fixed by its
//     median-of-3 partition + final insertion sort.
// Multi-threaded over reads (std::thread); no GPU work here.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "../../include/bwaflow_cuda.h"

namespace bwaflow {
void set_error(const char *fmt, ...);
}

namespace {

struct SeedNode { bwaflow_seed_t s; int next; };

struct Chain {
  int64_t pos;
  int n, rid, head, tail;       // seeds as a singly linked list in the per-read pool
  int w, kept, first, is_alt;
};

constexpr int BT_T = 5, BT_MAXK = 2 * BT_T - 1;   // kb_init(chn, 512): t = ((512-4-8)/(8+40)+1)>>1

struct Node { int internal, n; int key[BT_MAXK]; int ptr[BT_MAXK + 1]; };

struct Scratch {
  std::vector<SeedNode> seeds;
  std::vector<Chain> chains;
  std::vector<Node> nodes;
  std::vector<int> order, kept_idx, stack_n, stack_i;
  std::vector<Chain> sorted;
  int root;
};

struct Ctg { int n; const int64_t *off; const int32_t *len; const int32_t *alt; int64_t l_pac; };

// bns_pos2rid, bwa/bntseq.c:349-363
static int pos2rid(const Ctg &c, int64_t pos_f)
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
static inline int64_t depos(const Ctg &c, int64_t pos) { return pos >= c.l_pac ? (c.l_pac << 1) - 1 - pos : pos; }
// bns_intv2rid, bwa/bntseq.c:365-373
static int intv2rid(const Ctg &c, int64_t rb, int64_t re)
{
  if (rb < c.l_pac && re > c.l_pac) return -2;
  const int rid_b = pos2rid(c, depos(c, rb));
  const int rid_e = rb < re ? pos2rid(c, depos(c, re - 1)) : rid_b;
  return rid_b == rid_e ? rid_b : -1;
}

// ---- kbtree.h emulation over chain indices ------------------------------------------------------
// __kb_getp_aux: index of the last key <= k in the node (first of a run of equal keys), r = sign of the comparison
static inline int bt_find(const Scratch &S, const Node &x, int64_t k, int *r)
{
  if (x.n == 0) return -1;
  int begin = 0, end = x.n;
  while (begin < end) {
    const int mid = (begin + end) >> 1;
    if (S.chains[x.key[mid]].pos < k) begin = mid + 1; else end = mid;
  }
  if (begin == x.n) { *r = 1; return x.n - 1; }
  const int64_t kp = S.chains[x.key[begin]].pos;
  *r = (kp < k) - (k < kp);
  if (*r < 0) --begin;
  return begin;
}
// kb_intervalp: only `lower` is used by the caller
static int bt_lower(const Scratch &S, int64_t k)
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
static int bt_new_node(Scratch &S, int internal)
{
  Node nd; memset(&nd, 0, sizeof(nd)); nd.internal = internal;
  for (int i = 0; i <= BT_MAXK; ++i) nd.ptr[i] = -1;
  S.nodes.push_back(nd);
  return (int)S.nodes.size() - 1;
}
// __kb_split: child y = x.ptr[i] is full; its upper half moves to a new right sibling, the median moves up into x
static void bt_split(Scratch &S, int xi, int i, int yi)
{
  const int zi = bt_new_node(S, S.nodes[yi].internal);
  Node &x = S.nodes[xi], &y = S.nodes[yi], &z = S.nodes[zi];
  z.n = BT_T - 1;
  memcpy(z.key, y.key + BT_T, sizeof(int) * (BT_T - 1));
  if (y.internal) memcpy(z.ptr, y.ptr + BT_T, sizeof(int) * BT_T);
  y.n = BT_T - 1;
  memmove(x.ptr + i + 2, x.ptr + i + 1, sizeof(int) * (x.n - i));
  x.ptr[i + 1] = zi;
  memmove(x.key + i + 1, x.key + i, sizeof(int) * (x.n - i));
  x.key[i] = y.key[BT_T - 1];
  ++x.n;
}
static void bt_put_aux(Scratch &S, int xi, int c)
{
  const int64_t k = S.chains[c].pos;
  int r;
  for (;;) {
    if (!S.nodes[xi].internal) {
      Node &x = S.nodes[xi];
      const int i = bt_find(S, x, k, &r);
      if (i != x.n - 1) memmove(x.key + i + 2, x.key + i + 1, sizeof(int) * (x.n - i - 1));
      x.key[i + 1] = c;
      ++x.n;
      return;
    }
    int i = bt_find(S, S.nodes[xi], k, &r) + 1;
    if (S.nodes[S.nodes[xi].ptr[i]].n == BT_MAXK) {
      bt_split(S, xi, i, S.nodes[xi].ptr[i]);
      const int64_t kp = S.chains[S.nodes[xi].key[i]].pos;
      if (k > kp) ++i;
    }
    xi = S.nodes[xi].ptr[i];
  }
}
static void bt_put(Scratch &S, int c)
{
  int r = S.root;
  if (S.nodes[r].n == BT_MAXK) {
    const int s = bt_new_node(S, 1);
    S.root = s; S.nodes[s].n = 0; S.nodes[s].ptr[0] = r;
    bt_split(S, s, 0, r);
    r = s;
  }
  bt_put_aux(S, r, c);
}
// __kb_traverse: in-order
static void bt_inorder(const Scratch &S, int x, std::vector<int> &out)
{
  if (x < 0) return;
  const Node &nd = S.nodes[x];
  for (int i = 0; i <= nd.n; ++i) {
    if (nd.internal) bt_inorder(S, nd.ptr[i], out);
    if (i < nd.n) out.push_back(nd.key[i]);
  }
}

// test_and_merge, bwa/bwamem.c:199-220
static int test_and_merge(Scratch &S, const bwaflow_chain_opt_t &o, int64_t l_pac, int ci, const bwaflow_seed_t &p, int seed_rid)
{
  Chain &c = S.chains[ci];
  const bwaflow_seed_t &first = S.seeds[c.head].s, &last = S.seeds[c.tail].s;
  const int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
  if (seed_rid != c.rid) return 0;
  if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return 1;
  if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return 0;
  const int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
  if (y >= 0 && x - y <= o.w && y - x <= o.w && x - last.len < o.max_chain_gap && y - last.len < o.max_chain_gap) {
    SeedNode sn; sn.s = p; sn.next = -1;
    S.seeds.push_back(sn);
    S.seeds[c.tail].next = (int)S.seeds.size() - 1;
    c.tail = (int)S.seeds.size() - 1;
    ++c.n;
    return 1;
  }
  return 0;
}

// mem_chain_weight, bwa/bwamem.c:222-243
static int chain_weight(const Scratch &S, const Chain &c)
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

// ---- ksort.h:176-232 ks_introsort(mem_flt) with flt_lt(a,b) = a.w > b.w, behaviour-exact -----------------
#define FLT_LT(a, b) ((a).w > (b).w)
static void flt_insertsort(Chain *s, Chain *t)
{
  for (Chain *i = s + 1; i < t; ++i)
    for (Chain *j = i; j > s && FLT_LT(*j, *(j - 1)); --j) { Chain tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}
static void flt_combsort(size_t n, Chain *a)
{
  const double shrink = 1.2473309501039786540366528676643;
  int do_swap; size_t gap = n;
  do {
    if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
    do_swap = 0;
    for (Chain *i = a; i < a + n - gap; ++i) {
      Chain *j = i + gap;
      if (FLT_LT(*j, *i)) { Chain tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
    }
  } while (do_swap || gap > 2);
  if (gap != 1) flt_insertsort(a, a + n);
}
static void flt_introsort(size_t n, Chain *a)
{
  struct Frame { Chain *left, *right; int depth; };
  if (n < 1) return;
  if (n == 2) { if (FLT_LT(a[1], a[0])) { Chain t = a[0]; a[0] = a[1]; a[1] = t; } return; }
  int d;
  for (d = 2; 1ul << d < n; ++d) {}
  std::vector<Frame> stack; stack.reserve(sizeof(size_t) * d + 2);
  Chain *s = a, *t = a + (n - 1), *i, *j, *k, rp, tmp;
  d <<= 1;
  for (;;) {
    if (s < t) {
      if (--d == 0) { flt_combsort(t - s + 1, s); t = s; continue; }
      i = s; j = t; k = i + ((j - i) >> 1) + 1;
      if (FLT_LT(*k, *i)) { if (FLT_LT(*k, *j)) k = j; }
      else k = FLT_LT(*j, *i) ? i : j;
      rp = *k;
      if (k != t) { tmp = *k; *k = *t; *t = tmp; }
      for (;;) {
        do ++i; while (FLT_LT(*i, rp));
        do --j; while (i <= j && FLT_LT(rp, *j));
        if (j <= i) break;
        tmp = *i; *i = *j; *j = tmp;
      }
      tmp = *i; *i = *t; *t = tmp;
      if (i - s > t - i) {
        if (i - s > 16) stack.push_back(Frame{s, i - 1, d});
        s = t - i > 16 ? i + 1 : t;
      } else {
        if (t - i > 16) stack.push_back(Frame{i + 1, t, d});
        t = i - s > 16 ? i - 1 : s;
      }
    } else {
      if (stack.empty()) { flt_insertsort(a, a + n); return; }
      s = stack.back().left; t = stack.back().right; d = stack.back().depth; stack.pop_back();
    }
  }
}

static inline int chn_beg(const Scratch &S, const Chain &c) { return S.seeds[c.head].s.qbeg; }
static inline int chn_end(const Scratch &S, const Chain &c) { return S.seeds[c.tail].s.qbeg + S.seeds[c.tail].s.len; }

// mem_chain_flt, bwa/bwamem.c:336-394.  `a` holds the chains in traversal order; returns the survivors in order.
static void chain_flt(Scratch &S, const bwaflow_chain_opt_t &o, std::vector<Chain> &a)
{
  int n_chn = (int)a.size();
  if (n_chn == 0) return;
  int k = 0;
  for (int i = 0; i < n_chn; ++i) {
    Chain &c = a[i];
    c.first = -1; c.kept = 0;
    c.w = chain_weight(S, c);
    if (c.w >= o.min_chain_weight) a[k++] = c;
  }
  n_chn = k; a.resize(n_chn);
  if (n_chn == 0) return;   // (the reference would touch a[0] here; unreachable with min_chain_weight <= smallest weight)
  flt_introsort(n_chn, a.data());
  std::vector<int> &kept = S.kept_idx; kept.clear();
  a[0].kept = 3; kept.push_back(0);
  for (int i = 1; i < n_chn; ++i) {
    int large_ovlp = 0; size_t kk;
    for (kk = 0; kk < kept.size(); ++kk) {
      const int j = kept[kk];
      const int bi = chn_beg(S, a[i]), bj = chn_beg(S, a[j]), ei = chn_end(S, a[i]), ej = chn_end(S, a[j]);
      const int b_max = bj > bi ? bj : bi, e_min = ej < ei ? ej : ei;
      if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
        const int li = ei - bi, lj = ej - bj, min_l = li < lj ? li : lj;
        if (e_min - b_max >= min_l * o.mask_level && min_l < o.max_chain_gap) {
          large_ovlp = 1;
          if (a[j].first < 0) a[j].first = i;
          if (a[i].w < a[j].w * o.drop_ratio && a[j].w - a[i].w >= o.min_seed_len << 1) break;
        }
      }
    }
    if (kk == kept.size()) { kept.push_back(i); a[i].kept = large_ovlp ? 2 : 3; }
  }
  for (size_t i = 0; i < kept.size(); ++i) { const Chain &c = a[kept[i]]; if (c.first >= 0) a[c.first].kept = 1; }
  int i;
  for (i = k = 0; i < n_chn; ++i) {
    if (a[i].kept == 0 || a[i].kept == 3) continue;
    if (++k >= o.max_chain_extend) break;
  }
  for (; i < n_chn; ++i) if (a[i].kept < 3) a[i].kept = 0;
  for (i = k = 0; i < n_chn; ++i) if (a[i].kept != 0) a[k++] = a[i];
  a.resize(k);
}

struct Job {
  const bwaflow_chain_opt_t *opt; Ctg ctg;
  int64_t n_reads; const int64_t *read_off;
  const bwaflow_intv_t *intv; const int64_t *intv_off; const int64_t *pos; const int64_t *pos_off;
  std::vector<std::vector<bwaflow_chain_t>> out_chains;   // per block of reads
  std::vector<std::vector<bwaflow_seed_t>> out_seeds;
  std::vector<int64_t> n_chains_of_read;
  std::atomic<int64_t> next{0};
  std::atomic<int> err{0};
};

constexpr int64_t BLOCK = 512;

static void chain_read(Scratch &S, const Job &J, int64_t r, std::vector<bwaflow_chain_t> &oc, std::vector<bwaflow_seed_t> &os, int *err)
{
  const bwaflow_chain_opt_t &o = *J.opt;
  const int len = (int)(J.read_off[r + 1] - J.read_off[r]);
  const int64_t i0 = J.intv_off[r], i1 = J.intv_off[r + 1];
  if (len < o.min_seed_len) return;                                   // src/bwa_wrapper.cpp:270
  {
    const double min_l = o.min_chain_weight ? 1.1f * o.min_chain_weight : 5.5f * log((double)len);
    if (!(min_l > 0.05f * len)) { *err = 1; return; }                // mem_flt_chained_seeds would rescore seeds (bwa/bwamem.c:611)
  }
  S.seeds.clear(); S.chains.clear(); S.nodes.clear();
  S.root = bt_new_node(S, 0);
  int b = 0, e = 0, l_rep = 0;                                        // frac_rep, :273-280
  for (int64_t i = i0; i < i1; ++i) {
    const bwaflow_intv_t &p = J.intv[i];
    const int sb = (int)(p.info >> 32), se = (int)(uint32_t)p.info;
    if (p.x[2] <= (uint64_t)o.max_occ) continue;
    if (sb > e) { l_rep += e - b; b = sb; e = se; } else e = e > se ? e : se;
  }
  l_rep += e - b;
  int n_keys = 0;
  for (int64_t i = i0; i < i1; ++i) {                                 // :281-309
    const bwaflow_intv_t &p = J.intv[i];
    const int slen = (int)((uint32_t)p.info - (uint32_t)(p.info >> 32));
    for (int64_t t = J.pos_off[i]; t < J.pos_off[i + 1]; ++t) {
      bwaflow_seed_t s; s.rbeg = J.pos[t]; s.qbeg = (int)(p.info >> 32); s.len = s.score = slen; s._pad = 0;
      const int rid = intv2rid(J.ctg, s.rbeg, s.rbeg + s.len);
      if (rid < 0) continue;
      bool to_add = true;
      if (n_keys) {
        const int lower = bt_lower(S, s.rbeg);
        if (lower >= 0 && test_and_merge(S, o, J.ctg.l_pac, lower, s, rid)) to_add = false;
      }
      if (to_add) {
        SeedNode sn; sn.s = s; sn.next = -1; S.seeds.push_back(sn);
        Chain c; memset(&c, 0, sizeof(c));
        c.pos = s.rbeg; c.n = 1; c.rid = rid; c.head = c.tail = (int)S.seeds.size() - 1;
        c.is_alt = J.ctg.alt ? !!J.ctg.alt[rid] : 0;
        S.chains.push_back(c);
        bt_put(S, (int)S.chains.size() - 1);
        ++n_keys;
      }
    }
  }
  S.order.clear();
  bt_inorder(S, S.root, S.order);
  std::vector<Chain> &a = S.sorted; a.clear();
  for (int ci : S.order) a.push_back(S.chains[ci]);
  chain_flt(S, o, a);
  const float frac_rep = (float)l_rep / len;                          // :318
  for (const Chain &c : a) {
    bwaflow_chain_t f; f.pos = c.pos; f.n_seeds = c.n; f.rid = c.rid; f.frac_rep = frac_rep;
    f.flags = (uint32_t)c.w | (uint32_t)c.kept << 29 | (uint32_t)c.is_alt << 31; f.seed_off = (int64_t)os.size();
    for (int j = c.head; j >= 0; j = S.seeds[j].next) os.push_back(S.seeds[j].s);
    oc.push_back(f);
  }
}

static void worker(Job *J)
{
  Scratch S;
  for (;;) {
    const int64_t blk = J->next.fetch_add(1);
    const int64_t r0 = blk * BLOCK;
    if (r0 >= J->n_reads) break;
    const int64_t r1 = r0 + BLOCK < J->n_reads ? r0 + BLOCK : J->n_reads;
    std::vector<bwaflow_chain_t> &oc = J->out_chains[blk];
    std::vector<bwaflow_seed_t> &os = J->out_seeds[blk];
    for (int64_t r = r0; r < r1; ++r) {
      const size_t before = oc.size();
      int err = 0;
      chain_read(S, *J, r, oc, os, &err);
      if (err) J->err.store(err);
      J->n_chains_of_read[r] = (int64_t)(oc.size() - before);
    }
  }
}

} // namespace

struct bwaflow_chain_result {
  std::vector<bwaflow_chain_t> chains;
  std::vector<bwaflow_seed_t> seeds;
  std::vector<int64_t> chain_off;
};

extern "C" {

void bwaflow_chain_opt_default(bwaflow_chain_opt_t *o)
{ // mem_opt_init, bwa/bwamem.c:48-84
  o->w = 100; o->max_chain_gap = 10000; o->min_chain_weight = 0; o->max_chain_extend = 1 << 30;
  o->min_seed_len = 19; o->max_occ = 500; o->mask_level = 0.50f; o->drop_ratio = 0.50f;
}

int bwaflow_chain_build(const bwaflow_chain_opt_t *opt, int64_t l_pac, int32_t n_seqs, const int64_t *ctg_offset,
                        const int32_t *ctg_len, const int32_t *ctg_is_alt,
                        int64_t n_reads, const int64_t *read_off,
                        const bwaflow_intv_t *intv, const int64_t *intv_off, const int64_t *pos, const int64_t *pos_off,
                        int n_threads, bwaflow_chain_result_t **out)
{
  if (!opt || !out || (n_reads && (!read_off || !intv_off || !pos_off))) { bwaflow::set_error("bwaflow_chain_build: null argument"); return -22; }
  Job J;
  J.opt = opt; J.ctg = Ctg{n_seqs, ctg_offset, ctg_len, ctg_is_alt, l_pac};
  J.n_reads = n_reads; J.read_off = read_off; J.intv = intv; J.intv_off = intv_off; J.pos = pos; J.pos_off = pos_off;
  const int64_t n_blk = (n_reads + BLOCK - 1) / BLOCK;
  J.out_chains.resize(n_blk); J.out_seeds.resize(n_blk); J.n_chains_of_read.assign(n_reads, 0);
  if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > n_blk) n_threads = n_blk > 0 ? (int)n_blk : 1;
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(worker, &J);
  worker(&J);
  for (auto &t : th) t.join();
  if (J.err.load()) { bwaflow::set_error("read too long for host chaining: mem_flt_chained_seeds (bwa/bwamem.c:607) would rescore its seeds"); return -34; }
  bwaflow_chain_result *R = new bwaflow_chain_result();
  size_t nc = 0, ns = 0;
  for (int64_t b = 0; b < n_blk; ++b) { nc += J.out_chains[b].size(); ns += J.out_seeds[b].size(); }
  R->chains.reserve(nc); R->seeds.reserve(ns); R->chain_off.resize(n_reads + 1);
  for (int64_t b = 0; b < n_blk; ++b) {
    const int64_t sbase = (int64_t)R->seeds.size();
    for (bwaflow_chain_t c : J.out_chains[b]) { c.seed_off += sbase; R->chains.push_back(c); }
    R->seeds.insert(R->seeds.end(), J.out_seeds[b].begin(), J.out_seeds[b].end());
  }
  R->chain_off[0] = 0;
  for (int64_t r = 0; r < n_reads; ++r) R->chain_off[r + 1] = R->chain_off[r] + J.n_chains_of_read[r];
  *out = R;
  return 0;
}

void bwaflow_chain_result_view(const bwaflow_chain_result_t *r, const bwaflow_chain_t **chains, int64_t *n_chains,
                               const bwaflow_seed_t **seeds, int64_t *n_seeds, const int64_t **chain_off)
{
  if (chains) *chains = r->chains.data();
  if (n_chains) *n_chains = (int64_t)r->chains.size();
  if (seeds) *seeds = r->seeds.data();
  if (n_seeds) *n_seeds = (int64_t)r->seeds.size();
  if (chain_off) *chain_off = r->chain_off.data();
}

void bwaflow_chain_result_free(bwaflow_chain_result_t *r) { delete r; }

} // extern "C"
