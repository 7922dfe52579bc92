// Host halves of stage 5 (RegionsToSam::compute, src/Pipeline.cpp:546-648) shared by the mate-rescue batch (matesw.cu) and the SAM
// batch (sam.cu): the small sequential decisions the reference makes per read or per pair -- region de-duplication, primary marking,
// pairing, mapping quality.  Every dynamic-programming call they depend on (ksw_align2, ksw_global2) runs on the device; these
// functions only consume its results.  Double arithmetic (log, erfc) stays on the host so that it is the reference's libm.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>
#include "../../include/bwaflow_cuda.h"
#include "klib_sort.h"

#ifdef __CUDACC__
#define S5_HD __host__ __device__
#else
#define S5_HD
#endif

namespace bwaflow {

typedef bwaflow_alnreg_t Reg;
S5_HD static inline int reg_is_alt(const Reg &r) { return (int)(r.n_comp_is_alt >> 30); }
static inline int reg_n_comp(const Reg &r) { return (int)(r.n_comp_is_alt & 0x3FFFFFFFu); }
static inline void reg_set_n_comp(Reg &r, int n) { r.n_comp_is_alt = (r.n_comp_is_alt & 0xC0000000u) | ((uint32_t)n & 0x3FFFFFFFu); }

// mem_infer_dir, bwa/bwamem_pair.c:26-33
static inline int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
  const int r1 = b1 >= l_pac, r2 = b2 >= l_pac;
  const int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
  *dist = p2 > b1 ? p2 - b1 : b1 - p2;
  return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

struct RegEndLt { S5_HD bool operator()(const Reg &a, const Reg &b) const { return a.re < b.re; } };                       // alnreg_slt2, bwa/bwamem.c:400
struct RegScoreLt { S5_HD bool operator()(const Reg &a, const Reg &b) const
  { return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb))); } };          // alnreg_slt, :403
struct RegHashLt { S5_HD bool operator()(const Reg &a, const Reg &b) const                                                 // alnreg_hlt, :406
  { return a.score > b.score || (a.score == b.score && (reg_is_alt(a) < reg_is_alt(b) || (reg_is_alt(a) == reg_is_alt(b) && a.hash < b.hash))); } };
struct RegHash2Lt { S5_HD bool operator()(const Reg &a, const Reg &b) const                                                // alnreg_hlt2, :409
  { return reg_is_alt(a) < reg_is_alt(b) || (reg_is_alt(a) == reg_is_alt(b) && (a.score > b.score || (a.score == b.score && a.hash < b.hash))); } };

// mem_sort_dedup_patch, bwa/bwamem.c:446-498.  `patch(q, p, &w)` stands for mem_patch_reg (:415-444): the score of the merged hit
// (> 0), 0 for "do not merge", or a negative value when the answer is not available yet -- the scan is then abandoned and the
// function returns -1 (the caller restarts it from the unsorted input once the device has produced the score).
template <class Patch>
static int sort_dedup(const bwaflow_pe_opt_t &po, int n, Reg *a, Patch patch)
{
  int m, i, j;
  if (n <= 1) return n;
  klib_introsort(n, a, RegEndLt());
  for (i = 0; i < n; ++i) reg_set_n_comp(a[i], 1);
  for (i = 1; i < n; ++i) {
    Reg *p = &a[i];
    if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + po.max_chain_gap) continue;
    for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + po.max_chain_gap; --j) {
      Reg *q = &a[j];
      int score, w;
      if (q->qe == q->qb) continue;
      const int64_t orr = q->re - p->rb;
      const int64_t oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
      const int64_t mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
      const int64_t mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
      if (orr > po.mask_level_redun * mr && oq > po.mask_level_redun * mq) {
        if (p->score < q->score) { p->qe = p->qb; break; }
        else q->qe = q->qb;
      } else if (q->rb < p->rb && (score = patch(*q, *p, &w)) != 0) {
        if (score < 0) return -1;
        reg_set_n_comp(*p, reg_n_comp(*p) + reg_n_comp(*q) + 1);
        p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
        p->sub = p->sub > q->sub ? p->sub : q->sub;
        p->csub = p->csub > q->csub ? p->csub : q->csub;
        p->qb = q->qb; p->rb = q->rb;
        p->truesc = p->score = score;
        p->w = w;
        q->qb = q->qe;
      }
    }
  }
  for (i = 0, m = 0; i < n; ++i)
    if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
  n = m;
  klib_introsort(n, a, RegScoreLt());
  for (i = 1; i < n; ++i)
    if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
  for (i = 1, m = 1; i < n; ++i)
    if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
  return m;
}
struct NoPatch { int operator()(const Reg &, const Reg &, int *) const { return 0; } };   // how mem_matesw calls it (bns = pac = query = 0)
static inline int sort_dedup(const bwaflow_pe_opt_t &po, int n, Reg *a) { return sort_dedup(po, n, a, NoPatch()); }

// hash_64 of bwa/utils.h:98-109 (Thomas Wang's 64-bit integer mix): fixes the order of equal-score hits
static inline uint64_t mix64(uint64_t key)
{
  key += ~(key << 32); key ^= key >> 22; key += ~(key << 13); key ^= key >> 8;
  key += key << 3; key ^= key >> 15; key += ~(key << 27); key ^= key >> 31;
  return key;
}

// mem_approx_mapq_se, bwa/bwamem.c:967-991
static inline int approx_mapq_se(const bwaflow_opt_t &o, const bwaflow_pe_opt_t &po, const Reg &a)
{
  int mapq, l, sub = a.sub ? a.sub : o.min_seed_len * o.a;
  sub = a.csub > sub ? a.csub : sub;
  if (sub >= a.score) return 0;
  l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
  const double identity = 1. - (double)(l * o.a - a.score) / (o.a + o.b) / l;
  if (a.score == 0) mapq = 0;
  else if (po.mapq_coef_len > 0) {
    double tmp = l < po.mapq_coef_len ? 1. : po.mapq_coef_fac / log((double)l);
    tmp *= identity * identity;
    mapq = (int)(6.02 * (a.score - sub) / o.a * tmp * tmp + .499);
  } else {
    mapq = (int)(30.0 * (1. - (double)sub / a.score) * log((double)a.seedcov) + .499);
    mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
  }
  if (a.sub_n > 0) mapq -= (int)(4.343 * log((double)(a.sub_n + 1)) + .499);
  if (mapq > 60) mapq = 60;
  if (mapq < 0) mapq = 0;
  mapq = (int)(mapq * (1. - a.frac_rep) + .499);
  return mapq;
}

// mem_mark_primary_se_core, bwa/bwamem.c:502-528
static inline void mark_primary_core(const bwaflow_opt_t &o, float mask_level, int n, Reg *a, std::vector<int> &z)
{
  int tmp = o.a + o.b;
  tmp = o.o_del + o.e_del > tmp ? o.o_del + o.e_del : tmp;
  tmp = o.o_ins + o.e_ins > tmp ? o.o_ins + o.e_ins : tmp;
  z.clear();
  z.push_back(0);
  for (int i = 1; i < n; ++i) {
    size_t k;
    for (k = 0; k < z.size(); ++k) {
      const int j = z[k];
      const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
      const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
      if (e_min > b_max) {
        const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
        if (e_min - b_max >= min_l * mask_level) {
          if (a[j].sub == 0) a[j].sub = a[i].score;
          if (a[j].score - a[i].score <= tmp && (reg_is_alt(a[j]) || !reg_is_alt(a[i]))) ++a[j].sub_n;
          break;
        }
      }
    }
    if (k == z.size()) z.push_back(i);
    else a[i].secondary = z[k];
  }
}

// mem_mark_primary_se, bwa/bwamem.c:530-571; returns n_pri
static inline int mark_primary_se(const bwaflow_opt_t &o, float mask_level, int n, Reg *a, uint64_t id, std::vector<int> &z)
{
  int i, n_pri;
  if (n == 0) return 0;
  for (i = n_pri = 0; i < n; ++i) {
    a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = mix64(id + (uint64_t)i);
    if (!reg_is_alt(a[i])) ++n_pri;
  }
  klib_introsort(n, a, RegHashLt());
  mark_primary_core(o, mask_level, n, a, z);
  for (i = 0; i < n; ++i) {
    Reg *p = &a[i];
    p->secondary_all = i;
    if (!reg_is_alt(*p) && p->secondary >= 0 && reg_is_alt(a[p->secondary])) p->alt_sc = a[p->secondary].score;
  }
  if (n_pri >= 0 && n_pri < n) {
    z.resize((size_t)n);
    if (n_pri > 0) klib_introsort(n, a, RegHash2Lt());
    for (i = 0; i < n; ++i) z[(size_t)a[i].secondary_all] = i;
    for (i = 0; i < n; ++i) {
      if (a[i].secondary >= 0) {
        a[i].secondary_all = z[(size_t)a[i].secondary];
        if (reg_is_alt(a[i])) a[i].secondary = 0x7fffffff;
      } else a[i].secondary_all = -1;
    }
    if (n_pri > 0) {
      for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
      mark_primary_core(o, mask_level, n_pri, a, z);
    }
  } else {
    for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
  }
  return n_pri;
}

// mem_pair, bwa/bwamem_pair.c:185-247.  a[r] / n_pri[r]: the regions of read r of the pair after mem_mark_primary_se; ctg_off =
// bns->anns[].offset.  Returns the score of the best pair (0: none), z[] = its two region indices.
struct Pair64 { uint64_t x, y; };
struct Pair64Lt { S5_HD bool operator()(const Pair64 &a, const Pair64 &b) const { return a.x < b.x || (a.x == b.x && a.y < b.y); } };
static inline int pair_hits(const bwaflow_opt_t &o, int64_t l_pac, const int64_t *ctg_off, const bwaflow_pestat_t pes[4], const Reg *const a[2],
                            const int n_pri[2], uint64_t id64, int *sub, int *n_sub, int z[2], std::vector<Pair64> &v, std::vector<Pair64> &u)
{
  const int id = (int)id64;                                 // the reference narrows the pair id to int here (:185)
  const uint64_t idmix = (uint64_t)(int64_t)(int32_t)((uint32_t)id << 8);
  int r, i, k, y[4], ret;
  v.clear(); u.clear();
  for (r = 0; r < 2; ++r)
    for (i = 0; i < n_pri[r]; ++i) {
      Pair64 key;
      const Reg *e = &a[r][i];
      key.x = (uint64_t)(e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb);
      key.x = (uint64_t)e->rid << 32 | (key.x - (uint64_t)ctg_off[e->rid]);
      key.y = (uint64_t)e->score << 32 | (uint64_t)(i << 2 | (e->rb >= l_pac) << 1 | r);
      v.push_back(key);
    }
  klib_introsort((int64_t)v.size(), v.data(), Pair64Lt());
  y[0] = y[1] = y[2] = y[3] = -1;
  for (i = 0; i < (int)v.size(); ++i) {
    for (r = 0; r < 2; ++r) {
      const int dir = r << 1 | (int)(v[i].y >> 1 & 1);
      if (pes[dir].failed) continue;
      const int which = r << 1 | (int)((v[i].y & 1) ^ 1);
      if (y[which] < 0) continue;
      for (k = y[which]; k >= 0; --k) {
        if ((int)(v[k].y & 3) != which) continue;
        const int64_t dist = (int64_t)v[i].x - (int64_t)v[k].x;
        if (dist > pes[dir].high) break;
        if (dist < pes[dir].low) continue;
        const double ns = (dist - pes[dir].avg) / pes[dir].std;
        int q = (int)((v[i].y >> 32) + (v[k].y >> 32) + .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * o.a + .499);
        if (q < 0) q = 0;
        Pair64 p;
        p.y = (uint64_t)k << 32 | (uint64_t)i;
        p.x = (uint64_t)q << 32 | (mix64(p.y ^ idmix) & 0xffffffffU);
        u.push_back(p);
      }
    }
    y[v[i].y & 3] = i;
  }
  if (!u.empty()) {
    int tmp = o.a + o.b;
    tmp = tmp > o.o_del + o.e_del ? tmp : o.o_del + o.e_del;
    tmp = tmp > o.o_ins + o.e_ins ? tmp : o.o_ins + o.e_ins;
    klib_introsort((int64_t)u.size(), u.data(), Pair64Lt());
    const size_t n = u.size();
    i = (int)(u[n - 1].y >> 32); k = (int)(u[n - 1].y << 32 >> 32);
    z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
    z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
    ret = (int)(u[n - 1].x >> 32);
    *sub = n > 1 ? (int)(u[n - 2].x >> 32) : 0;
    *n_sub = 0;
    for (long x = (long)n - 2; x >= 0; --x)
      if (*sub - (int)(u[(size_t)x].x >> 32) <= tmp) ++*n_sub;
  } else { ret = 0; *sub = 0; *n_sub = 0; }
  return ret;
}

} // namespace bwaflow
