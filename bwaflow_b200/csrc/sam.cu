// Stage 5 for one batch: RegionsToSam::compute (src/Pipeline.cpp:546-648) -- regions of stage 4 in, per-read SAM text out.
//
//   phase 1  mem_sort_dedup_patch (bwa/bwamem.c:446-498) per read; the global alignments mem_patch_reg asks for run on the device
//            (cigar.cu, score-only mode), suspended scans are redone when their scores are back; ALT marking (Pipeline.cpp:570-574)
//   phase 2  mem_pestat over the batch (bwa/bwamem_pair.c:49-112), mate rescue for every pair (matesw.cu: ksw_align2 on the device)
//   phase 3  per pair, the decisions of mem_sam_pe_hts (src/bwa_wrapper.cpp:640-790): mem_mark_primary_se, mem_pair, mapping
//            qualities, which regions become records and which are XA candidates (mem_gen_alt, bwa/bwamem_extra.c:99-152).  Every
//            mem_reg2aln call becomes a CIGAR job; what it derives from the region alone (flag, mapq, score, sub) is kept beside it
//   phase 4  the CIGAR jobs on the device (cigar.cu: ksw_global2 + traceback, NM / MD, clipping, position)
//   phase 5  mem_aln2sam_hts (src/bwa_wrapper.cpp:452-590) for every planned record
//
// Phases 1, 3 and 5 are host loops over reads / pairs on worker threads: each is a short sequential decision procedure whose result
// the next phase's device batch needs as a whole.  No dynamic programming runs on the host.
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include "internal.cuh"
#include "stage5_host.h"

using namespace bwaflow;

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// fn(lo, hi, t) over [0, n) split into n_thr contiguous slices
void par_for(int n_thr, int64_t n, const std::function<void(int64_t, int64_t, int)> &fn)
{
  if (n_thr <= 1 || n < 64) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_thr; ++t) th.emplace_back([&, t]() { fn(n * t / n_thr, n * (t + 1) / n_thr, t); });
  for (auto &x : th) x.join();
}

// ---- mem_patch_reg, bwa/bwamem.c:415-444, split where it calls bwa_gen_cigar2 --------------------------------------------------
struct PatchReq { int32_t qb, qe, w; int64_t rb, re; int32_t score; };
// the tests before the global alignment; false = no merge.  *w = the band handed to bwa_gen_cigar2
bool patch_pre(const bwaflow_opt_t &o, int64_t l_pac, const Reg &a, const Reg &b, int *w_out)
{
  if (a.rb < l_pac && b.rb >= l_pac) return false;
  if (a.qb >= b.qb || a.qe >= b.qe || a.re >= b.re) return false;
  int w = (int)((a.re - b.rb) - (a.qe - b.qb));
  w = w > 0 ? w : -w;
  double r = (double)(a.re - b.rb) / (b.re - a.rb) - (double)(a.qe - b.qb) / (b.qe - a.qb);
  r = r > 0. ? r : -r;
  if (a.re < b.rb || a.qe < b.qb) {
    if (w > o.w << 1 || r >= 0.05f) return false;
  } else if (w > o.w << 2 || r >= 0.05f * 2) return false;
  w += a.w + b.w;
  w = w < o.w << 2 ? w : o.w << 2;
  *w_out = w;
  return true;
}
// the test after it; returns the merged score or 0
int patch_post(const Reg &a, const Reg &b, int score)
{
  const int q_s = (int)((double)(b.qe - a.qb) / ((b.qe - b.qb) + (a.qe - a.qb)) * (b.score + a.score) + .499);
  const int r_s = (int)((double)(b.re - a.rb) / ((b.re - b.rb) + (a.re - a.rb)) * (b.score + a.score) + .499);
  if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
  return score;
}

struct DedupWork {                                  // per worker thread
  std::vector<Reg> tmp;
};

// Phase 1 for a whole batch.  out: regions per read at the input offsets (n_out[r] <= input count).
int dedup_batch(bwaflow_cigar_batch_t *cb, const bwaflow_ctx_t *ctx, const bwaflow_pe_opt_t &po, int n_thr, int64_t n_reads,
                const uint8_t *bases, const int64_t *read_off, const Reg *regs, const int64_t *reg_off, Reg *out, int32_t *n_out,
                int64_t *n_jobs_total, int32_t *rounds)
{
  const bwaflow_opt_t &o = ctx->par.opt;
  const int64_t l_pac = ctx->ix.l_pac;
  std::vector<int64_t> todo;
  for (int64_t r = 0; r < n_reads; ++r) {
    const int64_t n = reg_off[r + 1] - reg_off[r];
    if (n <= 1) { if (n == 1) out[reg_off[r]] = regs[reg_off[r]]; n_out[r] = (int32_t)n; }
    else todo.push_back(r);
  }
  std::vector<std::vector<PatchReq>> cache((size_t)n_reads);       // scores known so far, per read (empty for almost all)
  std::vector<PatchReq> want((size_t)n_reads);
  std::vector<uint8_t> susp((size_t)n_reads, 0);
  int n_rounds = 0, uploaded = 0;
  int64_t jobs_total = 0;
  while (!todo.empty()) {
    par_for(n_thr, (int64_t)todo.size(), [&](int64_t lo, int64_t hi, int) {
      for (int64_t x = lo; x < hi; ++x) {
        const int64_t r = todo[(size_t)x];
        const int n = (int)(reg_off[r + 1] - reg_off[r]);
        Reg *a = out + reg_off[r];
        memcpy(a, regs + reg_off[r], (size_t)n * sizeof(Reg));      // every (re)start is from the stage-4 order: the sorts are unstable
        const std::vector<PatchReq> &known = cache[(size_t)r];
        PatchReq miss;
        const int m = sort_dedup(po, n, a, [&](const Reg &q, const Reg &p, int *w) -> int {
          if (!patch_pre(o, l_pac, q, p, w)) return 0;
          for (const PatchReq &k : known)
            if (k.qb == q.qb && k.qe == p.qe && k.rb == q.rb && k.re == p.re && k.w == *w) return patch_post(q, p, k.score);
          miss.qb = q.qb; miss.qe = p.qe; miss.rb = q.rb; miss.re = p.re; miss.w = *w; miss.score = 0;
          return -1;
        });
        if (m < 0) { susp[(size_t)r] = 1; want[(size_t)r] = miss; }
        else { susp[(size_t)r] = 0; n_out[r] = m; }
      }
    });
    std::vector<int64_t> again;
    std::vector<bwaflow_gscore_job_t> jobs;
    for (int64_t r : todo)
      if (susp[(size_t)r]) {
        const PatchReq &q = want[(size_t)r];
        bwaflow_gscore_job_t j;
        j.read = (int32_t)r; j.qb = q.qb; j.qe = q.qe; j.w = q.w; j.rb = q.rb; j.re = q.re;
        jobs.push_back(j);
        again.push_back(r);
      }
    if (jobs.empty()) break;
    std::vector<int32_t> sc(jobs.size());
    const int64_t cap = cigar_max_jobs(cb);
    for (size_t s = 0; s < jobs.size(); s += (size_t)cap) {
      const int64_t n = (int64_t)std::min(jobs.size() - s, (size_t)cap);
      const int rc = bwaflow_cigar_global_scores(cb, n_reads, bases, read_off, !uploaded, n, jobs.data() + s, sc.data() + s);
      if (rc) return rc;
      uploaded = 1;
    }
    for (size_t k = 0; k < jobs.size(); ++k) {
      PatchReq q = want[(size_t)jobs[k].read];
      q.score = sc[k];
      cache[(size_t)jobs[k].read].push_back(q);
    }
    jobs_total += (int64_t)jobs.size();
    ++n_rounds;
    todo.swap(again);
  }
  if (n_jobs_total) *n_jobs_total = jobs_total;
  if (rounds) *rounds = n_rounds;
  return 0;
}

// ---- phase 3: records in the making ---------------------------------------------------------------------------------------------
struct AlnP {                         // what mem_reg2aln (bwa/bwamem.c:1104-1174) derives from the region alone + the job that carries the rest
  int32_t job;                        // CIGAR job of the chunk; -1 = the unmapped record (ar == 0)
  int32_t rid, flag, mapq, score, sub, alt_sc, is_alt;
  int32_t xa_lo, xa_hi;               // XA candidates: jobs xa[xa_lo .. xa_hi) of the chunk
};
struct ReadPlan { int32_t lo, hi; AlnP mate; int32_t has_mate; };     // records alns[lo .. hi), written against `mate`
struct Chunk {                        // one worker's slice of the batch
  std::vector<int32_t> job_read; std::vector<Reg> job_reg;
  std::vector<int32_t> xa;
  std::vector<AlnP> alns;
  std::vector<ReadPlan> plan;
  int64_t job_base, first_read;
  std::string text; std::vector<int64_t> text_off;
  std::vector<int> z; std::vector<Pair64> v, u;
};

struct Stage5 {                       // read-only context of phases 3 and 5
  const bwaflow_ctx_t *ctx; bwaflow_pe_opt_t po; bwaflow_sam_opt_t so;
  std::vector<std::string> names, anno; std::vector<uint8_t> is_alt; std::string rg_id;
};

AlnP plan_reg2aln(const Stage5 &S, Chunk &ck, int64_t read, const Reg *ar)
{
  AlnP t;
  memset(&t, 0, sizeof(t));
  if (ar == 0 || ar->rb < 0 || ar->re < 0) { t.job = -1; t.rid = -1; t.flag = 0x4; return t; }
  t.job = (int32_t)ck.job_read.size();
  ck.job_read.push_back((int32_t)read);
  ck.job_reg.push_back(*ar);
  t.rid = ar->rid;
  t.mapq = ar->secondary < 0 ? approx_mapq_se(S.ctx->par.opt, S.po, *ar) : 0;
  if (ar->secondary >= 0) t.flag |= 0x100;
  t.score = ar->score; t.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
  t.is_alt = reg_is_alt(*ar); t.alt_sc = ar->alt_sc;
  return t;
}

// get_pri_idx, bwa/bwamem_extra.c:90-95
inline int pri_idx(double xa_drop, const Reg *a, int i)
{
  const int k = a[i].secondary_all;
  if (k >= 0 && a[i].score >= a[k].score * xa_drop) return k;
  return -1;
}
// XA[k] of mem_gen_alt for one k: the hits whose primary is k, unless there are too many of them (:112-121, 123-126)
void plan_xa(const Stage5 &S, Chunk &ck, int64_t read, const Reg *a, int n, int k, AlnP &q)
{
  q.xa_lo = q.xa_hi = (int32_t)ck.xa.size();
  if (S.so.flag & 0x8) return;                                       // MEM_F_ALL: no XA
  int cnt = 0, has_alt = 0;
  for (int i = 0; i < n; ++i)
    if (pri_idx(S.so.XA_drop_ratio, a, i) == k) { ++cnt; if (reg_is_alt(a[i])) has_alt = 1; }
  if (cnt == 0 || cnt > S.so.max_XA_hits_alt || (!has_alt && cnt > S.so.max_XA_hits)) return;
  for (int i = 0; i < n; ++i)
    if (pri_idx(S.so.XA_drop_ratio, a, i) == k) ck.xa.push_back(plan_reg2aln(S, ck, read, &a[i]).job);
  q.xa_hi = (int32_t)ck.xa.size();
}

// mem_reg2sam_hts, src/bwa_wrapper.cpp:592-637
void plan_reg2sam(const Stage5 &S, Chunk &ck, int64_t read, const Reg *a, int n, int extra_flag, const AlnP *m)
{
  ReadPlan rp;
  rp.lo = (int32_t)ck.alns.size();
  int l = 0, mapq0 = 0;
  for (int k = 0; k < n; ++k) {
    const Reg *p = &a[k];
    if (p->score < S.so.T) continue;
    if (p->secondary >= 0 && (reg_is_alt(*p) || !(S.so.flag & 0x8))) continue;
    if (p->secondary >= 0 && p->secondary < 0x7fffffff && p->score < a[p->secondary].score * S.so.drop_ratio) continue;
    AlnP q = plan_reg2aln(S, ck, read, p);
    plan_xa(S, ck, read, a, n, k, q);
    q.flag |= extra_flag;
    if (p->secondary >= 0) q.sub = -1;
    if (l && p->secondary < 0) q.flag |= (S.so.flag & 0x10) ? 0x10000 : 0x800;
    if (l && !reg_is_alt(*p) && q.mapq > mapq0) q.mapq = mapq0;
    if (l == 0) mapq0 = q.mapq;
    ck.alns.push_back(q);
    ++l;
  }
  if (l == 0) {
    AlnP t = plan_reg2aln(S, ck, read, 0);
    t.flag |= extra_flag;
    ck.alns.push_back(t);
  }
  rp.hi = (int32_t)ck.alns.size();
  rp.has_mate = m ? 1 : 0;
  if (m) rp.mate = *m; else memset(&rp.mate, 0, sizeof(rp.mate));
  ck.plan.push_back(rp);
}

inline int raw_mapq(int diff, int a) { return (int)(6.02 * diff / a + .499); }

// mem_sam_pe_hts after its rescue block (src/bwa_wrapper.cpp:669-790) for the pair (read0, read0 + 1)
void plan_pair(const Stage5 &S, Chunk &ck, int64_t read0, Reg *const a[2], const int n[2], const bwaflow_pestat_t pes[4], uint64_t id)
{
  const bwaflow_opt_t &o = S.ctx->par.opt;
  const int64_t l_pac = S.ctx->ix.l_pac;
  int z[2] = {0, 0}, n_pri[2], extra_flag = 1, opair = 0, subo = 0, n_sub = 0;
  AlnP h[2];
  n_pri[0] = mark_primary_se(o, S.po.mask_level, n[0], a[0], id << 1 | 0, ck.z);
  n_pri[1] = mark_primary_se(o, S.po.mask_level, n[1], a[1], id << 1 | 1, ck.z);
  bool paired = false;
  if (!(S.so.flag & 0x4) && n_pri[0] && n_pri[1]) {
    const Reg *ca[2] = {a[0], a[1]};
    opair = pair_hits(o, l_pac, S.ctx->h_ctg_off, pes, ca, n_pri, id, &subo, &n_sub, z, ck.v, ck.u);
    if (opair > 0) {
      int is_multi[2];
      for (int i = 0; i < 2; ++i) {
        int j;
        for (j = 1; j < n_pri[i]; ++j)
          if (a[i][j].secondary < 0 && a[i][j].score >= S.so.T) break;
        is_multi[i] = j < n_pri[i] ? 1 : 0;
      }
      paired = !(is_multi[0] || is_multi[1]);
    }
  }
  if (paired) {
    int q_pe, q_se[2];
    const int score_un = a[0][0].score + a[1][0].score - S.po.pen_unpaired;
    subo = subo > score_un ? subo : score_un;
    q_pe = raw_mapq(opair - subo, o.a);
    if (n_sub > 0) q_pe -= (int)(4.343 * log((double)(n_sub + 1)) + .499);
    if (q_pe < 0) q_pe = 0;
    if (q_pe > 60) q_pe = 60;
    q_pe = (int)(q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
    if (opair > score_un) {                                 // the paired alignment is preferred
      Reg *c[2] = {&a[0][z[0]], &a[1][z[1]]};
      for (int i = 0; i < 2; ++i) {
        if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
        q_se[i] = approx_mapq_se(o, S.po, *c[i]);
      }
      q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
      q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
      extra_flag |= 2;
      for (int i = 0; i < 2; ++i) {                         // cap at the tandem repeat score
        const int cap = raw_mapq(c[i]->score - c[i]->csub, o.a);
        q_se[i] = q_se[i] < cap ? q_se[i] : cap;
      }
    } else {                                                // the unpaired alignment is preferred
      z[0] = z[1] = 0;
      q_se[0] = approx_mapq_se(o, S.po, a[0][0]);
      q_se[1] = approx_mapq_se(o, S.po, a[1][0]);
    }
    for (int i = 0; i < 2; ++i) {
      const int k = a[i][z[i]].secondary_all;
      if (k >= 0 && k < n_pri[i]) {                         // switch secondary and primary if both of them are non-ALT
        for (int j = 0; j < n[i]; ++j)
          if (a[i][j].secondary_all == k || j == k) a[i][j].secondary_all = z[i];
        a[i][z[i]].secondary_all = -1;
      }
    }
    int lo[2], hi[2];
    for (int i = 0; i < 2; ++i) {
      lo[i] = (int)ck.alns.size();
      h[i] = plan_reg2aln(S, ck, read0 + i, &a[i][z[i]]);
      h[i].mapq = q_se[i];
      h[i].flag |= 0x40 << i | extra_flag;
      plan_xa(S, ck, read0 + i, a[i], n[i], z[i], h[i]);
      ck.alns.push_back(h[i]);
      if (n_pri[i] < n[i]) {                                // the read has ALT hits
        const Reg *p = &a[i][n_pri[i]];
        if (!(p->score < S.so.T || p->secondary >= 0 || !reg_is_alt(*p))) {
          AlnP g = plan_reg2aln(S, ck, read0 + i, p);
          g.flag |= 0x800 | 0x40 << i | extra_flag;
          plan_xa(S, ck, read0 + i, a[i], n[i], n_pri[i], g);
          ck.alns.push_back(g);
        }
      }
      hi[i] = (int)ck.alns.size();
    }
    for (int i = 0; i < 2; ++i) {
      ReadPlan rp;
      rp.lo = lo[i]; rp.hi = hi[i]; rp.mate = h[!i]; rp.has_mate = 1;
      ck.plan.push_back(rp);
    }
    return;
  }
  // no_pairing (:768-789)
  for (int i = 0; i < 2; ++i) {
    int which = -1;
    if (n[i]) {
      if (a[i][0].score >= S.so.T) which = 0;
      else if (n_pri[i] < n[i] && a[i][n_pri[i]].score >= S.so.T) which = n_pri[i];
    }
    h[i] = plan_reg2aln(S, ck, read0 + i, which >= 0 ? &a[i][which] : 0);
  }
  if (!(S.so.flag & 0x4) && h[0].rid == h[1].rid && h[0].rid >= 0) {
    int64_t dist;
    const int d = infer_dir(l_pac, a[0][0].rb, a[1][0].rb, &dist);
    if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
  }
  plan_reg2sam(S, ck, read0, a[0], n[0], 0x41 | extra_flag, &h[1]);
  plan_reg2sam(S, ck, read0 + 1, a[1], n[1], 0x81 | extra_flag, &h[0]);
}

// ---- phase 5: text ----------------------------------------------------------------------------------------------------------------
struct AlnR {                         // a complete mem_aln_t
  int64_t pos; int rid, flag, is_rev, is_alt, mapq, NM, n_cigar, score, sub, alt_sc;
  const uint32_t *cigar; const char *md; int xa_lo, xa_hi;
};
struct CigarOut {                     // results of all CIGAR jobs of the batch
  std::vector<bwaflow_aln_t> aln; std::vector<uint32_t> cig; std::vector<char> md;
};

AlnR resolve(const AlnP &p, const Chunk &ck, const CigarOut &co)
{
  AlnR r;
  memset(&r, 0, sizeof(r));
  r.rid = p.rid; r.flag = p.flag; r.mapq = p.mapq; r.score = p.score; r.sub = p.sub; r.alt_sc = p.alt_sc; r.is_alt = p.is_alt;
  r.xa_lo = p.xa_lo; r.xa_hi = p.xa_hi;
  if (p.job < 0) { r.pos = -1; return r; }
  const bwaflow_aln_t &d = co.aln[(size_t)(ck.job_base + p.job)];
  r.pos = d.pos; r.is_rev = d.is_rev; r.NM = d.NM; r.n_cigar = d.n_cigar;
  r.cigar = co.cig.data() + d.cigar_off; r.md = co.md.data() + d.md_off;
  return r;
}

inline void put_int(std::string &s, long long v)
{
  char buf[24]; int n = 0;
  unsigned long long x = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
  do { buf[n++] = (char)('0' + x % 10); x /= 10; } while (x);
  if (v < 0) buf[n++] = '-';
  while (n) s.push_back(buf[--n]);
}
inline int ref_len(int n_cigar, const uint32_t *cigar)    // get_rlen, bwa/bwamem.c:810-819
{
  int l = 0;
  for (int k = 0; k < n_cigar; ++k) { const int op = cigar[k] & 0xf; if (op == 0 || op == 2) l += (int)(cigar[k] >> 4); }
  return l;
}

struct ReadText { const char *name; const uint8_t *seq; const uint8_t *qual; const char *comment; int l_seq; };

// mem_aln2sam_hts, src/bwa_wrapper.cpp:452-590 (text path: the line ends with '\n')
void format_aln(const Stage5 &S, std::string &str, const ReadText &s, const std::vector<AlnR> &list, int which, const AlnR *m_, const Chunk &ck,
                const CigarOut &co)
{
  const int n = (int)list.size();
  AlnR ptmp = list[(size_t)which], *p = &ptmp, mtmp, *m = 0;
  if (m_) { mtmp = *m_; m = &mtmp; }
  p->flag |= m ? 0x1 : 0;
  p->flag |= p->rid < 0 ? 0x4 : 0;
  p->flag |= m && m->rid < 0 ? 0x8 : 0;
  if (p->rid < 0 && m && m->rid >= 0) { p->rid = m->rid; p->pos = m->pos; p->is_rev = m->is_rev; p->n_cigar = 0; }
  if (m && m->rid < 0 && p->rid >= 0) { m->rid = p->rid; m->pos = p->pos; m->is_rev = p->is_rev; m->n_cigar = 0; }
  p->flag |= p->is_rev ? 0x10 : 0;
  p->flag |= m && m->is_rev ? 0x20 : 0;
  const bool soft = (S.so.flag & 0x200) != 0;

  str += s.name; str.push_back('\t');
  put_int(str, (p->flag & 0xffff) | (p->flag & 0x10000 ? 0x100 : 0)); str.push_back('\t');
  if (p->rid >= 0) {
    str += S.names[(size_t)p->rid]; str.push_back('\t');
    put_int(str, p->pos + 1); str.push_back('\t');
    put_int(str, p->mapq); str.push_back('\t');
    if (p->n_cigar) {
      for (int i = 0; i < p->n_cigar; ++i) {
        int c = p->cigar[i] & 0xf;
        if (!soft && !p->is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
        put_int(str, p->cigar[i] >> 4); str.push_back("MIDSH"[c]);
      }
    } else str.push_back('*');
  } else str += "*\t0\t0\t*";
  str.push_back('\t');

  if (m && m->rid >= 0) {
    if (p->rid == m->rid) str.push_back('=');
    else str += S.names[(size_t)m->rid];
    str.push_back('\t');
    put_int(str, m->pos + 1); str.push_back('\t');
    if (p->rid == m->rid) {
      const int64_t p0 = p->pos + (p->is_rev ? ref_len(p->n_cigar, p->cigar) - 1 : 0);
      const int64_t p1 = m->pos + (m->is_rev ? ref_len(m->n_cigar, m->cigar) - 1 : 0);
      if (m->n_cigar == 0 || p->n_cigar == 0) str.push_back('0');
      else put_int(str, -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
    } else str.push_back('0');
  } else str += "*\t0\t0";
  str.push_back('\t');

  if (p->flag & 0x100) str += "*\t*";
  else {
    int qb = 0, qe = s.l_seq;
    if (p->n_cigar && which && !soft && !p->is_alt) {
      const int c0 = p->cigar[0] & 0xf, c1 = p->cigar[p->n_cigar - 1] & 0xf;
      if (!p->is_rev) {
        if (c0 == 4 || c0 == 3) qb += (int)(p->cigar[0] >> 4);
        if (c1 == 4 || c1 == 3) qe -= (int)(p->cigar[p->n_cigar - 1] >> 4);
      } else {
        if (c0 == 4 || c0 == 3) qe -= (int)(p->cigar[0] >> 4);
        if (c1 == 4 || c1 == 3) qb += (int)(p->cigar[p->n_cigar - 1] >> 4);
      }
    }
    if (!p->is_rev) {
      for (int i = qb; i < qe; ++i) str.push_back("ACGTN"[s.seq[i] < 4 ? s.seq[i] : 4]);
      str.push_back('\t');
      if (s.qual) str.append((const char *)s.qual + qb, (size_t)(qe - qb));
      else str.push_back('*');
    } else {
      for (int i = qe - 1; i >= qb; --i) str.push_back("TGCAN"[s.seq[i] < 4 ? s.seq[i] : 4]);
      str.push_back('\t');
      if (s.qual) for (int i = qe - 1; i >= qb; --i) str.push_back((char)s.qual[i]);
      else str.push_back('*');
    }
  }

  if (p->n_cigar) { str += "\tNM:i:"; put_int(str, p->NM); str += "\tMD:Z:"; str += p->md; }
  if (p->score >= 0) { str += "\tAS:i:"; put_int(str, p->score); }
  if (p->sub >= 0) { str += "\tXS:i:"; put_int(str, p->sub); }
  if (!S.rg_id.empty()) { str += "\tRG:Z:"; str += S.rg_id; }
  if (!(p->flag & 0x100)) {
    int i;
    for (i = 0; i < n; ++i)
      if (i != which && !(list[(size_t)i].flag & 0x100)) break;
    if (i < n) {
      str += "\tSA:Z:";
      for (i = 0; i < n; ++i) {
        const AlnR *r = &list[(size_t)i];
        if (i == which || (r->flag & 0x100)) continue;
        str += S.names[(size_t)r->rid]; str.push_back(',');
        put_int(str, r->pos + 1); str.push_back(',');
        str.push_back("+-"[r->is_rev]); str.push_back(',');
        for (int k = 0; k < r->n_cigar; ++k) { put_int(str, r->cigar[k] >> 4); str.push_back("MIDSH"[r->cigar[k] & 0xf]); }
        str.push_back(','); put_int(str, r->mapq);
        str.push_back(','); put_int(str, r->NM);
        str.push_back(';');
      }
    }
    if (p->alt_sc > 0) { char buf[64]; snprintf(buf, sizeof(buf), "\tpa:f:%.3f", (double)p->score / p->alt_sc); str += buf; }
  }
  if (p->xa_hi > p->xa_lo) {                                // mem_gen_alt's string, bwa/bwamem_extra.c:127-142
    str += "\tXA:Z:";
    for (int x = p->xa_lo; x < p->xa_hi; ++x) {
      const bwaflow_aln_t &t = co.aln[(size_t)(ck.job_base + ck.xa[(size_t)x])];
      str += S.names[(size_t)t.rid]; str.push_back(',');
      str.push_back("+-"[t.is_rev]); put_int(str, t.pos + 1); str.push_back(',');
      const uint32_t *cg = co.cig.data() + t.cigar_off;
      for (int k = 0; k < t.n_cigar; ++k) { put_int(str, cg[k] >> 4); str.push_back("MIDSHN"[cg[k] & 0xf]); }
      str.push_back(','); put_int(str, t.NM);
      if (S.so.flag & 0x2000) { str.push_back(','); put_int(str, t.score); }
      str.push_back(';');
    }
  }
  if (s.comment) { str.push_back('\t'); str += s.comment; }
  if ((S.so.flag & 0x100) && p->rid >= 0 && (size_t)p->rid < S.anno.size() && !S.anno[(size_t)p->rid].empty()) {
    str += "\tXR:Z:";
    const size_t tmp = str.size();
    str += S.anno[(size_t)p->rid];
    for (size_t i = tmp; i < str.size(); ++i) if (str[i] == '\t') str[i] = ' ';
  }
  str.push_back('\n');
}

} // namespace

struct bwaflow_sam_batch {
  Stage5 S;
  bwaflow_ctx_t *ctx;
  int64_t max_reads, max_bases;
  int n_thr;
  bwaflow_cigar_batch_t *cb; bwaflow_matesw_batch_t *mb;
  std::vector<Reg> r1, r2; std::vector<int32_t> n1; std::vector<int64_t> off1, off2;
  std::vector<Chunk> chunks;
  CigarOut co;
  std::vector<int32_t> job_read; std::vector<Reg> job_reg;
  std::string sam; std::vector<int64_t> sam_off;
};

extern "C" {

void bwaflow_sam_opt_default(bwaflow_sam_opt_t *so)
{
  memset(so, 0, sizeof(*so));
  so->flag = 0x2 | 0x10; so->T = 30; so->drop_ratio = 0.50f; so->XA_drop_ratio = 0.80f; so->max_XA_hits = 5; so->max_XA_hits_alt = 200;
  so->n_threads = 0;
}

int bwaflow_dedup_patch(bwaflow_cigar_batch_t *cb, int64_t n_reads, const uint8_t *bases, const int64_t *read_off,
                        const bwaflow_alnreg_t *regs, const int64_t *reg_off, bwaflow_alnreg_t *regs_out, int64_t *reg_off_out,
                        int64_t *n_patch_jobs, int32_t *rounds)
{
  BF_REQUIRE(cb && bases && read_off && reg_off && regs_out && reg_off_out && (regs || reg_off[n_reads] == 0), -22, "bwaflow_dedup_patch: null argument");
  int n_thr = (int)std::thread::hardware_concurrency();
  n_thr = n_thr < 1 ? 1 : n_thr > 32 ? 32 : n_thr;
  std::vector<Reg> tmp((size_t)reg_off[n_reads] + 1);
  std::vector<int32_t> n_out((size_t)n_reads + 1);
  const int rc = dedup_batch(cb, cigar_ctx(cb), *cigar_po(cb), n_thr, n_reads, bases, read_off, regs, reg_off, tmp.data(), n_out.data(), n_patch_jobs, rounds);
  if (rc) return rc;
  int64_t w = 0;
  for (int64_t r = 0; r < n_reads; ++r) {
    reg_off_out[r] = w;
    for (int k = 0; k < n_out[(size_t)r]; ++k) regs_out[w++] = tmp[(size_t)(reg_off[r] + k)];
  }
  reg_off_out[n_reads] = w;
  return 0;
}

int bwaflow_sam_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, const bwaflow_pe_opt_t *po,
                             const bwaflow_sam_opt_t *so, const char *const *ctg_names, const uint8_t *ctg_is_alt,
                             const char *const *ctg_anno, const char *rg_id, bwaflow_sam_batch_t **out)
{
  BF_REQUIRE(ctx && out && ctg_names && max_reads > 0 && max_bases > 0, -22, "bwaflow_sam_batch_create: bad argument");
  BF_REQUIRE(ctx->has_index, -22, "bwaflow_sam_batch_create: the context has no index");
  bwaflow_sam_batch *b = new bwaflow_sam_batch();
  b->ctx = ctx; b->max_reads = max_reads; b->max_bases = max_bases; b->cb = nullptr; b->mb = nullptr;
  b->S.ctx = ctx;
  if (po) b->S.po = *po; else bwaflow_pe_opt_default(&b->S.po);
  if (so) b->S.so = *so; else bwaflow_sam_opt_default(&b->S.so);
  if (b->S.so.flag & 0x20) b->S.po.no_rescue = 1;
  const int n_seqs = ctx->ix.n_seqs;
  for (int i = 0; i < n_seqs; ++i) {
    b->S.names.push_back(ctg_names[i] ? ctg_names[i] : "");
    b->S.is_alt.push_back(ctg_is_alt ? ctg_is_alt[i] : 0);
    b->S.anno.push_back(ctg_anno && ctg_anno[i] ? ctg_anno[i] : "");
  }
  b->S.rg_id = rg_id ? rg_id : "";
  int n_thr = b->S.so.n_threads;
  if (n_thr < 1) n_thr = (int)std::thread::hardware_concurrency();
  b->n_thr = n_thr < 1 ? 1 : n_thr > 64 ? 64 : n_thr;
  int rc = bwaflow_cigar_batch_create(ctx, max_reads, max_bases, 2 * max_reads + 4096, &b->S.po, &b->cb);
  if (!rc) rc = bwaflow_matesw_batch_create(ctx, max_reads, max_bases, &b->S.po, &b->mb);
  if (rc) { bwaflow_sam_batch_destroy(b); return rc; }
  *out = b;
  return 0;
}

void bwaflow_sam_batch_destroy(bwaflow_sam_batch_t *b)
{
  if (!b) return;
  if (b->cb) bwaflow_cigar_batch_destroy(b->cb);
  if (b->mb) bwaflow_matesw_batch_destroy(b->mb);
  delete b;
}

int bwaflow_sam_process(bwaflow_sam_batch_t *b, int64_t n_reads, const uint8_t *bases, const uint8_t *quals, const int64_t *read_off,
                        const char *names, const int64_t *name_off, const char *comments, const int64_t *comment_off,
                        const bwaflow_alnreg_t *regs, const int64_t *reg_off, int64_t start_idx,
                        const char **sam, const int64_t **sam_off, bwaflow_pestat_t pes_out[4], bwaflow_sam_timing_t *timing)
{
  BF_REQUIRE(b && bases && read_off && names && name_off && reg_off && sam && sam_off, -22, "bwaflow_sam_process: null argument");
  BF_REQUIRE(n_reads >= 0 && n_reads <= b->max_reads && read_off[n_reads] <= b->max_bases, -7, "bwaflow_sam_process: batch capacity exceeded");
  const Stage5 &S = b->S;
  const bool pe = (S.so.flag & 0x2) != 0;
  BF_REQUIRE(!pe || (n_reads & 1) == 0, -22, "bwaflow_sam_process: a paired batch needs an even number of reads");
  const bwaflow_opt_t &o = b->ctx->par.opt;
  bwaflow_sam_timing_t tm;
  memset(&tm, 0, sizeof(tm));
  const double t0 = now_ms();
  // ---- phase 1: mem_sort_dedup_patch + ALT marking
  const int64_t n_regs = reg_off[n_reads];
  b->r1.resize((size_t)n_regs + 1); b->n1.resize((size_t)n_reads + 1);
  int32_t rounds = 0;
  int rc = dedup_batch(b->cb, b->ctx, S.po, b->n_thr, n_reads, bases, read_off, regs, reg_off, b->r1.data(), b->n1.data(), &tm.n_patch_jobs, &rounds);
  if (rc) return rc;
  b->off1.resize((size_t)n_reads + 1);
  {
    int64_t w = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
      b->off1[(size_t)r] = w;
      for (int k = 0; k < b->n1[(size_t)r]; ++k) {
        Reg x = b->r1[(size_t)(reg_off[r] + k)];
        if (x.rid >= 0 && S.is_alt[(size_t)x.rid]) x.n_comp_is_alt = (x.n_comp_is_alt & 0x3FFFFFFFu) | (1u << 30);
        b->r1[(size_t)w++] = x;                              // w <= reg_off[r] + k: compaction in place
      }
    }
    b->off1[(size_t)n_reads] = w;
  }
  const double t1 = now_ms();
  tm.dedup_ms = (float)(t1 - t0);
  // ---- phase 2: mem_pestat + the rescue block of every pair
  bwaflow_pestat_t pes[4];
  memset(pes, 0, sizeof(pes));
  Reg *cur = b->r1.data();
  const int64_t *cur_off = b->off1.data();
  if (pe) {
    rc = bwaflow_pestat(&o, &S.po, b->ctx->ix.l_pac, n_reads, b->r1.data(), b->off1.data(), pes);
    if (rc) return rc;
    if (!S.po.no_rescue) {
      rc = bwaflow_matesw_submit(b->mb, n_reads, bases, read_off, b->r1.data(), b->off1.data(), pes);
      if (rc) return rc;
      const bwaflow_alnreg_t *mr; const int64_t *mo; int64_t mn, n_sw, n_jobs;
      bwaflow_timing_t mt;
      rc = bwaflow_matesw_collect(b->mb, &mr, &mo, &mn, &n_sw, &n_jobs, &mt);
      if (rc) return rc;
      tm.n_matesw_jobs = n_jobs; tm.n_matesw_used = n_sw; tm.matesw_kernel_ms = mt.kernel_ms;
      b->r2.assign(mr, mr + mn);
      b->off2.assign(mo, mo + n_reads + 1);
      cur = b->r2.data(); cur_off = b->off2.data();
    }
  }
  if (pes_out) memcpy(pes_out, pes, sizeof(pes));
  const double t2 = now_ms();
  tm.matesw_ms = (float)(t2 - t1);
  // ---- phase 3: decisions -> planned records + CIGAR jobs, one chunk of pairs (reads) per worker
  const int64_t n_units = pe ? n_reads / 2 : n_reads;
  const int n_chunks = n_units < 64 ? 1 : b->n_thr;
  b->chunks.resize((size_t)n_chunks);
  par_for(n_chunks, n_units, [&](int64_t lo, int64_t hi, int t) {
    Chunk &ck = b->chunks[(size_t)t];
    ck.job_read.clear(); ck.job_reg.clear(); ck.xa.clear(); ck.alns.clear(); ck.plan.clear(); ck.text.clear(); ck.text_off.clear();
    ck.first_read = pe ? lo * 2 : lo;
    for (int64_t u = lo; u < hi; ++u) {
      if (pe) {
        const int64_t r0 = u * 2;
        Reg *a[2] = {cur + cur_off[r0], cur + cur_off[r0 + 1]};
        const int n[2] = {(int)(cur_off[r0 + 1] - cur_off[r0]), (int)(cur_off[r0 + 2] - cur_off[r0 + 1])};
        plan_pair(S, ck, r0, a, n, pes, (uint64_t)(start_idx >> 1) + (uint64_t)u);
      } else {
        Reg *a = cur + cur_off[u];
        const int n = (int)(cur_off[u + 1] - cur_off[u]);
        mark_primary_se(o, S.po.mask_level, n, a, (uint64_t)start_idx + (uint64_t)u, ck.z);
        plan_reg2sam(S, ck, u, a, n, 0, 0);
      }
    }
  });
  b->job_read.clear(); b->job_reg.clear();
  for (Chunk &ck : b->chunks) {
    ck.job_base = (int64_t)b->job_read.size();
    b->job_read.insert(b->job_read.end(), ck.job_read.begin(), ck.job_read.end());
    b->job_reg.insert(b->job_reg.end(), ck.job_reg.begin(), ck.job_reg.end());
  }
  const double t3 = now_ms();
  tm.pair_ms = (float)(t3 - t2);
  // ---- phase 4: CIGAR jobs on the device, in slices of the batch's capacity
  const int64_t n_jobs = (int64_t)b->job_read.size();
  tm.n_cigar_jobs = n_jobs;
  CigarOut &co = b->co;
  co.aln.clear(); co.cig.clear(); co.md.clear();
  co.aln.reserve((size_t)n_jobs);
  const int64_t job_cap = cigar_max_jobs(b->cb);
  for (int64_t s = 0; s < n_jobs; s += job_cap) {
    const int64_t n = std::min(n_jobs - s, job_cap);
    rc = bwaflow_cigar_submit(b->cb, n_reads, bases, read_off, n, b->job_read.data() + s, b->job_reg.data() + s);
    if (rc) return rc;
    const bwaflow_aln_t *al; const uint32_t *cg; const char *md;
    bwaflow_timing_t ct;
    rc = bwaflow_cigar_collect(b->cb, &al, &cg, &md, &ct);
    if (rc) return rc;
    tm.cigar_kernel_ms += ct.kernel_ms;
    const int64_t cig_base = (int64_t)co.cig.size(), md_base = (int64_t)co.md.size();
    int64_t n_cig = 0, n_md = 0;
    for (int64_t j = 0; j < n; ++j) {
      bwaflow_aln_t x = al[j];
      n_cig = std::max(n_cig, x.cigar_off + x.n_cigar);
      n_md = std::max(n_md, x.md_off + (int64_t)strlen(md + x.md_off) + 1);
      x.cigar_off += cig_base; x.md_off += md_base;
      co.aln.push_back(x);
    }
    co.cig.insert(co.cig.end(), cg, cg + n_cig);
    co.md.insert(co.md.end(), md, md + n_md);
  }
  co.cig.push_back(0); co.md.push_back(0);
  const double t4 = now_ms();
  tm.cigar_ms = (float)(t4 - t3);
  // ---- phase 5: text, each worker over the chunk it planned
  par_for(n_chunks, n_units, [&](int64_t lo, int64_t hi, int t) {
    Chunk &ck = b->chunks[(size_t)t];
    std::vector<AlnR> list;
    const int64_t nr = (int64_t)ck.plan.size();
    ck.text_off.assign((size_t)nr + 1, 0);
    for (int64_t k = 0; k < nr; ++k) {
      const ReadPlan &rp = ck.plan[(size_t)k];
      const int64_t r = ck.first_read + k;
      ReadText s;
      s.name = names + name_off[r]; s.seq = bases + read_off[r]; s.qual = quals ? quals + read_off[r] : 0;
      s.comment = comments && comment_off && comment_off[r] >= 0 ? comments + comment_off[r] : 0;
      s.l_seq = (int)(read_off[r + 1] - read_off[r]);
      list.clear();
      for (int x = rp.lo; x < rp.hi; ++x) list.push_back(resolve(ck.alns[(size_t)x], ck, co));
      AlnR m;
      if (rp.has_mate) m = resolve(rp.mate, ck, co);
      for (int x = 0; x < (int)list.size(); ++x) format_aln(S, ck.text, s, list, x, rp.has_mate ? &m : 0, ck, co);
      ck.text_off[(size_t)k + 1] = (int64_t)ck.text.size();
    }
  });
  b->sam.clear();
  b->sam_off.assign((size_t)n_reads + 1, 0);
  {
    size_t total = 0;
    for (const Chunk &ck : b->chunks) total += ck.text.size();
    b->sam.reserve(total + 1);
    for (const Chunk &ck : b->chunks) {
      const int64_t base = (int64_t)b->sam.size();
      for (size_t k = 0; k + 1 < ck.text_off.size(); ++k) b->sam_off[(size_t)ck.first_read + k + 1] = base + ck.text_off[k + 1];
      b->sam += ck.text;
    }
  }
  const double t5 = now_ms();
  tm.text_ms = (float)(t5 - t4);
  tm.total_ms = (float)(t5 - t0);
  tm.sam_bytes = (int64_t)b->sam.size();
  if (timing) *timing = tm;
  *sam = b->sam.data(); *sam_off = b->sam_off.data();
  return 0;
}

int64_t bwaflow_sam_header(bwaflow_sam_batch_t *b, const char *hdr_line, const char *pg_line, char *buf, int64_t cap)
{
  if (!b) return -22;
  std::string s = "@HD\tVN:1.3\tSO:coordinate\n";
  for (size_t i = 0; i < b->S.names.size(); ++i) {
    s += "@SQ\tSN:"; s += b->S.names[i]; s += "\tLN:"; put_int(s, b->ctx->h_ctg_len[i]); s.push_back('\n');
  }
  if (hdr_line && hdr_line[0]) { s += hdr_line; s.push_back('\n'); }
  if (pg_line && pg_line[0]) { s += pg_line; s.push_back('\n'); }
  if (buf && cap > 0) memcpy(buf, s.data(), (size_t)std::min<int64_t>(cap, (int64_t)s.size()));
  return (int64_t)s.size();
}

} // extern "C"
