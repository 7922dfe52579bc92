// Mate rescue for a batch of pairs: the rescue block of mem_sam_pe (src/bwa_wrapper.cpp:657-668) = mem_matesw
// (bwa/bwamem_pair.c:114-183) for the best regions of either end, with every ksw_align2 call (bwa/ksw.c:343-365) it can make run
// ahead on the device (sw.cu) and the calls themselves replayed on the host in the reference's order.  Also mem_pestat
// (bwa/bwamem_pair.c:49-112), which the batch needs first (src/Pipeline.cpp:578-579).
//
// Why "run ahead + replay" is exact: the window of a call depends only on the anchoring region (a copy taken before any rescue,
// src/bwa_wrapper.cpp:660-663), the orientation, the insert-size bounds and the mate's length -- never on the hits rescued so
// far.  What does depend on them is WHETHER an orientation is tried (skip[], bwamem_pair.c:121-127).  Most pairs already have a
// properly placed mate and skip everything, so the device runs ahead only the windows that are not skipped given the hits the
// pair starts with; the replay consumes them in the reference's order, and a pair whose replay asks for a window that was not
// run (an earlier rescue of the same pair changed the skip flags) is suspended, its window goes to the device with those of
// the other suspended pairs, and its replay is redone -- rounds until no pair is suspended (almost always one).
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "internal.cuh"
#include "stage5_host.h"
#include "batches.cuh"
#include "klib_sort.h"

namespace bwaflow {

struct MsFill { int64_t rb; int64_t q_src; int32_t l_ms, tlen, is_rev, _pad; int64_t q_off, t_off; };

__device__ __forceinline__ int ms_ref_base(const DevIndex &ix, int64_t p)
{ // bns_get_seq, bwa/bntseq.c:398-419
  const bool rev = p >= ix.l_pac;
  const int64_t f = rev ? (ix.l_pac << 1) - 1 - p : p;
  const int b = (__ldg(ix.pac + (f >> 2)) >> ((~f & 3) << 1)) & 3;
  return rev ? 3 - b : b;
}

// one warp per job: the mate (reverse-complemented when the orientation asks for it, bwamem_pair.c:135-139) and the window
__global__ void matesw_fill_kernel(const MsFill *f, int64_t n, const uint8_t *bases, DevIndex ix, uint8_t *q, uint8_t *t)
{
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const MsFill d = f[j];
  const uint8_t *ms = bases + d.q_src;
  for (int k = lane; k < d.l_ms; k += 32) {
    uint8_t c;
    if (d.is_rev) { c = ms[d.l_ms - 1 - k]; c = c < 4 ? 3 - c : 4; } else c = ms[k];
    q[d.q_off + k] = c;
  }
  for (int k = lane; k < d.tlen; k += 32) t[d.t_off + k] = (uint8_t)ms_ref_base(ix, d.rb + k);
}

// ---- host pieces -------------------------------------------------------------------------------------------------------------

// bns_pos2rid, bwa/bntseq.c:349-363
static int pos2rid(const bwaflow_ctx_t *c, int64_t pos_f)
{
  const int n = c->ix.n_seqs;
  if (pos_f >= c->ix.l_pac) return -1;
  int left = 0, mid = 0, right = n;
  while (left < right) {
    mid = (left + right) >> 1;
    if (pos_f >= c->h_ctg_off[mid]) {
      if (mid == n - 1) break;
      if (pos_f < c->h_ctg_off[mid + 1]) break;
      left = mid + 1;
    } else right = mid;
  }
  return mid;
}

// The window mem_matesw aligns the mate to for orientation r (bwamem_pair.c:133-150, bns_fetch_seq bwa/bntseq.c:421-446).
// Returns false when there is no ksw_align2 call for it.
static bool matesw_window(const bwaflow_ctx_t *c, const bwaflow_opt_t &o, const bwaflow_pestat_t *pes, const Reg &a, int l_ms, int r,
                          int *is_rev, int64_t *rb_, int64_t *re_)
{
  const int64_t l_pac = c->ix.l_pac;
  *is_rev = (r >> 1) != (r & 1);
  const int is_larger = !(r >> 1);
  int64_t rb, re;
  if (!*is_rev) {
    rb = is_larger ? a.rb + pes[r].low : a.rb - pes[r].high;
    re = (is_larger ? a.rb + pes[r].high : a.rb - pes[r].low) + l_ms;
  } else {
    rb = (is_larger ? a.rb + pes[r].low : a.rb - pes[r].high) - l_ms;
    re = is_larger ? a.rb + pes[r].high : a.rb - pes[r].low;
  }
  if (rb < 0) rb = 0;
  if (re > l_pac << 1) re = l_pac << 1;
  if (!(rb < re)) return false;
  const int64_t mid = (rb + re) >> 1;
  const int rev = mid >= l_pac;
  const int rid = pos2rid(c, rev ? (l_pac << 1) - 1 - mid : mid);
  int64_t fb = c->h_ctg_off[rid], fe = fb + c->h_ctg_len[rid];
  if (rev) { const int64_t t = fb; fb = (l_pac << 1) - fe; fe = (l_pac << 1) - t; }
  rb = rb > fb ? rb : fb;
  re = re < fe ? re : fe;
  if (!(a.rid == rid && re - rb >= o.min_seed_len)) return false;
  *rb_ = rb; *re_ = re;
  return true;
}

} // namespace bwaflow

using namespace bwaflow;

struct MsJobRef { uint32_t key; int64_t res; };            // key = end << 20 | anchor index << 2 | orientation
struct bwaflow_matesw_batch {
  bwaflow_ctx_t *ctx;
  bwaflow_pe_opt_t po;
  cudaStream_t stream;
  Timer tm;
  int64_t max_reads, max_bases, n_reads;
  uint8_t *d_bases;
  bwaflow_sw_batch_t *sw; int64_t sw_jobs, sw_q, sw_t;
  MsFill *d_fill; int64_t fill_cap;
  // host side
  std::vector<Reg> regs; std::vector<int64_t> reg_off, read_off;
  bwaflow_pestat_t pes[4];
  std::vector<bwaflow_sw_job_t> jobs; std::vector<MsFill> fills;       // the round being run
  std::vector<int64_t> job_pair; std::vector<uint32_t> job_key;
  int max_q, max_t;
  std::vector<bwaflow_sw_result_t> res;                                 // results of all rounds
  std::vector<std::vector<MsJobRef>> pair_jobs;                         // per pair: which windows have a result
  std::vector<uint8_t> pair_live;                                       // pair has windows to replay
  std::vector<std::vector<Reg>> pair_out[2];
  std::vector<Reg> out; std::vector<int64_t> out_off;
  int64_t n_sw, n_run; int rounds;
  float kernel_ms; double cells;
  bool pending;
};

// one window of the current round
static void ms_add_job(bwaflow_matesw_batch *b, int64_t pr, uint32_t key, int64_t q_src, int l_ms, int is_rev, int64_t rb, int64_t re)
{
  const bwaflow_opt_t &o = b->ctx->par.opt;
  int64_t qo = 0, to = 0;
  if (!b->jobs.empty()) { qo = b->jobs.back().q_off + b->jobs.back().qlen; to = b->jobs.back().t_off + b->jobs.back().tlen; }
  bwaflow_sw_job_t jb;
  jb.qlen = l_ms; jb.tlen = (int)(re - rb); jb._pad = 0; jb.q_off = qo; jb.t_off = to;
  jb.xtra = 0x40000 | 0x80000 | (l_ms * o.a < 250 ? 0x10000 : 0) | (o.min_seed_len * o.a);   // bwamem_pair.c:153
  MsFill f; f.rb = rb; f.q_src = q_src; f.l_ms = l_ms; f.tlen = jb.tlen; f.is_rev = is_rev; f._pad = 0; f.q_off = qo; f.t_off = to;
  b->jobs.push_back(jb); b->fills.push_back(f); b->job_pair.push_back(pr); b->job_key.push_back(key);
  if (l_ms > b->max_q) b->max_q = l_ms;
  if (jb.tlen > b->max_t) b->max_t = jb.tlen;
}

// the current round's jobs on the device: upload, fill the byte pools from the reads and the reference, ksw_align2, results back
static int ms_enqueue_round(bwaflow_matesw_batch *b)
{
  const int64_t nj = (int64_t)b->jobs.size();
  const int64_t qo = nj ? b->jobs.back().q_off + b->jobs.back().qlen : 0, to = nj ? b->jobs.back().t_off + b->jobs.back().tlen : 0;
  if (!b->sw || nj > b->sw_jobs || qo > b->sw_q || to > b->sw_t) {
    if (b->sw) bwaflow_sw_batch_destroy(b->sw);
    b->sw_jobs = nj * 2 + 1024; b->sw_q = qo * 2 + 4096; b->sw_t = to * 2 + 4096;
    int r = bwaflow_sw_batch_create(b->ctx, b->sw_jobs, b->sw_q, b->sw_t, &b->sw);
    if (r) return r;
  }
  if (nj > b->fill_cap) {
    cudaFree(b->d_fill);
    b->fill_cap = nj * 2 + 1024;
    BF_CUDA(cudaMalloc(&b->d_fill, b->fill_cap * sizeof(MsFill)));
  }
  const SwDeviceView v = sw_device_view(b->sw);
  if (nj) {
    BF_CUDA(cudaMemcpyAsync(v.jobs, b->jobs.data(), nj * sizeof(bwaflow_sw_job_t), cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_fill, b->fills.data(), nj * sizeof(MsFill), cudaMemcpyHostToDevice, b->stream));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  if (nj) {
    matesw_fill_kernel<<<(unsigned)((nj * 32 + 255) / 256), 256, 0, b->stream>>>(b->d_fill, nj, b->d_bases, b->ctx->ix, v.q, v.t);
    int r = sw_launch_device(b->sw, nj, b->max_q, b->max_t, b->stream);
    if (r) return r;
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[2], b->stream));
  const size_t base = b->res.size();
  b->res.resize(base + (size_t)nj);
  if (nj) BF_CUDA(cudaMemcpyAsync(b->res.data() + base, v.out, nj * sizeof(bwaflow_sw_result_t), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaMemcpyAsync(v.h_ctl, v.d_ctl, 4 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[3], b->stream));
  return 0;
}

// waits for the round and files its results under their pairs
static int ms_finish_round(bwaflow_matesw_batch *b)
{
  BF_CUDA(cudaStreamSynchronize(b->stream));
  const SwDeviceView v = sw_device_view(b->sw);
  BF_REQUIRE(v.h_ctl[1] == 0, -34, "mate rescue: a ksw_align2 job is outside the supported range (device error %d)", v.h_ctl[1]);
  const size_t nj = b->jobs.size(), base = b->res.size() - nj;
  for (size_t k = 0; k < nj; ++k) b->pair_jobs[(size_t)b->job_pair[k]].push_back(MsJobRef{b->job_key[k], (int64_t)(base + k)});
  float ms = 0;
  cudaEventElapsedTime(&ms, b->tm.ev[1], b->tm.ev[2]);
  b->kernel_ms += ms;
  b->cells += (double)*(unsigned long long *)(v.h_ctl + 2);
  b->n_run += (int64_t)nj;
  if (nj) ++b->rounds;
  b->jobs.clear(); b->fills.clear(); b->job_pair.clear(); b->job_key.clear(); b->max_q = b->max_t = 1;
  return 0;
}

// The rescue block of mem_sam_pe for pair pr (src/bwa_wrapper.cpp:657-668 with mem_matesw, bwa/bwamem_pair.c:114-183, inside).
// enumerate = true: nothing is consumed; every window the calls would run GIVEN THE HITS THE PAIR STARTS WITH is added to the
// round.  enumerate = false: the replay proper; returns false (pair suspended, its missing window added to the round) when a
// call needs a window that has no result yet.
static bool ms_pair(bwaflow_matesw_batch *b, int64_t pr, bool enumerate, std::vector<Reg> ma[2], std::vector<Reg> bb[2], int64_t *n_total)
{
  const bwaflow_opt_t &o = b->ctx->par.opt;
  const int64_t l_pac = b->ctx->ix.l_pac;
  const bwaflow_pestat_t *pes = b->pes;
  for (int i = 0; i < 2; ++i) {
    const Reg *a = b->regs.data() + b->reg_off[2 * pr + i];
    const int64_t na = b->reg_off[2 * pr + i + 1] - b->reg_off[2 * pr + i];
    ma[i].assign(a, a + na);
    bb[i].clear();
    for (int64_t j = 0; j < na; ++j)
      if (a[j].score >= a[0].score - b->po.pen_unpaired) bb[i].push_back(a[j]);   // src/bwa_wrapper.cpp:660-663
  }
  const std::vector<MsJobRef> &have = b->pair_jobs[(size_t)pr];
  int64_t total = 0;
  for (int i = 0; i < 2; ++i) {
    const int64_t mate = 2 * pr + (i ^ 1);
    const int l_ms = (int)(b->read_off[mate + 1] - b->read_off[mate]);
    std::vector<Reg> &m = ma[i ^ 1];
    for (size_t j = 0; j < bb[i].size() && (int)j < b->po.max_matesw; ++j) {
      const Reg &a = bb[i][j];
      // ---- mem_matesw, bwa/bwamem_pair.c:114-183
      int skip[4], n = 0;
      for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
      for (size_t k = 0; k < m.size(); ++k) {
        int64_t dist;
        const int r = infer_dir(l_pac, a.rb, m[k].rb, &dist);
        if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
      }
      if (skip[0] + skip[1] + skip[2] + skip[3] == 4) continue;
      for (int r = 0; r < 4; ++r) {
        if (skip[r]) continue;
        int is_rev; int64_t rb, re;
        if (matesw_window(b->ctx, o, pes, a, l_ms, r, &is_rev, &rb, &re)) {
          const uint32_t key = (uint32_t)i << 20 | (uint32_t)j << 2 | (uint32_t)r;
          if (enumerate) { ms_add_job(b, pr, key, b->read_off[mate], l_ms, is_rev, rb, re); continue; }
          int64_t ri = -1;
          for (const MsJobRef &h : have) if (h.key == key) { ri = h.res; break; }
          if (ri < 0) { ms_add_job(b, pr, key, b->read_off[mate], l_ms, is_rev, rb, re); return false; }
          const bwaflow_sw_result_t aln = b->res[(size_t)ri];
          if (aln.score >= o.min_seed_len && aln.qb >= 0) {
            Reg g;
            memset(&g, 0, sizeof(g));
            g.rid = a.rid;
            g.n_comp_is_alt = (uint32_t)reg_is_alt(a) << 30;
            g.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
            g.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
            g.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
            g.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
            g.score = aln.score;
            g.csub = aln.score2;
            g.secondary = -1;
            g.seedcov = (int)((g.re - g.rb < g.qe - g.qb ? g.re - g.rb : g.qe - g.qb) >> 1);
            m.push_back(g);
            size_t k;
            for (k = 0; k + 1 < m.size(); ++k)
              if (m[k].score < g.score) break;
            const size_t at = k;
            for (k = m.size() - 1; k > at; --k) m[k] = m[k - 1];
            m[at] = g;
          }
          ++n;
        }
        if (n && !enumerate) m.resize((size_t)sort_dedup(b->po, (int)m.size(), m.data()));
      }
      total += n;
    }
  }
  *n_total += total;
  return true;
}

extern "C" {

void bwaflow_pe_opt_default(bwaflow_pe_opt_t *po)
{ // mem_opt_init, bwa/bwamem.c:48-84
  memset(po, 0, sizeof(*po));
  po->pen_unpaired = 17; po->max_matesw = 50; po->max_ins = 10000; po->max_chain_gap = 10000;
  po->mask_level = 0.50f; po->mask_level_redun = 0.95f; po->no_rescue = 0;
  po->mapq_coef_len = 50.f; po->mapq_coef_fac = (int)log(50.);
}

int bwaflow_pestat(const bwaflow_opt_t *opt, const bwaflow_pe_opt_t *po, int64_t l_pac, int64_t n_reads, const bwaflow_alnreg_t *regs,
                   const int64_t *reg_off, bwaflow_pestat_t pes[4])
{
  BF_REQUIRE(opt && po && regs && reg_off && pes, -22, "bwaflow_pestat: null argument");
  const double MIN_RATIO = 0.8, MIN_DIR_RATIO = 0.05, OUTLIER_BOUND = 2.0, MAPPING_BOUND = 3.0, MAX_STDDEV = 4.0;
  const int MIN_DIR_CNT = 10;
  std::vector<uint64_t> isize[4];
  memset(pes, 0, 4 * sizeof(bwaflow_pestat_t));
  auto cal_sub = [&](const Reg *a, int64_t n) {           // bwamem_pair.c:35-47
    int64_t j;
    for (j = 1; j < n; ++j) {
      const int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
      const int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
      if (e_min > b_max) {
        const int min_l = a[j].qe - a[j].qb < a[0].qe - a[0].qb ? a[j].qe - a[j].qb : a[0].qe - a[0].qb;
        if (e_min - b_max >= min_l * po->mask_level) break;
      }
    }
    return j < n ? a[j].score : opt->min_seed_len * opt->a;
  };
  for (int64_t i = 0; i < n_reads >> 1; ++i) {
    const Reg *r0 = regs + reg_off[2 * i], *r1 = regs + reg_off[2 * i + 1];
    const int64_t n0 = reg_off[2 * i + 1] - reg_off[2 * i], n1 = reg_off[2 * i + 2] - reg_off[2 * i + 1];
    if (n0 == 0 || n1 == 0) continue;
    if (cal_sub(r0, n0) > MIN_RATIO * r0[0].score) continue;
    if (cal_sub(r1, n1) > MIN_RATIO * r1[0].score) continue;
    if (r0[0].rid != r1[0].rid) continue;
    int64_t is;
    const int dir = infer_dir(l_pac, r0[0].rb, r1[0].rb, &is);
    if (is && is <= po->max_ins) isize[dir].push_back((uint64_t)is);
  }
  for (int d = 0; d < 4; ++d) {
    bwaflow_pestat_t *r = &pes[d];
    std::vector<uint64_t> &q = isize[d];
    if ((int)q.size() < MIN_DIR_CNT) { r->failed = 1; continue; }
    std::sort(q.begin(), q.end());                           // ks_introsort_64: plain values, any sort gives the same array
    const int p25 = (int)q[(int)(.25 * q.size() + .499)], p50 = (int)q[(int)(.50 * q.size() + .499)], p75 = (int)q[(int)(.75 * q.size() + .499)];
    (void)p50;
    r->low = (int)(p25 - OUTLIER_BOUND * (p75 - p25) + .499);
    if (r->low < 1) r->low = 1;
    r->high = (int)(p75 + OUTLIER_BOUND * (p75 - p25) + .499);
    size_t i; int x;
    for (i = 0, x = 0, r->avg = 0; i < q.size(); ++i)
      if (q[i] >= (uint64_t)r->low && q[i] <= (uint64_t)r->high) { r->avg += q[i]; ++x; }
    r->avg /= x;
    for (i = 0, r->std = 0; i < q.size(); ++i)
      if (q[i] >= (uint64_t)r->low && q[i] <= (uint64_t)r->high) r->std += (q[i] - r->avg) * (q[i] - r->avg);
    r->std = sqrt(r->std / x);
    r->low = (int)(p25 - MAPPING_BOUND * (p75 - p25) + .499);
    r->high = (int)(p75 + MAPPING_BOUND * (p75 - p25) + .499);
    if (r->low > r->avg - MAX_STDDEV * r->std) r->low = (int)(r->avg - MAX_STDDEV * r->std + .499);
    if (r->high < r->avg + MAX_STDDEV * r->std) r->high = (int)(r->avg + MAX_STDDEV * r->std + .499);
    if (r->low < 1) r->low = 1;
  }
  size_t max = 0;
  for (int d = 0; d < 4; ++d) max = max > isize[d].size() ? max : isize[d].size();
  for (int d = 0; d < 4; ++d)
    if (pes[d].failed == 0 && isize[d].size() < max * MIN_DIR_RATIO) pes[d].failed = 1;
  return 0;
}

int bwaflow_matesw_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, const bwaflow_pe_opt_t *po,
                                bwaflow_matesw_batch_t **out)
{
  BF_REQUIRE(ctx && out, -22, "bwaflow_matesw_batch_create: null argument");
  BF_REQUIRE(ctx->has_index, -22, "bwaflow_matesw_batch_create: the context has no index");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_matesw_batch *b = new bwaflow_matesw_batch();
  b->ctx = ctx; b->max_reads = max_reads; b->max_bases = max_bases;
  if (po) b->po = *po; else bwaflow_pe_opt_default(&b->po);
  BF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (b->tm.init()) return -5;
  BF_CUDA(cudaMalloc(&b->d_bases, max_bases + 16));
  b->sw = nullptr; b->sw_jobs = b->sw_q = b->sw_t = 0; b->d_fill = nullptr; b->fill_cap = 0; b->pending = false;
  *out = b;
  return 0;
}

void bwaflow_matesw_batch_destroy(bwaflow_matesw_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  if (b->sw) bwaflow_sw_batch_destroy(b->sw);
  cudaFree(b->d_bases); cudaFree(b->d_fill);
  b->tm.destroy();
  cudaStreamDestroy(b->stream);
  delete b;
}

int bwaflow_matesw_submit(bwaflow_matesw_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off,
                          const bwaflow_alnreg_t *regs, const int64_t *reg_off, const bwaflow_pestat_t pes[4])
{
  BF_REQUIRE(b && bases && read_off && reg_off && pes && (regs || reg_off[n_reads] == 0), -22, "bwaflow_matesw_submit: null argument");
  BF_REQUIRE(n_reads >= 0 && n_reads <= b->max_reads && (n_reads & 1) == 0 && read_off[n_reads] <= b->max_bases, -7,
             "bwaflow_matesw_submit: batch capacity exceeded or odd read count");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  b->n_reads = n_reads;
  b->regs.assign(regs, regs + reg_off[n_reads]);
  b->reg_off.assign(reg_off, reg_off + n_reads + 1);
  b->read_off.assign(read_off, read_off + n_reads + 1);
  memcpy(b->pes, pes, sizeof(b->pes));
  const int64_t n_pairs = n_reads >> 1;
  b->pair_jobs.resize((size_t)n_pairs);
  for (auto &v : b->pair_jobs) v.clear();
  b->res.clear(); b->jobs.clear(); b->fills.clear(); b->job_pair.clear(); b->job_key.clear(); b->max_q = b->max_t = 1;
  b->n_run = 0; b->rounds = 0; b->kernel_ms = 0; b->cells = 0;
  // round 0: the windows the calls run given the hits every pair starts with
  if (!b->po.no_rescue) {
    std::vector<Reg> ma[2], bb[2];
    int64_t dummy = 0;
    const int64_t l_pac = b->ctx->ix.l_pac;
    for (int64_t pr = 0; pr < n_pairs; ++pr) {
      // most pairs: every anchoring region already has a mate hit at a proper distance in every orientation that is not
      // failed, so mem_matesw returns at once (bwamem_pair.c:121-127) -- decided here on the regions in place, without the
      // copies the general path takes
      bool quiet = true;
      for (int i = 0; i < 2 && quiet; ++i) {
        const Reg *a = b->regs.data() + reg_off[2 * pr + i], *m = b->regs.data() + reg_off[2 * pr + (i ^ 1)];
        const int64_t na = reg_off[2 * pr + i + 1] - reg_off[2 * pr + i], nm = reg_off[2 * pr + (i ^ 1) + 1] - reg_off[2 * pr + (i ^ 1)];
        int taken = 0;
        for (int64_t j = 0; j < na && taken < b->po.max_matesw && quiet; ++j) {
          if (!(a[j].score >= a[0].score - b->po.pen_unpaired)) continue;
          ++taken;
          int skip[4];
          for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
          for (int64_t k = 0; k < nm; ++k) {
            int64_t dist;
            const int r = infer_dir(l_pac, a[j].rb, m[k].rb, &dist);
            if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
          }
          if (skip[0] + skip[1] + skip[2] + skip[3] != 4) quiet = false;
        }
      }
      if (!quiet) ms_pair(b, pr, true, ma, bb, &dummy);
    }
  }
  // a pair without a window in round 0 rescues nothing, so its hits never change and no later call can ask for one: it is final
  b->pair_live.assign((size_t)n_pairs, 0);
  for (int64_t pr : b->job_pair) b->pair_live[(size_t)pr] = 1;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  BF_CUDA(cudaMemcpyAsync(b->d_bases, bases, read_off[n_reads], cudaMemcpyHostToDevice, b->stream));
  int r = ms_enqueue_round(b);
  if (r) return r;
  b->pending = true;
  return 0;
}

int bwaflow_matesw_collect(bwaflow_matesw_batch_t *b, const bwaflow_alnreg_t **regs_out, const int64_t **reg_off_out, int64_t *n_regs,
                           int64_t *n_sw, int64_t *n_jobs, bwaflow_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_matesw_collect: nothing submitted");
  b->pending = false;
  int r;
  if ((r = ms_finish_round(b))) return r;
  float h2d_ms = 0;
  cudaEventElapsedTime(&h2d_ms, b->tm.ev[0], b->tm.ev[1]);
  const int64_t n_pairs = b->n_reads >> 1;
  for (int i = 0; i < 2; ++i) b->pair_out[i].resize((size_t)n_pairs);
  std::vector<int64_t> todo, again;
  for (int64_t pr = 0; pr < n_pairs; ++pr) if (b->pair_live[(size_t)pr]) todo.push_back(pr);
  int64_t total = 0;
  std::vector<Reg> ma[2], bb[2];
  for (int round = 0; !todo.empty(); ++round) {
    BF_REQUIRE(round < 64, -5, "mate rescue: the replay did not settle");
    again.clear();
    for (int64_t pr : todo) {
      int64_t n = 0;
      if (ms_pair(b, pr, false, ma, bb, &n)) {
        total += n;
        for (int i = 0; i < 2; ++i) b->pair_out[i][(size_t)pr].swap(ma[i]);
      } else again.push_back(pr);
    }
    if (again.empty()) break;
    if ((r = ms_enqueue_round(b))) return r;               // the windows the suspended pairs asked for
    if ((r = ms_finish_round(b))) return r;
    todo.swap(again);
  }
  b->out.clear(); b->out_off.assign(1, 0);
  b->out.reserve(b->regs.size() + 64);
  if (b->n_run == 0) {                                     // nothing was rescued anywhere: the regions leave as they came
    b->out = b->regs;
    b->out_off = b->reg_off;
  } else
  for (int64_t pr = 0; pr < n_pairs; ++pr)
    for (int i = 0; i < 2; ++i) {
      if (b->pair_live[(size_t)pr]) {
        const std::vector<Reg> &v = b->pair_out[i][(size_t)pr];
        b->out.insert(b->out.end(), v.begin(), v.end());
      } else b->out.insert(b->out.end(), b->regs.begin() + b->reg_off[2 * pr + i], b->regs.begin() + b->reg_off[2 * pr + i + 1]);
      b->out_off.push_back((int64_t)b->out.size());
    }
  b->n_sw = total;
  if (t) {
    memset(t, 0, sizeof(*t));
    t->h2d_ms = h2d_ms; t->kernel_ms = b->kernel_ms; t->total_ms = h2d_ms + b->kernel_ms;
    t->n_launches = 2 * b->rounds;
    t->work_units = b->cells;
  }
  if (regs_out) *regs_out = b->out.data();
  if (reg_off_out) *reg_off_out = b->out_off.data();
  if (n_regs) *n_regs = (int64_t)b->out.size();
  if (n_sw) *n_sw = total;
  if (n_jobs) *n_jobs = b->n_run;
  return 0;
}

} // extern "C"
