// Batched ksw_align2 jobs (bwa/ksw.c:343-365): local Smith-Waterman of a whole query against a target window with the second-best
// score and, by a reversed second pass, the start of the best hit -- what mem_matesw runs for every mate-rescue window
// (bwa/bwamem_pair.c:150-154).
//
// The reference computes this with Farrar's striped SSE2 kernels ksw_u8 / ksw_i16 (bwa/ksw.c:110-339), and two of their
// properties show in the output, so the device kernel reproduces the STRIPED computation itself rather than a textbook matrix:
//   * E(i+1, q) is taken from the H(i, q) of the main pass, BEFORE the lazy-F loop raises it ("we do not need to set E(i,j) as we
//     disallow adjacent insertion and then deletion", ksw.c:168).  The main pass carries F only inside a segment of slen =
//     ceil(qlen / p) consecutive query positions (p = 16 scores per vector for bytes, 8 for words), so which insertion-then-
//     deletion paths exist depends on the segmentation;
//   * the query is padded to slen * p positions that score 0 against everything; they take part in the row maximum, hence in the
//     list of candidate second-best rows (ksw.c:180-188).
// One warp per job: lane L < p owns segment L, i.e. query positions L*slen .. L*slen+slen-1 (the j-th score of SSE lane L), the
// row state H0/H1/E/Hmax lives in shared memory as [j][lane], the lane shift of the SSE code (_mm_slli_si128) is a shuffle, and the
// early exit of the lazy-F loop is a vote.  Scores are plain ints with the saturation of the byte / word kernels applied explicitly.
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "batches.cuh"

namespace bwaflow {

constexpr int KSW_XBYTE = 0x10000, KSW_XSTOP = 0x20000, KSW_XSUBO = 0x40000, KSW_XSTART = 0x80000;

struct SwArgs {
  const bwaflow_sw_job_t *jobs;
  const uint8_t *qpool, *tpool;
  bwaflow_sw_result_t *out;
  int *head; int n_jobs;
  int slen_cap;                 // segment length the shared-memory row state is sized for
  unsigned long long *bscr; int64_t b_cap;   // per warp: candidate rows of the second-best score (the `b` array of ksw.c:130)
  int8_t mat[25];
  int o_del, e_del, o_ins, e_ins, shift, maxsc;
  unsigned long long *cells;
  int *err;
};

struct SwOut { int score, te, qe, score2, te2; };

// ksw_u8 / ksw_i16 on the query `query[0..qlen)` read through qf(k) and the target through tf(i).  p = scores per vector.
template <class QF, class TF>
__device__ void sw_pass(const SwArgs &a, int p, bool bytes, int qlen, int tlen, int xtra, QF qf, TF tf, uint16_t *H0, uint16_t *H1, uint16_t *E,
                        uint16_t *HM, uint8_t *QC, unsigned long long *b, SwOut &r, unsigned long long &cells)
{
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool on = lane < p;
  const int slen = (qlen + p - 1) / p;
  const int minsc = (xtra & KSW_XSUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & KSW_XSTOP) ? xtra & 0xffff : 0x10000;
  const int cap = bytes ? 255 : 32767, sh = bytes ? a.shift : 0;
  const int oe_del = a.o_del + a.e_del, oe_ins = a.o_ins + a.e_ins;
  if (on)
    for (int j = 0; j < slen; ++j) {
      const int k = lane * slen + j;
      H0[j * 16 + lane] = 0; E[j * 16 + lane] = 0; HM[j * 16 + lane] = 0;
      QC[j * 16 + lane] = k < qlen ? (uint8_t)qf(k) : (uint8_t)5;            // 5 = padding, scores 0
    }
  __syncwarp();
  int gmax = 0, te = -1, n_b = 0, i;
  for (i = 0; i < tlen; ++i) {
    const int t = tf(i);
    // scores of this target base against the five query codes and the padding
    int srow[6];
#pragma unroll
    for (int c = 0; c < 5; ++c) srow[c] = a.mat[t * 5 + c];
    srow[5] = 0;
    int f = 0, mx = 0;
    int h = on && slen > 0 ? H0[(slen - 1) * 16 + lane] : 0;
    h = __shfl_up_sync(full, h, 1);                                          // _mm_slli_si128: lane L takes lane L-1's last score
    if (lane == 0) h = 0;
    if (on)
      for (int j = 0; j < slen; ++j) {
        const int q = QC[j * 16 + lane];
        int s = q == 0 ? srow[0] : q == 1 ? srow[1] : q == 2 ? srow[2] : q == 3 ? srow[3] : q == 4 ? srow[4] : 0;
        if (bytes) { h = h + s + sh; h = h > 255 ? 255 : h; h -= sh; h = h > 0 ? h : 0; }   // adds_epu8 then subs_epu8 (ksw.c:152-153)
        else { h = h + s; h = h > 32767 ? 32767 : h; }                                      // adds_epi16
        int e = E[j * 16 + lane];
        h = h > e ? h : e;
        h = h > f ? h : f;
        mx = mx > h ? mx : h;
        H1[j * 16 + lane] = (uint16_t)h;
        e -= a.e_del; e = e > 0 ? e : 0;
        int tt = h - oe_del; tt = tt > 0 ? tt : 0;
        e = e > tt ? e : tt;
        E[j * 16 + lane] = (uint16_t)e;
        f -= a.e_ins; f = f > 0 ? f : 0;
        tt = h - oe_ins; tt = tt > 0 ? tt : 0;
        f = f > tt ? f : tt;
        h = H0[j * 16 + lane];
      }
    cells += (unsigned long long)slen * p;
    // the lazy-F loop (ksw.c:169-182 / 279-290): F crosses the segment boundaries one lane per round
    {
      bool out = false;
      for (int k = 0; k < 16 && !out; ++k) {
        f = __shfl_up_sync(full, f, 1);
        if (lane == 0) f = 0;
        for (int j = 0; j < slen; ++j) {
          bool fin = true;
          if (on) {
            int hh = H1[j * 16 + lane];
            hh = hh > f ? hh : f;
            H1[j * 16 + lane] = (uint16_t)hh;
            hh -= oe_ins; hh = hh > 0 ? hh : 0;
            f -= a.e_ins; f = f > 0 ? f : 0;
            fin = f <= hh;
          }
          if (__all_sync(full, fin)) { out = true; break; }
        }
      }
    }
    for (int d = 8; d; d >>= 1) { const int o = __shfl_xor_sync(full, mx, d); mx = mx > o ? mx : o; }
    const int imax = __shfl_sync(full, mx, 0);                               // lanes 0..15 hold the maximum of lanes 0..p-1 (idle lanes contribute 0)
    if (imax >= minsc && lane == 0) {                                        // ksw.c:184-192
      if (n_b == 0 || (int)(uint32_t)b[n_b - 1] + 1 != i) { if (n_b < a.b_cap) b[n_b++] = (unsigned long long)imax << 32 | (uint32_t)i; else atomicExch(a.err, 7); }
      else if ((int)(b[n_b - 1] >> 32) < imax) b[n_b - 1] = (unsigned long long)imax << 32 | (uint32_t)i;
    }
    bool stop = false;
    if (imax > gmax) {
      gmax = imax; te = i;
      if (on) for (int j = 0; j < slen; ++j) HM[j * 16 + lane] = H1[j * 16 + lane];
      if ((bytes && gmax + a.shift >= 255) || gmax >= endsc) stop = true;
    }
    __syncwarp();
    if (stop) break;
    uint16_t *sw = H1; H1 = H0; H0 = sw;
  }
  r.score = bytes ? (gmax + a.shift < 255 ? gmax : 255) : gmax;
  r.te = te; r.qe = -1; r.score2 = -1; r.te2 = -1;
  if (!(bytes && r.score == 255)) {
    // end of the best hit on the query: the largest score of the saved row, smallest position among equals (ksw.c:201-204)
    int best = -1, bq = 0x7fffffff;
    if (on)
      for (int j = 0; j < slen; ++j) {
        const int v = HM[j * 16 + lane], q = lane * slen + j;
        if (v > best || (v == best && q < bq)) { best = v; bq = q; }
      }
    for (int d = 8; d; d >>= 1) {
      const int ob = __shfl_xor_sync(full, best, d), oq = __shfl_xor_sync(full, bq, d);
      if (ob > best || (ob == best && oq < bq)) { best = ob; bq = oq; }
    }
    r.qe = __shfl_sync(full, slen > 0 ? bq : -1, 0);
    if (slen == 0) r.qe = -1;
    // second-best score outside the band of rows the best hit can explain (ksw.c:206-214)
    int s2 = -1, t2 = -1;
    if (lane == 0 && n_b > 0) {
      const int w = (r.score + a.maxsc - 1) / a.maxsc, low = te - w, high = te + w;
      for (int k = 0; k < n_b; ++k) {
        const int e = (int)(uint32_t)b[k];
        if ((e < low || e > high) && (int)(b[k] >> 32) > s2) { s2 = (int)(b[k] >> 32); t2 = e; }
      }
    }
    r.score2 = __shfl_sync(full, s2, 0); r.te2 = __shfl_sync(full, t2, 0);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128) sw_kernel(const SwArgs a)
{
  extern __shared__ uint16_t s_sw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = (size_t)a.slen_cap * 16 * 4 + (size_t)a.slen_cap * 8;   // four u16 arrays + the query codes (bytes)
  uint16_t *H0 = s_sw + warp * per_warp, *H1 = H0 + a.slen_cap * 16, *E = H1 + a.slen_cap * 16, *HM = E + a.slen_cap * 16;
  uint8_t *QC = (uint8_t *)(HM + a.slen_cap * 16);
  unsigned long long *b = a.bscr + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * a.b_cap;
  unsigned long long cells = 0;
  for (;;) {
    int job = 0;
    if (lane == 0) job = atomicAdd(a.head, 1);
    job = __shfl_sync(0xffffffffu, job, 0);
    if (job >= a.n_jobs) break;
    const bwaflow_sw_job_t jb = a.jobs[job];
    const uint8_t *qp = a.qpool + jb.q_off, *tp = a.tpool + jb.t_off;
    const bool bytes = (jb.xtra & KSW_XBYTE) != 0;
    const int p = bytes ? 16 : 8;
    bwaflow_sw_result_t res = {0, -1, -1, -1, -1, -1, -1};                  // g_defr, ksw.c:44
    const int slen = (jb.qlen + p - 1) / p;
    if (jb.qlen <= 0 || jb.tlen < 0 || slen > a.slen_cap || jb.tlen > a.b_cap || (bytes && jb.qlen * a.maxsc >= 250)) {
      if (lane == 0) { atomicExch(a.err, 1); a.out[job] = res; }
      continue;
    }
    SwOut r;
    sw_pass(a, p, bytes, jb.qlen, jb.tlen, jb.xtra, [&](int k) { return (int)qp[k]; }, [&](int i) { return (int)tp[i]; }, H0, H1, E, HM, QC, b, r, cells);
    res.score = r.score; res.te = r.te; res.qe = r.qe; res.score2 = r.score2; res.te2 = r.te2;
    if ((jb.xtra & KSW_XSTART) && !((jb.xtra & KSW_XSUBO) && r.score < (jb.xtra & 0xffff)) && r.qe >= 0) {
      // second pass over the reversed prefixes (ksw.c:357-363): stops as soon as the same score is reached
      const int qe = r.qe, te = r.te;
      SwOut rr;
      sw_pass(a, p, bytes, qe + 1, jb.tlen, KSW_XSTOP | r.score, [&](int k) { return (int)qp[qe - k]; },
              [&](int i) { return (int)(i <= te ? tp[te - i] : tp[i]); }, H0, H1, E, HM, QC, b, rr, cells);
      if (r.score == rr.score) { res.tb = r.te - rr.te; res.qb = r.qe - rr.qe; }
    }
    if (lane == 0) a.out[job] = res;
  }
  if (lane == 0 && cells) atomicAdd(a.cells, cells);
}

} // namespace bwaflow

using namespace bwaflow;

struct bwaflow_sw_batch {
  bwaflow_ctx_t *ctx;
  cudaStream_t stream;
  Timer tm;
  int64_t max_jobs, max_q, max_t, n_jobs;
  bwaflow_sw_job_t *d_jobs; uint8_t *d_q, *d_t;
  bwaflow_sw_result_t *d_out, *h_out;
  int *d_ctl, *h_ctl;               // head | err | cells (u64)
  unsigned long long *d_bscr; int64_t b_cap; int n_warps;
  int blocks;
  bool pending;
};

namespace bwaflow {

static void sw_fill_args(bwaflow_ctx_t *ctx, SwArgs &a)
{
  const bwaflow_opt_t &o = ctx->par.opt;
  int mn = 127, mxs = 0;
  for (int i = 0; i < 25; ++i) { a.mat[i] = o.mat[i]; if (o.mat[i] < mn) mn = o.mat[i]; if (o.mat[i] > mxs) mxs = o.mat[i]; }
  a.o_del = o.o_del; a.e_del = o.e_del; a.o_ins = o.o_ins; a.e_ins = o.e_ins;
  a.shift = (uint8_t)(256 - (uint8_t)mn);                                    // ksw_qinit, ksw.c:79-85
  a.maxsc = (uint8_t)mxs;
}

// Runs n jobs already resident in d_jobs / pools on `st`.  max_qlen / max_tlen bound the jobs (they size the row state and the
// per-warp candidate lists); jobs beyond them are reported through the error word.
SwDeviceView sw_device_view(bwaflow_sw_batch *b)
{
  SwDeviceView v;
  v.jobs = b->d_jobs; v.q = b->d_q; v.t = b->d_t; v.out = b->d_out; v.d_ctl = b->d_ctl; v.h_ctl = b->h_ctl;
  return v;
}

int sw_launch_device(bwaflow_sw_batch *b, int64_t n_jobs, int max_qlen, int max_tlen, cudaStream_t st)
{
  BF_REQUIRE(n_jobs >= 0 && n_jobs <= b->max_jobs, -7, "sw: %lld jobs, capacity %lld", (long long)n_jobs, (long long)b->max_jobs);
  b->n_jobs = n_jobs;
  BF_CUDA(cudaMemsetAsync(b->d_ctl, 0, 4 * sizeof(int), st));
  if (n_jobs == 0) return 0;
  SwArgs a;
  sw_fill_args(b->ctx, a);
  a.jobs = b->d_jobs; a.qpool = b->d_q; a.tpool = b->d_t; a.out = b->d_out; a.head = b->d_ctl; a.err = b->d_ctl + 1;
  a.cells = (unsigned long long *)(b->d_ctl + 2); a.n_jobs = (int)n_jobs;
  a.slen_cap = (max_qlen + 7) / 8; if (a.slen_cap < 1) a.slen_cap = 1;
  const int warps = 4;
  const size_t smem = (size_t)warps * ((size_t)a.slen_cap * 16 * 4 * sizeof(uint16_t) + (size_t)a.slen_cap * 16);
  BF_REQUIRE(smem <= 200 * 1024, -34, "sw: queries of %d bases do not fit the row state", max_qlen);
  int64_t need = (int64_t)max_tlen + 1;
  if (need > b->b_cap) {
    BF_CUDA(cudaStreamSynchronize(st));
    cudaFree(b->d_bscr);
    b->b_cap = need * 2;
    BF_CUDA(cudaMalloc(&b->d_bscr, (size_t)b->n_warps * b->b_cap * sizeof(unsigned long long)));
  }
  a.bscr = b->d_bscr; a.b_cap = b->b_cap;
  BF_CUDA(cudaFuncSetAttribute(sw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int blocks = (int)((n_jobs + warps - 1) / warps);
  if (blocks > b->blocks) blocks = b->blocks;
  sw_kernel<<<blocks, warps * 32, smem, st>>>(a);
  BF_CUDA(cudaGetLastError());
  return 0;
}

} // namespace bwaflow

extern "C" {

int bwaflow_sw_batch_create(bwaflow_ctx_t *ctx, int64_t max_jobs, int64_t max_qbytes, int64_t max_tbytes, bwaflow_sw_batch_t **out)
{
  BF_REQUIRE(ctx && out, -22, "bwaflow_sw_batch_create: null argument");
  BF_REQUIRE(max_jobs > 0 && max_jobs < (1LL << 31), -22, "max_jobs out of range");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_sw_batch *b = (bwaflow_sw_batch *)calloc(1, sizeof(*b));
  b->ctx = ctx; b->max_jobs = max_jobs; b->max_q = max_qbytes; b->max_t = max_tbytes;
  BF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (b->tm.init()) return -5;
  BF_CUDA(cudaMalloc(&b->d_jobs, max_jobs * sizeof(bwaflow_sw_job_t)));
  BF_CUDA(cudaMalloc(&b->d_q, max_qbytes + 16));
  BF_CUDA(cudaMalloc(&b->d_t, max_tbytes + 16));
  BF_CUDA(cudaMalloc(&b->d_out, max_jobs * sizeof(bwaflow_sw_result_t)));
  BF_CUDA(cudaMallocHost(&b->h_out, max_jobs * sizeof(bwaflow_sw_result_t)));
  BF_CUDA(cudaMalloc(&b->d_ctl, 4 * sizeof(int)));
  BF_CUDA(cudaMallocHost(&b->h_ctl, 4 * sizeof(int)));
  b->blocks = ctx->n_sm * 4;
  b->n_warps = b->blocks * 4;
  b->b_cap = 4096;
  BF_CUDA(cudaMalloc(&b->d_bscr, (size_t)b->n_warps * b->b_cap * sizeof(unsigned long long)));
  *out = b;
  return 0;
}

void bwaflow_sw_batch_destroy(bwaflow_sw_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  cudaFree(b->d_jobs); cudaFree(b->d_q); cudaFree(b->d_t); cudaFree(b->d_out); cudaFree(b->d_ctl); cudaFree(b->d_bscr);
  cudaFreeHost(b->h_out); cudaFreeHost(b->h_ctl);
  b->tm.destroy();
  cudaStreamDestroy(b->stream);
  free(b);
}

int bwaflow_sw_submit(bwaflow_sw_batch_t *b, int64_t n_jobs, const bwaflow_sw_job_t *jobs, const uint8_t *qpool, int64_t qbytes,
                      const uint8_t *tpool, int64_t tbytes)
{
  BF_REQUIRE(b && (n_jobs == 0 || (jobs && qpool && tpool)), -22, "bwaflow_sw_submit: null argument");
  BF_REQUIRE(n_jobs >= 0 && n_jobs <= b->max_jobs && qbytes <= b->max_q && tbytes <= b->max_t, -7, "bwaflow_sw_submit: batch capacity exceeded");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  int max_q = 1, max_t = 1;
  for (int64_t i = 0; i < n_jobs; ++i) {
    BF_REQUIRE(jobs[i].qlen > 0 && jobs[i].tlen >= 0 && jobs[i].q_off >= 0 && jobs[i].t_off >= 0 && jobs[i].q_off + jobs[i].qlen <= qbytes &&
                   jobs[i].t_off + jobs[i].tlen <= tbytes, -22, "bwaflow_sw_submit: job %lld out of range", (long long)i);
    if (jobs[i].qlen > max_q) max_q = jobs[i].qlen;
    if (jobs[i].tlen > max_t) max_t = jobs[i].tlen;
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  if (n_jobs) {
    BF_CUDA(cudaMemcpyAsync(b->d_jobs, jobs, n_jobs * sizeof(bwaflow_sw_job_t), cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_q, qpool, qbytes, cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_t, tpool, tbytes, cudaMemcpyHostToDevice, b->stream));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  int r = sw_launch_device(b, n_jobs, max_q, max_t, b->stream);
  if (r) return r;
  BF_CUDA(cudaEventRecord(b->tm.ev[2], b->stream));
  if (n_jobs) BF_CUDA(cudaMemcpyAsync(b->h_out, b->d_out, n_jobs * sizeof(bwaflow_sw_result_t), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl, 4 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[3], b->stream));
  b->pending = true;
  return 0;
}

int bwaflow_sw_collect(bwaflow_sw_batch_t *b, const bwaflow_sw_result_t **out, bwaflow_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_sw_collect: nothing submitted");
  BF_CUDA(cudaStreamSynchronize(b->stream));
  b->pending = false;
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->tm.ev[0], b->tm.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, b->tm.ev[1], b->tm.ev[2]);
    cudaEventElapsedTime(&t->d2h_ms, b->tm.ev[2], b->tm.ev[3]);
    cudaEventElapsedTime(&t->total_ms, b->tm.ev[0], b->tm.ev[3]);
    t->n_launches = b->n_jobs ? 1 : 0;
    t->work_units = (double)*(unsigned long long *)(b->h_ctl + 2);          // striped cells computed (main passes, padding included)
  }
  BF_REQUIRE(b->h_ctl[1] != 1, -34, "ksw_align2 job outside the supported range (query too long for the row state, or a byte-mode job whose score can reach 250)");
  BF_REQUIRE(b->h_ctl[1] != 7, -34, "ksw_align2 job with more candidate rows than the second-best list holds");
  if (out) *out = b->h_out;
  return 0;
}

} // extern "C"
