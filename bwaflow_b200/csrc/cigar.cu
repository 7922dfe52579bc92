// CIGAR generation for reported alignments: mem_reg2aln (bwa/bwamem.c:1104-1174) -> bwa_gen_cigar2
// (bwa/bwa.c:121-207) -> ksw_global2 (bwa/ksw.c:504-606) for a batch of (read, region) jobs.
//
// One lane owns one job and runs what the reference runs for it: the band-width inference and the up-to-three band tries of
// mem_reg2aln, the banded global alignment with its backtrack matrix, the traceback, NM and the MD string, the squeeze of a leading /
// trailing deletion and the clipping.  Jobs are queued by band width and target length so that the lanes of a warp walk loops of the
// same shape.  The row state eh[] is a ring over the band in shared memory ([column & (R-1)][lane]: a lane's columns never leave
// the band, 2w+2 <= R), the backtrack matrix z (one byte per cell, f<<4|e<<2|h as in ksw.c:514) lives in HBM.
// Scores are the reference's int32 with its MINUS_INF, so every comparison -- hence every direction bit -- is the reference's.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include "internal.cuh"

namespace bwaflow {

constexpr int CG_MINUS_INF = -0x40000000;

struct CgJob {
  int64_t rb, re, q_src;          // reference interval; first base of the read in the batch's base array
  int32_t qb, qe, l_query, truesc, w_reg, w_cap, mode, _p1;   // mode 1: one score-only bwa_gen_cigar2 call with w_ = w_reg (mem_patch_reg)
  int64_t z_off, cig_off, md_off; // this job's slices of the backtrack / CIGAR / MD scratch
  int32_t cig_cap, md_cap;
};
struct CgOut { int64_t pos; int32_t score, n_cigar, NM, md_len, is_rev, err; };

struct CgArgs {
  const CgJob *jobs; const uint32_t *order; int n_jobs; int *head;
  const uint8_t *bases; uint8_t *z; uint32_t *cig; char *md; CgOut *out;
  int2 *geh; int64_t geh_stride;  // global row state for bands wider than the shared ring (per resident lane)
  DevIndex ix;
  int8_t mat[25];
  int a, o_del, e_del, o_ins, e_ins, w;
};

__device__ __forceinline__ int cg_ref_base(const DevIndex &ix, int64_t p)
{
  const bool rev = p >= ix.l_pac;
  const int64_t f = rev ? (ix.l_pac << 1) - 1 - p : p;
  const int b = (__ldg(ix.pac + (f >> 2)) >> ((~f & 3) << 1)) & 3;
  return rev ? 3 - b : b;
}

// infer_bw, bwa/bwamem.c:801-808
__host__ __device__ inline int cg_infer_bw(int l1, int l2, int score, int a, int q, int r)
{
  if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
  int w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
  const int d = l1 > l2 ? l1 - l2 : l2 - l1;
  if (w < d) w = d;
  return w;
}
// the band bwa_gen_cigar2 hands to ksw_global2 for the caller's w_ (bwa/bwa.c:150-158)
__host__ __device__ inline int cg_band(int l_query, int rlen, int a, int o_del, int e_del, int o_ins, int e_ins, int w_)
{
  int max_ins = (int)((double)(((l_query + 1) >> 1) * a - o_ins) / e_ins + 1.);
  int max_del = (int)((double)(((l_query + 1) >> 1) * a - o_del) / e_del + 1.);
  int max_gap = max_ins > max_del ? max_ins : max_del;
  max_gap = max_gap > 1 ? max_gap : 1;
  const int d = rlen > l_query ? rlen - l_query : l_query - rlen;
  int w = (max_gap + d + 1) >> 1;
  w = w < w_ ? w : w_;
  const int min_w = d + 3;
  return w > min_w ? w : min_w;
}

__device__ __forceinline__ void cg_putc(char *s, int &l, int cap, char c) { if (l < cap) s[l] = c; ++l; }
__device__ __forceinline__ void cg_putw(char *s, int &l, int cap, int c)
{ // kputw, bwa/kstring.h:62-77 (non-negative values only here)
  char buf[12]; int n = 0;
  if (c == 0) { cg_putc(s, l, cap, '0'); return; }
  for (int x = c; x > 0; x /= 10) buf[n++] = (char)('0' + x % 10);
  for (int x = n - 1; x >= 0; --x) cg_putc(s, l, cap, buf[x]);
}

template <int R>
__global__ void __launch_bounds__(32) cigar_kernel(const CgArgs a)
{
  extern __shared__ int2 s_eh[];
  const int lane = threadIdx.x;
  int2 *ring = R ? s_eh + lane : a.geh + (int64_t)blockIdx.x * 32 * a.geh_stride + lane;   // [column][lane], in shared memory or HBM
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(a.head, 32);
    slot = __shfl_sync(0xffffffffu, slot, 0) + lane;
    if (slot - lane >= a.n_jobs) break;
    if (slot < a.n_jobs) {
      const int job = (int)a.order[slot];
      const CgJob jb = a.jobs[job];
      const int lq = jb.qe - jb.qb, rlen = (int)(jb.re - jb.rb);
      const bool rev = jb.rb >= a.ix.l_pac;                 // both sequences reversed: indels end up leftmost on the forward strand
      const uint8_t *qp = a.bases + jb.q_src + jb.qb;
      uint8_t *z = a.z + jb.z_off;
      uint32_t *cig = a.cig + jb.cig_off;
      char *md = a.md + jb.md_off;
      CgOut o;
      o.err = 0; o.NM = -1; o.n_cigar = 0; o.md_len = 0; o.score = 0;
      auto Q = [&](int k) -> int { return rev ? qp[lq - 1 - k] : qp[k]; };
      auto T = [&](int i) -> int { return cg_ref_base(a.ix, rev ? jb.rb + (rlen - 1 - i) : jb.rb + i); };
      if (lq <= 0 || jb.rb >= jb.re || (jb.rb < a.ix.l_pac && jb.re > a.ix.l_pac)) o.err = 1;
      // ---- mem_reg2aln's band: inferred from the local score, capped by the region's and 4 * opt->w (:1123-1131)
      const bool sco = jb.mode == 1;                        // mem_patch_reg's call (bwa/bwamem.c:433): the caller's band, score only
      int w2 = cg_infer_bw(lq, rlen, jb.truesc, a.a, a.o_del, a.e_del);
      { const int t = cg_infer_bw(lq, rlen, jb.truesc, a.a, a.o_ins, a.e_ins); w2 = w2 > t ? w2 : t; }
      if (w2 > a.w) w2 = w2 < jb.w_reg ? w2 : jb.w_reg;
      if (sco) w2 = jb.w_reg;
      int score = 0, last_sc = -(1 << 30), n_cigar = 0, it = 0;
      if (!o.err)
        do {
          if (!sco) w2 = w2 < a.w << 2 ? w2 : a.w << 2;
          n_cigar = 0;
          if (lq == rlen && w2 == 0) {                      // bwa/bwa.c:141-149: no gap possible, no DP
            if (!sco) { cig[0] = (uint32_t)lq << 4; n_cigar = 1; }
            score = 0;
            for (int k = 0; k < lq; ++k) score += a.mat[T(k) * 5 + Q(k)];
          } else {
            const int w = cg_band(lq, rlen, a.mat[0], a.o_del, a.e_del, a.o_ins, a.e_ins, w2);
            if (w > jb.w_cap || (R && 2 * w + 2 > R)) { o.err = 2; break; }
            const int n_col = lq < 2 * w + 1 ? lq : 2 * w + 1;
            const int oe_del = a.o_del + a.e_del, oe_ins = a.o_ins + a.e_ins;
            const int M_ = R ? R - 1 : 0x7fffffff;
            // ---- ksw_global2, bwa/ksw.c:504-606.  Only columns of the band are ever read: row i reads eh[beg..end) and writes
            // eh[beg..end]; a column entering on the right (j = i + w) was last written as the previous row's eh[end] or, in row 0, is
            // row -1's value.  Row -1 therefore only needs columns 0..w (+1), which is what the ring holds.
            ring[0] = make_int2(0, CG_MINUS_INF);
            for (int j = 1; j <= lq && j <= w; ++j) ring[(size_t)(j & M_) * 32] = make_int2(-(a.o_ins + a.e_ins * j), CG_MINUS_INF);
            if (w + 1 <= lq) ring[(size_t)((w + 1) & M_) * 32] = make_int2(CG_MINUS_INF, CG_MINUS_INF);
            for (int i = 0; i < rlen; ++i) {
              int f = CG_MINUS_INF, h1;
              const int tb = T(i);
              const int8_t *mrow = a.mat + tb * 5;
              const int beg = i > w ? i - w : 0;
              const int end = i + w + 1 < lq ? i + w + 1 : lq;
              h1 = beg == 0 ? -(a.o_del + a.e_del * (i + 1)) : CG_MINUS_INF;
              uint8_t *zi = z + (size_t)i * n_col;
              for (int j = beg; j < end; ++j) {
                int2 *p = ring + (size_t)(j & M_) * 32;
                const int2 v = *p;
                int m = v.x, e = v.y, h, t;
                uint8_t d;
                m += mrow[Q(j)];
                d = m >= e ? 0 : 1;
                h = m >= e ? m : e;
                d = h >= f ? d : 2;
                h = h >= f ? h : f;
                t = m - oe_del;
                e -= a.e_del;
                d |= e > t ? 1 << 2 : 0;
                e = e > t ? e : t;
                *p = make_int2(h1, e);
                h1 = h;
                t = m - oe_ins;
                f -= a.e_ins;
                d |= f > t ? 2 << 4 : 0;
                f = f > t ? f : t;
                if (!sco) zi[j - beg] = d;
              }
              ring[(size_t)(end & M_) * 32] = make_int2(h1, CG_MINUS_INF);
            }
            score = ring[(size_t)(lq & M_) * 32].x;
            if (sco) break;
            // ---- backtrack
            int i = rlen - 1, k = (i + w + 1 < lq ? i + w + 1 : lq) - 1, which = 0;
            auto push = [&](int op, int len) {
              if (n_cigar == 0 || op != (int)(cig[n_cigar - 1] & 0xf)) { if (n_cigar < jb.cig_cap) cig[n_cigar] = (uint32_t)len << 4 | (uint32_t)op; else o.err = 3; ++n_cigar; }
              else if (n_cigar <= jb.cig_cap) cig[n_cigar - 1] += (uint32_t)len << 4;
            };
            while (i >= 0 && k >= 0) {
              which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
              if (which == 0) { push(0, 1); --i; --k; }
              else if (which == 1) { push(2, 1); --i; }
              else { push(1, 1); --k; }
            }
            if (i >= 0) push(2, i + 1);
            if (k >= 0) push(1, k + 1);
            if (o.err) break;
            for (int x = 0; x < n_cigar >> 1; ++x) { const uint32_t t = cig[x]; cig[x] = cig[n_cigar - 1 - x]; cig[n_cigar - 1 - x] = t; }
          }
          if (sco || score == last_sc || w2 == a.w << 2) break;      // :1134
          last_sc = score;
          w2 <<= 1;
        } while (++it < 3 && score < jb.truesc - a.a);
      // ---- NM and MD, bwa/bwa.c:164-197
      int NM = -1, l = 0;
      if (sco) { o.pos = 0; o.is_rev = 0; o.score = score; a.out[job] = o; continue; }
      if (!o.err) {
        int x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0;
        const char *int2base = jb.rb < a.ix.l_pac ? "ACGTN" : "TGCAN";
        for (int c = 0; c < n_cigar; ++c) {
          const int op = cig[c] & 0xf, len = (int)(cig[c] >> 4);
          if (op == 0) {
            for (int k = 0; k < len; ++k) {
              const int tbase = T(y + k);
              if (Q(x + k) != tbase) { cg_putw(md, l, jb.md_cap, u); cg_putc(md, l, jb.md_cap, int2base[tbase]); ++n_mm; u = 0; }
              else ++u;
            }
            x += len; y += len;
          } else if (op == 2) {
            if (c > 0 && c < n_cigar - 1) {
              cg_putw(md, l, jb.md_cap, u); cg_putc(md, l, jb.md_cap, '^');
              for (int k = 0; k < len; ++k) cg_putc(md, l, jb.md_cap, int2base[T(y + k)]);
              u = 0; n_gap += len;
            }
            y += len;
          } else if (op == 1) { x += len; n_gap += len; }
        }
        cg_putw(md, l, jb.md_cap, u);
        NM = n_mm + n_gap;
        if (l >= jb.md_cap) o.err = 4;
      }
      // ---- position, squeeze of a leading / trailing deletion, clipping (bwamem.c:1140-1166)
      int64_t pos = jb.rb < a.ix.l_pac ? jb.rb : jb.re - 1;
      const int is_rev = pos >= a.ix.l_pac;
      if (is_rev) pos = (a.ix.l_pac << 1) - 1 - pos;
      if (!o.err && n_cigar > 0) {
        if ((cig[0] & 0xf) == 2) { pos += cig[0] >> 4; --n_cigar; for (int c = 0; c < n_cigar; ++c) cig[c] = cig[c + 1]; }
        else if ((cig[n_cigar - 1] & 0xf) == 2) --n_cigar;
      }
      if (!o.err && (jb.qb != 0 || jb.qe != jb.l_query)) {
        const int clip5 = is_rev ? jb.l_query - jb.qe : jb.qb, clip3 = is_rev ? jb.qb : jb.l_query - jb.qe;
        if (n_cigar + 2 > jb.cig_cap) o.err = 3;
        else {
          if (clip5) { for (int c = n_cigar; c > 0; --c) cig[c] = cig[c - 1]; cig[0] = (uint32_t)clip5 << 4 | 3; ++n_cigar; }
          if (clip3) cig[n_cigar++] = (uint32_t)clip3 << 4 | 3;
        }
      }
      o.pos = pos; o.is_rev = is_rev; o.score = score; o.n_cigar = o.err ? 0 : n_cigar; o.NM = NM; o.md_len = o.err ? 0 : l;
      a.out[job] = o;
    }
  }
}

__global__ void cigar_sizes_kernel(const CgOut *out, int64_t n, int64_t *sz)
{
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  sz[j] = j < n ? out[j].n_cigar : 0;
  sz[n + 1 + j] = j < n ? out[j].md_len + 1 : 0;          // with the terminating NUL
}

__global__ void cigar_gather_kernel(const CgJob *jobs, const CgOut *out, int64_t n, const int64_t *off, const uint32_t *cig, const char *md,
                                    uint32_t *cig_out, char *md_out)
{
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const CgJob jb = jobs[j];
  const CgOut o = out[j];
  for (int k = lane; k < o.n_cigar; k += 32) cig_out[off[j] + k] = cig[jb.cig_off + k];
  const int64_t mo = off[n + 1 + j];
  for (int k = lane; k < o.md_len; k += 32) md_out[mo + k] = md[jb.md_off + k];
  if (lane == 0) md_out[mo + o.md_len] = 0;
}

} // namespace bwaflow

using namespace bwaflow;

struct bwaflow_cigar_batch {
  bwaflow_ctx_t *ctx;
  bwaflow_pe_opt_t po;
  cudaStream_t stream;
  Timer tm;
  int64_t max_reads, max_bases, max_jobs, n_jobs;
  uint8_t *d_bases;
  CgJob *d_jobs; CgOut *d_out; uint32_t *d_order[2], *d_keys[2];
  uint8_t *d_z; size_t z_cap; uint32_t *d_cig; size_t cig_cap; char *d_md; size_t md_cap;
  uint32_t *d_cig_out; char *d_md_out; size_t out_cig_cap, out_md_cap;
  int64_t *d_sz, *d_off;
  int2 *d_geh; size_t geh_cap;
  void *d_tmp; size_t tmp_bytes;
  int *d_ctl;
  std::vector<CgJob> jobs; std::vector<uint32_t> keys, vals;
  std::vector<CgOut> outs; std::vector<int64_t> offs;
  std::vector<bwaflow_alnreg_t> regs;
  std::vector<bwaflow_aln_t> aln; std::vector<uint32_t> cig; std::vector<char> md;
  bool pending;
};

namespace bwaflow {
int64_t cigar_max_jobs(const bwaflow_cigar_batch_t *b) { return b->max_jobs; }
bwaflow_ctx_t *cigar_ctx(const bwaflow_cigar_batch_t *b) { return b->ctx; }
const bwaflow_pe_opt_t *cigar_po(const bwaflow_cigar_batch_t *b) { return &b->po; }
}

// mem_approx_mapq_se, bwa/bwamem.c:967-991 (mapQ_coef_len > 0 branch and the MEM_MAPQ_COEF one)
static int approx_mapq_se(const bwaflow_opt_t &o, const bwaflow_pe_opt_t &po, const bwaflow_alnreg_t &a)
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

template <class T>
static int cg_grow(T **p, size_t *cap, size_t need)
{
  if (need <= *cap) return 0;
  cudaFree(*p);
  *cap = need + need / 4 + 4096;
  BF_CUDA(cudaMalloc(p, *cap * sizeof(T)));
  return 0;
}

// one launch of cigar_kernel over the batch's job array in queue order `order`: ring size by the widest band, row state in HBM
// when a band does not fit the shared ring
static int cg_launch(bwaflow_cigar_batch *b, int64_t n_jobs, const uint32_t *order, int w_max, int lq_max, int wide)
{
  const bwaflow_opt_t &o = b->ctx->par.opt;
  int r;
  CgArgs a;
  a.jobs = b->d_jobs; a.order = order; a.n_jobs = (int)n_jobs; a.head = b->d_ctl; a.bases = b->d_bases; a.z = b->d_z; a.cig = b->d_cig;
  a.md = b->d_md; a.out = b->d_out; a.ix = b->ctx->ix; a.geh = nullptr; a.geh_stride = 0;
  for (int i = 0; i < 25; ++i) a.mat[i] = o.mat[i];
  a.a = o.a; a.o_del = o.o_del; a.e_del = o.e_del; a.o_ins = o.o_ins; a.e_ins = o.e_ins; a.w = o.w;
  const int R = 2 * w_max + 2 <= 64 ? 64 : 256;
  const size_t smem = (size_t)R * 32 * sizeof(int2);
  int blocks = (int)((n_jobs + 31) / 32);
  if (wide) {                                             // bands beyond the shared ring: row state of lq + 1 columns per lane in HBM
    const int cap_blocks = b->ctx->n_sm * 8;
    if (blocks > cap_blocks) blocks = cap_blocks;
    a.geh_stride = lq_max + 2;
    if ((r = cg_grow(&b->d_geh, &b->geh_cap, (size_t)blocks * 32 * a.geh_stride))) return r;
    a.geh = b->d_geh;
    cigar_kernel<0><<<blocks, 32, 0, b->stream>>>(a);
  } else if (R == 64) {
    BF_CUDA(cudaFuncSetAttribute(cigar_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int cap_blocks = b->ctx->n_sm * 12;
    cigar_kernel<64><<<blocks < cap_blocks ? blocks : cap_blocks, 32, smem, b->stream>>>(a);
  } else {
    BF_CUDA(cudaFuncSetAttribute(cigar_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int cap_blocks = b->ctx->n_sm * 3;
    cigar_kernel<256><<<blocks < cap_blocks ? blocks : cap_blocks, 32, smem, b->stream>>>(a);
  }
  return 0;
}

extern "C" {

int bwaflow_cigar_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, int64_t max_jobs, const bwaflow_pe_opt_t *po,
                               bwaflow_cigar_batch_t **out)
{
  BF_REQUIRE(ctx && out && max_jobs > 0 && max_jobs < (1LL << 31), -22, "bwaflow_cigar_batch_create: bad argument");
  BF_REQUIRE(ctx->has_index, -22, "bwaflow_cigar_batch_create: the context has no index");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_cigar_batch *b = new bwaflow_cigar_batch();
  b->ctx = ctx; b->max_reads = max_reads; b->max_bases = max_bases; b->max_jobs = max_jobs;
  if (po) b->po = *po; else bwaflow_pe_opt_default(&b->po);
  BF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (b->tm.init()) return -5;
  BF_CUDA(cudaMalloc(&b->d_bases, max_bases + 16));
  BF_CUDA(cudaMalloc(&b->d_jobs, max_jobs * sizeof(CgJob)));
  BF_CUDA(cudaMalloc(&b->d_out, max_jobs * sizeof(CgOut)));
  for (int i = 0; i < 2; ++i) { BF_CUDA(cudaMalloc(&b->d_order[i], max_jobs * 4)); BF_CUDA(cudaMalloc(&b->d_keys[i], max_jobs * 4)); }
  BF_CUDA(cudaMalloc(&b->d_sz, 2 * (max_jobs + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_off, 2 * (max_jobs + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_ctl, 4 * sizeof(int)));
  size_t t1 = 0, t2 = 0;
  BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t1, b->d_sz, b->d_off, (int)(max_jobs + 1), b->stream));
  cub::DoubleBuffer<uint32_t> k(b->d_keys[0], b->d_keys[1]), v(b->d_order[0], b->d_order[1]);
  BF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, k, v, (int)max_jobs, 0, 32, b->stream));
  b->tmp_bytes = (t1 > t2 ? t1 : t2) + 256;
  BF_CUDA(cudaMalloc(&b->d_tmp, b->tmp_bytes));
  *out = b;
  return 0;
}

void bwaflow_cigar_batch_destroy(bwaflow_cigar_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  cudaFree(b->d_bases); cudaFree(b->d_jobs); cudaFree(b->d_out);
  for (int i = 0; i < 2; ++i) { cudaFree(b->d_order[i]); cudaFree(b->d_keys[i]); }
  cudaFree(b->d_sz); cudaFree(b->d_off); cudaFree(b->d_ctl); cudaFree(b->d_tmp);
  cudaFree(b->d_z); cudaFree(b->d_cig); cudaFree(b->d_md); cudaFree(b->d_cig_out); cudaFree(b->d_md_out); cudaFree(b->d_geh);
  b->tm.destroy();
  cudaStreamDestroy(b->stream);
  delete b;
}

int bwaflow_cigar_submit(bwaflow_cigar_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off, int64_t n_jobs,
                         const int32_t *job_read, const bwaflow_alnreg_t *job_reg)
{
  BF_REQUIRE(b && bases && read_off && (n_jobs == 0 || (job_read && job_reg)), -22, "bwaflow_cigar_submit: null argument");
  BF_REQUIRE(n_reads <= b->max_reads && read_off[n_reads] <= b->max_bases && n_jobs >= 0 && n_jobs <= b->max_jobs, -7,
             "bwaflow_cigar_submit: batch capacity exceeded");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  const bwaflow_opt_t &o = b->ctx->par.opt;
  b->n_jobs = n_jobs;
  b->jobs.resize((size_t)n_jobs); b->keys.resize((size_t)n_jobs); b->vals.resize((size_t)n_jobs);
  b->regs.assign(job_reg, job_reg + n_jobs);
  size_t zo = 0, co = 0, mo = 0;
  int wide = 0, w_max = 0, lq_max = 1;
  for (int64_t j = 0; j < n_jobs; ++j) {
    const bwaflow_alnreg_t &r = job_reg[j];
    const int32_t rd = job_read[j];
    BF_REQUIRE(rd >= 0 && rd < n_reads && r.rb >= 0 && r.re > r.rb && r.qb >= 0 && r.qe > r.qb && r.qe <= read_off[rd + 1] - read_off[rd], -22,
               "bwaflow_cigar_submit: job %lld is not a mapped region of its read", (long long)j);
    CgJob &g = b->jobs[(size_t)j];
    memset(&g, 0, sizeof(g));
    g.rb = r.rb; g.re = r.re; g.q_src = read_off[rd]; g.qb = r.qb; g.qe = r.qe; g.l_query = (int)(read_off[rd + 1] - read_off[rd]);
    g.truesc = r.truesc; g.w_reg = r.w;
    const int lq = r.qe - r.qb, rlen = (int)(r.re - r.rb);
    // widest band any of the tries can use
    g.w_cap = cg_band(lq, rlen, o.mat[0], o.o_del, o.e_del, o.o_ins, o.e_ins, o.w << 2);
    const int n_col = lq < 2 * g.w_cap + 1 ? lq : 2 * g.w_cap + 1;
    g.z_off = (int64_t)zo; zo += (size_t)n_col * rlen;
    g.cig_cap = lq + rlen + 4; g.cig_off = (int64_t)co; co += (size_t)g.cig_cap;
    g.md_cap = lq + 3 * rlen + 16; g.md_off = (int64_t)mo; mo += (size_t)g.md_cap;
    // queue order: band class, then target length (lanes of a warp walk loops of the same shape)
    int w_try = cg_infer_bw(lq, rlen, r.truesc, o.a, o.o_del, o.e_del);
    { const int t = cg_infer_bw(lq, rlen, r.truesc, o.a, o.o_ins, o.e_ins); w_try = w_try > t ? w_try : t; }
    const int w_first = cg_band(lq, rlen, o.mat[0], o.o_del, o.e_del, o.o_ins, o.e_ins, w_try < o.w << 2 ? w_try : o.w << 2);
    b->keys[(size_t)j] = (uint32_t)(w_first > 1023 ? 1023 : w_first) << 20 | (uint32_t)(rlen > 0xFFFFF ? 0xFFFFF : rlen);
    b->vals[(size_t)j] = (uint32_t)j;
    if (2 * g.w_cap + 2 > 256) wide = 1;
    if (g.w_cap > w_max) w_max = g.w_cap;
    if (lq > lq_max) lq_max = lq;
  }
  int r;
  if ((r = cg_grow(&b->d_z, &b->z_cap, zo + 16))) return r;
  if ((r = cg_grow(&b->d_cig, &b->cig_cap, co + 16))) return r;
  if ((r = cg_grow(&b->d_md, &b->md_cap, mo + 16))) return r;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  BF_CUDA(cudaMemcpyAsync(b->d_bases, bases, read_off[n_reads], cudaMemcpyHostToDevice, b->stream));
  BF_CUDA(cudaMemsetAsync(b->d_ctl, 0, 4 * sizeof(int), b->stream));
  if (n_jobs) {
    BF_CUDA(cudaMemcpyAsync(b->d_jobs, b->jobs.data(), n_jobs * sizeof(CgJob), cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_keys[0], b->keys.data(), n_jobs * 4, cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_order[0], b->vals.data(), n_jobs * 4, cudaMemcpyHostToDevice, b->stream));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  if (n_jobs) {
    cub::DoubleBuffer<uint32_t> k(b->d_keys[0], b->d_keys[1]), v(b->d_order[0], b->d_order[1]);
    size_t tmp = b->tmp_bytes;
    BF_CUDA(cub::DeviceRadixSort::SortPairs(b->d_tmp, tmp, k, v, (int)n_jobs, 0, 32, b->stream));
    if ((r = cg_launch(b, n_jobs, v.Current(), w_max, lq_max, wide))) return r;
    cigar_sizes_kernel<<<(unsigned)((n_jobs + 1 + 255) / 256), 256, 0, b->stream>>>(b->d_out, n_jobs, b->d_sz);
    for (int k2 = 0; k2 < 2; ++k2) {
      tmp = b->tmp_bytes;
      BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_sz + k2 * (n_jobs + 1), b->d_off + k2 * (n_jobs + 1), (int)(n_jobs + 1), b->stream));
    }
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[2], b->stream));
  BF_CUDA(cudaGetLastError());
  b->pending = true;
  return 0;
}

int bwaflow_cigar_collect(bwaflow_cigar_batch_t *b, const bwaflow_aln_t **aln, const uint32_t **cigar_pool, const char **md_pool,
                          bwaflow_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_cigar_collect: nothing submitted");
  const int64_t n = b->n_jobs;
  b->pending = false;
  b->outs.resize((size_t)n); b->offs.resize((size_t)(2 * (n + 1)));
  b->aln.resize((size_t)n);
  if (n) {
    BF_CUDA(cudaMemcpyAsync(b->outs.data(), b->d_out, n * sizeof(CgOut), cudaMemcpyDeviceToHost, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->offs.data(), b->d_off, 2 * (n + 1) * 8, cudaMemcpyDeviceToHost, b->stream));
  }
  BF_CUDA(cudaStreamSynchronize(b->stream));
  const size_t n_cig = n ? (size_t)b->offs[(size_t)n] : 0, n_md = n ? (size_t)b->offs[(size_t)(2 * n + 1)] : 0;
  if (n) {
    int r;
    if ((r = cg_grow(&b->d_cig_out, &b->out_cig_cap, n_cig + 16))) return r;
    if ((r = cg_grow(&b->d_md_out, &b->out_md_cap, n_md + 16))) return r;
    cigar_gather_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, b->stream>>>(b->d_jobs, b->d_out, n, b->d_off, b->d_cig, b->d_md, b->d_cig_out, b->d_md_out);
    b->cig.resize(n_cig + 1); b->md.resize(n_md + 1);
    BF_CUDA(cudaMemcpyAsync(b->cig.data(), b->d_cig_out, n_cig * 4, cudaMemcpyDeviceToHost, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->md.data(), b->d_md_out, n_md, cudaMemcpyDeviceToHost, b->stream));
    BF_CUDA(cudaEventRecord(b->tm.ev[3], b->stream));
    BF_CUDA(cudaStreamSynchronize(b->stream));
  }
  const bwaflow_opt_t &o = b->ctx->par.opt;
  const int n_seqs = b->ctx->ix.n_seqs;
  for (int64_t j = 0; j < n; ++j) {
    const CgOut &g = b->outs[(size_t)j];
    BF_REQUIRE(g.err == 0, -34, "CIGAR job %lld failed on the device (code %d: 1 invalid interval, 2 band beyond the planned capacity, 3/4 CIGAR / MD buffer)",
               (long long)j, g.err);
    const bwaflow_alnreg_t &r = b->regs[(size_t)j];
    bwaflow_aln_t &a = b->aln[(size_t)j];
    memset(&a, 0, sizeof(a));
    int rid = 0;                                           // bns_pos2rid of the (squeezed) forward position
    { int left = 0, mid = 0, right = n_seqs;
      while (left < right) { mid = (left + right) >> 1;
        if (g.pos >= b->ctx->h_ctg_off[mid]) { if (mid == n_seqs - 1) break; if (g.pos < b->ctx->h_ctg_off[mid + 1]) break; left = mid + 1; } else right = mid; }
      rid = mid; }
    a.rid = rid; a.pos = g.pos - b->ctx->h_ctg_off[rid];
    a.flag = r.secondary >= 0 ? 0x100 : 0;
    a.is_rev = g.is_rev; a.is_alt = (int32_t)(r.n_comp_is_alt >> 30);
    a.mapq = r.secondary < 0 ? approx_mapq_se(o, b->po, r) : 0;
    a.NM = g.NM; a.n_cigar = g.n_cigar; a.score = r.score; a.sub = r.sub > r.csub ? r.sub : r.csub; a.alt_sc = r.alt_sc;
    a.global_score = g.score;
    a.cigar_off = b->offs[(size_t)j]; a.md_off = b->offs[(size_t)(n + 1 + j)];
  }
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->tm.ev[0], b->tm.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, b->tm.ev[1], b->tm.ev[2]);
    if (n) cudaEventElapsedTime(&t->d2h_ms, b->tm.ev[2], b->tm.ev[3]);
    t->total_ms = t->h2d_ms + t->kernel_ms + t->d2h_ms;
    t->n_launches = n ? 5 : 0;
  }
  if (aln) *aln = b->aln.data();
  if (cigar_pool) *cigar_pool = b->cig.data();
  if (md_pool) *md_pool = b->md.data();
  return 0;
}

int bwaflow_cigar_global_scores(bwaflow_cigar_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off, int upload,
                                int64_t n_jobs, const bwaflow_gscore_job_t *jobs, int32_t *scores)
{
  BF_REQUIRE(b && !b->pending && read_off && (n_jobs == 0 || (jobs && scores)), -22, "bwaflow_cigar_global_scores: bad argument or a submit is pending");
  BF_REQUIRE(n_reads <= b->max_reads && read_off[n_reads] <= b->max_bases && n_jobs >= 0 && n_jobs <= b->max_jobs, -7,
             "bwaflow_cigar_global_scores: batch capacity exceeded");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  const bwaflow_opt_t &o = b->ctx->par.opt;
  if (upload) { BF_REQUIRE(bases, -22, "bwaflow_cigar_global_scores: no bases"); BF_CUDA(cudaMemcpyAsync(b->d_bases, bases, read_off[n_reads], cudaMemcpyHostToDevice, b->stream)); }
  if (n_jobs == 0) { BF_CUDA(cudaStreamSynchronize(b->stream)); return 0; }
  b->jobs.resize((size_t)n_jobs); b->vals.resize((size_t)n_jobs);
  int wide = 0, w_max = 0, lq_max = 1;
  for (int64_t j = 0; j < n_jobs; ++j) {
    const bwaflow_gscore_job_t &q = jobs[j];
    BF_REQUIRE(q.read >= 0 && q.read < n_reads && q.qb >= 0 && q.qe > q.qb && q.qe <= read_off[q.read + 1] - read_off[q.read] && q.rb >= 0 && q.re > q.rb && q.w >= 0,
               -22, "bwaflow_cigar_global_scores: job %lld is not an interval of its read", (long long)j);
    CgJob &g = b->jobs[(size_t)j];
    memset(&g, 0, sizeof(g));
    g.rb = q.rb; g.re = q.re; g.q_src = read_off[q.read]; g.qb = q.qb; g.qe = q.qe; g.l_query = (int)(read_off[q.read + 1] - read_off[q.read]);
    g.w_reg = q.w; g.mode = 1;
    const int lq = q.qe - q.qb, rlen = (int)(q.re - q.rb);
    g.w_cap = cg_band(lq, rlen, o.mat[0], o.o_del, o.e_del, o.o_ins, o.e_ins, q.w);
    b->vals[(size_t)j] = (uint32_t)j;
    if (2 * g.w_cap + 2 > 256) wide = 1;
    if (g.w_cap > w_max) w_max = g.w_cap;
    if (lq > lq_max) lq_max = lq;
  }
  BF_CUDA(cudaMemsetAsync(b->d_ctl, 0, 4 * sizeof(int), b->stream));
  BF_CUDA(cudaMemcpyAsync(b->d_jobs, b->jobs.data(), n_jobs * sizeof(CgJob), cudaMemcpyHostToDevice, b->stream));
  BF_CUDA(cudaMemcpyAsync(b->d_order[0], b->vals.data(), n_jobs * 4, cudaMemcpyHostToDevice, b->stream));
  int r;
  if ((r = cg_launch(b, n_jobs, b->d_order[0], w_max, lq_max, wide))) return r;
  BF_CUDA(cudaGetLastError());
  b->outs.resize((size_t)n_jobs);
  BF_CUDA(cudaMemcpyAsync(b->outs.data(), b->d_out, n_jobs * sizeof(CgOut), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaStreamSynchronize(b->stream));
  for (int64_t j = 0; j < n_jobs; ++j) {
    BF_REQUIRE(b->outs[(size_t)j].err == 0, -34, "global-score job %lld failed on the device (code %d)", (long long)j, b->outs[(size_t)j].err);
    scores[j] = b->outs[(size_t)j].score;
  }
  return 0;
}

} // extern "C"
