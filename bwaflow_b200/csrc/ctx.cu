// Context: one per GPU.  Uploads the FM-index and the 2-bit reference to HBM once, the way
// BWAOCLEnv's constructor does for the FPGA boards (src/fpga/BWAOCLEnv.h:67-216).
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "internal.cuh"

namespace bwaflow {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void fill_score(const bwaflow_opt_t &o, DevScore &s)
{
  s.o_del = o.o_del; s.e_del = o.e_del; s.o_ins = o.o_ins; s.e_ins = o.e_ins; s.zdrop = o.zdrop;
  s.smax = 0;
  for (int i = 0; i < 25; ++i) if (o.mat[i] > s.smax) s.smax = o.mat[i];
  for (int t = 0; t < 5; ++t) {
    uint32_t lo = 0;
    for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)o.mat[t * 5 + q] << (8 * q);
    s.row_lo[t] = lo;
    s.row_hi[t] = (uint32_t)(uint8_t)o.mat[t * 5 + 4];
  }
}

} // namespace bwaflow

using namespace bwaflow;
namespace bwaflow { int densify_sa(bwaflow_ctx *c, int dense_shift, uint64_t n_sa, void **d_dense); int build_level_tables(bwaflow_ctx *c, int K, void **d_tab); }

// forward + reverse complement text from the 2-bit pac, one thread per 32-symbol word
__global__ void ctx_text_kernel(const uint8_t *pac, int64_t L, uint64_t *T, int64_t n_words)
{
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const int64_t n = 2 * L;
  uint64_t v = 0;
  for (int s = 0; s < 32; ++s) {
    const int64_t p = w * 32 + s;
    int c = 0;
    if (p < L) c = (pac[p >> 2] >> ((~p & 3) << 1)) & 3;
    else if (p < n) { const int64_t f = n - 1 - p; c = 3 - ((pac[f >> 2] >> ((~f & 3) << 1)) & 3); }
    v = v << 2 | (uint64_t)c;
  }
  T[w] = v;
}
// bwa layout (4 x u64 counts + 8 x u32 symbols per 128 symbols, bwa/bwt.h:72-73) -> occ32 blocks (internal.cuh).
// Pass 1, one thread per 64-symbol block: absolute suffix counts S1,S2,S3 (symbols >0, >1, >2 before the block).
__global__ void occ32_count_kernel(const uint32_t *bwt, uint64_t n64, uint4 *occ, uint64_t *absS)
{
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n64) return;
  const uint32_t *blk = bwt + (b >> 1) * 16;
  uint64_t c1 = (uint64_t)blk[3] << 32 | blk[2], c2 = (uint64_t)blk[5] << 32 | blk[4], c3 = (uint64_t)blk[7] << 32 | blk[6];
  const uint32_t *sym = blk + 8;
  if (b & 1) {                                         // second half of the bwa block: add its first 64 symbols
    for (int j = 0; j < 4; ++j) {
      const uint32_t w = sym[j], lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
      c1 += __popc(~hi & lo & 0x55555555u); c2 += __popc(hi & ~lo); c3 += __popc(hi & lo);
    }
    sym += 4;
  }
  absS[b] = c1 + c2 + c3; absS[n64 + b] = c2 + c3; absS[2 * n64 + b] = c3;
  occ[2 * b + 1] = make_uint4(sym[0], sym[1], sym[2], sym[3]);
}
// Pass 2: counts relative to the block's 2^31-symbol superblock, as u32.
__global__ void occ32_pack_kernel(uint64_t n64, uint4 *occ, const uint64_t *absS, DevIndex ix)
{
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n64) return;
  const uint64_t *sb = ix.sbt + (b >> 25) * 5;
  occ[2 * b] = make_uint4((uint32_t)(absS[b] - sb[1]), (uint32_t)(absS[n64 + b] - sb[2]), (uint32_t)(absS[2 * n64 + b] - sb[3]), 0u);
}

// dependency-free DPX issue-rate probe: 8 independent accumulator chains per thread
template <int PACKED>
__global__ void dpx_peak_kernel(unsigned *out, int iters, unsigned seed)
{
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
  const unsigned b = seed * 2654435761u, c = seed ^ 0x01010101u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (PACKED) {
        a0 = __vimax3_s16x2(a0, b, c); a1 = __viaddmax_s16x2_relu(a1, b, c); a2 = __vimax3_s16x2(a2, b, c); a3 = __viaddmax_s16x2_relu(a3, b, c);
        a4 = __vimax3_s16x2(a4, b, c); a5 = __viaddmax_s16x2_relu(a5, b, c); a6 = __vimax3_s16x2(a6, b, c); a7 = __viaddmax_s16x2_relu(a7, b, c);
      } else {
        a0 = __vimax3_s32(a0, b, c); a1 = __viaddmax_s32_relu(a1, b, c); a2 = __vimax3_s32(a2, b, c); a3 = __viaddmax_s32_relu(a3, b, c);
        a4 = __vimax3_s32(a4, b, c); a5 = __viaddmax_s32_relu(a5, b, c); a6 = __vimax3_s32(a6, b, c); a7 = __viaddmax_s32_relu(a7, b, c);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

extern "C" {

int bwaflow_dpx_peak(bwaflow_ctx_t *ctx, double *packed, double *scalar)
{
  BF_REQUIRE(ctx, -22, "bwaflow_dpx_peak: null context");
  BF_CUDA(cudaSetDevice(ctx->device));
  const int threads = 256, blocks = ctx->n_sm * 8, iters = 4096;
  unsigned *d = 0;
  BF_CUDA(cudaMalloc(&d, (size_t)threads * blocks * 4));
  cudaEvent_t e0, e1;
  BF_CUDA(cudaEventCreate(&e0)); BF_CUDA(cudaEventCreate(&e1));
  for (int mode = 0; mode < 2; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      BF_CUDA(cudaEventRecord(e0, 0));
      if (mode == 0) dpx_peak_kernel<1><<<blocks, threads>>>(d, iters, 12345u + rep);
      else dpx_peak_kernel<0><<<blocks, threads>>>(d, iters, 12345u + rep);
      BF_CUDA(cudaEventRecord(e1, 0));
      BF_CUDA(cudaEventSynchronize(e1));
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)threads * blocks * iters * 64.0 / (best * 1e-3);
    if (mode == 0) { if (packed) *packed = ops; } else if (scalar) *scalar = ops;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  BF_CUDA(cudaGetLastError());
  return 0;
}

const char *bwaflow_last_error(void) { return g_err; }

int bwaflow_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void bwaflow_opt_default(bwaflow_opt_t *o)
{ // mem_opt_init, bwa/bwamem.c:48-84 ; bwa_fill_scmat bwa/bwa.c:109-119
  memset(o, 0, sizeof(*o));
  o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
  o->pen_clip5 = o->pen_clip3 = 5; o->w = 100; o->zdrop = 100;
  o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500; o->max_mem_intv = 20;
  o->split_factor = 1.5f;
  int k = 0;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) o->mat[k++] = (int8_t)(i == j ? o->a : -o->b);
    o->mat[k++] = -1;
  }
  for (int j = 0; j < 5; ++j) o->mat[k++] = -1;
}

void *bwaflow_host_alloc(uint64_t bytes)
{
  void *p = 0;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { set_error("cudaMallocHost(%llu) failed", (unsigned long long)bytes); cudaGetLastError(); return 0; }
  return p;
}
void bwaflow_host_free(void *p) { if (p) cudaFreeHost(p); }

static int upload(void **dst, const void *src, size_t bytes)
{
  BF_CUDA(cudaMalloc(dst, bytes ? bytes : 16));
  if (bytes) BF_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

int bwaflow_ctx_set_opt(bwaflow_ctx_t *ctx, const bwaflow_opt_t *opt)
{
  BF_REQUIRE(ctx && opt, -22, "bwaflow_ctx_set_opt: null argument");
  BF_REQUIRE(opt->e_del > 0 && opt->e_ins > 0, -22, "gap extension penalties must be positive");
  ctx->par.opt = *opt;
  fill_score(*opt, ctx->par.sc);
  return 0;
}

static int ctx_init(bwaflow_ctx_t *c, int device, const cudaDeviceProp &prop, const bwaflow_index_t *idx, const bwaflow_opt_t *opt)
{
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  bwaflow_opt_t def;
  if (!opt) { bwaflow_opt_default(&def); opt = &def; }
  int r = bwaflow_ctx_set_opt(c, opt);
  if (r) return r;

  if (const char *s = getenv("BWAFLOW_L2_FETCH")) BF_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(s)));
  if (idx) {
    c->ix.occ_cs = 1;                           // streaming loads: random index sectors must not evict the per-lane lists from L2
    if (const char *s = getenv("BWAFLOW_OCC_LOAD")) c->ix.occ_cs = strcmp(s, "ldg") != 0;
    if (getenv("BWAFLOW_DEBUG_COUNT_TEXT_STEPS")) c->ix.occ_cs |= 2;
    BF_REQUIRE(idx->sa_intv > 0 && (idx->sa_intv & (idx->sa_intv - 1)) == 0, -22, "sa_intv must be a power of two");
    DevIndex &d = c->ix;
    d.primary = idx->primary; d.seq_len = idx->seq_len;
    memcpy(d.L2, idx->L2, sizeof(d.L2));
    d.sa_shift = 0; while ((1LL << d.sa_shift) < idx->sa_intv) ++d.sa_shift;
    d.l_pac = idx->l_pac; d.n_seqs = idx->n_seqs;
    // bwt_size words; pad to whole 64-byte blocks so 128-bit block loads never run off the end
    size_t bwt_bytes = (size_t)idx->bwt_size * 4, bwt_pad = (bwt_bytes + 63) / 64 * 64 + 256;
    BF_CUDA(cudaMalloc(&c->d_bwt, bwt_pad));
    BF_CUDA(cudaMemset(c->d_bwt, 0, bwt_pad));
    BF_CUDA(cudaMemcpy(c->d_bwt, idx->bwt, bwt_bytes, cudaMemcpyHostToDevice));
    if ((r = upload(&c->d_sa, idx->sa, (size_t)idx->n_sa * 8))) return r;
    if ((r = upload(&c->d_pac, idx->pac, (size_t)(idx->l_pac / 4 + 1)))) return r;
    if ((r = upload(&c->d_ctg_off, idx->ctg_offset, (size_t)idx->n_seqs * 8))) return r;
    c->h_ctg_off = (int64_t *)malloc((size_t)idx->n_seqs * 8); c->h_ctg_len = (int32_t *)malloc((size_t)idx->n_seqs * 4);
    memcpy(c->h_ctg_off, idx->ctg_offset, (size_t)idx->n_seqs * 8); memcpy(c->h_ctg_len, idx->ctg_len, (size_t)idx->n_seqs * 4);
    if ((r = upload(&c->d_ctg_len, idx->ctg_len, (size_t)idx->n_seqs * 4))) return r;
    {
      // re-lay the occurrence array out for the device (occ32), then drop the upload of bwa's layout
      const uint64_t n64 = (idx->seq_len + 63) / 64 + 1;
      BF_REQUIRE(n64 <= (8ull << 25), -22, "reference too long for the superblock table (max 2^34 BWT symbols)");
      uint4 *d_occ = 0; uint64_t *d_abs = 0;
      BF_CUDA(cudaMalloc(&d_occ, (n64 + 2) * 32));
      BF_CUDA(cudaMemset(d_occ, 0, (n64 + 2) * 32));
      BF_CUDA(cudaMalloc(&d_abs, 3 * n64 * 8));
      occ32_count_kernel<<<(unsigned)((n64 + 255) / 256), 256>>>((const uint32_t *)c->d_bwt, n64, d_occ, d_abs);
      BF_CUDA(cudaGetLastError());
      memset(d.sbt, 0, sizeof(d.sbt));
      for (uint64_t j = 0; (j << 25) < n64; ++j) {
        d.sbt[j * 5 + 0] = j << 31;
        for (int t = 0; t < 3; ++t)
          BF_CUDA(cudaMemcpy(&d.sbt[j * 5 + 1 + t], d_abs + t * n64 + (j << 25), 8, cudaMemcpyDeviceToHost));
      }
      occ32_pack_kernel<<<(unsigned)((n64 + 255) / 256), 256>>>(n64, d_occ, d_abs, d);
      BF_CUDA(cudaGetLastError());
      BF_CUDA(cudaDeviceSynchronize());
      cudaFree(d_abs); cudaFree(c->d_bwt);
      c->d_bwt = d_occ;
      d.occ = d_occ;
    }
    d.sa = (const uint64_t *)c->d_sa; d.pac = (const uint8_t *)c->d_pac;
    d.ctg_offset = (const int64_t *)c->d_ctg_off; d.ctg_len = (const int32_t *)c->d_ctg_len;
    // Denser suffix-array sampling in HBM.  Default: the FULL suffix array (16 bytes per reference base, 49.6 GB of the 180 GB
    // at 3.1 Gbp) instead of bwa's every 32nd row, so that a seed position is one read instead of a walk of ~16 dependent
    // random index reads; BWAFLOW_SA_INTV=2,4,..,32 trades memory back.  bwt_sa values do not depend on the sampling.
    int dense_intv = 1;
    if (const char *s = getenv("BWAFLOW_SA_INTV")) dense_intv = atoi(s);
    int dense_shift = 0;
    while ((2 << dense_shift) <= dense_intv) ++dense_shift;
    if (dense_intv >= 1 && dense_shift < d.sa_shift) {
      void *d_dense = nullptr;
      if ((r = densify_sa(c, dense_shift, idx->n_sa, &d_dense))) return r;
      cudaFree(c->d_sa);
      c->d_sa = d_dense;
      d.sa = (const uint64_t *)d_dense; d.sa_shift = dense_shift;
    }
    // Level tables of the seeding kernels (smem_fast.cu): intervals of all patterns of 1..K bases, 16 bytes each.  K follows the
    // index size (about log4(rows) - 1: the deepest level whose patterns still occur several times), at most 15
    // (22.9 GB on a human-scale index).  BWAFLOW_KMER_K overrides; 0 disables.  If the tables do not fit, K shrinks.
    d.ktab = nullptr; d.ktab_k = 0;
    int kk = 0;
    while ((4ull << (2 * kk)) <= d.seq_len) ++kk;          // floor(log4(seq_len))
    kk = kk - 1 > 15 ? 15 : kk - 1;
    if (const char *s = getenv("BWAFLOW_KMER_K")) kk = atoi(s);
    if (kk > 15) kk = 15;
    for (; kk >= 2; --kk) {
      size_t free_b = 0, total_b = 0;
      BF_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const uint64_t need = (((1ull << (2 * (kk + 1))) - 4) / 3) * 16;
      if (need + (24ull << 30) > free_b && need > (1ull << 30)) continue;   // leave room for the batches
      if ((r = build_level_tables(c, kk, &c->d_ktab))) return r;
      d.ktab = (const uint4 *)c->d_ktab; d.ktab_k = kk;
      break;
    }
    // The 2-bit text (with the full suffix array only): unique intervals are extended by comparing bases against it.
    d.text = nullptr;
    int text_mode = 1;
    if (const char *s = getenv("BWAFLOW_TEXT_MODE")) text_mode = atoi(s);
    if (text_mode && d.sa_shift == 0) {
      const int64_t n = (int64_t)d.seq_len, n_words = (n + 31) / 32 + 4;
      BF_CUDA(cudaMalloc(&c->d_text, n_words * 8));
      ctx_text_kernel<<<(unsigned)((n_words + 255) / 256), 256>>>((const uint8_t *)c->d_pac, d.l_pac, (uint64_t *)c->d_text, n_words);
      BF_CUDA(cudaGetLastError());
      BF_CUDA(cudaDeviceSynchronize());
      d.text = (const uint64_t *)c->d_text;
    }
    c->has_index = true;
  }
  return 0;
}

int bwaflow_ctx_create(int device, const bwaflow_index_t *idx, const bwaflow_opt_t *opt, bwaflow_ctx_t **out)
{
  BF_REQUIRE(out, -22, "bwaflow_ctx_create: out is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); set_error("no CUDA device available (this backend has no CPU fallback)"); return -19; }
  BF_REQUIRE(device >= 0 && device < n, -19, "device %d out of range (have %d)", device, n);
  BF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BF_CUDA(cudaGetDeviceProperties(&prop, device));
  BF_REQUIRE(prop.major >= 10, -19, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

  bwaflow_ctx_t *c = (bwaflow_ctx_t *)calloc(1, sizeof(bwaflow_ctx_t));
  BF_REQUIRE(c, -12, "bwaflow_ctx_create: out of host memory");
  const int r = ctx_init(c, device, prop, idx, opt);
  if (r) { bwaflow_ctx_destroy(c); cudaGetLastError(); return r; }      // one cleanup path: everything uploaded so far is released
  *out = c;
  return 0;
}

struct bwaflow_stopwatch { int device; cudaEvent_t a, b; };

int bwaflow_stopwatch_create(bwaflow_ctx_t *ctx, bwaflow_stopwatch_t **out)
{
  BF_REQUIRE(ctx && out, -22, "bwaflow_stopwatch_create: null argument");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_stopwatch *w = (bwaflow_stopwatch *)calloc(1, sizeof(*w));
  w->device = ctx->device;
  BF_CUDA(cudaEventCreate(&w->a)); BF_CUDA(cudaEventCreate(&w->b));
  *out = w;
  return 0;
}
int bwaflow_stopwatch_start(bwaflow_stopwatch_t *w)
{
  BF_CUDA(cudaSetDevice(w->device));
  BF_CUDA(cudaDeviceSynchronize());
  BF_CUDA(cudaEventRecord(w->a, 0));
  return 0;
}
int bwaflow_stopwatch_stop(bwaflow_stopwatch_t *w, float *ms)
{
  BF_CUDA(cudaSetDevice(w->device));
  BF_CUDA(cudaDeviceSynchronize());
  BF_CUDA(cudaEventRecord(w->b, 0));
  BF_CUDA(cudaEventSynchronize(w->b));
  BF_CUDA(cudaEventElapsedTime(ms, w->a, w->b));
  return 0;
}
void bwaflow_stopwatch_destroy(bwaflow_stopwatch_t *w)
{
  if (!w) return;
  cudaEventDestroy(w->a); cudaEventDestroy(w->b); free(w);
}
int bwaflow_ctx_sync(bwaflow_ctx_t *ctx)
{
  BF_REQUIRE(ctx, -22, "bwaflow_ctx_sync: null context");
  BF_CUDA(cudaSetDevice(ctx->device));
  BF_CUDA(cudaDeviceSynchronize());
  return 0;
}

void bwaflow_ctx_destroy(bwaflow_ctx_t *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  free(c->h_ctg_off); free(c->h_ctg_len);
  cudaFree(c->d_bwt); cudaFree(c->d_sa); cudaFree(c->d_text); cudaFree(c->d_ktab); cudaFree(c->d_pac); cudaFree(c->d_ctg_off); cudaFree(c->d_ctg_len);
  free(c);
}

} // extern "C"
