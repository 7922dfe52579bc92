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

extern "C" {

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
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  bwaflow_opt_t def;
  if (!opt) { bwaflow_opt_default(&def); opt = &def; }
  int r = bwaflow_ctx_set_opt(c, opt);
  if (r) { free(c); return r; }

  if (idx) {
    BF_REQUIRE(idx->sa_intv > 0 && (idx->sa_intv & (idx->sa_intv - 1)) == 0, -22, "sa_intv must be a power of two");
    DevIndex &d = c->ix;
    d.primary = idx->primary; d.seq_len = idx->seq_len;
    memcpy(d.L2, idx->L2, sizeof(d.L2));
    d.sa_shift = 0; while ((1LL << d.sa_shift) < idx->sa_intv) ++d.sa_shift;
    d.l_pac = idx->l_pac; d.n_seqs = idx->n_seqs;
    // bwt_size words; pad to whole 64-byte blocks so 128-bit block loads never run off the end
    size_t bwt_bytes = (size_t)idx->bwt_size * 4, bwt_pad = (bwt_bytes + 63) / 64 * 64 + 64;
    BF_CUDA(cudaMalloc(&c->d_bwt, bwt_pad));
    BF_CUDA(cudaMemset(c->d_bwt, 0, bwt_pad));
    BF_CUDA(cudaMemcpy(c->d_bwt, idx->bwt, bwt_bytes, cudaMemcpyHostToDevice));
    if ((r = upload(&c->d_sa, idx->sa, (size_t)idx->n_sa * 8))) return r;
    if ((r = upload(&c->d_pac, idx->pac, (size_t)(idx->l_pac / 4 + 1)))) return r;
    if ((r = upload(&c->d_ctg_off, idx->ctg_offset, (size_t)idx->n_seqs * 8))) return r;
    if ((r = upload(&c->d_ctg_len, idx->ctg_len, (size_t)idx->n_seqs * 4))) return r;
    d.bwt = (const uint4 *)c->d_bwt; d.sa = (const uint64_t *)c->d_sa; d.pac = (const uint8_t *)c->d_pac;
    d.ctg_offset = (const int64_t *)c->d_ctg_off; d.ctg_len = (const int32_t *)c->d_ctg_len;
    c->has_index = true;
  }
  *out = c;
  return 0;
}

void bwaflow_ctx_destroy(bwaflow_ctx_t *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_bwt); cudaFree(c->d_sa); cudaFree(c->d_pac); cudaFree(c->d_ctg_off); cudaFree(c->d_ctg_len);
  free(c);
}

} // extern "C"
