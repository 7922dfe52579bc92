// Fused hot path: reads -> alignment regions with every intermediate resident in HBM.
//   stage 2   smem_kernel + seedpos_kernel                         (smem.cu)
//   chaining  chain_kernel: mem_chain_postprocess + mem_chain_flt  (chain_core.h, one read per thread)
//   stage 4   c2a_kernel                                            (chain2aln.cu)
// In kflow terms this is SeqsToChainsCUDA producing tag=1 records whose chains stay on the device, ChainsPipe passing
// them through (src/Pipeline.cpp:409), and ChainsToRegionsCUDA consuming them.  Only the reads go
// up and only the regions come down; the host sees three tiny size read-backs per batch.
#include <cub/device/device_scan.cuh>
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "batches.cuh"
#include "chain_core.h"

namespace bwaflow {

using namespace chain;

struct ChainArgs {
  int64_t n_reads; const int64_t *read_off;
  const bwaflow_intv_t *intv; const int64_t *intv_off; const int64_t *pos; const int64_t *pos_off;
  int64_t out_cap, pos_cap;          // capacities of intv / pos: a read beyond them is left to the redo (smem.cu, device-driven tail)
  // per-read scratch, carved by the read's first position index
  SeedNode *seeds; Chain *chains; Chain *sorted; int *order; int *kept; Node *nodes;
  int64_t *n_chains, *n_seeds; float *frac_rep; int *err;
  int *heavy;                        // [0] reads set aside for chain_heavy_kernel, [1] its queue head, [2..] their ids
  bwaflow_chain_opt_t opt; Ctg ctg;
};

__device__ __forceinline__ Scratch read_scratch(const ChainArgs &a, int64_t r)
{
  const int64_t pbase = a.pos_off[a.intv_off[r]], e = pbase + r;       // P_r + 1 entries per read
  Scratch S;
  S.seeds = a.seeds + e; S.chains = a.chains + e; S.sorted = a.sorted + e; S.order = a.order + e; S.kept = a.kept + e;
  S.nodes = a.nodes + (pbase / 4 + 4 * r);
  S.n_seeds = S.n_chains = S.n_nodes = 0; S.root = 0;
  return S;
}

// The overlap filter of mem_chain_flt (bwa/bwamem.c:355-380) for ONE read by a whole warp.  Chain i is tested against the chains kept
// so far, 32 at a time; the sequential scan stops at the first kept chain that drops it, so every test the scan would not have
// reached is discarded: `first` is only set (and large_ovlp only raised) up to and including that entry, exactly as the scan does.
// Same arrays and the same result as chain::flt_filter.
constexpr int CHAIN_FLT_WARP_MIN = 64;     // chains of a read from which the warp takes the filter over (quadratic work: a read inside a repeat has hundreds)
__device__ int flt_filter_warp(Chain *a, KeptEnt *K, int *F, int *kept, int n_chn, const bwaflow_chain_opt_t &o, int lane)
{
  const unsigned full = 0xffffffffu;
  if (lane == 0) { a[0].kept = 3; kept[0] = 0; K[0] = KeptEnt{a[0].beg, a[0].end, a[0].w, a[0].is_alt}; F[0] = -1; }
  int n_kept = 1;
  __syncwarp();
  for (int i = 1; i < n_chn; ++i) {
    const int bi = a[i].beg, ei = a[i].end, wi = a[i].w, alt_i = a[i].is_alt;     // one address for the warp
    int large = 0;
    bool dropped = false;
    for (int base = 0; base < n_kept && !dropped; base += 32) {
      const int kk = base + lane;
      const int t = kk < n_kept ? flt_test(o, K[kk], bi, ei, wi, alt_i) : 0;
      const unsigned ov = __ballot_sync(full, t & 1), br = __ballot_sync(full, t & 2);
      unsigned upto = full;
      if (br) { const int p = __ffs(br) - 1; upto = p == 31 ? full : (2u << p) - 1u; dropped = true; }
      if (ov & upto) large = 1;
      if ((t & 1) && ((upto >> lane) & 1u) && F[kk] < 0) F[kk] = i;
    }
    if (!dropped) {
      if (lane == 0) { kept[n_kept] = i; K[n_kept] = KeptEnt{bi, ei, wi, alt_i}; F[n_kept] = -1; a[i].kept = large ? 2 : 3; }
      ++n_kept;
    }
    __syncwarp();                                            // lane 0's append before the next chain's tests
  }
  return n_kept;
}

#ifndef CHAIN_MIN_BLOCKS
#define CHAIN_MIN_BLOCKS 12      // 40 registers: 48 warps per SM hide more of the B-tree walk's load latency (5.2 -> 4.8 ms; 16 blocks / 32 registers spill too much: 5.6 ms)
#endif
__global__ void __launch_bounds__(128, CHAIN_MIN_BLOCKS) chain_kernel(const ChainArgs a)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_reads) return;
  const int len = (int)(a.read_off[r + 1] - a.read_off[r]);
  if (len >= a.opt.min_seed_len && needs_seed_rescoring(a.opt, len)) { atomicExch(a.err, 1); a.n_chains[r] = 0; a.n_seeds[r] = 0; return; }
  const int64_t i0 = a.intv_off[r], i1 = a.intv_off[r + 1];
  if (i1 > a.out_cap || a.pos_off[i1] > a.pos_cap) { a.n_chains[r] = 0; a.n_seeds[r] = 0; return; }   // left to the redo with larger buffers
  Scratch S = read_scratch(a, r);
  float fr;
  int n = chain_seeds(S, a.opt, a.ctg, len, a.intv + i0, i1 - i0, a.pos, a.pos_off + i0, &fr);
  a.frac_rep[r] = fr;
  if (n >= CHAIN_FLT_WARP_MIN) {                             // the overlap filter of this read is a warp's work: chain_heavy_kernel
    a.n_chains[r] = n; a.n_seeds[r] = 0;
    a.heavy[2 + atomicAdd(a.heavy, 1)] = (int)r;
    return;
  }
  n = chain_filter(S, a.opt, n);
  int ns = 0;
  for (int c = 0; c < n; ++c) ns += S.sorted[c].n;
  a.n_chains[r] = n; a.n_seeds[r] = ns;
}

// The rest of mem_chain_flt for the reads chain_kernel set aside (64 chains or more), one warp per read: the overlap filter by
// the warp, marking / compaction and the counts by lane 0.  Usually there is none and the warps leave at once.
__global__ void __launch_bounds__(128) chain_heavy_kernel(const ChainArgs a)
{
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(a.heavy + 1, 1);
    slot = __shfl_sync(full, slot, 0);
    if (slot >= a.heavy[0]) break;
    const int64_t r = a.heavy[2 + slot];
    Scratch S = read_scratch(a, r);
    const int nn = (int)a.n_chains[r];
    const int nk = flt_filter_warp(S.sorted, flt_kept_ents(S), flt_kept_first(S, nn), S.kept, nn, a.opt, lane);
    if (lane == 0) {
      const int n = flt_finish(S, a.opt, nn, nk);
      int ns = 0;
      for (int c = 0; c < n; ++c) ns += S.sorted[c].n;
      a.n_chains[r] = n; a.n_seeds[r] = ns;
    }
    __syncwarp();
  }
}

// second pass: flat chains/seeds in the layout stage 4 takes (chains of read r at chain_off[r], seeds contiguous per chain)
__global__ void __launch_bounds__(128) chain_emit_kernel(const ChainArgs a, const int64_t *chain_off, const int64_t *seed_off,
                                                         bwaflow_chain_t *chains, bwaflow_seed_t *seeds)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_reads) return;
  const Scratch S = read_scratch(a, r);
  const int n = (int)a.n_chains[r];
  int64_t so = seed_off[r];
  const float fr = a.frac_rep[r];
  for (int c = 0; c < n; ++c) {
    const Chain ch = S.sorted[c];
    bwaflow_chain_t f; f.pos = ch.pos; f.n_seeds = ch.n; f.rid = ch.rid; f.frac_rep = fr;
    f.flags = (uint32_t)ch.w | (uint32_t)ch.kept << 29 | (uint32_t)ch.is_alt << 31; f.seed_off = so;
    chains[chain_off[r] + c] = f;
    for (int j = ch.head; j >= 0; j = S.seeds[j].next) seeds[so++] = S.seeds[j].s;
  }
}

__global__ void fill_last_zero(int64_t *a, int64_t n) { if (threadIdx.x == 0 && blockIdx.x == 0) a[n] = 0; }

} // namespace bwaflow

using namespace bwaflow;

struct bwaflow_hot_batch {
  bwaflow_ctx_t *ctx;
  bwaflow_smem_batch *s;
  bwaflow_c2a_batch *c;
  cudaStream_t stream;
  cudaEvent_t ev[7];
  int64_t max_reads, max_bases;
  // chaining scratch
  int64_t scr_entries, scr_nodes;
  chain::SeedNode *d_seeds; chain::Chain *d_chains, *d_sorted; int *d_order, *d_kept; chain::Node *d_nodes;
  int64_t *d_nch, *d_nsd, *d_choff, *d_sdoff, *h_tot; float *d_frac; int *d_err, *h_err;
  void *d_tmp; size_t tmp_bytes;
  int *d_heavy;
  int32_t *d_ctg_alt;
  bwaflow_chain_opt_t copt;
  int64_t n_intv, n_pos, n_chains, n_seeds;
  float ms_smem, ms_seedpos, ms_chain, ms_c2a, ms_h2d;
  bool count_cells, new_input, front_queued;
  bwaflow::ChainArgs cargs;
  double blocks_smem, blocks_pos;
  bool pending;
};

static int hot_grow_scratch(bwaflow_hot_batch *b, int64_t entries, int64_t nodes)
{
  if (entries > b->scr_entries) {
    cudaFree(b->d_seeds); cudaFree(b->d_chains); cudaFree(b->d_sorted); cudaFree(b->d_order); cudaFree(b->d_kept);
    b->scr_entries = entries + entries / 4;
    BF_CUDA(cudaMalloc(&b->d_seeds, b->scr_entries * sizeof(chain::SeedNode)));
    BF_CUDA(cudaMalloc(&b->d_chains, b->scr_entries * sizeof(chain::Chain)));
    BF_CUDA(cudaMalloc(&b->d_sorted, b->scr_entries * sizeof(chain::Chain)));
    BF_CUDA(cudaMalloc(&b->d_order, b->scr_entries * sizeof(int)));
    BF_CUDA(cudaMalloc(&b->d_kept, b->scr_entries * sizeof(int)));
  }
  if (nodes > b->scr_nodes) {
    cudaFree(b->d_nodes);
    b->scr_nodes = nodes + nodes / 4;
    BF_CUDA(cudaMalloc(&b->d_nodes, b->scr_nodes * sizeof(chain::Node)));
  }
  return 0;
}

// (re)creates the stage-4 batch with room for n_chains/n_seeds and points it at the stage-2 batch's resident reads
static int hot_make_c2a(bwaflow_hot_batch *b, int64_t max_chains, int64_t max_seeds)
{
  if (b->c) {
    b->c->d_bases = nullptr; b->c->d_read_off = nullptr;           // aliases, not owned
    b->c->stream = b->c->own_stream;
    bwaflow_c2a_batch_destroy(b->c);
    b->c = nullptr;
  }
  int r = bwaflow_c2a_batch_create(b->ctx, b->max_reads, 16, max_chains, max_seeds, &b->c);
  if (r) return r;
  cudaFree(b->c->d_bases); cudaFree(b->c->d_read_off);
  b->c->d_bases = b->s->d_bases; b->c->d_read_off = b->s->d_read_off;
  b->c->own_stream = b->c->stream;
  b->c->stream = b->stream;                                         // everything runs in order on the stage-2 stream
  return 0;
}

// Front half, nothing but enqueues: stage 2 (seeding, compaction, seed positions) and chaining on the batch's stream, every
// kernel bounded by the batch's capacities and sized from device-side counts, then the read-back of the batch's status words.
static int hot_enqueue_front(bwaflow_hot_batch *b)
{
  bwaflow_smem_batch *s = b->s;
  cudaStream_t st = b->stream;
  const int64_t nr = s->n_reads;
  int r;
  BF_CUDA(cudaEventRecord(b->ev[1], st));
  s->want_rows = false; s->count_blocks = b->count_cells;   // rows are not needed downstream: text-anchored seeds carry positions
  if ((r = smem_launch(s))) return r;
  BF_CUDA(cudaEventRecord(b->ev[2], st));
  if ((r = smem_tail_device(s))) return r;
  BF_CUDA(cudaEventRecord(b->ev[3], st));
  // ---- chaining
  ChainArgs a;
  a.n_reads = nr; a.read_off = s->d_read_off; a.intv = s->d_out; a.intv_off = s->d_off; a.pos = s->d_pos; a.pos_off = s->d_poff;
  a.out_cap = s->out_cap; a.pos_cap = s->pos_cap;
  a.seeds = b->d_seeds; a.chains = b->d_chains; a.sorted = b->d_sorted; a.order = b->d_order; a.kept = b->d_kept; a.nodes = b->d_nodes;
  a.n_chains = b->d_nch; a.n_seeds = b->d_nsd; a.frac_rep = b->d_frac; a.err = b->d_err; a.heavy = b->d_heavy; a.opt = b->copt;
  a.ctg = chain::Ctg{b->ctx->ix.n_seqs, b->ctx->ix.ctg_offset, b->ctx->ix.ctg_len, b->d_ctg_alt, b->ctx->ix.l_pac};
  b->cargs = a;
  BF_CUDA(cudaMemsetAsync(b->d_err, 0, sizeof(int), st));
  BF_CUDA(cudaMemsetAsync(b->d_heavy, 0, 2 * sizeof(int), st));
  BF_CUDA(cudaMemsetAsync(b->d_choff, 0, (nr + 1) * 8, st));
  BF_CUDA(cudaMemsetAsync(b->d_sdoff, 0, (nr + 1) * 8, st));
  if (nr) {
    const unsigned grid = (unsigned)((nr + 127) / 128);
    chain_kernel<<<grid, 128, 0, st>>>(a);
    chain_heavy_kernel<<<b->ctx->n_sm * 4, 128, 0, st>>>(a);
    fill_last_zero<<<1, 1, 0, st>>>(b->d_nch, nr);
    fill_last_zero<<<1, 1, 0, st>>>(b->d_nsd, nr);
    size_t tmp = b->tmp_bytes;
    BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_nch, b->d_choff, (int)(nr + 1), st));
    tmp = b->tmp_bytes;
    BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_nsd, b->d_sdoff, (int)(nr + 1), st));
  }
  BF_CUDA(cudaEventRecord(b->ev[4], st));
  BF_CUDA(cudaMemcpyAsync(b->h_tot, b->d_choff + nr, 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(b->h_tot + 1, b->d_sdoff + nr, 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(b->h_err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((r = smem_status_d2h(s))) return r;
  b->front_queued = true;
  return 0;
}

// Back half: the one host wait of a batch (status words of the front half), a redo of the front half if a capacity was overrun
// (first batches only: capacities keep their high-water marks), then stage 4 on the resident chains.
static int hot_finish(bwaflow_hot_batch *b)
{
  bwaflow_smem_batch *s = b->s;
  cudaStream_t st = b->stream;
  const int64_t nr = s->n_reads;
  int r;
  for (int round = 0;; ++round) {
    BF_CUDA(cudaStreamSynchronize(st));
    bool redo = false; int64_t n_pos = 0;
    if ((r = smem_check_capacity(s, &redo, &n_pos))) return r;
    if (!redo) {
      b->n_intv = s->n_intv; b->n_pos = n_pos;
      if (n_pos + nr + 8 > b->scr_entries) { if ((r = hot_grow_scratch(b, n_pos + nr + 8, n_pos / 4 + 4 * nr + 8))) return r; redo = true; }
    } else if ((r = hot_grow_scratch(b, s->pos_cap + nr + 8, s->pos_cap / 4 + 4 * nr + 8))) return r;
    if (!redo) break;
    BF_REQUIRE(round < 8, -5, "hot path: capacities did not settle");
    if ((r = hot_enqueue_front(b))) return r;
  }
  b->front_queued = false;
  b->blocks_smem = (double)*(unsigned long long *)(s->h_ctl + 4);
  BF_REQUIRE(!*b->h_err, -34, "read too long for chaining: mem_flt_chained_seeds (bwa/bwamem.c:607) would rescore its seeds");
  const int64_t n_chains = b->h_tot[0], n_seeds = b->h_tot[1];
  b->n_chains = n_chains; b->n_seeds = n_seeds;
  if (!b->c || n_chains > b->c->max_chains || n_seeds > b->c->max_seeds)
    if ((r = hot_make_c2a(b, n_chains + n_chains / 4 + 1024, n_seeds + n_seeds / 4 + 1024))) return r;
  bwaflow_c2a_batch *c = b->c;
  BF_CUDA(cudaEventRecord(b->ev[6], st));
  if (nr) {
    const unsigned grid = (unsigned)((nr + 127) / 128);
    chain_emit_kernel<<<grid, 128, 0, st>>>(b->cargs, b->d_choff, b->d_sdoff, c->d_chains, c->d_seeds);
    BF_CUDA(cudaMemcpyAsync(c->d_chain_off, b->d_choff, (nr + 1) * 8, cudaMemcpyDeviceToDevice, st));
  }
  // ---- stage 4 on the resident chains
  c->n_reads = nr; c->n_chains = n_chains; c->n_seeds = n_seeds;
  BF_CUDA(cudaEventRecord(c->tm.ev[0], st));
  BF_CUDA(cudaEventRecord(c->tm.ev[1], st));
  c->count_cells = b->count_cells;
  if (b->new_input) { c->cells_known = 0.0; b->new_input = false; }
  if ((r = c2a_launch(c))) return r;
  if ((r = c2a_finish_queue(c))) return r;
  BF_CUDA(cudaEventRecord(b->ev[5], st));
  return 0;
}

extern "C" {

int bwaflow_hot_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, const bwaflow_chain_opt_t *copt,
                             const int32_t *ctg_is_alt, bwaflow_hot_batch_t **out)
{
  BF_REQUIRE(ctx && out && ctx->has_index, -22, "bwaflow_hot_batch_create: bad argument");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_hot_batch *b = (bwaflow_hot_batch *)calloc(1, sizeof(*b));
  b->ctx = ctx; b->max_reads = max_reads; b->max_bases = max_bases;
  if (copt) b->copt = *copt; else bwaflow_chain_opt_default(&b->copt);
  // the fields mem_opt_t holds once (bwa/bwamem.h:26-58) come from the context, so the two option structs cannot disagree
  b->copt.min_seed_len = ctx->par.opt.min_seed_len; b->copt.max_occ = ctx->par.opt.max_occ; b->copt.w = ctx->par.opt.w;
  int r = bwaflow_smem_batch_create(ctx, max_reads, max_bases, &b->s);
  if (r) return r;
  b->stream = b->s->stream;
  for (int i = 0; i < 7; ++i) BF_CUDA(cudaEventCreate(&b->ev[i]));
  BF_CUDA(cudaMalloc(&b->d_nch, (max_reads + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_nsd, (max_reads + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_choff, (max_reads + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_sdoff, (max_reads + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_frac, (max_reads + 2) * 4));
  BF_CUDA(cudaMalloc(&b->d_err, 16));
  BF_CUDA(cudaMalloc(&b->d_heavy, (max_reads + 4) * sizeof(int)));
  BF_CUDA(cudaMallocHost(&b->h_tot, 32));
  BF_CUDA(cudaMallocHost(&b->h_err, 16));
  BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b->tmp_bytes, b->d_nch, b->d_choff, (int)(max_reads + 1), b->stream));
  b->tmp_bytes += 256;
  BF_CUDA(cudaMalloc(&b->d_tmp, b->tmp_bytes));
  if (ctg_is_alt) {
    BF_CUDA(cudaMalloc(&b->d_ctg_alt, ctx->ix.n_seqs * 4));
    BF_CUDA(cudaMemcpy(b->d_ctg_alt, ctg_is_alt, ctx->ix.n_seqs * 4, cudaMemcpyHostToDevice));
  }
  if ((r = hot_grow_scratch(b, b->s->pos_cap + max_reads + 8, b->s->pos_cap / 4 + 4 * max_reads + 8))) return r;
  if ((r = hot_make_c2a(b, 4 * max_reads + 1024, 12 * max_reads + 1024))) return r;
  *out = b;
  return 0;
}

void bwaflow_hot_batch_destroy(bwaflow_hot_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  if (b->c) { b->c->d_bases = nullptr; b->c->d_read_off = nullptr; b->c->stream = b->c->own_stream; bwaflow_c2a_batch_destroy(b->c); }
  bwaflow_smem_batch_destroy(b->s);
  cudaFree(b->d_seeds); cudaFree(b->d_chains); cudaFree(b->d_sorted); cudaFree(b->d_order); cudaFree(b->d_kept); cudaFree(b->d_nodes);
  cudaFree(b->d_nch); cudaFree(b->d_nsd); cudaFree(b->d_choff); cudaFree(b->d_sdoff); cudaFree(b->d_frac); cudaFree(b->d_err); cudaFree(b->d_heavy);
  cudaFreeHost(b->h_tot); cudaFreeHost(b->h_err); cudaFree(b->d_tmp); cudaFree(b->d_ctg_alt);
  for (int i = 0; i < 7; ++i) cudaEventDestroy(b->ev[i]);
  free(b);
}

int bwaflow_hot_submit(bwaflow_hot_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off)
{
  BF_REQUIRE(b && (n_reads == 0 || (bases && read_off)), -22, "bwaflow_hot_submit: null argument");
  bwaflow_smem_batch *s = b->s;
  const int64_t n_bases = n_reads ? read_off[n_reads] : 0;
  BF_REQUIRE(n_reads >= 0 && n_reads <= s->max_reads && n_bases <= s->max_bases, -7,
             "bwaflow_hot_submit: batch capacity exceeded (%lld reads, %lld bases)", (long long)n_reads, (long long)n_bases);
  BF_CUDA(cudaSetDevice(b->ctx->device));
  int64_t max_len = 0;
  for (int64_t i = 0; i < n_reads; ++i) { const int64_t l = read_off[i + 1] - read_off[i]; max_len = l > max_len ? l : max_len; }
  BF_REQUIRE(max_len < 1024, -34, "read of %lld bases: the device path supports reads shorter than 1024", (long long)max_len);
  s->n_reads = n_reads; s->max_len = max_len;
  BF_CUDA(cudaEventRecord(b->ev[0], b->stream));
  if (n_reads) {
    BF_CUDA(cudaMemcpyAsync(s->d_bases, bases, n_bases, cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(s->d_read_off, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, b->stream));
  }
  b->count_cells = false; b->new_input = true;
  b->pending = true;
  return hot_enqueue_front(b);
}

static int hot_rerun(bwaflow_hot_batch_t *b, bool counted)
{
  BF_REQUIRE(b, -22, "bwaflow_hot_rerun: null batch");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  BF_CUDA(cudaEventRecord(b->ev[0], b->stream));
  b->count_cells = counted;
  b->pending = true;
  return hot_enqueue_front(b);
}

int bwaflow_hot_second_round(bwaflow_hot_batch_t *b, int64_t *n_reads, int64_t *n_jobs)
{
  BF_REQUIRE(b && b->c, -22, "bwaflow_hot_second_round: call after collect");
  return bwaflow_c2a_second_round(b->c, n_reads, n_jobs);
}

int bwaflow_hot_exact_reads(bwaflow_hot_batch_t *b, int64_t *n_exact)
{
  BF_REQUIRE(b && b->c && n_exact, -22, "bwaflow_hot_exact_reads: call after collect");
  return bwaflow_c2a_exact_reads(b->c, n_exact);
}

int bwaflow_hot_rerun(bwaflow_hot_batch_t *b) { return hot_rerun(b, false); }
int bwaflow_hot_rerun_counted(bwaflow_hot_batch_t *b) { return hot_rerun(b, true); }

int bwaflow_hot_collect(bwaflow_hot_batch_t *b, const bwaflow_alnreg_t **regs, const int64_t **reg_off, int64_t *n_regs,
                        bwaflow_hot_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_hot_collect: nothing submitted");
  bwaflow_timing_t ct;
  BF_CUDA(cudaSetDevice(b->ctx->device));
  int r = hot_finish(b);
  if (r) { b->pending = false; return r; }
  r = bwaflow_c2a_collect(b->c, regs, reg_off, n_regs, &ct);
  b->pending = false;
  if (r) return r;
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->ev[0], b->ev[1]);
    cudaEventElapsedTime(&t->smem_ms, b->ev[1], b->ev[2]);
    cudaEventElapsedTime(&t->seedpos_ms, b->ev[2], b->ev[3]);
    cudaEventElapsedTime(&t->chain_ms, b->ev[3], b->ev[4]);
    cudaEventElapsedTime(&t->c2a_ms, b->ev[6], b->ev[5]);     // from the host's go-ahead (the batch's one wait) to the last stage-4 kernel
    t->d2h_ms = ct.d2h_ms;
    cudaEventElapsedTime(&t->total_ms, b->ev[0], b->c->tm.ev[3]);
    t->smem_blocks = b->blocks_smem;
    t->seedpos_blocks = (double)(unsigned long long)b->h_tot[2];
    t->cells = ct.work_units;
    t->n_intv = b->n_intv; t->n_pos = b->n_pos; t->n_chains = b->n_chains; t->n_seeds = b->n_seeds;
    t->n_launches = 6 + 4 + 8 + ct.n_launches;     // seeding 6, seed positions 4, chaining 8 (scans = 2 kernels each), stage 4
  }
  return 0;
}

} // extern "C"
