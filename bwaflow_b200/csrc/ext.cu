// Batched ksw_extend2 jobs (bwa/ksw.c:380-479), the device counterpart of the FPGA SWTask batch
// (src/fpga/SWTask.h:9-50): submit = async H2D + device-side length binning + persistent kernels
// + async D2H; collect = wait.
//
// Device pipeline per submit:
//   1. ext_key_kernel      one key per job = qlen class (sizes the shared-memory row state) | work estimate
//   2. cub radix sort      jobs ordered class-major, longest first (LPT order for the work queues)
//   3. ext_ranges_kernel   per-class [begin,end) + work-queue heads
//   4. ext_kernel<...>     one persistent launch per qlen class; every lane owns one job at a time and
//                          pulls the next from its class queue when enough lanes of the warp are idle
#include <cub/device/device_radix_sort.cuh>
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "ksw_lane.cuh"
#include "ksw_pair.cuh"
#include "ksw_cols.cuh"
#include "batches.cuh"

namespace bwaflow {

constexpr int EXT_NCLASS = 8;
constexpr int EXT_NQ = EXT_NCLASS + 1;
constexpr int CTL_HIST = 0, CTL_QUEUE = EXT_NCLASS, CTL_ERR = CTL_QUEUE + 2 * EXT_NQ, CTL_CELLS = CTL_ERR + 2, CTL_INTS = CTL_CELLS + 2;
constexpr int EXT_PAIR_CLASSES = 6;      // classes 0..5 (qlen <= 255) can run on the packed engines; which of them do is set per batch (ext_set_pair_classes, BWAFLOW_EXT_PAIR_CLASSES)
// qlen <= qmax runs in class c with `threads` lanes per block; row state = threads*(qmax+1)*4 bytes
__constant__ int c_ext_qmax[EXT_NCLASS] = {31, 63, 95, 127, 159, 255, 511, 1023};
static const int h_ext_qmax[EXT_NCLASS] = {31, 63, 95, 127, 159, 255, 511, 1023};
static const int h_ext_threads[EXT_NCLASS] = {64, 64, 64, 64, 64, 64, 64, 32};

struct ExtArgs {
  const bwaflow_ext_job_t *jobs;
  const uint8_t *qpool, *tpool;
  const uint32_t *order;
  int *queue;                 // [EXT_NQ] heads, [EXT_NQ] ends; queue EXT_NCLASS = jobs deferred to the int32 engine
  uint32_t *slow;             // job ids of that queue (appended by the packed kernel)
  bwaflow_ext_result_t *out;
  int *err;
  unsigned long long *cells;
  int cls, qmax;
  int refill;                 // refill idle lanes once at least this many are idle
  int pair_sync;              // packed kernel: refill the two slots of a lane together
  DevScore sc;
  PairConsts pc;
};

__global__ void ext_key_kernel(const bwaflow_ext_job_t *jobs, int64_t n, uint32_t *keys, uint32_t *vals, int *hist, int *err, int order_rows, int zdrop)
{
  __shared__ int sh[EXT_NCLASS];
  if (threadIdx.x < EXT_NCLASS) sh[threadIdx.x] = 0;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int qlen = jobs[i].qlen, tlen = jobs[i].tlen;
    int c = 0;
    while (c < EXT_NCLASS - 1 && qlen > c_ext_qmax[c]) ++c;
    if (qlen > c_ext_qmax[EXT_NCLASS - 1] || qlen < 0 || tlen < 0) { atomicExch(err, 1); }
    if (jobs[i].zdrop != zdrop) atomicExch(err, 4);        // mem_chain2aln passes opt->zdrop to every call; so does the context
    // queue order inside a class.  order_rows == 0 (raw job batches): widest rows first (qlen), then most rows (tlen): the lanes of
    // a warp then walk rows of the same width and finish at about the same time.  order_rows == 1 (the first-seed waves of
    // stage 4, a mix of strong seeds and weak chains): by estimated lifetime first -- an extension that starts from a small
    // h0 dies within about 2*h0 rows whatever its target length, every mismatch costing more than a match gains -- then by
    // width, so that a warp's lanes free up together (stage 4: 11.3 -> 8.9 ms; raw batches of long-lived jobs lose 5% with it).
    if (order_rows) {
      const int h0 = jobs[i].h0;
      int rows = 2 * (h0 > 0 ? h0 : 0) + 8;
      rows = rows < tlen ? rows : tlen;
      keys[i] = (uint32_t)c << 24 | (uint32_t)(rows > 0x3FFF ? 0x3FFF : rows) << 10 | (uint32_t)(qlen & 0x3FF);
    } else
      keys[i] = (uint32_t)c << 24 | (uint32_t)(qlen & 0x3FF) << 14 | (uint32_t)(tlen > 0x3FFF ? 0x3FFF : tlen);
    vals[i] = (uint32_t)i;
    atomicAdd(&sh[c], 1);
  }
  __syncthreads();
  if (threadIdx.x < EXT_NCLASS && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void ext_ranges_kernel(const int *hist, int *queue)
{ // keys are sorted descending, so the highest class comes first
  int pos = 0;
  for (int c = EXT_NCLASS - 1; c >= 0; --c) {
    queue[c] = pos;
    pos += hist[c];
    queue[EXT_NQ + c] = pos;
  }
  queue[EXT_NCLASS] = 0; queue[EXT_NQ + EXT_NCLASS] = 0;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) ext_kernel(const ExtArgs a)
{
  extern __shared__ uint32_t s_rows[];
  const int stride = THREADS;
  uint32_t *col = s_rows + threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int *head = a.queue + a.cls;
  const int q_end = a.queue[EXT_NQ + a.cls];

  KswLane k;
  bool active = false, drained = false;
  int job = -1;
  const uint8_t *tp = nullptr;

  for (;;) {
    const unsigned idle = __ballot_sync(full, !active);
    if (idle == full || (!drained && __popc(idle) >= a.refill)) {
      const int n_idle = __popc(idle);
      int base = 0;
      if (!drained) {
        if (lane == 0) base = atomicAdd(head, n_idle);
        base = __shfl_sync(full, base, 0);
        if (base + n_idle >= q_end) drained = true;
      } else base = q_end;
      if (!active) {
        const int slot = base + __popc(idle & ((1u << lane) - 1));
        if (slot < q_end) {
          job = (int)a.order[slot];
          const bwaflow_ext_job_t jb = a.jobs[job];
          const uint8_t *qp = a.qpool + jb.q_off;
          tp = a.tpool + jb.t_off;
          const bool ok = jb.qlen >= 0 && jb.qlen <= a.qmax &&
                          k.begin(col, stride, a.sc, jb.qlen, jb.tlen, jb.w, jb.h0, jb.end_bonus,
                                  [&](int j) { return (uint32_t)qp[j]; });
          if (!ok) {
            atomicExch(a.err, 2);
            a.out[job] = bwaflow_ext_result_t{0, 0, 0, 0, 0, 0};
          } else if (jb.tlen <= 0) {
            a.out[job] = bwaflow_ext_result_t{k.best, 0, 0, 0, -1, 0};
          } else active = true;
        }
      }
      if (__ballot_sync(full, active) == 0) break;
    }
    if (active) {
      if (k.row(col, stride, a.sc, (int)tp[k.i])) {
        a.out[job] = bwaflow_ext_result_t{k.best, k.best_j + 1, k.best_i + 1, k.end_i + 1, k.gscore, k.max_off};
        atomicAdd(a.cells, (unsigned long long)k.cells);
        active = false;
      }
    }
  }
}

// Packed engine: every lane owns two slots (ksw_pair.cuh); one warp per block, row state of both slots interleaved in
// shared memory.  Jobs the packed form cannot represent are appended to the deferred queue for ext_kernel.
// EXACT = also count the DP cells the reference executes (the GCUPS work unit); results are identical either way.
template <bool EXACT>
__global__ void __launch_bounds__(32) ext_pair_kernel(const ExtArgs a)
{
  extern __shared__ uint2 s_pair[];
  const int stride = 32;
  const int lane = threadIdx.x;
  const int ncol = a.qmax + 1, ngrp = (a.qmax >> 3) + 1;
  uint2 *st = s_pair + lane;
  uint32_t *qa = (uint32_t *)(s_pair + (size_t)ncol * stride) + lane;
  for (int j = 0; j < ncol; ++j) st[(size_t)j * stride] = make_uint2(a.pc.B2, a.pc.B2);
  for (int g = 0; g < ngrp; ++g) qa[(size_t)g * stride] = 0xFFFFFFFFu;
  const unsigned full = 0xffffffffu;
  int *head = a.queue + a.cls;
  const int q_end = a.queue[EXT_NQ + a.cls];

  KswSlot k[2];
  bool act[2] = {false, false};
  int job[2] = {-1, -1};
  const uint8_t *tp[2] = {nullptr, nullptr};
  int tcur[2] = {0, 0};
  bool drained = false;

  for (;;) {
    // Refill.  pair_sync: a lane takes two queue-adjacent jobs (same qlen class, nearly the same qlen and tlen after the
    // sort) when BOTH its slots are idle, so the two jobs of a lane walk the same band windows row for row; otherwise
    // every idle slot is refilled on its own.
    unsigned idle0 = __ballot_sync(full, !act[0]), idle1 = __ballot_sync(full, !act[1]);
    if (a.pair_sync) idle0 = idle1 = idle0 & idle1;
    const int n_idle = __popc(idle0) + __popc(idle1);
    if (__popc(__ballot_sync(full, !act[0] && !act[1])) == 32 || (!drained && n_idle >= a.refill)) {
      int base = q_end;
      if (!drained) {
        if (lane == 0) base = atomicAdd(head, n_idle);
        base = __shfl_sync(full, base, 0);
        if (base + n_idle >= q_end) drained = true;
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (act[s] || !((s ? idle1 : idle0) >> lane & 1)) continue;
        const int below = __popc(idle0 & ((1u << lane) - 1));
        const int slot = base + (a.pair_sync ? 2 * below + s : s ? __popc(idle0) + __popc(idle1 & ((1u << lane) - 1)) : below);
        if (slot >= q_end) continue;
        job[s] = (int)a.order[slot];
        const bwaflow_ext_job_t jb = a.jobs[job[s]];
        const uint8_t *qp = a.qpool + jb.q_off;
        tp[s] = a.tpool + jb.t_off;
        int r = PAIR_BAD;
        if (jb.qlen >= 0 && jb.qlen <= a.qmax)
          r = pair_begin(k[s], s, st, qa, stride, a.sc, a.pc, jb.qlen, jb.tlen, jb.w, jb.h0, jb.end_bonus,
                         [&](int j) { return (uint32_t)qp[j]; });
        if (r == PAIR_BAD) {
          atomicExch(a.err, 2);
          a.out[job[s]] = bwaflow_ext_result_t{0, 0, 0, 0, 0, 0};
        } else if (jb.tlen <= 0) {
          a.out[job[s]] = bwaflow_ext_result_t{k[s].best, 0, 0, 0, -1, 0};
        } else if (r == PAIR_SLOW) {
          a.slow[atomicAdd(a.queue + EXT_NQ + EXT_NCLASS, 1)] = (uint32_t)job[s];
        } else {
          act[s] = true;
          tcur[s] = tp[s][0];
        }
        if (!act[s]) pair_retire(s, qa, stride, jb.qlen <= a.qmax ? jb.qlen : a.qmax);
      }
      if (__ballot_sync(full, act[0] || act[1]) == 0) break;
    }
    // an ambiguous target base sends the job to the int32 engine
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (act[s] && tcur[s] > 3) {
        a.slow[atomicAdd(a.queue + EXT_NQ + EXT_NCLASS, 1)] = (uint32_t)job[s];
        act[s] = false;
        pair_retire(s, qa, stride, k[s].qlen);
      }
    {
      bool d0, d1;
      const int t0 = act[0] ? tcur[0] : 0, t1 = act[1] ? tcur[1] : 0;
      // next rows' target bases are fetched before the row so that the load overlaps it
      if (act[0] && k[0].i + 1 < k[0].tlen) tcur[0] = tp[0][k[0].i + 1];
      if (act[1] && k[1].i + 1 < k[1].tlen) tcur[1] = tp[1][k[1].i + 1];
      pair_step<EXACT>(st, qa, stride, a.sc, a.pc, k[0], k[1], act[0], act[1], t0, t1, d0, d1);
#pragma unroll
      for (int s = 0; s < 2; ++s)
        if (s ? d1 : d0) {
          a.out[job[s]] = bwaflow_ext_result_t{k[s].best, k[s].best_j + 1, k[s].best_i + 1, k[s].end_i + 1, k[s].gscore, k[s].max_off};
          if (EXACT) atomicAdd(a.cells, (unsigned long long)k[s].cells);
          act[s] = false;
          pair_retire(s, qa, stride, k[s].qlen);
        }
    }
  }
}

// Two-columns-per-word engine (ksw_cols.cuh): every lane owns ONE job and walks its own window; one warp per block, the row
// state of its 32 jobs interleaved in shared memory.  Jobs the packed form cannot represent go to the deferred queue.
__global__ void __launch_bounds__(32) ext_cols_kernel(const ExtArgs a)
{
  extern __shared__ uint2 s_cols[];
  const int lane = threadIdx.x;
  const int npmax = (a.qmax >> 1) + 1;
  uint2 *st = s_cols + lane;
  uint8_t *qb = (uint8_t *)(s_cols + (size_t)npmax * 32) + lane;
  const unsigned full = 0xffffffffu;
  int *head = a.queue + a.cls;
  const int q_end = a.queue[EXT_NQ + a.cls];

  __shared__ uint32_t s_stage[16];
  ColsK ck;
  ck.load(a.pc, s_stage);
  ColJob k;
  bool active = false, drained = false;
  int job = -1, tcur = 0;
  const uint8_t *tp = nullptr;

  for (;;) {
    const unsigned idle = __ballot_sync(full, !active);
    if (idle == full || (!drained && __popc(idle) >= a.refill)) {
      const int n_idle = __popc(idle);
      int base = q_end;
      if (!drained) {
        if (lane == 0) base = atomicAdd(head, n_idle);
        base = __shfl_sync(full, base, 0);
        if (base + n_idle >= q_end) drained = true;
      }
      if (!active) {
        const int slot = base + __popc(idle & ((1u << lane) - 1));
        if (slot < q_end) {
          job = (int)a.order[slot];
          const bwaflow_ext_job_t jb = a.jobs[job];
          const uint8_t *qp = a.qpool + jb.q_off;
          tp = a.tpool + jb.t_off;
          int r = PAIR_BAD;
          if (jb.qlen >= 0 && jb.qlen <= a.qmax)
            r = cols_begin(k, st, qb, a.sc, a.pc, jb.qlen, jb.tlen, jb.w, jb.h0, jb.end_bonus, [&](int j) { return (uint32_t)qp[j]; });
          if (r == PAIR_BAD) {
            atomicExch(a.err, 2);
            a.out[job] = bwaflow_ext_result_t{0, 0, 0, 0, 0, 0};
          } else if (jb.tlen <= 0) {
            a.out[job] = bwaflow_ext_result_t{k.best, 0, 0, 0, -1, 0};
          } else if (r == PAIR_SLOW) {
            a.slow[atomicAdd(a.queue + EXT_NQ + EXT_NCLASS, 1)] = (uint32_t)job;
          } else {
            active = true;
            tcur = tp[0];
          }
        }
      }
      if (__ballot_sync(full, active) == 0) break;
    }
    if (active && tcur > 3) {                              // an ambiguous target base sends the job to the int32 engine
      a.slow[atomicAdd(a.queue + EXT_NQ + EXT_NCLASS, 1)] = (uint32_t)job;
      active = false;
    }
    const int t = active ? tcur : 0;
    if (active && k.i + 1 < k.tlen) tcur = tp[k.i + 1];    // fetched before the row so that the load overlaps it
    if (cols_row(k, st, qb, a.sc, ck, active, t)) {
      a.out[job] = bwaflow_ext_result_t{k.best, k.best_j + 1, k.best_i + 1, k.end_i + 1, k.gscore, k.max_off};
      active = false;
    }
  }
}

} // namespace bwaflow

using namespace bwaflow;

struct bwaflow_ext_batch {
  bwaflow_ctx_t *ctx;
  cudaStream_t stream;
  Timer tm;
  int64_t max_jobs, max_q, max_t, n_jobs;
  bwaflow_ext_job_t *d_jobs;
  uint8_t *d_q, *d_t;
  uint32_t *d_keys[2], *d_vals[2];
  void *d_tmp; size_t tmp_bytes;
  int *d_ctl;                       // hist[8] | queue[16] | err | pad | cells (u64)
  bwaflow_ext_result_t *d_out, *h_out;
  int *h_err;                       // err | pad | cells (u64)
  int blocks[EXT_NCLASS], blocks_pair[EXT_NCLASS], blocks_cols[EXT_NCLASS];
  bool use_cols;               // packed classes run on the two-columns-per-word engine (the two-jobs-per-lane engine keeps the counting variant)
  int refill_cols;
  int refill, refill_pair, pair_sync, pair_classes, order_rows;
  bool use_pair, multi_stream;
  double cells_known;
  bool count_cells;            // this launch runs the cell-counting (EXACT) variant of the packed kernel
  PairConsts pc;
  uint32_t *d_slow;
  cudaStream_t cstream[EXT_NCLASS];
  cudaEvent_t ev_fork, ev_join[EXT_NCLASS];
  bool pending;
  double cells_hint;
};

static size_t pair_smem(int qmax) { return (size_t)32 * ((qmax + 1) * sizeof(uint2) + ((qmax >> 3) + 1) * sizeof(uint32_t)); }

static int ext_launch(bwaflow_ext_batch *b)
{
  const int64_t n = b->n_jobs;
  int *hist = b->d_ctl + CTL_HIST, *queue = b->d_ctl + CTL_QUEUE, *err = b->d_ctl + CTL_ERR;
  BF_CUDA(cudaMemsetAsync(b->d_ctl, 0, CTL_INTS * sizeof(int), b->stream));
  if (n == 0) return 0;
  ext_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->d_jobs, n, b->d_keys[0], b->d_vals[0], hist, err, b->order_rows, b->ctx->par.sc.zdrop);
  cub::DoubleBuffer<uint32_t> keys(b->d_keys[0], b->d_keys[1]), vals(b->d_vals[0], b->d_vals[1]);
  size_t tmp = b->tmp_bytes;
  BF_CUDA(cub::DeviceRadixSort::SortPairsDescending(b->d_tmp, tmp, keys, vals, (int)n, 0, 27, b->stream));
  ext_ranges_kernel<<<1, 1, 0, b->stream>>>(hist, queue);
  ExtArgs a;
  a.jobs = b->d_jobs; a.qpool = b->d_q; a.tpool = b->d_t; a.order = vals.Current(); a.queue = queue; a.slow = b->d_slow;
  a.out = b->d_out; a.err = err; a.cells = (unsigned long long *)(b->d_ctl + CTL_CELLS); a.refill = b->refill; a.sc = b->ctx->par.sc;
  a.pc = b->pc;
  // one persistent launch per class, each on its own stream: a class's blocks take the SMs as the previous class drains,
  // so the tails (few long jobs left) overlap with the next class instead of idling the GPU
  BF_CUDA(cudaEventRecord(b->ev_fork, b->stream));
  for (int c = EXT_NCLASS - 1; c >= 0; --c) {
    a.cls = c; a.qmax = h_ext_qmax[c];
    cudaStream_t cs = b->multi_stream ? b->cstream[c] : b->stream;
    if (b->multi_stream) BF_CUDA(cudaStreamWaitEvent(cs, b->ev_fork, 0));
    if (b->use_pair && c < b->pair_classes) {
      a.refill = b->refill_pair; a.pair_sync = b->pair_sync;
      if (b->count_cells) ext_pair_kernel<true><<<b->blocks_pair[c], 32, pair_smem(h_ext_qmax[c]), cs>>>(a);
      else if (b->use_cols) { a.refill = b->refill_cols; ext_cols_kernel<<<b->blocks_cols[c], 32, cols_smem(h_ext_qmax[c]), cs>>>(a); }
      else ext_pair_kernel<false><<<b->blocks_pair[c], 32, pair_smem(h_ext_qmax[c]), cs>>>(a);
    } else {
      a.refill = b->refill;
      const size_t smem = (size_t)h_ext_threads[c] * (h_ext_qmax[c] + 1) * sizeof(uint32_t);
      if (h_ext_threads[c] == 64) ext_kernel<64><<<b->blocks[c], 64, smem, cs>>>(a);
      else ext_kernel<32><<<b->blocks[c], 32, smem, cs>>>(a);
    }
    if (b->multi_stream) {
      BF_CUDA(cudaEventRecord(b->ev_join[c], cs));
      BF_CUDA(cudaStreamWaitEvent(b->stream, b->ev_join[c], 0));
    }
  }
  if (b->use_pair) {                                   // jobs the packed engine deferred (ambiguous bases, 16-bit range)
    const int c = EXT_NCLASS - 1;
    a.cls = EXT_NCLASS; a.qmax = h_ext_qmax[c]; a.order = b->d_slow; a.refill = b->refill;
    ext_kernel<32><<<b->blocks[c], 32, (size_t)32 * (h_ext_qmax[c] + 1) * sizeof(uint32_t), b->stream>>>(a);
  }
  BF_CUDA(cudaGetLastError());
  return 0;
}

// ---- device-side use by another stage (chain2aln.cu): jobs and byte pools are written by kernels, results stay in HBM
namespace bwaflow {
ExtDeviceView ext_device_view(bwaflow_ext_batch *b)
{
  ExtDeviceView v;
  v.jobs = b->d_jobs; v.q = b->d_q; v.t = b->d_t; v.out = b->d_out; v.max_jobs = b->max_jobs; v.max_q = b->max_q; v.max_t = b->max_t;
  return v;
}
void ext_set_pair_classes(bwaflow_ext_batch *b, int n)
{
  if (!getenv("BWAFLOW_EXT_PAIR_CLASSES")) b->pair_classes = n < EXT_PAIR_CLASSES ? n : EXT_PAIR_CLASSES;
}
void ext_set_order_rows(bwaflow_ext_batch *b, int on)
{
  if (!getenv("BWAFLOW_EXT_ORDER_ROWS")) b->order_rows = on;
}
int ext_launch_device(bwaflow_ext_batch *b, int64_t n_jobs, cudaStream_t st)
{
  BF_REQUIRE(n_jobs >= 0 && n_jobs <= b->max_jobs, -7, "ext_launch_device: %lld jobs, capacity %lld", (long long)n_jobs, (long long)b->max_jobs);
  cudaStream_t own = b->stream;
  b->stream = st; b->n_jobs = n_jobs; b->count_cells = false;
  const int r = ext_launch(b);
  b->stream = own;
  return r;
}
} // namespace bwaflow

extern "C" {

int bwaflow_ext_batch_create(bwaflow_ctx_t *ctx, int64_t max_jobs, int64_t max_qbytes, int64_t max_tbytes,
                             bwaflow_ext_batch_t **out)
{
  BF_REQUIRE(ctx && out, -22, "bwaflow_ext_batch_create: null argument");
  BF_REQUIRE(max_jobs > 0 && max_jobs < (1LL << 31), -22, "max_jobs out of range");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_ext_batch *b = (bwaflow_ext_batch *)calloc(1, sizeof(*b));
  b->ctx = ctx; b->max_jobs = max_jobs; b->max_q = max_qbytes; b->max_t = max_tbytes;
  b->refill = 8;
  if (const char *s = getenv("BWAFLOW_EXT_REFILL")) b->refill = atoi(s) > 0 ? atoi(s) : 8;
  b->refill_pair = 16;
  if (const char *s = getenv("BWAFLOW_EXT_REFILL_PAIR")) b->refill_pair = atoi(s) > 0 ? atoi(s) : 16;
  b->pair_sync = 0;
  b->order_rows = 0;
  if (const char *s = getenv("BWAFLOW_EXT_ORDER_ROWS")) b->order_rows = atoi(s) != 0;
  b->pair_classes = 5;                            // raw job batches: packed up to qlen 159 (measured crossover on long-lived jobs)
  if (const char *s = getenv("BWAFLOW_EXT_PAIR_CLASSES")) b->pair_classes = atoi(s) < EXT_PAIR_CLASSES ? atoi(s) : EXT_PAIR_CLASSES;
  if (const char *s = getenv("BWAFLOW_EXT_PAIR_SYNC")) b->pair_sync = atoi(s) != 0;
  b->multi_stream = true;
  if (const char *s = getenv("BWAFLOW_EXT_STREAMS")) b->multi_stream = atoi(s) != 1;
  fill_pair_consts(ctx->par.opt, ctx->par.sc, b->pc);
  b->use_pair = b->pc.ok != 0;
  b->use_cols = true; b->refill_cols = 8;
  if (const char *s = getenv("BWAFLOW_EXT_ENGINE")) { if (!strcmp(s, "int32")) b->use_pair = false; if (!strcmp(s, "pair")) b->use_cols = false; }
  if (const char *s = getenv("BWAFLOW_EXT_REFILL_COLS")) b->refill_cols = atoi(s) > 0 ? atoi(s) : 8;
  BF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (b->tm.init()) return -5;
  BF_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
  for (int c = 0; c < EXT_NCLASS; ++c) {
    BF_CUDA(cudaStreamCreateWithFlags(&b->cstream[c], cudaStreamNonBlocking));
    BF_CUDA(cudaEventCreateWithFlags(&b->ev_join[c], cudaEventDisableTiming));
  }
  BF_CUDA(cudaMalloc(&b->d_jobs, max_jobs * sizeof(bwaflow_ext_job_t)));
  BF_CUDA(cudaMalloc(&b->d_q, max_qbytes + 16));
  BF_CUDA(cudaMalloc(&b->d_t, max_tbytes + 16));
  for (int i = 0; i < 2; ++i) {
    BF_CUDA(cudaMalloc(&b->d_keys[i], max_jobs * 4));
    BF_CUDA(cudaMalloc(&b->d_vals[i], max_jobs * 4));
  }
  cub::DoubleBuffer<uint32_t> keys(b->d_keys[0], b->d_keys[1]), vals(b->d_vals[0], b->d_vals[1]);
  BF_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, b->tmp_bytes, keys, vals, (int)max_jobs, 0, 27, b->stream));
  BF_CUDA(cudaMalloc(&b->d_tmp, b->tmp_bytes));
  BF_CUDA(cudaMalloc(&b->d_ctl, CTL_INTS * sizeof(int)));
  BF_CUDA(cudaMalloc(&b->d_slow, max_jobs * 4));
  BF_CUDA(cudaMalloc(&b->d_out, max_jobs * sizeof(bwaflow_ext_result_t)));
  BF_CUDA(cudaMallocHost(&b->h_out, max_jobs * sizeof(bwaflow_ext_result_t)));
  BF_CUDA(cudaMallocHost(&b->h_err, 4 * sizeof(int)));
  for (int c = 0; c < EXT_NCLASS; ++c) {
    const size_t smem = (size_t)h_ext_threads[c] * (h_ext_qmax[c] + 1) * sizeof(uint32_t);
    int per_sm = 0;
    if (h_ext_threads[c] == 64) {
      BF_CUDA(cudaFuncSetAttribute(ext_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_kernel<64>, 64, smem));
    } else {
      BF_CUDA(cudaFuncSetAttribute(ext_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_kernel<32>, 32, smem));
    }
    BF_REQUIRE(per_sm > 0, -12, "extension kernel class %d does not fit on an SM", c);
    b->blocks[c] = per_sm * ctx->n_sm;
    b->blocks_pair[c] = 0;
    if (c < EXT_PAIR_CLASSES) {
      BF_CUDA(cudaFuncSetAttribute(ext_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      BF_CUDA(cudaFuncSetAttribute(ext_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int per_sm2 = 0;
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, ext_pair_kernel<true>, 32, pair_smem(h_ext_qmax[c])));
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_pair_kernel<false>, 32, pair_smem(h_ext_qmax[c])));
      if (per_sm2 < per_sm) per_sm = per_sm2;
      BF_REQUIRE(per_sm > 0, -12, "packed extension kernel class %d does not fit on an SM", c);
      b->blocks_pair[c] = per_sm * ctx->n_sm;
      BF_CUDA(cudaFuncSetAttribute(ext_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_cols_kernel, 32, cols_smem(h_ext_qmax[c])));
      BF_REQUIRE(per_sm > 0, -12, "two-column extension kernel class %d does not fit on an SM", c);
      b->blocks_cols[c] = per_sm * ctx->n_sm;
    }
  }
  *out = b;
  return 0;
}

void bwaflow_ext_batch_destroy(bwaflow_ext_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  cudaFree(b->d_jobs); cudaFree(b->d_q); cudaFree(b->d_t);
  for (int i = 0; i < 2; ++i) { cudaFree(b->d_keys[i]); cudaFree(b->d_vals[i]); }
  cudaFree(b->d_tmp); cudaFree(b->d_ctl); cudaFree(b->d_out); cudaFree(b->d_slow);
  cudaFreeHost(b->h_out); cudaFreeHost(b->h_err);
  b->tm.destroy();
  cudaEventDestroy(b->ev_fork);
  for (int c = 0; c < EXT_NCLASS; ++c) { cudaStreamDestroy(b->cstream[c]); cudaEventDestroy(b->ev_join[c]); }
  cudaStreamDestroy(b->stream);
  free(b);
}

static int ext_queue_d2h(bwaflow_ext_batch *b)
{
  BF_CUDA(cudaEventRecord(b->tm.ev[2], b->stream));
  if (b->n_jobs) BF_CUDA(cudaMemcpyAsync(b->h_out, b->d_out, b->n_jobs * sizeof(bwaflow_ext_result_t), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaMemcpyAsync(b->h_err, b->d_ctl + CTL_ERR, 4 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[3], b->stream));
  b->pending = true;
  return 0;
}

int bwaflow_ext_submit(bwaflow_ext_batch_t *b, int64_t n_jobs, const bwaflow_ext_job_t *jobs,
                       const uint8_t *qpool, int64_t qbytes, const uint8_t *tpool, int64_t tbytes)
{
  BF_REQUIRE(b && (n_jobs == 0 || (jobs && qpool && tpool)), -22, "bwaflow_ext_submit: null argument");
  BF_REQUIRE(n_jobs >= 0 && n_jobs <= b->max_jobs && qbytes <= b->max_q && tbytes <= b->max_t, -7,
             "bwaflow_ext_submit: batch capacity exceeded (%lld jobs, %lld+%lld bytes)", (long long)n_jobs, (long long)qbytes, (long long)tbytes);
  BF_CUDA(cudaSetDevice(b->ctx->device));
  b->n_jobs = n_jobs;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  if (n_jobs) {
    BF_CUDA(cudaMemcpyAsync(b->d_jobs, jobs, n_jobs * sizeof(bwaflow_ext_job_t), cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_q, qpool, qbytes, cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_t, tpool, tbytes, cudaMemcpyHostToDevice, b->stream));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  b->count_cells = false; b->cells_known = 0.0;
  int r = ext_launch(b);
  if (r) return r;
  return ext_queue_d2h(b);
}

static int ext_rerun(bwaflow_ext_batch *b, bool counted)
{
  BF_REQUIRE(b, -22, "bwaflow_ext_rerun: null batch");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  b->count_cells = counted;
  int r = ext_launch(b);
  if (r) return r;
  return ext_queue_d2h(b);
}

int bwaflow_ext_rerun(bwaflow_ext_batch_t *b) { return ext_rerun(b, false); }
int bwaflow_ext_rerun_counted(bwaflow_ext_batch_t *b) { return ext_rerun(b, true); }

int bwaflow_ext_collect(bwaflow_ext_batch_t *b, const bwaflow_ext_result_t **out, bwaflow_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_ext_collect: nothing submitted");
  BF_CUDA(cudaStreamSynchronize(b->stream));
  b->pending = false;
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->tm.ev[0], b->tm.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, b->tm.ev[1], b->tm.ev[2]);
    cudaEventElapsedTime(&t->d2h_ms, b->tm.ev[2], b->tm.ev[3]);
    cudaEventElapsedTime(&t->total_ms, b->tm.ev[0], b->tm.ev[3]);
    t->n_launches = b->n_jobs ? 3 + EXT_NCLASS + 1 + 7 : 0;   // key + ranges + 8 class kernels + deferred queue + cub passes
    if (b->count_cells || !b->use_pair) b->cells_known = (double)*(unsigned long long *)(b->h_err + 2);
    t->work_units = b->cells_known;                      // DP cells the reference executes for these jobs (0 = not counted yet)
  }
  BF_REQUIRE(*b->h_err != 1, -34, "extension job with qlen > %d (or negative length)", h_ext_qmax[EXT_NCLASS - 1]);
  BF_REQUIRE(*b->h_err != 4, -22, "extension job whose zdrop differs from the context's (bwaflow_ctx_set_opt): ksw_extend2 is always called with opt->zdrop (bwa/bwamem.c:734,762)");
  BF_REQUIRE(*b->h_err != 2, -34, "extension job whose scores may exceed %d (h0 + qlen*max_mat) or with h0 <= 0", KSW_MAX_SCORE);
  if (out) *out = b->h_out;
  return 0;
}

} // extern "C"
