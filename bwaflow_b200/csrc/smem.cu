// Stage 2 on the device: reads -> SMEM intervals, exact mem_collect_intv_new semantics
// (src/bwa_wrapper.cpp:207-246: bwt_smem1a_new :325-450, pass-2 re-seeding, bwt_seed_strategy1
// bwa/bwt.c:358-379, final sort by info), plus SA lookups (bwt_sa, bwa/bwt.c:86-96).
// Mirrors SeqsToChainsFPGA::compute's role (src/fpga/FPGAPipeline.cpp:582) behind an SMemTask-like
// submit/collect pair; there is no 256-interval cap and no CPU redo (:706-714) -- a full per-read
// slot is detected on the device and the batch is re-run with larger slots.
//
// The FM-index (occ-interleaved BWT, 64-byte blocks = 4 x u64 counts + 128 two-bit symbols,
// bwa/bwt.h:72-73) is re-laid out at upload into 32-byte blocks of 64 symbols (internal.cuh); an occurrence lookup is
// two 128-bit loads of one aligned 32-byte sector.
// Seeding is a chain of dependent random block reads, so the kernel keeps one read per lane and many
// lanes in flight, and the irregular control flow of bwt_smem1a_new is written as a per-lane flow
// machine whose ONLY expensive operation -- bwt_extend (bwa/bwt.c:262-275) -- sits at one convergent
// call site: every trip of the warp loop performs one interval extension for every live lane.
#include <cub/device/device_scan.cuh>
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "batches.cuh"
#include "fmindex.cuh"
#include "smem_args.cuh"

namespace bwaflow {


// One read per lane.  Every trip of the warp loop performs ONE interval extension for every live lane at a single
// convergent call site (the expensive part: index loads + popcounts), followed by a short consume step for the four
// kinds of inner loop of the algorithm (forward sweep, back-enlarge, forward-enlarge, pass-3 forward seed) that in
// the common case just advances a cursor and requests the next extension.  Only loop exits (phase changes) go
// through the divergent transition machine below.  The read and the end coordinates of the forward list live in
// shared memory (one row per lane); the 16-byte interval entries of the two lists are per-lane local arrays.
enum { PH_F = 0, PH_BE = 1, PH_FE = 2, PH_P3 = 3 };
enum { T_NONE = 0, T_FETCH, T_NEXT, T_F_DONE, T_B_SELECT, T_BE_DONE, T_FE_PICK, T_FE_FAIL, T_FE_REACHED, T_B_NEXT, T_CALL_DONE,
       T_FINISH, T_EXIT };

#ifndef SMEM_MIN_BLOCKS
#define SMEM_MIN_BLOCKS 6
#endif
constexpr int SMEM_FEND = 64;   // forward-list end coordinates kept in shared memory; later entries read their local-memory copy

template <int MAXL>
__global__ void __launch_bounds__(128, MAXL <= 256 ? SMEM_MIN_BLOCKS : 1) smem_kernel(const SmemArgs a)
{
  constexpr int QROW = MAXL + 4;                 // bytes per lane for the read (odd word count: conflict-free rows)
  constexpr bool QS = MAXL <= 256;               // longer reads stay in global memory
  extern __shared__ uint8_t s_mem[];
  uint8_t *q_s = s_mem + (size_t)threadIdx.x * QROW;
  uint8_t *fend = s_mem + (QS ? (size_t)blockDim.x * QROW : 0) + (size_t)threadIdx.x * SMEM_FEND;   // forward-list end coordinates
  uint4 fwd[MAXL], back[MAXL];
  const DevIndex &ix = a.ix;
  const unsigned full = 0xffffffffu;
  const int msl = a.min_seed_len;
  unsigned long long blocks = 0;
  // end coordinate of forward-list entry k: a byte in shared memory for reads <= 256, else the entry's own field
  auto get_end = [&](int k) -> int { return QS && k < SMEM_FEND ? (int)fend[k] : entry_pos(fwd[k]); };

  int ts = T_FETCH;                              // transition state; T_NONE = an extension is pending
  int64_t r = -1; int len = 0; const uint8_t *q = nullptr; bwaflow_intv_t *mem = nullptr; int n_mem = 0;
  int pass = 1, x = 0, k2 = 0, old_n = 0;
  int sx = 0, mi = 1, fwd_n = 0, back_n = 0, fi = 0, start = 0, stop = 0, ret = 0, px = 0, bk = 0, f_end = 0, ik_beg = 0;
  int phase = PH_F, cur = 0, c = 0;
  Intv ik = {0, 0, 0}, cand = {0, 0, 0}; int cand_beg = 0, cand_end = 0;

  auto push_mem = [&](const Intv &tv, int beg, int end) {
    if (n_mem < a.cap) {
      const Intv v = tv;
      bwaflow_intv_t o; o.x[0] = v.x0; o.x[1] = v.x1; o.x[2] = v.x2; o.info = (uint64_t)(uint32_t)beg << 32 | (uint32_t)end;
      mem[n_mem] = o;
    }
    ++n_mem;
  };
  auto push_fwd = [&](const Intv &v, int end) { fwd[fwd_n] = pack_entry(v, end); if (QS && fwd_n < SMEM_FEND) fend[fwd_n] = (uint8_t)end; ++fwd_n; };
  auto take_cand = [&](int beg, bool first) {   // src/bwa_wrapper.cpp:392-400 / 418-427
    if (!first && beg > cand_beg && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
    cand = ik; cand_beg = beg; cand_end = f_end;
  };

  for (;;) {
    // ---- transitions: run until this lane has an extension pending (or no reads are left)
    while (ts != T_NONE && ts != T_EXIT) {
      switch (ts) {
      case T_FETCH: {
        r = (int64_t)atomicAdd(a.ctl, 1);
        if (r >= a.n_reads) { ts = T_EXIT; break; }
        const int64_t ro = a.read_off[r];
        len = (int)(a.read_off[r + 1] - ro);
        mem = a.slots + r * a.cap; n_mem = 0; pass = 1; x = 0;
        if (len >= MAXL) { atomicExch(a.ctl + 2, 1); a.counts[r] = 0; break; }
        if (QS) { for (int t = 0; t < len; ++t) q_s[t] = a.bases[ro + t]; q = q_s; } else q = a.bases + ro;
        ts = len < msl ? T_FINISH : T_NEXT;                  // src/bwa_wrapper.cpp:150
        break;
      }
      case T_NEXT: {                                         // the three loops of mem_collect_intv_new (:214-243)
        bool call = false;
        if (pass == 1) {
          while (x < len && q[x] > 3) ++x;
          if (x < len) { sx = x; mi = 1; call = true; }
          else { pass = 2; old_n = n_mem < a.cap ? n_mem : a.cap; k2 = 0; }
        }
        if (pass == 2 && !call) {
          while (k2 < old_n) {
            const bwaflow_intv_t p = mem[k2++];
            const int s = (int)(p.info >> 32), e = (int)(uint32_t)p.info;
            if (e - s < a.split_len || p.x[2] > (uint64_t)a.split_width) continue;
            if (q[(s + e) >> 1] > 3) continue;               // bwt_smem1a_new :338
            sx = (s + e) >> 1; mi = (int)p.x[2] + 1; call = true;
            break;
          }
          if (!call) { pass = 3; x = 0; }
        }
        if (call) {                                          // bwt_smem1a_new entry (:338-346)
          fwd_n = 0; back_n = 0; ik = set_intv(ix, q[sx]); cur = sx + 1; phase = PH_F;
          if (cur < len && q[cur] < 4) { c = 3 - q[cur]; ts = T_NONE; }
          else { push_fwd(ik, cur); ts = T_F_DONE; }
          break;
        }
        ts = T_FINISH;                                       // pass 3 runs in smem_p3_kernel
        break;
      }
      case T_F_DONE:                                         // :364-366
        ret = get_end(fwd_n - 1); start = sx; stop = sx; fi = 0;
        ts = T_B_SELECT;
        break;
      case T_B_SELECT: {                                     // :367-371
        if (fi >= fwd_n) { ts = T_CALL_DONE; break; }
        int dummy;
        ik = unpack_entry(fwd[fi], &dummy); f_end = get_end(fi); ik_beg = sx;
        if (back_n == 0 || stop - start >= 3) {              // back-enlarge
          back[0] = pack_entry(ik, 0); back_n = 1; cur = sx - 1; phase = PH_BE;
          if (cur >= 0 && q[cur] < 4) { c = q[cur]; ts = T_NONE; } else ts = T_BE_DONE;
        } else { stop = f_end; bk = back_n - 1; ts = T_FE_PICK; }   // forward-enlarge
        break;
      }
      case T_BE_DONE:                                        // :387-400
        start = f_end; stop = fi == fwd_n - 1 ? len : get_end(fi + 1);
        take_cand(ik_beg, fi == 0);
        ts = T_B_NEXT;
        break;
      case T_FE_FAIL:
        --bk;
        // fall through
      case T_FE_PICK: {                                      // :405-408
        if (bk < 0) { ts = T_B_NEXT; break; }
        int dummy;
        ik = unpack_entry(back[bk], &dummy); cur = start; phase = PH_FE; c = 3 - q[cur];
        ts = T_NONE;
        break;
      }
      case T_FE_REACHED:                                     // :413-427
        take_cand(sx - bk, false);
        ts = T_B_NEXT;
        break;
      case T_B_NEXT:                                         // :431-443
        ++fi;
        if (fi < fwd_n) {
          int max_len = cand_beg + get_end(fi);
          while (max_len < msl && fi < fwd_n) {
            ++fi;
            if (fi < fwd_n) stop = get_end(fi);
            max_len = cand_beg + stop;
          }
        }
        if (fi >= fwd_n && cand_end - cand_beg >= msl) push_mem(cand, cand_beg, cand_end);
        ts = T_B_SELECT;
        break;
      case T_CALL_DONE:
        if (pass == 1) x = ret;
        ts = T_NEXT;
        break;
      case T_FINISH:                                         // passes 1 and 2 done: pass 3, the sort and the final count follow
        a.counts[r] = n_mem;
        if (n_mem > a.cap) atomicExch(a.ctl + 1, 1);
        ts = T_FETCH;
        break;
      }
    }
    if (__ballot_sync(full, ts == T_NONE) == 0) break;
    if (ts == T_NONE) {
      // ---- the one call site of bwt_extend, then the consume step of the four inner loops
      const Intv o = extend1(ix, ik, phase == PH_BE, c, &blocks);
      const bool small = o.x2 < (uint64_t)mi;
      const bool changed = o.x2 != ik.x2;
      if (phase == PH_F) {                                   // :350-357
        if (changed) push_fwd(ik, cur);
        if (changed && small) ts = T_F_DONE;
        else {
          ik = o; ++cur;
          if (cur < len && q[cur] < 4) c = 3 - q[cur];
          else { push_fwd(ik, cur); ts = T_F_DONE; }         // N or end of read (:358-363)
        }
      } else if (phase == PH_BE) {                           // :374-386
        if (small) ts = T_BE_DONE;
        else {
          ik = o; ik_beg = cur; back[back_n++] = pack_entry(ik, 0); --cur;
          if (cur >= 0 && q[cur] < 4) c = q[cur]; else ts = T_BE_DONE;
        }
      } else {                                               // forward-enlarge :407-417
        if (small) ts = T_FE_FAIL;
        else {
          ik = o; ++cur;
          if (cur == stop) ts = T_FE_REACHED; else c = 3 - q[cur];
        }
      }
    }
  }
  if (blocks) atomicAdd(a.blocks, blocks);
}

// ---- SA -----------------------------------------------------------------------------------------

// Pass 3 of mem_collect_intv_new (src/bwa_wrapper.cpp:228-243): bwt_seed_strategy1 (bwa/bwt.c:358-379) from every
// position the previous seed ended at -- a plain forward loop, so it gets its own kernel: one read per lane, one
// interval extension per trip for every live lane, nothing but a cursor between two extensions, no local memory.
// It appends to the read's slot after the SMEMs of passes 1-2, sorts the slot by info (:245) and writes the count.
template <int MAXL>
__global__ void __launch_bounds__(128) smem_p3_kernel(const SmemArgs a)
{
  constexpr int QROW = MAXL + 4;
  constexpr bool QS = MAXL <= 256;
  extern __shared__ uint8_t s_mem[];
  uint8_t *q_s = s_mem + (size_t)threadIdx.x * QROW;
  const DevIndex &ix = a.ix;
  const unsigned full = 0xffffffffu;
  const int msl = a.min_seed_len;
  unsigned long long blocks = 0;
  int64_t r = -1; int len = 0, x = 0, px = 0, cur = 0, c = 0, n_mem = 0; const uint8_t *q = nullptr; bwaflow_intv_t *mem = nullptr;
  Intv ik = {0, 0, 0};
  bool pending = false, live = true;
  for (;;) {
    while (live && !pending) {                               // find the next extension of this lane
      if (r < 0) {
        r = (int64_t)atomicAdd(a.ctl + 3, 1);
        if (r >= a.n_reads) { live = false; break; }
        const int64_t ro = a.read_off[r];
        len = (int)(a.read_off[r + 1] - ro);
        mem = a.slots + r * a.cap; n_mem = a.counts[r]; x = 0;
        if (len >= MAXL || len < msl || a.max_mem_intv <= 0) x = len;   // nothing to add (too-long reads are flagged by pass 1)
        else if (QS) { for (int t = 0; t < len; ++t) q_s[t] = a.bases[ro + t]; q = q_s; } else q = a.bases + ro;
      }
      while (x < len && q[x] > 3) ++x;
      if (x >= len) {                                        // read finished: sort by info, final count
        const int n = n_mem < a.cap ? n_mem : a.cap;
        for (int u = 1; u < n; ++u) {
          const bwaflow_intv_t v = mem[u];
          int t = u - 1;
          while (t >= 0 && mem[t].info > v.info) { mem[t + 1] = mem[t]; --t; }
          mem[t + 1] = v;
        }
        a.counts[r] = n_mem;
        if (n_mem > a.cap) atomicExch(a.ctl + 1, 1);
        r = -1;
        continue;
      }
      px = x; ik = set_intv(ix, q[x]); cur = x + 1;          // bwa/bwt.c:364-366
      if (cur >= len) { x = len; continue; }
      if (q[cur] > 3) { x = cur + 1; continue; }
      c = 3 - q[cur]; pending = true;
    }
    if (__ballot_sync(full, pending) == 0) break;
    if (pending) {
      const Intv o = extend1(ix, ik, 0, c, &blocks);         // bwa/bwt.c:367-376
      if (o.x2 < (uint64_t)a.max_mem_intv && cur - px >= msl) {
        if (o.x2 > 0) {
          if (n_mem < a.cap) {
            bwaflow_intv_t w; w.x[0] = o.x0; w.x[1] = o.x1; w.x[2] = o.x2; w.info = (uint64_t)(uint32_t)px << 32 | (uint32_t)(cur + 1);
            mem[n_mem] = w;
          }
          ++n_mem;
        }
        x = cur + 1; pending = false;
      } else {
        ik = o; ++cur;
        if (cur >= len) { x = len; pending = false; }
        else if (q[cur] > 3) { x = cur + 1; pending = false; }
        else c = 3 - q[cur];
      }
    }
  }
  if (blocks) atomicAdd(a.blocks, blocks);
}


// One thread per row sampled by bwa (every 2^ref_shift-th): walks invPsi until the next sampled row, handing its value minus
// the number of steps to every row it passes whose index is a multiple of 2^dense_shift.  The walks partition all rows.
__global__ void sa_densify_kernel(DevIndex ix, uint64_t n_sa, int dense_shift, uint64_t *dense)
{
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_sa) return;
  const uint64_t ref_mask = (1ull << ix.sa_shift) - 1, dmask = (1ull << dense_shift) - 1;
  uint64_t r = j << ix.sa_shift, v = ix.sa[j];
  dense[r >> dense_shift] = v;
  if (j == 0) v = ix.seq_len;                  // sa[0] is stored as -1 == seq_len modulo the number of rows
  for (;;) {
    r = inv_psi(ix, r);
    --v;
    if ((r & ref_mask) == 0) break;
    if ((r & dmask) == 0) dense[r >> dense_shift] = v;
  }
}

int densify_sa(bwaflow_ctx *c, int dense_shift, uint64_t n_sa, void **d_dense)
{
  const uint64_t rows = c->ix.seq_len + 1;
  const uint64_t n_dense = ((rows - 1) >> dense_shift) + 1;
  BF_CUDA(cudaMalloc(d_dense, n_dense * 8));
  sa_densify_kernel<<<(unsigned)((n_sa + 127) / 128), 128>>>(c->ix, n_sa, dense_shift, (uint64_t *)*d_dense);
  BF_CUDA(cudaGetLastError());
  BF_CUDA(cudaDeviceSynchronize());
  return 0;
}

__global__ void sa_kernel(DevIndex ix, int64_t n, const uint64_t *k, uint64_t *out, unsigned long long *blocks)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nb = 0;
  if (i < n) out[i] = sa_lookup(ix, k[i], &nb);
  if (nb) atomicAdd(blocks, nb);
}

__global__ void smem_compact_counts(const int *counts, int64_t n, int cap, int64_t *out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = counts[i] < cap ? counts[i] : cap;
  if (i == n) out[i] = 0;
}

__global__ void smem_compact(const bwaflow_intv_t *slots, int cap, const int64_t *off, int64_t n_reads, bwaflow_intv_t *out)
{ // one warp per read, 32-byte intervals copied as uint4 halves
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_reads) return;
  const int64_t lo = off[r], n = off[r + 1] - lo;
  const uint4 *src = (const uint4 *)(slots + r * cap);
  uint4 *dst = (uint4 *)(out + lo);
  for (int64_t t = lane; t < 2 * n; t += 32) dst[t] = src[t];
}

// number of SA rows sampled per interval: the loop of src/bwa_wrapper.cpp:286-287 visits min(x2, max_occ) rows
__global__ void seedpos_counts(const bwaflow_intv_t *intv, int64_t n, int max_occ, int64_t *out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = intv[i].x[2] < (uint64_t)max_occ ? (int64_t)intv[i].x[2] : max_occ;
  if (i == n) out[i] = 0;
}

__global__ void seedpos_kernel(DevIndex ix, const bwaflow_intv_t *intv, int64_t n_intv, const int64_t *off, int64_t total,
                               int max_occ, int64_t *pos, unsigned long long *blocks)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nb = 0;
  if (t < total) {
    int64_t lo = 0, hi = n_intv;                       // last interval with off[iv] <= t
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (off[mid] <= t) lo = mid; else hi = mid; }
    const bwaflow_intv_t p = intv[lo];
    const uint64_t step = p.x[2] > (uint64_t)max_occ ? p.x[2] / max_occ : 1;
    pos[t] = p.x[1] == TEXT_MARK ? (int64_t)p.x[0]                    // text-anchored interval (smem_fast.cu): x[0] is the position
                                 : (int64_t)sa_lookup(ix, p.x[0] + (uint64_t)(t - off[lo]) * step, &nb);
  }
  if (nb) atomicAdd(blocks, nb);
}

// ---- device-driven tail for the fused hot path: nothing is read back between the stages.  Every kernel is bounded by the batch's
// capacities, takes its item count from device memory, and reports an overrun in ctl[6] (bit 0: intervals, bit 1: positions);
// the host looks at the flags once per batch (hotpath.cu) and only then grows the buffers and redoes the batch.
__global__ void smem_compact_dev(const bwaflow_intv_t *slots, int cap, const int64_t *off, int64_t n_reads, bwaflow_intv_t *out,
                                 int64_t out_cap, int *flags)
{
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_reads) return;
  const int64_t lo = off[r], n = off[r + 1] - lo;
  if (lo + n > out_cap) { if (lane == 0 && n) atomicOr(flags, 1); return; }
  const uint4 *src = (const uint4 *)(slots + r * cap);
  uint4 *dst = (uint4 *)(out + lo);
  for (int64_t t = lane; t < 2 * n; t += 32) dst[t] = src[t];
}

__global__ void seedpos_counts_dev(const bwaflow_intv_t *intv, const int64_t *n_ptr, int64_t out_cap, int max_occ, int64_t *out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > out_cap) return;
  const int64_t n = *n_ptr < out_cap ? *n_ptr : out_cap;
  out[i] = i < n ? (intv[i].x[2] < (uint64_t)max_occ ? (int64_t)intv[i].x[2] : max_occ) : 0;
}

__global__ void seedpos_dev_kernel(DevIndex ix, const bwaflow_intv_t *intv, const int64_t *n_ptr, int64_t out_cap, const int64_t *off,
                                   int64_t pos_cap, int max_occ, int64_t *pos, int *flags)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_intv = *n_ptr < out_cap ? *n_ptr : out_cap;
  int64_t total = off[n_intv];
  if (total > pos_cap) { if (t == 0) atomicOr(flags, 2); total = pos_cap; }
  if (t >= total) return;
  int64_t lo = 0, hi = n_intv;                         // last interval with off[iv] <= t
  while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (off[mid] <= t) lo = mid; else hi = mid; }
  const bwaflow_intv_t p = intv[lo];
  const uint64_t step = p.x[2] > (uint64_t)max_occ ? p.x[2] / max_occ : 1;
  unsigned long long nb = 0;
  pos[t] = p.x[1] == TEXT_MARK ? (int64_t)p.x[0] : (int64_t)sa_lookup(ix, p.x[0] + (uint64_t)(t - off[lo]) * step, &nb);
}

} // namespace bwaflow

using namespace bwaflow;


static int smem_alloc_slots(bwaflow_smem_batch *b, int cap)
{
  if (b->d_slots) cudaFree(b->d_slots);
  b->d_slots = nullptr;
  b->cap = cap; b->slots_cap = b->max_reads * cap;
  // every read of the batch gets a slot of `cap` intervals (the slots grow x4 for all reads when one read fills its own): say
  // so when that does not fit instead of failing in cudaMalloc
  size_t free_b = 0, total_b = 0;
  BF_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t need = (size_t)b->slots_cap * sizeof(bwaflow_intv_t);
  BF_REQUIRE(need <= free_b, -12, "interval slots of %lld reads x %d intervals need %.1f GB of device memory, %.1f GB are free: a read of this batch "
             "has more than %d intervals; use smaller batches", (long long)b->max_reads, cap, need / 1e9, free_b / 1e9, cap / 4);
  BF_CUDA(cudaMalloc(&b->d_slots, need));
  return 0;
}

static int smem_grow_out(bwaflow_smem_batch *b, int64_t total)
{
  cudaFree(b->d_out); cudaFreeHost(b->h_out); cudaFree(b->d_pcnt); cudaFree(b->d_poff); cudaFreeHost(b->h_poff); cudaFree(b->d_tmp);
  b->out_cap = total + total / 4;
  BF_CUDA(cudaMalloc(&b->d_out, b->out_cap * sizeof(bwaflow_intv_t)));
  BF_CUDA(cudaMallocHost(&b->h_out, b->out_cap * sizeof(bwaflow_intv_t)));
  BF_CUDA(cudaMalloc(&b->d_pcnt, (b->out_cap + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_poff, (b->out_cap + 2) * 8));
  BF_CUDA(cudaMallocHost(&b->h_poff, (b->out_cap + 2) * 8));
  BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b->tmp_bytes, b->d_cnt64, b->d_off, (int)(b->out_cap + 1), b->stream));
  b->tmp_bytes += 256;
  BF_CUDA(cudaMalloc(&b->d_tmp, b->tmp_bytes));
  return 0;
}

int bwaflow::smem_launch(bwaflow_smem_batch *b)
{
  const int64_t nr = b->n_reads;
  cudaStream_t st = b->stream;
  BF_CUDA(cudaMemsetAsync(b->d_ctl, 0, 8 * sizeof(int), st));
  BF_CUDA(cudaMemsetAsync(b->d_defer, 0, 2 * sizeof(int), st));
  BF_CUDA(cudaMemsetAsync(b->d_off, 0, (nr + 1) * 8, st));
  if (nr == 0) return 0;
  const bwaflow_opt_t &o = b->ctx->par.opt;
  SmemArgs a;
  a.bases = b->d_bases; a.read_off = b->d_read_off; a.n_reads = nr; a.slots = b->d_slots; a.cap = b->cap;
  a.counts = b->d_counts; a.ctl = b->d_ctl; a.blocks = (unsigned long long *)(b->d_ctl + 4); a.ix = b->ctx->ix;
  a.min_seed_len = o.min_seed_len; a.split_len = (int)(o.min_seed_len * o.split_factor + .499);
  a.split_width = o.split_width; a.max_mem_intv = o.max_mem_intv;
  a.packed = b->d_packed; a.x_min = b->x_min; a.defer = b->d_defer; a.heavy_min = b->heavy_min;
  const bool count = b->count_blocks;
  if (b->max_len < 256 && !getenv("BWAFLOW_SMEM_LEGACY")) {
    // reads shorter than 256 bases: table / text-anchored kernels (smem_fast.cu).  Text anchoring needs the full suffix
    // array and yields positions instead of rows, so the interval API (want_rows) and the counting runs go without it.
    const bool text = !b->want_rows && !count && a.ix.text && a.ix.sa_shift == 0;
    int r = smem_fast_launch(b, a, text, count);
    if (r) return r;
  } else {
    int64_t want = (nr + 127) / 128;
    const int grid = (int)(want < b->blocks_per_launch ? want : b->blocks_per_launch);
    const int grid3 = (int)(want < (int64_t)b->ctx->n_sm * 10 ? want : (int64_t)b->ctx->n_sm * 10);
    // dynamic shared memory: per lane one row for the read (<= 256 bases) and one for the forward-list end coordinates
    if (b->max_len < 160) {
      smem_kernel<160><<<grid, 128, 128 * (164 + SMEM_FEND), st>>>(a);
      smem_p3_kernel<160><<<grid3, 128, 128 * 164, st>>>(a);
    } else if (b->max_len < 256) {
      smem_kernel<256><<<grid, 128, 128 * (260 + SMEM_FEND), st>>>(a);
      smem_p3_kernel<256><<<grid3, 128, 128 * 260, st>>>(a);
    } else {
      smem_kernel<1024><<<grid, 128, 0, st>>>(a);
      smem_p3_kernel<1024><<<grid3, 128, 0, st>>>(a);
    }
  }
  smem_compact_counts<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, st>>>(b->d_counts, nr, b->cap, b->d_cnt64);
  size_t tmp = b->tmp_bytes;
  BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_cnt64, b->d_off, (int)(nr + 1), st));
  BF_CUDA(cudaGetLastError());
  return 0;
}

static int smem_queue_tail(bwaflow_smem_batch *b)
{
  BF_CUDA(cudaEventRecord(b->tm.ev[2], b->stream));
  BF_CUDA(cudaMemcpyAsync(b->h_off, b->d_off, (b->n_reads + 1) * 8, cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  b->pending = true;
  return 0;
}


// Waits for the seeding kernel, redoes the batch with larger slots if some read overflowed its slot, then compacts
// the per-read slots into d_out / d_off (device) and h_off (host).  Leaves b->n_intv set.
int bwaflow::smem_finish_device(bwaflow_smem_batch *b)
{
  BF_CUDA(cudaSetDevice(b->ctx->device));
  BF_CUDA(cudaStreamSynchronize(b->stream));
  while (b->h_ctl[1]) {                      // some read filled its slot: grow the slots and redo the batch on the GPU
    BF_REQUIRE(b->cap < 4096, -34, "a read produced more than 4096 intervals");
    int r = smem_alloc_slots(b, b->cap * 4);
    if (r) return r;
    if ((r = smem_launch(b))) return r;
    if ((r = smem_queue_tail(b))) return r;
    BF_CUDA(cudaStreamSynchronize(b->stream));
  }
  b->pending = false;
  BF_REQUIRE(!b->h_ctl[2], -34, "a read is too long for the seeding kernel instance chosen for this batch");
  const int64_t total = b->n_reads ? b->h_off[b->n_reads] : 0;
  if (total > b->out_cap) { int r = smem_grow_out(b, total); if (r) return r; }
  if (total)
    smem_compact<<<(unsigned)((b->n_reads * 32 + 255) / 256), 256, 0, b->stream>>>(b->d_slots, b->cap, b->d_off, b->n_reads, b->d_out);
  BF_CUDA(cudaGetLastError());
  b->n_intv = total;
  return 0;
}

// Hot path, all asynchronous: compacted intervals, seed-position offsets and positions for the batch on its stream.
int bwaflow::smem_tail_device(bwaflow_smem_batch *b)
{
  cudaStream_t st = b->stream;
  const int64_t nr = b->n_reads;
  const int max_occ = b->ctx->par.opt.max_occ;
  if (nr == 0) { BF_CUDA(cudaMemsetAsync(b->d_poff, 0, 16, st)); return 0; }
  smem_compact_dev<<<(unsigned)((nr * 32 + 255) / 256), 256, 0, st>>>(b->d_slots, b->cap, b->d_off, nr, b->d_out, b->out_cap, b->d_ctl + 6);
  seedpos_counts_dev<<<(unsigned)((b->out_cap + 1 + 255) / 256), 256, 0, st>>>(b->d_out, b->d_off + nr, b->out_cap, max_occ, b->d_pcnt);
  size_t tmp = b->tmp_bytes;
  BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_pcnt, b->d_poff, (int)(b->out_cap + 1), st));
  seedpos_dev_kernel<<<(unsigned)((b->pos_cap + 127) / 128), 128, 0, st>>>(b->ctx->ix, b->d_out, b->d_off + nr, b->out_cap, b->d_poff, b->pos_cap,
                                                                         max_occ, b->d_pos, b->d_ctl + 6);
  BF_CUDA(cudaGetLastError());
  return 0;
}

// queues the read-back of the batch's status words: h_stat = {n_intv, n_pos}, h_ctl = the control block
int bwaflow::smem_status_d2h(bwaflow_smem_batch *b)
{
  cudaStream_t st = b->stream;
  BF_CUDA(cudaMemcpyAsync(b->h_stat, b->d_off + b->n_reads, 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(b->h_stat + 1, b->d_poff + b->out_cap, 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  return 0;
}

// After the stream has been synchronised: did the batch fit?  Grows what overflowed (slots x4, interval / position arrays to 1.25x
// the true counts) and sets *redo when the batch has to be run again; otherwise leaves n_intv set and *n_pos.
int bwaflow::smem_check_capacity(bwaflow_smem_batch *b, bool *redo, int64_t *n_pos)
{
  *redo = false;
  BF_REQUIRE(!b->h_ctl[2], -34, "a read is too long for the seeding kernel instance chosen for this batch");
  if (b->n_reads == 0) { b->n_intv = 0; *n_pos = 0; return 0; }
  if (b->h_ctl[1]) {                         // some read filled its slot
    BF_REQUIRE(b->cap < 4096, -34, "a read produced more than 4096 intervals");
    int r = smem_alloc_slots(b, b->cap * 4);
    if (r) return r;
    *redo = true;
  }
  const int64_t total = b->h_stat[0];
  if (total > b->out_cap) {
    int r = smem_grow_out(b, total);
    if (r) return r;
    *redo = true;
  }
  if (!*redo && ((b->h_ctl[6] & 2) || b->h_stat[1] > b->pos_cap)) {
    cudaFree(b->d_pos);
    b->pos_cap = b->h_stat[1] + b->h_stat[1] / 4 + 1024;
    BF_CUDA(cudaMalloc(&b->d_pos, b->pos_cap * 8));
    *redo = true;
  }
  b->n_intv = total; *n_pos = b->h_stat[1];
  return 0;
}

// SA positions of the current intervals into d_pos / d_poff (device); *total = number of positions.
int bwaflow::smem_seedpos_device(bwaflow_smem_batch *b, int64_t *total_out)
{
  BF_CUDA(cudaSetDevice(b->ctx->device));
  cudaStream_t st = b->stream;
  const int64_t ni = b->n_intv;
  const int max_occ = b->ctx->par.opt.max_occ;
  BF_CUDA(cudaMemsetAsync(b->d_ctl + 4, 0, 8, st));
  BF_CUDA(cudaMemsetAsync(b->d_poff, 0, (ni + 1) * 8, st));
  if (ni) {
    seedpos_counts<<<(unsigned)((ni + 1 + 255) / 256), 256, 0, st>>>(b->d_out, ni, max_occ, b->d_pcnt);
    size_t tmp = b->tmp_bytes;
    BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_pcnt, b->d_poff, (int)(ni + 1), st));
  }
  BF_CUDA(cudaMemcpyAsync(b->h_poff, b->d_poff, (ni + 1) * 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  const int64_t total = b->h_poff[ni];
  if (total > b->pos_cap) {
    cudaFree(b->d_pos);
    b->pos_cap = total + total / 4 + 1024;
    BF_CUDA(cudaMalloc(&b->d_pos, b->pos_cap * 8));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], st));
  if (total) {
    seedpos_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(b->ctx->ix, b->d_out, ni, b->d_poff, total, max_occ, b->d_pos,
                                                                      (unsigned long long *)(b->d_ctl + 4));
    BF_CUDA(cudaGetLastError());
  }
  *total_out = total;
  return 0;
}

extern "C" {

int bwaflow_smem_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, bwaflow_smem_batch_t **out)
{
  BF_REQUIRE(ctx && out, -22, "bwaflow_smem_batch_create: null argument");
  BF_REQUIRE(ctx->has_index, -22, "bwaflow_smem_batch_create: the context has no index");
  BF_REQUIRE(max_reads > 0 && max_reads < (1LL << 31) - 1, -22, "max_reads out of range");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_smem_batch *b = (bwaflow_smem_batch *)calloc(1, sizeof(*b));
  b->ctx = ctx; b->max_reads = max_reads; b->max_bases = max_bases; b->want_rows = true;
  BF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (b->tm.init()) return -5;
  BF_CUDA(cudaMalloc(&b->d_bases, max_bases + 16));
  BF_CUDA(cudaMalloc(&b->d_read_off, (max_reads + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_counts, (max_reads + 1) * 4));
  BF_CUDA(cudaMalloc(&b->d_cnt64, (max_reads + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_off, (max_reads + 2) * 8));
  BF_CUDA(cudaMallocHost(&b->h_off, (max_reads + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_ctl, 8 * sizeof(int)));
  BF_CUDA(cudaMalloc(&b->d_defer, (max_reads + 4) * sizeof(int)));
  // forward matches of one call from which a read goes to the warp-per-read kernel.  An ordinary read has about log4(index rows)
  // + 3 -- the interval shrinks at every base until it is unique, i.e. over about as many bases as the level tables are deep --
  // a read inside a repeat several times that.  Measured at 3.1 Gbp (K = 14): 16 hands every read over (seeding 128 ms), 20
  // some ordinary ones (14.2 ms), 24 none (13.7 ms); on the repeat-rich variant 19.3 ms at 40, 16.9 at 24, 15.8 at 20.
  b->heavy_min = (ctx->ix.ktab_k > 0 ? ctx->ix.ktab_k : 16) + 10;
  if (const char *e = getenv("BWAFLOW_SMEM_HEAVY_MIN")) b->heavy_min = atoi(e);
  BF_CUDA(cudaMallocHost(&b->h_ctl, 8 * sizeof(int)));
  int r = smem_alloc_slots(b, 32);
  if (r) return r;
  b->out_cap = max_reads * 16;
  BF_CUDA(cudaMalloc(&b->d_out, b->out_cap * sizeof(bwaflow_intv_t)));
  BF_CUDA(cudaMallocHost(&b->h_out, b->out_cap * sizeof(bwaflow_intv_t)));
  BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b->tmp_bytes, b->d_cnt64, b->d_off, (int)(b->out_cap + 1), b->stream));
  b->tmp_bytes += 256;
  BF_CUDA(cudaMalloc(&b->d_tmp, b->tmp_bytes));
  BF_CUDA(cudaMalloc(&b->d_pcnt, (b->out_cap + 2) * 8));
  BF_CUDA(cudaMalloc(&b->d_poff, (b->out_cap + 2) * 8));
  BF_CUDA(cudaMallocHost(&b->h_poff, (b->out_cap + 2) * 8));
  b->pos_cap = max_reads * 12 + 1024; b->h_pos = 0;     // grown to 1.25x the need when a batch overruns it
  BF_CUDA(cudaMalloc(&b->d_pos, b->pos_cap * 8));
  BF_CUDA(cudaMallocHost(&b->h_stat, 4 * 8));
  int per_sm = 0;
  BF_CUDA(cudaFuncSetAttribute(smem_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * (164 + SMEM_FEND)));
  BF_CUDA(cudaFuncSetAttribute(smem_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * (260 + SMEM_FEND)));
  BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smem_kernel<160>, 128, 128 * (164 + SMEM_FEND)));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  if (const char *s = getenv("BWAFLOW_SMEM_BLOCKS_PER_SM")) per_sm = atoi(s) > 0 ? atoi(s) : per_sm;
  b->blocks_per_launch = per_sm * ctx->n_sm;
  if ((r = smem_fast_setup(b))) return r;
  *out = b;
  return 0;
}

void bwaflow_smem_batch_destroy(bwaflow_smem_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  cudaFree(b->d_bases); cudaFree(b->d_read_off); cudaFree(b->d_counts); cudaFree(b->d_cnt64); cudaFree(b->d_off);
  cudaFreeHost(b->h_off); cudaFree(b->d_ctl); cudaFree(b->d_defer); cudaFreeHost(b->h_ctl); cudaFree(b->d_slots); cudaFree(b->d_out);
  cudaFreeHost(b->h_out); cudaFree(b->d_tmp); cudaFree(b->d_pcnt); cudaFree(b->d_poff); cudaFreeHost(b->h_poff);
  cudaFree(b->d_pos); cudaFreeHost(b->h_pos); cudaFree(b->d_packed); cudaFreeHost(b->h_stat);
  b->tm.destroy();
  cudaStreamDestroy(b->stream);
  free(b);
}


int bwaflow_smem_submit(bwaflow_smem_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off)
{
  BF_REQUIRE(b && (n_reads == 0 || (bases && read_off)), -22, "bwaflow_smem_submit: null argument");
  const int64_t n_bases = n_reads ? read_off[n_reads] : 0;
  BF_REQUIRE(n_reads >= 0 && n_reads <= b->max_reads && n_bases <= b->max_bases, -7,
             "bwaflow_smem_submit: batch capacity exceeded (%lld reads, %lld bases)", (long long)n_reads, (long long)n_bases);
  BF_CUDA(cudaSetDevice(b->ctx->device));
  int64_t max_len = 0;
  for (int64_t i = 0; i < n_reads; ++i) { const int64_t l = read_off[i + 1] - read_off[i]; max_len = l > max_len ? l : max_len; }
  BF_REQUIRE(max_len < 1024, -34, "read of %lld bases: the seeding kernel supports reads shorter than 1024", (long long)max_len);
  b->n_reads = n_reads; b->max_len = max_len;
  b->count_blocks = false;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  if (n_reads) {
    BF_CUDA(cudaMemcpyAsync(b->d_bases, bases, n_bases, cudaMemcpyHostToDevice, b->stream));
    BF_CUDA(cudaMemcpyAsync(b->d_read_off, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, b->stream));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  int r = smem_launch(b);
  if (r) return r;
  return smem_queue_tail(b);
}

static int smem_rerun(bwaflow_smem_batch_t *b, bool counted)
{
  BF_REQUIRE(b, -22, "bwaflow_smem_rerun: null batch");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  b->count_blocks = counted;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  int r = smem_launch(b);
  if (r) return r;
  return smem_queue_tail(b);
}

int bwaflow_smem_set_anchored(bwaflow_smem_batch_t *b, int on)
{
  BF_REQUIRE(b, -22, "bwaflow_smem_set_anchored: null batch");
  b->want_rows = !on;
  return 0;
}

int bwaflow_smem_rerun(bwaflow_smem_batch_t *b) { return smem_rerun(b, false); }
int bwaflow_smem_rerun_counted(bwaflow_smem_batch_t *b) { return smem_rerun(b, true); }

int bwaflow_smem_collect(bwaflow_smem_batch_t *b, const bwaflow_intv_t **intv, const int64_t **intv_off,
                         int64_t *n_intv, bwaflow_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_smem_collect: nothing submitted");
  int r = smem_finish_device(b);
  if (r) return r;
  const int64_t total = b->n_intv;
  if (total) BF_CUDA(cudaMemcpyAsync(b->h_out, b->d_out, total * sizeof(bwaflow_intv_t), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[3], b->stream));
  BF_CUDA(cudaStreamSynchronize(b->stream));
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->tm.ev[0], b->tm.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, b->tm.ev[1], b->tm.ev[2]);
    cudaEventElapsedTime(&t->d2h_ms, b->tm.ev[2], b->tm.ev[3]);
    cudaEventElapsedTime(&t->total_ms, b->tm.ev[0], b->tm.ev[3]);
    t->n_launches = b->n_reads ? 5 : 0;
    t->work_units = (double)*(unsigned long long *)(b->h_ctl + 4);   // occ blocks touched (64 B each)
  }
  if (intv) *intv = b->h_out;
  if (intv_off) *intv_off = b->h_off;
  if (n_intv) *n_intv = total;
  return 0;
}

int bwaflow_smem_seed_positions(bwaflow_smem_batch_t *b, const int64_t **pos, const int64_t **pos_off,
                                int64_t *n_pos, bwaflow_timing_t *t)
{
  BF_REQUIRE(b && !b->pending, -22, "bwaflow_smem_seed_positions: collect the batch first");
  cudaStream_t st = b->stream;
  int64_t total = 0;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], st));
  int r = smem_seedpos_device(b, &total);
  if (r) return r;
  BF_CUDA(cudaEventRecord(b->tm.ev[2], st));
  if (total > b->hpos_cap) {
    cudaFreeHost(b->h_pos);
    b->hpos_cap = total + total / 4 + 1024;
    BF_CUDA(cudaMallocHost(&b->h_pos, b->hpos_cap * 8));
  }
  if (total) BF_CUDA(cudaMemcpyAsync(b->h_pos, b->d_pos, total * 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaEventRecord(b->tm.ev[3], st));
  BF_CUDA(cudaStreamSynchronize(st));
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->tm.ev[0], b->tm.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, b->tm.ev[1], b->tm.ev[2]);
    cudaEventElapsedTime(&t->d2h_ms, b->tm.ev[2], b->tm.ev[3]);
    cudaEventElapsedTime(&t->total_ms, b->tm.ev[0], b->tm.ev[3]);
    t->n_launches = total ? 3 : 0;
    t->work_units = (double)*(unsigned long long *)(b->h_ctl + 4);
  }
  if (pos) *pos = b->h_pos;
  if (pos_off) *pos_off = b->h_poff;
  if (n_pos) *n_pos = total;
  return 0;
}

int bwaflow_sa_lookup(bwaflow_ctx_t *ctx, int64_t n, const uint64_t *k, uint64_t *out)
{
  BF_REQUIRE(ctx && ctx->has_index && (n == 0 || (k && out)), -22, "bwaflow_sa_lookup: bad argument");
  BF_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  uint64_t *dk = 0, *dout = 0; unsigned long long *db = 0;
  BF_CUDA(cudaMalloc(&dk, n * 8)); BF_CUDA(cudaMalloc(&dout, n * 8)); BF_CUDA(cudaMalloc(&db, 8));
  BF_CUDA(cudaMemset(db, 0, 8));
  BF_CUDA(cudaMemcpy(dk, k, n * 8, cudaMemcpyHostToDevice));
  sa_kernel<<<(unsigned)((n + 127) / 128), 128>>>(ctx->ix, n, dk, dout, db);
  BF_CUDA(cudaGetLastError());
  BF_CUDA(cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dk); cudaFree(dout); cudaFree(db);
  return 0;
}

} // extern "C"
