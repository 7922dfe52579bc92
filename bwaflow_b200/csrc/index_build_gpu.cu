// FM-index construction on the GPU: the same arrays as index_build.cpp / `bwa index` (bwa/bwtindex.c:256-324), built in
// HBM so that the 3.1 Gbp configuration (6.2 G suffixes) takes seconds instead of hours.  The BWT of a text is unique, so
// any correct suffix sort gives bit-identical arrays (tests compare against the host builder and against bwa's own).
//
//   1. text      forward + reverse complement, 2 bits per symbol, 32 symbols per u64 (MSB first), zero padded
//   2. partition all suffix positions into 16 buckets by their first two symbols (one counting pass, one scatter pass)
//   3. per bucket: key = the next 32 symbols, cub radix sort of (key, position) pairs; runs of equal keys (repeats) are finished
//      by one thread per run comparing the suffixes word by word ("$" sorts first = the shorter suffix first)
//   4. per bucket: BWT symbol of every row (text[pos-1]), the row of suffix 0 (primary), every 32nd suffix-array row
//   5. occurrence counts per 128 symbols (scan) and bwa's interleaved layout (bwt_bwtupdate_core bwa/bwtindex.c:151-173)
//
// Workload-preparation tool (index construction is `bwa index`, outside the timed path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>
#include "built_index.h"
#include "internal.cuh"

namespace bwaflow {
namespace {

constexpr int IDX_NB = 16;                 // buckets = first two symbols
constexpr int IDX_PREFIX = 2;

__device__ __forceinline__ uint64_t text32(const uint64_t *w, int64_t p)
{
  const int64_t i = p >> 5;
  const int sh = (int)(p & 31) << 1;
  return sh ? (w[i] << sh) | (w[i + 1] >> (64 - sh)) : w[i];
}
__device__ __forceinline__ int text_at(const uint64_t *w, int64_t p) { return (int)(w[p >> 5] >> ((~p & 31) << 1)) & 3; }
__device__ __forceinline__ int pac_at(const uint8_t *pac, int64_t p) { return (pac[p >> 2] >> ((~p & 3) << 1)) & 3; }

// one thread per text word
__global__ void idx_text_kernel(const uint8_t *pac, int64_t L, uint64_t *T, int64_t n_words)
{
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const int64_t n = 2 * L;
  uint64_t v = 0;
  for (int s = 0; s < 32; ++s) {
    const int64_t p = w * 32 + s;
    int c = 0;
    if (p < L) c = pac_at(pac, p);
    else if (p < n) c = 3 - pac_at(pac, n - 1 - p);
    v = v << 2 | (uint64_t)c;
  }
  T[w] = v;
}

// padded first-two-symbols bucket of suffix p (symbols past the end read as 0; the text is zero padded)
__device__ __forceinline__ int bucket_of(const uint64_t *T, int64_t p) { return (int)(text32(T, p) >> (64 - 2 * IDX_PREFIX)); }

constexpr int PART_THREADS = 256, PART_PER = 16;

__global__ void idx_hist_kernel(const uint64_t *T, int64_t n, unsigned long long *hist)
{
  __shared__ unsigned int sh[IDX_NB];
  if (threadIdx.x < IDX_NB) sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = ((int64_t)blockIdx.x * PART_THREADS + threadIdx.x) * PART_PER;
  uint64_t c0 = 0, c1 = 0;                                 // 16 one-byte counters
  for (int k = 0; k < PART_PER; ++k)
    if (base + k < n) {
      const int b = bucket_of(T, base + k);
      if (b < 8) c0 += 1ull << (b * 8); else c1 += 1ull << ((b - 8) * 8);
    }
  for (int b = 0; b < IDX_NB; ++b) {
    const unsigned v = __reduce_add_sync(0xffffffffu, (unsigned)((b < 8 ? c0 >> (b * 8) : c1 >> ((b - 8) * 8)) & 0xff));
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[b], v);
  }
  __syncthreads();
  if (threadIdx.x < IDX_NB && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// scatter every position into its bucket's range of pos[] (order inside a bucket is irrelevant: it is sorted next)
__global__ void idx_scatter_kernel(const uint64_t *T, int64_t n, unsigned long long *cursor, uint64_t *pos)
{
  __shared__ unsigned int cnt[IDX_NB];
  __shared__ unsigned long long start[IDX_NB];
  if (threadIdx.x < IDX_NB) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = ((int64_t)blockIdx.x * PART_THREADS + threadIdx.x) * PART_PER;
  unsigned int slot[PART_PER];
  int bk[PART_PER];
  for (int k = 0; k < PART_PER; ++k)
    if (base + k < n) { bk[k] = bucket_of(T, base + k); slot[k] = atomicAdd(&cnt[bk[k]], 1u); }
  __syncthreads();
  if (threadIdx.x < IDX_NB && cnt[threadIdx.x]) start[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
  __syncthreads();
  for (int k = 0; k < PART_PER; ++k)
    if (base + k < n) pos[start[bk[k]] + slot[k]] = (uint64_t)(base + k);
}

__global__ void idx_key_kernel(const uint64_t *T, const uint64_t *pos, int64_t m, uint64_t *keys)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) keys[i] = text32(T, (int64_t)pos[i] + IDX_PREFIX);
}

// suffix a < suffix b by the padded text, ties (equal up to both ends) = shorter suffix first
__device__ bool idx_suf_less(const uint64_t *T, int64_t n, int64_t a, int64_t b)
{
  for (int64_t off = 0;; off += 32) {
    if (a + off >= n || b + off >= n) {
      // at least one has ended: the rest of it is padding (zeros)
      const uint64_t wa = a + off < n ? text32(T, a + off) : 0, wb = b + off < n ? text32(T, b + off) : 0;
      if (wa != wb) return wa < wb;
      if (a + off + 32 >= n && b + off + 32 >= n) return a > b;
      continue;
    }
    const uint64_t wa = text32(T, a + off), wb = text32(T, b + off);
    if (wa != wb) return wa < wb;
  }
}

// one thread per run of equal keys: heap sort of its positions with the full comparison
__global__ void idx_ties_kernel(const uint64_t *T, int64_t n, const uint64_t *keys, uint64_t *pos, int64_t m)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  if (i > 0 && keys[i - 1] == keys[i]) return;             // not a run start
  if (i + 1 >= m || keys[i + 1] != keys[i]) return;        // run of one
  int64_t len = 2;
  while (i + len < m && keys[i + len] == keys[i]) ++len;
  uint64_t *a = pos + i;
  auto less = [&](uint64_t x, uint64_t y) { return idx_suf_less(T, n, (int64_t)x, (int64_t)y); };
  auto sift = [&](int64_t root, int64_t end) {
    for (;;) {
      int64_t child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && less(a[child], a[child + 1])) ++child;
      if (!less(a[root], a[child])) break;
      const uint64_t t = a[root]; a[root] = a[child]; a[child] = t;
      root = child;
    }
  };
  for (int64_t s = len / 2 - 1; s >= 0; --s) sift(s, len);
  for (int64_t e = len - 1; e > 0; --e) {
    const uint64_t t = a[0]; a[0] = a[e]; a[e] = t;
    sift(0, e);
  }
}

// rows[] = BWT symbol of every row of the (n+1)-row matrix, 16 per u32, MSB first.  This launch covers sorted suffixes
// g in [g0, g0+m) = rows g+1.  One thread per aligned group of 16 rows; groups shared with a neighbouring bucket use atomicOr.
__global__ void idx_emit_kernel(const uint64_t *T, const uint64_t *pos, int64_t g0, int64_t m, uint32_t *rows, uint64_t *sa,
                                unsigned long long *primary)
{
  const int64_t r_lo = g0 + 1, r_hi = g0 + 1 + m;           // rows of this bucket
  const int64_t w = (r_lo >> 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 16 >= r_hi) return;
  uint32_t v = 0;
  bool whole = true;
  for (int s = 0; s < 16; ++s) {
    const int64_t r = w * 16 + s;
    if (r < r_lo || r >= r_hi) { whole = false; continue; }
    const int64_t p = (int64_t)pos[r - r_lo];
    int c = 0;
    if (p == 0) *primary = (unsigned long long)r;
    else c = text_at(T, p - 1);
    v |= (uint32_t)c << ((15 - s) << 1);
    if ((r & 31) == 0) sa[r >> 5] = (uint64_t)p;
  }
  if (whole) rows[w] = v; else atomicOr(&rows[w], v);
}

// final BWT symbol i (the '$' row removed): row i before primary, row i+1 after
__device__ __forceinline__ int idx_sym(const uint32_t *rows, int64_t i, int64_t primary)
{
  const int64_t r = i < primary ? i : i + 1;
  return (int)(rows[r >> 4] >> ((~r & 15) << 1)) & 3;
}

// one thread per 128-symbol block: its four symbol counts
__global__ void idx_count_kernel(const uint32_t *rows, int64_t n, int64_t primary, int64_t n_blk, uint64_t *cnt)
{
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blk) return;
  uint64_t c[4] = {0, 0, 0, 0};
  const int64_t e = (b + 1) * 128 < n ? (b + 1) * 128 : n;
  for (int64_t i = b * 128; i < e; ++i) ++c[idx_sym(rows, i, primary)];
  for (int k = 0; k < 4; ++k) cnt[k * (n_blk + 1) + b] = c[k];
}

// bwa's layout: before every 128 symbols four running u64 counts, then the symbols (16 per u32); a last count block at the end
__global__ void idx_layout_kernel(const uint32_t *rows, int64_t n, int64_t primary, int64_t n_blk, const uint64_t *cum, uint32_t *bwt)
{
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blk) return;
  uint32_t *o = bwt + b * 16;
  if (b == n_blk && (n & 127) != 0) o = bwt + (n_blk - 1) * 16 + 8 + (((n & 127) + 15) >> 4);   // after a partial last block
  for (int k = 0; k < 4; ++k) {
    const uint64_t v = cum[k * (n_blk + 1) + b];
    o[2 * k] = (uint32_t)v; o[2 * k + 1] = (uint32_t)(v >> 32);
  }
  if (b == n_blk) return;
  const int64_t e = (b + 1) * 128 < n ? (b + 1) * 128 : n;
  for (int64_t i0 = b * 128, k = 0; i0 < e; i0 += 16, ++k) {
    uint32_t v = 0;
    for (int s = 0; s < 16 && i0 + s < n; ++s) v |= (uint32_t)idx_sym(rows, i0 + s, primary) << ((15 - s) << 1);
    o[8 + k] = v;
  }
}

struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { BF_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); return 0; }
  template <class X> X *as() { return (X *)p; }
};

} // namespace
} // namespace bwaflow

using namespace bwaflow;

extern "C" int bwaflow_index_build_gpu(int device, const uint8_t *codes, int64_t n_bases, int32_t n_contigs,
                                       const int64_t *contig_len, bwaflow_built_index_t **out)
{
  BF_REQUIRE(codes && out && n_bases > 0 && n_contigs > 0 && contig_len, -22, "bwaflow_index_build_gpu: bad argument");
  BF_CUDA(cudaSetDevice(device));
  bwaflow_built_index *B = new bwaflow_built_index();
  struct Guard { bwaflow_built_index *b; ~Guard() { delete b; } } guard{B};
  int64_t tot = 0;
  for (int i = 0; i < n_contigs; ++i) { B->ctg_offset.push_back(tot); B->ctg_len.push_back((int32_t)contig_len[i]); tot += contig_len[i]; }
  BF_REQUIRE(tot == n_bases, -22, "bwaflow_index_build_gpu: contig lengths do not add up to n_bases");
  const int64_t L = n_bases, n = 2 * L;
  // pac on the host: N -> lrand48() & 3 in reference order after srand48(11)   (bwa/bntseq.c:260-262, 289-290)
  B->pac.assign(L / 4 + 1, 0);
  srand48(11);
  for (int64_t p = 0; p < L; ++p) {
    int c = codes[p];
    if (c >= 4) c = (int)(lrand48() & 3);
    B->pac[p >> 2] |= (uint8_t)(c << ((~p & 3) << 1));
  }
  cudaStream_t st = 0;
  DevBuf d_pac, d_T, d_pos, d_keys, d_keys2, d_pos2, d_tmp, d_rows, d_sa, d_ctl, d_cnt, d_bwt;
  const int64_t n_words = (n + 31) / 32 + 4;
  if (d_pac.alloc(B->pac.size()) || d_T.alloc(n_words * 8) || d_pos.alloc(n * 8) || d_ctl.alloc(64 * 8)) return -5;
  BF_CUDA(cudaMemcpyAsync(d_pac.p, B->pac.data(), B->pac.size(), cudaMemcpyHostToDevice, st));
  uint64_t *T = d_T.as<uint64_t>();
  idx_text_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(d_pac.as<uint8_t>(), L, T, n_words);
  // bucket sizes
  unsigned long long *hist = d_ctl.as<unsigned long long>(), *cursor = hist + IDX_NB, *d_primary = hist + 2 * IDX_NB;
  BF_CUDA(cudaMemsetAsync(d_ctl.p, 0, 64 * 8, st));
  const unsigned part_blocks = (unsigned)((n + (int64_t)PART_THREADS * PART_PER - 1) / ((int64_t)PART_THREADS * PART_PER));
  idx_hist_kernel<<<part_blocks, PART_THREADS, 0, st>>>(T, n, hist);
  unsigned long long h_hist[IDX_NB], h_start[IDX_NB + 1];
  BF_CUDA(cudaMemcpyAsync(h_hist, hist, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  h_start[0] = 0;
  int64_t m_max = 0;
  for (int b = 0; b < IDX_NB; ++b) { h_start[b + 1] = h_start[b] + h_hist[b]; if ((int64_t)h_hist[b] > m_max) m_max = (int64_t)h_hist[b]; }
  BF_REQUIRE((int64_t)h_start[IDX_NB] == n, -5, "bwaflow_index_build_gpu: bucket histogram does not add up");
  BF_REQUIRE(m_max < (1LL << 31), -34, "bwaflow_index_build_gpu: a first-two-symbols bucket holds 2^31 suffixes or more");
  BF_CUDA(cudaMemcpyAsync(cursor, h_start, IDX_NB * 8, cudaMemcpyHostToDevice, st));
  idx_scatter_kernel<<<part_blocks, PART_THREADS, 0, st>>>(T, n, cursor, d_pos.as<uint64_t>());
  // per-bucket sort
  size_t tmp_bytes = 0;
  {
    cub::DoubleBuffer<uint64_t> k(nullptr, nullptr), v(nullptr, nullptr);
    BF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k, v, (int)m_max, 0, 64, st));
  }
  const int64_t n_rows_words = (n + 1 + 15) / 16 + 1;
  const uint64_t n_sa = (uint64_t)(n + 32) / 32;
  if (d_keys.alloc(m_max * 8) || d_keys2.alloc(m_max * 8) || d_pos2.alloc(m_max * 8) || d_tmp.alloc(tmp_bytes) ||
      d_rows.alloc(n_rows_words * 4) || d_sa.alloc(n_sa * 8)) return -5;
  BF_CUDA(cudaMemsetAsync(d_rows.p, 0, n_rows_words * 4, st));
  uint32_t *rows = d_rows.as<uint32_t>();
  for (int b = 0; b < IDX_NB; ++b) {
    const int64_t m = (int64_t)h_hist[b], g0 = (int64_t)h_start[b];
    if (!m) continue;
    uint64_t *pos = d_pos.as<uint64_t>() + g0;
    idx_key_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(T, pos, m, d_keys.as<uint64_t>());
    cub::DoubleBuffer<uint64_t> k(d_keys.as<uint64_t>(), d_keys2.as<uint64_t>()), v(pos, d_pos2.as<uint64_t>());
    size_t tb = tmp_bytes;
    BF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, k, v, (int)m, 0, 64, st));
    if (v.Current() != pos) BF_CUDA(cudaMemcpyAsync(pos, v.Current(), m * 8, cudaMemcpyDeviceToDevice, st));
    idx_ties_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(T, n, k.Current(), pos, m);
    const int64_t n_w = ((g0 + 1 + m + 15) >> 4) - ((g0 + 1) >> 4);
    idx_emit_kernel<<<(unsigned)((n_w + 255) / 256), 256, 0, st>>>(T, pos, g0, m, rows, d_sa.as<uint64_t>(), d_primary);
  }
  BF_CUDA(cudaGetLastError());
  unsigned long long h_primary = 0;
  BF_CUDA(cudaMemcpyAsync(&h_primary, d_primary, 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  BF_REQUIRE(h_primary > 0, -5, "bwaflow_index_build_gpu: suffix 0 not found");
  // the sort buffers are no longer needed
  cudaFree(d_pos.p); d_pos.p = nullptr; cudaFree(d_pos2.p); d_pos2.p = nullptr;
  cudaFree(d_keys.p); d_keys.p = nullptr; cudaFree(d_keys2.p); d_keys2.p = nullptr;
  // row 0 = the '$' suffix, preceded by the last symbol
  {
    const uint32_t c = (uint32_t)(3 - ((B->pac[0] >> 6) & 3));      // last symbol of the text = complement of the first base
    uint32_t w0;
    BF_CUDA(cudaMemcpy(&w0, rows, 4, cudaMemcpyDeviceToHost));
    w0 |= c << 30;
    BF_CUDA(cudaMemcpy(rows, &w0, 4, cudaMemcpyHostToDevice));
  }
  const int64_t n_blk = (n + 127) / 128;
  const uint64_t n_occ = (uint64_t)n_blk + 1;
  const uint64_t bwt_size = (uint64_t)((n + 15) >> 4) + n_occ * 8;
  if (d_cnt.alloc(4 * (n_blk + 1) * 8) || d_bwt.alloc(bwt_size * 4)) return -5;
  BF_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4 * (n_blk + 1) * 8, st));
  idx_count_kernel<<<(unsigned)((n_blk + 255) / 256), 256, 0, st>>>(rows, n, (int64_t)h_primary, n_blk, d_cnt.as<uint64_t>());
  {
    size_t sb = 0;
    BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, d_cnt.as<uint64_t>(), d_cnt.as<uint64_t>(), (int)(n_blk + 1), st));
    DevBuf d_scan;
    if (d_scan.alloc(sb)) return -5;
    for (int k = 0; k < 4; ++k) {
      uint64_t *c = d_cnt.as<uint64_t>() + (size_t)k * (n_blk + 1);
      BF_CUDA(cub::DeviceScan::ExclusiveSum(d_scan.p, sb, c, c, (int)(n_blk + 1), st));
    }
    BF_CUDA(cudaStreamSynchronize(st));
  }
  idx_layout_kernel<<<(unsigned)((n_blk + 1 + 255) / 256), 256, 0, st>>>(rows, n, (int64_t)h_primary, n_blk, d_cnt.as<uint64_t>(), d_bwt.as<uint32_t>());
  BF_CUDA(cudaGetLastError());
  B->bwt.resize(bwt_size);
  B->sa.resize(n_sa);
  BF_CUDA(cudaMemcpyAsync(B->bwt.data(), d_bwt.p, bwt_size * 4, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(B->sa.data(), d_sa.p, n_sa * 8, cudaMemcpyDeviceToHost, st));
  uint64_t tot_cnt[4];
  for (int k = 0; k < 4; ++k)
    BF_CUDA(cudaMemcpyAsync(&tot_cnt[k], d_cnt.as<uint64_t>() + (size_t)k * (n_blk + 1) + n_blk, 8, cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  B->sa[0] = (uint64_t)-1;
  bwaflow_index_t &v = B->view;
  memset(&v, 0, sizeof(v));
  v.primary = h_primary;
  v.L2[0] = 0;
  for (int c = 0; c < 4; ++c) v.L2[c + 1] = v.L2[c] + tot_cnt[c];     // the BWT is a permutation of the text
  v.seq_len = (uint64_t)n; v.bwt_size = bwt_size; v.bwt = B->bwt.data();
  v.n_sa = n_sa; v.sa_intv = 32; v.sa = B->sa.data();
  v.l_pac = L; v.pac = B->pac.data(); v.n_seqs = n_contigs; v._pad = 0;
  v.ctg_offset = B->ctg_offset.data(); v.ctg_len = B->ctg_len.data();
  guard.b = nullptr;
  *out = B;
  return 0;
}
