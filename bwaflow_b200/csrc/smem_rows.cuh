// Packed read rows of the fast seeding kernels (smem_fast.cu, smem_sweep.cu): 2-bit codes + N mask per read, and the
// level-table addressing.
#pragma once
#include "internal.cuh"

namespace bwaflow {

template <int MAXL> struct Row {
  static constexpr int W = MAXL / 32;                       // 2-bit words holding bases
  static constexpr int QW = W + 1;                          // + one zero word so a 32-base window never runs off the row
  static constexpr int PACK_BYTES = QW * 8 + (W * 4 + 7) / 8 * 8;
  static constexpr int ROW_BYTES = PACK_BYTES + 64;         // + forward-list end coordinates; an odd number of 8-byte words
  static_assert((ROW_BYTES / 8) % 2 == 1, "row stride must be an odd number of 64-bit words");
};

// One warp per read: 2-bit codes (MSB first, N as 0) and an N bit mask in the Row<MAXL> layout.
template <int MAXL>
__global__ void pack_reads_kernel(const uint8_t *bases, const int64_t *read_off, int64_t n_reads, uint64_t *packed)
{
  using R = Row<MAXL>;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_reads) return;
  const int64_t ro = read_off[r];
  const int len = (int)(read_off[r + 1] - ro);
  uint32_t p16 = 0, n8 = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int pos = lane * 8 + t;
    uint32_t b = pos < len && pos < MAXL ? bases[ro + pos] : 0;
    if (b > 3) { n8 |= 1u << t; b = 0; }
    p16 = p16 << 2 | b;
  }
  uint64_t piece = (uint64_t)p16 << (48 - 16 * (lane & 3));
  piece |= __shfl_xor_sync(0xffffffffu, piece, 1); piece |= __shfl_xor_sync(0xffffffffu, piece, 2);
  uint32_t np = n8 << (8 * (lane & 3));
  np |= __shfl_xor_sync(0xffffffffu, np, 1); np |= __shfl_xor_sync(0xffffffffu, np, 2);
  uint64_t *row = packed + r * (R::PACK_BYTES / 8);
  uint32_t *nm = (uint32_t *)(row + R::QW);
  if ((lane & 3) == 0 && (lane >> 2) < R::W) row[lane >> 2] = piece;
  if ((lane & 3) == 1 && (lane >> 2) < R::W) nm[lane >> 2] = np;
  if (lane == 2) row[R::W] = 0;
  if (lane == 3 && (R::W & 1)) nm[R::W] = 0;
}

__host__ __device__ __forceinline__ uint64_t ktab_base(int L) { return ((1ull << (2 * L)) - 4) / 3; }     // entries of levels 1..L-1

// The lane's view of its read row in shared memory.
template <int MAXL> struct LaneRead {
  using R = Row<MAXL>;
  uint64_t *q2; const uint32_t *nm; int len; bool anyN;
  __device__ __forceinline__ void bind(uint64_t *row) { q2 = row; nm = (const uint32_t *)(row + R::QW); }
  __device__ __forceinline__ void load(const uint64_t *packed, int64_t r, int len_)
  {
    const uint64_t *src = packed + r * (R::PACK_BYTES / 8);
#pragma unroll
    for (int t = 0; t < R::PACK_BYTES / 8; ++t) q2[t] = __ldg(src + t);
    uint32_t any = 0;
#pragma unroll
    for (int t = 0; t < R::W; ++t) any |= nm[t];
    anyN = any != 0; len = len_;
  }
  __device__ __forceinline__ int base(int i) const { return (int)(q2[i >> 5] >> ((~i & 31) << 1)) & 3; }
  __device__ __forceinline__ bool isN(int i) const { return anyN && ((nm[i >> 5] >> (i & 31)) & 1u); }
  __device__ __forceinline__ uint64_t win(int i) const                    // 32 bases from i, MSB first
  {
    const int s = (i & 31) << 1; const uint64_t w0 = q2[i >> 5], w1 = q2[(i >> 5) + 1];
    return s ? (w0 << s | w1 >> (64 - s)) : w0;
  }
  __device__ __forceinline__ uint64_t win_end(int i) const                // bases [i-32, i), the last one at bit 0
  { return i >= 32 ? win(i - 32) : (i > 0 ? q2[0] >> (64 - 2 * i) : 0ull); }
  __device__ __forceinline__ uint64_t code(int b, int L) const { return win(b) >> (64 - 2 * L); }   // bases [b, b+L), 1 <= L <= 32
  __device__ int next_n(int i) const                                      // first N at or after i, else len
  {
    if (!anyN) return len;
    for (int w = i >> 5; w < R::W; ++w) {
      uint32_t m = nm[w];
      if (w == (i >> 5)) m &= ~0u << (i & 31);
      if (m) return min(len, w * 32 + __ffs(m) - 1);
    }
    return len;
  }
  __device__ int prev_n(int i) const                                      // last N before i, else -1
  {
    if (!anyN || i <= 0) return -1;
    for (int w = (i - 1) >> 5; w >= 0; --w) {
      uint32_t m = nm[w];
      if (w == ((i - 1) >> 5)) m &= 0xffffffffu >> (31 - ((i - 1) & 31));
      if (m) return w * 32 + 31 - __clz(m);
    }
    return -1;
  }
};

// 32 text symbols from position p (MSB first) / the 32 symbols before p (the last one at bit 0); the text is zero padded.
__device__ __forceinline__ uint64_t text_win(const uint64_t *text, uint64_t p)
{
  const int s = (int)(p & 31) << 1; const uint64_t w0 = __ldg(text + (p >> 5)), w1 = __ldg(text + (p >> 5) + 1);
  return s ? (w0 << s | w1 >> (64 - s)) : w0;
}
__device__ __forceinline__ uint64_t text_win_end(const uint64_t *text, uint64_t p)
{ return p >= 32 ? text_win(text, p - 32) : (p > 0 ? __ldg(text) >> (64 - 2 * (int)p) : 0ull); }

} // namespace bwaflow
