// FM-index device primitives shared by the seeding kernels (smem.cu, smem_fast.cu): occurrence lookups on the occ32
// layout (internal.cuh), bwt_extend for the one consumed base (bwa/bwt.c:262-275), bwt_set_intv (bwa/bwt.h:80),
// bwt_invPsi / bwt_sa (bwa/bwt.c:53-59,86-96) and the 16-byte packed interval entry.
#pragma once
#include "internal.cuh"

namespace bwaflow {

struct Intv { uint64_t x0, x1, x2; };

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, int sh)
{ // PTX shl clamps shift amounts above 31 (result 0); C++ << would be undefined there
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(sh));
  return r;
}

// Among the first n symbols (1..64) of an occ32 block, whose 64 two-bit symbols are the four 32-bit words of `sy`
// (16 symbols each, MSB first): how many equal base c (*eq) and how many are greater than c (*gt).  32-bit
// single-instruction logic; the per-word prefix mask is one clamped shift.  POPC runs on the narrow XU pipe, so the
// masks of two words (which only use even bit positions) are interleaved before counting: 2+2 popcounts.
__device__ __forceinline__ void prefix_counts(const uint4 sy, int n, int c, int *eq, int *gt)
{
  const uint32_t w[4] = {sy.x, sy.y, sy.z, sy.w};
  const uint32_t ODD = 0x55555555u;
  const uint32_t cpat = (uint32_t)c * ODD;                       // c in every 2-bit field
  const uint32_t g_or = c == 0 ? ~0u : 0u, g_hi = c == 1 ? ~0u : 0u, g_and = c == 2 ? ~0u : 0u;
  const int base = 32 - 2 * n;                                   // word j keeps its top 2*clamp(n-16j,0,16) bits
  uint32_t e[4], g[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t keep = shl_clamp(ODD, max(base + 32 * j, 0));   // even bit of every kept symbol
    const uint32_t x = w[j] ^ cpat, lo = w[j], hi = w[j] >> 1;
    e[j] = ~(x | x >> 1) & keep;
    g[j] = (((hi | lo) & g_or) | (hi & g_hi) | (hi & lo & g_and)) & keep;
  }
  *eq = __popc(e[0] | e[1] << 1) + __popc(e[2] | e[3] << 1);
  *gt = __popc(g[0] | g[1] << 1) + __popc(g[2] | g[3] << 1);
}

// Occ(c, k) and Occ(>c, k) over BWT rows [0, k] (k already adjusted for the '$' row): one 32-byte occ32 block.
// Header S1,S2,S3 = symbols >0,>1,>2 before the block (relative to its superblock); S0 = the block start, S4 = 0:
//   Occ(>c) = S[c+1] + gt,   Occ(c) = S[c] - S[c+1] + eq.
__device__ __forceinline__ void occ_eq_gt(const DevIndex &ix, uint64_t k, int c, uint64_t *eq, uint64_t *gt)
{
  const uint4 *b = ix.occ + ((k >> 6) << 1);
  const uint4 hd = occ_ld(ix, b), sy = occ_ld(ix, b + 1);
  int e, g;
  prefix_counts(sy, (int)(k & 63) + 1, c, &e, &g);
  const uint64_t *sb = ix.sbt + (k >> 31) * 5;
  const uint32_t r0 = (uint32_t)(k & 0x7fffffffull) & ~63u;    // block start relative to the superblock
  const uint32_t lo_rel = c == 0 ? r0 : c == 1 ? hd.x : c == 2 ? hd.y : hd.z;        // S[c]
  const uint32_t hi_rel = c == 0 ? hd.x : c == 1 ? hd.y : c == 2 ? hd.z : 0u;        // S[c+1]
  const uint64_t Sc = sb[c] + lo_rel, Sc1 = sb[c + 1] + hi_rel;
  *gt = Sc1 + (uint64_t)g;
  *eq = Sc - Sc1 + (uint64_t)e;
}

// ok[c] of bwt_extend (bwa/bwt.c:262-275) for the one base c the caller wants: Occ of c and of all bases > c at both
// ends of the interval (the part of bwt_2occ4, bwa/bwt.c:189-220, that is consumed).  Straight-line code: all lanes of
// the warp run the two lookups together.  `blocks` counts touches of the REFERENCE's 64-byte blocks (the algorithmic
// unit of SURVEY 8d: one if both ends share a 128-symbol block, else two), not of our 32-byte device blocks.
__device__ __forceinline__ Intv extend1(const DevIndex &ix, const Intv ik, int is_back, int c, unsigned long long *blocks)
{
  const uint64_t xf = is_back ? ik.x0 : ik.x1, xg = is_back ? ik.x1 : ik.x0;
  const uint64_t k = xf - 1, l = xf - 1 + ik.x2;
  const uint64_t _k = k - (k >= ix.primary), _l = l - (l >= ix.primary);
  uint64_t ek, gk, el, gl;
  occ_eq_gt(ix, _k, c, &ek, &gk);
  occ_eq_gt(ix, _l, c, &el, &gl);
  if (!(ix.occ_cs & 2)) *blocks += (_k >> 7) != (_l >> 7) ? 2 : 1;   // (bit 1 = developer switch: count text-anchored steps only)
  Intv o;
  const uint64_t acc = xg + (xf <= ix.primary && xf + ik.x2 - 1 >= ix.primary) + (gl - gk);
  const uint64_t nf = ix.L2[c] + 1 + ek;
  o.x2 = el - ek;
  if (is_back) { o.x0 = nf; o.x1 = acc; } else { o.x1 = nf; o.x0 = acc; }
  return o;
}

__device__ __forceinline__ Intv set_intv(const DevIndex &ix, int c)
{ // bwt_set_intv, bwa/bwt.h:80
  Intv ik; ik.x0 = ix.L2[c] + 1; ik.x2 = ix.L2[c + 1] - ix.L2[c]; ik.x1 = ix.L2[3 - c] + 1; return ik;
}

// One list entry of bwt_smem1a_new packed into 16 bytes: three 39-bit interval fields and an 11-bit query coordinate
// (seq_len < 2^39, reads < 2^11); a single 128-bit local-memory access instead of four.
__device__ __forceinline__ uint4 pack_entry(const Intv &v, int pos)
{
  const uint64_t lo = v.x0 | v.x1 << 39, hi = v.x1 >> 25 | v.x2 << 14 | (uint64_t)pos << 53;
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}
__device__ __forceinline__ Intv unpack_entry(const uint4 e, int *pos)
{
  const uint64_t lo = (uint64_t)e.y << 32 | e.x, hi = (uint64_t)e.w << 32 | e.z, m = (1ull << 39) - 1;
  Intv v; v.x0 = lo & m; v.x1 = (lo >> 39 | hi << 25) & m; v.x2 = (hi >> 14) & m; *pos = (int)(hi >> 53);
  return v;
}
__device__ __forceinline__ int entry_pos(const uint4 e) { return (int)(e.w >> 21); }

// bwt_sa, bwa/bwt.c:86-96 with bwt_invPsi :53-59 (symbol and its Occ come from the same occ32 block)
// bwt_invPsi (bwa/bwt.c:53-59): the row of the suffix one position to the left; symbol and Occ come from the same occ32 block.
__device__ __forceinline__ uint64_t inv_psi(const DevIndex &ix, uint64_t k)
{
  if (k == ix.primary) return 0;
  const uint64_t x = k - (k > ix.primary);
  const uint4 *b = ix.occ + ((x >> 6) << 1);
  const uint4 hd = occ_ld(ix, b), sy = occ_ld(ix, b + 1);
  const int r = (int)(x & 63);
  const uint32_t word = (r >> 4) == 0 ? sy.x : (r >> 4) == 1 ? sy.y : (r >> 4) == 2 ? sy.z : sy.w;
  const int c = (word >> ((~r & 15) << 1)) & 3;
  int e, g;
  prefix_counts(sy, r + 1, c, &e, &g);
  const uint64_t *sb = ix.sbt + (x >> 31) * 5;
  const uint32_t r0 = (uint32_t)(x & 0x7fffffffull) & ~63u;
  const uint32_t lo_rel = c == 0 ? r0 : c == 1 ? hd.x : c == 2 ? hd.y : hd.z;
  const uint32_t hi_rel = c == 0 ? hd.x : c == 1 ? hd.y : c == 2 ? hd.z : 0u;
  return ix.L2[c] + (sb[c] + lo_rel) - (sb[c + 1] + hi_rel) + (uint64_t)e;
}

// bwt_sa (bwa/bwt.c:86-96) over the device-side sampling (every 2^sa_shift-th row; bwa_ctx_create may have made it denser than
// bwa's 32, which does not change any value).  `blocks` counts the occ-block touches actually made.
__device__ __forceinline__ uint64_t sa_lookup(const DevIndex &ix, uint64_t k, unsigned long long *blocks)
{
  uint64_t steps = 0;
  const uint64_t mask = (1ull << ix.sa_shift) - 1;
  while (k & mask) {
    ++steps;
    k = inv_psi(ix, k);
  }
  *blocks += steps;
  return steps + ix.sa[k >> ix.sa_shift];
}

// Text-anchored form of an interval that has become unique (x2 == 1): the pattern occurs exactly once in the
// forward + reverse-complement text, at position x0 here; x1 holds a marker no row can equal (fields are 39 bits).
// Extending such an interval needs no index walk: the next base either equals the text base next to the occurrence or
// the extension fails (bwt_extend on a one-row interval computes exactly that), and its rows are recovered on demand from
// the inverse suffix array: x[0] = isa[pos], x[1] = isa[n - pos - length] (the reverse complement occurs at the mirrored
// position).  Only available with the full suffix array in HBM (DevIndex::isa).
constexpr uint64_t TEXT_MARK = (1ull << 39) - 1;
__device__ __forceinline__ bool is_text(const Intv &v) { return v.x1 == TEXT_MARK; }
__device__ __forceinline__ Intv text_intv(uint64_t tpos) { Intv v; v.x0 = tpos; v.x1 = TEXT_MARK; v.x2 = 1; return v; }

} // namespace bwaflow
