// Shared internals of libbwaflow_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/bwaflow_cuda.h"

namespace bwaflow {

void set_error(const char *fmt, ...);

#define BF_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      bwaflow::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return -5; /* -EIO */                                                                    \
    }                                                                                          \
  } while (0)

#define BF_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      bwaflow::set_error(__VA_ARGS__);   \
      return (code);                     \
    }                                    \
  } while (0)

// Index as the kernels see it (all pointers are device pointers into HBM).
struct DevIndex {
  uint64_t primary, seq_len;
  uint64_t L2[5];
  // Device-side occurrence array ("occ32"): one 32-byte block per 64 BWT symbols =
  //   u32 S1,S2,S3 (symbols >0, >1, >2 before the block, relative to its 2^31-symbol superblock), u32 pad, 64 two-bit symbols.
  // Same information as bwa's 64-byte/128-symbol blocks (bwa/bwt.h:72-73), re-laid out at upload: one 32-byte sector per
  // lookup and half the symbols to count.  sbt[j*5+t] = absolute count of symbols >= t before superblock j (t=0: its start).
  const uint4 *occ;
  int occ_cs;              // load occ blocks with the streaming (evict-first) policy
  uint64_t sbt[8 * 5];
  const uint64_t *sa;
  int sa_shift;            // log2(sa_intv)
  // Level tables of the seeding kernels (smem_fast.cu): the bidirectional interval of every pattern of 1..ktab_k bases as a
  // 16-byte packed entry (fmindex.cuh), level L at entry offset (4^L - 4) / 3, code = the pattern's bases MSB first.
  const uint4 *ktab; int ktab_k;
  // Forward + reverse-complement text, 32 two-bit symbols per u64 (MSB first), zero padded: lets the seeding kernel extend
  // an interval that has become unique by comparing bases (smem_fast.cu); used only with the full suffix array (sa_shift == 0).
  const uint64_t *text;
  int64_t l_pac;
  const uint8_t *pac;
  int n_seqs;
  const int64_t *ctg_offset;
  const int32_t *ctg_len;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint4 occ_ld(const DevIndex &ix, const uint4 *p) { return (ix.occ_cs & 1) ? __ldcs(p) : __ldg(p); }
#endif

// Scoring as the extension kernels see it.  Matrix row t (target base) as 5 packed int8: lo = q0..q3, hi = q4.
struct DevScore {
  int o_del, e_del, o_ins, e_ins, zdrop, smax;
  uint32_t row_lo[5], row_hi[5];
};

struct Params {
  bwaflow_opt_t opt;
  DevScore sc;
};

} // namespace bwaflow

struct bwaflow_ctx {
  int device;
  int n_sm;
  bool has_index;
  bwaflow::DevIndex ix;
  bwaflow::Params par;
  // owned device allocations
  void *d_bwt, *d_sa, *d_pac, *d_ctg_off, *d_ctg_len, *d_text, *d_ktab;
  // host copies of the contig table (bns->anns[].offset / .len) for the host halves of the pairing stages
  int64_t *h_ctg_off; int32_t *h_ctg_len;
};

namespace bwaflow {
// what sam.cu needs to know about a CIGAR batch (cigar.cu)
int64_t cigar_max_jobs(const bwaflow_cigar_batch_t *b);
bwaflow_ctx_t *cigar_ctx(const bwaflow_cigar_batch_t *b);
const bwaflow_pe_opt_t *cigar_po(const bwaflow_cigar_batch_t *b);
void fill_score(const bwaflow_opt_t &o, DevScore &s);

struct Timer {
  cudaEvent_t ev[4];
  bool ok = false;
  int init() {
    for (int i = 0; i < 4; ++i) BF_CUDA(cudaEventCreate(&ev[i]));
    ok = true; return 0;
  }
  void destroy() { if (ok) for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]); ok = false; }
};
} // namespace bwaflow
