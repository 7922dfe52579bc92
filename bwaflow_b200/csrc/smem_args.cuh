// Arguments of the seeding kernels (smem.cu, smem_fast.cu).
#pragma once
#include "internal.cuh"

namespace bwaflow {

struct SmemArgs {
  const uint8_t *bases; const int64_t *read_off; int64_t n_reads;
  bwaflow_intv_t *slots; int cap;            // per-read output slot of `cap` intervals
  int *counts;                               // per-read interval count (or cap+1 on overflow)
  int *ctl;                                  // [0] queue head, [1] overflow flag, [2] too-long flag, [3] pass-3 queue head
  unsigned long long *blocks;                // occ-block touches (64 B each)
  const uint64_t *packed;                    // smem_fast.cu: 2-bit reads + N masks, one Row<MAXL> record per read
  int x_min;                                 // smem_fast.cu: lanes that must want an index extension before the warp runs one
  int *defer; int heavy_min;                 // smem_sweep.cu: [0] reads handed to the warp-per-read kernel, [1] its queue head, [2..] their ids; forward matches of one call from which a read is handed over
  DevIndex ix;
  int min_seed_len, split_len, split_width, max_mem_intv;
};

} // namespace bwaflow
