// Batch objects of the device stages (shared by smem.cu, chain2aln.cu, chain_dev.cu and hotpath.cu).
#pragma once
#include "internal.cuh"

namespace bwaflow {
constexpr int C2A_NCLASS = 8;
}

struct bwaflow_smem_batch {
  bwaflow_ctx_t *ctx;
  cudaStream_t stream;
  bwaflow::Timer tm;
  int64_t max_reads, max_bases, n_reads, max_len;
  uint8_t *d_bases; int64_t *d_read_off;
  bwaflow_intv_t *d_slots; int cap; int64_t slots_cap;
  int *d_counts; int64_t *d_cnt64, *d_off, *h_off;
  bwaflow_intv_t *d_out, *h_out; int64_t out_cap;
  int *d_ctl, *h_ctl;               // head | slot overflow | too-long | pass-3 head | blocks(u64) | capacity flags | pad
  int64_t *h_stat;                  // {n_intv, n_pos} of the last device-driven run (hot path)
  void *d_tmp; size_t tmp_bytes;
  int64_t n_intv;
  // seed positions
  int64_t *d_pcnt, *d_poff, *h_poff, *d_pos, *h_pos; int64_t pos_cap, hpos_cap;
  int blocks_per_launch;
  // smem_fast.cu
  uint64_t *d_packed; int fast_blocks_per_sm, sweep_blocks_per_sm, x_min;
  int *d_defer; int heavy_min;      // smem_sweep.cu: reads handed to the warp-per-read kernel
  bool want_rows;                   // intervals leave the device as BWT rows (interval API): no text-anchored entries
  bool count_blocks;                // this launch counts the reference's occ-block touches (exact instances)
  bool pending;
};

struct bwaflow_c2a_batch {
  bwaflow_ctx_t *ctx;
  cudaStream_t stream, own_stream;   // own_stream: saved while the fused hot path runs this batch on its stream
  bwaflow::Timer tm;
  int64_t max_reads, max_bases, max_chains, max_seeds;
  int64_t n_reads, n_chains, n_seeds;   // n_seeds = sum of chains[].n_seeds (seeds actually referenced)
  uint8_t *d_bases; int64_t *d_read_off, *d_chain_off;
  bwaflow_chain_t *d_chains; bwaflow_seed_t *d_seeds, *d_sseeds;
  int64_t *d_nseeds, *d_sbase, *d_chain_read, *d_win;
  uint64_t *d_skeys[2]; uint32_t *d_svals[2];
  uint32_t *d_rkeys[2], *d_rvals[2];
  void *d_tmp; size_t tmp_bytes;
  uint8_t *d_skipped;
  bwaflow_alnreg_t *d_regs, *d_regs_out, *h_regs;
  int *d_nregs; int64_t *d_nregs64, *d_reg_off, *h_reg_off;
  int *d_ctl, *h_ctl;               // hist[8] | queue[16] | err | pad | cells(u64)
  int blocks[bwaflow::C2A_NCLASS];
  int refill;
  bool pending;
  // speculative first-seed extension on the packed batch engine (chain2aln.cu)
  bool spec;                        // enabled for this batch
  bool count_cells;                 // this launch runs every read on the exact lane-per-read kernel and counts its DP cells
  double cells_known;
  bwaflow_ext_batch_t *xb;
  bwaflow_ext_result_t *d_resL;     // left-flank results per chain (right-flank results stay in the extension batch)
  int64_t *d_jsz, *d_joff;          // [2*(max_chains+1)] query / target bytes per chain job and their exclusive prefix
  void *d_jdesc;                    // per chain: 32-byte job descriptor of the current wave (chain2aln.cu SpecDesc)
  uint8_t *d_cflag;                 // per chain: 1 = no speculative result (pool overflow, unrepresentable job)
  uint8_t *d_strag;                 // per read: 1 = must be redone by the exact kernel
  uint8_t *d_skipped2;              // the replay's own copy of the per-seed "skipped" flags
  // second round of waves: every later seed of the reads the first replay could not finish (chain2aln.cu, "wave B")
  bwaflow_ext_batch_t *xb2;
  bwaflow_ext_result_t *d_resL2;    // per wave-B job
  int32_t *d_slot_chain;            // per seed slot (processing order): its chain
  int32_t *d_jslot, *d_slot_job;    // wave-B job -> seed slot and back (-1 = none)
  uint8_t *d_cand, *d_cflag2;       // per slot: candidate for wave B; per wave-B job: 1 = no result
  int64_t *d_jsz2, *d_joff2; void *d_jdesc2;
  int32_t *d_nsel; int *h_sel;      // {wave-B jobs, reads left by the first replay}
  int64_t waveb_jobs, waveb_reads;  // of the last launch
};

namespace bwaflow {
// ext.cu: an extension batch driven from the device side of another stage
struct ExtDeviceView { bwaflow_ext_job_t *jobs; uint8_t *q, *t; bwaflow_ext_result_t *out; int64_t max_jobs, max_q, max_t; };
ExtDeviceView ext_device_view(bwaflow_ext_batch_t *b);
int ext_launch_device(bwaflow_ext_batch_t *b, int64_t n_jobs, cudaStream_t st);
void ext_set_pair_classes(bwaflow_ext_batch_t *b, int n);   // qlen classes below n run on the packed engine
void ext_set_order_rows(bwaflow_ext_batch_t *b, int on);    // queue order by estimated lifetime first (see ext_key_kernel)
// sw.cu: a ksw_align2 batch driven from the device side of another stage (matesw.cu)
struct SwDeviceView { bwaflow_sw_job_t *jobs; uint8_t *q, *t; bwaflow_sw_result_t *out; int *d_ctl, *h_ctl; };
SwDeviceView sw_device_view(bwaflow_sw_batch_t *b);
int sw_launch_device(bwaflow_sw_batch_t *b, int64_t n_jobs, int max_qlen, int max_tlen, cudaStream_t st);
// smem_fast.cu
struct SmemArgs;
int smem_fast_launch(bwaflow_smem_batch *b, const SmemArgs &a, bool text, bool count);
int smem_fast_setup(bwaflow_smem_batch *b);
// smem_sweep.cu
int smem_sweep_launch(bwaflow_smem_batch *b, const SmemArgs &a, bool text);
int smem_sweep_setup(bwaflow_smem_batch *b);
int build_level_tables(bwaflow_ctx *c, int K, void **d_tab);
// smem.cu: device-side halves of submit/collect (no interval D2H); used by the fused hot path
int smem_launch(bwaflow_smem_batch *b);
int smem_tail_device(bwaflow_smem_batch *b);                   // hot path: compact + seed positions, no host read-back
int smem_status_d2h(bwaflow_smem_batch *b);
int smem_check_capacity(bwaflow_smem_batch *b, bool *redo, int64_t *n_pos);
int smem_finish_device(bwaflow_smem_batch *b);                 // waits, regrows slots on overflow, compacts into d_out/d_off; sets n_intv
int smem_seedpos_device(bwaflow_smem_batch *b, int64_t *total); // d_pos/d_poff for the current intervals
// chain2aln.cu
int c2a_launch(bwaflow_c2a_batch *b);
int c2a_finish_queue(bwaflow_c2a_batch *b);
} // namespace bwaflow
