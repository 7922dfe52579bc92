/* bwaflow_cuda.h -- C ABI of the B200 (sm_100a) backend for bwa-flow's per-read compute path.
 *
 * This is the boundary a `src/cuda` backend of falcon-computing/bwa-flow binds, in the same
 * place `src/fpga` plugs in (kflow stage slots 2 and 4, src/main.cpp:302-335).  Every entry
 * point cites the reference interface it replaces.  Plain pointers and sizes only; no C++,
 * no torch, no reference headers needed.  All functions return 0 on success and a negative
 * errno-style code on failure (message via bwaflow_last_error()); nothing throws across the
 * ABI and there is NO CPU fallback: a missing GPU is an error.
 *
 * Shape of the interface = the FPGA batch interface it mirrors:
 *   BWAOCLEnv ctor (src/fpga/BWAOCLEnv.h:67-216: upload pac + BWT once)   -> bwaflow_ctx_create
 *   SMemTask{start,finish} (src/fpga/SMemTask.h:9-47)                     -> bwaflow_smem_{submit,collect}
 *   SWTask{start,finish}   (src/fpga/SWTask.h:9-50)                       -> bwaflow_c2a_{submit,collect}
 *                                                                            bwaflow_ext_{submit,collect}
 *   both slots fused, chains never leaving HBM (ChainsRecord::tag = 1)      -> bwaflow_hot_{submit,collect}
 *   RegionsToSam::compute (src/Pipeline.cpp:546-648), slot 5               -> bwaflow_sam_process, built from
 *                                                                            bwaflow_dedup_patch, bwaflow_pestat,
 *                                                                            bwaflow_matesw_*, bwaflow_cigar_*, bwaflow_sw_*
 * submit is asynchronous (H2D + kernels + D2H queued on the batch's own stream), collect blocks;
 * a worker keeps two batch objects to double-buffer, exactly as FPGAPipeline.cpp:374-386 does.
 * The kflow stage classes over these calls are bwaflow_b200/host/CUDAPipeline.{h,cpp}.
 */
#ifndef BWAFLOW_CUDA_H
#define BWAFLOW_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plain-data mirrors of the reference structs that cross the boundary ------------------ */

/* bwtintv_t, bwa/bwt.h:60-62 (32 B): x[0] fwd SA start, x[1] rev-comp SA start, x[2] size, info = qbeg<<32|qend */
typedef struct { uint64_t x[3], info; } bwaflow_intv_t;

/* mem_seed_t, src/bwa_wrapper.h:64-68 (24 B) */
typedef struct { int64_t rbeg; int32_t qbeg, len, score, _pad; } bwaflow_seed_t;

/* the fields of mem_chain_t (src/bwa_wrapper.h:70-77) that mem_chain2aln reads, flattened:
 * seeds of chain c are seeds[seed_off .. seed_off+n_seeds) of the batch-wide seed array.
 * flags = w | kept<<29 | is_alt<<31 (carried, not interpreted). */
typedef struct { int64_t pos; int32_t n_seeds, rid; float frac_rep; uint32_t flags; int64_t seed_off; } bwaflow_chain_t;

/* mem_alnreg_t, bwa/bwamem.h:60-78 (88 B, identical layout; the device fills what mem_chain2aln
 * fills -- rb re qb qe rid score truesc w seedcov seedlen0 frac_rep -- and zeroes the rest) */
typedef struct {
  int64_t rb, re; int32_t qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary,
  secondary_all, seedlen0; uint32_t n_comp_is_alt; float frac_rep; uint64_t hash;
} bwaflow_alnreg_t;

/* One ksw_extend2 call (bwa/ksw.c:380): query = qpool[q_off..+qlen) codes 0..4, target = tpool[t_off..+tlen)
 * codes 0..3(4).  w is the caller's (unclamped) band; the clamp of ksw.c:399-407 happens on the device.  zdrop must equal the
 * context's opt.zdrop (mem_chain2aln passes opt->zdrop to every call, bwa/bwamem.c:734,762); other values are rejected. */
typedef struct { int32_t qlen, tlen, w, h0, end_bonus, zdrop; int64_t q_off, t_off; } bwaflow_ext_job_t;
/* its six outputs: return value, *_qle, *_tle, *_gtle, *_gscore, *_max_off (bwa/ksw.c:473-478) */
typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } bwaflow_ext_result_t;

/* Index arrays exactly as bwa holds them in memory (bwt_t bwa/bwt.h:46-58, bntseq_t bwa/bntseq.h:56-64,
 * bwaidx_t::pac bwa/bwa.h:22-30).  Host pointers; copied to HBM once by bwaflow_ctx_create. */
typedef struct {
  uint64_t primary, L2[5], seq_len, bwt_size; const uint32_t *bwt;   /* occ-interleaved BWT, bwt_size u32 words */
  uint64_t n_sa; int64_t sa_intv; const uint64_t *sa;                /* sampled SA, sa[0] = (u64)-1 */
  int64_t l_pac; const uint8_t *pac; int32_t n_seqs, _pad;           /* 2-bit forward reference, l_pac/4+1 bytes */
  const int64_t *ctg_offset; const int32_t *ctg_len;                 /* bns->anns[i].offset / .len, n_seqs each */
} bwaflow_index_t;

/* The mem_opt_t fields the path reads (bwa/bwamem.h:26-58; defaults bwa/bwamem.c:48-84). */
typedef struct {
  int32_t a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop;
  int32_t min_seed_len, split_width, max_occ, max_mem_intv;
  float split_factor; int8_t mat[25]; int8_t _pad[3];
} bwaflow_opt_t;

typedef struct bwaflow_ctx bwaflow_ctx_t;
typedef struct bwaflow_ext_batch bwaflow_ext_batch_t;
typedef struct bwaflow_c2a_batch bwaflow_c2a_batch_t;
typedef struct bwaflow_smem_batch bwaflow_smem_batch_t;

/* per-call device timings (CUDA events on the batch stream), milliseconds */
typedef struct { float h2d_ms, kernel_ms, d2h_ms, total_ms; int32_t n_launches, _pad; double work_units; } bwaflow_timing_t;

/* ---- context ------------------------------------------------------------------------------ */

const char *bwaflow_last_error(void);
int bwaflow_device_count(void);
void bwaflow_opt_default(bwaflow_opt_t *opt);                       /* mem_opt_init, bwa/bwamem.c:48-84 */

/* replaces `new BWAOCLEnv()` (src/main.cpp:212, src/fpga/BWAOCLEnv.h:67-216).  idx may be NULL for a
 * context that only runs raw extension jobs. */
int bwaflow_ctx_create(int device, const bwaflow_index_t *idx, const bwaflow_opt_t *opt, bwaflow_ctx_t **out);
void bwaflow_ctx_destroy(bwaflow_ctx_t *ctx);
int bwaflow_ctx_set_opt(bwaflow_ctx_t *ctx, const bwaflow_opt_t *opt);

/* pinned host staging memory (the FPGA tasks use clCreateBuffer+map for the same purpose, SWTask.cpp:30-60) */
void *bwaflow_host_alloc(uint64_t bytes);
void bwaflow_host_free(void *p);

/* device-side stopwatch (CUDA events on the context's device) for callers that time whole steps */
typedef struct bwaflow_stopwatch bwaflow_stopwatch_t;
int bwaflow_stopwatch_create(bwaflow_ctx_t *ctx, bwaflow_stopwatch_t **out);
int bwaflow_stopwatch_start(bwaflow_stopwatch_t *w);          /* device sync + event record */
int bwaflow_stopwatch_stop(bwaflow_stopwatch_t *w, float *ms); /* device sync + event record + elapsed */
void bwaflow_stopwatch_destroy(bwaflow_stopwatch_t *w);
int bwaflow_ctx_sync(bwaflow_ctx_t *ctx);
/* Measured integer peak for the extension roofline: chip-wide issue rate of dependency-free packed
 * DPX instructions (VIMNMX3.S16x2 / VIADDMNMX.S16x2.RELU), in lane-instructions per second, and of their 32-bit forms. */
int bwaflow_dpx_peak(bwaflow_ctx_t *ctx, double *s16x2_lane_ops_per_s, double *s32_lane_ops_per_s);

/* ---- extension jobs: one ksw_extend2 each (replaces the per-seed SW stream of SWTask, src/fpga/SWTask.h) */

int bwaflow_ext_batch_create(bwaflow_ctx_t *ctx, int64_t max_jobs, int64_t max_qbytes, int64_t max_tbytes,
                             bwaflow_ext_batch_t **out);
void bwaflow_ext_batch_destroy(bwaflow_ext_batch_t *b);
/* SWTask::start: async H2D of jobs+pools, device-side length binning, kernels, async D2H of results */
int bwaflow_ext_submit(bwaflow_ext_batch_t *b, int64_t n_jobs, const bwaflow_ext_job_t *jobs,
                       const uint8_t *qpool, int64_t qbytes, const uint8_t *tpool, int64_t tbytes);
/* SWTask::finish: blocks; (*out)[n_jobs] in job order, pinned host memory owned by the batch, valid until
 * its next submit */
int bwaflow_ext_collect(bwaflow_ext_batch_t *b, const bwaflow_ext_result_t **out, bwaflow_timing_t *timing);
/* device-resident re-run of the last submitted inputs (bench: inputs already in HBM); results stay on device
 * until bwaflow_ext_collect.  */
int bwaflow_ext_rerun(bwaflow_ext_batch_t *b);
/* same, with the kernel variant that also counts the DP cells ksw_extend2 executes for these jobs (the GCUPS work unit,
 * bwa/ksw.c:424-448 iterations); the count is reported as timing.work_units by this and later collects of the same inputs.
 * Results are identical to the uncounted variant. */
int bwaflow_ext_rerun_counted(bwaflow_ext_batch_t *b);

/* ---- stage 4: chains -> alignment regions (replaces ChainsToRegionsFPGA::compute, src/fpga/FPGAPipeline.cpp:367;
 *      arithmetic = mem_chain2aln bwa/bwamem.c:641-795 incl. the containment skip the FPGA path drops) */

int bwaflow_c2a_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, int64_t max_chains,
                             int64_t max_seeds, bwaflow_c2a_batch_t **out);
void bwaflow_c2a_batch_destroy(bwaflow_c2a_batch_t *b);
/* bases = nt4 codes of all reads concatenated (bseq1_t::seq after src/bwa_wrapper.cpp:124-125);
 * read_off[n_reads+1]; chains of read r = chains[chain_off[r] .. chain_off[r+1]) in mem_chain_v order.
 * bases may be NULL to reuse the reads already resident from the previous submit (same read_off). */
int bwaflow_c2a_submit(bwaflow_c2a_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off,
                       const bwaflow_chain_t *chains, const int64_t *chain_off,
                       const bwaflow_seed_t *seeds, int64_t n_seeds);
/* blocks; regions of read r = (*regs)[(*reg_off)[r] .. (*reg_off)[r+1]) in the order mem_chain2aln appends them.
 * The returned arrays are pinned host memory owned by the batch, valid until its next submit. */
int bwaflow_c2a_collect(bwaflow_c2a_batch_t *b, const bwaflow_alnreg_t **regs, const int64_t **reg_off,
                        int64_t *n_regs, bwaflow_timing_t *timing);
int bwaflow_c2a_rerun(bwaflow_c2a_batch_t *b);
/* same on the exact lane-per-read kernel for every read, which also counts the DP cells ksw_extend2 executes (reported as
 * timing.work_units by this and later collects of the same input); regions are identical */
int bwaflow_c2a_rerun_counted(bwaflow_c2a_batch_t *b);
/* after a collect: how many reads of the last launch were (re)done by the exact lane-per-read kernel rather than served by the
 * speculative first-seed waves + replay (all of them when speculation is off or for a counted pass) */
int bwaflow_c2a_exact_reads(bwaflow_c2a_batch_t *b, int64_t *n_exact);
/* after a collect: how many reads the first replay left and how many extension jobs (their later seeds) the second round of
 * waves ran ahead for them before the second replay; both 0 when every read was served by the first-seed waves */
int bwaflow_c2a_second_round(bwaflow_c2a_batch_t *b, int64_t *n_reads, int64_t *n_jobs);

/* ---- stage 2: reads -> SMEM intervals (replaces SeqsToChainsFPGA::compute, src/fpga/FPGAPipeline.cpp:582;
 *      arithmetic = mem_collect_intv_new src/bwa_wrapper.cpp:207-246).  Output per read: bwtintv_t list sorted
 *      by info, what ChainsPipe::compute (src/Pipeline.cpp:408) consumes with tag=0.  No per-read cap. */

int bwaflow_smem_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, bwaflow_smem_batch_t **out);
void bwaflow_smem_batch_destroy(bwaflow_smem_batch_t *b);
int bwaflow_smem_submit(bwaflow_smem_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off);
int bwaflow_smem_collect(bwaflow_smem_batch_t *b, const bwaflow_intv_t **intv, const int64_t **intv_off,
                         int64_t *n_intv, bwaflow_timing_t *timing);
int bwaflow_smem_rerun(bwaflow_smem_batch_t *b);
/* on != 0: an interval that is unique (x[2] == 1) may be returned "anchored" as x[0] = its position in the forward +
 * reverse-complement text (what bwt_sa(x[0]) would give), x[1] = BWAFLOW_ANCHORED, x[2] = 1 -- the form the fused hot
 * path uses internally: such intervals are extended by comparing bases against the 2-bit text and their seed
 * position needs no suffix-array read.  bwaflow_smem_seed_positions understands both forms.  Default off (BWT rows). */
#define BWAFLOW_ANCHORED ((1ull << 39) - 1)
int bwaflow_smem_set_anchored(bwaflow_smem_batch_t *b, int on);
/* same on the kernel instances that also count the 64-byte occ-block touches the reference makes for these reads (bwt_2occ4,
 * bwa/bwt.c:189-220: one block if both interval ends share it, else two); reported as timing.work_units.  Same intervals. */
int bwaflow_smem_rerun_counted(bwaflow_smem_batch_t *b);

/* Seed positions for chaining (replaces the bwt_sa loop of mem_chain_postprocess, src/bwa_wrapper.cpp:281-291):
 * for the intervals of the last collect, pos[] holds bwt_sa(x[0]+k) for k = 0, step, 2*step.. (count < max_occ)
 * in interval order; pos_off[n_intv+1].  Pinned memory owned by the batch. */
int bwaflow_smem_seed_positions(bwaflow_smem_batch_t *b, const int64_t **pos, const int64_t **pos_off,
                                int64_t *n_pos, bwaflow_timing_t *timing);

/* plain SA lookups (bwt_sa, bwa/bwt.c:86-96) */
int bwaflow_sa_lookup(bwaflow_ctx_t *ctx, int64_t n, const uint64_t *k, uint64_t *out);

/* ---- host chaining between the two device stages (the role of ChainsPipe::compute, src/Pipeline.cpp:408-438:
 *      mem_chain_postprocess src/bwa_wrapper.cpp:248-323 with the bwt_sa results supplied by
 *      bwaflow_smem_seed_positions, then mem_chain_flt bwa/bwamem.c:336-394).  CPU code, multi-threaded over
 *      reads; offered so that a backend can go reads -> regions without the reference's host objects.  A drop-in
 *      integration may keep using the reference's own ChainsPipe instead (tag=0 records). */
typedef struct {
  int32_t w, max_chain_gap, min_chain_weight, max_chain_extend, min_seed_len, max_occ;
  float mask_level, drop_ratio;
} bwaflow_chain_opt_t;
typedef struct bwaflow_chain_result bwaflow_chain_result_t;
void bwaflow_chain_opt_default(bwaflow_chain_opt_t *opt);
/* Longest read the path accepts with these options: below 1 024 bases (row state of the seeding kernels) and short enough that
 * mem_flt_chained_seeds (bwa/bwamem.c:607-624) does not re-score seeds -- 724 bases at default options.  A batch holding a longer
 * read is refused as a whole (-34); a caller that must take such reads checks lengths first and routes the record elsewhere. */
int32_t bwaflow_max_read_len(const bwaflow_chain_opt_t *opt);
int bwaflow_chain_build(const bwaflow_chain_opt_t *opt, int64_t l_pac, int32_t n_seqs, const int64_t *ctg_offset,
                        const int32_t *ctg_len, const int32_t *ctg_is_alt /* may be NULL */,
                        int64_t n_reads, const int64_t *read_off,
                        const bwaflow_intv_t *intv, const int64_t *intv_off, const int64_t *pos, const int64_t *pos_off,
                        int n_threads /* <1: all cores */, bwaflow_chain_result_t **out);
/* chains of read r = chains[chain_off[r] .. chain_off[r+1]); seed_off indexes `seeds`; mem_chain_v order */
void bwaflow_chain_result_view(const bwaflow_chain_result_t *r, const bwaflow_chain_t **chains, int64_t *n_chains,
                               const bwaflow_seed_t **seeds, int64_t *n_seeds, const int64_t **chain_off);
void bwaflow_chain_result_free(bwaflow_chain_result_t *r);

/* ---- fused hot path: reads -> regions with intervals, SA positions and chains resident in HBM (stage 2, the chaining
 *      hop of ChainsPipe src/Pipeline.cpp:408-438 and stage 4 back to back on one stream).  Results are identical
 *      to running the three pieces separately.  ctg_is_alt may be NULL. */
typedef struct bwaflow_hot_batch bwaflow_hot_batch_t;
typedef struct {
  float h2d_ms, smem_ms, seedpos_ms, chain_ms, c2a_ms, d2h_ms, total_ms; int32_t n_launches;
  double smem_blocks, seedpos_blocks, cells;           /* 64-byte occ-block touches; executed DP cells */
  int64_t n_intv, n_pos, n_chains, n_seeds;
} bwaflow_hot_timing_t;
int bwaflow_hot_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, const bwaflow_chain_opt_t *copt,
                             const int32_t *ctg_is_alt, bwaflow_hot_batch_t **out);
void bwaflow_hot_batch_destroy(bwaflow_hot_batch_t *b);
int bwaflow_hot_submit(bwaflow_hot_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off);
int bwaflow_hot_rerun(bwaflow_hot_batch_t *b);        /* reads of the last submit, already resident */
int bwaflow_hot_exact_reads(bwaflow_hot_batch_t *b, int64_t *n_exact);   /* as bwaflow_c2a_exact_reads, after bwaflow_hot_collect */
int bwaflow_hot_second_round(bwaflow_hot_batch_t *b, int64_t *n_reads, int64_t *n_jobs);   /* as bwaflow_c2a_second_round */
int bwaflow_hot_rerun_counted(bwaflow_hot_batch_t *b); /* same on the counting instances of stage 2 (timing.smem_blocks) and stage 4 (timing.cells, as bwaflow_c2a_rerun_counted) */
int bwaflow_hot_collect(bwaflow_hot_batch_t *b, const bwaflow_alnreg_t **regs, const int64_t **reg_off, int64_t *n_regs,
                        bwaflow_hot_timing_t *timing);

/* ---- mate rescue (stage 5) ------------------------------------------------------------------------
 *      ksw_align2 (bwa/ksw.c:343-365 over ksw_u8 / ksw_i16 :110-339): local Smith-Waterman of a whole query against a target
 *      window, second-best score, start by the reversed pass.  xtra as in bwa/ksw.h:6-9: KSW_XBYTE 0x10000, KSW_XSTOP 0x20000,
 *      KSW_XSUBO 0x40000, KSW_XSTART 0x80000, low 16 bits = the threshold.  Scoring matrix and gap costs are the context's. */
typedef struct { int32_t qlen, tlen, xtra, _pad; int64_t q_off, t_off; } bwaflow_sw_job_t;
typedef struct { int32_t score, te, qe, score2, te2, tb, qb; } bwaflow_sw_result_t;       /* kswr_t, bwa/ksw.h:14-20 */
typedef struct bwaflow_sw_batch bwaflow_sw_batch_t;
int bwaflow_sw_batch_create(bwaflow_ctx_t *ctx, int64_t max_jobs, int64_t max_qbytes, int64_t max_tbytes, bwaflow_sw_batch_t **out);
void bwaflow_sw_batch_destroy(bwaflow_sw_batch_t *b);
int bwaflow_sw_submit(bwaflow_sw_batch_t *b, int64_t n_jobs, const bwaflow_sw_job_t *jobs, const uint8_t *qpool, int64_t qbytes,
                      const uint8_t *tpool, int64_t tbytes);
int bwaflow_sw_collect(bwaflow_sw_batch_t *b, const bwaflow_sw_result_t **out, bwaflow_timing_t *timing);

/* mem_pestat_t, bwa/bwamem.h:80-86 (32 B, identical layout) */
typedef struct { int32_t low, high, failed, _pad; double avg, std; } bwaflow_pestat_t;
/* the mem_opt_t fields the pairing steps read besides bwaflow_opt_t (defaults bwa/bwamem.c:59-78) */
typedef struct {
  int32_t pen_unpaired, max_matesw, max_ins, max_chain_gap; float mask_level, mask_level_redun; int32_t no_rescue;
  float mapq_coef_len; int32_t mapq_coef_fac, _pad;     /* mapQ_coef_len (float) and mapQ_coef_fac (int: log(50) truncated), bwa/bwamem.h:52-53 */
} bwaflow_pe_opt_t;
void bwaflow_pe_opt_default(bwaflow_pe_opt_t *po);
/* mem_pestat (bwa/bwamem_pair.c:49-112) over the regions of one whole batch (src/Pipeline.cpp:578-579): reads 2k and 2k+1 are
 * mates, regions of read r = regs[reg_off[r] .. reg_off[r+1]) after mem_sort_dedup_patch.  Host code. */
int bwaflow_pestat(const bwaflow_opt_t *opt, const bwaflow_pe_opt_t *po, int64_t l_pac, int64_t n_reads, const bwaflow_alnreg_t *regs,
                   const int64_t *reg_off, bwaflow_pestat_t pes[4]);
/* The rescue block of mem_sam_pe (src/bwa_wrapper.cpp:657-668) for every pair of a batch: mem_matesw (bwa/bwamem_pair.c:114-183)
 * for the up-to-max_matesw best regions of either end against the mate.  Every window any of those calls can ask for is aligned on
 * the device ahead of time (ksw_align2 kernel); collect then replays the calls in the reference's order on the host -- which
 * windows are skipped depends on the hits rescued so far -- inserting hits and re-running mem_sort_dedup_patch (no patching, as
 * mem_matesw calls it) exactly as the reference does.  Input regions: after mem_sort_dedup_patch + ALT marking
 * (src/Pipeline.cpp:560-575).  Output: the regions mem_mark_primary_se / mem_pair see.  *n_sw = sum of mem_matesw's return values. */
typedef struct bwaflow_matesw_batch bwaflow_matesw_batch_t;
int bwaflow_matesw_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, const bwaflow_pe_opt_t *po,
                                bwaflow_matesw_batch_t **out);
void bwaflow_matesw_batch_destroy(bwaflow_matesw_batch_t *b);
int bwaflow_matesw_submit(bwaflow_matesw_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off,
                          const bwaflow_alnreg_t *regs, const int64_t *reg_off, const bwaflow_pestat_t pes[4]);
int bwaflow_matesw_collect(bwaflow_matesw_batch_t *b, const bwaflow_alnreg_t **regs, const int64_t **reg_off, int64_t *n_regs,
                           int64_t *n_sw, int64_t *n_jobs, bwaflow_timing_t *timing);

/* ---- CIGAR generation (stage 5): mem_reg2aln (bwa/bwamem.c:1104-1174) for a batch of (read, region) jobs --
 *      band inference and up to three band tries, bwa_gen_cigar2 (bwa/bwa.c:121-207) with ksw_global2 (bwa/ksw.c:504-606) and its
 *      traceback, NM and MD, the squeeze of a leading / trailing deletion, clipping, position.  One job = one call of mem_reg2aln on a
 *      mapped region (the caller decides which regions are reported: primary, supplementary, the XA candidates of mem_gen_alt). */
typedef struct {                                  /* mem_aln_t, bwa/bwamem.h:88-99, with the CIGAR and MD as offsets into the pools */
  int64_t pos; int32_t rid, flag, is_rev, is_alt, mapq, NM, n_cigar, score, sub, alt_sc;
  int32_t global_score;                           /* score of the last ksw_global2 / ungapped pass (mem_reg2aln's local `score`) */
  int64_t cigar_off, md_off;                      /* cigar = cigar_pool + cigar_off (n_cigar words, opLen<<4|op, MIDSH = 01234); MD = md_pool + md_off, NUL-terminated */
} bwaflow_aln_t;
typedef struct bwaflow_cigar_batch bwaflow_cigar_batch_t;
int bwaflow_cigar_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, int64_t max_jobs, const bwaflow_pe_opt_t *po,
                               bwaflow_cigar_batch_t **out);
void bwaflow_cigar_batch_destroy(bwaflow_cigar_batch_t *b);
/* job j: region job_reg[j] (as it stands when mem_reg2aln would be called: sub, csub, secondary .. filled) of read job_read[j] */
int bwaflow_cigar_submit(bwaflow_cigar_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off, int64_t n_jobs,
                         const int32_t *job_read, const bwaflow_alnreg_t *job_reg);
int bwaflow_cigar_collect(bwaflow_cigar_batch_t *b, const bwaflow_aln_t **aln, const uint32_t **cigar_pool, const char **md_pool,
                          bwaflow_timing_t *timing);

/* ---- region de-duplication (stage 5, first step): mem_sort_dedup_patch (bwa/bwamem.c:446-498) for every read of a batch, as
 *      RegionsToSam::compute runs it with the real bns / pac / query (src/Pipeline.cpp:560-568), i.e. WITH mem_patch_reg (:415-444):
 *      two colinear hits are merged when a banded global alignment over their union (bwa_gen_cigar2 with n_cigar = 0, bwa/bwa.c:121,
 *      -> ksw_global2 score only) scores at least PATCH_MIN_SC_RATIO of what the hits predict.  The scans run on the host; every
 *      global alignment they ask for runs on the device through the CIGAR batch's kernel in its score-only mode.  A read whose scan
 *      needs a score that is not known yet is suspended, the requests of all suspended reads go to the device as one launch, and the
 *      scans are redone (a merge changes what later tests ask for, so the requests cannot be enumerated up front); *rounds = number
 *      of launches.  regs_out needs room for reg_off[n_reads] regions.  mask_level_redun / max_chain_gap are the CIGAR batch's. */
typedef struct { int32_t read, qb, qe, w; int64_t rb, re; } bwaflow_gscore_job_t;   /* query[qb..qe) of `read` vs reference [rb, re), band w_ */
/* scores[j] = *score of bwa_gen_cigar2(mat, gaps, w, l_pac, pac, qe - qb, query + qb, rb, re, &score, 0, 0).  upload != 0 copies the
 * batch's bases to the device first; 0 = those of the previous call are still resident.  Blocks until the scores are back. */
int bwaflow_cigar_global_scores(bwaflow_cigar_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off, int upload,
                                int64_t n_jobs, const bwaflow_gscore_job_t *jobs, int32_t *scores);
int bwaflow_dedup_patch(bwaflow_cigar_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off,
                        const bwaflow_alnreg_t *regs, const int64_t *reg_off, bwaflow_alnreg_t *regs_out, int64_t *reg_off_out,
                        int64_t *n_patch_jobs, int32_t *rounds);

/* ---- stage 5 as a whole: RegionsToSam::compute (src/Pipeline.cpp:546-648) for one batch = the role a RegionsToSamCUDA stage
 *      plays behind slot 5 of compute_flow.  Regions of stage 4 in, per-read SAM text out (the string mem_aln2sam_hts builds before
 *      sam_parse1, src/bwa_wrapper.cpp:452-590 -- the comparison point of SURVEY 8c):
 *        mem_sort_dedup_patch + ALT marking -> mem_pestat -> per pair mem_sam_pe_hts (src/bwa_wrapper.cpp:640-790): mate rescue,
 *        mem_mark_primary_se, mem_pair, mapping qualities, mem_reg2aln for every reported hit and every XA candidate of mem_gen_alt
 *        (bwa/bwamem_extra.c:99-152), mem_aln2sam_hts; single-end batches: mem_mark_primary_se + mem_reg2sam_hts (:592-637).
 *      Device: ksw_global2 (patch scores, CIGARs) and ksw_align2 (mate rescue).  Host (worker threads over pairs): the sequential
 *      decisions in between and the text. */
typedef struct {
  int32_t flag;                     /* MEM_F_* of bwa/bwamem.h:14-24: PE 0x2, NOPAIRING 0x4, ALL 0x8, NO_MULTI 0x10 (-M, bwa-flow's
                                       default, src/config.cpp:16), NO_RESCUE 0x20, REF_HDR 0x100, SOFTCLIP 0x200, XB 0x2000 */
  int32_t T;                        /* minimum score to output, default 30 */
  float drop_ratio, XA_drop_ratio;  /* 0.50, 0.80 */
  int32_t max_XA_hits, max_XA_hits_alt;   /* 5, 200 */
  int32_t n_threads, _pad;          /* host threads of the decision / text phases; < 1: all cores */
} bwaflow_sam_opt_t;
void bwaflow_sam_opt_default(bwaflow_sam_opt_t *so);       /* flag = PE | NO_MULTI */
typedef struct {
  float dedup_ms, matesw_ms, pair_ms, cigar_ms, text_ms, total_ms;     /* host wall clock per phase (the device work is inside) */
  int64_t n_patch_jobs, n_matesw_jobs, n_matesw_used, n_cigar_jobs, sam_bytes;
  float matesw_kernel_ms, cigar_kernel_ms;                             /* of which on the device (CUDA events): ksw_align2 / mem_reg2aln kernels */
} bwaflow_sam_timing_t;
typedef struct bwaflow_sam_batch bwaflow_sam_batch_t;
/* ctg_names: n_seqs NUL-terminated names (bns->anns[i].name); ctg_is_alt / ctg_anno (XR:Z with MEM_F_REF_HDR) may be NULL;
 * rg_id: the read-group ID printed as RG:Z (bwa_rg_id), NULL or "" for none.  The strings are copied. */
int bwaflow_sam_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, const bwaflow_pe_opt_t *po,
                             const bwaflow_sam_opt_t *so, const char *const *ctg_names, const uint8_t *ctg_is_alt,
                             const char *const *ctg_anno, const char *rg_id, bwaflow_sam_batch_t **out);
void bwaflow_sam_batch_destroy(bwaflow_sam_batch_t *b);
/* One batch.  bases: nt4 codes, read r = bases[read_off[r] .. read_off[r+1]); quals: same offsets, NULL = '*'; names / comments:
 * NUL-terminated strings at names + name_off[r] (comment_off[r] < 0 or comments == NULL: none); reads 2k and 2k+1 are mates when
 * MEM_F_PE is set.  regs / reg_off: stage 4's output.  start_idx = SeqsRecord::start_idx (the pair id (start_idx>>1)+k seeds the
 * tie-breaking hashes, src/Pipeline.cpp:589).  Output: sam = the reads' SAM lines concatenated in read order, read r's lines =
 * sam[sam_off[r] .. sam_off[r+1]); valid until the next call on this batch.  pes_out (may be NULL) = mem_pestat of the batch. */
int bwaflow_sam_process(bwaflow_sam_batch_t *b, int64_t n_reads, const uint8_t *bases, const uint8_t *quals, const int64_t *read_off,
                        const char *names, const int64_t *name_off, const char *comments, const int64_t *comment_off,
                        const bwaflow_alnreg_t *regs, const int64_t *reg_off, int64_t start_idx,
                        const char **sam, const int64_t **sam_off, bwaflow_pestat_t pes_out[4], bwaflow_sam_timing_t *timing);
/* @HD / @SQ / @PG header lines as bwa_format_sam_hdr writes them (src/bwa_wrapper.cpp:25-35); returns the length, copies at most cap bytes */
int64_t bwaflow_sam_header(bwaflow_sam_batch_t *b, const char *hdr_line /* @RG line etc., may be NULL */, const char *pg_line, char *buf, int64_t cap);

/* ---- index construction for synthetic references (workload preparation; the arrays equal what `bwa index`
 *      + bwa_idx_load leave in memory, bwa/bwtindex.c:256-324).  codes: 0..3 = ACGT, >=4 = N, contigs concatenated. */
typedef struct bwaflow_built_index bwaflow_built_index_t;
int bwaflow_index_build(const uint8_t *codes, int64_t n_bases, int32_t n_contigs, const int64_t *contig_len,
                        int n_threads /* <1: all cores */, bwaflow_built_index_t **out);
/* the same arrays built on GPU `device` (suffix sort, BWT, occurrence layout and SA sampling in HBM); needs about 12 bytes of
 * device memory per suffix (2 * n_bases suffixes).  Bit-identical to bwaflow_index_build. */
int bwaflow_index_build_gpu(int device, const uint8_t *codes, int64_t n_bases, int32_t n_contigs, const int64_t *contig_len,
                            bwaflow_built_index_t **out);
const bwaflow_index_t *bwaflow_built_index_view(const bwaflow_built_index_t *b);
void bwaflow_built_index_free(bwaflow_built_index_t *b);

#ifdef __cplusplus
}
#endif
#endif
