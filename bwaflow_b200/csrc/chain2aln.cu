// Stage 4 on the device: chains -> alignment regions, exact mem_chain2aln semantics
// (bwa/bwamem.c:641-795) for every chain of every read of a batch, including the containment skip
// (:680-715) that the FPGA path drops.  Mirrors ChainsToRegionsFPGA::compute's role
// (src/fpga/FPGAPipeline.cpp:367) behind a SWTask-like submit/collect pair.
//
// One lane owns one read at a time and walks its chains and seeds in the reference's order, because
// whether seed k is extended depends on the regions earlier seeds (and earlier chains) of the same
// read produced.  Each extension is run by the lane itself with the row-stepped engine in
// ksw_lane.cuh (left flank, then right flank seeded with the left score, up to two band tries each).
// The target is read straight from the 2-bit reference in HBM, the query from the read.
//
// Device pipeline per submit:
//   c2a_chain_prep   per chain: reference window [lo,hi) (:651-668 + bns_fetch_seq clipping), sort keys
//   cub scan + sort  seeds of each chain permuted into processing order (score descending, :671-676)
//   c2a_read_prep    per read: LPT key (total seeds) and qlen class
//   c2a_kernel       persistent lanes pulling reads from per-class queues
//   scan + gather    compact the per-read region lists
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <string.h>
#include <stdlib.h>
#include "internal.cuh"
#include "ksw_lane.cuh"
#include "batches.cuh"

namespace bwaflow {

__constant__ int c_c2a_qmax[C2A_NCLASS] = {31, 63, 95, 127, 159, 255, 511, 1023};
static const int h_c2a_qmax[C2A_NCLASS] = {31, 63, 95, 127, 159, 255, 511, 1023};
static const int h_c2a_threads[C2A_NCLASS] = {64, 64, 64, 64, 64, 64, 64, 32};

struct C2AOpt { int a, w, pen_clip5, pen_clip3, o_del, e_del, o_ins, e_ins; };

struct C2AArgs {
  const uint8_t *bases; const int64_t *read_off;
  const bwaflow_chain_t *chains; const int64_t *chain_off;
  const int64_t *sbase;              // [n_chains+1] exclusive prefix of n_seeds in chain order
  const bwaflow_seed_t *sseeds;      // seeds in processing order, chain c at sbase[c]
  const int64_t *win;                // [2*n_chains] reference window of each chain
  uint8_t *skipped;                  // per seed (processing order): 1 = extension skipped (:710)
  bwaflow_alnreg_t *regs;            // region slots, read r at sbase[chain_off[r]]
  int *n_regs;                       // [n_reads]
  const uint32_t *order; int *queue; int *err; unsigned long long *cells;
  int cls, qmax, refill;
  // results of the speculative first-seed waves (null when they did not run): chain c's first seed, left / right flank, first band
  // try; cflag[c] != 0 = not run ahead.  The exact kernel takes them instead of recomputing, so a read that was handed to it because
  // ONE of its chains needs more (a repeat read has hundreds of chains) only runs the extensions that are actually missing.
  const uint8_t *cflag; const bwaflow_ext_result_t *resL, *resR;
  // wave B: the same for every later seed of the reads the first replay left (slot_job[slot] = job or -1; null when it did not run)
  const int32_t *slot_job; const uint8_t *cflag2; const bwaflow_ext_result_t *resL2, *resR2;
  DevIndex ix; DevScore sc; C2AOpt o;
};

__device__ __forceinline__ void lane_take_result(KswLane &k, const bwaflow_ext_result_t &x)
{
  k.best = x.score; k.best_j = x.qle - 1; k.best_i = x.tle - 1; k.end_i = x.gtle - 1; k.gscore = x.gscore; k.max_off = x.max_off;
  k.cells = 0;
}

// cal_max_gap, bwa/bwamem.c:630-637 (integer-exact form of the double arithmetic, see gap_limit)
__device__ __forceinline__ int max_gap(const C2AOpt &o, int qlen)
{
  const int l_del = gap_limit(qlen * o.a - o.o_del, o.e_del);
  const int l_ins = gap_limit(qlen * o.a - o.o_ins, o.e_ins);
  const int l = l_del > l_ins ? l_del : l_ins;
  return l < o.w << 1 ? l : o.w << 1;
}

// base at coordinate p of the forward + reverse-complement text (bns_get_seq, bwa/bntseq.c:398-419)
__device__ __forceinline__ int ref_base(const DevIndex &ix, int64_t p)
{
  const bool rev = p >= ix.l_pac;
  const int64_t f = rev ? (ix.l_pac << 1) - 1 - p : p;
  const int b = (__ldg(ix.pac + (f >> 2)) >> ((~f & 3) << 1)) & 3;
  return rev ? 3 - b : b;
}

__global__ void c2a_chain_prep(const bwaflow_chain_t *chains, int64_t n_chains, const int64_t *chain_read,
                               const int64_t *read_off, const bwaflow_seed_t *seeds, const int64_t *sbase,
                               DevIndex ix, C2AOpt o, int64_t *win, uint64_t *keys, uint32_t *vals, int32_t *slot_chain, int *err)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const bwaflow_chain_t ch = chains[c];
  const int64_t r = chain_read[c];
  const int lq = (int)(read_off[r + 1] - read_off[r]);
  const bwaflow_seed_t *sd = seeds + ch.seed_off;
  const int64_t L = ix.l_pac;
  int64_t lo = L << 1, hi = 0;
  if ((uint64_t)c >= (1ull << 24) || ch.n_seeds >= (1 << 20)) atomicExch(err, 3);
  for (int i = 0; i < ch.n_seeds; ++i) {
    const bwaflow_seed_t s = sd[i];
    const int tail = lq - s.qbeg - s.len;
    const int64_t b = s.rbeg - (s.qbeg + max_gap(o, s.qbeg));
    const int64_t e = s.rbeg + s.len + (tail + max_gap(o, tail));
    lo = b < lo ? b : lo;
    hi = e > hi ? e : hi;
    if ((unsigned)s.score >= (1u << 20)) atomicExch(err, 3);
    // ascending sort => chain-major, then (score, index) descending  (srt[] walked from the top, :671-676)
    const uint64_t inner = ((uint64_t)(uint32_t)s.score << 20 | (uint32_t)i) ^ 0xFFFFFFFFFFull;
    keys[sbase[c] + i] = (uint64_t)c << 40 | inner;
    vals[sbase[c] + i] = (uint32_t)(ch.seed_off + i);
    slot_chain[sbase[c] + i] = (int32_t)c;
  }
  if (ch.n_seeds > 0) {
    lo = lo > 0 ? lo : 0;
    hi = hi < L << 1 ? hi : L << 1;
    if (lo < L && L < hi) { if (sd[0].rbeg < L) hi = L; else lo = L; }
    // bns_fetch_seq: clip to the contig of the first seed (bwa/bntseq.c:421-446)
    int64_t fb = ix.ctg_offset[ch.rid], fe = fb + ix.ctg_len[ch.rid];
    if (sd[0].rbeg >= L) { const int64_t t = fb; fb = (L << 1) - fe; fe = (L << 1) - t; }
    lo = lo > fb ? lo : fb;
    hi = hi < fe ? hi : fe;
  }
  win[2 * c] = lo; win[2 * c + 1] = hi;
}

__global__ void c2a_mark_reads(const int64_t *chain_off, int64_t n_reads, int64_t *chain_read)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  for (int64_t c = chain_off[r]; c < chain_off[r + 1]; ++c) chain_read[c] = r;
}

__global__ void c2a_nseeds(const bwaflow_chain_t *chains, int64_t n_chains, int64_t *out)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_chains) out[c] = chains[c].n_seeds;
  if (c == n_chains) out[c] = 0;
}

__global__ void c2a_gather_seeds(const bwaflow_seed_t *seeds, const uint32_t *vals, int64_t n, bwaflow_seed_t *out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = seeds[vals[i]];
}

// `only` (optional): queue only the reads it marks; the others sort behind every class range and are never pulled
__global__ void c2a_read_prep(const int64_t *read_off, const int64_t *chain_off, const int64_t *sbase, const bwaflow_seed_t *sseeds,
                              int64_t n_reads, const uint8_t *only, uint32_t *keys, uint32_t *vals, int *hist, int *err)
{
  __shared__ int sh[C2A_NCLASS];
  if (threadIdx.x < C2A_NCLASS) sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads && only && !only[r]) { keys[r] = 0; vals[r] = (uint32_t)r; }
  else if (r < n_reads) {
    const int64_t lq = read_off[r + 1] - read_off[r];
    int c = 0;
    while (c < C2A_NCLASS - 1 && lq > c_c2a_qmax[c]) ++c;
    if (lq > c_c2a_qmax[C2A_NCLASS - 1]) atomicExch(err, 1);
    // queue order inside a class: reads with many seeds first (longest-processing-time first), and among equals the
    // ones whose first extension has a similar left-flank length next to each other, so that the lanes of a warp
    // walk rows of similar width
    int64_t ns = sbase[chain_off[r + 1]] - sbase[chain_off[r]];
    const int flank = ns > 0 ? sseeds[sbase[chain_off[r]]].qbeg : 0;
    ns = ns > 0xFFF ? 0xFFF : ns;
    keys[r] = 1u << 27 | (uint32_t)c << 24 | (uint32_t)ns << 12 | (uint32_t)(flank & 0xFFF);
    vals[r] = (uint32_t)r;
    atomicAdd(&sh[c], 1);
  }
  __syncthreads();
  if (threadIdx.x < C2A_NCLASS && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void c2a_ranges_kernel(const int *hist, int *queue)
{
  int pos = 0;
  for (int c = C2A_NCLASS - 1; c >= 0; --c) { queue[c] = pos; pos += hist[c]; queue[C2A_NCLASS + c] = pos; }
}

// Run-ahead results of the seed at processing position p of chain ci (slot = its index in processing order): the first seed of a
// chain comes from the first-seed waves, later seeds from wave B.  Returns false when there is none.
__device__ __forceinline__ bool spec_lookup(const C2AArgs &a, int64_t ci, int p, int64_t slot, const bwaflow_ext_result_t **L,
                                            const bwaflow_ext_result_t **R)
{
  if (p == 0) {
    if (!a.resL || a.cflag[ci]) return false;
    *L = a.resL + ci; *R = a.resR + ci;
    return true;
  }
  if (!a.slot_job) return false;
  const int j = a.slot_job[slot];
  if (j < 0 || a.cflag2[j]) return false;
  *L = a.resL2 + j; *R = a.resR2 + j;
  return true;
}

enum { PH_NEXT_SEED = 0, PH_LEFT = 1, PH_RIGHT_START = 2, PH_RIGHT = 3 };

template <int THREADS>
__global__ void __launch_bounds__(THREADS) c2a_kernel(const C2AArgs a)
{
  extern __shared__ uint32_t s_rows[];
  const int stride = THREADS;
  uint32_t *col = s_rows + threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int *head = a.queue + a.cls;
  const int q_end = a.queue[C2A_NCLASS + a.cls];
  const C2AOpt o = a.o;

  KswLane k;
  bool running = false, drained = false;
  int r = -1;                         // read owned by this lane (-1: idle)
  // read
  int lq = 0; const uint8_t *rd = nullptr; int64_t ci = 0, c_end = 0; bwaflow_alnreg_t *av = nullptr; int n_av = 0;
  // chain
  int64_t s_base = 0, wlo = 0, whi = 0; int n_seeds = 0, p = 0, rid = 0; float frac_rep = 0.f;
  // seed + region under construction
  int64_t s_rbeg = 0, a_rb = 0, a_re = 0, tpos = 0;
  int s_qbeg = 0, s_len = 0, a_score = 0, a_truesc = 0, a_qb = 0, a_qe = 0, aw0 = 0, aw1 = 0, sc0 = 0, prev = 0, tri = 0;
  int phase = PH_NEXT_SEED, tstep = 1;
  unsigned long long cells = 0;

  for (;;) {
    const unsigned stalled = __ballot_sync(full, !running);
    if (stalled == full || __popc(stalled) >= a.refill) {
      // ---- hand reads to idle lanes
      const unsigned idle = __ballot_sync(full, r < 0);
      if (idle && !drained) {
        const int n_idle = __popc(idle);
        int base = 0;
        if (lane == 0) base = atomicAdd(head, n_idle);
        base = __shfl_sync(full, base, 0);
        if (base + n_idle >= q_end) drained = true;
        if (r < 0) {
          const int slot = base + __popc(idle & ((1u << lane) - 1));
          if (slot < q_end) {
            r = (int)a.order[slot];
            const int64_t ro = a.read_off[r];
            lq = (int)(a.read_off[r + 1] - ro); rd = a.bases + ro;
            ci = a.chain_off[r] - 1; c_end = a.chain_off[r + 1];
            av = a.regs + a.sbase[a.chain_off[r]]; n_av = 0;
            n_seeds = 0; p = 0; phase = PH_NEXT_SEED;
            if (lq > a.qmax) { a.n_regs[r] = 0; r = -1; }           // reported through err by c2a_read_prep
          }
        }
      }
      // ---- driver: runs until this lane has an extension in flight or its read is finished
      if (r >= 0 && !running) {
        for (;;) {
          if (phase == PH_LEFT || phase == PH_RIGHT) {            // an extension just finished
            cells += k.cells;
            const int score = k.best;
            const int aw = phase == PH_LEFT ? aw0 : aw1;
            if (!(tri == 1 || score == prev || k.max_off < (aw >> 1) + (aw >> 2))) {
              // second band try (bwa/bwamem.c:732-743, 760-771)
              prev = score; tri = 1;
              int qlen, tlen, h0, eb;
              if (phase == PH_LEFT) { aw0 = o.w << 1; qlen = s_qbeg; tlen = (int)(s_rbeg - wlo); h0 = s_len * o.a; eb = o.pen_clip5; tpos = s_rbeg - 1; tstep = -1; }
              else { aw1 = o.w << 1; qlen = lq - (s_qbeg + s_len); tlen = (int)(whi - (s_rbeg + s_len)); h0 = sc0; eb = o.pen_clip3; tpos = s_rbeg + s_len; tstep = 1; }
              const uint8_t *qp = phase == PH_LEFT ? rd + s_qbeg - 1 : rd + s_qbeg + s_len;
              const int qs = tstep;
              if (!k.begin(col, stride, a.sc, qlen, tlen, phase == PH_LEFT ? aw0 : aw1, h0, eb, [&](int j) { return (uint32_t)qp[j * qs]; })) atomicExch(a.err, 2);
              if (tlen > 0) { running = true; break; }
              continue;
            }
            if (phase == PH_LEFT) {                                  // :745-751
              a_score = score;
              if (k.gscore <= 0 || k.gscore <= a_score - o.pen_clip5) { a_qb = s_qbeg - (k.best_j + 1); a_rb = s_rbeg - (k.best_i + 1); a_truesc = a_score; }
              else { a_qb = 0; a_rb = s_rbeg - (k.end_i + 1); a_truesc = k.gscore; }
              phase = PH_RIGHT_START;
            } else {                                                 // :773-779
              a_score = score;
              const int qe = s_qbeg + s_len;
              if (k.gscore <= 0 || k.gscore <= a_score - o.pen_clip3) { a_qe = qe + k.best_j + 1; a_re = s_rbeg + s_len + k.best_i + 1; a_truesc += a_score - sc0; }
              else { a_qe = lq; a_re = s_rbeg + s_len + k.end_i + 1; a_truesc += k.gscore - sc0; }
              phase = -1;                                            // region complete
            }
          }
          if (phase == PH_RIGHT_START) {
            if (s_qbeg + s_len != lq) {                              // :755-771
              sc0 = a_score; prev = a_score; tri = 0; aw1 = o.w;
              {
                const bwaflow_ext_result_t *xl, *xr;
                if (spec_lookup(a, ci, p, s_base + p, &xl, &xr) && (s_qbeg == 0 || xl->score == sc0)) {   // run ahead from this very h0
                  lane_take_result(k, *xr);
                  phase = PH_RIGHT;
                  continue;
                }
              }
              const int qlen = lq - (s_qbeg + s_len), tlen = (int)(whi - (s_rbeg + s_len));
              tpos = s_rbeg + s_len; tstep = 1;
              const uint8_t *qp = rd + s_qbeg + s_len;
              if (!k.begin(col, stride, a.sc, qlen, tlen, aw1, sc0, o.pen_clip3, [&](int j) { return (uint32_t)qp[j]; })) atomicExch(a.err, 2);
              phase = PH_RIGHT;
              if (tlen > 0) { running = true; break; }
              continue;
            }
            a_qe = lq; a_re = s_rbeg + s_len;                        // :780
            phase = -1;
          }
          if (phase == -1) {                                         // :783-792 append the region
            const bwaflow_seed_t *sd = a.sseeds + s_base;
            int cov = 0;
            for (int i = 0; i < n_seeds; ++i) {
              const bwaflow_seed_t t = sd[i];
              if (t.qbeg >= a_qb && t.qbeg + t.len <= a_qe && t.rbeg >= a_rb && t.rbeg + t.len <= a_re) cov += t.len;
            }
            bwaflow_alnreg_t g;
            memset(&g, 0, sizeof(g));
            g.rb = a_rb; g.re = a_re; g.qb = a_qb; g.qe = a_qe; g.rid = rid; g.score = a_score; g.truesc = a_truesc;
            g.w = aw0 > aw1 ? aw0 : aw1; g.seedcov = cov; g.seedlen0 = s_len; g.frac_rep = frac_rep;
            av[n_av++] = g;
            phase = PH_NEXT_SEED; ++p;
          }
          // ---- PH_NEXT_SEED: advance to the next seed that needs extending (:676-715)
          while (p >= n_seeds) {
            if (++ci >= c_end) break;
            const bwaflow_chain_t ch = a.chains[ci];
            n_seeds = ch.n_seeds; rid = ch.rid; frac_rep = ch.frac_rep; s_base = a.sbase[ci];
            wlo = a.win[2 * ci]; whi = a.win[2 * ci + 1]; p = 0;
          }
          if (ci >= c_end) { a.n_regs[r] = n_av; r = -1; break; }    // read finished
          {
            const bwaflow_seed_t s = a.sseeds[s_base + p];
            s_rbeg = s.rbeg; s_qbeg = s.qbeg; s_len = s.len;
            int i;
            for (i = 0; i < n_av; ++i) {                             // :680-696
              const bwaflow_alnreg_t *q = &av[i];
              const int64_t prb = q->rb, pre = q->re; const int pqb = q->qb, pqe = q->qe;
              if (s_rbeg < prb || s_rbeg + s_len > pre || s_qbeg < pqb || s_qbeg + s_len > pqe) continue;
              if ((double)(s_len - q->seedlen0) > .1 * lq) continue;
              const int pw = q->w;
              int qd = s_qbeg - pqb; int64_t rdd = s_rbeg - prb;
              int mg = max_gap(o, qd < rdd ? qd : (int)rdd);
              int w = mg < pw ? mg : pw;
              if (qd - rdd < w && rdd - qd < w) break;
              qd = pqe - (s_qbeg + s_len); rdd = pre - (s_rbeg + s_len);
              mg = max_gap(o, qd < rdd ? qd : (int)rdd);
              w = mg < pw ? mg : pw;
              if (qd - rdd < w && rdd - qd < w) break;
            }
            if (i < n_av) {                                          // :697-712
              const bwaflow_seed_t *sd = a.sseeds + s_base;
              int e;
              for (e = p - 1; e >= 0; --e) {
                if (a.skipped[s_base + e]) continue;
                const bwaflow_seed_t t = sd[e];
                if ((double)t.len < s_len * .95) continue;
                if (s_qbeg <= t.qbeg && s_qbeg + s_len - t.qbeg >= s_len >> 2 && t.qbeg - s_qbeg != t.rbeg - s_rbeg) break;
                if (t.qbeg <= s_qbeg && t.qbeg + t.len - s_qbeg >= s_len >> 2 && s_qbeg - t.qbeg != s_rbeg - t.rbeg) break;
              }
              if (e < 0) { a.skipped[s_base + p] = 1; ++p; continue; }
            }
            aw0 = aw1 = o.w; a_score = a_truesc = -1;                // :717-721
            if (s_qbeg) {                                            // :724-743
              prev = -1; tri = 0;
              {
                const bwaflow_ext_result_t *xl, *xr;
                if (spec_lookup(a, ci, p, s_base + p, &xl, &xr)) {    // this call was run ahead
                  lane_take_result(k, *xl);
                  phase = PH_LEFT;
                  continue;
                }
              }
              const int tlen = (int)(s_rbeg - wlo);
              tpos = s_rbeg - 1; tstep = -1;
              const uint8_t *qp = rd + s_qbeg - 1;
              if (!k.begin(col, stride, a.sc, s_qbeg, tlen, aw0, s_len * o.a, o.pen_clip5, [&](int j) { return (uint32_t)qp[-j]; })) atomicExch(a.err, 2);
              phase = PH_LEFT;
              if (tlen > 0) { running = true; break; }
              continue;
            }
            a_score = a_truesc = s_len * o.a; a_qb = 0; a_rb = s_rbeg;   // :753
            phase = PH_RIGHT_START;
          }
        }
      }
      if (drained && __ballot_sync(full, r >= 0) == 0) break;
    }
    if (running) {
      const int t = ref_base(a.ix, tpos + (int64_t)k.i * tstep);
      if (k.row(col, stride, a.sc, t)) running = false;
    }
  }
  if (cells) atomicAdd(a.cells, cells);
}


// ---- speculative first-seed extension -----------------------------------------------------------
// The first seed of a chain (in processing order) is extended unless an earlier chain's region already contains it, and
// what its extension computes depends only on the seed, the read and the chain's reference window -- not on the regions
// produced so far.  So the left and the right flank of every chain's first seed are run ahead of time as two waves of
// plain ksw_extend2 jobs on the packed batch engine (ext.cu), and c2a_replay_kernel then walks every read in the
// reference's order consuming those results.  A read that needs anything else -- a second band try, a later seed of a
// chain that is not contained, a job the packed form cannot hold -- is left to the exact lane-per-read kernel above.

// side 0 = left flank (reversed query prefix vs reversed reference), 1 = right flank.  Flank sizes of chain c's first seed.
__device__ __forceinline__ void spec_flank(int side, const bwaflow_seed_t &s, int lq, int64_t wlo, int64_t whi, int *qlen, int *tlen)
{
  if (side == 0) { *qlen = s.qbeg; *tlen = (int)(s.rbeg - wlo); }
  else { *qlen = lq - (s.qbeg + s.len); *tlen = (int)(whi - (s.rbeg + s.len)); }
  if (*tlen < 0) *tlen = 0;
}

// What c2a_spec_fill needs to know about chain c's job, gathered by one thread per chain so that the warp-per-chain fill kernel
// starts from two loads instead of a chain of six dependent ones (it is latency-bound: 1.7M warps with little work each).
struct SpecDesc { int64_t q_src, tpos; int32_t qlen, tlen, h0; int8_t qstep, tstep, nospec, end_bonus; };
static_assert(sizeof(SpecDesc) == 32, "SpecDesc is allocated as 32 bytes per chain");

// Job j of a wave: the first seed of chain j (jslot == null: the first-seed waves) or the seed at slot jslot[j] (wave B).
__global__ void c2a_spec_size(int side, const bwaflow_chain_t *chains, int64_t n_jobs, const int32_t *jslot, const int32_t *slot_chain,
                              const int64_t *chain_read, const int64_t *read_off,
                              const int64_t *sbase, const bwaflow_seed_t *sseeds, const int64_t *win, DevScore sc, C2AOpt o,
                              const bwaflow_ext_result_t *resL, const uint8_t *cflag, SpecDesc *desc, int64_t *jsz)
{
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_jobs) return;
  SpecDesc d;
  memset(&d, 0, sizeof(d));
  d.nospec = 1; d.h0 = 1; d.qstep = d.tstep = 1;
  if (j < n_jobs) {
    const int64_t c = jslot ? slot_chain[jslot[j]] : j;
    if (chains[c].n_seeds > 0 && !(side == 1 && cflag[j])) {
      const int64_t r = chain_read[c];
      const int64_t ro = read_off[r];
      const int lq = (int)(read_off[r + 1] - ro);
      const bwaflow_seed_t s = sseeds[jslot ? (int64_t)jslot[j] : sbase[c]];
      int qlen, tlen;
      spec_flank(side, s, lq, win[2 * c], win[2 * c + 1], &qlen, &tlen);
      if (qlen == 0) tlen = 0;                                  // no call at all (bwa/bwamem.c:753, 780)
      d.qlen = qlen; d.tlen = tlen;
      if (side == 0) { d.h0 = s.len * o.a; d.end_bonus = (int8_t)o.pen_clip5; d.q_src = ro + s.qbeg - 1; d.qstep = -1; d.tpos = s.rbeg - 1; d.tstep = -1; }
      else { d.h0 = s.qbeg ? resL[j].score : s.len * o.a; d.end_bonus = (int8_t)o.pen_clip3; d.q_src = ro + s.qbeg + s.len; d.tpos = s.rbeg + s.len; }
      d.nospec = d.h0 <= 0 || d.h0 + qlen * sc.smax > KSW_MAX_SCORE || o.pen_clip5 > 127 || o.pen_clip3 > 127;
    }
  }
  if (d.nospec) { d.qlen = 0; d.tlen = 0; d.h0 = 1; }
  if (j < n_jobs) desc[j] = d;
  jsz[j] = d.qlen; jsz[n_jobs + 1 + j] = d.tlen;
}

// one warp per chain: job record and its query / target bytes
__global__ void c2a_spec_fill(int side, int64_t n_chains, const SpecDesc *desc, const uint8_t *bases, const int64_t *joff, DevIndex ix,
                              DevScore sc, C2AOpt o, ExtDeviceView x, uint8_t *cflag)
{
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= n_chains) return;
  SpecDesc d = desc[c];
  const int64_t qo = joff[c], to = joff[n_chains + 1 + c];
  if (!d.nospec && (qo + d.qlen > x.max_q || to + d.tlen > x.max_t)) { d.nospec = 1; d.qlen = 0; d.tlen = 0; d.h0 = 1; }
  if (lane == 0) {
    bwaflow_ext_job_t jb;
    memset(&jb, 0, sizeof(jb));
    jb.w = o.w; jb.zdrop = sc.zdrop; jb.h0 = d.h0; jb.end_bonus = d.end_bonus; jb.qlen = d.qlen; jb.tlen = d.tlen;
    if (!d.nospec) { jb.q_off = qo; jb.t_off = to; }
    x.jobs[c] = jb;
    if (d.nospec) cflag[c] = 1;
  }
  const uint8_t *qp = bases + d.q_src;
  for (int j = lane; j < d.qlen; j += 32) x.q[qo + j] = qp[j * d.qstep];
  for (int j = lane; j < d.tlen; j += 32) x.t[to + j] = (uint8_t)ref_base(ix, d.tpos + (int64_t)j * d.tstep);
}

// One thread per read: mem_chain2aln's seed loop (bwa/bwamem.c:676-792) in the reference's order, with the two extensions of
// a chain's first seed taken from the speculative waves.  Sets strag[r] (and leaves no regions) when the read needs more.
// WARP = false: the read on the calling thread.  WARP = true: all 32 lanes of the warp work on the same read (a read inside a repeat
// has hundreds of chains, and the containment scan over the regions found so far makes the loop quadratic): every lane follows the
// same control flow on the same values, the scan for the first containing region is split 32 regions at a time, lane 0 stores.
template <bool WARP>
__device__ void c2a_replay_read(const C2AArgs &a, int64_t r, uint8_t *skipped, uint8_t *strag, int lane)
{
  const unsigned full = 0xffffffffu;
  const bool writer = !WARP || lane == 0;
  const C2AOpt o = a.o;
  const int lq = (int)(a.read_off[r + 1] - a.read_off[r]);
  bwaflow_alnreg_t *av = a.regs + a.sbase[a.chain_off[r]];
  int n_av = 0;
  bool bail = lq > c_c2a_qmax[C2A_NCLASS - 1];
  for (int64_t ci = a.chain_off[r]; ci < a.chain_off[r + 1] && !bail; ++ci) {
    const bwaflow_chain_t ch = a.chains[ci];
    const int64_t s_base = a.sbase[ci];
    const bwaflow_seed_t *sd = a.sseeds + s_base;
    for (int p = 0; p < ch.n_seeds; ++p) {
      const bwaflow_seed_t s = sd[p];
      auto contains = [&](int i) -> bool {                     // :680-696: region i contains the seed and the seed is on its diagonal band
        const bwaflow_alnreg_t *q = &av[i];
        const int64_t prb = q->rb, pre = q->re; const int pqb = q->qb, pqe = q->qe;
        if (s.rbeg < prb || s.rbeg + s.len > pre || s.qbeg < pqb || s.qbeg + s.len > pqe) return false;
        if ((double)(s.len - q->seedlen0) > .1 * lq) return false;
        const int pw = q->w;
        int qd = s.qbeg - pqb; int64_t rdd = s.rbeg - prb;
        int mg = max_gap(o, qd < rdd ? qd : (int)rdd);
        int w = mg < pw ? mg : pw;
        if (qd - rdd < w && rdd - qd < w) return true;
        qd = pqe - (s.qbeg + s.len); rdd = pre - (s.rbeg + s.len);
        mg = max_gap(o, qd < rdd ? qd : (int)rdd);
        w = mg < pw ? mg : pw;
        return qd - rdd < w && rdd - qd < w;
      };
      bool found = false;
      if (WARP) {
        for (int base = 0; base < n_av && !found; base += 32)
          found = __ballot_sync(full, base + lane < n_av && contains(base + lane)) != 0;
      } else {
        for (int i = 0; i < n_av && !found; ++i) found = contains(i);
      }
      if (found) {                                             // :697-712
        int e;
        for (e = p - 1; e >= 0; --e) {
          if (skipped[s_base + e]) continue;
          const bwaflow_seed_t t = sd[e];
          if ((double)t.len < s.len * .95) continue;
          if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
          if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
        }
        if (e < 0) {
          if (writer) skipped[s_base + p] = 1;
          if (WARP) __syncwarp();
          continue;
        }
      }
      const bwaflow_ext_result_t *xl, *xr;
      if (!spec_lookup(a, ci, p, s_base + p, &xl, &xr)) { bail = true; break; }   // not run ahead
      const int wtest = (o.w >> 1) + (o.w >> 2);
      int a_score, a_truesc, a_qb, a_qe; int64_t a_rb, a_re;
      if (s.qbeg) {                                            // :724-751
        const bwaflow_ext_result_t x = *xl;
        if (!(x.max_off < wtest)) { bail = true; break; }      // would take the second band try
        a_score = x.score;
        if (x.gscore <= 0 || x.gscore <= a_score - o.pen_clip5) { a_qb = s.qbeg - x.qle; a_rb = s.rbeg - x.tle; a_truesc = a_score; }
        else { a_qb = 0; a_rb = s.rbeg - x.gtle; a_truesc = x.gscore; }
      } else { a_score = a_truesc = s.len * o.a; a_qb = 0; a_rb = s.rbeg; }
      if (s.qbeg + s.len != lq) {                              // :755-779
        const bwaflow_ext_result_t x = *xr;
        const int sc0 = a_score;
        if (!(x.score == sc0 || x.max_off < wtest)) { bail = true; break; }
        a_score = x.score;
        const int qe = s.qbeg + s.len;
        if (x.gscore <= 0 || x.gscore <= a_score - o.pen_clip3) { a_qe = qe + x.qle; a_re = s.rbeg + s.len + x.tle; a_truesc += a_score - sc0; }
        else { a_qe = lq; a_re = s.rbeg + s.len + x.gtle; a_truesc += x.gscore - sc0; }
      } else { a_qe = lq; a_re = s.rbeg + s.len; }
      int cov = 0;                                             // :783-792
      for (int k = 0; k < ch.n_seeds; ++k) {
        const bwaflow_seed_t t = sd[k];
        if (t.qbeg >= a_qb && t.qbeg + t.len <= a_qe && t.rbeg >= a_rb && t.rbeg + t.len <= a_re) cov += t.len;
      }
      if (writer) {
        bwaflow_alnreg_t g;
        memset(&g, 0, sizeof(g));
        g.rb = a_rb; g.re = a_re; g.qb = a_qb; g.qe = a_qe; g.rid = ch.rid; g.score = a_score; g.truesc = a_truesc;
        g.w = o.w; g.seedcov = cov; g.seedlen0 = s.len; g.frac_rep = ch.frac_rep;
        av[n_av] = g;
      }
      ++n_av;
      if (WARP) __syncwarp();
    }
  }
  if (writer) {
    strag[r] = bail ? 1 : 0;
    if (bail) atomicAdd(a.err + 1, 1);                     // reads left to the exact kernel (reported, not an error; reset before a second replay)
    a.n_regs[r] = bail ? 0 : n_av;
  }
}

// `again`: second replay, only for the reads the first one left (strag[r] != 0), now with the wave-B results in `a`.
constexpr int C2A_REPLAY_WARP_MIN = 64;     // seeds of a read from which the warp replays it together
__global__ void c2a_replay_kernel(C2AArgs a, int64_t n_reads, uint8_t *skipped, uint8_t *strag, int again)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const bool act = r < n_reads && !(again && !strag[r]);
  bool heavy = false;
  if (act) {
    heavy = a.sbase[a.chain_off[r + 1]] - a.sbase[a.chain_off[r]] >= C2A_REPLAY_WARP_MIN;
    if (!heavy) c2a_replay_read<false>(a, r, skipped, strag, lane);
  }
  for (unsigned hm = __ballot_sync(full, heavy); hm; hm &= hm - 1) {
    const int64_t rr = __shfl_sync(full, (long long)r, __ffs(hm) - 1);
    c2a_replay_read<true>(a, rr, skipped, strag, lane);
    __syncwarp();
  }
}

// wave B candidates: every later seed of the chains of the reads the first replay left
__global__ void c2a_cand_kernel(int64_t n_slots, const int32_t *slot_chain, const int64_t *chain_read, const int64_t *sbase, const uint8_t *strag,
                                uint8_t *cand, int32_t *slot_job)
{
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int64_t c = slot_chain[s];
  cand[s] = strag[chain_read[c]] && s != sbase[c];
  slot_job[s] = -1;
}

__global__ void c2a_slot_job_kernel(const int32_t *n_jobs, int64_t cap, const int32_t *jslot, int32_t *slot_job)
{
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < *n_jobs && j < cap) slot_job[jslot[j]] = (int32_t)j;
}

__global__ void c2a_count_strag(const uint8_t *strag, int64_t n_reads, int32_t *out)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = __ballot_sync(0xffffffffu, r < n_reads && strag[r]);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, __popc(m));
}

__global__ void c2a_compact(const int64_t *chain_off, const int64_t *sbase, const int *n_regs, const int64_t *reg_off,
                            int64_t n_reads, const bwaflow_alnreg_t *regs, bwaflow_alnreg_t *out)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const bwaflow_alnreg_t *src = regs + sbase[chain_off[r]];
  bwaflow_alnreg_t *dst = out + reg_off[r];
  for (int i = 0; i < n_regs[r]; ++i) dst[i] = src[i];
}

__global__ void c2a_widen(const int *n_regs, int64_t n, int64_t *out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = n_regs[i];
  if (i == n) out[i] = 0;
}

} // namespace bwaflow

using namespace bwaflow;


static size_t max3(size_t a, size_t b, size_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

// The left and the right flank of n jobs (job j = first seed of chain j, or the seed at slot jslot[j]) as two waves on an extension
// batch: sizes -> offsets -> job records and byte pools -> kernels.  Left results end in resL, right results stay in xb's output.
static int c2a_waves(bwaflow_c2a_batch *b, const C2AOpt &o, cudaStream_t st, bwaflow_ext_batch_t *xb, int64_t n, const int32_t *jslot,
                     uint8_t *cflag, bwaflow_ext_result_t *resL, void *jdesc, int64_t *jsz, int64_t *joff)
{
  const ExtDeviceView x = ext_device_view(xb);
  for (int side = 0; side < 2; ++side) {
    if (side == 0) BF_CUDA(cudaMemsetAsync(cflag, 0, n, st));
    c2a_spec_size<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(side, b->d_chains, n, jslot, b->d_slot_chain, b->d_chain_read, b->d_read_off, b->d_sbase,
                                                               b->d_sseeds, b->d_win, b->ctx->par.sc, o, resL, cflag, (SpecDesc *)jdesc, jsz);
    for (int k = 0; k < 2; ++k) {
      size_t tmp = b->tmp_bytes;
      BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, jsz + k * (n + 1), joff + k * (n + 1), (int)(n + 1), st));
    }
    c2a_spec_fill<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(side, n, (const SpecDesc *)jdesc, b->d_bases, joff, b->ctx->ix, b->ctx->par.sc, o, x, cflag);
    int r = ext_launch_device(xb, n, st);
    if (r) return r;
    if (side == 0) BF_CUDA(cudaMemcpyAsync(resL, x.out, n * sizeof(bwaflow_ext_result_t), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int bwaflow::c2a_launch(bwaflow_c2a_batch *b)
{
  const int64_t nr = b->n_reads, nc = b->n_chains, ns = b->n_seeds;
  cudaStream_t st = b->stream;
  int *hist = b->d_ctl, *queue = b->d_ctl + C2A_NCLASS, *err = b->d_ctl + 3 * C2A_NCLASS;
  BF_CUDA(cudaMemsetAsync(b->d_ctl, 0, (3 * C2A_NCLASS + 4) * sizeof(int), st));
  BF_CUDA(cudaMemsetAsync(b->d_reg_off, 0, (nr + 1) * sizeof(int64_t), st));
  if (nr == 0) return 0;
  const bwaflow_opt_t &op = b->ctx->par.opt;
  C2AOpt o{op.a, op.w, op.pen_clip5, op.pen_clip3, op.o_del, op.e_del, op.o_ins, op.e_ins};
  size_t tmp;
  // per-chain seed prefix
  c2a_nseeds<<<(unsigned)((nc + 1 + 255) / 256), 256, 0, st>>>(b->d_chains, nc, b->d_nseeds);
  tmp = b->tmp_bytes;
  BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_nseeds, b->d_sbase, (int)(nc + 1), st));
  c2a_mark_reads<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(b->d_chain_off, nr, b->d_chain_read);
  cub::DoubleBuffer<uint64_t> sk(b->d_skeys[0], b->d_skeys[1]);
  cub::DoubleBuffer<uint32_t> sv(b->d_svals[0], b->d_svals[1]);
  if (nc) {
    c2a_chain_prep<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(b->d_chains, nc, b->d_chain_read, b->d_read_off, b->d_seeds, b->d_sbase,
                                                               b->ctx->ix, o, b->d_win, b->d_skeys[0], b->d_svals[0], b->d_slot_chain, err);
    if (ns) {
      tmp = b->tmp_bytes;
      BF_CUDA(cub::DeviceRadixSort::SortPairs(b->d_tmp, tmp, sk, sv, (int)ns, 0, 64, st));
      c2a_gather_seeds<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(b->d_seeds, sv.Current(), ns, b->d_sseeds);
      BF_CUDA(cudaMemsetAsync(b->d_skipped, 0, ns, st));
    }
  }
  C2AArgs a;
  a.bases = b->d_bases; a.read_off = b->d_read_off; a.chains = b->d_chains; a.chain_off = b->d_chain_off;
  a.sbase = b->d_sbase; a.sseeds = b->d_sseeds; a.win = b->d_win; a.skipped = b->d_skipped; a.regs = b->d_regs;
  a.n_regs = b->d_nregs; a.order = nullptr; a.queue = queue; a.err = err;
  a.cells = (unsigned long long *)(b->d_ctl + 3 * C2A_NCLASS + 2);
  a.refill = b->refill; a.ix = b->ctx->ix; a.sc = b->ctx->par.sc; a.o = o;
  a.cls = 0; a.qmax = 0;
  a.cflag = nullptr; a.resL = a.resR = nullptr;
  a.slot_job = nullptr; a.cflag2 = nullptr; a.resL2 = a.resR2 = nullptr;
  // speculative waves + replay; what is left goes to the exact kernel below
  const uint8_t *only = nullptr;
  b->waveb_jobs = b->waveb_reads = 0;
  if (b->spec && !b->count_cells && nc > 0 && ns > 0) {
    const ExtDeviceView x = ext_device_view(b->xb);
    int r = c2a_waves(b, o, st, b->xb, nc, nullptr, b->d_cflag, b->d_resL, b->d_jdesc, b->d_jsz, b->d_joff);
    if (r) return r;
    a.cflag = b->d_cflag; a.resL = b->d_resL; a.resR = x.out;
    BF_CUDA(cudaMemsetAsync(b->d_skipped2, 0, ns, st));
    c2a_replay_kernel<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(a, nr, b->d_skipped2, b->d_strag, 0);
    only = b->d_strag;
    // Wave B.  A read the replay left (a later seed of a chain is not contained, typically one chain out of the hundreds of a
    // repeat read) would otherwise run ALL its missing extensions one after the other on one lane of the exact kernel.  Every
    // later seed of such reads is run ahead as well -- two more waves on the batch engine -- and the reads are replayed once
    // more; only second band tries and jobs the engine cannot hold are then left.  One small read-back decides whether there
    // is anything to do (nothing on the i.i.d. workloads).
    if (b->xb2) {
      BF_CUDA(cudaMemsetAsync(b->d_nsel, 0, 2 * sizeof(int32_t), st));
      c2a_count_strag<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(b->d_strag, nr, b->d_nsel + 1);
      c2a_cand_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(ns, b->d_slot_chain, b->d_chain_read, b->d_sbase, b->d_strag, b->d_cand, b->d_slot_job);
      tmp = b->tmp_bytes;
      BF_CUDA(cub::DeviceSelect::Flagged(b->d_tmp, tmp, cub::CountingInputIterator<int32_t>(0), b->d_cand, b->d_jslot, b->d_nsel, (int)ns, st));
      BF_CUDA(cudaMemcpyAsync(b->h_sel, b->d_nsel, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      BF_CUDA(cudaStreamSynchronize(st));
      const ExtDeviceView x2 = ext_device_view(b->xb2);
      int64_t nb = b->h_sel[0];
      b->waveb_reads = b->h_sel[1];
      if (nb > x2.max_jobs) nb = x2.max_jobs;
      if (b->waveb_reads > 0 && nb > 0) {
        b->waveb_jobs = nb;
        c2a_slot_job_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(b->d_nsel, nb, b->d_jslot, b->d_slot_job);
        r = c2a_waves(b, o, st, b->xb2, nb, b->d_jslot, b->d_cflag2, b->d_resL2, b->d_jdesc2, b->d_jsz2, b->d_joff2);
        if (r) return r;
        a.slot_job = b->d_slot_job; a.cflag2 = b->d_cflag2; a.resL2 = b->d_resL2; a.resR2 = x2.out;
        BF_CUDA(cudaMemsetAsync(err + 1, 0, sizeof(int), st));
        c2a_replay_kernel<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(a, nr, b->d_skipped2, b->d_strag, 1);
      }
    }
  }
  // read queue
  c2a_read_prep<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(b->d_read_off, b->d_chain_off, b->d_sbase, b->d_sseeds, nr, only, b->d_rkeys[0], b->d_rvals[0], hist, err);
  cub::DoubleBuffer<uint32_t> rk(b->d_rkeys[0], b->d_rkeys[1]), rv(b->d_rvals[0], b->d_rvals[1]);
  tmp = b->tmp_bytes;
  BF_CUDA(cub::DeviceRadixSort::SortPairsDescending(b->d_tmp, tmp, rk, rv, (int)nr, 0, 28, st));
  c2a_ranges_kernel<<<1, 1, 0, st>>>(hist, queue);
  a.order = rv.Current();
  for (int c = C2A_NCLASS - 1; c >= 0; --c) {
    a.cls = c; a.qmax = h_c2a_qmax[c];
    const size_t smem = (size_t)h_c2a_threads[c] * (h_c2a_qmax[c] + 1) * sizeof(uint32_t);
    if (h_c2a_threads[c] == 64) c2a_kernel<64><<<b->blocks[c], 64, smem, st>>>(a);
    else c2a_kernel<32><<<b->blocks[c], 32, smem, st>>>(a);
  }
  // compact
  c2a_widen<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, st>>>(b->d_nregs, nr, b->d_nregs64);
  tmp = b->tmp_bytes;
  BF_CUDA(cub::DeviceScan::ExclusiveSum(b->d_tmp, tmp, b->d_nregs64, b->d_reg_off, (int)(nr + 1), st));
  c2a_compact<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(b->d_chain_off, b->d_sbase, b->d_nregs, b->d_reg_off, nr, b->d_regs, b->d_regs_out);
  BF_CUDA(cudaGetLastError());
  return 0;
}

int bwaflow::c2a_finish_queue(bwaflow_c2a_batch *b)
{
  BF_CUDA(cudaEventRecord(b->tm.ev[2], b->stream));
  BF_CUDA(cudaMemcpyAsync(b->h_reg_off, b->d_reg_off, (b->n_reads + 1) * 8, cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl + 3 * C2A_NCLASS, 4 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  b->pending = true;
  return 0;
}

extern "C" {

int bwaflow_c2a_batch_create(bwaflow_ctx_t *ctx, int64_t max_reads, int64_t max_bases, int64_t max_chains,
                             int64_t max_seeds, bwaflow_c2a_batch_t **out)
{
  BF_REQUIRE(ctx && out, -22, "bwaflow_c2a_batch_create: null argument");
  BF_REQUIRE(ctx->has_index, -22, "bwaflow_c2a_batch_create: the context has no index");
  BF_REQUIRE(max_reads > 0 && max_reads < (1LL << 31) - 1 && max_chains < (1LL << 24) && max_seeds < (1LL << 31) - 1, -22,
             "bwaflow_c2a_batch_create: capacity out of range (reads < 2^31, chains < 2^24, seeds < 2^31)");
  BF_CUDA(cudaSetDevice(ctx->device));
  bwaflow_c2a_batch *b = (bwaflow_c2a_batch *)calloc(1, sizeof(*b));
  b->ctx = ctx; b->max_reads = max_reads; b->max_bases = max_bases; b->max_chains = max_chains; b->max_seeds = max_seeds;
  b->refill = 8;
  if (const char *s = getenv("BWAFLOW_C2A_REFILL")) b->refill = atoi(s) > 0 ? atoi(s) : 8;
  BF_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  if (b->tm.init()) return -5;
  const int64_t mc = max_chains + 1, ms = max_seeds + 1, mr = max_reads + 1;
  BF_CUDA(cudaMalloc(&b->d_bases, max_bases + 16));
  BF_CUDA(cudaMalloc(&b->d_read_off, mr * 8));
  BF_CUDA(cudaMalloc(&b->d_chain_off, mr * 8));
  BF_CUDA(cudaMalloc(&b->d_chains, mc * sizeof(bwaflow_chain_t)));
  BF_CUDA(cudaMalloc(&b->d_seeds, ms * sizeof(bwaflow_seed_t)));
  BF_CUDA(cudaMalloc(&b->d_sseeds, ms * sizeof(bwaflow_seed_t)));
  BF_CUDA(cudaMalloc(&b->d_nseeds, (mc + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_sbase, (mc + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_chain_read, mc * 8));
  BF_CUDA(cudaMalloc(&b->d_win, mc * 16));
  for (int i = 0; i < 2; ++i) {
    BF_CUDA(cudaMalloc(&b->d_skeys[i], ms * 8));
    BF_CUDA(cudaMalloc(&b->d_svals[i], ms * 4));
    BF_CUDA(cudaMalloc(&b->d_rkeys[i], mr * 4));
    BF_CUDA(cudaMalloc(&b->d_rvals[i], mr * 4));
  }
  size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  {
    cub::DoubleBuffer<uint64_t> sk(b->d_skeys[0], b->d_skeys[1]);
    cub::DoubleBuffer<uint32_t> sv(b->d_svals[0], b->d_svals[1]), rk(b->d_rkeys[0], b->d_rkeys[1]), rv(b->d_rvals[0], b->d_rvals[1]);
    BF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, sk, sv, (int)ms, 0, 64, b->stream));
    BF_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, t2, rk, rv, (int)mr, 0, 28, b->stream));
    BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t3, b->d_nseeds, b->d_sbase, (int)(mc + 1), b->stream));
    BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t4, b->d_nseeds, b->d_sbase, (int)(mr + 1), b->stream));
    size_t t5 = 0, t6 = 0;
    BF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t5, b->d_nseeds, b->d_sbase, (int)(ms + 1), b->stream));
    BF_CUDA(cub::DeviceSelect::Flagged(nullptr, t6, cub::CountingInputIterator<int32_t>(0), (uint8_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)ms, b->stream));
    t4 = max3(t4, t5, t6);
  }
  b->tmp_bytes = max3(t1, t2, t3 > t4 ? t3 : t4) + 256;
  BF_CUDA(cudaMalloc(&b->d_tmp, b->tmp_bytes));
  BF_CUDA(cudaMalloc(&b->d_skipped, ms));
  BF_CUDA(cudaMalloc(&b->d_skipped2, ms));
  BF_CUDA(cudaMalloc(&b->d_strag, mr));
  BF_CUDA(cudaMalloc(&b->d_cflag, mc));
  BF_CUDA(cudaMalloc(&b->d_jsz, 2 * (mc + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_joff, 2 * (mc + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_resL, mc * sizeof(bwaflow_ext_result_t)));
  BF_CUDA(cudaMalloc(&b->d_jdesc, mc * 32));
  b->spec = true;
  if (const char *s = getenv("BWAFLOW_C2A_SPEC")) b->spec = atoi(s) != 0;
  if (b->spec) {
    // byte pools of the speculative jobs: about twice the mean flank (44 / 131 bases at 2x150 bp) per chain; a chain whose
    // job does not fit is simply not run ahead
    int r = bwaflow_ext_batch_create(ctx, mc, 96 * mc + 4096, 288 * mc + 4096, &b->xb);
    if (r) return r;
    // every class the two-columns-per-word engine can hold (qlen <= 255): no loss on 2x150 bp, stage 4 27.4 -> 23.9 ms on 2x250 bp
    ext_set_pair_classes(b->xb, 6);
    ext_set_order_rows(b->xb, 1);
  }
  BF_CUDA(cudaMalloc(&b->d_slot_chain, ms * 4));
  if (b->spec && !(getenv("BWAFLOW_C2A_WAVEB") && atoi(getenv("BWAFLOW_C2A_WAVEB")) == 0)) {
    // wave B: up to one job per seed slot, pools as above
    int r = bwaflow_ext_batch_create(ctx, ms, 96 * ms + 4096, 288 * ms + 4096, &b->xb2);
    if (r) return r;
    ext_set_pair_classes(b->xb2, 6);
    ext_set_order_rows(b->xb2, 1);
    BF_CUDA(cudaMalloc(&b->d_resL2, ms * sizeof(bwaflow_ext_result_t)));
    BF_CUDA(cudaMalloc(&b->d_jslot, ms * 4));
    BF_CUDA(cudaMalloc(&b->d_slot_job, ms * 4));
    BF_CUDA(cudaMalloc(&b->d_cand, ms));
    BF_CUDA(cudaMalloc(&b->d_cflag2, ms));
    BF_CUDA(cudaMalloc(&b->d_jsz2, 2 * (ms + 1) * 8));
    BF_CUDA(cudaMalloc(&b->d_joff2, 2 * (ms + 1) * 8));
    BF_CUDA(cudaMalloc(&b->d_jdesc2, ms * 32));
    BF_CUDA(cudaMalloc(&b->d_nsel, 2 * sizeof(int32_t)));
    BF_CUDA(cudaMallocHost(&b->h_sel, 2 * sizeof(int)));
  }
  BF_CUDA(cudaMalloc(&b->d_regs, ms * sizeof(bwaflow_alnreg_t)));
  BF_CUDA(cudaMalloc(&b->d_regs_out, ms * sizeof(bwaflow_alnreg_t)));
  BF_CUDA(cudaMallocHost(&b->h_regs, ms * sizeof(bwaflow_alnreg_t)));
  BF_CUDA(cudaMalloc(&b->d_nregs, mr * 4));
  BF_CUDA(cudaMalloc(&b->d_nregs64, (mr + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_reg_off, (mr + 1) * 8));
  BF_CUDA(cudaMallocHost(&b->h_reg_off, (mr + 1) * 8));
  BF_CUDA(cudaMalloc(&b->d_ctl, (3 * C2A_NCLASS + 4) * sizeof(int)));
  BF_CUDA(cudaMallocHost(&b->h_ctl, 4 * sizeof(int)));
  for (int c = 0; c < C2A_NCLASS; ++c) {
    const size_t smem = (size_t)h_c2a_threads[c] * (h_c2a_qmax[c] + 1) * sizeof(uint32_t);
    int per_sm = 0;
    if (h_c2a_threads[c] == 64) {
      BF_CUDA(cudaFuncSetAttribute(c2a_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c2a_kernel<64>, 64, smem));
    } else {
      BF_CUDA(cudaFuncSetAttribute(c2a_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c2a_kernel<32>, 32, smem));
    }
    BF_REQUIRE(per_sm > 0, -12, "chain2aln kernel class %d does not fit on an SM", c);
    b->blocks[c] = per_sm * ctx->n_sm;
  }
  *out = b;
  return 0;
}

void bwaflow_c2a_batch_destroy(bwaflow_c2a_batch_t *b)
{
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->stream);
  cudaFree(b->d_bases); cudaFree(b->d_read_off); cudaFree(b->d_chain_off); cudaFree(b->d_chains); cudaFree(b->d_seeds);
  cudaFree(b->d_sseeds); cudaFree(b->d_nseeds); cudaFree(b->d_sbase); cudaFree(b->d_chain_read); cudaFree(b->d_win);
  for (int i = 0; i < 2; ++i) { cudaFree(b->d_skeys[i]); cudaFree(b->d_svals[i]); cudaFree(b->d_rkeys[i]); cudaFree(b->d_rvals[i]); }
  if (b->xb) bwaflow_ext_batch_destroy(b->xb);
  if (b->xb2) bwaflow_ext_batch_destroy(b->xb2);
  cudaFree(b->d_slot_chain); cudaFree(b->d_resL2); cudaFree(b->d_jslot); cudaFree(b->d_slot_job); cudaFree(b->d_cand); cudaFree(b->d_cflag2);
  cudaFree(b->d_jsz2); cudaFree(b->d_joff2); cudaFree(b->d_jdesc2); cudaFree(b->d_nsel); if (b->h_sel) cudaFreeHost(b->h_sel);
  cudaFree(b->d_skipped2); cudaFree(b->d_strag); cudaFree(b->d_cflag); cudaFree(b->d_jsz); cudaFree(b->d_joff); cudaFree(b->d_resL); cudaFree(b->d_jdesc);
  cudaFree(b->d_tmp); cudaFree(b->d_skipped); cudaFree(b->d_regs); cudaFree(b->d_regs_out); cudaFreeHost(b->h_regs);
  cudaFree(b->d_nregs); cudaFree(b->d_nregs64); cudaFree(b->d_reg_off); cudaFreeHost(b->h_reg_off);
  cudaFree(b->d_ctl); cudaFreeHost(b->h_ctl);
  b->tm.destroy();
  cudaStreamDestroy(b->stream);
  free(b);
}


int bwaflow_c2a_submit(bwaflow_c2a_batch_t *b, int64_t n_reads, const uint8_t *bases, const int64_t *read_off,
                       const bwaflow_chain_t *chains, const int64_t *chain_off,
                       const bwaflow_seed_t *seeds, int64_t n_seeds)
{
  BF_REQUIRE(b && (n_reads == 0 || (read_off && chain_off)), -22, "bwaflow_c2a_submit: null argument");
  const int64_t n_bases = n_reads ? read_off[n_reads] : 0, n_chains = n_reads ? chain_off[n_reads] : 0;
  BF_REQUIRE(n_reads >= 0 && n_reads <= b->max_reads && n_bases <= b->max_bases && n_chains <= b->max_chains && n_seeds <= b->max_seeds, -7,
             "bwaflow_c2a_submit: batch capacity exceeded (%lld reads, %lld bases, %lld chains, %lld seeds)",
             (long long)n_reads, (long long)n_bases, (long long)n_chains, (long long)n_seeds);
  BF_REQUIRE(n_reads == 0 || (read_off[0] == 0 && chain_off[0] == 0), -22, "bwaflow_c2a_submit: offsets must start at 0");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  int64_t used = 0;
  for (int64_t c = 0; c < n_chains; ++c) {
    BF_REQUIRE(chains[c].n_seeds >= 0 && chains[c].seed_off >= 0 && chains[c].seed_off + chains[c].n_seeds <= n_seeds, -22,
               "bwaflow_c2a_submit: chain %lld references seeds outside [0,%lld)", (long long)c, (long long)n_seeds);
    used += chains[c].n_seeds;
  }
  BF_REQUIRE(used <= b->max_seeds, -7, "bwaflow_c2a_submit: chains reference %lld seeds, capacity %lld", (long long)used, (long long)b->max_seeds);
  b->n_reads = n_reads; b->n_chains = n_chains; b->n_seeds = used;
  cudaStream_t st = b->stream;
  BF_CUDA(cudaEventRecord(b->tm.ev[0], st));
  if (n_reads) {
    if (bases) {   // bases == NULL: the reads of the previous submit are still resident and are reused
      BF_CUDA(cudaMemcpyAsync(b->d_bases, bases, n_bases, cudaMemcpyHostToDevice, st));
      BF_CUDA(cudaMemcpyAsync(b->d_read_off, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    BF_CUDA(cudaMemcpyAsync(b->d_chain_off, chain_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_chains) BF_CUDA(cudaMemcpyAsync(b->d_chains, chains, n_chains * sizeof(bwaflow_chain_t), cudaMemcpyHostToDevice, st));
    if (n_seeds) BF_CUDA(cudaMemcpyAsync(b->d_seeds, seeds, n_seeds * sizeof(bwaflow_seed_t), cudaMemcpyHostToDevice, st));
  }
  BF_CUDA(cudaEventRecord(b->tm.ev[1], st));
  b->count_cells = false; b->cells_known = 0.0;
  int r = c2a_launch(b);
  if (r) return r;
  return c2a_finish_queue(b);
}

static int c2a_rerun(bwaflow_c2a_batch_t *b, bool counted)
{
  BF_REQUIRE(b, -22, "bwaflow_c2a_rerun: null batch");
  BF_CUDA(cudaSetDevice(b->ctx->device));
  BF_CUDA(cudaEventRecord(b->tm.ev[0], b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[1], b->stream));
  b->count_cells = counted;
  int r = c2a_launch(b);
  if (r) return r;
  return c2a_finish_queue(b);
}

int bwaflow_c2a_second_round(bwaflow_c2a_batch_t *b, int64_t *n_reads, int64_t *n_jobs)
{
  BF_REQUIRE(b && !b->pending, -22, "bwaflow_c2a_second_round: call after collect");
  if (n_reads) *n_reads = b->waveb_reads;
  if (n_jobs) *n_jobs = b->waveb_jobs;
  return 0;
}

int bwaflow_c2a_exact_reads(bwaflow_c2a_batch_t *b, int64_t *n_exact)
{
  BF_REQUIRE(b && n_exact && !b->pending, -22, "bwaflow_c2a_exact_reads: call after collect");
  *n_exact = (b->spec && !b->count_cells) ? (int64_t)b->h_ctl[1] : b->n_reads;
  return 0;
}

int bwaflow_c2a_rerun(bwaflow_c2a_batch_t *b) { return c2a_rerun(b, false); }
int bwaflow_c2a_rerun_counted(bwaflow_c2a_batch_t *b) { return c2a_rerun(b, true); }

int bwaflow_c2a_collect(bwaflow_c2a_batch_t *b, const bwaflow_alnreg_t **regs, const int64_t **reg_off,
                        int64_t *n_regs, bwaflow_timing_t *t)
{
  BF_REQUIRE(b && b->pending, -22, "bwaflow_c2a_collect: nothing submitted");
  BF_CUDA(cudaStreamSynchronize(b->stream));
  b->pending = false;
  const int err = b->h_ctl[0];
  BF_REQUIRE(err != 1, -34, "read longer than %d bases", h_c2a_qmax[C2A_NCLASS - 1]);
  BF_REQUIRE(err != 2, -34, "extension whose scores may exceed %d (h0 + qlen*max_mat)", KSW_MAX_SCORE);
  BF_REQUIRE(err != 3, -34, "chain/seed counts or seed scores exceed the sort-key fields (2^24 chains, 2^20 seeds per chain, score < 2^20)");
  const int64_t total = b->n_reads ? b->h_reg_off[b->n_reads] : 0;
  if (total) BF_CUDA(cudaMemcpyAsync(b->h_regs, b->d_regs_out, total * sizeof(bwaflow_alnreg_t), cudaMemcpyDeviceToHost, b->stream));
  BF_CUDA(cudaEventRecord(b->tm.ev[3], b->stream));
  BF_CUDA(cudaStreamSynchronize(b->stream));
  if (t) {
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->h2d_ms, b->tm.ev[0], b->tm.ev[1]);
    cudaEventElapsedTime(&t->kernel_ms, b->tm.ev[1], b->tm.ev[2]);
    cudaEventElapsedTime(&t->d2h_ms, b->tm.ev[2], b->tm.ev[3]);
    cudaEventElapsedTime(&t->total_ms, b->tm.ev[0], b->tm.ev[3]);
    // kernels per launch of this stage (cub calls counted by the kernels they run): prep 6 + seed sort 10 + gather 1 +
    // read queue 8 + C2A_NCLASS exact kernels + compaction 4 = 37; the speculative waves add 2 x (size 1 + scans 4 + fill 1 +
    // extension batch: key 1 + sort 6 + ranges 1 + 8 class kernels + deferred queue 1) + replay 1 = 45
    t->n_launches = b->n_reads ? 29 + C2A_NCLASS + ((b->spec && !b->count_cells) ? 45 : 0) : 0;
    if (b->count_cells || !b->spec) b->cells_known = (double)*(unsigned long long *)(b->h_ctl + 2);
    t->work_units = b->cells_known;        // DP cells the reference executes for this input (0 = not counted yet)
  }
  if (regs) *regs = b->h_regs;
  if (reg_off) *reg_off = b->h_reg_off;
  if (n_regs) *n_regs = total;
  return 0;
}

} // extern "C"
