/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's hot-path arithmetic, written from the
 * behaviour of the reference (each function cites the file:line it follows).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this file's library; the product path (bwaflow_b200/, include/) never
 * does and fails loudly without its CUDA library.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks every function here
 * against oracle/_ref (the reference's own bwa/ *.c + src/bwa_wrapper.cpp
 * compiled in place) and against the committed fixtures in tests/golden/
 * that oracle/_ref produced (tests/golden/make_golden.py).
 *
 * Besides results, the restatement counts the work units SURVEY.md 8(d)
 * defines: executed DP cells for extension, occ-block touches for seeding.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t primary, L2[5], seq_len, bwt_size;
  const uint32_t *bwt;          /* occ-interleaved, 64-byte blocks (bwa/bwt.h:72-73) */
  uint64_t n_sa; int64_t sa_intv; const uint64_t *sa;
  int64_t l_pac; const uint8_t *pac; int32_t n_seqs, pad;
} orc_index_t;                  /* same layout as ref_index_view_t in oracle/ref_driver.cpp */

typedef struct { uint64_t x[3], info; } orc_intv_t;               /* bwtintv_t, bwa/bwt.h:60-62 */
typedef struct { int64_t rbeg; int32_t qbeg, len, score, pad; } orc_seed_t;
typedef struct { int64_t pos; int32_t n_seeds, rid; float frac_rep; uint32_t flags; int64_t seed_off; } orc_chain_t;
typedef struct {                                                  /* mem_alnreg_t, bwa/bwamem.h:60-78 (88 B) */
  int64_t rb, re; int32_t qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary,
  secondary_all, seedlen0; uint32_t n_comp_is_alt; float frac_rep; uint64_t hash;
} orc_alnreg_t;
typedef struct {
  int32_t a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop;
  int32_t min_seed_len, split_width, max_occ, max_mem_intv;
  float split_factor; int8_t mat[25]; int8_t pad[3];
} orc_opt_t;

static uint64_t g_blocks;   /* occ-block (64 B) touches since the last reset */
static uint64_t g_cells;    /* DP cells executed since the last reset */
void orc_counters_reset(void) { g_blocks = 0; g_cells = 0; }
uint64_t orc_blocks(void) { return g_blocks; }
uint64_t orc_cells(void) { return g_cells; }

/* ------------------------------------------------------------------ FM index */

/* Occ of all four bases among BWT positions [0,k] -- bwa/bwt.c:169-186.
 * A block is 4 x u64 running counts followed by 8 words of 16 two-bit symbols, MSB first. */
static void occ4(const orc_index_t *ix, uint64_t k, uint64_t cnt[4])
{
  if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
  if (k >= ix->primary) --k;                                   /* '$' is not stored */
  const uint32_t *blk = ix->bwt + ((k >> 7) << 4);
  ++g_blocks;
  memcpy(cnt, blk, 32);
  const uint32_t *sym = blk + 8;
  int r = (int)(k & 127);                                      /* symbols 0..r of the block are counted */
  for (int wi = 0; wi <= (r >> 4); ++wi) {
    uint32_t W = sym[wi];
    uint32_t keep = (wi < (r >> 4)) ? 0xffffffffu : (0xffffffffu << (2 * (15 - (r & 15))));
    uint32_t lo = W & 0x55555555u, hi = (W >> 1) & 0x55555555u;
    keep &= 0x55555555u;
    cnt[0] += __builtin_popcount(~hi & ~lo & keep);
    cnt[1] += __builtin_popcount(~hi & lo & keep);
    cnt[2] += __builtin_popcount(hi & ~lo & keep);
    cnt[3] += __builtin_popcount(hi & lo & keep);
  }
}

/* bwa/bwt.c:189-220 computes both ends from one block when they share it; the result is the
 * same as two independent lookups, only the block-touch count differs. */
static void occ4_pair(const orc_index_t *ix, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4])
{
  uint64_t before = g_blocks;
  occ4(ix, k, ck); occ4(ix, l, cl);
  if (k != (uint64_t)-1 && l != (uint64_t)-1) {
    uint64_t _k = k - (k >= ix->primary), _l = l - (l >= ix->primary);
    if ((_k >> 7) == (_l >> 7)) g_blocks = before + 1;
  }
}

/* bwa/bwt.c:262-275 */
static void extend(const orc_index_t *ix, const orc_intv_t *ik, orc_intv_t ok[4], int is_back)
{
  uint64_t tk[4], tl[4];
  int f = !is_back, g = is_back;
  occ4_pair(ix, ik->x[f] - 1, ik->x[f] - 1 + ik->x[2], tk, tl);
  for (int c = 0; c < 4; ++c) {
    ok[c].x[f] = ix->L2[c] + 1 + tk[c];
    ok[c].x[2] = tl[c] - tk[c];
  }
  uint64_t acc = ik->x[g] + (ik->x[f] <= ix->primary && ik->x[f] + ik->x[2] - 1 >= ix->primary);
  for (int c = 3; c >= 0; --c) { ok[c].x[g] = acc; acc += ok[c].x[2]; }
}
void orc_bwt_extend(const orc_index_t *ix, const uint64_t ik[4], int is_back, uint64_t ok[16])
{
  orc_intv_t in, out[4]; memcpy(&in, ik, 32); memset(out, 0, sizeof(out));
  extend(ix, &in, out, is_back);
  for (int c = 0; c < 4; ++c) { out[c].info = 0; }
  memcpy(ok, out, 128);
}

static void set_intv(const orc_index_t *ix, int c, orc_intv_t *ik)
{ /* bwa/bwt.h:80 */
  ik->x[0] = ix->L2[c] + 1; ik->x[2] = ix->L2[c + 1] - ix->L2[c]; ik->x[1] = ix->L2[3 - c] + 1; ik->info = 0;
}

/* single-base Occ and LF step: bwa/bwt.c:107-129, 53-59 */
static uint64_t occ1(const orc_index_t *ix, uint64_t k, int c)
{
  if (k == ix->seq_len) return ix->L2[c + 1] - ix->L2[c];
  if (k == (uint64_t)-1) return 0;
  uint64_t cnt[4]; occ4(ix, k, cnt); return cnt[c];
}
static uint64_t inv_psi(const orc_index_t *ix, uint64_t k)
{
  if (k == ix->primary) return 0;
  uint64_t x = k - (k > ix->primary);
  int c = (ix->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)] >> ((~x & 15) << 1)) & 3;
  return ix->L2[c] + occ1(ix, k, c);     /* same block as the symbol read: one touch */
}
/* bwa/bwt.c:86-96 */
uint64_t orc_bwt_sa(const orc_index_t *ix, uint64_t k)
{
  uint64_t steps = 0, mask = (uint64_t)ix->sa_intv - 1;
  while (k & mask) { ++steps; k = inv_psi(ix, k); }
  return steps + ix->sa[k / ix->sa_intv];
}

typedef struct { orc_intv_t *a; int n, m; } ivec_t;
static void ipush(ivec_t *v, const orc_intv_t *x)
{
  if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (orc_intv_t *)realloc(v->a, v->m * sizeof(orc_intv_t)); }
  v->a[v->n++] = *x;
}
#define INFO_BEG(p) ((int)((p).info >> 32))
#define INFO_END(p) ((int)(uint32_t)(p).info)

/* bwa-flow's SMEM search, src/bwa_wrapper.cpp:325-450.  Appends to `mem` every SMEM through
 * position x whose interval is >= min_intv and whose length is >= min_seed_len; returns the
 * position the caller continues from.  max_intv is always 0 on the reference path. */
static int smem_new(const orc_index_t *ix, int len, const uint8_t *q, int x, int min_intv, int min_seed_len, ivec_t *mem)
{
  if (q[x] > 3) return x + 1;
  if (min_intv < 1) min_intv = 1;
  ivec_t fwd = {0, 0, 0}, back = {0, 0, 0};
  orc_intv_t ik, ok[4], cand;
  int i, ret;
  memset(&cand, 0, sizeof(cand));

  /* forward sweep: remember the interval each time its size changes (:346-363) */
  set_intv(ix, q[x], &ik); ik.info = (uint64_t)(x + 1);
  for (i = x + 1; i < len; ++i) {
    if (q[i] > 3) { ipush(&fwd, &ik); break; }
    int c = 3 - q[i];
    extend(ix, &ik, ok, 0);
    if (ok[c].x[2] != ik.x[2]) {
      ipush(&fwd, &ik);
      if (ok[c].x[2] < (uint64_t)min_intv) break;
    }
    ik = ok[c]; ik.info = (uint64_t)(i + 1);
  }
  if (i == len) ipush(&fwd, &ik);
  ret = INFO_END(fwd.a[fwd.n - 1]);

  /* backward phase over the forward list, shortest match first (:366-444) */
  int start = x, stop = x;
  i = 0;
  while (i < fwd.n) {
    ik = fwd.a[i]; ik.info |= (uint64_t)x << 32;
    if (back.n == 0 || stop - start >= 3) {
      /* "back-enlarge": extend this forward match leftwards base by base, keeping the trail (:371-401) */
      back.n = 0; ipush(&back, &ik);
      for (int k = x - 1; k >= 0; --k) {
        if (q[k] > 3) break;
        extend(ix, &ik, ok, 1);
        if (ok[q[k]].x[2] < (uint64_t)min_intv) break;
        ik = ok[q[k]];
        ik.info = fwd.a[i].info | (uint64_t)k << 32;
        ipush(&back, &ik);
      }
      start = INFO_END(fwd.a[i]);
      stop = (i == fwd.n - 1) ? len : INFO_END(fwd.a[i + 1]);
      if (i > 0 && INFO_BEG(ik) > INFO_BEG(cand) && INFO_END(cand) - INFO_BEG(cand) >= min_seed_len) ipush(mem, &cand);
      cand = ik;
    } else {
      /* "forward-enlarge": re-extend the saved leftward trail to the right up to this match's end (:403-430) */
      stop = INFO_END(fwd.a[i]);
      for (int k = back.n - 1; k >= 0; --k) {
        int m, reached = 0;
        ik = back.a[k];
        for (m = start + 1; m <= stop; ++m) {
          int c = 3 - q[m - 1];
          extend(ix, &ik, ok, 0);
          if (ok[c].x[2] < (uint64_t)min_intv) break;
          ik = ok[c];
          if (m == stop) { ik.info = fwd.a[i].info | (uint64_t)(uint32_t)(x - k) << 32; reached = 1; }
        }
        if (reached) {
          if ((uint64_t)(int64_t)(x - k) > (cand.info >> 32) && INFO_END(cand) - INFO_BEG(cand) >= min_seed_len) ipush(mem, &cand);
          cand = ik;
          break;
        }
      }
    }
    ++i;
    if (i < fwd.n) { /* skip forward matches that cannot reach min_seed_len (:433-441, formula kept as is) */
      int max_len = INFO_BEG(cand) + INFO_END(fwd.a[i]);
      while (max_len < min_seed_len && i < fwd.n) {
        ++i;
        if (i < fwd.n) stop = INFO_END(fwd.a[i]);
        max_len = INFO_BEG(cand) + stop;
      }
    }
    if (i >= fwd.n && INFO_END(cand) - INFO_BEG(cand) >= min_seed_len) ipush(mem, &cand);
  }
  free(fwd.a); free(back.a);
  return ret;
}

/* bwa/bwt.c:358-379 */
static int seed_strategy1(const orc_index_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, orc_intv_t *out)
{
  orc_intv_t ik, ok[4];
  memset(out, 0, sizeof(*out));
  if (q[x] > 3) return x + 1;
  set_intv(ix, q[x], &ik);
  for (int i = x + 1; i < len; ++i) {
    if (q[i] > 3) return i + 1;
    int c = 3 - q[i];
    extend(ix, &ik, ok, 0);
    if (ok[c].x[2] < (uint64_t)max_intv && i - x >= min_len) {
      *out = ok[c]; out->info = (uint64_t)x << 32 | (uint32_t)(i + 1);
      return i + 1;
    }
    ik = ok[c];
  }
  return len;
}

static int intv_cmp(const void *a, const void *b)
{
  uint64_t x = ((const orc_intv_t *)a)->info, y = ((const orc_intv_t *)b)->info;
  return (x > y) - (x < y);
}

/* src/bwa_wrapper.cpp:207-246 (three passes + sort by info; equal keys are identical elements) */
int orc_collect_intv(const orc_index_t *ix, const orc_opt_t *opt, const uint8_t *seq, int len, orc_intv_t *out, int cap)
{
  ivec_t mem = {0, 0, 0};
  if (len < opt->min_seed_len) return 0;              /* src/bwa_wrapper.cpp:150 */
  int split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
  int x = 0;
  while (x < len) x = seq[x] < 4 ? smem_new(ix, len, seq, x, 1, opt->min_seed_len, &mem) : x + 1;
  int old_n = mem.n;
  for (int k = 0; k < old_n; ++k) {
    int s = INFO_BEG(mem.a[k]), e = INFO_END(mem.a[k]);
    uint64_t occ = mem.a[k].x[2];
    if (e - s < split_len || occ > (uint64_t)opt->split_width) continue;
    smem_new(ix, len, seq, (s + e) >> 1, (int)occ + 1, opt->min_seed_len, &mem);
  }
  if (opt->max_mem_intv > 0) {
    x = 0;
    while (x < len) {
      if (seq[x] < 4) {
        orc_intv_t m;
        x = seed_strategy1(ix, len, seq, x, opt->min_seed_len, opt->max_mem_intv, &m);
        if (m.x[2] > 0) ipush(&mem, &m);
      } else ++x;
    }
  }
  qsort(mem.a, mem.n, sizeof(orc_intv_t), intv_cmp);
  int n = mem.n;
  for (int i = 0; i < n && i < cap; ++i) out[i] = mem.a[i];
  free(mem.a);
  return n;
}

/* The same search stated as a function (what bwaflow_b200/csrc/smem_sweep.cu computes; kept here so that the equivalence
 * is pinned on the CPU, tests/test_oracle_golden.py): both routes of the backward loop above ("back-enlarge" and
 * "forward-enlarge") return, for the forward match q[x, e), the leftmost begin b whose interval still has min_intv rows
 * (not crossing an N or the read start), because interval sizes only shrink as a pattern grows.  So the backward loop is:
 * for every visited forward match, extend it leftwards as far as it goes, then run the candidate / skip rules. */
static int smem_leftmost(const orc_index_t *ix, int len, const uint8_t *q, int x, int min_intv, int min_seed_len, ivec_t *mem)
{
  if (q[x] > 3) return x + 1;
  if (min_intv < 1) min_intv = 1;
  ivec_t fwd = {0, 0, 0};
  orc_intv_t ik, ok[4], cand;
  int i, ret, lb = 0;
  memset(&cand, 0, sizeof(cand));
  set_intv(ix, q[x], &ik); ik.info = (uint64_t)(x + 1);
  for (i = x + 1; i < len; ++i) {
    if (q[i] > 3) { ipush(&fwd, &ik); break; }
    int c = 3 - q[i];
    extend(ix, &ik, ok, 0);
    if (ok[c].x[2] != ik.x[2]) {
      ipush(&fwd, &ik);
      if (ok[c].x[2] < (uint64_t)min_intv) break;
    }
    ik = ok[c]; ik.info = (uint64_t)(i + 1);
  }
  if (i == len) ipush(&fwd, &ik);
  ret = INFO_END(fwd.a[fwd.n - 1]);
  for (lb = x; lb > 0 && q[lb - 1] < 4; --lb) {}              /* leftmost begin any match through x can have */
  i = 0;
  while (i < fwd.n) {
    int beg = x;
    ik = fwd.a[i];
    while (beg > lb) {                                       /* begins never move left from one forward match to the next */
      extend(ix, &ik, ok, 1);
      if (ok[q[beg - 1]].x[2] < (uint64_t)min_intv) break;
      ik = ok[q[beg - 1]]; --beg;
    }
    ik.info = (uint64_t)beg << 32 | (uint32_t)INFO_END(fwd.a[i]);
    if (i > 0 && beg > INFO_BEG(cand) && INFO_END(cand) - INFO_BEG(cand) >= min_seed_len) ipush(mem, &cand);
    cand = ik; lb = beg;
    ++i;
    if (i < fwd.n) {
      int stop = 0, max_len = INFO_BEG(cand) + INFO_END(fwd.a[i]);
      while (max_len < min_seed_len && i < fwd.n) {
        ++i;
        if (i < fwd.n) stop = INFO_END(fwd.a[i]);
        max_len = INFO_BEG(cand) + stop;
      }
    }
    if (i >= fwd.n && INFO_END(cand) - INFO_BEG(cand) >= min_seed_len) ipush(mem, &cand);
  }
  free(fwd.a);
  return ret;
}

/* passes 1 and 2 of mem_collect_intv_new through smem_leftmost, pass 3 and the sort as in orc_collect_intv */
static int seed_strategy1(const orc_index_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, orc_intv_t *out);
static int intv_cmp(const void *a, const void *b);
int orc_collect_intv_leftmost(const orc_index_t *ix, const orc_opt_t *opt, const uint8_t *seq, int len, orc_intv_t *out, int cap)
{
  ivec_t mem = {0, 0, 0};
  if (len < opt->min_seed_len) return 0;
  int split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
  int x = 0;
  while (x < len) x = seq[x] < 4 ? smem_leftmost(ix, len, seq, x, 1, opt->min_seed_len, &mem) : x + 1;
  int old_n = mem.n;
  for (int k = 0; k < old_n; ++k) {
    int s = INFO_BEG(mem.a[k]), e = INFO_END(mem.a[k]);
    uint64_t occ = mem.a[k].x[2];
    if (e - s < split_len || occ > (uint64_t)opt->split_width) continue;
    smem_leftmost(ix, len, seq, (s + e) >> 1, (int)occ + 1, opt->min_seed_len, &mem);
  }
  if (opt->max_mem_intv > 0) {
    x = 0;
    while (x < len) {
      if (seq[x] < 4) {
        orc_intv_t m;
        x = seed_strategy1(ix, len, seq, x, opt->min_seed_len, opt->max_mem_intv, &m);
        if (m.x[2] > 0) ipush(&mem, &m);
      } else ++x;
    }
  }
  qsort(mem.a, mem.n, sizeof(orc_intv_t), intv_cmp);
  int n = mem.n;
  for (int i = 0; i < n && i < cap; ++i) out[i] = mem.a[i];
  free(mem.a);
  return n;
}

/* --------------------------------------------------------- banded extension */

/* bwa/ksw.c:380-479.  out = {score, qle, tle, gtle, gscore, max_off}.  Adds executed cells to the counter. */
int orc_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, int out[6])
{
  int32_t *H = (int32_t *)calloc(qlen + 1, sizeof(int32_t));   /* H[j] = H(i-1, j-1) */
  int32_t *E = (int32_t *)calloc(qlen + 1, sizeof(int32_t));   /* E[j] = E(i, j)     */
  const int gd = o_del + e_del, gi = o_ins + e_ins;
  int best = h0, best_i = -1, best_j = -1, end_i = -1, gscore = -1, max_off = 0;
  int i, j;

  /* row -1 (:395-397) */
  H[0] = h0;
  if (qlen >= 1) H[1] = h0 > gi ? h0 - gi : 0;
  for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;

  /* band clamp in double (:399-407) */
  int smax = 0;
  for (i = 0; i < 25; ++i) if (mat[i] > smax) smax = mat[i];
  int lim = (int)((double)(qlen * smax + end_bonus - o_ins) / e_ins + 1.);
  if (lim < 1) lim = 1;
  if (w > lim) w = lim;
  lim = (int)((double)(qlen * smax + end_bonus - o_del) / e_del + 1.);
  if (lim < 1) lim = 1;
  if (w > lim) w = lim;

  int lo = 0, hi = qlen;
  for (i = 0; i < tlen; ++i) {
    const int8_t *srow = mat + target[i] * 5;
    int f = 0, rowmax = 0, rowmax_j = -1, left;
    if (lo < i - w) lo = i - w;
    if (hi > i + w + 1) hi = i + w + 1;
    if (hi > qlen) hi = qlen;
    left = 0;
    if (lo == 0) { left = h0 - (o_del + e_del * (i + 1)); if (left < 0) left = 0; }
    for (j = lo; j < hi; ++j) {
      int diag = H[j], e = E[j], M, h, t;
      H[j] = left;
      M = diag ? diag + srow[query[j]] : 0;
      h = M > e ? M : e; if (f > h) h = f;
      left = h;
      if (h >= rowmax) rowmax_j = j;
      if (h > rowmax) rowmax = h;
      t = M - gd; if (t < 0) t = 0;
      e -= e_del; if (t > e) e = t;
      E[j] = e;
      t = M - gi; if (t < 0) t = 0;
      f -= e_ins; if (t > f) f = t;
    }
    if (hi > lo) g_cells += (uint64_t)(hi - lo);
    H[hi] = left; E[hi] = 0;
    if (j == qlen) {
      if (left >= gscore) end_i = i;
      if (left > gscore) gscore = left;
    }
    if (rowmax == 0) break;
    if (rowmax > best) {
      int d = rowmax_j - i; if (d < 0) d = -d;
      best = rowmax; best_i = i; best_j = rowmax_j;
      if (d > max_off) max_off = d;
    } else if (zdrop > 0) {
      int di = i - best_i, dj = rowmax_j - best_j;
      if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
      else if (best - rowmax - (dj - di) * e_ins > zdrop) break;
    }
    for (j = lo; j < hi && H[j] == 0 && E[j] == 0; ++j) {}
    lo = j;
    for (j = hi; j >= lo && H[j] == 0 && E[j] == 0; --j) {}
    hi = j + 2 < qlen ? j + 2 : qlen;
  }
  free(H); free(E);
  out[0] = best; out[1] = best_j + 1; out[2] = best_i + 1; out[3] = end_i + 1; out[4] = gscore; out[5] = max_off;
  return best;
}

/* ---------------------------------------------------------------- reference sequence */

static inline int pac_base(const uint8_t *pac, int64_t p) { return pac[p >> 2] >> ((~p & 3) << 1) & 3; }   /* bwa/bntseq.c:224-225 */
/* base at coordinate p of the forward+reverse-complement text (bwa/bntseq.c:398-419) */
static inline int ref_base(const orc_index_t *ix, int64_t p)
{
  return p < ix->l_pac ? pac_base(ix->pac, p) : 3 - pac_base(ix->pac, (ix->l_pac << 1) - 1 - p);
}

/* ---------------------------------------------------------------- chain -> alignment regions */

static int max_gap(const orc_opt_t *o, int qlen)
{ /* bwa/bwamem.c:630-637 */
  int l_del = (int)((double)(qlen * o->a - o->o_del) / o->e_del + 1.);
  int l_ins = (int)((double)(qlen * o->a - o->o_ins) / o->e_ins + 1.);
  int l = l_del > l_ins ? l_del : l_ins;
  if (l < 1) l = 1;
  return l < o->w << 1 ? l : o->w << 1;
}

static int u64_cmp(const void *a, const void *b)
{
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return (x > y) - (x < y);
}

/* One side of a seed: up to two band tries (bwa/bwamem.c:732-743, 760-771). Returns the last score. */
static int flank(const orc_opt_t *o, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int end_bonus, int h0,
                 int prev, int *aw, int res[6])
{
  int sc = prev;
  for (int i = 0; i < 2; ++i) {
    int before = sc;
    *aw = o->w << i;
    sc = orc_ksw_extend2(qlen, q, tlen, t, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, *aw, end_bonus, o->zdrop, h0, res);
    if (sc == before || res[5] < (*aw >> 1) + (*aw >> 2)) break;
  }
  return sc;
}

/* bwa/bwamem.c:641-795 for every chain of one read, in order (src/Pipeline.cpp:515-527).
 * ctg_off/ctg_len describe the contigs (bns->anns[].offset/.len).  Returns the number of regions. */
int orc_chains2aln(const orc_index_t *ix, const orc_opt_t *o, const int64_t *ctg_off, const int32_t *ctg_len,
                   int l_query, const uint8_t *query, int n_chains, const orc_chain_t *chains, const orc_seed_t *seeds,
                   orc_alnreg_t *av, int cap)
{
  int n_av = 0;
  const int64_t L = ix->l_pac;
  for (int ci = 0; ci < n_chains; ++ci) {
    const orc_chain_t *c = &chains[ci];
    const orc_seed_t *sd = seeds + c->seed_off;
    int n = c->n_seeds, k, i;
    if (n == 0) continue;
    /* widest window any seed of the chain can reach (:651-666) */
    int64_t lo = L << 1, hi = 0;
    for (i = 0; i < n; ++i) {
      int tail = l_query - sd[i].qbeg - sd[i].len;
      int64_t b = sd[i].rbeg - (sd[i].qbeg + max_gap(o, sd[i].qbeg));
      int64_t e = sd[i].rbeg + sd[i].len + (tail + max_gap(o, tail));
      if (b < lo) lo = b;
      if (e > hi) hi = e;
    }
    if (lo < 0) lo = 0;
    if (hi > L << 1) hi = L << 1;
    if (lo < L && L < hi) { if (sd[0].rbeg < L) hi = L; else lo = L; }
    /* clip to the contig holding the first seed (bwa/bntseq.c:421-446) */
    {
      int rid = c->rid, rev = sd[0].rbeg >= L;
      int64_t fb = ctg_off[rid], fe = fb + ctg_len[rid];
      if (rev) { int64_t t = fb; fb = (L << 1) - fe; fe = (L << 1) - t; }
      if (lo < fb) lo = fb;
      if (hi > fe) hi = fe;
    }
    int64_t wlen = hi - lo;
    uint8_t *rseq = (uint8_t *)malloc(wlen > 0 ? wlen : 1);
    for (int64_t p = 0; p < wlen; ++p) rseq[p] = (uint8_t)ref_base(ix, lo + p);

    uint64_t *ord = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (i = 0; i < n; ++i) ord[i] = (uint64_t)sd[i].score << 32 | (uint32_t)i;
    qsort(ord, n, sizeof(uint64_t), u64_cmp);               /* keys are unique: any sort gives the order of :674 */

    for (k = n - 1; k >= 0; --k) {
      const orc_seed_t *s = &sd[(uint32_t)ord[k]];
      /* is the seed already explained by an earlier region of this read? (:680-696) */
      for (i = 0; i < n_av; ++i) {
        const orc_alnreg_t *p = &av[i];
        if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
        if (s->len - p->seedlen0 > .1 * l_query) continue;
        int qd = s->qbeg - p->qb; int64_t rd = s->rbeg - p->rb;
        int mg = max_gap(o, qd < rd ? qd : (int)rd);
        int w = mg < p->w ? mg : p->w;
        if (qd - rd < w && rd - qd < w) break;
        qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
        mg = max_gap(o, qd < rd ? qd : (int)rd);
        w = mg < p->w ? mg : p->w;
        if (qd - rd < w && rd - qd < w) break;
      }
      if (i < n_av) { /* contained: extend anyway only if a longer, inconsistent seed overlaps (:697-712) */
        for (i = k + 1; i < n; ++i) {
          if (ord[i] == 0) continue;
          const orc_seed_t *t = &sd[(uint32_t)ord[i]];
          if (t->len < s->len * .95) continue;
          if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
          if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
        }
        if (i == n) { ord[k] = 0; continue; }
      }
      if (n_av >= cap) { free(ord); free(rseq); return -1; }
      orc_alnreg_t *a = &av[n_av++];
      int aw[2], res[6];
      memset(a, 0, sizeof(*a));
      a->w = aw[0] = aw[1] = o->w;
      a->score = a->truesc = -1;
      a->rid = c->rid;
      if (s->qbeg) { /* left flank, both strings reversed (:724-752) */
        int ql = s->qbeg; int64_t tl = s->rbeg - lo;
        uint8_t *qs = (uint8_t *)malloc(ql), *ts = (uint8_t *)malloc(tl > 0 ? tl : 1);
        for (i = 0; i < ql; ++i) qs[i] = query[ql - 1 - i];
        for (int64_t p = 0; p < tl; ++p) ts[p] = rseq[tl - 1 - p];
        a->score = flank(o, ql, qs, (int)tl, ts, o->pen_clip5, s->len * o->a, a->score, &aw[0], res);
        if (res[4] <= 0 || res[4] <= a->score - o->pen_clip5) { a->qb = s->qbeg - res[1]; a->rb = s->rbeg - res[2]; a->truesc = a->score; }
        else { a->qb = 0; a->rb = s->rbeg - res[3]; a->truesc = res[4]; }
        free(qs); free(ts);
      } else { a->score = a->truesc = s->len * o->a; a->qb = 0; a->rb = s->rbeg; }
      if (s->qbeg + s->len != l_query) { /* right flank, seeded with the left score (:755-779) */
        int qe = s->qbeg + s->len, sc0 = a->score;
        int64_t re = s->rbeg + s->len - lo;
        a->score = flank(o, l_query - qe, query + qe, (int)(wlen - re), rseq + re, o->pen_clip3, sc0, a->score, &aw[1], res);
        if (res[4] <= 0 || res[4] <= a->score - o->pen_clip3) { a->qe = qe + res[1]; a->re = lo + re + res[2]; a->truesc += a->score - sc0; }
        else { a->qe = l_query; a->re = lo + re + res[3]; a->truesc += res[4] - sc0; }
      } else { a->qe = l_query; a->re = s->rbeg + s->len; }
      for (i = 0, a->seedcov = 0; i < n; ++i) /* (:784-788) */
        if (sd[i].qbeg >= a->qb && sd[i].qbeg + sd[i].len <= a->qe && sd[i].rbeg >= a->rb && sd[i].rbeg + sd[i].len <= a->re)
          a->seedcov += sd[i].len;
      a->w = aw[0] > aw[1] ? aw[0] : aw[1];
      a->seedlen0 = s->len;
      a->frac_rep = c->frac_rep;
    }
    free(ord); free(rseq);
  }
  return n_av;
}
