// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Never linked into, imported by or
// executed from the product path (bwaflow_b200/, include/).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the library this file builds.
//
// Thin C driver around the UNMODIFIED reference objects (bwa/*.c and
// src/bwa_wrapper.cpp compiled in place from /root/reference by
// oracle/build_ref.sh).  It only marshals flat arrays in and out of the
// reference's own functions so that Python (ctypes) can call them:
//
//   stage 2  SeqsToChains::compute       src/Pipeline.cpp:333-348
//            mem_seq2chain_wrapper       src/bwa_wrapper.cpp:118-132
//            mem_collect_intv_new        src/bwa_wrapper.cpp:207-246
//   stage 2' ChainsPipe::compute         src/Pipeline.cpp:408-438  (accelerator return path)
//   stage 4  ChainsToRegions::compute    src/Pipeline.cpp:503-544  -> mem_chain2aln bwa/bwamem.c:641-795
//   stage 5  RegionsToSam::compute       src/Pipeline.cpp:546-648
//   straight-line twin                   test/src/TestCommon.h:51-136
//
// The ksw_extend2 call sites inside mem_chain2aln are observed through a
// link-time rename (bwamem.c is compiled with -Dksw_extend2=bwaref_ksw_hook);
// the hook records the job and forwards to the real bwa/ksw.c:380 function.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>

#include "bwa_wrapper.h"   // reference header (src/), pulls bwa/*.h

extern "C" {
mem_opt_t *mem_opt_init(void);
void mem_pestat(const mem_opt_t *opt, int64_t l_pac, int n, const mem_alnreg_v *regs, mem_pestat_t pes[4]);
int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off);
bwtint_t bwt_sa(const bwt_t *bwt, bwtint_t k);
void bwt_extend(const bwt_t *bwt, const bwtintv_t *ik, bwtintv_t ok[4], int is_back);
extern char *bwa_pg;
}

// ---- symbols the reference objects expect from the application shell --------
// src/allocation_wrapper.h:8-11 declares these with C++ linkage (src/bwa_wrapper.cpp includes it).
void *calloc_wrapper(size_t n, size_t s, const char *, unsigned int, const char *) { return (calloc)(n, s); }
void *malloc_wrapper(size_t s, const char *, unsigned int, const char *) { return (malloc)(s); }
void *realloc_wrapper(void *p, size_t s, const char *, unsigned int, const char *) { return (realloc)(p, s); }
char *strdup_wrapper(const char *s, const char *, unsigned int, const char *) { return (strdup)(s); }
extern "C" {
bam1_t *bam_init1(void) { return (bam1_t *)(calloc)(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { (free)(b->data); (free)(b); } }
}

// ---- flat PODs shared with the tests (mirrored in include/bwaflow_cuda.h) ---
struct flat_seed_t { int64_t rbeg; int32_t qbeg, len, score, pad; };            // == mem_seed_t (24 B)
struct flat_chain_t { int64_t pos; int32_t n_seeds, rid; float frac_rep; uint32_t flags; int64_t seed_off; };
struct flat_job_t {                       // one observed ksw_extend2 call
  int32_t qlen, tlen, w, h0, end_bonus, zdrop;
  int32_t score, qle, tle, gtle, gscore, max_off;
  int64_t q_off, t_off;                   // into the byte pools
};

// ---- ksw_extend2 observation hook ------------------------------------------
struct hook_state_t {
  bool on = false;
  std::vector<flat_job_t> jobs;
  std::vector<uint8_t> qpool, tpool;
};
static thread_local hook_state_t g_hook;

extern "C" int bwaref_ksw_hook(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                               int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                               int *qle, int *tle, int *gtle, int *gscore, int *max_off)
{
  int _qle, _tle, _gtle, _gs, _mo;
  int sc = ksw_extend2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                       &_qle, &_tle, &_gtle, &_gs, &_mo);
  if (g_hook.on) {
    flat_job_t j;
    j.qlen = qlen; j.tlen = tlen; j.w = w; j.h0 = h0; j.end_bonus = end_bonus; j.zdrop = zdrop;
    j.score = sc; j.qle = _qle; j.tle = _tle; j.gtle = _gtle; j.gscore = _gs; j.max_off = _mo;
    j.q_off = (int64_t)g_hook.qpool.size(); j.t_off = (int64_t)g_hook.tpool.size();
    g_hook.qpool.insert(g_hook.qpool.end(), query, query + qlen);
    g_hook.tpool.insert(g_hook.tpool.end(), target, target + tlen);
    g_hook.jobs.push_back(j);
  }
  if (qle) *qle = _qle; if (tle) *tle = _tle; if (gtle) *gtle = _gtle;
  if (gscore) *gscore = _gs; if (max_off) *max_off = _mo;
  return sc;
}

// ---- ksw_align2 observation hook (mem_matesw's calls; bwamem_pair.c is compiled with -Dksw_align2=bwaref_sw_hook) -------------
struct flat_swjob_t {                     // one observed ksw_align2 call and its kswr_t
  int32_t qlen, tlen, xtra, pad;
  int64_t q_off, t_off;
  int32_t score, te, qe, score2, te2, tb, qb, pad2;
};
struct swhook_state_t { bool on = false; std::vector<flat_swjob_t> jobs; std::vector<uint8_t> qpool, tpool; };
static thread_local swhook_state_t g_swhook;
extern "C" kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat, int o_del, int e_del, int o_ins,
                             int e_ins, int xtra, kswq_t **qry);
extern "C" kswr_t bwaref_sw_hook(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat, int o_del, int e_del,
                                 int o_ins, int e_ins, int xtra, kswq_t **qry)
{
  flat_swjob_t j;
  if (g_swhook.on) {
    j.qlen = qlen; j.tlen = tlen; j.xtra = xtra; j.pad = j.pad2 = 0;
    j.q_off = (int64_t)g_swhook.qpool.size(); j.t_off = (int64_t)g_swhook.tpool.size();
    g_swhook.qpool.insert(g_swhook.qpool.end(), query, query + qlen);
    g_swhook.tpool.insert(g_swhook.tpool.end(), target, target + tlen);
  }
  kswr_t r = ksw_align2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra, qry);
  if (g_swhook.on) {
    j.score = r.score; j.te = r.te; j.qe = r.qe; j.score2 = r.score2; j.te2 = r.te2; j.tb = r.tb; j.qb = r.qb;
    g_swhook.jobs.push_back(j);
  }
  return r;
}

// ---- handle ------------------------------------------------------------------
struct ref_t {
  bwaidx_t *idx;
  mem_opt_t *opt;
  ktp_aux_t aux;
};

struct batch_t {
  ref_t *ref;
  int n;
  uint64_t start_idx;
  bseq1_t *seqs;
  mem_chain_v *chains;     // stage-2 output (owned until consumed by stage 4)
  mem_alnreg_v *alnreg;    // stage-4 output
  bool nt4;
  mem_pestat_t pes[4]; bool have_pes;
  // flat export scratch
  std::vector<bwtintv_t> intv; std::vector<int64_t> intv_off;
  std::vector<flat_chain_t> fchains; std::vector<flat_seed_t> fseeds; std::vector<int64_t> chain_off;
  std::vector<mem_alnreg_t> fregs; std::vector<int64_t> reg_off;
  std::string sam; std::vector<int64_t> sam_off;
};

static void to_nt4(bseq1_t *s)
{ // src/bwa_wrapper.cpp:124-125
  for (int i = 0; i < s->l_seq; ++i)
    s->seq[i] = s->seq[i] < 4 ? s->seq[i] : nst_nt4_table[(int)s->seq[i]];
}

static void free_chain_v(mem_chain_v *c)
{
  for (size_t j = 0; j < c->n; ++j) free(c->a[j].seeds);
  free(c->a); c->a = 0; c->n = c->m = 0;
}

extern "C" {

int bwaref_index_build(const char *fasta, const char *prefix)
{
  int v = bwa_verbose; bwa_verbose = 1;
  int r = bwa_idx_build(fasta, prefix, BWTALGO_AUTO, 10000000);
  bwa_verbose = v;
  return r;
}

void *bwaref_open(const char *prefix, int paired)
{
  bwa_verbose = 1;
  bwaidx_t *idx = bwa_idx_load(prefix, BWA_IDX_ALL);
  if (!idx) return 0;
  ref_t *r = new ref_t();
  r->idx = idx;
  r->opt = mem_opt_init();
  if (paired) r->opt->flag |= MEM_F_PE;
  r->opt->flag |= MEM_F_NO_MULTI;                 // -M is on by default in bwa-flow: src/config.cpp:16, src/main.cpp:161-163
  memset(&r->aux, 0, sizeof(r->aux));
  r->aux.opt = r->opt; r->aux.idx = idx;
  if (!bwa_pg) bwa_pg = strdup("@PG\tID:bwa\tPN:bwa\tVN:0.7.17-r1194-dirty\tCL:bwaref");
  return r;
}

void bwaref_close(void *h)
{
  ref_t *r = (ref_t *)h;
  if (!r) return;
  bwa_idx_destroy(r->idx); free(r->opt); delete r;
}

// name -> int option access (enough for the tests: scoring, band, zdrop, seeding knobs)
int bwaref_opt(void *h, const char *key, int set, int value)
{
  mem_opt_t *o = ((ref_t *)h)->opt;
  int *p = 0;
#define K(name) if (!strcmp(key, #name)) p = &o->name;
  K(a) K(b) K(o_del) K(e_del) K(o_ins) K(e_ins) K(pen_clip5) K(pen_clip3) K(w) K(zdrop) K(T) K(flag)
  K(min_seed_len) K(split_width) K(max_occ) K(max_chain_gap) K(min_chain_weight) K(max_chain_extend) K(pen_unpaired)
#undef K
  if (!strcmp(key, "max_mem_intv")) { if (set) o->max_mem_intv = (uint64_t)value; return (int)o->max_mem_intv; }
  if (!p) return -2147483647;
  if (set) { *p = value; if (!strcmp(key, "a") || !strcmp(key, "b")) bwa_fill_scmat(o->a, o->b, o->mat); }
  return *p;
}
float bwaref_opt_split_factor(void *h) { return ((ref_t *)h)->opt->split_factor; }
void bwaref_opt_mat(void *h, int8_t out[25]) { memcpy(out, ((ref_t *)h)->opt->mat, 25); }

// raw index views (pointers stay owned by the reference library)
struct ref_index_view_t {
  uint64_t primary, L2[5], seq_len, bwt_size; const uint32_t *bwt;
  uint64_t n_sa; int64_t sa_intv; const uint64_t *sa;
  int64_t l_pac; const uint8_t *pac; int32_t n_seqs, pad;
};
void bwaref_index_view(void *h, ref_index_view_t *v)
{
  ref_t *r = (ref_t *)h; const bwt_t *b = r->idx->bwt;
  v->primary = b->primary; memcpy(v->L2, b->L2, sizeof(v->L2)); v->seq_len = b->seq_len; v->bwt_size = b->bwt_size; v->bwt = b->bwt;
  v->n_sa = b->n_sa; v->sa_intv = b->sa_intv; v->sa = b->sa;
  v->l_pac = r->idx->bns->l_pac; v->pac = r->idx->pac; v->n_seqs = r->idx->bns->n_seqs; v->pad = 0;
}
void bwaref_contigs(void *h, int64_t *offset, int32_t *len, int32_t *is_alt)
{
  const bntseq_t *bns = ((ref_t *)h)->idx->bns;
  for (int i = 0; i < bns->n_seqs; ++i) { offset[i] = bns->anns[i].offset; len[i] = bns->anns[i].len; is_alt[i] = bns->anns[i].is_alt; }
}
int bwaref_contig_name(void *h, int i, char *buf, int cap)
{
  const bntseq_t *bns = ((ref_t *)h)->idx->bns;
  snprintf(buf, cap, "%s", bns->anns[i].name); return (int)strlen(bns->anns[i].name);
}
int bwaref_sam_header(void *h, char *buf, int cap)
{ // src/bwa_wrapper.cpp:25-35 with the default read group of src/config.cpp:10 left out (rg_line = 0)
  kstring_t s = {0, 0, 0};                       // bwa_format_sam_hdr resets l and s but not m
  bwa_format_sam_hdr(((ref_t *)h)->idx->bns, 0, &s);
  int l = (int)s.l; if (buf && cap > l) memcpy(buf, s.s, l + 1);
  free(s.s); return l;
}

// ---- single-call primitives --------------------------------------------------
uint64_t bwaref_bwt_sa(void *h, uint64_t k) { return bwt_sa(((ref_t *)h)->idx->bwt, k); }

void bwaref_bwt_extend(void *h, const uint64_t ik[4], int is_back, uint64_t ok[16])
{
  bwtintv_t in, out[4]; memcpy(&in, ik, 32);
  bwt_extend(((ref_t *)h)->idx->bwt, &in, out, is_back);
  memcpy(ok, out, 128);
}

int bwaref_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                       int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, int out[6])
{
  int qle, tle, gtle, gscore, max_off;
  int sc = ksw_extend2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                       &qle, &tle, &gtle, &gscore, &max_off);
  out[0] = sc; out[1] = qle; out[2] = tle; out[3] = gtle; out[4] = gscore; out[5] = max_off;
  return sc;
}

// nt4 read -> sorted interval list, exactly what stage 2 feeds to chaining (src/bwa_wrapper.cpp:207-246)
int bwaref_collect_intv(void *h, const uint8_t *nt4, int len, bwtintv_t *out, int cap)
{
  ref_t *r = (ref_t *)h;
  smem_aux_t *a = smem_aux_init();
  int n = 0;
  if (len >= r->opt->min_seed_len) {            // mem_chain_new's early return, src/bwa_wrapper.cpp:150
    mem_collect_intv_new(r->opt, r->idx->bwt, len, nt4, a);
    n = (int)a->mem.n;
    for (int i = 0; i < n && i < cap; ++i) out[i] = a->mem.a[i];
  }
  smem_aux_destroy(a);
  return n;
}

// ---- ksw hook control --------------------------------------------------------
void bwaref_hook_begin(void) { g_hook.on = true; g_hook.jobs.clear(); g_hook.qpool.clear(); g_hook.tpool.clear(); }
void bwaref_hook_end(void) { g_hook.on = false; }
int64_t bwaref_hook_count(int64_t *qbytes, int64_t *tbytes)
{ *qbytes = (int64_t)g_hook.qpool.size(); *tbytes = (int64_t)g_hook.tpool.size(); return (int64_t)g_hook.jobs.size(); }
void bwaref_hook_fetch(flat_job_t *jobs, uint8_t *q, uint8_t *t)
{
  if (!g_hook.jobs.empty()) memcpy(jobs, g_hook.jobs.data(), g_hook.jobs.size() * sizeof(flat_job_t));
  if (!g_hook.qpool.empty()) memcpy(q, g_hook.qpool.data(), g_hook.qpool.size());
  if (!g_hook.tpool.empty()) memcpy(t, g_hook.tpool.data(), g_hook.tpool.size());
}

// ---- batch object: mirrors the records flowing between the kflow stages ----
void *bwaref_batch_new(void *h, int n, const char *names, const int64_t *name_off,
                       const char *bases, const char *quals, const int64_t *seq_off, uint64_t start_idx)
{
  batch_t *b = new batch_t();
  b->ref = (ref_t *)h; b->n = n; b->start_idx = start_idx; b->chains = 0; b->alnreg = 0; b->nt4 = false;
  b->seqs = (bseq1_t *)calloc(n, sizeof(bseq1_t));
  for (int i = 0; i < n; ++i) {
    bseq1_t *s = &b->seqs[i];
    int64_t l = seq_off[i + 1] - seq_off[i], ln = name_off[i + 1] - name_off[i];
    s->l_seq = (int)l;
    s->name = (char *)malloc(ln + 1); memcpy(s->name, names + name_off[i], ln); s->name[ln] = 0;
    s->seq = (char *)malloc(l + 1); memcpy(s->seq, bases + seq_off[i], l); s->seq[l] = 0;
    if (quals) { s->qual = (char *)malloc(l + 1); memcpy(s->qual, quals + seq_off[i], l); s->qual[l] = 0; }
  }
  return b;
}

static void batch_drop_chains(batch_t *b)
{
  if (!b->chains) return;
  for (int i = 0; i < b->n; ++i) free_chain_v(&b->chains[i]);
  free(b->chains); b->chains = 0;
}
static void batch_drop_regs(batch_t *b)
{
  if (!b->alnreg) return;
  for (int i = 0; i < b->n; ++i) free(b->alnreg[i].a);
  free(b->alnreg); b->alnreg = 0;
}

void bwaref_batch_free(void *hb)
{
  batch_t *b = (batch_t *)hb;
  batch_drop_chains(b); batch_drop_regs(b);
  for (int i = 0; i < b->n; ++i) { free(b->seqs[i].name); free(b->seqs[i].comment); free(b->seqs[i].seq); free(b->seqs[i].qual); free(b->seqs[i].sam); }
  free(b->seqs);
  delete b;
}

static void batch_nt4(batch_t *b)
{
  if (b->nt4) return;
  for (int i = 0; i < b->n; ++i) to_nt4(&b->seqs[i]);
  b->nt4 = true;
}

// nt4 bases of the whole batch, concatenated (what a device stage uploads)
void bwaref_batch_nt4(void *hb, uint8_t *out)
{
  batch_t *b = (batch_t *)hb; batch_nt4(b);
  int64_t o = 0;
  for (int i = 0; i < b->n; ++i) { memcpy(out + o, b->seqs[i].seq, b->seqs[i].l_seq); o += b->seqs[i].l_seq; }
}

// K1 contract: per-read sorted bwtintv_t lists
int64_t bwaref_batch_intervals(void *hb, const bwtintv_t **intv, const int64_t **off)
{
  batch_t *b = (batch_t *)hb; batch_nt4(b);
  b->intv.clear(); b->intv_off.assign(1, 0);
  smem_aux_t *a = smem_aux_init();
  for (int i = 0; i < b->n; ++i) {
    bseq1_t *s = &b->seqs[i];
    a->mem.n = 0;
    if (s->l_seq >= b->ref->opt->min_seed_len)
      mem_collect_intv_new(b->ref->opt, b->ref->idx->bwt, s->l_seq, (uint8_t *)s->seq, a);
    b->intv.insert(b->intv.end(), a->mem.a, a->mem.a + a->mem.n);
    b->intv_off.push_back((int64_t)b->intv.size());
  }
  smem_aux_destroy(a);
  *intv = b->intv.data(); *off = b->intv_off.data();
  return (int64_t)b->intv.size();
}

// stage 2 on the CPU path: SeqsToChains::compute (src/Pipeline.cpp:333-348)
void bwaref_batch_seq2chains(void *hb)
{
  batch_t *b = (batch_t *)hb;
  batch_drop_chains(b);
  b->chains = (mem_chain_v *)malloc(b->n * sizeof(mem_chain_v));
  for (int i = 0; i < b->n; ++i) b->chains[i] = mem_seq2chain_wrapper(&b->ref->aux, &b->seqs[i]);
  b->nt4 = true;
}

// stage 2' accelerator return path: ChainsPipe::compute (src/Pipeline.cpp:408-438)
void bwaref_batch_chainspipe(void *hb, const bwtintv_t *intv, const int64_t *off)
{
  batch_t *b = (batch_t *)hb; batch_nt4(b);
  batch_drop_chains(b);
  const mem_opt_t *opt = b->ref->opt; const bwaidx_t *idx = b->ref->idx;
  b->chains = (mem_chain_v *)malloc(b->n * sizeof(mem_chain_v));
  for (int i = 0; i < b->n; ++i) {
    bseq1_t *s = &b->seqs[i];
    smem_aux_t *a = smem_aux_init();
    size_t n = (size_t)(off[i + 1] - off[i]);
    a->mem.a = (bwtintv_t *)malloc((n ? n : 1) * sizeof(bwtintv_t)); a->mem.n = a->mem.m = n;
    memcpy(a->mem.a, intv + off[i], n * sizeof(bwtintv_t));
    b->chains[i] = mem_chain_postprocess(opt, idx->bwt, idx->bns, s->l_seq, (uint8_t *)s->seq, a);
    smem_aux_destroy(a);
    b->chains[i].n = mem_chain_flt(opt, b->chains[i].n, b->chains[i].a);
    mem_flt_chained_seeds(opt, idx->bns, idx->pac, s->l_seq, (uint8_t *)s->seq, b->chains[i].n, b->chains[i].a);
  }
}

int64_t bwaref_batch_chains(void *hb, const flat_chain_t **chains, const flat_seed_t **seeds, const int64_t **off, int64_t *n_seeds)
{
  batch_t *b = (batch_t *)hb;
  b->fchains.clear(); b->fseeds.clear(); b->chain_off.assign(1, 0);
  for (int i = 0; i < b->n; ++i) {
    for (size_t j = 0; b->chains && j < b->chains[i].n; ++j) {
      const mem_chain_t *c = &b->chains[i].a[j];
      flat_chain_t f; f.pos = c->pos; f.n_seeds = c->n; f.rid = c->rid; f.frac_rep = c->frac_rep;
      f.flags = (uint32_t)c->w | (uint32_t)c->kept << 29 | (uint32_t)c->is_alt << 31; f.seed_off = (int64_t)b->fseeds.size();
      for (int k = 0; k < c->n; ++k) {
        flat_seed_t s; s.rbeg = c->seeds[k].rbeg; s.qbeg = c->seeds[k].qbeg; s.len = c->seeds[k].len; s.score = c->seeds[k].score; s.pad = 0;
        b->fseeds.push_back(s);
      }
      b->fchains.push_back(f);
    }
    b->chain_off.push_back((int64_t)b->fchains.size());
  }
  *chains = b->fchains.data(); *seeds = b->fseeds.data(); *off = b->chain_off.data(); *n_seeds = (int64_t)b->fseeds.size();
  return (int64_t)b->fchains.size();
}

// stage 4 on the CPU path: ChainsToRegions::compute (src/Pipeline.cpp:503-544).  Chains are kept (not freed).
void bwaref_batch_chains2regions(void *hb)
{
  batch_t *b = (batch_t *)hb;
  batch_drop_regs(b);
  const mem_opt_t *opt = b->ref->opt; const bwaidx_t *idx = b->ref->idx;
  b->alnreg = (mem_alnreg_v *)malloc(b->n * sizeof(mem_alnreg_v));
  for (int i = 0; i < b->n; ++i) {
    kv_init(b->alnreg[i]);
    for (size_t j = 0; j < b->chains[i].n; ++j)
      mem_chain2aln(opt, idx->bns, idx->pac, b->seqs[i].l_seq, (uint8_t *)b->seqs[i].seq, &b->chains[i].a[j], &b->alnreg[i]);
  }
}

// inject stage-4 results computed elsewhere (the device stage under test)
void bwaref_batch_set_regions(void *hb, const mem_alnreg_t *regs, const int64_t *off)
{
  batch_t *b = (batch_t *)hb; batch_nt4(b);
  batch_drop_regs(b);
  b->alnreg = (mem_alnreg_v *)malloc(b->n * sizeof(mem_alnreg_v));
  for (int i = 0; i < b->n; ++i) {
    size_t n = (size_t)(off[i + 1] - off[i]);
    b->alnreg[i].n = b->alnreg[i].m = n;
    b->alnreg[i].a = (mem_alnreg_t *)malloc((n ? n : 1) * sizeof(mem_alnreg_t));
    memcpy(b->alnreg[i].a, regs + off[i], n * sizeof(mem_alnreg_t));
  }
}

int64_t bwaref_batch_regions(void *hb, const mem_alnreg_t **regs, const int64_t **off)
{
  batch_t *b = (batch_t *)hb;
  b->fregs.clear(); b->reg_off.assign(1, 0);
  for (int i = 0; i < b->n; ++i) {
    if (b->alnreg) b->fregs.insert(b->fregs.end(), b->alnreg[i].a, b->alnreg[i].a + b->alnreg[i].n);
    b->reg_off.push_back((int64_t)b->fregs.size());
  }
  *regs = b->fregs.data(); *off = b->reg_off.data();
  return (int64_t)b->fregs.size();
}

// stage 5: RegionsToSam::compute (src/Pipeline.cpp:546-648), text path (USE_HTSLIB undefined)
int64_t bwaref_batch_regions2sam(void *hb, const char **sam, const int64_t **off)
{
  batch_t *b = (batch_t *)hb;
  const mem_opt_t *opt = b->ref->opt; const bwaidx_t *idx = b->ref->idx;
  batch_drop_chains(b);
  for (int i = 0; i < b->n; ++i) {
    b->alnreg[i].n = mem_sort_dedup_patch(opt, idx->bns, idx->pac, (uint8_t *)b->seqs[i].seq, (int)b->alnreg[i].n, b->alnreg[i].a);
    for (size_t j = 0; j < b->alnreg[i].n; ++j) {
      mem_alnreg_t *p = &b->alnreg[i].a[j];
      if (p->rid >= 0 && idx->bns->anns[p->rid].is_alt) p->is_alt = 1;
    }
  }
  if (opt->flag & MEM_F_PE) {
    mem_pestat_t pes[4];
    mem_pestat(opt, idx->bns->l_pac, b->n, b->alnreg, pes);
    for (int i = 0; i < b->n / 2; ++i)
      mem_sam_pe_hts(opt, idx->bns, idx->pac, pes, (b->start_idx >> 1) + i, &b->seqs[i << 1], &b->alnreg[i << 1], 0);
  } else {
    for (int i = 0; i < b->n; ++i) {
      mem_mark_primary_se(opt, (int)b->alnreg[i].n, b->alnreg[i].a, b->start_idx + i);
      mem_reg2sam_hts(opt, idx->bns, idx->pac, &b->seqs[i], &b->alnreg[i], 0, 0, 0);
    }
  }
  batch_drop_regs(b);
  b->sam.clear(); b->sam_off.assign(1, 0);
  for (int i = 0; i < b->n; ++i) {
    if (b->seqs[i].sam) b->sam += b->seqs[i].sam;
    b->sam_off.push_back((int64_t)b->sam.size());
  }
  *sam = b->sam.data(); *off = b->sam_off.data();
  return (int64_t)b->sam.size();
}

// ---- pieces of stage 5 taken one at a time (RegionsToSam::compute, src/Pipeline.cpp:560-604) -------------------------------------
// (1) mem_sort_dedup_patch + ALT marking of every read (:560-575), then mem_pestat over the whole batch (:578-579).
//     pes_out: 4 x {low, high, failed, avg, std} as doubles.
void bwaref_batch_dedup_pestat(void *hb, double *pes_out)
{
  batch_t *b = (batch_t *)hb;
  const mem_opt_t *opt = b->ref->opt; const bwaidx_t *idx = b->ref->idx;
  batch_nt4(b);
  for (int i = 0; i < b->n; ++i) {
    b->alnreg[i].n = mem_sort_dedup_patch(opt, idx->bns, idx->pac, (uint8_t *)b->seqs[i].seq, (int)b->alnreg[i].n, b->alnreg[i].a);
    for (size_t j = 0; j < b->alnreg[i].n; ++j) {
      mem_alnreg_t *p = &b->alnreg[i].a[j];
      if (p->rid >= 0 && idx->bns->anns[p->rid].is_alt) p->is_alt = 1;
    }
  }
  mem_pestat(opt, idx->bns->l_pac, b->n, b->alnreg, b->pes);
  b->have_pes = true;
  for (int d = 0; d < 4; ++d) {
    pes_out[d * 5 + 0] = b->pes[d].low; pes_out[d * 5 + 1] = b->pes[d].high; pes_out[d * 5 + 2] = b->pes[d].failed;
    pes_out[d * 5 + 3] = b->pes[d].avg; pes_out[d * 5 + 4] = b->pes[d].std;
  }
}

// (2) the mate-rescue block of mem_sam_pe (src/bwa_wrapper.cpp:657-668) for every pair, calling the reference's own mem_matesw.
//     Returns the sum of its return values; the regions afterwards are read with bwaref_batch_regions.
extern "C" int mem_matesw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], const mem_alnreg_t *a,
                          int l_ms, const uint8_t *ms, mem_alnreg_v *ma);
int64_t bwaref_batch_matesw(void *hb)
{
  batch_t *b = (batch_t *)hb;
  const mem_opt_t *opt = b->ref->opt; const bwaidx_t *idx = b->ref->idx;
  int64_t n = 0;
  if (!b->have_pes) return -1;
  for (int pr = 0; pr < b->n / 2; ++pr) {
    bseq1_t *s = &b->seqs[pr << 1]; mem_alnreg_v *a = &b->alnreg[pr << 1];
    if (opt->flag & MEM_F_NO_RESCUE) continue;
    mem_alnreg_v bb[2];
    kv_init(bb[0]); kv_init(bb[1]);
    for (int i = 0; i < 2; ++i)
      for (size_t j = 0; j < a[i].n; ++j)
        if (a[i].a[j].score >= a[i].a[0].score - opt->pen_unpaired) kv_push(mem_alnreg_t, bb[i], a[i].a[j]);
    for (int i = 0; i < 2; ++i)
      for (size_t j = 0; j < bb[i].n && (int)j < opt->max_matesw; ++j)
        n += mem_matesw(opt, idx->bns, idx->pac, b->pes, &bb[i].a[j], s[!i].l_seq, (uint8_t *)s[!i].seq, &a[!i]);
    free(bb[0].a); free(bb[1].a);
  }
  return n;
}

// ksw_align2 as the library exports it (query / target are restored by the function itself)
void bwaref_ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                       int xtra, int out[7])
{
  kswr_t r = ksw_align2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, xtra, 0);
  out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = r.tb; out[6] = r.qb;
}

void bwaref_swhook_begin(void) { g_swhook.on = true; g_swhook.jobs.clear(); g_swhook.qpool.clear(); g_swhook.tpool.clear(); }
void bwaref_swhook_end(void) { g_swhook.on = false; }
int64_t bwaref_swhook_count(int64_t *qbytes, int64_t *tbytes)
{ *qbytes = (int64_t)g_swhook.qpool.size(); *tbytes = (int64_t)g_swhook.tpool.size(); return (int64_t)g_swhook.jobs.size(); }
void bwaref_swhook_fetch(flat_swjob_t *jobs, uint8_t *q, uint8_t *t)
{
  if (!g_swhook.jobs.empty()) memcpy(jobs, g_swhook.jobs.data(), g_swhook.jobs.size() * sizeof(flat_swjob_t));
  if (!g_swhook.qpool.empty()) memcpy(q, g_swhook.qpool.data(), g_swhook.qpool.size());
  if (!g_swhook.tpool.empty()) memcpy(t, g_swhook.tpool.data(), g_swhook.tpool.size());
}

// mem_reg2aln (bwa/bwamem.c:1104-1174) on one region: the mem_aln_t fields, the CIGAR and the MD string stored behind it.
struct flat_aln_t { int64_t pos; int32_t rid, flag, is_rev, is_alt, mapq, NM, n_cigar, score, sub, alt_sc, pad; };
extern "C" mem_aln_t mem_reg2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const char *query, const mem_alnreg_t *ar);
int bwaref_reg2aln(void *h, int l_query, const char *query, const mem_alnreg_t *ar, flat_aln_t *out, uint32_t *cigar, int cigar_cap, char *md, int md_cap)
{
  ref_t *r = (ref_t *)h;
  mem_aln_t a = mem_reg2aln(r->opt, r->idx->bns, r->idx->pac, l_query, query, ar);
  out->pos = a.pos; out->rid = a.rid; out->flag = a.flag; out->is_rev = a.is_rev; out->is_alt = a.is_alt; out->mapq = a.mapq; out->NM = a.NM;
  out->n_cigar = a.n_cigar; out->score = a.score; out->sub = a.sub; out->alt_sc = a.alt_sc; out->pad = 0;
  int ok = a.n_cigar <= cigar_cap;
  if (a.cigar) {
    if (ok) memcpy(cigar, a.cigar, a.n_cigar * 4);
    const char *m = (const char *)(a.cigar + a.n_cigar);
    if ((int)strlen(m) + 1 <= md_cap) strcpy(md, m); else ok = 0;
    free(a.cigar);
  } else md[0] = 0;
  return ok ? 0 : -1;
}

} // extern "C"

// ---- ingest: the reference's own FASTA/FASTQ parser and batching, for the parity test of the host-side FastqSource -----------------
// bseq_read (bwa/bwa.c:53-85) is the stock twin of getKseqBatch + KseqsToBseqs (src/Pipeline.cpp:98-202): kseq_read's parsing rules,
// whole pairs until the record holds >= chunk_size bases, trim_readno, an empty comment is no comment.  Dump: one line
// "#record <start_idx> <batch_num>" per record, then "name\tcomment\tseq\tqual" per read ("" for no comment, "*" for no quality).
// (kseq_t / kseq_init / kseq_destroy: KSEQ_DECLARE(gzFile) in src/bwa_wrapper.h:29)
extern "C" int64_t bwaref_fastq_dump(const char *p1, const char *p2, int chunk_size, char **out)
{
  gzFile f1 = gzopen(p1, "r"), f2 = p2 ? gzopen(p2, "r") : 0;
  if (!f1 || (p2 && !f2)) return -1;
  kseq_t *ks = kseq_init(f1), *ks2 = f2 ? kseq_init(f2) : 0;
  std::string o;
  int64_t start = 0;
  for (;;) {
    int n = 0;
    bseq1_t *seqs = bseq_read(chunk_size, &n, ks, ks2);
    if (n == 0) { free(seqs); break; }
    o += "#record " + std::to_string(start) + " " + std::to_string(n) + "\n";
    for (int i = 0; i < n; ++i) {
      o += seqs[i].name; o += '\t';
      if (seqs[i].comment) o += seqs[i].comment;
      o += '\t'; o += seqs[i].seq; o += '\t';
      o += seqs[i].qual ? seqs[i].qual : "*";
      o += '\n';
      free(seqs[i].name); free(seqs[i].comment); free(seqs[i].seq); free(seqs[i].qual);
    }
    free(seqs);
    start += n;
  }
  kseq_destroy(ks); gzclose(f1);
  if (ks2) { kseq_destroy(ks2); gzclose(f2); }
  *out = (char *)malloc(o.size() + 1);
  memcpy(*out, o.data(), o.size() + 1);
  return (int64_t)o.size();
}
extern "C" void bwaref_free(void *p) { free(p); }
