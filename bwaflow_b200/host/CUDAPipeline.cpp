// src/cuda/CUDAPipeline.cpp -- worker loops shaped like SeqsToChainsFPGA::compute / ChainsToRegionsFPGA::compute
// (src/fpga/FPGAPipeline.cpp:582-800, 367-555) without their CPU fallbacks: errors are fatal (std::runtime_error, src/util.h:15-31).
#include <stdlib.h>
#include <string.h>
#include "CUDAPipeline.h"

static void check(int rc) { if (rc) throw std::runtime_error(bwaflow_last_error()); }

CUDABatchPool::CUDABatchPool(bwaflow_ctx_t *ctx, int n_batches, int64_t max_reads, int64_t max_bases, const bwaflow_chain_opt_t *copt,
                             const int32_t *ctg_is_alt)
{
  for (int i = 0; i < n_batches; ++i) {
    bwaflow_hot_batch_t *b = nullptr;
    check(bwaflow_hot_batch_create(ctx, max_reads, max_bases, copt, ctg_is_alt, &b));
    all_.push_back(b);
    free_.push(b);
  }
}
CUDABatchPool::~CUDABatchPool() { for (auto *b : all_) bwaflow_hot_batch_destroy(b); }
bwaflow_hot_batch_t *CUDABatchPool::take()
{
  bwaflow_hot_batch_t *b = nullptr;
  while (!free_.async_pop(b)) std::this_thread::sleep_for(std::chrono::microseconds(10));
  return b;
}
void CUDABatchPool::give(bwaflow_hot_batch_t *b) { free_.push(b); }

CUDAEnv::CUDAEnv(const bwaflow_index_t *idx, const bwaflow_opt_t *opt, const bwaflow_chain_opt_t *copt, const int32_t *ctg_is_alt, int n_devices,
                 int batches_per_device, int64_t max_reads, int64_t max_bases)
    : max_reads_(max_reads), max_bases_(max_bases)
{
  if (copt) copt_ = *copt; else bwaflow_chain_opt_default(&copt_);
  const int visible = bwaflow_device_count();
  if (visible <= 0) throw std::runtime_error("CUDAEnv: no CUDA device (there is no CPU fallback on this path)");
  if (n_devices <= 0 || n_devices > visible) n_devices = visible;
  for (int d = 0; d < n_devices; ++d) {                        // the index is replicated: every GPU gets its own copy in HBM
    PE pe; pe.device = d; pe.ctx = nullptr; pe.pool = nullptr;
    check(bwaflow_ctx_create(d, idx, opt, &pe.ctx));
    if (batches_per_device > 0) {
      pe.pool = new CUDABatchPool(pe.ctx, batches_per_device, max_reads, max_bases, &copt_, ctg_is_alt);
      for (auto *b : pe.pool->all()) owner_[b] = pe.pool;
    }
    pes_.push_back(pe);
  }
}
CUDAEnv::~CUDAEnv()
{
  for (auto &pe : pes_) { delete pe.pool; bwaflow_ctx_destroy(pe.ctx); }
}

void SeedPositionTable::put(uint64_t start_idx, Entry &&e)
{
  std::lock_guard<std::mutex> lk(m_);
  tab_[start_idx] = std::move(e);
}
SeedPositionTable::Entry SeedPositionTable::take(uint64_t start_idx)
{
  std::lock_guard<std::mutex> lk(m_);
  auto it = tab_.find(start_idx);
  if (it == tab_.end()) throw std::runtime_error("ChainsPipe: no seed positions for this record");
  Entry e = std::move(it->second);
  tab_.erase(it);
  return e;
}

// nst_nt4_table (bwa/bntseq.c:47-64): A,C,G,T (either case) -> 0..3, everything else 4; codes 0..3 stay (stage 2 may see a record twice)
struct Nt4Table {
  uint8_t t[256];
  Nt4Table() { for (int c = 0; c < 256; ++c) t[c] = c < 4 ? (uint8_t)c : 4; t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; }
};
static const Nt4Table nt4_table;
static inline uint8_t nt4(char c) { return nt4_table.t[(uint8_t)c]; }

void SeqsToChainsCUDA::compute(int wid)
{
  CUDABatchPool *pool_ = this->pool_;
  bwaflow_ctx_t *ctx_ = this->ctx_;
  if (env_) { CUDAEnv::PE &pe = env_->getPE(wid); pool_ = pe.pool; ctx_ = pe.ctx; }    // this worker's GPU
  std::vector<uint8_t> bases;
  std::vector<int64_t> off;
  bwaflow_smem_batch_t *smem = nullptr;                        // interval mode: this worker's SMemTask
  if (!pool_) check(bwaflow_smem_batch_create(ctx_, max_reads_, max_bases_, &smem));
  for (;;) {
    SeqsRecord rec = SeqsRecord();
    bool ready = this->getInput(rec);
    while (!this->isFinal() && !ready) {                       // FPGAPipeline.cpp:601-605
      std::this_thread::sleep_for(std::chrono::microseconds(10));
      ready = this->getInput(rec);
    }
    if (!ready && !(ready = this->getInput(rec))) break;
    {                                                          // reads this backend cannot take
      const int max_len = bwaflow_max_read_len(env_ ? &env_->chain_opt() : nullptr);
      int bad = -1;
      for (int i = 0; i < rec.batch_num && bad < 0; ++i) if (rec.seqs[i].l_seq > max_len) bad = i;
      if (bad >= 0) {
        if (unsupported_) { this->pushOutput(unsupported_(rec)); continue; }
        throw std::runtime_error("SeqsToChainsCUDA: read " + std::to_string(rec.start_idx + (uint64_t)bad) + " has " + std::to_string(rec.seqs[bad].l_seq) +
                                 " bases; this backend takes reads of up to " + std::to_string(max_len) +
                                 " (mem_flt_chained_seeds re-scoring and the 1 024-base row state are not implemented): route the record to the CPU stages");
      }
    }
    bases.clear(); off.assign(1, 0);
    for (int i = 0; i < rec.batch_num; ++i) {                  // in-place nt4, as stage 2 must (src/bwa_wrapper.cpp:124-125)
      bseq1_t *s = &rec.seqs[i];
      for (int j = 0; j < s->l_seq; ++j) s->seq[j] = (char)nt4(s->seq[j]);
      bases.insert(bases.end(), (uint8_t *)s->seq, (uint8_t *)s->seq + s->l_seq);
      off.push_back((int64_t)bases.size());
    }
    ChainsRecord out;
    out.start_idx = rec.start_idx; out.batch_num = rec.batch_num; out.seqs = rec.seqs;
    out.chains = NULL; out.bwtintvs = NULL; out.bwtintv_nums = NULL; out.alnreg = NULL; out.chain_ref = NULL;
    if (pool_) {
      bwaflow_hot_batch_t *task = pool_->take();
      check(bwaflow_hot_submit(task, rec.batch_num, bases.data(), off.data()));
      out.chain_ref = (mem_chainref_t **)task;                 // the chains stay in HBM; the handle travels in the unused field
      out.tag = 1;
    } else {
      check(bwaflow_smem_submit(smem, rec.batch_num, bases.data(), off.data()));
      const bwaflow_intv_t *iv; const int64_t *ivoff; int64_t n_intv;
      check(bwaflow_smem_collect(smem, &iv, &ivoff, &n_intv, nullptr));
      out.bwtintvs = (bwtintv_t **)malloc(rec.batch_num * sizeof(bwtintv_t *));
      out.bwtintv_nums = (size_t *)malloc(rec.batch_num * sizeof(size_t));
      for (int i = 0; i < rec.batch_num; ++i) {                // ownership moves downstream, as FPGAPipeline.cpp:716-719
        const size_t k = (size_t)(ivoff[i + 1] - ivoff[i]);
        out.bwtintv_nums[i] = k;
        out.bwtintvs[i] = (bwtintv_t *)malloc((k ? k : 1) * sizeof(bwtintv_t));
        memcpy(out.bwtintvs[i], iv + ivoff[i], k * sizeof(bwtintv_t));             // bwaflow_intv_t == bwtintv_t (32 bytes)
      }
      const int64_t *pos, *poff; int64_t n_pos;
      check(bwaflow_smem_seed_positions(smem, &pos, &poff, &n_pos, nullptr));
      SeedPositionTable::Entry e;
      e.pos.assign(pos, pos + n_pos); e.pos_off.assign(poff, poff + n_intv + 1);
      pos_->put(rec.start_idx, std::move(e));
      out.tag = 0;
    }
    this->pushOutput(out);
  }
  if (smem) bwaflow_smem_batch_destroy(smem);
}

// tag == 1: pass through (src/Pipeline.cpp:409).  tag == 0: intervals -> chains on the host (src/Pipeline.cpp:410-438:
// mem_chain_postprocess + mem_chain_flt + mem_flt_chained_seeds), here through bwaflow_chain_build with the seed positions the
// device already looked up; the interval lists are freed as the reference does (:436-437).
ChainsRecord ChainsPipe::compute(ChainsRecord const &record)
{
  if (record.tag == 1) return record;
  if (!ctg_ || !pos_) throw std::runtime_error("ChainsPipe: interval-mode records need the contig table and the seed-position table");
  ChainsRecord out = record;
  const int n = record.batch_num;
  std::vector<int64_t> read_off(1, 0), intv_off(1, 0);
  std::vector<bwaflow_intv_t> intv;
  for (int i = 0; i < n; ++i) {
    read_off.push_back(read_off.back() + record.seqs[i].l_seq);
    const bwaflow_intv_t *p = (const bwaflow_intv_t *)record.bwtintvs[i];
    intv.insert(intv.end(), p, p + record.bwtintv_nums[i]);
    intv_off.push_back((int64_t)intv.size());
    free(record.bwtintvs[i]);
  }
  free(record.bwtintvs); free(record.bwtintv_nums);
  out.bwtintvs = NULL; out.bwtintv_nums = NULL;
  SeedPositionTable::Entry sp = pos_->take(record.start_idx);
  bwaflow_chain_result_t *res = nullptr;
  check(bwaflow_chain_build(&copt_, ctg_->l_pac, ctg_->n_seqs, ctg_->offset, ctg_->len, ctg_->is_alt, n, read_off.data(), intv.data(),
                            intv_off.data(), sp.pos.data(), sp.pos_off.data(), 1, &res));
  const bwaflow_chain_t *ch; const bwaflow_seed_t *sd; const int64_t *coff; int64_t n_ch, n_sd;
  bwaflow_chain_result_view(res, &ch, &n_ch, &sd, &n_sd, &coff);
  out.chains = (mem_chain_v *)malloc(n * sizeof(mem_chain_v));                  // as SeqsToChains::compute (Pipeline.cpp:344)
  for (int i = 0; i < n; ++i) {
    mem_chain_v &v = out.chains[i];
    v.n = v.m = (size_t)(coff[i + 1] - coff[i]);
    v.a = (mem_chain_t *)malloc((v.n ? v.n : 1) * sizeof(mem_chain_t));
    for (size_t j = 0; j < v.n; ++j) {
      const bwaflow_chain_t &f = ch[coff[i] + j];
      mem_chain_t &c = v.a[j];
      memset(&c, 0, sizeof(c));
      c.n = c.m = f.n_seeds; c.rid = f.rid; c.frac_rep = f.frac_rep; c.pos = f.pos;
      c.w = f.flags & 0x1FFFFFFFu; c.kept = (f.flags >> 29) & 3u; c.is_alt = f.flags >> 31;
      c.seeds = (mem_seed_t *)malloc((f.n_seeds ? f.n_seeds : 1) * sizeof(mem_seed_t));
      for (int k = 0; k < f.n_seeds; ++k) {
        const bwaflow_seed_t &s = sd[f.seed_off + k];
        c.seeds[k].rbeg = s.rbeg; c.seeds[k].qbeg = s.qbeg; c.seeds[k].len = s.len; c.seeds[k].score = s.score;
      }
    }
  }
  bwaflow_chain_result_free(res);
  return out;
}

void ChainsToRegionsCUDA::compute(int wid)
{
  bwaflow_ctx_t *ctx_ = env_ ? env_->getPE(wid).ctx : this->ctx_;      // interval mode: this worker's GPU
  bwaflow_c2a_batch_t *c2a = nullptr;                          // interval mode: this worker's SWTask, created on first use
  int64_t c2a_chains = 0, c2a_seeds = 0;
  for (;;) {
    ChainsRecord rec = ChainsRecord();
    bool ready = this->getInput(rec);
    while (!this->isFinal() && !ready) {                       // FPGAPipeline.cpp:394-406
      std::this_thread::sleep_for(std::chrono::microseconds(10));
      ready = this->getInput(rec);
    }
    if (!ready && !(ready = this->getInput(rec))) break;
    bwaflow_hot_batch_t *task = (bwaflow_hot_batch_t *)rec.chain_ref;
    const bwaflow_alnreg_t *regs; const int64_t *roff; int64_t n_regs;
    if (rec.tag == 1) check(bwaflow_hot_collect(task, &regs, &roff, &n_regs, nullptr));
    else {                                                     // mem_chain_v lists -> flat SoA -> bwaflow_c2a_*, then free the chains
      std::vector<uint8_t> bases; std::vector<int64_t> off(1, 0), coff(1, 0);
      std::vector<bwaflow_chain_t> ch; std::vector<bwaflow_seed_t> sd;
      for (int i = 0; i < rec.batch_num; ++i) {
        bases.insert(bases.end(), (uint8_t *)rec.seqs[i].seq, (uint8_t *)rec.seqs[i].seq + rec.seqs[i].l_seq);
        off.push_back((int64_t)bases.size());
        for (size_t j = 0; j < rec.chains[i].n; ++j) {         // mem_chain_t -> flat (src/bwa_wrapper.h:64-77)
          const mem_chain_t &c = rec.chains[i].a[j];
          bwaflow_chain_t f;
          f.pos = c.pos; f.n_seeds = c.n; f.rid = c.rid; f.frac_rep = c.frac_rep;
          f.flags = (uint32_t)c.w | (uint32_t)c.kept << 29 | (uint32_t)c.is_alt << 31; f.seed_off = (int64_t)sd.size();
          for (int k = 0; k < c.n; ++k) { bwaflow_seed_t s; s.rbeg = c.seeds[k].rbeg; s.qbeg = c.seeds[k].qbeg; s.len = c.seeds[k].len; s.score = c.seeds[k].score; s._pad = 0; sd.push_back(s); }
          ch.push_back(f);
          free(c.seeds);
        }
        coff.push_back((int64_t)ch.size());
        free(rec.chains[i].a);                                 // Pipeline.cpp:526-530
      }
      free(rec.chains);
      if (!c2a || (int64_t)ch.size() > c2a_chains || (int64_t)sd.size() > c2a_seeds) {
        if (c2a) bwaflow_c2a_batch_destroy(c2a);
        c2a_chains = (int64_t)ch.size() * 2 + 1024; c2a_seeds = (int64_t)sd.size() * 2 + 1024;
        check(bwaflow_c2a_batch_create(ctx_, max_reads_, max_bases_, c2a_chains, c2a_seeds, &c2a));
      }
      check(bwaflow_c2a_submit(c2a, rec.batch_num, bases.data(), off.data(), ch.data(), coff.data(), sd.data(), (int64_t)sd.size()));
      check(bwaflow_c2a_collect(c2a, &regs, &roff, &n_regs, nullptr));
    }
    RegionsRecord out;
    out.start_idx = rec.start_idx; out.batch_num = rec.batch_num; out.seqs = rec.seqs; out.chains = NULL;
    out.alnreg = (mem_alnreg_v *)malloc(rec.batch_num * sizeof(mem_alnreg_v));    // ownership moves downstream (Pipeline.cpp:516-544)
    for (int i = 0; i < rec.batch_num; ++i) {
      const size_t k = (size_t)(roff[i + 1] - roff[i]);
      out.alnreg[i].n = out.alnreg[i].m = k;
      out.alnreg[i].a = (mem_alnreg_t *)malloc((k ? k : 1) * sizeof(mem_alnreg_t));
      memcpy(out.alnreg[i].a, regs + roff[i], k * sizeof(mem_alnreg_t));           // bwaflow_alnreg_t == mem_alnreg_t (88 bytes)
    }
    if (rec.tag == 1) { if (env_) env_->give(task); else pool_->give(task); }
    this->pushOutput(out);
  }
  if (c2a) bwaflow_c2a_batch_destroy(c2a);
}

void RegionsToSamCUDA::compute(int wid)
{
  bwaflow_ctx_t *ctx = env_->getPE(wid).ctx;
  bwaflow_sam_batch_t *sam = nullptr;
  check(bwaflow_sam_batch_create(ctx, env_->max_reads(), env_->max_bases(), &cfg_.po, &cfg_.so, cfg_.ctg_names, cfg_.ctg_is_alt, cfg_.ctg_anno,
                                 cfg_.rg_id, &sam));
  std::vector<uint8_t> bases, quals;
  std::vector<char> names, comments;
  std::vector<int64_t> off, name_off, comment_off, reg_off;
  std::vector<bwaflow_alnreg_t> regs;
  for (;;) {
    RegionsRecord rec = RegionsRecord();
    bool ready = this->getInput(rec);
    while (!this->isFinal() && !ready) {
      std::this_thread::sleep_for(std::chrono::microseconds(10));
      ready = this->getInput(rec);
    }
    if (!ready && !(ready = this->getInput(rec))) break;
    const int n = rec.batch_num;
    bases.clear(); quals.clear(); names.clear(); comments.clear(); regs.clear();
    off.assign(1, 0); reg_off.assign(1, 0); name_off.clear(); comment_off.clear();
    bool have_qual = true, have_comment = false;
    for (int i = 0; i < n; ++i) if (!rec.seqs[i].qual) have_qual = false;
    for (int i = 0; i < n; ++i) {
      const bseq1_t *s = &rec.seqs[i];                         // s->seq is nt4 already (stage 2 converted it in place)
      bases.insert(bases.end(), (const uint8_t *)s->seq, (const uint8_t *)s->seq + s->l_seq);
      if (have_qual) quals.insert(quals.end(), (const uint8_t *)s->qual, (const uint8_t *)s->qual + s->l_seq);
      off.push_back((int64_t)bases.size());
      name_off.push_back((int64_t)names.size());
      names.insert(names.end(), s->name, s->name + strlen(s->name) + 1);
      if (s->comment) { comment_off.push_back((int64_t)comments.size()); comments.insert(comments.end(), s->comment, s->comment + strlen(s->comment) + 1); have_comment = true; }
      else comment_off.push_back(-1);
      const mem_alnreg_v &v = rec.alnreg[i];
      const bwaflow_alnreg_t *a = (const bwaflow_alnreg_t *)v.a;          // bwaflow_alnreg_t == mem_alnreg_t (88 bytes)
      regs.insert(regs.end(), a, a + v.n);
      reg_off.push_back((int64_t)regs.size());
    }
    if (!have_qual) for (int i = 0; i < n; ++i) if (rec.seqs[i].qual) throw std::runtime_error("RegionsToSamCUDA: a record mixes reads with and without qualities");
    const char *text; const int64_t *toff;
    check(bwaflow_sam_process(sam, n, bases.data(), have_qual ? quals.data() : nullptr, off.data(), names.data(), name_off.data(),
                              have_comment ? comments.data() : nullptr, have_comment ? comment_off.data() : nullptr, regs.data(), reg_off.data(),
                              (int64_t)rec.start_idx, &text, &toff, nullptr, nullptr));
    for (int i = 0; i < n; ++i) {
      bseq1_t *s = &rec.seqs[i];
      const size_t l = (size_t)(toff[i + 1] - toff[i]);
      s->sam = (char *)malloc(l + 1);
      memcpy(s->sam, text + toff[i], l);
      s->sam[l] = 0;
      free(s->name); free(s->comment); free(s->seq); free(s->qual);      // Pipeline.cpp:632-638
      s->name = s->comment = s->seq = s->qual = NULL;
      if (rec.alnreg[i].a) free(rec.alnreg[i].a);                         // freeAligns
    }
    free(rec.alnreg);
    SeqsRecord out;
    out.start_idx = rec.start_idx; out.batch_num = n; out.seqs = rec.seqs;
    this->pushOutput(out);
  }
  bwaflow_sam_batch_destroy(sam);
}
