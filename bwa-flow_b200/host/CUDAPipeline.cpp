// src/cuda/CUDAPipeline.cpp -- worker loops shaped like SeqsToChainsFPGA::compute / ChainsToRegionsFPGA::compute
// (src/fpga/FPGAPipeline.cpp:582-800, 367-555) without their CPU fallbacks: errors are fatal (std::runtime_error, src/util.h:15-31).
#include <stdlib.h>
#include <string.h>
#include "CUDAPipeline.h"

static void check(int rc) { if (rc) throw std::runtime_error(bwaflow_last_error()); }

CUDABatchPool::CUDABatchPool(bwaflow_ctx_t *ctx, int n_batches, int64_t max_reads, int64_t max_bases)
{
  for (int i = 0; i < n_batches; ++i) {
    bwaflow_hot_batch_t *b = nullptr;
    check(bwaflow_hot_batch_create(ctx, max_reads, max_bases, nullptr, nullptr, &b));
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

// nst_nt4_table (bwa/bntseq.c:47-64): A,C,G,T (either case) -> 0..3, everything else 4
static inline uint8_t nt4(char c)
{
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return (uint8_t)c < 4 ? (uint8_t)c : 4; }
}

void SeqsToChainsCUDA::compute(int wid)
{
  (void)wid;
  std::vector<uint8_t> bases;
  std::vector<int64_t> off;
  for (;;) {
    SeqsRecord rec = SeqsRecord();
    bool ready = this->getInput(rec);
    while (!this->isFinal() && !ready) {                       // FPGAPipeline.cpp:601-605
      std::this_thread::sleep_for(std::chrono::microseconds(10));
      ready = this->getInput(rec);
    }
    if (!ready && !(ready = this->getInput(rec))) break;
    bases.clear(); off.assign(1, 0);
    for (int i = 0; i < rec.batch_num; ++i) {                  // in-place nt4, as stage 2 must (src/bwa_wrapper.cpp:124-125)
      bseq1_t *s = &rec.seqs[i];
      for (int j = 0; j < s->l_seq; ++j) s->seq[j] = (char)nt4(s->seq[j]);
      bases.insert(bases.end(), (uint8_t *)s->seq, (uint8_t *)s->seq + s->l_seq);
      off.push_back((int64_t)bases.size());
    }
    bwaflow_hot_batch_t *task = pool_->take();
    check(bwaflow_hot_submit(task, rec.batch_num, bases.data(), off.data()));
    ChainsRecord out;
    out.start_idx = rec.start_idx; out.batch_num = rec.batch_num; out.seqs = rec.seqs;
    out.chains = NULL; out.bwtintvs = NULL; out.bwtintv_nums = NULL; out.alnreg = NULL;
    out.chain_ref = (mem_chainref_t **)task;                   // the chains stay in HBM; the handle travels in the unused field
    out.tag = 1;
    this->pushOutput(out);
  }
}

ChainsRecord ChainsPipe::compute(ChainsRecord const &record)
{
  if (record.tag == 1) return record;                          // src/Pipeline.cpp:409
  throw std::runtime_error("ChainsPipe: tag 0 records (host chaining) are produced by the bwaflow_smem_* variant of stage 2");
}

void ChainsToRegionsCUDA::compute(int wid)
{
  (void)wid;
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
    check(bwaflow_hot_collect(task, &regs, &roff, &n_regs, nullptr));
    RegionsRecord out;
    out.start_idx = rec.start_idx; out.batch_num = rec.batch_num; out.seqs = rec.seqs; out.chains = NULL;
    out.alnreg = (mem_alnreg_v *)malloc(rec.batch_num * sizeof(mem_alnreg_v));    // ownership moves downstream (Pipeline.cpp:516-544)
    for (int i = 0; i < rec.batch_num; ++i) {
      const size_t k = (size_t)(roff[i + 1] - roff[i]);
      out.alnreg[i].n = out.alnreg[i].m = k;
      out.alnreg[i].a = (mem_alnreg_t *)malloc((k ? k : 1) * sizeof(mem_alnreg_t));
      memcpy(out.alnreg[i].a, regs + roff[i], k * sizeof(mem_alnreg_t));           // bwaflow_alnreg_t == mem_alnreg_t (88 bytes)
    }
    pool_->give(task);
    this->pushOutput(out);
  }
}
