// Stand-alone harness for the src/cuda stage classes: the three-stage compute_flow slice of src/main.cpp:302-335
// (SeqsToChains -> ChainsPipe -> ChainsToRegions) behind a source that cuts the reads into kflow records and a sink that restores
// read order, over the boost-free kflow stand-in.  Input: a case file written by tests/test_host_cpp_gpu.py (index arrays as bwa
// holds them + ASCII reads); output: the mem_alnreg_t lists of all reads.  Everything below the stage classes is the C ABI.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include "CUDAPipeline.h"

struct CaseHeader {
  int64_t l_pac, n_seqs; uint64_t primary, L2[5], seq_len, bwt_size, n_sa; int64_t sa_intv, n_reads, n_bases, reads_per_record;
};

template <class T> static void rd(FILE *f, std::vector<T> &v, size_t n)
{
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short case file\n"); exit(2); }
}

class ReadsSource : public kestrelFlow::SourceStage<SeqsRecord, INPUT_DEPTH> {
 public:
  ReadsSource(std::vector<bseq1_t> *seqs, int per_record) : seqs_(seqs), per_(per_record) {}
  void compute()
  {
    const int n = (int)seqs_->size();
    for (int s = 0; s < n; s += per_) {                       // getKseqBatch cuts whole records (src/Pipeline.cpp:98-130)
      SeqsRecord r; r.start_idx = (uint64_t)s; r.batch_num = n - s < per_ ? n - s : per_; r.seqs = seqs_->data() + s;
      pushOutput(r);
    }
  }
 private:
  std::vector<bseq1_t> *seqs_; int per_;
};

class RegionsSink : public kestrelFlow::SinkStage<RegionsRecord, COMPUTE_DEPTH> {
 public:
  std::map<uint64_t, RegionsRecord> got;                       // by start_idx: what SamsReorder does (src/Pipeline.cpp:651-737)
  void compute()
  {
    for (;;) {
      RegionsRecord r = RegionsRecord();
      if (getInput(r)) { got[r.start_idx] = r; continue; }
      if (isFinal()) { if (getInput(r)) { got[r.start_idx] = r; continue; } break; }
      std::this_thread::sleep_for(std::chrono::microseconds(10));
    }
  }
};

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: %s case.bin regions.bin [n_workers] [fast|intervals] [n_gpus (0 = all)] [--use_gpu]\n", argv[0]); return 2; }
  const int n_workers = argc > 3 ? atoi(argv[3]) : 2;
  const bool fast = !(argc > 4 && !strcmp(argv[4], "intervals"));
  const int n_gpus = argc > 5 ? atoi(argv[5]) : 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  CaseHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1) return 2;
  std::vector<uint32_t> bwt; std::vector<uint64_t> sa; std::vector<uint8_t> pac; std::vector<int64_t> coff, roff; std::vector<int32_t> clen;
  std::vector<char> bases;
  rd(f, bwt, h.bwt_size); rd(f, sa, h.n_sa); rd(f, pac, (size_t)(h.l_pac / 4 + 1)); rd(f, coff, h.n_seqs); rd(f, clen, h.n_seqs);
  rd(f, roff, h.n_reads + 1); rd(f, bases, h.n_bases);
  fclose(f);
  bwaflow_index_t ix;
  memset(&ix, 0, sizeof(ix));
  ix.primary = h.primary; memcpy(ix.L2, h.L2, sizeof(ix.L2)); ix.seq_len = h.seq_len; ix.bwt_size = h.bwt_size; ix.bwt = bwt.data();
  ix.n_sa = h.n_sa; ix.sa_intv = h.sa_intv; ix.sa = sa.data(); ix.l_pac = h.l_pac; ix.pac = pac.data(); ix.n_seqs = (int32_t)h.n_seqs;
  ix.ctg_offset = coff.data(); ix.ctg_len = clen.data();
  std::vector<bseq1_t> seqs(h.n_reads);
  int64_t max_rec_bases = 0;
  for (int64_t i = 0; i < h.n_reads; ++i) {
    memset(&seqs[i], 0, sizeof(bseq1_t));
    seqs[i].l_seq = (int)(roff[i + 1] - roff[i]); seqs[i].seq = bases.data() + roff[i]; seqs[i].id = (int)i;
  }
  for (int64_t s = 0; s < h.n_reads; s += h.reads_per_record) {
    const int64_t e = s + h.reads_per_record < h.n_reads ? s + h.reads_per_record : h.n_reads;
    if (roff[e] - roff[s] > max_rec_bases) max_rec_bases = roff[e] - roff[s];
  }
  int rc = 0;
  try {
    // what main() does under --use_gpu (the place of `new BWAOCLEnv()` under --use_fpga, src/main.cpp:201-231): options from aux->opt,
    // one index replica and two batches per worker on every GPU
    bwaflow_opt_t opt; bwaflow_opt_default(&opt);
    bwaflow_chain_opt_t copt; bwaflow_chain_opt_default(&copt);
    CUDAEnv cuda_env(&ix, &opt, &copt, /*bns->anns[].is_alt*/ nullptr, n_gpus, fast ? 2 * n_workers : 0, h.reads_per_record, max_rec_bases + 16);
    SeedPositionTable seed_pos;
    ContigTable ctg = {h.l_pac, (int32_t)h.n_seqs, coff.data(), clen.data(), nullptr};
    ReadsSource src(&seqs, (int)h.reads_per_record);
    SeqsToChainsCUDA seq2chain_cuda_stage(n_workers, &cuda_env, &seed_pos);
    ChainsPipe chainpipe_stage(1, &ctg, &seed_pos, &copt);
    ChainsToRegionsCUDA chain2reg_cuda_stage(n_workers, &cuda_env);
    fprintf(stderr, "CUDAEnv: %d GPU(s), %d workers per stage, %s mode\n", cuda_env.n_devices(), n_workers, fast ? "fast" : "interval");
    RegionsSink sink;
    kestrelFlow::Pipeline compute_flow(5, n_workers);
    compute_flow.addStage(0, &src);
    compute_flow.addStage(1, &seq2chain_cuda_stage);           // slot 2 of the full pipeline (src/main.cpp:302-316)
    compute_flow.addStage(2, &chainpipe_stage);                // slot 3
    compute_flow.addStage(3, &chain2reg_cuda_stage);           // slot 4 (src/main.cpp:318-335)
    compute_flow.addStage(4, &sink);
    compute_flow.start();
    compute_flow.wait();
    FILE *o = fopen(argv[2], "wb");
    std::vector<int64_t> off(1, 0);
    std::vector<mem_alnreg_t> all;
    for (auto &kv : sink.got)
      for (int i = 0; i < kv.second.batch_num; ++i) {
        const mem_alnreg_v &v = kv.second.alnreg[i];
        all.insert(all.end(), v.a, v.a + v.n);
        off.push_back((int64_t)all.size());
        free(v.a);
      }
    for (auto &kv : sink.got) free(kv.second.alnreg);
    const int64_t n = (int64_t)off.size() - 1;
    fwrite(&n, 8, 1, o); fwrite(off.data(), 8, off.size(), o); fwrite(all.data(), sizeof(mem_alnreg_t), all.size(), o);
    fclose(o);
    fprintf(stderr, "%lld reads in %zu records -> %zu regions\n", (long long)n, sink.got.size(), all.size());
  } catch (std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); rc = 1; }
  return rc;
}
