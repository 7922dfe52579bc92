// src/cuda/CUDAPipeline.h -- the CUDA counterparts of SeqsToChainsFPGA / ChainsToRegionsFPGA (src/fpga/FPGAPipeline.h:14-48):
// MapPartitionStage subclasses for kflow stage slots 2 and 4, driving libbwaflow_cuda.so through its C ABI only.
// "Fast mode": stage 2 submits the whole hot path (bwaflow_hot_submit) and pushes a ChainsRecord with tag = 1 whose chain_ref
// carries the device batch; ChainsPipe passes tag == 1 records through (src/Pipeline.cpp:409); stage 4 collects the regions.
#pragma once
#include <functional>
#include <map>
#include <string>
#include "kflow_shim.h"      // <kflow/MapPartitionStage.h> etc. in a bwa-flow build
#include "records.h"         // "Pipeline.h" in a bwa-flow build
#include "../../include/bwaflow_cuda.h"

// Device batches shared by the two stages: stage 2 takes a free one, stage 4 hands it back after the collect -- the role of the
// two Task buffers every FPGA worker owns (src/fpga/FPGAPipeline.cpp:374-386).
class CUDABatchPool {
 public:
  // copt / ctg_is_alt: the chaining options (from aux->opt) and bns->anns[].is_alt of the device-side ChainsPipe hop; null = defaults / no ALT contigs
  CUDABatchPool(bwaflow_ctx_t *ctx, int n_batches, int64_t max_reads, int64_t max_bases, const bwaflow_chain_opt_t *copt = nullptr,
                const int32_t *ctg_is_alt = nullptr);
  ~CUDABatchPool();
  bwaflow_hot_batch_t *take();
  void give(bwaflow_hot_batch_t *b);
  const std::vector<bwaflow_hot_batch_t *> &all() const { return all_; }

 private:
  kestrelFlow::Queue<bwaflow_hot_batch_t *, 16> free_;
  std::vector<bwaflow_hot_batch_t *> all_;
};

// "Interval mode" (the FPGA flow: SMemTask results go through ChainsPipe on the host, tag = 0): the seed positions the device
// computed for a record (bwaflow_smem_seed_positions, the bwt_sa loop of mem_chain_postprocess) wait here for ChainsPipe.
class SeedPositionTable {
 public:
  struct Entry { std::vector<int64_t> pos, pos_off; };
  void put(uint64_t start_idx, Entry &&e);
  Entry take(uint64_t start_idx);

 private:
  std::mutex m_;
  std::map<uint64_t, Entry> tab_;
};

// What the host chaining needs from bntseq_t (bwa/bntseq.h:56-64); is_alt may be null (no ALT contigs)
struct ContigTable { int64_t l_pac; int32_t n_seqs; const int64_t *offset; const int32_t *len; const int32_t *is_alt; };

// The role of BWAOCLEnv (src/fpga/BWAOCLEnv.h:67-258): owns the device side -- one context (index replica) and one batch pool per
// GPU -- and hands every stage worker its processing element: getPE(wid) binds worker wid to device wid % n_devices, as
// BWAOCLEnv::getPE binds a worker thread to an FPGA device/bank.  Records are dealt to the GPUs by whichever worker pops them from
// the kflow queue; a batch goes back to the pool it came from.
class CUDAEnv {
 public:
  struct PE { int device; bwaflow_ctx_t *ctx; CUDABatchPool *pool; };
  // n_devices <= 0: every visible GPU.  batches_per_device = device batches in flight per GPU (0: interval mode, no fused batches).
  CUDAEnv(const bwaflow_index_t *idx, const bwaflow_opt_t *opt, const bwaflow_chain_opt_t *copt, const int32_t *ctg_is_alt, int n_devices,
          int batches_per_device, int64_t max_reads, int64_t max_bases);
  ~CUDAEnv();
  int n_devices() const { return (int)pes_.size(); }
  PE &getPE(int wid) { return pes_[(size_t)wid % pes_.size()]; }
  void give(bwaflow_hot_batch_t *b) { owner_.at(b)->give(b); }
  const bwaflow_chain_opt_t &chain_opt() const { return copt_; }
  int64_t max_reads() const { return max_reads_; }
  int64_t max_bases() const { return max_bases_; }

 private:
  std::vector<PE> pes_;
  std::map<bwaflow_hot_batch_t *, CUDABatchPool *> owner_;
  bwaflow_chain_opt_t copt_;
  int64_t max_reads_, max_bases_;
};

class SeqsToChainsCUDA : public kestrelFlow::MapPartitionStage<SeqsRecord, ChainsRecord, INPUT_DEPTH, COMPUTE_DEPTH> {
 public:
  // fast mode: the whole hot path is submitted here, chains stay in HBM (tag = 1)
  SeqsToChainsCUDA(int n_workers, CUDABatchPool *pool)
      : kestrelFlow::MapPartitionStage<SeqsRecord, ChainsRecord, INPUT_DEPTH, COMPUTE_DEPTH>(n_workers, false), pool_(pool), ctx_(nullptr), pos_(nullptr),
        max_reads_(0), max_bases_(0) {}
  // interval mode: SMEM intervals per read as SeqsToChainsFPGA returns them (tag = 0, src/fpga/FPGAPipeline.cpp:765-788)
  SeqsToChainsCUDA(int n_workers, bwaflow_ctx_t *ctx, SeedPositionTable *pos, int64_t max_reads, int64_t max_bases)
      : kestrelFlow::MapPartitionStage<SeqsRecord, ChainsRecord, INPUT_DEPTH, COMPUTE_DEPTH>(n_workers, false), pool_(nullptr), ctx_(ctx), pos_(pos),
        max_reads_(max_reads), max_bases_(max_bases) {}
  // several GPUs: worker wid drives env->getPE(wid); fast mode when the environment holds fused batches, else interval mode
  SeqsToChainsCUDA(int n_workers, CUDAEnv *env, SeedPositionTable *pos = nullptr)
      : kestrelFlow::MapPartitionStage<SeqsRecord, ChainsRecord, INPUT_DEPTH, COMPUTE_DEPTH>(n_workers, false), env_(env), pool_(nullptr), ctx_(nullptr),
        pos_(pos), max_reads_(env->max_reads()), max_bases_(env->max_bases()) {}
  void compute(int wid);
  // Records holding a read longer than bwaflow_max_read_len() cannot run on this backend.  Without a handler such a record is
  // a std::runtime_error that names the read; an integration that keeps the reference's CPU stage registers it here
  // (e.g. [&](SeqsRecord const &r) { return cpu_seq2chain_stage.compute(r); }) and the record takes that route instead.
  void setUnsupportedHandler(std::function<ChainsRecord(SeqsRecord const &)> h) { unsupported_ = h; }

 private:
  CUDAEnv *env_ = nullptr;
  CUDABatchPool *pool_;
  bwaflow_ctx_t *ctx_;
  SeedPositionTable *pos_;
  int64_t max_reads_, max_bases_;
  std::function<ChainsRecord(SeqsRecord const &)> unsupported_;
};

class ChainsPipe : public kestrelFlow::MapStage<ChainsRecord, ChainsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH> {
 public:
  // copt: the chaining options of aux->opt (w, max_chain_gap, mask_level, drop_ratio, min_seed_len, max_occ ..); null = bwa's defaults
  ChainsPipe(int n = 1, const ContigTable *ctg = nullptr, SeedPositionTable *pos = nullptr, const bwaflow_chain_opt_t *copt = nullptr)
      : kestrelFlow::MapStage<ChainsRecord, ChainsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(n, false), ctg_(ctg), pos_(pos)
  { if (copt) copt_ = *copt; else bwaflow_chain_opt_default(&copt_); }
  ChainsRecord compute(ChainsRecord const &record);

 private:
  const ContigTable *ctg_;
  SeedPositionTable *pos_;
  bwaflow_chain_opt_t copt_;
};

class ChainsToRegionsCUDA : public kestrelFlow::MapPartitionStage<ChainsRecord, RegionsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH> {
 public:
  ChainsToRegionsCUDA(int n_workers, CUDABatchPool *pool, bwaflow_ctx_t *ctx = nullptr, int64_t max_reads = 0, int64_t max_bases = 0)
      : kestrelFlow::MapPartitionStage<ChainsRecord, RegionsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(n_workers, false), pool_(pool), ctx_(ctx),
        max_reads_(max_reads), max_bases_(max_bases) {}
  ChainsToRegionsCUDA(int n_workers, CUDAEnv *env)
      : kestrelFlow::MapPartitionStage<ChainsRecord, RegionsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(n_workers, false), env_(env), pool_(nullptr), ctx_(nullptr),
        max_reads_(env->max_reads()), max_bases_(env->max_bases()) {}
  void compute(int wid);

 private:
  CUDAEnv *env_ = nullptr;
  CUDABatchPool *pool_;
  bwaflow_ctx_t *ctx_;
  int64_t max_reads_, max_bases_;
};

// Slot 5: RegionsToSam (src/Pipeline.cpp:546-648) on the device where it is dynamic programming (patch scores, mate rescue, CIGARs)
// and on this worker's thread where it is a per-pair decision or text; every worker owns one bwaflow_sam_batch on its GPU.
// Output as the reference's: seqs[i].sam set, name / comment / seq / qual and the regions freed.
struct SamStageConfig {
  bwaflow_pe_opt_t po; bwaflow_sam_opt_t so;
  const char *const *ctg_names; const uint8_t *ctg_is_alt; const char *const *ctg_anno; const char *rg_id;
};
class RegionsToSamCUDA : public kestrelFlow::MapPartitionStage<RegionsRecord, SeqsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH> {
 public:
  RegionsToSamCUDA(int n_workers, CUDAEnv *env, const SamStageConfig &cfg)
      : kestrelFlow::MapPartitionStage<RegionsRecord, SeqsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(n_workers, false), env_(env), cfg_(cfg) {}
  void compute(int wid);

 private:
  CUDAEnv *env_;
  SamStageConfig cfg_;
};
