// src/cuda/CUDAPipeline.h -- the CUDA counterparts of SeqsToChainsFPGA / ChainsToRegionsFPGA (src/fpga/FPGAPipeline.h:14-48):
// MapPartitionStage subclasses for kflow stage slots 2 and 4, driving libbwaflow_cuda.so through its C ABI only.
// "Fast mode": stage 2 submits the whole hot path (bwaflow_hot_submit) and pushes a ChainsRecord with tag = 1 whose chain_ref
// carries the device batch; ChainsPipe passes tag == 1 records through (src/Pipeline.cpp:409); stage 4 collects the regions.
#pragma once
#include "kflow_shim.h"      // <kflow/MapPartitionStage.h> etc. in a bwa-flow build
#include "records.h"         // "Pipeline.h" in a bwa-flow build
#include "../../include/bwaflow_cuda.h"

// Device batches shared by the two stages: stage 2 takes a free one, stage 4 hands it back after the collect -- the role of the
// two Task buffers every FPGA worker owns (src/fpga/FPGAPipeline.cpp:374-386).
class CUDABatchPool {
 public:
  CUDABatchPool(bwaflow_ctx_t *ctx, int n_batches, int64_t max_reads, int64_t max_bases);
  ~CUDABatchPool();
  bwaflow_hot_batch_t *take();
  void give(bwaflow_hot_batch_t *b);

 private:
  kestrelFlow::Queue<bwaflow_hot_batch_t *, 16> free_;
  std::vector<bwaflow_hot_batch_t *> all_;
};

class SeqsToChainsCUDA : public kestrelFlow::MapPartitionStage<SeqsRecord, ChainsRecord, INPUT_DEPTH, COMPUTE_DEPTH> {
 public:
  SeqsToChainsCUDA(int n_workers, CUDABatchPool *pool)
      : kestrelFlow::MapPartitionStage<SeqsRecord, ChainsRecord, INPUT_DEPTH, COMPUTE_DEPTH>(n_workers, false), pool_(pool) {}
  void compute(int wid);

 private:
  CUDABatchPool *pool_;
};

class ChainsPipe : public kestrelFlow::MapStage<ChainsRecord, ChainsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH> {
 public:
  ChainsPipe(int n = 1) : kestrelFlow::MapStage<ChainsRecord, ChainsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(n, false) {}
  ChainsRecord compute(ChainsRecord const &record);
};

class ChainsToRegionsCUDA : public kestrelFlow::MapPartitionStage<ChainsRecord, RegionsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH> {
 public:
  ChainsToRegionsCUDA(int n_workers, CUDABatchPool *pool)
      : kestrelFlow::MapPartitionStage<ChainsRecord, RegionsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(n_workers, false), pool_(pool) {}
  void compute(int wid);

 private:
  CUDABatchPool *pool_;
};
