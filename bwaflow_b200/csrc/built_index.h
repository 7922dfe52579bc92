// Host-side product of the index builders (index_build.cpp: host threads; index_build_gpu.cu: sm_100a).
#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/bwaflow_cuda.h"

struct bwaflow_built_index {
  std::vector<uint32_t> bwt;
  std::vector<uint64_t> sa;
  std::vector<uint8_t> pac;
  std::vector<int64_t> ctg_offset;
  std::vector<int32_t> ctg_len;
  bwaflow_index_t view;
};
