// bwa's on-disk index (prefix.bwt .sa .pac .ann .alt) -> the arrays bwa_idx_load leaves in memory (bwa/bwa.c:246-290), which is what
// bwaflow_ctx_create takes.  In a bwa-flow build this is `aux->idx` and the file is not needed; the stand-alone driver has no bwa
// objects to link, so it reads the files itself.  Formats: bwt_restore_bwt / bwt_restore_sa (bwa/bwt.c:421-462), bns_restore_core
// (bwa/bntseq.c:96-166), the .pac trailer (bwa/bntseq.c:309-318), the .alt list (bwa/bntseq.c:181-203).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/bwaflow_cuda.h"

struct IndexFiles {
  std::vector<uint32_t> bwt;
  std::vector<uint64_t> sa;
  std::vector<uint8_t> pac;
  std::vector<int64_t> ctg_offset;
  std::vector<int32_t> ctg_len, is_alt;
  std::vector<uint8_t> is_alt_u8;
  std::vector<std::string> names, anno;
  std::vector<const char *> name_ptr, anno_ptr;
  bwaflow_index_t ix;
  // throws std::runtime_error on a missing or inconsistent file
  void load(const std::string &prefix);
};
