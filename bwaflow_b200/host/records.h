// The record and bwa structs the two stage slots exchange, layout-compatible with the reference on x86-64
// (src/Pipeline.h:39-66; bwa/bwa.h:33-46 with Falcon's `bams`; src/bwa_wrapper.h:64-77; bwa/bwamem.h:60-78).
// In a bwa-flow build these come from the reference's own headers; this file exists only because boost/htslib are absent here.
#pragma once
#include <stddef.h>
#include <stdint.h>

typedef struct { int l_seq, id; char *name, *comment, *seq, *qual, *sam; void *bams; } bseq1_t;          // 56 bytes
typedef struct { int64_t rbeg; int32_t qbeg, len; int score; } mem_seed_t;                                 // 24 bytes
typedef struct { int n, m, first, rid; uint32_t w : 29, kept : 2, is_alt : 1; float frac_rep; int64_t pos; mem_seed_t *seeds; } mem_chain_t;   // 40 bytes
typedef struct { size_t n, m; mem_chain_t *a; } mem_chain_v;
typedef struct {
  int64_t rb, re; int qb, qe; int rid; int score; int truesc; int sub; int alt_sc; int csub; int sub_n; int w; int seedcov; int secondary;
  int secondary_all; int seedlen0; int n_comp : 30, is_alt : 2; float frac_rep; uint64_t hash;
} mem_alnreg_t;                                                                                             // 88 bytes
typedef struct { size_t n, m; mem_alnreg_t *a; } mem_alnreg_v;
typedef struct { uint64_t x[3], info; } bwtintv_t;                                                          // 32 bytes
typedef struct mem_chainref_s mem_chainref_t;

static_assert(sizeof(bseq1_t) == 56 && sizeof(mem_seed_t) == 24 && sizeof(mem_chain_t) == 40 && sizeof(mem_alnreg_t) == 88 && sizeof(bwtintv_t) == 32,
              "record structs must keep the reference's layout");

struct SeqsRecord { uint64_t start_idx; int batch_num; bseq1_t *seqs; const char *name = "SeqsRecord"; };
struct ChainsRecord {
  uint64_t start_idx; int batch_num; bseq1_t *seqs; mem_chain_v *chains; bwtintv_t **bwtintvs; size_t *bwtintv_nums;
  mem_alnreg_v *alnreg; mem_chainref_t **chain_ref; const char *name = "ChainsRecord"; int tag;
};
struct RegionsRecord { uint64_t start_idx; int batch_num; bseq1_t *seqs; mem_chain_v *chains; mem_alnreg_v *alnreg; const char *name = "RegionsRecord"; };

#define INPUT_DEPTH 64       // src/Pipeline.h:20-22
#define COMPUTE_DEPTH 64
