// Ingest and egress around the CUDA stages: the stages of src/Pipeline.cpp that sit before slot 2 and after
// slot 5 -- KseqsRead + KseqsToBseqs (:98-202: FASTQ/FASTA parsing with kseq's rules, records of >= 10 Mbp holding whole pairs,
// trim_readno, comments dropped unless asked for), SamsReorder (:651-737: records released in start_idx order under
// --inorder_output) and a text writer.  Host code: gz inflation and line splitting have no device form worth the copy; the base
// conversion (nt4) happens in SeqsToChainsCUDA as in the reference's stage 2.
#pragma once
#include <stdio.h>
#include <zlib.h>
#include <map>
#include <string>
#include "kflow_shim.h"
#include "records.h"

// One FASTA/FASTQ stream read with kseq_read's rules (bwa/kseq.h:177-216): a record starts at '@' or '>', the name ends at the first
// white space, the rest of the header line is the comment, the sequence may span lines up to a line starting with '+', '>' or '@',
// the quality string is read until it is as long as the sequence.
class FastxReader {
 public:
  explicit FastxReader(const char *path);
  ~FastxReader();
  // >= 0: sequence length; -1: end of file; -2: truncated quality string
  int read(std::string &name, std::string &comment, std::string &seq, std::string &qual);

 private:
  int getc_();
  void rest_of_line(std::string &s, bool append);
  gzFile f_;
  unsigned char buf_[1 << 18];
  int begin_, end_, eof_, last_char_;
};

class FastqSource : public kestrelFlow::SourceStage<SeqsRecord, INPUT_DEPTH> {
 public:
  // in2 == NULL: single file (interleaved pairs or single-end); chunk_size = bases per record (10 000 000, src/Pipeline.cpp:146)
  FastqSource(const char *in1, const char *in2, bool copy_comment, int chunk_size = 10000000)
      : r1_(in1), r2_(in2 ? new FastxReader(in2) : nullptr), copy_comment_(copy_comment), chunk_size_(chunk_size) {}
  ~FastqSource() { delete r2_; }
  void compute();
  uint64_t n_reads() const { return produced_; }

 private:
  FastxReader r1_, *r2_;
  bool copy_comment_;
  int chunk_size_;
  uint64_t produced_ = 0;
};

class SamsReorder : public kestrelFlow::MapPartitionStage<SeqsRecord, SeqsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH> {
 public:
  explicit SamsReorder(bool inorder_output)
      : kestrelFlow::MapPartitionStage<SeqsRecord, SeqsRecord, COMPUTE_DEPTH, COMPUTE_DEPTH>(1, false), inorder_(inorder_output) {}
  void compute(int wid);

 private:
  bool inorder_;
};

// writes seqs[i].sam of every record and frees the record (the text path of SamsPrint / the BAM writer's place)
class SamsPrint : public kestrelFlow::SinkStage<SeqsRecord, COMPUTE_DEPTH> {
 public:
  explicit SamsPrint(FILE *out) : out_(out) {}
  void compute();
  uint64_t n_reads = 0, n_bytes = 0;

 private:
  FILE *out_;
};
