// Stand-alone check of the ingest stage: FastqSource (host/Ingest.cpp) behind a two-stage kflow pipeline, every record dumped as
// "#record <start_idx> <batch_num>" + one "name\tcomment\tseq\tqual" line per read.  tests/test_ingest.py compares the dump with
// the reference's own kseq_read + bseq_read on the same files (oracle/_ref).  No GPU involved.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Ingest.h"

class DumpSink : public kestrelFlow::SinkStage<SeqsRecord, COMPUTE_DEPTH> {
 public:
  void compute()
  {
    for (;;) {
      SeqsRecord r = SeqsRecord();
      bool ready = getInput(r);
      if (!ready) {
        if (isFinal()) { if (!(ready = getInput(r))) break; }
        else { std::this_thread::sleep_for(std::chrono::microseconds(10)); continue; }
      }
      printf("#record %llu %d\n", (unsigned long long)r.start_idx, r.batch_num);
      for (int i = 0; i < r.batch_num; ++i) {
        const bseq1_t *s = &r.seqs[i];
        printf("%s\t%s\t%s\t%s\n", s->name, s->comment ? s->comment : "", s->seq, s->qual ? s->qual : "*");
        free(s->name); free(s->comment); free(s->seq); free(s->qual);
      }
      free(r.seqs);
    }
  }
};

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: %s chunk_size in1 [in2]\n", argv[0]); return 2; }
  try {
    FastqSource src(argv[2], argc > 3 ? argv[3] : nullptr, /*copy_comment=*/true, atoi(argv[1]));
    DumpSink sink;
    kestrelFlow::Pipeline flow(2, 1);
    flow.addStage(0, &src);
    flow.addStage(1, &sink);
    flow.start();
    flow.wait();
  } catch (std::exception &e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
