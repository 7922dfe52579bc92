// CPU-only self-test of kflow_shim.h: a source, a MapStage, a multi-worker MapPartitionStage and a sink over small queues; every
// record must come out exactly once whatever the interleaving, and the pipeline must terminate (isFinal race, bounded queues).
#include <stdio.h>
#include <stdlib.h>
#include <set>
#include "kflow_shim.h"

struct Rec { long id; long v; };

class Src : public kestrelFlow::SourceStage<Rec, 4> {
 public:
  explicit Src(long n) : n_(n) {}
  void compute() { for (long i = 0; i < n_; ++i) { Rec r = {i, i}; pushOutput(r); } }
 private:
  long n_;
};
class Twice : public kestrelFlow::MapStage<Rec, Rec, 4, 4> {
 public:
  Twice() : kestrelFlow::MapStage<Rec, Rec, 4, 4>(2, false) {}
  Rec compute(Rec const &r) { Rec o = {r.id, r.v * 2}; return o; }
};
class PlusOne : public kestrelFlow::MapPartitionStage<Rec, Rec, 4, 4> {
 public:
  PlusOne() : kestrelFlow::MapPartitionStage<Rec, Rec, 4, 4>(3, false) {}
  void compute(int)
  {
    for (;;) {
      Rec r = {0, 0};
      bool ready = getInput(r);
      while (!isFinal() && !ready) { std::this_thread::sleep_for(std::chrono::microseconds(10)); ready = getInput(r); }
      if (!ready && !(ready = getInput(r))) break;
      r.v += 1;
      pushOutput(r);
    }
  }
};
class Snk : public kestrelFlow::SinkStage<Rec, 4> {
 public:
  std::set<long> seen; long bad = 0;
  void compute()
  {
    for (;;) {
      Rec r = {0, 0};
      if (getInput(r)) { if (r.v != 2 * r.id + 1 || !seen.insert(r.id).second) ++bad; continue; }
      if (isFinal()) { if (getInput(r)) { if (r.v != 2 * r.id + 1 || !seen.insert(r.id).second) ++bad; continue; } break; }
      std::this_thread::sleep_for(std::chrono::microseconds(10));
    }
  }
};

int main(int argc, char **argv)
{
  const long n = argc > 1 ? atol(argv[1]) : 20000;
  for (int rep = 0; rep < 3; ++rep) {
    Src src(n); Twice a; PlusOne b; Snk snk;
    kestrelFlow::Pipeline p(4, 8);
    p.addStage(0, &src); p.addStage(1, &a); p.addStage(2, &b); p.addStage(3, &snk);
    p.start();
    p.wait();
    if (snk.bad || (long)snk.seen.size() != n) { fprintf(stderr, "FAIL: %ld records, %ld bad\n", (long)snk.seen.size(), snk.bad); return 1; }
  }
  printf("ok %ld records x 3\n", n);
  return 0;
}
