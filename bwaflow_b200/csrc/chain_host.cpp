// Host entry point of the chaining core (chain_core.h): the role of ChainsPipe::compute (src/Pipeline.cpp:408-438) for
// integrations that keep the hop between the two device stages on the CPU, exactly like the FPGA flow.  Multi-threaded
// over reads (std::thread), results written in parallel into the final flat arrays.  The device path runs the same
// core inside chain_dev.cu.
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "chain_core.h"

namespace bwaflow {
void set_error(const char *fmt, ...);
}
using namespace bwaflow::chain;

struct bwaflow_chain_result {
  std::vector<bwaflow_chain_t> chains;
  std::vector<bwaflow_seed_t> seeds;
  std::vector<int64_t> chain_off;
};

namespace {

constexpr int64_t BLOCK = 256;

struct Job {
  const bwaflow_chain_opt_t *opt; Ctg ctg;
  int64_t n_reads; const int64_t *read_off;
  const bwaflow_intv_t *intv; const int64_t *intv_off; const int64_t *pos; const int64_t *pos_off;
  // pass 1 output, per block of reads: chains (seed_off relative to the block's seeds) and seeds
  std::vector<std::vector<bwaflow_chain_t>> blk_chains;
  std::vector<std::vector<bwaflow_seed_t>> blk_seeds;
  std::vector<int64_t> n_chains_of_read, blk_chain_base, blk_seed_base;
  bwaflow_chain_result *res;
  std::atomic<int64_t> next{0};
  std::atomic<int> err{0};
};

struct HostScratch {
  std::vector<SeedNode> seeds; std::vector<Chain> chains, sorted; std::vector<Node> nodes; std::vector<int> order, kept;
  Scratch view(int64_t P)
  {
    if ((int64_t)seeds.size() < P + 1) { seeds.resize(P + 1); chains.resize(P + 1); sorted.resize(P + 1); order.resize(P + 1); kept.resize(P + 1); }
    if ((int64_t)nodes.size() < node_cap(P)) nodes.resize(node_cap(P));
    Scratch S; S.seeds = seeds.data(); S.chains = chains.data(); S.nodes = nodes.data(); S.order = order.data();
    S.sorted = sorted.data(); S.kept = kept.data(); S.n_seeds = S.n_chains = S.n_nodes = 0; S.root = 0;
    return S;
  }
};

static void pass1(Job *J)
{
  HostScratch H;
  const bwaflow_chain_opt_t &o = *J->opt;
  for (;;) {
    const int64_t blk = J->next.fetch_add(1), r0 = blk * BLOCK;
    if (r0 >= J->n_reads) break;
    const int64_t r1 = r0 + BLOCK < J->n_reads ? r0 + BLOCK : J->n_reads;
    std::vector<bwaflow_chain_t> &oc = J->blk_chains[blk];
    std::vector<bwaflow_seed_t> &os = J->blk_seeds[blk];
    for (int64_t r = r0; r < r1; ++r) {
      const int len = (int)(J->read_off[r + 1] - J->read_off[r]);
      const int64_t i0 = J->intv_off[r], i1 = J->intv_off[r + 1];
      if (len >= o.min_seed_len && needs_seed_rescoring(o, len)) { J->err.store(1); continue; }
      Scratch S = H.view(J->pos_off[i1] - J->pos_off[i0]);
      float frac_rep;
      const int n = chain_read(S, o, J->ctg, len, J->intv + i0, i1 - i0, J->pos, J->pos_off + i0, &frac_rep);
      for (int c = 0; c < n; ++c) {
        const Chain &ch = S.sorted[c];
        bwaflow_chain_t f; f.pos = ch.pos; f.n_seeds = ch.n; f.rid = ch.rid; f.frac_rep = frac_rep;
        f.flags = (uint32_t)ch.w | (uint32_t)ch.kept << 29 | (uint32_t)ch.is_alt << 31; f.seed_off = (int64_t)os.size();
        for (int j = ch.head; j >= 0; j = S.seeds[j].next) os.push_back(S.seeds[j].s);
        oc.push_back(f);
      }
      J->n_chains_of_read[r] = n;
    }
  }
}

static void pass2(Job *J)
{ // copy every block into its final place
  const int64_t n_blk = (int64_t)J->blk_chains.size();
  for (;;) {
    const int64_t blk = J->next.fetch_add(1);
    if (blk >= n_blk) break;
    const int64_t cb = J->blk_chain_base[blk], sb = J->blk_seed_base[blk];
    bwaflow_chain_t *dc = J->res->chains.data() + cb;
    const std::vector<bwaflow_chain_t> &oc = J->blk_chains[blk];
    for (size_t i = 0; i < oc.size(); ++i) { dc[i] = oc[i]; dc[i].seed_off += sb; }
    if (!J->blk_seeds[blk].empty()) memcpy(J->res->seeds.data() + sb, J->blk_seeds[blk].data(), J->blk_seeds[blk].size() * sizeof(bwaflow_seed_t));
  }
}

static void run_parallel(void (*fn)(Job *), Job *J, int n_threads)
{
  J->next.store(0);
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(fn, J);
  fn(J);
  for (auto &t : th) t.join();
}

} // namespace

extern "C" {

void bwaflow_chain_opt_default(bwaflow_chain_opt_t *o)
{ // mem_opt_init, bwa/bwamem.c:48-84
  o->w = 100; o->max_chain_gap = 10000; o->min_chain_weight = 0; o->max_chain_extend = 1 << 30;
  o->min_seed_len = 19; o->max_occ = 500; o->mask_level = 0.50f; o->drop_ratio = 0.50f;
}

int32_t bwaflow_max_read_len(const bwaflow_chain_opt_t *opt)
{
  bwaflow_chain_opt_t d;
  if (!opt) { bwaflow_chain_opt_default(&d); opt = &d; }
  int32_t len = 1023;                                        // row state of the seeding kernels
  while (len > 0 && needs_seed_rescoring(*opt, len)) --len;  // monotone: longer reads are re-scored, shorter ones are not
  return len;
}

int bwaflow_chain_build(const bwaflow_chain_opt_t *opt, int64_t l_pac, int32_t n_seqs, const int64_t *ctg_offset,
                        const int32_t *ctg_len, const int32_t *ctg_is_alt,
                        int64_t n_reads, const int64_t *read_off,
                        const bwaflow_intv_t *intv, const int64_t *intv_off, const int64_t *pos, const int64_t *pos_off,
                        int n_threads, bwaflow_chain_result_t **out)
{
  if (!opt || !out || (n_reads && (!read_off || !intv_off || !pos_off))) { bwaflow::set_error("bwaflow_chain_build: null argument"); return -22; }
  Job J;
  J.opt = opt; J.ctg = Ctg{n_seqs, ctg_offset, ctg_len, ctg_is_alt, l_pac};
  J.n_reads = n_reads; J.read_off = read_off; J.intv = intv; J.intv_off = intv_off; J.pos = pos; J.pos_off = pos_off;
  const int64_t n_blk = (n_reads + BLOCK - 1) / BLOCK;
  J.blk_chains.resize(n_blk); J.blk_seeds.resize(n_blk); J.n_chains_of_read.assign(n_reads, 0);
  if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > n_blk) n_threads = n_blk > 0 ? (int)n_blk : 1;
  run_parallel(pass1, &J, n_threads);
  if (J.err.load()) { bwaflow::set_error("read too long for chaining: mem_flt_chained_seeds (bwa/bwamem.c:607) would rescore its seeds"); return -34; }
  bwaflow_chain_result *R = new bwaflow_chain_result();
  J.res = R;
  J.blk_chain_base.resize(n_blk); J.blk_seed_base.resize(n_blk);
  int64_t nc = 0, ns = 0;
  for (int64_t b = 0; b < n_blk; ++b) { J.blk_chain_base[b] = nc; J.blk_seed_base[b] = ns; nc += (int64_t)J.blk_chains[b].size(); ns += (int64_t)J.blk_seeds[b].size(); }
  R->chains.resize(nc); R->seeds.resize(ns); R->chain_off.resize(n_reads + 1);
  run_parallel(pass2, &J, n_threads);
  R->chain_off[0] = 0;
  for (int64_t r = 0; r < n_reads; ++r) R->chain_off[r + 1] = R->chain_off[r] + J.n_chains_of_read[r];
  *out = R;
  return 0;
}

void bwaflow_chain_result_view(const bwaflow_chain_result_t *r, const bwaflow_chain_t **chains, int64_t *n_chains,
                               const bwaflow_seed_t **seeds, int64_t *n_seeds, const int64_t **chain_off)
{
  if (chains) *chains = r->chains.data();
  if (n_chains) *n_chains = (int64_t)r->chains.size();
  if (seeds) *seeds = r->seeds.data();
  if (n_seeds) *n_seeds = (int64_t)r->seeds.size();
  if (chain_off) *chain_off = r->chain_off.data();
}

void bwaflow_chain_result_free(bwaflow_chain_result_t *r) { delete r; }

} // extern "C"
