// FM-index construction for synthetic references, producing the same arrays `bwa index` leaves in a loaded
// bwaidx_t (bwa/bwtindex.c:256-324):
//   pac   2-bit forward reference, N replaced by lrand48()&3 after srand48(11)   (bwa/bntseq.c:260-262, 289-290)
//   bwt   BWT of forward + reverse complement with the '$' row removed, occ-interleaved every 128 symbols
//         (bwt_pac2bwt bwa/bwtindex.c:64-134, bwt_bwtupdate_core :151-173), primary, L2[]
//   sa    every 32nd row of the suffix array, sa[0] = (u64)-1                     (bwt_cal_sa bwa/bwt.c:62-84)
// The BWT of a text is unique, so any correct suffix sort gives bit-identical arrays.  This one buckets the suffixes
// by a k-mer prefix and sorts each bucket with word-wise comparisons on the 2-bit text, in parallel over
// buckets -- fast on the (near-random) synthetic references of the benchmark configs; it is a workload-preparation
// tool, not part of the timed path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "built_index.h"

namespace bwaflow {
void set_error(const char *fmt, ...);
}

namespace {

struct Text {
  int64_t n;                       // symbols
  std::vector<uint64_t> w;         // 32 symbols per word, MSB first, zero padded by two words
  inline int at(int64_t p) const { return (int)(w[p >> 5] >> ((~p & 31) << 1)) & 3; }
  inline uint64_t get32(int64_t p) const
  {
    const int64_t i = p >> 5; const int sh = (int)(p & 31) << 1;
    return sh ? (w[i] << sh) | (w[i + 1] >> (64 - sh)) : w[i];
  }
};

// suffix a < suffix b, both known to share their first `from` symbols (the end of the text sorts first)
static inline bool suf_less(const Text &T, int64_t a, int64_t b, int64_t from)
{
  for (int64_t off = from;; off += 32) {
    const int64_t la = T.n - a - off, lb = T.n - b - off;
    if (la <= 0 || lb <= 0) return la < lb;
    uint64_t wa = T.get32(a + off), wb = T.get32(b + off);
    const int64_t m = std::min<int64_t>(32, std::min(la, lb));
    if (m < 32) { const uint64_t mask = ~0ull << (64 - 2 * m); wa &= mask; wb &= mask; }
    if (wa != wb) return wa < wb;
    if (m < 32) return la < lb;
  }
}

template <class IDX>
static void build_sa(const Text &T, int n_threads, std::vector<IDX> &sa)
{
  const int64_t n = T.n;
  int k = 1;
  while (k < 13 && (1LL << (2 * (k + 1))) <= n / 4) ++k;      // ~4+ suffixes per bucket on random text
  const int64_t n_buckets = 1LL << (2 * k);
  std::vector<int64_t> start(n_buckets + 1, 0);
  auto kmer_at = [&](int64_t p) -> int64_t {                   // first k symbols, missing ones read as 0
    const int64_t rem = n - p;
    uint64_t v = T.get32(p) >> (64 - 2 * k);
    if (rem < k) v &= ~0ull << (2 * (k - rem));
    return (int64_t)v;
  };
  for (int64_t p = 0; p < n; ++p) ++start[kmer_at(p) + 1];
  for (int64_t b = 0; b < n_buckets; ++b) start[b + 1] += start[b];
  sa.resize(n);
  {
    std::vector<int64_t> fill(start.begin(), start.end() - 1);
    for (int64_t p = 0; p < n; ++p) sa[fill[kmer_at(p)]++] = (IDX)p;
  }
  std::atomic<int64_t> next{0};
  const int64_t chunk = std::max<int64_t>(1, n_buckets / (64 * (int64_t)n_threads));
  auto work = [&]() {
    for (;;) {
      const int64_t b0 = next.fetch_add(chunk);
      if (b0 >= n_buckets) break;
      const int64_t b1 = std::min(n_buckets, b0 + chunk);
      for (int64_t b = b0; b < b1; ++b) {
        IDX *lo = sa.data() + start[b], *hi = sa.data() + start[b + 1];
        if (hi - lo < 2) continue;
        std::sort(lo, hi, [&](IDX x, IDX y) {
          const int64_t a = (int64_t)x, c = (int64_t)y;
          const int64_t from = (n - a >= k && n - c >= k) ? k : 0;
          return suf_less(T, a, c, from);
        });
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}

template <class IDX>
static void finish(const Text &T, const std::vector<IDX> &sa, bwaflow_built_index *B)
{
  const int64_t n = T.n;
  // rows of the (n+1)-row matrix: row 0 is the '$' suffix, row i+1 is sa[i].  BWT symbol of a row = T[suffix-1].
  uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0};
  std::vector<uint32_t> plain((n + 15) >> 4, 0);
  int64_t out = 0;
  {
    const int c = T.at(n - 1);                               // row 0: suffix n ('$'), preceded by the last symbol
    plain[out >> 4] |= (uint32_t)c << ((15 - (out & 15)) << 1); ++out;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int64_t s = (int64_t)sa[i];
    if (s == 0) { primary = (uint64_t)(i + 1); continue; }   // this row would hold '$'
    const int c = T.at(s - 1);
    plain[out >> 4] |= (uint32_t)c << ((15 - (out & 15)) << 1); ++out;
  }
  for (int64_t p = 0; p < n; ++p) ++L2[1 + T.at(p)];
  for (int c = 2; c <= 4; ++c) L2[c] += L2[c - 1];
  // occ-interleave: before every 128 symbols, the four running counts as u64 (bwa/bwtindex.c:151-173)
  const uint64_t n_occ = (uint64_t)(n + 127) / 128 + 1;
  const uint64_t bwt_size = (uint64_t)((n + 15) >> 4) + n_occ * 8;
  B->bwt.assign(bwt_size, 0);
  uint64_t cnt[4] = {0, 0, 0, 0}, kk = 0;
  for (int64_t i = 0; i < n; ++i) {
    if ((i & 127) == 0) { memcpy(&B->bwt[kk], cnt, 32); kk += 8; }
    if ((i & 15) == 0) B->bwt[kk++] = plain[i >> 4];
    ++cnt[(plain[i >> 4] >> ((~i & 15) << 1)) & 3];
  }
  memcpy(&B->bwt[kk], cnt, 32);
  // sampled SA (bwa/bwt.c:62-84)
  const int intv = 32;
  const uint64_t n_sa = (uint64_t)(n + intv) / intv;
  B->sa.assign(n_sa, 0);
  for (uint64_t j = 1; j < n_sa; ++j) B->sa[j] = (uint64_t)sa[j * intv - 1];
  B->sa[0] = (uint64_t)-1;
  bwaflow_index_t &v = B->view;
  v.primary = primary; memcpy(v.L2, L2, sizeof(L2)); v.seq_len = (uint64_t)n; v.bwt_size = bwt_size; v.bwt = B->bwt.data();
  v.n_sa = n_sa; v.sa_intv = intv; v.sa = B->sa.data();
}

} // namespace

extern "C" {

int bwaflow_index_build(const uint8_t *codes, int64_t n_bases, int32_t n_contigs, const int64_t *contig_len,
                        int n_threads, bwaflow_built_index_t **out)
{
  if (!codes || !out || n_bases <= 0 || n_contigs <= 0 || !contig_len) { bwaflow::set_error("bwaflow_index_build: bad argument"); return -22; }
  if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  bwaflow_built_index *B = new bwaflow_built_index();
  int64_t tot = 0;
  for (int i = 0; i < n_contigs; ++i) { B->ctg_offset.push_back(tot); B->ctg_len.push_back((int32_t)contig_len[i]); tot += contig_len[i]; }
  if (tot != n_bases) { delete B; bwaflow::set_error("bwaflow_index_build: contig lengths do not add up to n_bases"); return -22; }
  const int64_t L = n_bases, n = 2 * L;
  B->pac.assign(L / 4 + 1, 0);
  Text T; T.n = n; T.w.assign((n + 31) / 32 + 2, 0);
  srand48(11);                                               // bns->seed, bwa/bntseq.c:289-290
  for (int64_t p = 0; p < L; ++p) {
    int c = codes[p];
    if (c >= 4) c = (int)(lrand48() & 3);
    B->pac[p >> 2] |= (uint8_t)(c << ((~p & 3) << 1));
    T.w[p >> 5] |= (uint64_t)c << ((~p & 31) << 1);
    const int64_t q = n - 1 - p;                             // reverse complement strand
    T.w[q >> 5] |= (uint64_t)(3 - c) << ((~q & 31) << 1);
  }
  if (n < (1LL << 32)) { std::vector<uint32_t> sa; build_sa(T, n_threads, sa); finish(T, sa, B); }
  else { std::vector<uint64_t> sa; build_sa(T, n_threads, sa); finish(T, sa, B); }
  bwaflow_index_t &v = B->view;
  v.l_pac = L; v.pac = B->pac.data(); v.n_seqs = n_contigs; v._pad = 0;
  v.ctg_offset = B->ctg_offset.data(); v.ctg_len = B->ctg_len.data();
  *out = B;
  return 0;
}

const bwaflow_index_t *bwaflow_built_index_view(const bwaflow_built_index_t *b) { return &b->view; }
void bwaflow_built_index_free(bwaflow_built_index_t *b) { delete b; }

} // extern "C"
