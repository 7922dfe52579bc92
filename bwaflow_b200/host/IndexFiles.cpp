#include "IndexFiles.h"
#include <stdio.h>
#include <string.h>
#include <map>
#include <stdexcept>

namespace {
struct File {
  FILE *f;
  File(const std::string &p, const char *mode, bool must = true) : f(fopen(p.c_str(), mode))
  { if (!f && must) throw std::runtime_error("cannot open " + p); }
  ~File() { if (f) fclose(f); }
};
template <class T> void rd(FILE *f, T *p, size_t n, const std::string &what)
{
  // large reads in slices, as fread_fix does (bwa/bwt.c:409-419)
  char *c = (char *)p; size_t left = n * sizeof(T);
  while (left) {
    const size_t k = left < (size_t)0x1000000 ? left : (size_t)0x1000000;
    if (fread(c, 1, k, f) != k) throw std::runtime_error("short read in " + what);
    c += k; left -= k;
  }
}
}

void IndexFiles::load(const std::string &prefix)
{
  memset(&ix, 0, sizeof(ix));
  { // .bwt: primary, L2[1..4], occ-interleaved BWT words
    File fp(prefix + ".bwt", "rb");
    fseek(fp.f, 0, SEEK_END);
    const long sz = ftell(fp.f);
    if (sz < 40) throw std::runtime_error(prefix + ".bwt is truncated");
    ix.bwt_size = (uint64_t)(sz - 40) >> 2;
    fseek(fp.f, 0, SEEK_SET);
    rd(fp.f, &ix.primary, 1, ".bwt");
    ix.L2[0] = 0;
    rd(fp.f, ix.L2 + 1, 4, ".bwt");
    bwt.resize(ix.bwt_size);
    rd(fp.f, bwt.data(), bwt.size(), ".bwt");
    ix.seq_len = ix.L2[4];
    ix.bwt = bwt.data();
  }
  { // .sa: primary, 4 skipped words, sa_intv, seq_len, sa[1..n_sa)
    File fp(prefix + ".sa", "rb");
    uint64_t primary, skip[4], intv, seq_len;
    rd(fp.f, &primary, 1, ".sa"); rd(fp.f, skip, 4, ".sa"); rd(fp.f, &intv, 1, ".sa"); rd(fp.f, &seq_len, 1, ".sa");
    if (primary != ix.primary || seq_len != ix.seq_len || intv == 0) throw std::runtime_error("SA-BWT inconsistency in " + prefix);
    ix.sa_intv = (int64_t)intv;
    ix.n_sa = (ix.seq_len + intv) / intv;
    sa.assign(ix.n_sa, 0);
    sa[0] = (uint64_t)-1;
    rd(fp.f, sa.data() + 1, sa.size() - 1, ".sa");
    ix.sa = sa.data();
  }
  { // .ann: l_pac n_seqs seed / per contig: "gi name anno" and "offset len n_ambs"
    File fp(prefix + ".ann", "r");
    long long xx; int n_seqs; unsigned seed;
    if (fscanf(fp.f, "%lld%d%u", &xx, &n_seqs, &seed) != 3) throw std::runtime_error("parse error in " + prefix + ".ann");
    ix.l_pac = xx; ix.n_seqs = n_seqs;
    char str[8192];
    for (int i = 0; i < n_seqs; ++i) {
      unsigned gi; int len, n_ambs, c = 0;
      if (fscanf(fp.f, "%u%8191s", &gi, str) != 2) throw std::runtime_error("parse error in " + prefix + ".ann");
      names.push_back(str);
      char *q = str;
      while (q - str < (long)sizeof(str) - 1 && (c = fgetc(fp.f)) != '\n' && c != EOF) *q++ = (char)c;
      while (c != '\n' && c != EOF) c = fgetc(fp.f);
      *q = 0;
      anno.push_back(q - str > 1 && strcmp(str, " (null)") != 0 ? std::string(str + 1) : std::string());
      if (fscanf(fp.f, "%lld%d%d", &xx, &len, &n_ambs) != 3) throw std::runtime_error("parse error in " + prefix + ".ann");
      ctg_offset.push_back(xx); ctg_len.push_back(len);
    }
    is_alt.assign((size_t)n_seqs, 0);
  }
  { // .alt (optional): first column = names of ALT contigs, '@' lines skipped
    File fp(prefix + ".alt", "r", false);
    if (fp.f) {
      std::map<std::string, int> by_name;
      for (size_t i = 0; i < names.size(); ++i) by_name[names[i]] = (int)i;
      std::string s; int c;
      while ((c = fgetc(fp.f)) != EOF) {
        if (c == '\t' || c == '\n' || c == '\r') {
          if (!s.empty() && s[0] != '@') { auto it = by_name.find(s); if (it != by_name.end()) is_alt[(size_t)it->second] = 1; }
          while (c != '\n' && c != EOF) c = fgetc(fp.f);
          s.clear();
        } else s.push_back((char)c);
      }
    }
  }
  { // .pac: 2-bit forward strand, 4 bases per byte MSB first
    File fp(prefix + ".pac", "rb");
    pac.assign((size_t)(ix.l_pac / 4 + 2), 0);
    const size_t n = (size_t)(ix.l_pac / 4 + (ix.l_pac % 4 ? 1 : 0));
    rd(fp.f, pac.data(), n, ".pac");
    ix.pac = pac.data();
  }
  for (size_t i = 0; i < names.size(); ++i) { name_ptr.push_back(names[i].c_str()); anno_ptr.push_back(anno[i].c_str()); is_alt_u8.push_back((uint8_t)is_alt[i]); }
  ix.ctg_offset = ctg_offset.data(); ix.ctg_len = ctg_len.data();
}
