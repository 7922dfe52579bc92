#include "Ingest.h"
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

FastxReader::FastxReader(const char *path) : begin_(0), end_(0), eof_(0), last_char_(0)
{
  f_ = strcmp(path, "-") ? gzopen(path, "r") : gzdopen(fileno(stdin), "r");
  if (!f_) throw std::runtime_error(std::string("cannot open ") + path);
}
FastxReader::~FastxReader() { if (f_) gzclose(f_); }

int FastxReader::getc_()
{
  if (begin_ >= end_) {
    if (eof_) return -1;
    begin_ = 0;
    end_ = gzread(f_, buf_, sizeof(buf_));
    if (end_ <= 0) { eof_ = 1; end_ = 0; return -1; }
  }
  return buf_[begin_++];
}

// up to the end of the line (the newline is consumed, a trailing '\r' dropped); whole buffer stretches at a time
void FastxReader::rest_of_line(std::string &s, bool append)
{
  if (!append) s.clear();
  const size_t l0 = s.size();
  for (;;) {
    if (begin_ >= end_) {
      if (eof_) break;
      begin_ = 0;
      end_ = gzread(f_, buf_, sizeof(buf_));
      if (end_ <= 0) { eof_ = 1; end_ = 0; break; }
    }
    const unsigned char *p = buf_ + begin_;
    const unsigned char *nl = (const unsigned char *)memchr(p, '\n', (size_t)(end_ - begin_));
    if (nl) { s.append((const char *)p, (size_t)(nl - p)); begin_ = (int)(nl - buf_) + 1; break; }
    s.append((const char *)p, (size_t)(end_ - begin_));
    begin_ = end_;
  }
  if (s.size() > l0 && s.size() > 1 && s.back() == '\r') s.pop_back();
}

int FastxReader::read(std::string &name, std::string &comment, std::string &seq, std::string &qual)
{
  int c;
  if (last_char_ == 0) {
    while ((c = getc_()) != -1 && c != '>' && c != '@') {}
    if (c == -1) return -1;
    last_char_ = c;
  }
  comment.clear(); seq.clear(); qual.clear(); name.clear();
  while ((c = getc_()) != -1 && !isspace(c)) name.push_back((char)c);
  if (c == -1 && name.empty()) return -1;
  if (c != '\n' && c != -1) rest_of_line(comment, false);
  while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
    if (c == '\n') continue;
    seq.push_back((char)c);
    rest_of_line(seq, true);
  }
  if (c == '>' || c == '@') last_char_ = c;
  if (c != '+') return (int)seq.size();
  while ((c = getc_()) != -1 && c != '\n') {}
  if (c == -1) return -2;
  for (;;) {
    if (begin_ >= end_ && eof_) break;
    const size_t l0 = qual.size();
    rest_of_line(qual, true);
    if (qual.size() == l0 && begin_ >= end_ && eof_) break;
    if (qual.size() >= seq.size()) break;
  }
  last_char_ = 0;
  if (seq.size() != qual.size()) return -2;
  return (int)seq.size();
}

// trim_readno, bwa/bwa.c:27-31
static void trim_readno(std::string &s)
{
  if (s.size() > 2 && s[s.size() - 2] == '/' && isdigit((unsigned char)s[s.size() - 1])) s.resize(s.size() - 2);
}

static bseq1_t to_bseq(std::string &name, const std::string &comment, const std::string &seq, const std::string &qual, bool copy_comment)
{
  bseq1_t b;
  memset(&b, 0, sizeof(b));
  trim_readno(name);
  b.name = strdup(name.c_str());
  b.comment = copy_comment && !comment.empty() ? strdup(comment.c_str()) : 0;
  b.seq = strdup(seq.c_str());
  b.qual = !qual.empty() ? strdup(qual.c_str()) : 0;
  b.l_seq = (int)strlen(b.seq);
  return b;
}

// getKseqBatch + KseqsRead::compute + KseqsToBseqs::compute, src/Pipeline.cpp:98-202
void FastqSource::compute()
{
  std::string name, comment, seq, qual;
  for (;;) {
    std::vector<bseq1_t> v;
    int size = 0;
    for (;;) {
      if (r1_.read(name, comment, seq, qual) < 0) break;
      v.push_back(to_bseq(name, comment, seq, qual, copy_comment_));
      size += v.back().l_seq;
      if (r2_) {
        if (r2_->read(name, comment, seq, qual) < 0) { fprintf(stderr, "[W::%s] the 2nd file has fewer sequences.\n", __func__); break; }
        v.push_back(to_bseq(name, comment, seq, qual, copy_comment_));
        size += v.back().l_seq;
      }
      if (size >= chunk_size_ && (v.size() & 1) == 0) break;
    }
    if (size == 0 && r2_ && r2_->read(name, comment, seq, qual) >= 0) fprintf(stderr, "[W::%s] the 1st file has fewer sequences.\n", __func__);
    if (v.empty()) break;
    SeqsRecord rec;
    rec.start_idx = produced_;
    rec.batch_num = (int)v.size();
    rec.seqs = (bseq1_t *)calloc(v.size(), sizeof(bseq1_t));
    memcpy(rec.seqs, v.data(), v.size() * sizeof(bseq1_t));
    pushOutput(rec);
    produced_ += v.size();
  }
}

void SamsReorder::compute(int)
{
  uint64_t n_processed = 0;
  std::map<uint64_t, SeqsRecord> held;
  for (;;) {
    SeqsRecord in = SeqsRecord();
    bool ready = this->getInput(in);
    while (!this->isFinal() && !ready) {
      std::this_thread::sleep_for(std::chrono::microseconds(10));
      ready = this->getInput(in);
    }
    if (!ready && !(ready = this->getInput(in))) break;
    if (!inorder_) { this->pushOutput(in); continue; }
    held[in.start_idx] = in;
    for (auto it = held.find(n_processed); it != held.end(); it = held.find(n_processed)) {
      SeqsRecord r = it->second;
      held.erase(it);
      n_processed += (uint64_t)r.batch_num;
      this->pushOutput(r);
    }
  }
  if (!held.empty()) throw std::runtime_error("SamsReorder: records missing from the sequence");
}

void SamsPrint::compute()
{
  for (;;) {
    SeqsRecord r = SeqsRecord();
    bool ready = getInput(r);
    if (!ready) {
      if (isFinal()) { if (!(ready = getInput(r))) break; }
      else { std::this_thread::sleep_for(std::chrono::microseconds(10)); continue; }
    }
    for (int i = 0; i < r.batch_num; ++i) {
      if (r.seqs[i].sam) {
        const size_t l = strlen(r.seqs[i].sam);
        if (fwrite(r.seqs[i].sam, 1, l, out_) != l) throw std::runtime_error("SamsPrint: write failed");
        n_bytes += l;
        free(r.seqs[i].sam);
      }
      ++n_reads;
    }
    free(r.seqs);
  }
  fflush(out_);
}
