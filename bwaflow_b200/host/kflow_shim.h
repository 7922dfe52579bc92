// A boost-free stand-in for the slice of kflow (namespace kestrelFlow) that accelerator stages are written against:
// Queue (kflow/include/kflow/Queue.h:16-70), Stage/MapStage/MapPartitionStage/SourceStage/SinkStage
// (kflow/include/kflow/{Stage,MapStage,MapPartitionStage,SourceStage,SinkStage}.h) and Pipeline::addStage/start/wait
// (kflow/include/kflow/Pipeline.h:98-183).  Same class names, template parameters and protected member functions
// (compute, getInput, pushOutput, isFinal), so that src/cuda/CUDAPipeline.{h,cpp} compile unchanged against the real kflow
// headers where boost and glog exist.  Scheduling is the static-stage model only: n_workers dedicated threads per stage
// (kflow/src/Stage.cpp:46-60); the dynamic worker pool and the accelerator token counters are not reproduced.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace kestrelFlow {

class QueueBase {
 public:
  virtual ~QueueBase() {}
};

template <typename U, int DEPTH = 64>
class Queue : public QueueBase {
 public:
  void push(U const &item)
  { // blocking when full (Queue.h:38-48)
    std::unique_lock<std::mutex> lk(m_);
    not_full_.wait(lk, [&] { return (int)q_.size() < DEPTH; });
    q_.push_back(item);
  }
  bool async_pop(U &item)
  { // non-blocking (Queue.h:50-60)
    std::lock_guard<std::mutex> lk(m_);
    if (q_.empty()) return false;
    item = q_.front();
    q_.pop_front();
    not_full_.notify_one();
    return true;
  }
  bool empty()
  {
    std::lock_guard<std::mutex> lk(m_);
    return q_.empty();
  }

 private:
  std::mutex m_;
  std::condition_variable not_full_;
  std::deque<U> q_;
};

class StageBase {
 public:
  StageBase(int n_workers) : n_workers_(n_workers), prev_(nullptr), done_(false), live_(0) {}
  virtual ~StageBase() {}
  virtual QueueBase *outputQueue() = 0;
  virtual void setInputQueue(QueueBase *q) = 0;
  virtual void worker(int wid) = 0;
  void start()
  {
    live_ = n_workers_;
    for (int i = 0; i < n_workers_; ++i)
      threads_.emplace_back([this, i] {
        try { this->worker(i); } catch (std::exception &e) { error_ = e.what(); }
        if (--live_ == 0) done_ = true;
      });
  }
  void wait() { for (auto &t : threads_) t.join(); threads_.clear(); }
  bool finished() const { return done_; }
  const std::string &error() const { return error_; }
  void setPrev(StageBase *p) { prev_ = p; }

 protected:
  int n_workers_;
  StageBase *prev_;
  std::atomic<bool> done_;
  std::atomic<int> live_;
  std::vector<std::thread> threads_;
  std::string error_;
};

template <typename U, typename V, int IN_DEPTH = 64, int OUT_DEPTH = 64>
class Stage : public StageBase {
 public:
  Stage(int n_workers, bool /*is_dyn*/) : StageBase(n_workers), in_(nullptr) {}
  QueueBase *outputQueue() { return &out_; }
  void setInputQueue(QueueBase *q) { in_ = static_cast<Queue<U, IN_DEPTH> *>(q); }

 protected:
  // no more input will ever arrive (Stage.h isFinal(): the upstream stage has finished and the input queue is drained)
  bool isFinal() { return (!prev_ || prev_->finished()) && (!in_ || in_->empty()); }
  Queue<U, IN_DEPTH> *getInputQueue() { return in_; }
  Queue<V, OUT_DEPTH> *getOutputQueue() { return &out_; }
  Queue<U, IN_DEPTH> *in_;
  Queue<V, OUT_DEPTH> out_;
};

// compute(wid) owns the loop (MapPartitionStage.h:19-41); used by the accelerator stages
template <typename U, typename V, int IN_DEPTH = 64, int OUT_DEPTH = 64>
class MapPartitionStage : public Stage<U, V, IN_DEPTH, OUT_DEPTH> {
 public:
  MapPartitionStage(int n_workers = 1, bool is_dyn = true) : Stage<U, V, IN_DEPTH, OUT_DEPTH>(n_workers, is_dyn) {}
  void worker(int wid) { compute(wid); }

 protected:
  virtual void compute(int wid) = 0;
  bool getInput(U &item) { return this->in_ ? this->in_->async_pop(item) : false; }   // MapPartitionStage.h:48-56
  void pushOutput(V const &item) { this->out_.push(item); }                             // :58-63
};

// one record in, one record out (MapStage.h:33)
template <typename U, typename V, int IN_DEPTH = 64, int OUT_DEPTH = 64>
class MapStage : public Stage<U, V, IN_DEPTH, OUT_DEPTH> {
 public:
  MapStage(int n_workers = 1, bool is_dyn = true) : Stage<U, V, IN_DEPTH, OUT_DEPTH>(n_workers, is_dyn) {}
  void worker(int)
  {
    U item;
    for (;;) {
      if (this->in_->async_pop(item)) { this->out_.push(compute(item)); continue; }
      if (this->isFinal()) { if (this->in_->async_pop(item)) { this->out_.push(compute(item)); continue; } break; }
      std::this_thread::sleep_for(std::chrono::microseconds(10));
    }
  }

 protected:
  virtual V compute(U const &input) = 0;
};

template <typename V, int OUT_DEPTH = 64>
class SourceStage : public Stage<int, V, 1, OUT_DEPTH> {
 public:
  SourceStage() : Stage<int, V, 1, OUT_DEPTH>(1, false) {}
  void worker(int) { compute(); }

 protected:
  virtual void compute() = 0;
  void pushOutput(V const &item) { this->out_.push(item); }
};

template <typename U, int IN_DEPTH = 64>
class SinkStage : public Stage<U, int, IN_DEPTH, 1> {
 public:
  SinkStage() : Stage<U, int, IN_DEPTH, 1>(1, false) {}
  void worker(int) { compute(); }

 protected:
  virtual void compute() = 0;
  bool getInput(U &item) { return this->in_->async_pop(item); }
};

// linear pipeline: stage i's output queue is stage i+1's input queue (Pipeline.h:98-148)
class Pipeline {
 public:
  Pipeline(int n_stages, int /*n_threads*/) : stages_(n_stages, nullptr) {}
  bool addStage(int idx, StageBase *s)
  {
    if (idx < 0 || idx >= (int)stages_.size() || stages_[idx]) return false;
    stages_[idx] = s;
    return true;
  }
  void start()
  {
    for (size_t i = 0; i < stages_.size(); ++i) {
      if (!stages_[i]) throw std::runtime_error("pipeline stage missing");
      if (i) { stages_[i]->setInputQueue(stages_[i - 1]->outputQueue()); stages_[i]->setPrev(stages_[i - 1]); }
    }
    for (auto *s : stages_) s->start();
  }
  void wait()
  {
    for (auto *s : stages_) s->wait();
    for (auto *s : stages_) if (!s->error().empty()) throw std::runtime_error(s->error());
  }

 private:
  std::vector<StageBase *> stages_;
};

} // namespace kestrelFlow
