// Host-side consumer of grouping results with the interface the reference's callers use
// (/root/reference/src/SaverQueue.h:16-42: construct with the sequence manager, start(), addRequest(path, list),
// stop()).  One worker thread drains a FIFO of jobs and writes each list with save_all_frag_pairs; an unwritable
// path falls back to represults-<n>.csv like the reference (SaverQueue.cpp:16-20).  Two defects of the reference are
// not reproduced: front() on an empty queue when the worker is woken by stop() (SaverQueue.cpp:7-12) and the leak of
// every FragsGroup (SaverQueue.cpp:21).
#pragma once

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>

#include "commonFunctions.h"
#include "structs.h"

class SaverQueue {
 public:
  explicit SaverQueue(const sequence_manager &sequences) : sequences_(sequences) {}
  ~SaverQueue() { stop(); }
  SaverQueue(const SaverQueue &) = delete;
  SaverQueue &operator=(const SaverQueue &) = delete;

  void start();                                           // idempotent
  void stop();                                            // drains the FIFO, then joins; idempotent
  void addRequest(const std::string &path, FGList *fgl);  // takes ownership of fgl and of its groups

 private:
  struct Job {
    std::string out_path;
    FGList *groups;
  };
  enum class Phase { idle, accepting, draining };

  void work();
  void write_job(const Job &job);

  const sequence_manager &sequences_;
  std::mutex lock_;
  std::condition_variable wake_;
  std::deque<Job> fifo_;
  std::thread worker_;
  Phase phase_ = Phase::idle;
  size_t fallback_files_ = 0;
};
