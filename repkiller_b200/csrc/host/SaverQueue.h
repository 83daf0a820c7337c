// Host-side consumer of device results, same interface as /root/reference/src/SaverQueue.h:16-42.  The two
// defects of the reference are not reproduced: front() on an empty queue when woken by stop()
// (SaverQueue.cpp:7-12) and the leak of every FragsGroup (SaverQueue.cpp:21).
#pragma once

#include <condition_variable>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>

#include "commonFunctions.h"
#include "structs.h"

class SaverQueue {
  struct SaveRequest {
    std::string path;
    FGList *fgl;
  };
  size_t count_ = 0;
  bool running_ = false;
  const sequence_manager &seq_mngr;
  std::mutex mutex_;
  std::condition_variable cond_;
  std::queue<SaveRequest> queue_;
  std::unique_ptr<std::thread> thread_ptr_;
  void run();

 public:
  explicit SaverQueue(const sequence_manager &seq_mngr) : seq_mngr(seq_mngr) {}
  ~SaverQueue();
  void start();
  void stop();  // drains the queue, then joins
  void addRequest(const std::string &path, FGList *fgl);  // takes ownership of fgl and its groups
};
