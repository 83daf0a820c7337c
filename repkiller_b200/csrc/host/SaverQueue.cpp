#include "SaverQueue.h"

#include <iostream>
#include <stdexcept>

void SaverQueue::run() {
  for (;;) {
    SaveRequest sr;
    {
      std::unique_lock<std::mutex> lck(mutex_);
      cond_.wait(lck, [&] { return !queue_.empty() || !running_; });
      if (queue_.empty()) return;  // stopped and drained
      sr = queue_.front();
      queue_.pop();
    }
    try {
      save_all_frag_pairs(sr.path, seq_mngr, *sr.fgl);
    } catch (const std::runtime_error &) {  // reference: SaverQueue.cpp:16-20
      const std::string default_path = "represults-" + std::to_string(++count_) + ".csv";
      std::cerr << "Couldn't access " << sr.path << ", saving into " << default_path << "\n" << std::flush;
      save_all_frag_pairs(default_path, seq_mngr, *sr.fgl);
    }
    free_groups(sr.fgl);
  }
}

SaverQueue::~SaverQueue() {
  if (running_) stop();
}

void SaverQueue::start() {
  std::lock_guard<std::mutex> lck(mutex_);
  if (!running_) {
    running_ = true;
    thread_ptr_.reset(new std::thread(&SaverQueue::run, this));
  }
}

void SaverQueue::stop() {
  {
    std::lock_guard<std::mutex> lck(mutex_);
    if (!running_) return;
    running_ = false;
  }
  cond_.notify_all();
  thread_ptr_->join();
}

void SaverQueue::addRequest(const std::string &path, FGList *fgl) {
  {
    std::lock_guard<std::mutex> lck(mutex_);
    queue_.push(SaveRequest{path, fgl});
  }
  cond_.notify_all();
}
