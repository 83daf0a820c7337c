#include "SaverQueue.h"

#include <iostream>
#include <stdexcept>
#include <utility>

void SaverQueue::write_job(const Job &job) {
  try {
    save_all_frag_pairs(job.out_path, sequences_, *job.groups);
  } catch (const std::runtime_error &) {  // reference: SaverQueue.cpp:16-20
    const std::string fallback = "represults-" + std::to_string(++fallback_files_) + ".csv";
    std::cerr << "Couldn't access " << job.out_path << ", saving into " << fallback << "\n" << std::flush;
    save_all_frag_pairs(fallback, sequences_, *job.groups);
  }
  free_groups(job.groups);
}

void SaverQueue::work() {
  std::unique_lock<std::mutex> hold(lock_);
  for (;;) {
    wake_.wait(hold, [this] { return !fifo_.empty() || phase_ == Phase::draining; });
    if (fifo_.empty()) return;  // draining and nothing left
    Job job = std::move(fifo_.front());
    fifo_.pop_front();
    hold.unlock();
    write_job(job);
    hold.lock();
  }
}

void SaverQueue::start() {
  std::lock_guard<std::mutex> hold(lock_);
  if (phase_ != Phase::idle) return;
  phase_ = Phase::accepting;
  worker_ = std::thread(&SaverQueue::work, this);
}

void SaverQueue::stop() {
  {
    std::lock_guard<std::mutex> hold(lock_);
    if (phase_ != Phase::accepting) return;
    phase_ = Phase::draining;
  }
  wake_.notify_all();
  worker_.join();
  std::lock_guard<std::mutex> hold(lock_);
  phase_ = Phase::idle;
}

void SaverQueue::addRequest(const std::string &path, FGList *fgl) {
  {
    std::lock_guard<std::mutex> hold(lock_);
    fifo_.push_back(Job{path, fgl});
  }
  wake_.notify_one();
}
