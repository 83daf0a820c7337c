// Drop-in for the reference's FragmentsDatabase (/root/reference/src/FragmentsDatabase.h:15-35): same
// constructor and accessors, but the per-bucket vector<FragFile>[] is replaced by the records in file order
// (pinned host memory) plus the device-side processing order built by rk_load_aos.
#pragma once

#include <cstdint>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "rk_b200.h"
#include "structs.h"

bool readFragment(FragFile *frag, const char *line, size_t len);  // reference: FragmentsDatabase.cpp:17-50
// the strtof fast path of readFragment (digits[.digits]): true and the correctly rounded float, or false = ask strtof
bool parse_plain_float(const char *p, size_t n, float *out);
// data rows of text[pos, size) parsed by nthreads threads; the accepted records of consecutive ranges, in file order
std::vector<std::vector<FragFile>> parse_rows_parallel(const char *text, size_t size, size_t pos, unsigned nthreads);

// what the stream holds: the GECKO CSV the reference reads, or GECKO's binary container (GeckoFrags.h)
enum class FragsInput { csv, gecko_binary };
// What the CLI uses to choose: RK_INPUT_FORMAT=csv|frags decides; otherwise a file is taken as binary only when its name
// ends in ".frags", its size is a header plus whole records and it starts with a zero byte (a sequence length below 2^56;
// the CSV starts with text).  Everything else is read as the reference reads it.
FragsInput detect_frags_input(const std::string &path, std::ifstream &stream);
// The input stage of the constructor (no device involved).  Head: the header text the output file will start with, the
// sequence lengths as the reference holds them (file value + 1, FragmentsDatabase.cpp:62,65), the record count the file
// states, where the records start.  Records: the accepted records of consecutive ranges of the file, in file order,
// read by nthreads threads (0: all cores / RK_PARSE_THREADS).
struct FragsHead {
  std::string header;
  uint64_t seqx_len = 0, seqy_len = 0, total_frags = 0;
  size_t body_pos = 0;
  bool binary = false;
};
FragsHead read_frags_head(const char *data, size_t size, FragsInput input);
std::vector<std::vector<FragFile>> read_frags_records(const char *data, size_t size, const FragsHead &head, unsigned nthreads);

class FragmentsDatabase {
  FragFile *records_ = nullptr;  // file order
  unsigned char *packed_ = nullptr;  // pinned: the compact arrays handed to rk_load_packed
  uint64_t count_ = 0, cap_ = 0;
  size_t vsize = 0;
  rk_ctx *ctx_ = nullptr;
  rk_multi *multi_ = nullptr;  // several GPUs: one comparison partitioned over them (rk_create_multi)
  std::string header;
  rk_load_stats load_stats_{};
  double ms_read_ = 0, ms_parse_ = 0, ms_device_load_ = 0;  // host wall clock of the three ingest phases
  // the reference's bucket array, built on first use from the processing order the device computed (K2a)
  mutable std::unique_ptr<std::vector<FragFile>[]> buckets_;
  mutable std::once_flag buckets_once_;
  void build_buckets() const;

 public:
  // Parses the GECKO CSV exactly like the reference (16 header lines, Frag rows, readFragment's accept/pad
  // rules), fills seq_manager, then hands the records to the GPU.  Throws std::runtime_error like the
  // reference on "Unexpected number of fragments"; also on device errors (message from rk_last_error).
  FragmentsDatabase(std::ifstream &frags_file, sequence_manager &seq_manager, int device = 0, FragsInput input = FragsInput::csv)
      : FragmentsDatabase(frags_file, seq_manager, std::vector<int>{device}, input) {}
  // several devices: the comparison is partitioned over them (rk_multi_*); ctx() is then null and multi() is set
  // input = gecko_binary: the stream is a .frags file; every record loads as readFragment would load its CSV row
  FragmentsDatabase(std::ifstream &frags_file, sequence_manager &seq_manager, const std::vector<int> &devices,
                    FragsInput input = FragsInput::csv);
  ~FragmentsDatabase();
  FragmentsDatabase(const FragmentsDatabase &) = delete;
  FragmentsDatabase &operator=(const FragmentsDatabase &) = delete;

  size_t getA() const { return vsize; }                 // reference: FragmentsDatabase.h:23-25
  uint64_t getTotalFrags() const { return count_; }      // reference: FragmentsDatabase.h:32-34
  // reference: FragmentsDatabase.h:26-31 — the xStart/10 bucket array; end() is begin() + getA() - 1, i.e. the last
  // bucket is never visited.  Built lazily (a host copy of every record, like the reference holds) from the device's
  // processing order; the grouping functions of this repo do not need it.
  const std::vector<FragFile> *begin() const;
  const std::vector<FragFile> *end() const { return begin() + vsize - 1; }
  const FragFile *records() const { return records_; }   // the records in file order
  rk_ctx *ctx() const { return ctx_; }
  rk_multi *multi() const { return multi_; }
  const rk_load_stats &load_stats() const { return load_stats_; }
  double ms_read() const { return ms_read_; }
  double ms_parse() const { return ms_parse_; }
  double ms_device_load() const { return ms_device_load_; }
};
