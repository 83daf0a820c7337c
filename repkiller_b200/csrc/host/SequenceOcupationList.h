// Drop-in for /root/reference/src/SequenceOcupationList.h:13-33: same constructor, same two methods, same results.
// The occupation lists live in device memory behind the C ABI (rk_sol_*, csrc/sol.cu); there is no host copy and no
// CPU evaluation.  Code that drives the lists call by call — the reference's own generate_fragment_groups body
// (src/commonFunctions.cpp:41-80) — compiles against this class unchanged; whole databases should call
// generate_fragment_groups of this repo (rk_group), which runs all queries of an axis in parallel.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>

#include "rk_b200.h"
#include "structs.h"

#define DIVISOR 100  // reference: SequenceOcupationList.h:11

class SequenceOcupationList {
  rk_sol *sol_;

 public:
  // reference: SequenceOcupationList.cpp:3-8.  RK_DEVICE selects the GPU (default 0).
  SequenceOcupationList(double len_pos_ratio, double pos_ratio, uint64_t max_length);
  ~SequenceOcupationList();
  SequenceOcupationList(const SequenceOcupationList &) = delete;
  SequenceOcupationList &operator=(const SequenceOcupationList &) = delete;
  // reference: SequenceOcupationList.cpp:33-91
  FragsGroup *get_associated_group(uint64_t center, uint64_t length) const;
  // reference: SequenceOcupationList.cpp:93-96
  void insert(uint64_t center, uint64_t length, FragsGroup *group);
};
