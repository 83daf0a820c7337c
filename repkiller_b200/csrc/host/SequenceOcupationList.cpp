#include "SequenceOcupationList.h"

#include <cstdlib>

static int sol_device() {
  const char *e = getenv("RK_DEVICE");
  return e ? atoi(e) : 0;
}

SequenceOcupationList::SequenceOcupationList(double len_ratio, double pos_ratio, uint64_t seq_size)
    : sol_(rk_sol_create(sol_device(), len_ratio, pos_ratio, seq_size)) {
  if (!sol_) throw std::runtime_error(std::string("repkiller-b200: cannot create the device occupation list: ") + rk_create_error());
}

SequenceOcupationList::~SequenceOcupationList() { rk_sol_destroy(sol_); }

FragsGroup *SequenceOcupationList::get_associated_group(uint64_t center, uint64_t length) const {
  uint64_t tag = 0;
  if (rk_sol_get_associated(sol_, center, length, &tag) != RK_OK)
    throw std::runtime_error(std::string("repkiller-b200: ") + rk_sol_last_error(sol_));
  return reinterpret_cast<FragsGroup *>(static_cast<uintptr_t>(tag));
}

void SequenceOcupationList::insert(uint64_t center, uint64_t length, FragsGroup *group) {
  if (rk_sol_insert(sol_, center, length, static_cast<uint64_t>(reinterpret_cast<uintptr_t>(group))) != RK_OK)
    throw std::runtime_error(std::string("repkiller-b200: ") + rk_sol_last_error(sol_));
}
