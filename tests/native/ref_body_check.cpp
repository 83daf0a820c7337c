// TEST INFRASTRUCTURE.  The reference's OWN generate_fragment_groups body, compiled unchanged against this repo's
// drop-in classes: FragmentsDatabase (begin()/end() bucket view from the device's processing order),
// SequenceOcupationList (device-resident lists, rk_sol_*), FragsGroup/FGList, sequence_manager.
//
// The body is not copied into the repository: oracle/Makefile extracts lines 41-80 of
// /root/reference/src/commonFunctions.cpp into the git-ignored oracle/_ref/gfg_body.inc when the reference is present,
// and this file includes it.  Then the repo's generate_diagonal_func, sort_groups (a pure function of the list and the
// table) and writer run on the groups that body produced; the output must be the reference's golden bytes.
//
// usage: ref_body_check <in.csv> <out.csv> <len_ratio> <pos_ratio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "FragmentsDatabase.h"
#include "SequenceOcupationList.h"
#include "commonFunctions.h"

namespace refbody {
using namespace std;
#include "gfg_body.inc"
}  // namespace refbody

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  std::ifstream in(argv[1], std::ifstream::in | std::ifstream::binary);
  if (!in) return 3;
  sequence_manager sm;
  FragmentsDatabase db(in, sm);
  FGList groups;
  refbody::generate_fragment_groups(db, groups, sm, std::stod(argv[3]), std::stod(argv[4]));
  std::vector<size_t> diag(db.getA());
  generate_diagonal_func(db, diag.data());
  sort_groups(groups, diag.data());
  save_all_frag_pairs(argv[2], sm, groups);
  return 0;
}
