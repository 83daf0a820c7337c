// TEST INFRASTRUCTURE — not part of the product path.
//
// Synchronous driver around the UNMODIFIED reference translation units
// (/root/reference/src/{FragmentsDatabase,SequenceOcupationList,commonFunctions,class_structs}.cpp).
// It performs the same calls, in the same order, as the reference's
// execWithParams (src/repkiller.cpp:80-97) and the saver thread
// (src/SaverQueue.cpp:15), but on one thread: the reference's own main cannot
// be rebuilt with g++ 13 because the file-scope `#pragma pack(1)`
// (src/structs.h:2) misaligns SaverQueue's std::mutex.
// No reference source is copied; the sources are compiled where they lie.
//
// usage: repkiller_ref <in.csv> <out.csv> <len_ratio> <pos_ratio>
//   env RK_REF_NOSAVE=1  skip save_all_frag_pairs (timing runs)
//   env RK_REF_REPEAT=K  repeat the timed calls K times on the once-loaded database (one JSON line each)
// prints one JSON line with per-phase milliseconds on stderr.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "commonFunctions.h"

static double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <in.csv> <out.csv> <len_ratio> <pos_ratio>\n", argv[0]);
    return 1;
  }
  const std::string out_path = argv[2];
  const double len_ratio = std::stod(argv[3]);
  const double pos_ratio = std::stod(argv[4]);
  ifstream in(argv[1], ifstream::in | ifstream::binary);
  if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }

  sequence_manager sm;
  auto t0 = std::chrono::steady_clock::now();
  FragmentsDatabase db(in, sm);
  in.close();
  const double load_ms = ms_since(t0);

  // RK_REF_REPEAT=K (bench.py --impl reference): the database is loaded once and the timed calls are repeated K
  // times, one JSON line per repetition; RK_REF_BUDGET_S stops the loop early once that much wall time is spent.
  const int repeat = getenv("RK_REF_REPEAT") ? atoi(getenv("RK_REF_REPEAT")) : 1;
  const double budget_ms = getenv("RK_REF_BUDGET_S") ? 1e3 * atof(getenv("RK_REF_BUDGET_S")) : 0.0;
  const auto t_start = std::chrono::steady_clock::now();
  for (int rep = 0; rep < repeat; ++rep) {
    FGList *groups = new FGList;
    t0 = std::chrono::steady_clock::now();
    generate_fragment_groups(db, *groups, sm, len_ratio, pos_ratio);
    const double group_ms = ms_since(t0);

    t0 = std::chrono::steady_clock::now();
    size_t *diag_func = new size_t[db.getA()];
    generate_diagonal_func(db, diag_func);
    const double diag_ms = ms_since(t0);

    t0 = std::chrono::steady_clock::now();
    sort_groups(*groups, diag_func);
    const double sort_ms = ms_since(t0);
    delete[] diag_func;

    double save_ms = 0.0;
    if (!getenv("RK_REF_NOSAVE") && rep == repeat - 1) {
      t0 = std::chrono::steady_clock::now();
      save_all_frag_pairs(out_path, sm, *groups);
      save_ms = ms_since(t0);
    }
    fprintf(stderr,
            "{\"n_frags\": %llu, \"n_groups\": %zu, \"load_ms\": %.3f, \"group_ms\": %.3f, "
            "\"diag_ms\": %.3f, \"sort_ms\": %.3f, \"save_ms\": %.3f}\n",
            (unsigned long long)db.getTotalFrags(), groups->size(), load_ms, group_ms, diag_ms,
            sort_ms, save_ms);
    fflush(stderr);
    for (FragsGroup *g : *groups) delete g;  // (the reference's saver leaks them; a repeat loop cannot)
    delete groups;
    if (budget_ms > 0 && ms_since(t_start) > budget_ms) break;
  }
  return 0;
}
