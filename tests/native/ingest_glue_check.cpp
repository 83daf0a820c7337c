// The FragmentsDatabase constructor against a RECORDING stand-in for librk_b200 (test only: nothing is computed; the
// stand-in notes what the constructor hands to the library).  Checks the host glue without a GPU: the arguments of the load
// call (count, sequence lengths, the compact arrays) must be the same whether the input was the CSV or the .frags file,
// and must equal what the test computes from the records.  Build: g++ this file + host/FragmentsDatabase.cpp +
// host/GeckoFrags.cpp (no librk_b200).
//   ingest_glue_check <input> [devices]     prints one line per fact
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "../../repkiller_b200/csrc/host/FragmentsDatabase.h"

namespace {
uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
int g_live_ctx = 0, g_live_multi = 0, g_live_host = 0;
}  // namespace

struct rk_ctx { int device; };
struct rk_multi { int n; };

extern "C" {
rk_ctx *rk_create(int device) { ++g_live_ctx; return new rk_ctx{device}; }
void rk_destroy(rk_ctx *c) { --g_live_ctx; delete c; }
rk_multi *rk_create_multi(const int *, int n) { ++g_live_multi; return new rk_multi{n}; }
void rk_destroy_multi(rk_multi *m) { --g_live_multi; delete m; }
const char *rk_create_error(void) { return ""; }
const char *rk_last_error(const rk_ctx *) { return ""; }
const char *rk_multi_last_error(const rk_multi *) { return ""; }
void *rk_host_alloc(size_t bytes) { ++g_live_host; return malloc(bytes); }
void rk_host_free(void *p) { --g_live_host; free(p); }
int64_t rk_debug_fetch(rk_ctx *, const char *, void *, uint64_t) { return -1; }
int rk_load_packed(rk_ctx *, const uint32_t *key4, const uint8_t *strand, const uint32_t *rest4, uint64_t n, uint64_t lx, uint64_t ly,
                   unsigned, rk_load_stats *st) {
  printf("call rk_load_packed n=%llu seqx_len=%llu seqy_len=%llu key4=%016llx strand=%016llx rest4=%016llx\n", (unsigned long long)n,
         (unsigned long long)lx, (unsigned long long)ly, (unsigned long long)fnv(key4, n * 16), (unsigned long long)fnv(strand, n),
         (unsigned long long)fnv(rest4, n * 16));
  memset(st, 0, sizeof *st);
  st->n_loaded = n, st->n_kept = n;
  return 0;
}
int rk_load_aos(rk_ctx *, const void *frags, uint64_t n, uint64_t lx, uint64_t ly, unsigned, rk_load_stats *st) {
  printf("call rk_load_aos n=%llu seqx_len=%llu seqy_len=%llu records=%016llx\n", (unsigned long long)n, (unsigned long long)lx,
         (unsigned long long)ly, (unsigned long long)fnv(frags, n * 109));
  memset(st, 0, sizeof *st);
  st->n_loaded = n, st->n_kept = n;
  return 0;
}
int rk_multi_load_aos(rk_multi *m, const void *frags, uint64_t n, uint64_t lx, uint64_t ly, unsigned, rk_load_stats *st) {
  printf("call rk_multi_load_aos ranks=%d n=%llu seqx_len=%llu seqy_len=%llu records=%016llx\n", m->n, (unsigned long long)n,
         (unsigned long long)lx, (unsigned long long)ly, (unsigned long long)fnv(frags, n * 109));
  memset(st, 0, sizeof *st);
  st->n_loaded = n, st->n_kept = n;
  return 0;
}
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::vector<int> devices(argc >= 3 ? atoi(argv[2]) : 1);
  for (size_t i = 0; i < devices.size(); ++i) devices[i] = (int)i;
  try {
    std::ifstream in(argv[1], std::ifstream::in | std::ifstream::binary);
    sequence_manager sm;
    {
      FragmentsDatabase db(in, sm, devices, detect_frags_input(argv[1], in));
      printf("getA=%zu total=%llu seqs=%llu len0=%llu len1=%llu max=%llu\n", db.getA(), (unsigned long long)db.getTotalFrags(),
             (unsigned long long)sm.get_number_of_sequences(), (unsigned long long)sm.get_sequence_by_label(0).len,
             (unsigned long long)sm.get_sequence_by_label(1).len, (unsigned long long)sm.get_maximum_length());
      printf("records=%016llx\n", (unsigned long long)fnv(db.records(), db.getTotalFrags() * sizeof(FragFile)));
      printf("header_lines=%zu header=%016llx\n", (size_t)std::count(sm.raw_header().begin(), sm.raw_header().end(), '\n'),
             (unsigned long long)fnv(sm.raw_header().data(), sm.raw_header().size()));
    }
    printf("live ctx=%d multi=%d host=%d\n", g_live_ctx, g_live_multi, g_live_host);
  } catch (const std::exception &e) {
    printf("exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
