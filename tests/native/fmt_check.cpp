// CPU check of repkiller_b200/csrc/rk_fmt.cuh against printf("%g") (what the reference's ostream << float prints).
// usage: fmt_check <random patterns> ; exits 0 when every value agrees
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <initializer_list>
#include "../../repkiller_b200/csrc/rk_fmt.cuh"

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint64_t next() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }

static long bad = 0, total = 0;
static void check(uint32_t bits) {
  float f;
  memcpy(&f, &bits, 4);
  char want[64], got[64];
  snprintf(want, sizeof want, "%g", (double)f);
  rkfmt::BufSink s(got);
  rkfmt::put_g6(s, bits);
  got[s.n] = 0;
  rkfmt::CountSink c;
  rkfmt::put_g6(c, bits);
  ++total;
  if (strcmp(want, got) != 0 || c.n != s.n) {
    if (bad < 20) printf("MISMATCH bits=%08x want=%s got=%s (count %u)\n", bits, want, got, c.n);
    ++bad;
  }
}

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  const uint32_t edge[] = {0u, 0x80000000u, 1u, 0x007FFFFFu, 0x00800000u, 0x7F7FFFFFu, 0x7F800000u, 0xFF800000u, 0x7FC00000u,
                           0xFFC00000u, 0x3F800000u, 0x42C80000u, 0x49742400u, 0x49742408u, 0x497423F8u, 0x38D1B717u, 0x38D1B718u,
                           0x38D1B716u, 0x3A83126Fu, 0x4B7FFFFFu, 0x4B800000u, 0x5F000000u};
  for (uint32_t b : edge) check(b), check(b ^ 0x80000000u);
  for (long i = 0; i < n; ++i) check((uint32_t)next());                       // all magnitudes
  for (long i = 0; i < n; ++i) {                                               // percentages: ident*100/len and parsed "dd.dd"
    const uint32_t len = 1 + (uint32_t)(next() % 5000), id = (uint32_t)(next() % (len + 1));
    const float v = (float)id * 100 / (float)len;
    uint32_t b;
    memcpy(&b, &v, 4);
    check(b);
    char t[32];
    snprintf(t, sizeof t, "%u.%02u", (unsigned)(next() % 101), (unsigned)(next() % 100));
    const float p = strtof(t, nullptr);
    memcpy(&b, &p, 4);
    check(b);
  }
  for (uint32_t ex = 0; ex < 255; ++ex)                                        // around every power of two, ties
    for (uint32_t fr : {0u, 1u, 2u, 0x400000u, 0x7FFFFEu, 0x7FFFFFu}) check((ex << 23) | fr);
  for (int X = -44; X <= 38; ++X) {                                            // neighbours of 10^X and of d.ddddd5 * 10^X
    char t[32];
    for (const char *mant : {"1", "9.999995", "9.99999", "1.000005", "1.234565", "1.234575", "9.999994", "9.999996"}) {
      snprintf(t, sizeof t, "%se%d", mant, X);
      const float v = strtof(t, nullptr);
      uint32_t b;
      memcpy(&b, &v, 4);
      for (int d = -3; d <= 3; ++d) check(b + (uint32_t)d);
    }
  }
  printf("%ld values, %ld mismatches\n", total, bad);
  return bad != 0;
}
