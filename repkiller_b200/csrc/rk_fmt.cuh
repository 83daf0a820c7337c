// Exact "%g" (precision 6) of a binary32 value and decimal integers, for the device-side output writer (K6) and its
// CPU test (tests/test_fmt_cpu.py compiles this header with g++ and compares it with printf).
//
// Reference: /root/reference/src/commonFunctions.cpp:101-104 streams `similarity` and `identity` (floats) through
// ostream::operator<<, i.e. printf("%g") with 6 significant digits in the default rounding mode: the EXACT binary value
// is rounded half-to-even to 6 significant decimal digits, fixed notation when -4 <= exponent < 6, trailing zeros
// stripped.  glibc does this with multi-precision arithmetic; so does this code, with integers:
//   x = m * 2^e  (m < 2^24).  Candidate decimal exponent X, then N = round_half_even(x * 10^(5-X)) must land in
//   [10^5, 10^6).  For 10^-5 <= x < 10^7 every quantity fits in 64 bits (the common case: percentages); anything
//   else goes through a small base-2^32 big-number routine (at most 6 limbs).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RK_HD __host__ __device__
#else
#define RK_HD
#endif

namespace rkfmt {

struct CountSink {  // pass 1: only the length
  uint32_t n = 0;
  RK_HD void put(char) { ++n; }
};
struct BufSink {  // pass 2 / tests: bytes into memory
  char *p;
  uint32_t n = 0;
  RK_HD explicit BufSink(char *q) : p(q) {}
  RK_HD void put(char c) { p[n++] = c; }
};

template <class Sink>
RK_HD inline void put_u64(Sink &s, uint64_t v) {
  char tmp[20];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + (int)(v % 10));
    v /= 10;
  } while (v);
  while (n) s.put(tmp[--n]);
}

// ---- big numbers: little-endian base-2^32 limbs, only what %g needs ------------------------------------
struct Big {
  uint32_t w[8];
  int n;  // used limbs
};
RK_HD inline void big_from_u64(Big &b, uint64_t v) {
  for (int i = 0; i < 8; ++i) b.w[i] = 0;
  b.w[0] = (uint32_t)v, b.w[1] = (uint32_t)(v >> 32);
  b.n = b.w[1] ? 2 : 1;
}
RK_HD inline void big_mul_small(Big &b, uint32_t k) {
  uint64_t carry = 0;
  for (int i = 0; i < b.n; ++i) {
    const uint64_t t = (uint64_t)b.w[i] * k + carry;
    b.w[i] = (uint32_t)t;
    carry = t >> 32;
  }
  if (carry && b.n < 8) b.w[b.n++] = (uint32_t)carry;
}
RK_HD inline void big_shl(Big &b, int sh) {  // sh < 32 * (8 - n)
  const int ws = sh >> 5, bs = sh & 31;
  for (int i = 7; i >= 0; --i) {
    uint64_t v = 0;
    if (i - ws >= 0) v = (uint64_t)b.w[i - ws] << bs;
    if (bs && i - ws - 1 >= 0) v |= (uint64_t)b.w[i - ws - 1] >> (32 - bs);
    b.w[i] = (uint32_t)v;
  }
  b.n = 8;
  while (b.n > 1 && b.w[b.n - 1] == 0) --b.n;
}
RK_HD inline int big_cmp(const Big &a, const Big &b) {
  for (int i = 7; i >= 0; --i) {
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  }
  return 0;
}
RK_HD inline void big_sub(Big &a, const Big &b) {  // a -= b, a >= b
  int64_t borrow = 0;
  for (int i = 0; i < 8; ++i) {
    int64_t t = (int64_t)a.w[i] - b.w[i] - borrow;
    borrow = t < 0;
    a.w[i] = (uint32_t)(t + (borrow << 32));
  }
}
RK_HD inline bool big_is_zero(const Big &a) {
  for (int i = 0; i < 8; ++i)
    if (a.w[i]) return false;
  return true;
}

// decimal digits of num/den: the integer part must be < 10; returns it and leaves the remainder in num
RK_HD inline int big_div_digit(Big &num, const Big &den) {
  int d = 0;
  while (big_cmp(num, den) >= 0) {
    big_sub(num, den);
    ++d;
  }
  return d;
}

// Six significant digits of x = m * 2^e (m != 0), rounded half-to-even on the exact value: digits in [100000, 999999],
// X = decimal exponent of the first digit.
RK_HD inline void six_digits(uint32_t m, int e, uint32_t &digits, int &X) {
  // x = num / den with num = m * 2^max(e,0), den = 2^max(-e,0); scale by powers of ten until 1 <= num/den < 10
  Big num, den;
  big_from_u64(num, m);
  big_from_u64(den, 1);
  if (e >= 0) big_shl(num, e);
  else big_shl(den, -e);  // -e <= 149 + 23: fits 6 limbs
  X = 0;
  // estimate X from the binary exponent of x (bit length of m plus e): log10(2) ~ 1233/4096
  int bl = 0;
  for (uint32_t t = m; t; t >>= 1) ++bl;
  int est = ((bl + e - 1) * 1233) >> 12;  // floor(log10(x)) or one less
  if (est > 0) {
    for (int i = 0; i < est; ++i) big_mul_small(den, 10);
  } else {
    for (int i = 0; i < -est; ++i) big_mul_small(num, 10);
  }
  X = est;
  Big ten_den = den;
  big_mul_small(ten_den, 10);
  while (big_cmp(num, ten_den) >= 0) {  // x / 10^X >= 10
    den = ten_den;
    big_mul_small(ten_den, 10);
    ++X;
  }
  while (big_cmp(num, den) < 0) {  // x / 10^X < 1
    big_mul_small(num, 10);
    --X;
  }
  uint32_t d = 0;
  for (int i = 0; i < 6; ++i) {
    d = d * 10 + (uint32_t)big_div_digit(num, den);
    big_mul_small(num, 10);
  }
  // remainder: num / den is (10 x the discarded fraction); round half to even: compare num with 5 * den
  Big half = den;
  big_mul_small(half, 5);
  const int c = big_cmp(num, half);
  if (c > 0 || (c == 0 && (d & 1u))) ++d;
  if (d == 1000000u) {
    d = 100000u;
    ++X;
  }
  digits = d;
}

// fast path for 2^-17 <= x < 2^23 (covers 1e-5 .. 8e6): everything fits in 64 bits
RK_HD inline bool six_digits_fast(uint32_t m, int e, uint32_t &digits, int &X) {
  // m has 24 significant bits for normal numbers; x = m * 2^e
  if (e > 0 || e < -40) return false;
  const int k = -e;  // x = m / 2^k, k in [0, 40]
  // candidate s = 5 - X such that N = x * 10^s in [1e5, 1e6)
  const uint64_t ip = k < 64 ? ((uint64_t)m >> k) : 0;  // integer part
  int Xc;
  if (ip >= 100000u) Xc = ip >= 1000000u ? 6 : 5;
  else if (ip >= 10000u) Xc = 4;
  else if (ip >= 1000u) Xc = 3;
  else if (ip >= 100u) Xc = 2;
  else if (ip >= 10u) Xc = 1;
  else if (ip >= 1u) Xc = 0;
  else {
    // x < 1: count leading decimal zeros by scaling; at most 5 steps in this range
    Xc = -1;
    uint64_t t = (uint64_t)m * 10;  // x * 10
    while ((t >> k) == 0 && Xc > -6) {
      t *= 10;
      --Xc;
    }
    if (Xc <= -6) return false;
  }
  if (Xc > 5) return false;
  static const uint64_t P10[12] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                   1000000000ull, 10000000000ull, 100000000000ull};
  const int s = 5 - Xc;  // 0 .. 10
  const uint64_t N = (uint64_t)m * P10[s];  // < 2^24 * 10^10 < 2^58
  uint64_t q = k ? (N >> k) : N;
  const uint64_t rem = k ? (N & ((1ull << k) - 1)) : 0;
  const uint64_t halfv = k ? (1ull << (k - 1)) : 0;
  if (k && (rem > halfv || (rem == halfv && (q & 1)))) ++q;
  if (q >= 1000000u) {  // rounding carried into a seventh digit: 999999.5 -> 1.00000e+06
    q = 100000u;
    ++Xc;
  }
  if (q < 100000u) return false;  // (cannot happen when Xc is right; let the exact path decide)
  digits = (uint32_t)q;
  X = Xc;
  return true;
}

// printf("%g", (double)f)
template <class Sink>
RK_HD inline void put_g6(Sink &s, uint32_t bits) {
  const bool neg = (bits >> 31) != 0;
  const uint32_t ex = (bits >> 23) & 0xFFu, fr = bits & 0x7FFFFFu;
  if (ex == 0xFFu) {
    if (neg) s.put('-');
    if (fr) s.put('n'), s.put('a'), s.put('n');
    else s.put('i'), s.put('n'), s.put('f');
    return;
  }
  if (neg) s.put('-');
  if (ex == 0 && fr == 0) {
    s.put('0');
    return;
  }
  const uint32_t m = ex ? (fr | 0x800000u) : fr;
  const int e = (ex ? (int)ex : 1) - 150;
  uint32_t d;
  int X;
  if (!six_digits_fast(m, e, d, X)) six_digits(m, e, d, X);
  char dig[6];
  for (int i = 5; i >= 0; --i) {
    dig[i] = (char)('0' + (int)(d % 10));
    d /= 10;
  }
  int nd = 6;
  while (nd > 1 && dig[nd - 1] == '0') --nd;  // %g strips trailing zeros
  if (X < -4 || X >= 6) {  // d.ddddde+XX
    s.put(dig[0]);
    if (nd > 1) {
      s.put('.');
      for (int i = 1; i < nd; ++i) s.put(dig[i]);
    }
    s.put('e');
    int ax = X;
    if (ax < 0) s.put('-'), ax = -ax;
    else s.put('+');
    s.put((char)('0' + ax / 10));
    s.put((char)('0' + ax % 10));
  } else if (X >= 0) {
    for (int i = 0; i <= X; ++i) s.put(i < nd ? dig[i] : '0');
    if (nd > X + 1) {
      s.put('.');
      for (int i = X + 1; i < nd; ++i) s.put(dig[i]);
    }
  } else {
    s.put('0');
    s.put('.');
    for (int i = 0; i < -X - 1; ++i) s.put('0');
    for (int i = 0; i < nd; ++i) s.put(dig[i]);
  }
}

}  // namespace rkfmt
