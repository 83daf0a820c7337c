// Synthetic workload generator on the device (tooling for benchmarks at sizes the host generator is too slow for:
// BASELINE config 5, 1e9 fragments).  Bit-for-bit the same function of (seed, i) as repkiller_b200/gen.py — the host
// generator is the definition, tests/test_gpu_parity.py::test_device_generator compares the two.
#include "rk_common.cuh"

namespace rk {

struct GenParams {
  u64 seed, lx, ly, families, ax, ay, tandem_every;
  double p_rep;
};

__device__ __forceinline__ u64 splitmix64(u64 x) {
  u64 z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ void put_u64(u8 *p, u64 v) {
#pragma unroll
  for (int b = 0; b < 8; ++b) p[b] = (u8)(v >> (8 * b));
}

__global__ void __launch_bounds__(256) k_gen(GenParams g, u64 start, u64 count, u8 *__restrict__ out) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const u64 i = start + t;
  const u64 base = g.seed * 0x9E3779B97F4A7C15ull + 16 * i;
  auto u = [&](int k) { return splitmix64(base + (u64)k); };
  auto fam = [&](u64 j, u64 tt) { return splitmix64((g.seed * 0x9E3779B97F4A7C15ull) ^ (0xD1B54A32D192ED03ull + 64 * j + tt)); };
  const double inv53 = 1.0 / 9007199254740992.0;
  const bool is_rep = __dmul_rn((double)(u(0) >> 11), inv53) < g.p_rep;
  u64 len, xs, ys;
  if (!is_rep) {
    len = 40 + u(7) % 2961;
    xs = u(8) % (g.lx - len - 1);
    ys = u(9) % (g.ly - len - 1);
  } else {
    const u64 j = u(1) % g.families;
    const u64 fl = 40 + fam(j, 0) % 1961;
    const u64 a = u(2) % g.ax, b = u(3) % g.ay;
    const u64 span_x = g.lx - fl * (g.ax + 1) - 32, span_y = g.ly - fl * (g.ay + 1) - 32;
    const bool tandem = g.tandem_every && (j % g.tandem_every) == 0;
    const u64 xa = tandem ? 8 + fam(j, 1) % span_x + a * fl : 8 + fam(j, 1 + a) % span_x;
    const u64 ya = tandem ? 8 + fam(j, 17) % span_y + b * fl : 8 + fam(j, 17 + b) % span_y;
    len = fl + u(6) % 7 - 3;
    xs = xa + u(4) % 11 - 5;
    ys = ya + u(5) % 11 - 5;
  }
  const u8 strand = (u(10) & 1) == 0 ? 'f' : 'r';
  const double frac = __dadd_rn(0.65, __dmul_rn(0.35, __dmul_rn((double)(u(11) >> 11), inv53)));
  const u64 ident_true = (u64)floor(__dmul_rn((double)len, frac));
  const float sim = __double2float_rn(__ddiv_rn(__dmul_rn(100.0, (double)ident_true), (double)len));
  u8 *p = out + t * FRAG_BYTES;
  put_u64(p + 0, xs - ys);           // diag (two's complement)
  put_u64(p + 8, xs);
  put_u64(p + 16, ys);
  put_u64(p + 24, xs + len - 1);
  put_u64(p + 32, ys + len - 1);
  put_u64(p + 40, len);
  put_u64(p + 48, (u64)sim);         // the reference's (uint64_t) stof(similarity)
  put_u64(p + 56, 4 * ident_true);
  const u32 sb = __float_as_uint(sim);
  p[64] = (u8)sb, p[65] = (u8)(sb >> 8), p[66] = (u8)(sb >> 16), p[67] = (u8)(sb >> 24);
  put_u64(p + 68, 0);                // seqX
  put_u64(p + 76, 1);                // seqY
  put_u64(p + 84, 0);                // block
  p[92] = strand;
#pragma unroll
  for (int b = 0; b < 16; ++b) p[93 + b] = 0;  // evalue
}

int launch_gen(u64 seed, u64 lx, u64 ly, double p_rep, u64 families, u64 ax, u64 ay, u64 tandem_every, u64 start, u64 count,
               u8 *out, cudaStream_t st) {
  if (count == 0) return 0;
  GenParams g{seed, lx, ly, families, ax, ay, tandem_every, p_rep};
  k_gen<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(g, start, count, out);
  return 1;
}

}  // namespace rk
