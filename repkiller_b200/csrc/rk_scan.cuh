// Exclusive prefix sums over u32 (block reduce -> scan of block sums -> block apply), generic in the load
// functor so that predicates (e.g. "is a root") are scanned without materialising a flag array.
#pragma once
#include "rk_common.cuh"

namespace rk {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

struct LoadU32 {
  const u32 *p;
  __device__ __forceinline__ u32 operator()(u64 i) const { return p[i]; }
};

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane_id() >= (u32)d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a CTA of NT threads; returns the exclusive prefix
template <int NT>
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *total) {
  __shared__ u32 warp_sums[NT / 32];
  __shared__ u32 s_total;
  const u32 incl = warp_incl_scan(v);
  const u32 w = threadIdx.x >> 5;
  if (lane_id() == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    u32 s = lane_id() < NT / 32 ? warp_sums[lane_id()] : 0;
    const u32 si = warp_incl_scan(s);
    if (lane_id() < NT / 32) warp_sums[lane_id()] = si - s;
    if (lane_id() == 31) s_total = si;
  }
  __syncthreads();
  const u32 r = warp_sums[w] + incl - v;
  if (total) *total = s_total;
  __syncthreads();  // shared scratch is reused by the next call
  return r;
}

template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(Load in, u64 n, u32 *__restrict__ bsum) {
  const u64 base = (u64)blockIdx.x * SCAN_CHUNK;
  u32 s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const u64 i = base + threadIdx.x + (u64)j * SCAN_THREADS;
    if (i < n) s += in(i);
  }
  u32 total;
  block_excl_scan<SCAN_THREADS>(s, &total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single CTA: exclusive scan of bsum[0..nb) in place, grand total to bsum[nb]
static __global__ void __launch_bounds__(1024) k_scan_blocksums(u32 *bsum, u32 nb) {
  __shared__ u32 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (u32 base = 0; base < nb; base += 1024) {
    const u32 i = base + threadIdx.x;
    const u32 v = i < nb ? bsum[i] : 0;
    u32 total;
    const u32 ex = block_excl_scan<1024>(v, &total);
    const u32 c = carry;
    if (i < nb) bsum[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

// `out` may alias the array `in` reads: every CTA loads its chunk before it stores
template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Load in, u32 *out, u64 n, const u32 *bsum) {
  const u64 base = (u64)blockIdx.x * SCAN_CHUNK + (u64)threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS];
  u32 s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = (base + j < n) ? in(base + j) : 0;
    s += v[j];
  }
  u32 ex = block_excl_scan<SCAN_THREADS>(s, nullptr) + bsum[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = ex;
    ex += v[j];
  }
}

static inline u64 scan_work_words(u64 n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 2; }

// bsum needs scan_work_words(n) words; the grand total ends up in bsum[ceil(n/SCAN_CHUNK)]; returns launches
template <class Load>
static int exclusive_scan_u32(Load in, u32 *out, u64 n, u32 *bsum, cudaStream_t st) {
  if (n == 0) return 0;
  const u32 nb = (u32)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  KScope ks(KID_SCAN, st, n);
  k_scan_reduce<Load><<<nb, SCAN_THREADS, 0, st>>>(in, n, bsum);
  k_scan_blocksums<<<1, 1024, 0, st>>>(bsum, nb);
  k_scan_apply<Load><<<nb, SCAN_THREADS, 0, st>>>(in, out, n, bsum);
  return 3;
}

}  // namespace rk
