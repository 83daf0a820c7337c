// K2 — stable LSD radix sort of (u32 key, u32 value) pairs, 8-bit digits.
//
// Replaces the reference's bucket arrays: FragmentsDatabase's vector<FragFile>[vsize] filled by push_back
// (/root/reference/src/FragmentsDatabase.cpp:84-85,96-97: a stable bucketing by xStart/10) and the
// per-bucket forward_lists of SequenceOcupationList (/root/reference/src/SequenceOcupationList.cpp:3-8,93-96:
// one list per center/100).  Stability is what carries the reference's visiting order (file order inside an
// X bucket, processing order inside an occupation bucket).
//
// Per digit pass: (1) per-tile digit histogram, (2) exclusive scan of the [digit][tile] count matrix,
// (3) scatter: a 4096-key tile is ranked stably inside the CTA (warp-level match_any multisplit + per-warp
// digit counters), reordered through shared memory so that each digit's run leaves as one contiguous,
// coalesced store, and placed at the scanned global offset.  HBM traffic per pass and key: 4 B (histogram)
// + 8 B read + 8 B written.
#include "rk_common.cuh"
#include "rk_scan.cuh"

namespace rk {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_TILE = 32 * RS_ITEMS;  // 512 consecutive keys per warp
constexpr int RADIX = 256;

// ---- radix passes -------------------------------------------------------------------------------------

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const u32 *__restrict__ keys, u64 n, int shift, u32 mask,
                                                           u32 *__restrict__ counts, u32 num_tiles) {
  __shared__ u32 hist[RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u64 i = base + threadIdx.x + (u64)j * RS_THREADS;
    if (i < n) atomicAdd(&hist[(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  counts[(u64)threadIdx.x * num_tiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
    k_radix_scatter(const u32 *__restrict__ kin, const u32 *__restrict__ vin, u32 *__restrict__ kout, u32 *__restrict__ vout,
                    u64 n, int shift, u32 mask, const u32 *__restrict__ offsets, u32 num_tiles) {
  __shared__ u32 warp_cnt[RS_WARPS][RADIX];
  __shared__ u32 digit_base[RADIX];
  __shared__ u32 gbase[RADIX];
  __shared__ u32 skeys[RS_TILE];
  __shared__ u32 svals[RS_TILE];

  const u32 tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const u64 tile_start = (u64)blockIdx.x * RS_TILE;
  const u32 nvalid = (u32)((n - tile_start < (u64)RS_TILE) ? (n - tile_start) : RS_TILE);

#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) warp_cnt[q][tid] = 0;

  // tile order: warp w owns keys [w*512, w*512+512), item j of lane l is key w*512 + j*32 + l
  u32 key[RS_ITEMS], val[RS_ITEMS], rnk[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 local = w * RS_WARP_TILE + j * 32 + lane;
    const u64 i = tile_start + local;
    if (local < nvalid) {
      key[j] = kin[i];
      val[j] = vin ? vin[i] : (u32)i;
    } else {
      key[j] = 0xFFFFFFFFu;  // past the end: largest digit, last in tile order => ranks after every valid key
      val[j] = 0;
    }
  }
  __syncthreads();

  const u32 lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 d = (key[j] >> shift) & mask;
    const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
    const int leader = __ffs(peers) - 1;
    u32 old = 0;
    if ((int)lane == leader) {
      old = warp_cnt[w][d];
      warp_cnt[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xFFFFFFFFu, old, leader);
    rnk[j] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();

  // thread tid = digit tid: exclusive prefix over the warps, then over the digits
  u32 total = 0;
#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) {
    const u32 c = warp_cnt[q][tid];
    warp_cnt[q][tid] = total;
    total += c;
  }
  const u32 dbase = block_excl_scan<RS_THREADS>(total, nullptr);
  digit_base[tid] = dbase;
  gbase[tid] = offsets[(u64)tid * num_tiles + blockIdx.x] - dbase;
  __syncthreads();

#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 d = (key[j] >> shift) & mask;
    const u32 pos = digit_base[d] + warp_cnt[w][d] + rnk[j];
    skeys[pos] = key[j];
    svals[pos] = val[j];
  }
  __syncthreads();

  for (u32 p = tid; p < nvalid; p += RS_THREADS) {
    const u32 k = skeys[p];
    const u32 dst = gbase[(k >> shift) & mask] + p;
    kout[dst] = k;
    vout[dst] = svals[p];
  }
}

__global__ void k_iota_copy(const u32 *__restrict__ kin, const u32 *__restrict__ vin, u32 *__restrict__ kout,
                            u32 *__restrict__ vout, u64 n) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    kout[i] = kin[i];
    vout[i] = vin ? vin[i] : (u32)i;
  }
}

static inline u64 align_up(u64 x, u64 a) { return (x + a - 1) / a * a; }

u64 sort_work_bytes(u64 n) {
  const u64 tiles = (n + RS_TILE - 1) / RS_TILE;
  const u64 counts = tiles * RADIX;
  const u64 nb = (counts + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return align_up(counts * 4, 256) + align_up((nb + 2) * 4, 256);
}

int launch_sort_pairs(const u32 *keys_in, const u32 *vals_in, u32 *keys_out, u32 *vals_out, u32 *keys_tmp,
                      u32 *vals_tmp, u64 n, int key_bits, void *work, cudaStream_t st) {
  if (n == 0) return 0;
  if (key_bits < 1) key_bits = 1;
  if (key_bits > 32) key_bits = 32;
  const int passes = (key_bits + 7) / 8;
  const u32 tiles = (u32)((n + RS_TILE - 1) / RS_TILE);
  u32 *counts = reinterpret_cast<u32 *>(work);
  u32 *bsum = reinterpret_cast<u32 *>(reinterpret_cast<char *>(work) + align_up((u64)tiles * RADIX * 4, 256));
  int launches = 0;
  const u32 *ksrc = keys_in, *vsrc = vals_in;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    const int bits = (key_bits - shift) < 8 ? (key_bits - shift) : 8;
    const u32 mask = (1u << bits) - 1;
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    u32 *kdst = to_out ? keys_out : keys_tmp;
    u32 *vdst = to_out ? vals_out : vals_tmp;
    {
      KScope ks(KID_RADIX_HIST, st);
      k_radix_hist<<<tiles, RS_THREADS, 0, st>>>(ksrc, n, shift, mask, counts, tiles);
    }
    launches += 1 + exclusive_scan_u32(LoadU32{counts}, counts, (u64)tiles * RADIX, bsum, st);
    {
      KScope ks(KID_RADIX_SCATTER, st);
      k_radix_scatter<<<tiles, RS_THREADS, 0, st>>>(ksrc, vsrc, kdst, vdst, n, shift, mask, counts, tiles);
    }
    launches += 1;
    ksrc = kdst;
    vsrc = vdst;
  }
  return launches;
}

}  // namespace rk
