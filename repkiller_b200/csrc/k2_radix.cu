// K2 — stable LSD radix sort of (u32 key, u32 value) pairs, 8-bit digits.
//
// Replaces the reference's bucket arrays: FragmentsDatabase's vector<FragFile>[vsize] filled by push_back
// (/root/reference/src/FragmentsDatabase.cpp:84-85,96-97: a stable bucketing by xStart/10) and the
// per-bucket forward_lists of SequenceOcupationList (/root/reference/src/SequenceOcupationList.cpp:3-8,93-96:
// one list per center/100).  Stability is what carries the reference's visiting order (file order inside an
// X bucket, processing order inside an occupation bucket).
//
// One-sweep passes (the path every call below 2^30 elements takes): the digit counts of all passes come from the
// kernel that produced the keys (HistOut in rk_common.cuh) or from k_onesweep_hist; then one kernel per digit: a
// 4096-key tile is ranked stably inside the CTA (ballot multisplit + per-warp digit counters), reordered through
// shared memory so that each digit's run leaves as one contiguous store, and placed at the offset its tile obtains
// by decoupled look-back over the preceding tiles' counts.  Algorithmic traffic per pass and key: 8 B read + 8 B
// written (served by the 126 MB L2 for the 10M-element sorts of config 2).
// The three-kernel variant (per-tile histogram, scan of the [digit][tile] matrix, scatter) is kept for n >= 2^30 and
// as a tuning reference (RK_SORT_3K).
#include <cstdlib>

#include "rk_common.cuh"
#include "rk_scan.cuh"

namespace rk {

constexpr int RS_THREADS = 256;
#ifndef RK_RS_ITEMS
#define RK_RS_ITEMS 16
#endif
#ifndef RK_RS_MINBLOCKS
#define RK_RS_MINBLOCKS 3
#endif
constexpr int RS_ITEMS = RK_RS_ITEMS;
#ifndef RK_OS_THREADS
#define RK_OS_THREADS 256
#endif
#ifndef RK_OS_LB
#define RK_OS_LB 8
#endif
constexpr int OS_THREADS = RK_OS_THREADS;  // one-sweep pass CTA (a multiple of 256: one thread per digit does the prefix work)
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_TILE = OS_THREADS * RS_ITEMS;
constexpr int OS_SMEM_BYTES = (OS_WARPS * 256 + 2 * 256 + 2 * OS_TILE) * 4;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_TILE = 32 * RS_ITEMS;  // 512 consecutive keys per warp
constexpr int RADIX = 256;

// ---- radix passes -------------------------------------------------------------------------------------

// lanes of the warp whose 8-bit digit equals this lane's
__device__ __forceinline__ u32 digit_peers(u32 d) {
  // (measured: __match_any_sync here instead of the eight ballots makes a pass 113 us instead of 75 us)
  u32 peers = 0xFFFFFFFFu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const u32 bit = (d >> b) & 1u;
    const u32 bal = __ballot_sync(0xFFFFFFFFu, bit);
    peers &= bal ^ (bit - 1u);  // bit ? bal : ~bal
  }
  return peers;
}

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

  // stable rank of every key among the keys of its warp: first the peer masks of all 16 rounds (8 independent
  // ballots per round pipeline; match.any is a long-latency instruction and would serialise the rounds), then the
  // short dependent chain through the per-warp digit counters.
  const u32 lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) rnk[j] = digit_peers((key[j] >> shift) & mask);
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 d = (key[j] >> shift) & mask;
    const u32 peers = rnk[j];
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

// ---- one-sweep path: one global histogram for all digit positions, then per pass a single kernel whose tiles
// chain their per-digit counts with decoupled look-back (flag+value packed in one 32-bit word; tile ids come from
// an atomic ticket counter, so a tile only ever waits on tiles that are already running).  Measured alternatives
// (B200, 10M pairs, one pass): serial look-back 100 us, batched x16 86 us, a dedicated scan-agent CTA 194 us,
// histogram+scan+scatter as three kernels 28+21+62 us ---------------------------------------------------------

constexpr u32 OS_AGG = 1u << 30;  // the tile's own digit count is published
constexpr u32 OS_PFX = 2u << 30;  // the inclusive prefix over tiles 0..t is published
constexpr u32 OS_VAL = (1u << 30) - 1;
constexpr u64 OS_MAX_N = 1ull << 30;
constexpr int OS_MAX_PASSES = 4;

// Keys of a warp often share a digit (the top digits of nearly sorted keys: gids, bucket keys in processing order), which
// would serialise 32 shared-memory atomics on one address; a uniform digit is counted once by lane 0.
template <int VEC>
__global__ void __launch_bounds__(256) k_onesweep_hist(const u32 *__restrict__ keys, u64 n, int passes, int key_bits,
                                                       u32 *__restrict__ ghist) {
  __shared__ u32 h[OS_MAX_PASSES][RADIX];
  for (int p = 0; p < OS_MAX_PASSES; ++p) h[p][threadIdx.x] = 0;
  __syncthreads();
  const u32 last_mask = (1u << (key_bits - 8 * (passes - 1))) - 1;
  const u32 lane = threadIdx.x & 31;
  const u64 nv = n / VEC;  // whole vectors; the loop bounds are warp-uniform so every lane takes part in the votes
  const u64 warp0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull;
  for (u64 base = warp0; base < nv; base += (u64)gridDim.x * blockDim.x) {
    const u64 i = base + lane;
    const bool valid = i < nv;
    u32 k[VEC];
    if (VEC == 4) {
      const uint4 v = valid ? reinterpret_cast<const uint4 *>(keys)[i] : make_uint4(0, 0, 0, 0);
      k[0] = v.x, k[VEC > 1 ? 1 : 0] = v.y, k[VEC > 2 ? 2 : 0] = v.z, k[VEC > 3 ? 3 : 0] = v.w;
    } else {
      k[0] = valid ? keys[i] : 0;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
#pragma unroll
      for (int p = 0; p < OS_MAX_PASSES; ++p) {
        if (p < passes) {
          const u32 d = (k[j] >> (8 * p)) & (p == passes - 1 ? last_mask : 0xFFu);
          const u32 d0 = __shfl_sync(0xFFFFFFFFu, d, 0);
          if (__all_sync(0xFFFFFFFFu, valid && d == d0)) {
            if (lane == 0) atomicAdd(&h[p][d], 32u);
          } else if (valid) {
            atomicAdd(&h[p][d], 1u);
          }
        }
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nv * VEC) {  // the n % VEC keys after the last whole vector
    const u32 kk = keys[nv * VEC + threadIdx.x];
    for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(kk >> (8 * p)) & (p == passes - 1 ? last_mask : 0xFFu)], 1u);
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const u32 c = h[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * RADIX + threadIdx.x], c);
  }
}

// block p: exclusive scan of the 256 bins of pass p, in place
__global__ void __launch_bounds__(RADIX) k_onesweep_bases(u32 *ghist) {
  u32 *g = ghist + blockIdx.x * RADIX;
  const u32 v = g[threadIdx.x];
  g[threadIdx.x] = block_excl_scan<RADIX>(v, nullptr);
}

__global__ void __launch_bounds__(OS_THREADS, RK_RS_MINBLOCKS)
    k_onesweep_pass(const u32 *__restrict__ kin, const u32 *__restrict__ vin, u32 *__restrict__ kout, u32 *__restrict__ vout,
                    u64 n, int shift, u32 mask, const u32 *__restrict__ digit_start, u32 *state, u32 *tile_counter, u32 *err) {
  extern __shared__ u32 os_smem[];
  u32(*warp_cnt)[RADIX] = reinterpret_cast<u32(*)[RADIX]>(os_smem);  // [OS_WARPS][RADIX]
  u32 *digit_base = os_smem + OS_WARPS * RADIX;
  u32 *gbase = digit_base + RADIX;
  u32 *skeys = gbase + RADIX;
  u32 *svals = skeys + OS_TILE;
  __shared__ u32 s_tile;

  const u32 tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (u32 q = tid; q < (u32)(OS_WARPS * RADIX); q += OS_THREADS) os_smem[q] = 0;
  __syncthreads();
  const u32 tile = s_tile;
  const u64 tile_start = (u64)tile * OS_TILE;
  const u32 nvalid = (u32)((n - tile_start < (u64)OS_TILE) ? (n - tile_start) : OS_TILE);

  u32 key[RS_ITEMS], val[RS_ITEMS], rnk[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 local = w * RS_WARP_TILE + j * 32 + lane;
    const u64 i = tile_start + local;
    if (local < nvalid) {
      key[j] = kin[i];
      val[j] = vin ? vin[i] : (u32)i;
    } else {
      key[j] = 0xFFFFFFFFu;
      val[j] = 0;
    }
  }

  // stable rank of every key among the keys of its warp: first the peer masks of all 16 rounds (8 independent
  // ballots per round pipeline; match.any is a long-latency instruction and would serialise the rounds), then the
  // short dependent chain through the per-warp digit counters.
  const u32 lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) rnk[j] = digit_peers((key[j] >> shift) & mask);
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 d = (key[j] >> shift) & mask;
    const u32 peers = rnk[j];
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

  // thread tid = digit tid: count of VALID keys with this digit (padding sits in the top digit, after them)
  const bool digit_thread = tid < (u32)RADIX;
  u32 total = 0;
  if (digit_thread) {
#pragma unroll
    for (int q = 0; q < OS_WARPS; ++q) {
      const u32 c = warp_cnt[q][tid];
      warp_cnt[q][tid] = total;
      total += c;
    }
  }
  u32 valid_total = total;
  if (tid == mask) valid_total -= (u32)OS_TILE - nvalid;

  // Publish this tile's counts at once, reorder the tile in shared memory (which needs no global offsets), and only
  // then look back over the predecessors: by then more of them have published their inclusive prefix, so the walk
  // is shorter, and its latency no longer sits between the ranking and the reordering of this tile.
  volatile u32 *my_state = state + (u64)tile * RADIX + tid;
  if (digit_thread && tile != 0) *my_state = OS_AGG | valid_total;

  const u32 dbase = block_excl_scan<OS_THREADS>(total, nullptr);
  if (digit_thread) digit_base[tid] = dbase;
  __syncthreads();

#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const u32 d = (key[j] >> shift) & mask;
    const u32 pos = digit_base[d] + warp_cnt[w][d] + rnk[j];
    skeys[pos] = key[j];
    svals[pos] = val[j];
  }

  u32 excl = 0;
  if (!digit_thread) {
    // threads beyond the 256 digits only help with loading, ranking and writing
  } else if (tile == 0) {
    *my_state = OS_PFX | valid_total;
  } else {
    // Batches of LB independent loads: the tiles of one wave start together and all sit in the AGG state, so the
    // walk back to the last published prefix is as long as the wave; one load in flight would cost an L2 round
    // trip per predecessor.
    constexpr int LB = RK_OS_LB;
    u32 next = tile;  // predecessors next-1, next-2, ... are still to be added
    u32 spins = 0;
    bool done = false;
    while (!done) {
      u32 v[LB];
      const volatile u32 *row = state + (u64)next * RADIX + tid;  // row of tile `next`; predecessors are below it
      if (next >= (u32)LB) {
#pragma unroll
        for (int j = 0; j < LB; ++j) v[j] = *(row - (j + 1) * RADIX);
        // fast path: all LB predecessors have published their count and none its prefix yet
        u32 all_and = v[0], all_or = v[0];
#pragma unroll
        for (int j = 1; j < LB; ++j) all_and &= v[j], all_or |= v[j];
        if ((all_and & OS_AGG) && !(all_or & OS_PFX)) {
          u32 sum = 0;
#pragma unroll
          for (int j = 0; j < LB; ++j) sum += v[j];
          excl += sum - LB * OS_AGG;  // every word carries exactly the AGG flag
          next -= LB;
          continue;
        }
      } else {
#pragma unroll
        for (int j = 0; j < LB; ++j) v[j] = ((u32)j < next) ? *(row - (j + 1) * RADIX) : (2u << 30) /* OS_PFX | 0 */;
      }
      u32 used = 0;
#pragma unroll
      for (int j = 0; j < LB; ++j) {
        if (!done && used == (u32)j) {
          const u32 flag = v[j] & ~OS_VAL;
          if (flag != 0) {
            excl += v[j] & OS_VAL;
            ++used;
            if (flag == OS_PFX) done = true;  // tile 0 always publishes OS_PFX; past it the filler is OS_PFX|0
          }
        }
      }
      next -= used < next ? used : next;
      if (!done && used == 0 && ++spins > (1u << 24)) {  // predecessors are always running: fail loudly, never hang
        atomicOr(err, ERR_SPIN);
        break;
      }
    }
    *my_state = OS_PFX | (excl + valid_total);
  }

  if (digit_thread) gbase[tid] = digit_start[tid] + excl - dbase;
  __syncthreads();

  for (u32 p = tid; p < nvalid; p += OS_THREADS) {
    const u32 k = skeys[p];
    const u32 dst = gbase[(k >> shift) & mask] + p;
    kout[dst] = k;
    vout[dst] = svals[p];
  }
}

static inline u64 align_up(u64 x, u64 a) { return (x + a - 1) / a * a; }

static inline u64 onesweep_state_words(u64 n) {
  const u64 tiles = (n + OS_TILE - 1) / OS_TILE;
  return OS_MAX_PASSES * RADIX + 64 + OS_MAX_PASSES * tiles * RADIX;
}

u64 sort_work_bytes(u64 n) {
  const u64 tiles = (n + RS_TILE - 1) / RS_TILE;
  const u64 counts = tiles * RADIX;
  const u64 nb = (counts + SCAN_CHUNK - 1) / SCAN_CHUNK;
  const u64 three_kernel = align_up(counts * 4, 256) + align_up((nb + 2) * 4, 256);
  const u64 onesweep = align_up(onesweep_state_words(n) * 4, 256);
  return three_kernel > onesweep ? three_kernel : onesweep;
}

cudaError_t sort_init_device() {
  return cudaFuncSetAttribute(k_onesweep_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_SMEM_BYTES);
}

static int launch_onesweep(const u32 *keys_in, const u32 *vals_in, u32 *keys_out, u32 *vals_out, u32 *keys_tmp, u32 *vals_tmp,
                           u64 n, int key_bits, void *work, cudaStream_t st, u32 *err_word, u32 *prehist) {
  const int passes = (key_bits + 7) / 8;
  const u32 tiles = (u32)((n + OS_TILE - 1) / OS_TILE);
  u32 *ghist = reinterpret_cast<u32 *>(work);          // [4][256] (or the producer's counters)
  u32 *counters = ghist + OS_MAX_PASSES * RADIX;          // [4] tile counters, [4] = error word when none was set
  u32 *state = counters + 64;                            // [passes][tiles][256]
  cudaMemsetAsync(work, 0, (OS_MAX_PASSES * RADIX + 64 + (u64)passes * tiles * RADIX) * 4, st);
  u32 *err = err_word ? err_word : counters + 8;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (prehist) {  // the producer of the keys counted the digits (HistOut): only the bin bases are left to do
    ghist = prehist;
    k_onesweep_bases<<<passes, RADIX, 0, st>>>(ghist);
  } else {
    KScope ks(KID_RADIX_HIST, st, n);
    u64 blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > (u64)sms * 8) blocks = (u64)sms * 8;
    if (((uintptr_t)keys_in & 15) == 0) k_onesweep_hist<4><<<(unsigned)blocks, 256, 0, st>>>(keys_in, n, passes, key_bits, ghist);
    else k_onesweep_hist<1><<<(unsigned)blocks, 256, 0, st>>>(keys_in, n, passes, key_bits, ghist);
    k_onesweep_bases<<<passes, RADIX, 0, st>>>(ghist);
  }
  const u32 *ksrc = keys_in, *vsrc = vals_in;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    const int bits = (key_bits - shift) < 8 ? (key_bits - shift) : 8;
    const u32 mask = (1u << bits) - 1;
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    u32 *kdst = to_out ? keys_out : keys_tmp;
    u32 *vdst = to_out ? vals_out : vals_tmp;
    KScope ks(KID_RADIX_SCATTER, st, n);
    k_onesweep_pass<<<tiles, OS_THREADS, OS_SMEM_BYTES, st>>>(ksrc, vsrc, kdst, vdst, n, shift, mask, ghist + p * RADIX,
                                                  state + (u64)p * tiles * RADIX, counters + p, err);
    ksrc = kdst;
    vsrc = vdst;
  }
  return 2 + passes;
}

int launch_sort_pairs(const u32 *keys_in, const u32 *vals_in, u32 *keys_out, u32 *vals_out, u32 *keys_tmp,
                      u32 *vals_tmp, u64 n, int key_bits, void *work, cudaStream_t st, u32 *err_word, u32 *prehist) {
  if (n == 0) return 0;
  if (key_bits < 1) key_bits = 1;
  if (key_bits > 32) key_bits = 32;
  static const bool force_3k = getenv("RK_SORT_3K") != nullptr;  // tuning switch: histogram + scan + scatter per pass
  if (n < OS_MAX_N && !force_3k) return launch_onesweep(keys_in, vals_in, keys_out, vals_out, keys_tmp, vals_tmp, n, key_bits, work, st, err_word, prehist);
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
      KScope ks(KID_RADIX_HIST, st, n);
      k_radix_hist<<<tiles, RS_THREADS, 0, st>>>(ksrc, n, shift, mask, counts, tiles);
    }
    launches += 1 + exclusive_scan_u32(LoadU32{counts}, counts, (u64)tiles * RADIX, bsum, st);
    {
      KScope ks(KID_RADIX_SCATTER, st, n);
      k_radix_scatter<<<tiles, RS_THREADS, 0, st>>>(ksrc, vsrc, kdst, vdst, n, shift, mask, counts, tiles);
    }
    launches += 1;
    ksrc = kdst;
    vsrc = vdst;
  }
  return launches;
}

}  // namespace rk
