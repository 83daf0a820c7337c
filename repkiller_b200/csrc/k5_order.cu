// K5 — generate_diagonal_func, sort_groups and the per-fragment output labels.
//
// Reference: /root/reference/src/commonFunctions.cpp:161-177 (diag_func[b] = yStart of the LAST fragment of X
// bucket b: `nh < oh` with oh = +inf never updated is always true; carried forward over empty buckets),
// :148-159 (sort_groups: std::sort of every group with > 1 member by h = |yStart - diag_func[xStart/10]|),
// :106-115 (repval) and :103 (identity).
//
// h only ever reads diag_func at non-empty buckets, so K5a evaluates it per fragment straight from the
// rank-ordered arrays (rank order IS xStart/10 bucket order): the bucket's last fragment is found by galloping
// over the sorted bucket keys.  std::sort is not stable, and equal h values are common inside repeat groups,
// so the member order of a group is whatever libstdc++'s introsort does with that input; K5b therefore runs
// the same algorithm (bits/stl_algo.h: __introsort_loop with the median-of-three to *first, unguarded
// partition, depth limit 2*lg(n) with heap-sort fallback, threshold 16, final insertion sort) on
// (h << 32 | index) words, comparing the high halves only — in parallel: groups of <= 16 by a stable rank per member,
// larger ones by a warp whose partition step is computed from ballot masks (see warp_partition), the largest split
// across warps after the top of the recursion.  Only __partial_sort (the heap-sort fallback at the depth limit) runs on
// one lane.
#include "rk_common.cuh"

namespace rk {

// ---- K5a ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_hkey(const u32 *__restrict__ k0_r, const u32 *__restrict__ ys_r, u32 m,
                                              u32 *__restrict__ h, const u32 *__restrict__ fidx_r,
                                              const float *__restrict__ identity_r, uint4 *__restrict__ hfi_r) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const u32 key = k0_r[i];
  u32 lo = i, step = 1;
  while (lo + step < m && k0_r[lo + step] == key) {
    lo += step;
    step <<= 1;
  }
  u32 hi = lo + step < m ? lo + step : m;  // first position known to be outside the bucket (or m)
  while (hi - lo > 1) {
    const u32 mid = lo + (hi - lo) / 2;
    if (k0_r[mid] == key) lo = mid;
    else hi = mid;
  }
  const u32 hv = absdiff(ys_r[i], ys_r[lo]);  // commonFunctions.cpp:152-155
  if (h) h[i] = hv;
  // everything K5b/c needs about a fragment in one 16-byte word (one gather by rank instead of three)
  if (hfi_r) hfi_r[i] = make_uint4(hv, fidx_r[i], __float_as_uint(identity_r[i]), 0u);
}

int launch_hkey(const u32 *k0_r, const u32 *ys_r, u32 m, u32 *h, cudaStream_t st, const u32 *fidx_r, const float *identity_r,
                uint4 *hfi_r) {
  if (m == 0) return 0;
  KScope ks(KID_HKEY, st, m);
  k_hkey<<<(m + 255) / 256, 256, 0, st>>>(k0_r, ys_r, m, h, fidx_r, identity_r, hfi_r);
  return 1;
}

// full table for rk_diagonal_func: last fragment with bucket key <= b, 0 when there is none (:166)
__global__ void __launch_bounds__(256) k_diag_table(const u32 *__restrict__ k0_r, const u32 *__restrict__ ys_r, u32 m,
                                                    u32 nb, u64 *__restrict__ diag) {
  const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  u32 lo = 0, hi = m;  // first index with key > b
  while (lo < hi) {
    const u32 mid = lo + (hi - lo) / 2;
    if (k0_r[mid] <= b) lo = mid + 1;
    else hi = mid;
  }
  diag[b] = lo ? (u64)ys_r[lo - 1] : 0;
}

int launch_diag_table(const u32 *k0_r, const u32 *ys_r, u32 m, u32 vsize, u64 *diag, void *, cudaStream_t st) {
  if (vsize <= 1) return 0;
  const u32 nb = vsize - 1;
  KScope ks(KID_DIAG, st, nb);
  k_diag_table<<<(nb + 255) / 256, 256, 0, st>>>(k0_r, ys_r, m, nb, diag);
  return 1;
}

// ---- K5b: pieces of libstdc++'s std::sort over an indexable array of packed (h, index) words ---------------

__device__ __forceinline__ bool hless(u64 a, u64 b) { return (u32)(a >> 32) < (u32)(b >> 32); }

template <class P>
__device__ __forceinline__ void swp(P a, int i, int j) {
  const u64 t = a[i];
  a[i] = a[j];
  a[j] = t;
}

template <class P>
__device__ void dev_push_heap(P a, int hole, int top, u64 value) {
  int parent = (hole - 1) / 2;
  while (hole > top && hless(a[parent], value)) {
    a[hole] = a[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[hole] = value;
}

template <class P>
__device__ void dev_adjust_heap(P a, int hole, int len, u64 value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (hless(a[child], a[child - 1])) child--;
    a[hole] = a[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[hole] = a[child - 1];
    hole = child - 1;
  }
  dev_push_heap(a, hole, top, value);
}

// __partial_sort(first, last, last) == make_heap + sort_heap
template <class P>
__device__ void dev_heap_sort(P a, int len) {
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      const u64 v = a[parent];
      dev_adjust_heap(a, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = len;
  while (last > 1) {
    --last;
    const u64 v = a[last];
    a[last] = a[0];
    dev_adjust_heap(a, 0, last, v);
  }
}

struct SubArray {  // a[first..] view so the heap code can stay zero-based
  u64 *p;
  __device__ __forceinline__ u64 &operator[](int i) const { return p[i]; }
};

constexpr int GS_STABLE = 16;        // std::sort of <= 16 elements is one insertion sort == a stable sort by h
constexpr int GS_WARP_CAP0 = 128;    // groups of 17..128 members: one warp each, 2 KB of shared memory per warp
constexpr int GS_WARP_CAP = 1024;    // groups of 129..1024 members: one warp each, 17 KB per warp
constexpr int OT_HEADS = 256;        // a CTA owns the groups whose head lies in its first 256 positions
constexpr int OT_TILE = OT_HEADS + 32;  // one thread per position; a group of <= 16 that starts before 256 ends before 272
constexpr int OT_WARPS = OT_TILE / 32;

// what the output line of a fragment needs besides its group: {h, file index, identity bits}
__device__ __forceinline__ void load_member(const OrderArgs &a, u32 j, u32 &r, u32 &h, u32 &fidx, u32 &ident) {
  if (a.srank) {  // single GPU: one 16-byte gather by processing rank
    r = a.srank[j];
    const uint4 rec = a.hfi_r[r];
    h = rec.x, fidx = rec.y, ident = rec.z;
  } else {  // multi-GPU stages: the three arrays are already in this order
    r = j;
    h = a.h[j], fidx = a.fidx_r[j], ident = __float_as_uint(a.identity_r[j]);
  }
}

// K5b+c for groups of <= 16 members (all but a few per cent).  One thread per position of the gid-sorted list:
//   * group boundaries from a ballot of gid changes (32 positions back and forth are enough to tell "<= 16");
//   * every member counts the members that sort before it (smaller h, or equal h and earlier) — the stable order
//     libstdc++'s insertion sort produces — and drops its source position at that output slot;
//   * larger groups go to the worklist of k_groupsort_warp;
// then every thread writes the output line of its position.
__global__ void __launch_bounds__(OT_TILE) k_order_tile(OrderArgs a) {
  __shared__ u32 s_gid[OT_TILE], s_h[OT_TILE], s_fidx[OT_TILE], s_ident[OT_TILE], s_perm[OT_TILE];
  __shared__ u32 s_heads[OT_WARPS];
  __shared__ u32 s_prev;
  const u32 e = threadIdx.x, lane = e & 31, w = e >> 5;
  const u32 bs = blockIdx.x * OT_HEADS;
  const u32 count = min((u32)OT_TILE, a.m - bs);
  const bool valid = e < count;
  u32 gid = 0xFFFFFFFFu;  // no real group id (gid < m <= 2^32-16): padding never continues a group
  if (valid) {
    u32 r, h, fidx, ident;
    gid = a.sgid[bs + e];
    load_member(a, bs + e, r, h, fidx, ident);
    s_h[e] = h, s_fidx[e] = fidx, s_ident[e] = ident;
  }
  s_gid[e] = gid;
  if (e == 0) s_prev = bs ? a.sgid[bs - 1] : 0;
  __syncthreads();

  const bool is_head = e == 0 || s_gid[e - 1] != gid;
  const u32 hb = __ballot_sync(0xFFFFFFFFu, is_head);
  if (lane == 0) s_heads[w] = hb;
  const bool foreign0 = bs != 0 && s_prev == s_gid[0];
  __syncthreads();
  // start: last head at or before e; end: first head after e.  A head that is more than a word away on either side
  // means more than 16 members.
  u32 start = 0xFFFFFFFFu, end = 0xFFFFFFFFu;
  {
    const u32 own = hb & (0xFFFFFFFFu >> (31 - lane));
    u32 x;
    if (own) start = (w << 5) + 31 - __clz(own);
    else if (w >= 1 && (x = s_heads[w - 1]) != 0) start = ((w - 1) << 5) + 31 - __clz(x);
    const u32 above = hb & (0xFFFFFFFEu << lane);
    if (above) end = (w << 5) + __ffs(above) - 1;
    else if (w + 1 < (u32)OT_WARPS && (x = s_heads[w + 1]) != 0) end = ((w + 1) << 5) + __ffs(x) - 1;
  }
  // (a group that starts before position 256 and shows no end inside the tile has more than 32 members)
  const bool big = start == 0xFFFFFFFFu || end == 0xFFFFFFFFu || end - start > (u32)GS_STABLE;
  if (valid && is_head && big && e < (u32)OT_HEADS && !(e == 0 && foreign0)) {
    const u32 slot = atomicAdd(a.work_count, 1u);
    if (slot < a.work_cap[0]) a.worklist[0][slot] = bs + e;
    else atomicOr(a.err, ERR_WORKLIST);
  }
  const bool mine = valid && !big && start < (u32)OT_HEADS && !(start == 0 && foreign0);
  const u32 n = big ? 0 : end - start;
  if (mine) {
    if (n == 1 || !a.do_sort) {
      s_perm[e] = e;
    } else {
      const u32 h = s_h[e], me = e - start;
      u32 before = 0;
      for (u32 q = 0; q < n; ++q) {
        const u32 o = s_h[start + q];
        before += (o < h || (o == h && q < me)) ? 1u : 0u;
      }
      s_perm[start + before] = e;
    }
  }
  __syncthreads();
  if (mine) {
    const u32 src = s_perm[e], j = bs + e;
    a.out_order[j] = s_fidx[src];
    a.out_gid[j] = gid + a.gid_base;
    a.out_repval[j] = (u8)(n == 1 ? 0 : (e == start ? 1 : 2));  // commonFunctions.cpp:106-115
    a.out_identity[j] = __uint_as_float(s_ident[src]);
  }
}

// ---- the same std::sort, run by a whole warp on a group of 17..1024 members held in shared memory ------------------
// __introsort_loop only ever compares against a pivot, so one __unguarded_partition is a pure function of two bit
// masks over the range: GE (not less than the pivot) and LE (not greater).  Its two pointers stop at the k-th GE
// position from the left and the k-th LE position from the right, swap them, and go on while the first lies left of
// the second.  So, with f(c) = #GE before position c and g(c) = #LE at or after c:
//     number of swaps K = max over c of min(f(c), g(c)),   swap k pairs listA[k] with listB[k]   (k < K),
//     the returned cut  = min(listA[K], listB[K-1])         (listB[-1] = last; a pointer that finds nothing new stops on
//                                                            the position of the last swap, which holds a stopping value).
// All of it is ballots, popcounts and one scan over the 32-position chunks of the range.
// __final_insertion_sort is a stable sort of whatever the partitions left, and it never moves an element out of its
// final partition range (<= 16 elements, everything left of a cut is <= everything right of it): every element counts
// the elements of its own range that must precede it.
struct SortPtrs {  // where one warp keeps a range while it partitions it (shared or global memory)
  u64 *a;
  u32 *listA, *listB;           // positions of the GE / LE elements, by rank from the left / from the right
  u32 *gew, *lew, *gep, *les;   // per chunk: GE / LE masks, #GE in earlier chunks, #LE in later chunks
};

// __unguarded_partition_pivot(first, last) by a warp: median of three to *first, partition of (first, last) around it;
// returns the cut.  Every lane gets the same result.
__device__ int warp_partition(const SortPtrs &m, int first, int last, u32 lane) {
  u64 *a = m.a;
  // __move_median_to_first(first, first+1, mid, last-1)
  const int mid = first + (last - first) / 2;
  const u64 vx = a[first + 1], vy = a[mid], vz = a[last - 1];
  int pick;
  if (hless(vx, vy)) pick = hless(vy, vz) ? mid : (hless(vx, vz) ? last - 1 : first + 1);
  else pick = hless(vx, vz) ? first + 1 : (hless(vy, vz) ? last - 1 : mid);
  __syncwarp();
  if (lane == 0) swp(a, first, pick);
  __syncwarp();
  const u64 pivot = a[first];
  // pass 1: masks of the range (first, last) per chunk of 32 positions
  const int c0 = (first + 1) >> 5, c1 = (last - 1) >> 5;
  const int nch = c1 - c0 + 1;
  for (int c = 0; c < nch; ++c) {
    const int p = ((c0 + c) << 5) + (int)lane;
    const bool in = p > first && p < last;
    const u64 v = in ? a[p] : 0;
    const u32 ge = __ballot_sync(0xFFFFFFFFu, in && !hless(v, pivot));
    const u32 le = __ballot_sync(0xFFFFFFFFu, in && !hless(pivot, v));
    if (lane == 0) m.gew[c] = ge, m.lew[c] = le;
  }
  __syncwarp();
  // exclusive prefix of #GE from the left, exclusive suffix of #LE from the right, over the chunks
  int tot_ge = 0;
  for (int base = 0; base < nch; base += 32) {
    const int c = base + (int)lane;
    int v = c < nch ? __popc(m.gew[c]) : 0;
    const int own = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if ((int)lane >= o) v += t;
    }
    if (c < nch) m.gep[c] = tot_ge + v - own;
    tot_ge += __shfl_sync(0xFFFFFFFFu, v, 31);
  }
  int tot_le = 0;
  for (int base = 0; base < nch; base += 32) {
    const int c = nch - 1 - (base + (int)lane);  // from the right
    int v = c >= 0 ? __popc(m.lew[c]) : 0;
    const int own = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if ((int)lane >= o) v += t;
    }
    if (c >= 0) m.les[c] = tot_le + v - own;
    tot_le += __shfl_sync(0xFFFFFFFFu, v, 31);
  }
  __syncwarp();
  // pass 2: position lists and K
  int kmax = 0;
  const u32 lt = lanemask_lt();
  for (int c = 0; c < nch; ++c) {
    const u32 ge = m.gew[c], le = m.lew[c];
    const int p = ((c0 + c) << 5) + (int)lane;
    const int f = (int)m.gep[c] + __popc(ge & lt);            // #GE before p
    const int g = (int)m.les[c] + __popc(le & ~lt);           // #LE at or after p
    if (p > first && p <= last) kmax = max(kmax, min(f, g));  // split points first+1 .. last
    if ((ge >> lane) & 1u) m.listA[f] = (u32)p;
    if ((le >> lane) & 1u) m.listB[g - 1] = (u32)p;           // g-1 = #LE after p
  }
  // (a split point `last` that falls into a further chunk has g = 0 and cannot raise the maximum)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, o));
  __syncwarp();
  const int K = kmax;
  const int cutA = K < tot_ge ? (int)m.listA[K] : 0x7FFFFFFF;
  const int cutB = K > 0 ? (int)m.listB[K - 1] : last;
  __syncwarp();
  for (int k = (int)lane; k < K; k += 32) swp(a, (int)m.listA[k], (int)m.listB[k]);
  __syncwarp();
  return min(cutA, cutB);
}

template <int CAP>
struct WarpSortMem {
  u64 a[CAP];      // the range, (h << 32 | index in the group)
  u64 b[CAP];      // listA / listB during the partitions, the sorted range at the end
  u32 gew[CAP / 32 + 1], lew[CAP / 32 + 1], gep[CAP / 32 + 1], les[CAP / 32 + 1];
  u32 bounds[CAP / 32 + 1];  // bit p: a final partition range starts at p
};

// std::sort of m.a[0..n) into m.b[0..n).  depth < 0: a whole group (depth limit 2*lg n); otherwise a range that the
// giant-group kernel split off, with the depth limit it had reached.
template <int CAP>
__device__ void warp_std_sort(WarpSortMem<CAP> &m, int n, u32 lane, int depth0) {
  u64 *a = m.a;
  SortPtrs sp_{m.a, reinterpret_cast<u32 *>(m.b), reinterpret_cast<u32 *>(m.b) + CAP, m.gew, m.lew, m.gep, m.les};
  for (int c = (int)lane; c <= (n >> 5); c += 32) m.bounds[c] = c == 0 ? 1u : 0u;
  __syncwarp();
  if (n > 16) {
    int st_first[40], st_last[40], st_depth[40];  // explicit recursion: the larger side is stacked, depth <= lg n
    if (depth0 < 0) {
      int lg = 0;
      for (int t = n; t > 1; t >>= 1) ++lg;
      depth0 = 2 * lg;
    }
    int sp = 1;
    st_first[0] = 0, st_last[0] = n, st_depth[0] = depth0;
    while (sp > 0) {
      --sp;
      int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
      while (last - first > 16) {
        if (depth == 0) {  // __partial_sort(first, last, last): heap sort, sequential (adversarial inputs only)
          if (lane == 0) {
            SubArray sub{a + first};
            dev_heap_sort(sub, last - first);
          }
          __syncwarp();
          for (int p = first + (int)lane; p < last; p += 32) atomicOr(&m.bounds[p >> 5], 1u << (p & 31));  // keep its order
          __syncwarp();
          break;
        }
        --depth;
        const int cut = warp_partition(sp_, first, last, lane);
        if (lane == 0) atomicOr(&m.bounds[cut >> 5], 1u << (cut & 31));
        __syncwarp();
        // __introsort_loop(cut, last, depth) and last = cut: the two sides are independent, keep the smaller one
        if (cut - first < last - cut) {
          if (last - cut > 16 && sp < 40) st_first[sp] = cut, st_last[sp] = last, st_depth[sp] = depth, ++sp;
          last = cut;
        } else {
          if (cut - first > 16 && sp < 40) st_first[sp] = first, st_last[sp] = cut, st_depth[sp] = depth, ++sp;
          first = cut;
        }
      }
    }
  }
  // __final_insertion_sort: stable by h inside every final range
  for (int base = 0; base < n; base += 32) {
    const int p = base + (int)lane;
    if (p < n) {
      const int wd = p >> 5;
      const u32 own = m.bounds[wd] & (0xFFFFFFFFu >> (31 - (p & 31)));
      int rs;  // ranges that are not kept verbatim have <= 16 elements: the start is in this word or the previous one
      if (own) rs = (wd << 5) + 31 - __clz(own);
      else rs = ((wd - 1) << 5) + 31 - __clz(m.bounds[wd - 1]);
      const u64 v = a[p];
      const u32 h = (u32)(v >> 32);
      int before = 0, q = rs;
      for (; q < p; ++q) before += ((u32)(a[q] >> 32) <= h) ? 1 : 0;          // earlier: precedes when <=
      for (q = p + 1; q < n && !((m.bounds[q >> 5] >> (q & 31)) & 1u); ++q)  // later, same range: when <
        before += ((u32)(a[q] >> 32) < h) ? 1 : 0;
      m.b[rs + before] = v;
    }
  }
  __syncwarp();
}

// members of the group [start, end) of the gid-sorted list
__device__ __forceinline__ u32 group_end(const OrderArgs &a, u32 start, u32 lane, u32 stop_after) {
  const u32 g = a.sgid[start];
  u32 end = start;
  for (;;) {
    const u32 idx = end + lane;
    const bool same = idx < a.m && a.sgid[idx] == g;
    const u32 bal = __ballot_sync(0xFFFFFFFFu, same);
    end += __popc(bal);  // sorted: the matching lanes are a prefix
    if (bal != 0xFFFFFFFFu || end - start > stop_after) break;
  }
  return end;
}

// output lines gstart+first .. of a group from the sorted (h, member index) words
__device__ __forceinline__ void emit_sorted(const OrderArgs &a, const u64 *sorted, u32 gstart, u32 first, u32 n, u32 g, u32 lane) {
  for (u32 t = lane; t < n; t += 32) {
    u32 r, h, fidx, ident;
    load_member(a, gstart + (u32)sorted[t], r, h, fidx, ident);
    const u32 j = gstart + first + t;
    a.out_order[j] = fidx;
    a.out_gid[j] = g + a.gid_base;
    a.out_repval[j] = (u8)(first + t == 0 ? 1 : 2);  // commonFunctions.cpp:106-115
    a.out_identity[j] = __uint_as_float(ident);
  }
}

// one warp per group of worklist LIST: up to CAP members here, larger ones are passed on to the next worklist
template <int CAP, int WARPS, int LIST>
__global__ void __launch_bounds__(WARPS * 32) k_groupsort_warp(OrderArgs a) {
  extern __shared__ __align__(16) unsigned char gw_smem[];
  const u32 lane = threadIdx.x & 31;
  WarpSortMem<CAP> &mem = reinterpret_cast<WarpSortMem<CAP> *>(gw_smem)[threadIdx.x >> 5];
  u32 *const count_in = a.work_count + 2 * LIST, *const count_out = a.work_count + 2 * (LIST + 1);
  const u32 *const list_in = a.worklist[LIST];
  u32 *const list_out = a.worklist[LIST + 1];
  const u32 nseg = min(*count_in, a.work_cap[LIST]);
  for (;;) {
    u32 seg = 0;
    if (lane == 0) seg = atomicAdd(count_in + 1, 1u);
    seg = __shfl_sync(0xFFFFFFFFu, seg, 0);
    if (seg >= nseg) return;
    const u32 start = list_in[seg];
    const u32 n = group_end(a, start, lane, CAP) - start;
    if (n > (u32)CAP) {
      if (lane == 0) {
        const u32 slot = atomicAdd(count_out, 1u);
        if (slot < a.work_cap[LIST + 1]) list_out[slot] = start;
        else atomicOr(a.err, ERR_WORKLIST);
      }
      continue;
    }
    for (u32 t = lane; t < n; t += 32) {
      u32 r, h, fidx, ident;
      load_member(a, start + t, r, h, fidx, ident);
      mem.a[t] = ((u64)h << 32) | t;
    }
    __syncwarp();
    const u64 *sorted = mem.a;
    if (a.do_sort) {
      warp_std_sort(mem, (int)n, lane, -1);
      sorted = mem.b;
    }
    emit_sorted(a, sorted, start, 0, n, a.sgid[start], lane);
    __syncwarp();
  }
}

// Groups of more than 1024 members (repeat families of 10^4 and more fragments).  One warp runs the top of the
// introsort recursion on the group in global memory — the same partition routine, its lists and masks in global
// scratch — and hands every range of <= 1024 elements, with the depth limit reached there, to k_rangesort_warp.
constexpr u32 RANGE_FROZEN = 0x80000000u;  // the range is final as it stands (heap-sorted, or do_sort == 0)

__device__ __forceinline__ void push_range(const OrderArgs &a, u32 gstart, u32 first, u32 last, u32 depth, u32 lane) {
  if (lane == 0) {
    const u32 slot = atomicAdd(a.work_count + 6, 1u);
    if (slot < a.range_cap) a.ranges[slot] = make_uint4(gstart, first, last, depth);
    else atomicOr(a.err, ERR_WORKLIST);
  }
}

template <int LEAF>
__global__ void __launch_bounds__(128) k_giant_split(OrderArgs a) {
  const u32 lane = threadIdx.x & 31;
  const u32 nseg = min(a.work_count[4], a.work_cap[2]);
  for (;;) {
    u32 seg = 0;
    if (lane == 0) seg = atomicAdd(a.work_count + 5, 1u);
    seg = __shfl_sync(0xFFFFFFFFu, seg, 0);
    if (seg >= nseg) return;
    const u32 start = a.worklist[2][seg];
    const u32 n = group_end(a, start, lane, 0xFFFFFFFFu) - start;
    u64 *arr = a.packed + start;
    for (u32 t = lane; t < n; t += 32) {
      u32 r, h, fidx, ident;
      load_member(a, start + t, r, h, fidx, ident);
      arr[t] = ((u64)h << 32) | t;
    }
    __syncwarp();
    if (!a.do_sort) {
      push_range(a, start, 0, n, RANGE_FROZEN, lane);
      continue;
    }
    // scratch of this group: n/16 >= n/32 + 2 chunk slots from start/16 on never reach the next giant group's
    const u64 cs = start >> 4;
    SortPtrs sp_{arr, reinterpret_cast<u32 *>(a.packed2 + start), reinterpret_cast<u32 *>(a.packed2 + start) + n,
                 a.chunk_words + cs, a.chunk_words + a.chunk_stride + cs, a.chunk_words + 2 * a.chunk_stride + cs,
                 a.chunk_words + 3 * a.chunk_stride + cs};
    int st_first[40], st_last[40], st_depth[40];
    int lg = 0;
    for (u32 t = n; t > 1; t >>= 1) ++lg;
    int sp = 1;
    st_first[0] = 0, st_last[0] = (int)n, st_depth[0] = 2 * lg;
    while (sp > 0) {
      --sp;
      int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
      for (;;) {
        if (last - first <= LEAF) {
          push_range(a, start, (u32)first, (u32)last, (u32)depth, lane);
          break;
        }
        if (depth == 0) {
          if (lane == 0) {
            SubArray sub{arr + first};
            dev_heap_sort(sub, last - first);
          }
          __syncwarp();
          push_range(a, start, (u32)first, (u32)last, RANGE_FROZEN, lane);
          break;
        }
        --depth;
        const int cut = warp_partition(sp_, first, last, lane);
        // the larger side is stacked (or handed over when it is a leaf), the smaller one continues
        int of, ol;
        if (cut - first < last - cut) of = cut, ol = last, last = cut;
        else of = first, ol = cut, first = cut;
        if (ol - of <= LEAF) push_range(a, start, (u32)of, (u32)ol, (u32)depth, lane);
        else if (sp < 40) st_first[sp] = of, st_last[sp] = ol, st_depth[sp] = depth, ++sp;
      }
    }
  }
}

// one warp per range handed over by k_giant_split: the rest of its introsort recursion, the final insertion sort,
// and the output lines of the range
template <int CAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_rangesort_warp(OrderArgs a) {
  extern __shared__ __align__(16) unsigned char gw_smem[];
  const u32 lane = threadIdx.x & 31;
  WarpSortMem<CAP> &mem = reinterpret_cast<WarpSortMem<CAP> *>(gw_smem)[threadIdx.x >> 5];
  const u32 nseg = min(a.work_count[6], a.range_cap);
  for (;;) {
    u32 seg = 0;
    if (lane == 0) seg = atomicAdd(a.work_count + 7, 1u);
    seg = __shfl_sync(0xFFFFFFFFu, seg, 0);
    if (seg >= nseg) return;
    const uint4 it = a.ranges[seg];
    const u32 gstart = it.x, first = it.y, n = it.z - it.y;
    const u32 g = a.sgid[gstart];
    const u64 *src = a.packed + gstart + first;
    if (it.w == RANGE_FROZEN || n > (u32)CAP) {  // (n > CAP only for frozen ranges)
      emit_sorted(a, src, gstart, first, n, g, lane);
      continue;
    }
    for (u32 t = lane; t < n; t += 32) mem.a[t] = src[t];
    __syncwarp();
    warp_std_sort(mem, (int)n, lane, (int)it.w);
    emit_sorted(a, mem.b, gstart, first, n, g, lane);
    __syncwarp();
  }
}

constexpr int GW0_WARPS = 8, GW1_WARPS = 4;

// scratch of the order stage: packed (h, index) words of the giant groups, their lists, chunk words, range and group lists
u64 order_scratch_bytes(u64 m) {
  const u64 m1 = m ? m : 1;
  return m1 * 8 + m1 * 8 + (m1 / 8 + 64) * 16 + 4 * (m1 / 16 + 2) * 4 + (m1 / 16 + 2 + m1 / 128 + 2 + m1 / 1024 + 2) * 4 + 256;
}
void order_carve(OrderArgs &a, void *scratch, u64 m) {
  const u64 m1 = m ? m : 1;
  u8 *p = (u8 *)scratch;
  a.packed = (u64 *)p, p += m1 * 8;
  a.packed2 = (u64 *)p, p += m1 * 8;
  a.ranges = (uint4 *)p, a.range_cap = (u32)(m1 / 8 + 64), p += (u64)a.range_cap * 16;
  a.chunk_stride = m1 / 16 + 2;
  a.chunk_words = (u32 *)p, p += 4 * a.chunk_stride * 4;
  a.work_cap[0] = (u32)(m1 / 16 + 2), a.work_cap[1] = (u32)(m1 / 128 + 2), a.work_cap[2] = (u32)(m1 / 1024 + 2);
  a.worklist[0] = (u32 *)p, p += (u64)a.work_cap[0] * 4;
  a.worklist[1] = (u32 *)p, p += (u64)a.work_cap[1] * 4;
  a.worklist[2] = (u32 *)p;
}

cudaError_t order_init_device() {
  const int smem1 = GW1_WARPS * (int)sizeof(WarpSortMem<GS_WARP_CAP>);
  cudaError_t e = cudaFuncSetAttribute(k_groupsort_warp<GS_WARP_CAP, GW1_WARPS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_rangesort_warp<GS_WARP_CAP, GW1_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
}

int launch_order(const OrderArgs &a, cudaStream_t st) {
  if (a.m == 0) return 0;
  const u32 m = a.m;
  cudaMemsetAsync(a.work_count, 0, 8 * sizeof(u32), st);
  {
    KScope ks(KID_GSORT_SMALL, st, m);
    k_order_tile<<<(m + OT_HEADS - 1) / OT_HEADS, OT_TILE, 0, st>>>(a);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem0 = GW0_WARPS * (int)sizeof(WarpSortMem<GS_WARP_CAP0>), smem1 = GW1_WARPS * (int)sizeof(WarpSortMem<GS_WARP_CAP>);
  {
    KScope ks(KID_GSORT_WARP, st, 0);
    k_groupsort_warp<GS_WARP_CAP0, GW0_WARPS, 0><<<sms * 6, GW0_WARPS * 32, smem0, st>>>(a);
    k_groupsort_warp<GS_WARP_CAP, GW1_WARPS, 1><<<sms * 3, GW1_WARPS * 32, smem1, st>>>(a);
  }
  KScope ks(KID_GSORT_LARGE, st, 0);
  k_giant_split<GS_WARP_CAP><<<sms * 4, 128, 0, st>>>(a);
  k_rangesort_warp<GS_WARP_CAP, GW1_WARPS><<<sms * 3, GW1_WARPS * 32, smem1, st>>>(a);
  return 5;
}

}  // namespace rk
