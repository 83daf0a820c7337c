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
// (h << 32 | rank) words, comparing the high halves only.
#include "rk_common.cuh"

namespace rk {

// ---- K5a ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_hkey(const u32 *__restrict__ k0_r, const u32 *__restrict__ ys_r, u32 m,
                                              u32 *__restrict__ h) {
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
  h[i] = absdiff(ys_r[i], ys_r[lo]);  // commonFunctions.cpp:152-155
}

int launch_hkey(const u32 *k0_r, const u32 *ys_r, u32 m, u32 *h, cudaStream_t st) {
  if (m == 0) return 0;
  KScope ks(KID_HKEY, st, m);
  k_hkey<<<(m + 255) / 256, 256, 0, st>>>(k0_r, ys_r, m, h);
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

// ---- K5b: libstdc++ std::sort, restated over an indexable array of packed (h,rank) words ----------------

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

template <class P>
__device__ __forceinline__ void dev_unguarded_linear_insert(P a, int last) {
  const u64 val = a[last];
  int next = last - 1;
  while (hless(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

template <class P>
__device__ void dev_insertion_sort(P a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (hless(a[i], a[first])) {
      const u64 val = a[i];
      for (int k = i; k > first; --k) a[k] = a[k - 1];  // move_backward
      a[first] = val;
    } else {
      dev_unguarded_linear_insert(a, i);
    }
  }
}

struct SubArray {  // a[first..] view so the heap code can stay zero-based
  u64 *p;
  __device__ __forceinline__ u64 &operator[](int i) const { return p[i]; }
};

template <class P>
__device__ void dev_std_sort(P a, int n) {
  if (n <= 1) return;
  if (n > 16) {
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    // __introsort_loop; the two sides of a partition are independent, so the recursion order is free: keep
    // iterating on the smaller side and stack the larger one (stack depth <= lg n).
    int st_first[40], st_last[40], st_depth[40];
    int sp = 0;
    st_first[0] = 0, st_last[0] = n, st_depth[0] = 2 * lg;
    sp = 1;
    while (sp > 0) {
      --sp;
      int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
      while (last - first > 16) {
        if (depth == 0) {
          SubArray sub{&a[first]};
          dev_heap_sort(sub, last - first);
          break;
        }
        --depth;
        // __unguarded_partition_pivot: median of (first+1, mid, last-1) moved to first
        const int mid = first + (last - first) / 2;
        {
          const int x = first + 1, y = mid, z = last - 1;
          if (hless(a[x], a[y])) {
            if (hless(a[y], a[z])) swp(a, first, y);
            else if (hless(a[x], a[z])) swp(a, first, z);
            else swp(a, first, x);
          } else if (hless(a[x], a[z])) swp(a, first, x);
          else if (hless(a[y], a[z])) swp(a, first, z);
          else swp(a, first, y);
        }
        int lo = first + 1, hi = last;
        const u64 pivot = a[first];
        for (;;) {
          while (hless(a[lo], pivot)) ++lo;
          --hi;
          while (hless(pivot, a[hi])) --hi;
          if (!(lo < hi)) break;
          swp(a, lo, hi);
          ++lo;
        }
        const int cut = lo;
        // left = [first, cut), right = [cut, last), both continue with `depth`
        if (cut - first < last - cut) {
          if (sp < 40) { st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; ++sp; }
          last = cut;
        } else {
          if (sp < 40) { st_first[sp] = first; st_last[sp] = cut; st_depth[sp] = depth; ++sp; }
          first = cut;
        }
      }
    }
    // __final_insertion_sort
    dev_insertion_sort(a, 0, 16);
    for (int i = 16; i != n; ++i) dev_unguarded_linear_insert(a, i);
  } else {
    dev_insertion_sort(a, 0, n);
  }
}

constexpr int GS_SMALL = 64;         // groups up to this size are ordered inside the tile kernel
constexpr int GS_SMEM_ELEMS = 6144;  // larger groups up to this size are sorted in shared memory (48 KB)
constexpr int OT_THREADS = 256;
constexpr int OT_TILE = OT_THREADS + GS_SMALL;  // groups that start in the first 256 positions end before 320

__device__ __forceinline__ void emit_line(const OrderArgs &a, u32 j, u64 pk, u32 g, u32 rep) {
  const u32 r = (u32)pk;
  const u32 f = a.fidx_r[r];
  a.out_order[j] = f;
  a.out_gid[j] = g;
  a.out_repval[j] = (u8)rep;  // commonFunctions.cpp:106-115
  a.out_identity[j] = a.identity_r ? a.identity_r[r] : a.identity_f[f];
}

// K5b+c for groups of <= 64 members (all but a handful): a tile of the gid-sorted list is packed into shared memory as
// (h << 32 | rank), the thread at a group's head orders the group there, and every thread writes the output line
// of its position.  Larger groups go to the worklist of k_groupsort_large.
__global__ void __launch_bounds__(OT_THREADS) k_order_tile(OrderArgs a) {
  __shared__ u64 s_pk[OT_TILE];
  __shared__ u32 s_gid[OT_TILE];
  __shared__ u8 s_rep[OT_TILE];  // repval, 0xFF: not this tile's (continuation of the previous tile's group, or a large group)
  __shared__ u32 s_prev;
  const u32 tid = threadIdx.x;
  const u32 bs = blockIdx.x * OT_THREADS;
  const u32 count = min((u32)OT_TILE, a.m - bs);
  for (u32 e = tid; e < (u32)OT_TILE; e += OT_THREADS) {
    s_rep[e] = 0xFF;
    if (e < count) {
      const u32 j = bs + e;
      const u32 r = a.srank ? a.srank[j] : j;  // direct layout: h/fidx/identity are already in this order
      s_pk[e] = ((u64)a.h[r] << 32) | r;
      s_gid[e] = a.sgid[j];
    }
  }
  if (tid == 0) s_prev = bs ? a.sgid[bs - 1] : 0;
  __syncthreads();
  if (tid < count) {
    const u32 g = s_gid[tid];
    const bool head = tid == 0 ? (bs == 0 || s_prev != g) : s_gid[tid - 1] != g;
    if (head) {
      u32 n = 1;
      while (n <= (u32)GS_SMALL && tid + n < count && s_gid[tid + n] == g) ++n;
      if (n > (u32)GS_SMALL) {  // the tile holds head+64, so this is exact
        const u32 slot = atomicAdd(a.work_count, 1u);
        if (slot < a.work_cap) a.worklist[slot] = bs + tid;
        else atomicOr(a.err, ERR_WORKLIST);
      } else {
        if (n > 1 && a.do_sort) dev_std_sort(s_pk + tid, (int)n);
        s_rep[tid] = n == 1 ? 0 : 1;
        for (u32 p = 1; p < n; ++p) s_rep[tid + p] = 2;
      }
    }
  }
  __syncthreads();
  for (u32 e = tid; e < count; e += OT_THREADS)
    if (s_rep[e] != 0xFF) emit_line(a, bs + e, s_pk[e], s_gid[e], s_rep[e]);
}

// one warp per large group; lane 0 runs the (inherently sequential) introsort, all lanes move the data
__global__ void __launch_bounds__(32) k_groupsort_large(OrderArgs a) {
  __shared__ u64 buf[GS_SMEM_ELEMS];
  const u32 lane = threadIdx.x;
  const u32 nseg = min(*a.work_count, a.work_cap);
  for (;;) {
    u32 seg = 0;
    if (lane == 0) seg = atomicAdd(a.work_count + 1, 1u);
    seg = __shfl_sync(0xFFFFFFFFu, seg, 0);
    if (seg >= nseg) return;
    const u32 start = a.worklist[seg];
    const u32 g = a.sgid[start];
    u32 end = start;
    for (;;) {
      const u32 idx = end + lane;
      const bool same = idx < a.m && a.sgid[idx] == g;
      const u32 bal = __ballot_sync(0xFFFFFFFFu, same);
      end += __popc(bal);
      if (bal != 0xFFFFFFFFu) break;
    }
    const u32 n = end - start;
    u64 *arr = n <= (u32)GS_SMEM_ELEMS ? buf : a.packed + start;  // beyond 48 KB: in the global scratch
    for (u32 t = lane; t < n; t += 32) {
      const u32 r = a.srank ? a.srank[start + t] : start + t;
      arr[t] = ((u64)a.h[r] << 32) | r;
    }
    __syncwarp();
    if (lane == 0 && a.do_sort) dev_std_sort(arr, (int)n);
    __syncwarp();
    for (u32 t = lane; t < n; t += 32) emit_line(a, start + t, arr[t], g, t == 0 ? 1 : 2);
    __syncwarp();
  }
}

int launch_order(const OrderArgs &a, cudaStream_t st) {
  if (a.m == 0) return 0;
  const u32 m = a.m;
  cudaMemsetAsync(a.work_count, 0, 2 * sizeof(u32), st);
  {
    KScope ks(KID_GSORT_SMALL, st, m);
    k_order_tile<<<(m + OT_THREADS - 1) / OT_THREADS, OT_THREADS, 0, st>>>(a);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  KScope ks(KID_GSORT_LARGE, st, 0);
  k_groupsort_large<<<sms * 4, 32, 0, st>>>(a);
  return 2;
}

}  // namespace rk
