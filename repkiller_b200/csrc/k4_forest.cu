// K4 — group membership from the owner forest.
//
// Reference: /root/reference/src/commonFunctions.cpp:56-76.  A fragment that matched an entry joins the group
// that entry points to (agx->push_back / agy->push_back, :58,:66); a fragment that matched nothing founds
// group number efrags_groups.size() (:72-74).  With parent[f] = the fragment that inserted the matched entry,
// the groups are the trees of a forest whose parents always have a smaller processing rank, and
//   gid(f) = number of roots with rank < root(f)                      (creation order, the printed block id).
// Kernels: an exclusive scan of the predicate parent == NONE (group id of every root), then one pointer chase
// per fragment to its root (observed depth <= 13; the hook step itself is the K3 store into parent[]).
#include "rk_common.cuh"
#include "rk_scan.cuh"

namespace rk {

struct LoadIsRoot {
  const u32 *parent;
  __device__ __forceinline__ u32 operator()(u64 i) const { return parent[i] == RK_NONE32 ? 1u : 0u; }
};

__global__ void __launch_bounds__(256) k_chase(const u32 *__restrict__ parent, const u32 *__restrict__ gid_of_root, u32 m,
                                               u32 *__restrict__ gid_rank, const u32 *__restrict__ total, u32 *n_groups,
                                               u32 lo, HistOut ho) {
  // ranks lo .. lo+m-1 are resolved (lo = 0, m = all on a single GPU; a rank's slice in the multi-GPU stages).
  // The group ids are the keys of the K5b sort: their digit counts are gathered here.
  __shared__ u32 s_h[HIST_PASSES][HIST_RADIX];
  const bool do_hist = ho.ghist != nullptr;
  if (do_hist) {
    hist_zero(s_h);
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_groups = *total;
  for (u64 base = (u64)blockIdx.x * blockDim.x; base < m; base += (u64)gridDim.x * blockDim.x) {
    const u32 t = (u32)base + threadIdx.x;
    const bool valid = t < m;
    u32 gidv = 0;
    if (valid) {
      u32 r = lo + t;
      for (;;) {
        const u32 p = parent[r];
        if (p == RK_NONE32) break;
        r = p;  // p < r always: terminates
      }
      gidv = gid_of_root[r];
      gid_rank[t] = gidv;
    }
    if (do_hist) hist_add(s_h, gidv, valid, ho);
  }
  if (do_hist) {
    __syncthreads();
    hist_flush(s_h, ho);
  }
}

u64 forest_work_bytes(u32 m) { return ((u64)m + scan_work_words(m)) * 4 + 256; }

int launch_forest(const u32 *parent, u32 m, u32 *gid_rank, u32 *n_groups, void *work, cudaStream_t st, u32 lo, u32 cnt,
                  HistOut hist) {
  if (cnt == 0xFFFFFFFFu) cnt = m;
  if (m == 0) {
    cudaMemsetAsync(n_groups, 0, sizeof(u32), st);
    return 0;
  }
  u32 *gid_of_root = reinterpret_cast<u32 *>(work);
  u32 *bsum = gid_of_root + m;
  int launches = exclusive_scan_u32(LoadIsRoot{parent}, gid_of_root, m, bsum, st);
  const u32 nb = (u32)(((u64)m + SCAN_CHUNK - 1) / SCAN_CHUNK);
  KScope ks(KID_CHASE, st, cnt);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u32 blocks = (cnt + 255) / 256 + (cnt == 0);
  if (hist.ghist && blocks > (u32)sms * 8) blocks = (u32)sms * 8;
  k_chase<<<blocks, 256, 0, st>>>(parent, gid_of_root, cnt, gid_rank, bsum + nb, n_groups, lo, hist);
  return launches + 1;
}

}  // namespace rk
