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
                                               u32 lo) {
  // ranks lo .. lo+m-1 are resolved (lo = 0, m = all on a single GPU; a rank's slice in the multi-GPU stages)
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) *n_groups = *total;
  if (t >= m) return;
  const u32 i = lo + t;
  u32 r = i;
  for (;;) {
    const u32 p = parent[r];
    if (p == RK_NONE32) break;
    r = p;  // p < r always: terminates
  }
  gid_rank[t] = gid_of_root[r];
}

u64 forest_work_bytes(u32 m) { return ((u64)m + scan_work_words(m)) * 4 + 256; }

int launch_forest(const u32 *parent, u32 m, u32 *gid_rank, u32 *n_groups, void *work, cudaStream_t st, u32 lo, u32 cnt) {
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
  k_chase<<<(cnt + 255) / 256 + (cnt == 0), 256, 0, st>>>(parent, gid_of_root, cnt, gid_rank, bsum + nb, n_groups, lo);
  return launches + 1;
}

}  // namespace rk
