// K3 — the X pass and the Y pass of generate_fragment_groups, and the kernel that prepares their input.
//
// Reference: /root/reference/src/commonFunctions.cpp:51-77 (the greedy loop) over
// /root/reference/src/SequenceOcupationList.cpp:20-31 (deviation), :33-91 (get_associated_group), :93-96 (insert).
//
// The loop is sequential, but it decomposes exactly (validated against the reference):
//   X pass  per strand class, in processing order: xo[f] = best earlier X entry or none; f becomes an X entry
//           iff it found none (commonFunctions.cpp:55-61 never inserts a matched fragment into solx, :67,:75 do).
//   Y pass  fragments with an X match insert unconditionally (:59); the others query (:63) and insert iff
//           they found nothing (:76).
// A query touches its own center/100 bucket and at most one neighbour (only for center%100 in {0,1} -> previous,
// {98,99} and center < max_index -> next), so buckets joined by such probes form short runs ("super-buckets",
// the link bits raised by K1).  Fragments arrive here stably sorted by (strand class, first bucket of their
// run): every run is a contiguous segment in processing order and segments are independent of one another.
//
//   tier 1: one thread per segment of <= 32 fragments (the typical segment holds 3); inserted entries are a
//           32-bit mask over the segment, scanned newest first.
//   tier 2: one warp per longer segment: 32 queries at a time are scored against the entry list (uniform
//           loads, each lane its own query), then the insertions inside the chunk are replayed in order with
//           ballots; work per query is O(entries), not O(segment).
// deviation() is evaluated in binary64 with explicit round-to-nearest intrinsics (no FMA contraction), the
// reference's operation order, because the comparison of two scores decides group membership.
#include "rk_common.cuh"

namespace rk {

constexpr int T1_MAX = 32;
constexpr u32 NO_BUCKET = 0xFFFFFFFFu;

__device__ __forceinline__ u32 run_start(const u32 *__restrict__ bm, u32 k) {
  // largest k' <= k whose link bit is clear (bit 0 of every strand class is never set)
  u32 w = k >> 5;
  u32 m = 0xFFFFFFFFu >> (31 - (k & 31));
  for (;;) {
    const u32 z = ~bm[w] & m;
    if (z) return (w << 5) + (31 - __clz(z));
    if (w == 0) return 0;
    --w;
    m = 0xFFFFFFFFu;
  }
}

__global__ void __launch_bounds__(256)
    k_keys(const u32 *__restrict__ fidx_r, u32 m, Geometry g, const u32 *__restrict__ xs, const u32 *__restrict__ ys,
           const u32 *__restrict__ len, const u8 *__restrict__ flags, const u32 *__restrict__ link_x,
           const u32 *__restrict__ link_y, u32 *__restrict__ cx_r, u32 *__restrict__ cy_r, u32 *__restrict__ len_r,
           u32 *__restrict__ ys_r, u32 *__restrict__ kx, u32 *__restrict__ ky) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const u32 f = fidx_r[i];
  const u32 x = xs[f], y = ys[f], l = len[f];
  const u32 sc = flags[f] & FL_REVERSE;
  const u32 cx = x + l / 2, cy = y + l / 2;  // commonFunctions.cpp:55,59
  cx_r[i] = cx;
  cy_r[i] = cy;
  len_r[i] = l;
  ys_r[i] = y;
  kx[i] = run_start(link_x, sc * g.nbx + cx / DIVISOR);
  ky[i] = run_start(link_y, sc * g.nby + cy / DIVISOR);
}

int launch_keys(const u32 *fidx_r, u32 m, Geometry g, const u32 *xs, const u32 *ys, const u32 *len, const u8 *flags,
                const u32 *link_x, const u32 *link_y, u32 *cx_r, u32 *cy_r, u32 *len_r, u32 *ys_r, u32 *kx, u32 *ky,
                cudaStream_t st) {
  if (m == 0) return 0;
  KScope ks(KID_KEYS, st);
  k_keys<<<(m + 255) / 256, 256, 0, st>>>(fidx_r, m, g, xs, ys, len, flags, link_x, link_y, cx_r, cy_r, len_r, ys_r, kx, ky);
  return 1;
}

// SequenceOcupationList::deviation (SequenceOcupationList.cpp:20-31).  t_len = length*len_ratio and
// t_pos = length*pos_ratio are the query's (the relation is asymmetric).
__device__ __forceinline__ double deviation(u32 ec, u32 el, u32 c, u32 len, double t_len, double t_pos) {
  const u32 dif_len = len > el ? len - el : el - len;
  const double sim_len = __dadd_rn(-fabs(__ddiv_rn((double)dif_len, t_len)), 1.0);
  if (sim_len < 0) return 0.0;
  const u32 dif_cen = c > ec ? c - ec : ec - c;
  const double sim_pos = __dadd_rn(-fabs(__ddiv_rn((double)dif_cen, t_pos)), 1.0);
  if (sim_pos < 0) return 0.0;
  return __dadd_rn(__dmul_rn(sim_len, 0.4), __dmul_rn(sim_pos, 0.6));
}

// the one neighbour bucket get_associated_group can reach from center c (:47-89), or NO_BUCKET
__device__ __forceinline__ u32 neighbour_bucket(u32 c, u32 max_index) {
  const u32 b = c / DIVISOR;
  if (probes_prev(c)) return b - 1;
  if (probes_next(c, max_index)) return b + 1;
  return NO_BUCKET;
}

__global__ void __launch_bounds__(128) k_match_small(MatchArgs a) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.m) return;
  const u32 key = a.skey[i];
  if (i > 0 && a.skey[i - 1] == key) return;  // not the head of a segment
  u32 n = 1;
  while (n <= (u32)T1_MAX && i + n < a.m && a.skey[i + n] == key) ++n;
  if (n > (u32)T1_MAX) {
    const u32 slot = atomicAdd(a.work_count, 1u);
    if (slot < a.work_cap) a.worklist[slot] = i;
    else atomicOr(a.err, ERR_WORKLIST);
    return;
  }
  u32 ec[T1_MAX], el[T1_MAX];
  u32 inserted = 0;
  for (u32 j = 0; j < n; ++j) {
    const u32 r = a.srank[i + j];
    const u32 c = a.c_r[r], len = a.len_r[r];
    ec[j] = c;
    el[j] = len;
    if (a.is_y && a.parent[r] != RK_NONE32) {  // X-matched: Y-insert without a query (commonFunctions.cpp:59)
      inserted |= 1u << j;
      continue;
    }
    const u32 b = c / DIVISOR;
    const u32 nbk = neighbour_bucket(c, a.max_index);
    const double t_len = __dmul_rn((double)len, a.len_ratio);
    const double t_pos = __dmul_rn((double)len, a.pos_ratio);
    double best_sc = 0.0;
    int best = -1;
    bool best_own = false;
    // newest entry first; the own bucket is scanned before the neighbour, `>` is strict (:40)
    for (u32 msk = inserted; msk;) {
      const int k = 31 - __clz(msk);
      msk &= ~(1u << k);
      const u32 bk = ec[k] / DIVISOR;
      const bool own = bk == b;
      if (!own && bk != nbk) continue;
      const double sc = deviation(ec[k], el[k], c, len, t_len, t_pos);
      if (sc > best_sc || (sc == best_sc && best >= 0 && own && !best_own)) {
        best_sc = sc;
        best = k;
        best_own = own;
      }
    }
    if (best >= 0) {
      a.parent[r] = a.srank[i + best];
    } else {
      inserted |= 1u << j;
      if (!a.is_y) a.parent[r] = RK_NONE32;
    }
  }
}

__global__ void __launch_bounds__(128) k_match_long(MatchArgs a) {
  const u32 lane = threadIdx.x & 31;
  const u32 nseg = min(*a.work_count, a.work_cap);
  for (;;) {
    u32 seg = 0;
    if (lane == 0) seg = atomicAdd(a.work_count + 1, 1u);
    seg = __shfl_sync(0xFFFFFFFFu, seg, 0);
    if (seg >= nseg) return;
    const u32 start = a.worklist[seg];
    const u32 key = a.skey[start];
    u32 end = start;
    for (;;) {
      const u32 idx = end + lane;
      const bool same = idx < a.m && a.skey[idx] == key;
      const u32 bal = __ballot_sync(0xFFFFFFFFu, same);
      end += __popc(bal);  // sorted keys: the matching lanes are a prefix
      if (bal != 0xFFFFFFFFu) break;
    }
    u32 n_ent = 0;
    for (u32 base = start; base < end; base += 32) {
      const u32 q = base + lane;
      const bool valid = q < end;
      u32 r = 0, c = 0, len = 0;
      bool xm = false;
      if (valid) {
        r = a.srank[q];
        c = a.c_r[r];
        len = a.len_r[r];
        xm = a.is_y && a.parent[r] != RK_NONE32;
      }
      const bool needq = valid && !xm;
      const u32 b = c / DIVISOR;
      const u32 nbk = neighbour_bucket(c, a.max_index);
      const double t_len = __dmul_rn((double)len, a.len_ratio);
      const double t_pos = __dmul_rn((double)len, a.pos_ratio);
      double best_sc = 0.0;
      u32 best = RK_NONE32;
      bool best_own = false;
      // phase A: entries inserted before this chunk, newest first (uniform loads)
      for (u32 t = n_ent; t-- > 0;) {
        const u32 e_c = a.ent_c[start + t], e_l = a.ent_len[start + t];
        if (needq) {
          const u32 bk = e_c / DIVISOR;
          const bool own = bk == b;
          if (own || bk == nbk) {
            const double sc = deviation(e_c, e_l, c, len, t_len, t_pos);
            if (sc > best_sc || (sc == best_sc && best != RK_NONE32 && own && !best_own)) {
              best_sc = sc;
              best = a.ent_rank[start + t];
              best_own = own;
            }
          }
        }
      }
      // phase B: replay the insertions of this chunk in order
      u32 pending = __ballot_sync(0xFFFFFFFFu, valid);
      while (pending) {
        const u32 wants = __ballot_sync(0xFFFFFFFFu, valid && (xm || best == RK_NONE32)) & pending;
        if (!wants) break;  // everything still pending has a match and nothing is inserted before it
        const int L = __ffs(wants) - 1;
        const u32 Lc = __shfl_sync(0xFFFFFFFFu, c, L);
        const u32 Ll = __shfl_sync(0xFFFFFFFFu, len, L);
        const u32 Lr = __shfl_sync(0xFFFFFFFFu, r, L);
        if ((int)lane == L) {
          a.ent_rank[start + n_ent] = r;
          a.ent_c[start + n_ent] = c;
          a.ent_len[start + n_ent] = len;
        }
        ++n_ent;
        if ((int)lane > L && needq) {
          // the new entry is the newest: it is scanned before every older entry of its bucket class
          const u32 bk = Lc / DIVISOR;
          const bool own = bk == b;
          if (own || bk == nbk) {
            const double sc = deviation(Lc, Ll, c, len, t_len, t_pos);
            if (sc > best_sc || (sc == best_sc && best != RK_NONE32 && (own || !best_own))) {
              best_sc = sc;
              best = Lr;
              best_own = own;
            }
          }
        }
        pending &= ~((2u << L) - 1u);  // lanes <= L are final
      }
      if (needq) {
        if (best != RK_NONE32) a.parent[r] = best;
        else if (!a.is_y) a.parent[r] = RK_NONE32;
      }
      __syncwarp();  // entry stores of this chunk are read by every lane in the next one
    }
  }
}

int launch_match(const MatchArgs &a, cudaStream_t st) {
  if (a.m == 0) return 0;
  cudaMemsetAsync(a.work_count, 0, 2 * sizeof(u32), st);
  {
    KScope ks(KID_MATCH_SMALL, st);
    k_match_small<<<(a.m + 127) / 128, 128, 0, st>>>(a);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  KScope ks(KID_MATCH_LONG, st);
  k_match_long<<<sms * 4, 128, 0, st>>>(a);
  return 2;
}

}  // namespace rk
