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
//   tier 1: segments of <= 32 fragments (the typical segment holds 3; a repeat family about a dozen): independent
//           warps on 64-position windows — candidate masks from a pair-balanced search, then the insertion status of
//           every member in a few ballot rounds, then the owners (see k_match_small).
//   tier 2: one warp per longer segment: 32 queries at a time are scored against the entry list (uniform
//           loads, each lane its own query), then the insertions inside the chunk are replayed in order with
//           ballots; work per query is O(entries), not O(segment).
// deviation() is evaluated in binary64 with explicit round-to-nearest intrinsics (no FMA contraction), the
// reference's operation order, because the comparison of two scores decides group membership.
#include "rk_common.cuh"

namespace rk {

constexpr int T1_MAX = 32;
constexpr u32 NO_BUCKET = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
    k_keys(const u32 *__restrict__ fidx_r, u32 m, Geometry g, const uint4 *__restrict__ rec4, const u32 *__restrict__ link_x,
           const u32 *__restrict__ link_y, uint2 *__restrict__ xl_r, uint2 *__restrict__ yl_r, u32 *__restrict__ ys_r,
           u32 *__restrict__ kx, u32 *__restrict__ ky, float *__restrict__ identity_r, HistOut hx, HistOut hy,
           u32 *__restrict__ gfidx_r, u32 own_bit, const uint2 *__restrict__ rec6) {
  // the digit counts of the two sort keys are gathered here (the keys would otherwise be read again by each sort)
  __shared__ u32 s_hx[HIST_PASSES][HIST_RADIX], s_hy[HIST_PASSES][HIST_RADIX];
  const bool do_hist = hx.ghist != nullptr;
  if (do_hist) {
    hist_zero(s_hx);
    hist_zero(s_hy);
    __syncthreads();
  }
  // warp-uniform loop bounds: every lane takes part in the votes of hist_add
  for (u64 base = (u64)blockIdx.x * blockDim.x; base < m; base += (u64)gridDim.x * blockDim.x) {
    const u32 i = (u32)base + threadIdx.x;
    const bool valid = i < m;
    u32 kxv = 0, kyv = 0;
    if (valid) {
      // one 32-byte gather (one sector): {xStart, yStart, length, flags} {identity bits, file index, 0, 0} — or, multi-GPU,
      // the 24-byte row that arrived in exchange 1 (three 8-byte words, one or two sectors)
      uint4 rec, rec1;
      if (rec6) {
        const uint2 *src = rec6 + 3 * (u64)fidx_r[i];
        const uint2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        rec = make_uint4(a.x, a.y, b.x, b.y);
        rec1 = make_uint4(c.x, c.y, 0u, 0u);
      } else {
        const uint4 *src = rec4 + 2 * (u64)fidx_r[i];
        rec = ldg_gather_u4(src), rec1 = ldg_gather_u4(src + 1);
      }
      identity_r[i] = __uint_as_float(rec1.x);
      if (gfidx_r) gfidx_r[i] = rec1.y;
      const u32 x = rec.x, y = rec.y, l = rec.z;
      const u32 sc = rec.w & FL_REVERSE;
      const u32 cx = x + l / 2, cy = y + l / 2;  // commonFunctions.cpp:55,59
      xl_r[i] = make_uint2(cx, l);
      yl_r[i] = make_uint2(cy, l);
      ys_r[i] = y;
      kxv = run_start(link_x, sc * g.nbx + cx / DIVISOR);
      kyv = run_start(link_y, sc * g.nby + cy / DIVISOR);
      if (own_bit) kxv = 2 * kxv + 1;
      kx[i] = kxv;
      ky[i] = kyv;
    }
    if (do_hist) {
      hist_add(s_hx, kxv, valid, hx);
      hist_add(s_hy, kyv, valid, hy);
    }
  }
  if (do_hist) {
    __syncthreads();
    hist_flush(s_hx, hx);
    hist_flush(s_hy, hy);
  }
}

int launch_keys(const u32 *fidx_r, u32 m, Geometry g, const uint4 *rec4, const u32 *link_x, const u32 *link_y, uint2 *xl_r,
                uint2 *yl_r, u32 *ys_r, u32 *kx, u32 *ky, float *identity_r, cudaStream_t st, HistOut hist_x, HistOut hist_y,
                u32 *gfidx_r, u32 own_bit, const uint2 *rec6) {
  if (m == 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u32 blocks = (m + 255) / 256;
  if (hist_x.ghist && blocks > (u32)sms * 8) blocks = (u32)sms * 8;  // few CTAs: few histogram flushes
  KScope ks(KID_KEYS, st, m);
  k_keys<<<blocks, 256, 0, st>>>(fidx_r, m, g, rec4, link_x, link_y, xl_r, yl_r, ys_r, kx, ky, identity_r, hist_x, hist_y, gfidx_r, own_bit, rec6);
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

// ---- tier 1: candidate masks -------------------------------------------------------------------------------
// Whether entry k is a candidate for query j (deviation(k, j) > 0) does not depend on what was inserted, only
// on the two fragments.  So the sequential part of the greedy loop shrinks to bit operations:
//   cand[j]  = { k < j in j's segment, bucket-compatible, deviation > 0 }     computed for all j in parallel
//   in order: hit = cand[j] & inserted;  hit == 0 -> j is inserted;  one bit -> that entry is the owner;
//             several bits (2 % of the queries) -> exact scores decide (greatest, own bucket first, newest first).
// deviation > 0  <=>  q_len <= 1 and q_pos <= 1 and not both == 1, with q = fl(dif / fl(length*ratio)).  The
// comparison of the correctly rounded quotient with 1 is decided without dividing whenever dif is not within
// 2^-40 (relative) of the threshold; inside that band the division is done.

constexpr double BAND_LO = 1.0 - 0x1p-40;
constexpr double BAND_HI = 1.0 + 0x1p-40;

// Integer thresholds of one similarity term for one query (t = fl(length*ratio) > 0):
//   d <  pass  =>  fl(d/t) < 1  (term > 0)          pass = ceil(t*(1-2^-40)), clamped to u32
//   d >  rej   =>  fl(d/t) > 1  (term < 0)          rej  = floor(t*(1+2^-40)), clamped to u32
// anything in between is decided by the real division.  Clamping only widens the band.
struct Thresh {
  u32 pass, rej;
  double t;
};
__device__ __forceinline__ Thresh make_thresh(u32 len, double ratio) {
  Thresh th;
  th.t = __dmul_rn((double)len, ratio);
  th.pass = __double2uint_ru(fmin(th.t * BAND_LO, 4294967295.0));
  th.rej = __double2uint_rd(fmin(th.t * BAND_HI, 4294967295.0));
  return th;
}
// 0: fl(d/t) > 1 (similarity < 0) ; 1: fl(d/t) < 1 ; 2: fl(d/t) == 1 (similarity exactly 0)
__device__ __forceinline__ int quotient_vs_one(u32 d, const Thresh &th) {
  if (d < th.pass) return 1;
  if (d > th.rej) return 0;
  const double q = __ddiv_rn((double)d, th.t);
  return q > 1.0 ? 0 : (q == 1.0 ? 2 : 1);
}

// srank[i] is the fragment's index in the pass's working list (single GPU: the processing rank; multi-GPU: the position
// in [own fragments ++ halo] / in arrival order); centers/lengths are gathered from the list-ordered cl_r and the owner is
// stored at parent[index].
__device__ __forceinline__ void load_elem(const MatchArgs &a, u32 pos, u32 &r, u32 &c, u32 &len, bool &xm) {
  r = a.srank[pos];
  const uint2 cl = a.cl_r[r];
  c = cl.x;
  len = cl.y;
  // X-matched: Y-insert without a query (commonFunctions.cpp:59); one bit per rank, set by the X pass (the 1.25 MB
  // map of 10M fragments stays in L2, a gather of parent[r] would cost a DRAM sector)
  xm = a.is_y && (a.xm_bytes ? a.xm_bytes[r] != 0 : (((a.xm_bits[r >> 5] >> (r & 31)) & 1u) != 0));
}
__device__ __forceinline__ void store_owner(const MatchArgs &a, u32 r, u32 v) {
  a.parent[r] = v;
  if (!a.is_y && v != RK_NONE32) atomicOr(&a.xm_bits[r >> 5], 1u << (r & 31));
}
// "no match" is stored by the X pass only: in the Y pass parent[] already holds the X result (multi-GPU: NONE).
__device__ __forceinline__ bool stores_none(const MatchArgs &a) { return !a.is_y; }

constexpr int MT_HEADS = 256;           // a CTA stages 256 positions + 32 of halo; warp w owns the segments whose head
constexpr int MT_TILE = MT_HEADS + 32;  // lies in positions [32w, 32w+32) — they end before 32w+64 (the warp's window)
#ifndef RK_MT_MINB
#define RK_MT_MINB 6
#endif

// The rare tails of the candidate search are kept out of line so that the hot loop stays small.
// In-band pair: both differences are within Thresh::rej but at least one is not below Thresh::pass; the correctly
// rounded quotients decide (deviation > 0 <=> both <= 1 and not both == 1).
__device__ __noinline__ bool band_candidate(u32 dl, u32 dc, u32 len, double len_ratio, double pos_ratio) {
  const double ql = __ddiv_rn((double)dl, __dmul_rn((double)len, len_ratio));
  if (ql > 1.0) return false;
  const double qp = __ddiv_rn((double)dc, __dmul_rn((double)len, pos_ratio));
  if (qp > 1.0) return false;
  return !(ql == 1.0 && qp == 1.0);
}

// owner among several inserted candidates: greatest score, own bucket before neighbour, newest first (:40 strict >).
// hit: bit q <-> window position q; returns the window position of the winner
__device__ __noinline__ u32 best_of(unsigned long long hit, const uint2 *s_cl_win, u32 c, u32 len, double len_ratio,
                                    double pos_ratio) {
  const u32 b = c / DIVISOR;
  const double t_len = __dmul_rn((double)len, len_ratio);
  const double t_pos = __dmul_rn((double)len, pos_ratio);
  double best_sc = 0.0;
  int best = -1;
  bool best_own = false;
  while (hit) {
    const int k = 63 - __clzll(hit);
    hit &= ~(1ull << k);
    const uint2 ecl = s_cl_win[k];
    const bool own = ecl.x / DIVISOR == b;
    const double sc = deviation(ecl.x, ecl.y, c, len, t_len, t_pos);
    if (sc > best_sc || (sc == best_sc && best >= 0 && own && !best_own)) {
      best_sc = sc;
      best = k;
      best_own = own;
    }
  }
  return (u32)best;
}

// what a query needs for the candidate test, worked out once per tile position while the tile is staged
struct QueryTh {
  u32 pass_len, rej_len, pass_pos, rej_pos;  // Thresh::pass / Thresh::rej of the two similarity terms
};

// deviation(entry, query) > 0 ?
__device__ __forceinline__ bool is_candidate(uint2 ecl, u32 c, u32 len, u32 b, u32 nbk, const uint4 &th, const MatchArgs &a) {
  const u32 dl = absdiff(len, ecl.y);
  if (dl > th.y) return false;  // rejects almost every unrelated pair after one 8-byte shared load
  const u32 dc = absdiff(c, ecl.x);
  if (dc > th.w) return false;
  const u32 bk = ecl.x / DIVISOR;
  if (bk != b && bk != nbk) return false;
  if (dl < th.x && dc < th.z) return true;
  return band_candidate(dl, dc, len, a.len_ratio, a.pos_ratio);
}

// Tier 1.  After the tile is staged in shared memory the warps work independently, each on a window of 64 positions:
// lane l holds window positions l ("A") and 32+l ("B"; only the segment that straddles the middle of the window has B
// members).  All segment state lives in ballots:
//   phase 0: segment heads = ballot of key changes;
//   phase 1: candidate mask of every member over its predecessors in the segment (bit q <-> window position q);
//   phase 2: status in rounds — UNK / INS (an entry) / NOT.  A member with an inserted candidate is NOT at once; a member
//            whose candidates are all decided and none inserted is INS.  The first undecided member of a segment can
//            always decide, and a family of mutual candidates settles in two rounds (first member INS, all others NOT);
//   phase 3: owners from the final INS mask: one inserted candidate -> that entry; several -> exact scores.
__global__ void __launch_bounds__(MT_HEADS, RK_MT_MINB) k_match_small(MatchArgs a) {
  __shared__ u32 s_key[MT_TILE + 1];  // s_key[1 + e]; s_key[0] = key before the tile
  __shared__ u32 s_rank[MT_TILE];
  __shared__ uint2 s_cl[MT_TILE];
  __shared__ uint4 s_th[MT_TILE];     // {pass_len, rej_len, pass_pos, rej_pos}
  __shared__ uint2 s_bn[MT_TILE];     // {own bucket, neighbour bucket | NO_BUCKET}; x == NO_BUCKET: not a query
  __shared__ u32 s_pref[MT_HEADS / 32][64];     // per warp: exclusive prefix of the pair counts of the 64 window slots
  __shared__ u32 s_cand[MT_HEADS / 32][64][2];  // per warp: candidate mask of every slot (bit q <-> window position q)

  const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const u32 bs = blockIdx.x * MT_HEADS;
  const u32 count = min((u32)MT_TILE, a.m - bs);  // valid tile positions
  for (u32 e = tid; e < (u32)MT_TILE; e += MT_HEADS) {
    u32 key = 0xFFFFFFFFu;  // no real key (bucket indices are < 2^31): padding never continues a segment
    if (e < count) {
      u32 r, c, len;
      bool xm;
      key = a.skey[bs + e] >> a.key_shift;
      load_elem(a, bs + e, r, c, len, xm);
      s_rank[e] = r;
      s_cl[e] = make_uint2(c, len);
      const Thresh tl = make_thresh(len, a.len_ratio), tp = make_thresh(len, a.pos_ratio);
      s_th[e] = make_uint4(tl.pass, tl.rej, tp.pass, tp.rej);
      // X-matched fragments (Y pass) and length 0 (every score is NaN or 0) never query; they are entries
      s_bn[e] = make_uint2((xm || len == 0) ? NO_BUCKET : c / DIVISOR, neighbour_bucket(c, a.max_index));
    }
    s_key[1 + e] = key;
  }
  if (tid == 0) s_key[0] = bs ? a.skey[bs - 1] >> a.key_shift : 0xFFFFFFFEu;  // a segment running in from the previous tile is that CTA's
  __syncthreads();

  // phase 0
  const u32 wb = w << 5;  // tile position of window position 0
  const u32 eA = wb + lane, eB = eA + 32;
  const u32 hbA = __ballot_sync(0xFFFFFFFFu, s_key[1 + eA] != s_key[eA]);
  const u32 hbB = __ballot_sync(0xFFFFFFFFu, s_key[1 + eB] != s_key[eB]);
  const u32 le = 0xFFFFFFFFu >> (31 - lane);  // lanes <= this one
  const u32 lastA = hbA ? 31 - __clz(hbA) : 0;             // head of the straddling segment (when hbA != 0)
  const u32 firstB = hbB ? __ffs(hbB) - 1 : 32;            // its end, as a B index
  const bool straddle_long = firstB > lastA;               // 32 + firstB - lastA > 32 members
  const u32 ownA = hbA & le;
  const u32 hpA = ownA ? 31 - __clz(ownA) : 0;             // head (window position) of A's segment
  const bool actA = eA < count && ownA != 0 && !(hpA == lastA && straddle_long);
  const bool actB = eB < count && hbA != 0 && lane < firstB && !straddle_long;
  if (eA < count && hbA != 0 && lane == lastA && straddle_long) {  // more than 32 fragments: tier 2
    const u32 slot = atomicAdd(a.work_count, 1u);
    if (slot < a.work_cap) a.worklist[slot] = bs + eA;
    else atomicOr(a.err, ERR_WORKLIST);
  }

  // phase 1: the (query, predecessor) pairs of the window are spread evenly over the lanes — a family of 13 mutual
  // candidates has 78 pairs, and a lane-per-query loop would run as long as the longest segment of the window.
  // Pair t belongs to the last slot whose exclusive prefix of pair counts is <= t (binary search in shared memory).
  const uint2 *win = s_cl + wb;
  uint2 bnA = make_uint2(NO_BUCKET, 0), bnB = make_uint2(NO_BUCKET, 0);
  if (actA) bnA = s_bn[eA];
  if (actB) bnB = s_bn[eB];
  const u32 nA = bnA.x != NO_BUCKET ? lane - hpA : 0;  // predecessors to test
  const u32 nB = bnB.x != NO_BUCKET ? 32 + lane - lastA : 0;
  u32 incA = nA, incB = nB;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 va = __shfl_up_sync(0xFFFFFFFFu, incA, o), vb = __shfl_up_sync(0xFFFFFFFFu, incB, o);
    if ((int)lane >= o) incA += va, incB += vb;
  }
  const u32 totA = __shfl_sync(0xFFFFFFFFu, incA, 31);
  const u32 tot = totA + __shfl_sync(0xFFFFFFFFu, incB, 31);
  u32 candA = 0;
  unsigned long long candB = 0;
  if (tot) {
    u32 *pref = s_pref[w];
    u32(*cw)[2] = s_cand[w];
    pref[lane] = incA - nA;
    pref[32 + lane] = totA + incB - nB;
    cw[lane][0] = 0, cw[lane][1] = 0, cw[32 + lane][0] = 0, cw[32 + lane][1] = 0;
    __syncwarp();
#pragma unroll 1
    for (u32 t = lane; t < tot; t += 32) {
      u32 i = 0, base = 0;
#pragma unroll
      for (int step = 32; step >= 1; step >>= 1) {
        const u32 v = pref[i + step];
        if (v <= t) i += step, base = v;
      }
      const u32 q = i - (t - base + 1);  // window position of the predecessor
      const uint2 qcl = win[i], bn = s_bn[wb + i];
      if (is_candidate(win[q], qcl.x, qcl.y, bn.x, bn.y, s_th[wb + i], a)) atomicOr(&cw[i][q >> 5], 1u << (q & 31));
    }
    __syncwarp();
    candA = cw[lane][0];
    candB = ((unsigned long long)cw[32 + lane][1] << 32) | cw[32 + lane][0];
  }
  const uint2 clA = win[lane], clB = win[32 + lane];

  // phase 2
  enum { UNK = 0, INS = 1, NOT = 2 };
  int stA = candA ? UNK : (actA ? INS : NOT);
  int stB = candB ? UNK : (actB ? INS : NOT);
  u32 insA, insB;
  for (;;) {
    insA = __ballot_sync(0xFFFFFFFFu, stA == INS);
    insB = __ballot_sync(0xFFFFFFFFu, stB == INS);
    const u32 unkA = __ballot_sync(0xFFFFFFFFu, stA == UNK);
    const u32 unkB = __ballot_sync(0xFFFFFFFFu, stB == UNK);
    if ((unkA | unkB) == 0) break;
    if (stA == UNK) {
      if (candA & insA) stA = NOT;
      else if ((candA & unkA) == 0) stA = INS;
    }
    if (stB == UNK) {
      const unsigned long long ins = ((unsigned long long)insB << 32) | insA, unk = ((unsigned long long)unkB << 32) | unkA;
      if (candB & ins) stB = NOT;
      else if ((candB & unk) == 0) stB = INS;
    }
  }

  // phase 3 (a fragment that does not query keeps what it has: X-matched in the Y pass; length 0 gets "none")
  const bool y_pass = !stores_none(a);
  if (actA && !(y_pass && bnA.x == NO_BUCKET)) {
    const u32 hit = candA & insA;
    u32 owner = RK_NONE32;
    if (hit) owner = s_rank[wb + ((hit & (hit - 1)) == 0 ? (u32)__ffs(hit) - 1 : best_of(hit, win, clA.x, clA.y, a.len_ratio, a.pos_ratio))];
    if (hit || !y_pass) store_owner(a, s_rank[eA], owner);
  }
  if (actB && !(y_pass && bnB.x == NO_BUCKET)) {
    const unsigned long long hit = candB & (((unsigned long long)insB << 32) | insA);
    u32 owner = RK_NONE32;
    if (hit) owner = s_rank[wb + ((hit & (hit - 1)) == 0 ? (u32)__ffsll((long long)hit) - 1 : best_of(hit, win, clB.x, clB.y, a.len_ratio, a.pos_ratio))];
    if (hit || !y_pass) store_owner(a, s_rank[eB], owner);
  }
}

// ---- tier 2: one warp per segment of more than 32 fragments ------------------------------------------------------
// 32 queries at a time are checked against the entry list (uniform loads; every lane its own query), then the
// insertions inside the chunk are replayed in order with ballots.  Two exact reductions keep the work per query at
// O(distinct entries) instead of O(segment):
//  * the candidate test (deviation > 0) uses the integer thresholds of tier 1; the two divisions are only done for
//    candidates, to rank them;
//  * an entry with the same (center, length) as a newer entry can never win: its score is identical and the newer
//    one is scanned first (`>` is strict, SequenceOcupationList.cpp:40).  Inserting therefore retires the older
//    duplicate (tombstone: rank = NONE; the list is compacted, order preserved, when half of it is dead).  In a
//    repeat family thousands of X-matched fragments Y-insert the same few (center, length) pairs.
__global__ void __launch_bounds__(128) k_match_long(MatchArgs a) {
  const u32 lane = threadIdx.x & 31;
  const u32 nseg = min(*a.work_count, a.work_cap);
  for (;;) {
    u32 seg = 0;
    if (lane == 0) seg = atomicAdd(a.work_count + 1, 1u);
    seg = __shfl_sync(0xFFFFFFFFu, seg, 0);
    if (seg >= nseg) return;
    const u32 start = a.worklist[seg];
    const u32 key = a.skey[start] >> a.key_shift;
    u32 end = start;
    for (;;) {
      const u32 idx = end + lane;
      const bool same = idx < a.m && (a.skey[idx] >> a.key_shift) == key;
      const u32 bal = __ballot_sync(0xFFFFFFFFu, same);
      end += __popc(bal);  // sorted keys: the matching lanes are a prefix
      if (bal != 0xFFFFFFFFu) break;
    }
    u32 *const e_rank = a.ent_rank + start, *const e_c = a.ent_c + start, *const e_len = a.ent_len + start;
    u32 n_ent = 0, n_dead = 0;
    for (u32 base = start; base < end; base += 32) {
      const u32 q = base + lane;
      const bool valid = q < end;
      u32 r = 0, c = 0, len = 0;
      bool xm = false;
      if (valid) load_elem(a, q, r, c, len, xm);
      const bool needq = valid && !xm && len != 0;  // length 0 never matches (every score is NaN or 0)
      const u32 b = c / DIVISOR;
      const u32 nbk = neighbour_bucket(c, a.max_index);
      const Thresh tl = make_thresh(len, a.len_ratio), tp = make_thresh(len, a.pos_ratio);
      double best_sc = 0.0;
      u32 best = RK_NONE32;
      bool best_own = false;
      u32 best_dl = 0, best_dc = 0;  // the two differences of the best entry so far
      // Exact scores cost two fp64 divisions; most of them are avoidable.  For one query the score is a function of
      // (dl, dc) that falls STRICTLY when either difference grows (every rounding step is monotone, and with
      // t = length*ratio < 2^30 one unit of difference moves a term by far more than an ulp).  So an entry whose
      // differences are both >= the best's is either the same pair of differences — the same score, only the tie rule
      // speaks — or strictly worse, and nothing has to be computed.
      const bool prune_ok = tl.t < 0x1p30 && tp.t < 0x1p30;
      // is entry (ec, el) a candidate (deviation > 0) of this lane's query?
      auto candidate = [&](u32 ec, u32 el, bool &own, u32 &dl, u32 &dc) -> bool {
        const u32 bk = ec / DIVISOR;
        own = bk == b;
        if (!own && bk != nbk) return false;
        dl = len > el ? len - el : el - len;
        dc = c > ec ? c - ec : ec - c;
        if (dl > tl.rej || dc > tp.rej) return false;
        const int ql = quotient_vs_one(dl, tl);
        if (ql == 0) return false;
        const int qp = quotient_vs_one(dc, tp);
        return !(qp == 0 || (ql == 2 && qp == 2));
      };
      // phase A: entries inserted before this chunk, newest first (uniform loads).  Every lane scores its query and,
      // because any lane may become an entry in phase B, notes the live entry with its own (center, length).
      // The list is read 32 entries at a time, one entry per lane (coalesced), and handed round with shuffles: the
      // inner loop then waits for no memory at all (a uniform load per entry cost two dependent cache round trips).
      u32 dup_idx = RK_NONE32;
      for (u32 hi = n_ent; hi > 0;) {
        const u32 cnt = hi < 32 ? hi : 32, base_t = hi - cnt;
        u32 er_l = RK_NONE32, ec_l = 0, el_l = 0;
        if (lane < cnt) er_l = e_rank[base_t + lane], ec_l = e_c[base_t + lane], el_l = e_len[base_t + lane];
        u32 live = __ballot_sync(0xFFFFFFFFu, er_l != RK_NONE32);  // retired duplicates are skipped
        while (live) {
          const int j = 31 - __clz(live);  // newest first
          live &= ~(1u << j);
          const u32 t = base_t + (u32)j;
          const u32 er = __shfl_sync(0xFFFFFFFFu, er_l, j), ec = __shfl_sync(0xFFFFFFFFu, ec_l, j), el = __shfl_sync(0xFFFFFFFFu, el_l, j);
          if (ec == c && el == len) dup_idx = t;  // at most one entry per (center, length) is live
          if (needq) {
            bool own;
            u32 dl, dc;
            if (candidate(ec, el, own, dl, dc)) {
              if (prune_ok && best != RK_NONE32 && dl >= best_dl && dc >= best_dc) {
                // same differences: same score, an older entry only wins with the own-bucket rule; else strictly worse
                if (dl == best_dl && dc == best_dc && own && !best_own) best = er, best_own = true;
              } else {
                const double sc = deviation(ec, el, c, len, tl.t, tp.t);
                if (sc > best_sc || (sc == best_sc && best != RK_NONE32 && own && !best_own)) {
                  best_sc = sc, best = er, best_own = own, best_dl = dl, best_dc = dc;
                }
              }
            }
          }
        }
        hi = base_t;
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
        // retire the live entry with the same (center, length), if any: the one phase A found, or the one an earlier
        // lane of this chunk has appended since (every lane tracks the live entry of its own pair)
        const u32 kill = __shfl_sync(0xFFFFFFFFu, dup_idx, L);
        if (kill != RK_NONE32) ++n_dead;
        if ((int)lane == L) {
          if (kill != RK_NONE32) e_rank[kill] = RK_NONE32;
          e_rank[n_ent] = r;
          e_c[n_ent] = c;
          e_len[n_ent] = len;
        }
        if (c == Lc && len == Ll) dup_idx = n_ent;
        ++n_ent;
        if ((int)lane > L && needq) {
          // the new entry is the newest: it is scanned before every older entry of its bucket class
          bool own;
          u32 dl, dc;
          if (candidate(Lc, Ll, own, dl, dc)) {
            if (prune_ok && best != RK_NONE32 && dl >= best_dl && dc >= best_dc) {
              if (dl == best_dl && dc == best_dc && (own || !best_own)) best = Lr, best_own = own;  // same score, newer
            } else {
              const double sc = deviation(Lc, Ll, c, len, tl.t, tp.t);
              if (sc > best_sc || (sc == best_sc && sc > 0.0 && (own || !best_own))) {
                best_sc = sc, best = Lr, best_own = own, best_dl = dl, best_dc = dc;
              }
            }
          }
        }
        pending &= ~((2u << L) - 1u);  // lanes <= L are final
      }
      if (valid && !xm) {
        if (best != RK_NONE32) store_owner(a, r, best);
        else if (!a.is_y) store_owner(a, r, RK_NONE32);  // the X pass stores "none" for every fragment it owns
      }
      __syncwarp();  // entry stores of this chunk are read by every lane in the next one
      // stable compaction once half of the list is dead
      if (n_dead > 32 && 2 * n_dead > n_ent) {
        u32 w = 0;
        for (u32 t0 = 0; t0 < n_ent; t0 += 32) {
          const u32 t = t0 + lane;
          u32 er = RK_NONE32, ec = 0, el = 0;
          if (t < n_ent) {
            er = e_rank[t];
            ec = e_c[t];
            el = e_len[t];
          }
          const u32 live = __ballot_sync(0xFFFFFFFFu, er != RK_NONE32);
          __syncwarp();           // every lane holds its entry before any lane stores into this batch's range
          if (er != RK_NONE32) {  // destination index <= t: never overwrites an entry that is still to be read
            const u32 dst = w + __popc(live & lanemask_lt());
            e_rank[dst] = er;
            e_c[dst] = ec;
            e_len[dst] = el;
          }
          w += __popc(live);
          __syncwarp();
        }
        n_ent = w;
        n_dead = 0;
      }
    }
  }
}

int launch_match(const MatchArgs &a, cudaStream_t st) {
  if (a.m == 0) return 0;
  cudaMemsetAsync(a.work_count, 0, 2 * sizeof(u32), st);
  if (!a.is_y) cudaMemsetAsync(a.xm_bits, 0, ((size_t)a.m + 31) / 32 * sizeof(u32), st);
  {
    KScope ks(KID_MATCH_SMALL, st, a.m);
    k_match_small<<<(a.m + MT_HEADS - 1) / MT_HEADS, MT_HEADS, 0, st>>>(a);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  KScope ks(KID_MATCH_LONG, st, 0);
  k_match_long<<<sms * 4, 128, 0, st>>>(a);
  return 2;
}

}  // namespace rk
