// One comparison partitioned over several GPUs: the orchestration behind rk_dist_* (one process per GPU) and
// rk_create_multi / rk_multi_* (one process, one host thread per GPU) of include/rk_b200.h.  SURVEY.md §8e.
//
// The reference has no distributed path; what is reproduced is generate_fragment_groups +
// generate_diagonal_func + sort_groups (/root/reference/src/commonFunctions.cpp:41-80,161-177,148-159) on ONE fragment
// file whose records start out spread over the GPUs in file order.  The kernels and the argument for exactness are in
// k7_dist.cu; this file owns the buffers, the two transports and the order of the launches:
//
//   load   K1 on the local slice -> [all-gather: xStart/10 histogram + link maps] -> cuts -> exchange 1: the split kernel
//          stores every 32-byte record into its owner's arena (peer memory) -> processing order (K2a) -> keys (K2) -> h (K5a)
//          -> X halo rows to the higher GPUs; Y rows (16 B) to the owners of their Y super-bucket range, by copy engines
//          beside the X bucket sort -> X and Y bucket sorts (K2b/c)                              [2 host syncs: counts]
//   group  X pass (K3) -> halo owners stored at their home rank (4 B) -> "matched in X" bytes stored at the Y owners
//          -> Y pass (K3) -> the Y owners stored at the home ranks -> root scan; local chains; chains that leave the GPU in
//          bulk ask/answer rounds (many) or by walking the peers' res[] (few) (K4) -> output rows (16 B) stored at the
//          owner of their group-id range -> K5b/c                                                [2 host syncs: counts]
//
// Every rank's partition arena is carved identically and mapped into every peer (cudaIpc handles between processes, raw
// pointers + cudaDeviceEnablePeerAccess inside one), so a local buffer address translates to the same buffer on any rank;
// kernels store rows straight into it (k7_dist.cu: k_push_rows, scatter_store), copy engines push the Y rows.
// NCCL (dlopen'ed: the process may already hold torch's libnccl.so.2) carries the small tables (ncclAllGather of counts,
// histograms, link maps) and is the barrier behind every push.  Ranks that are threads of one process, possibly on the same
// device (tests on a one-GPU box), exchange the small tables through host memory between host barriers instead.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "rk_ctx.cuh"

using namespace rk;

namespace rk {

// ---------------------------------------------------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------------------------------------------------
struct NcclApi {
  void *lib = nullptr;
  std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok() const { return lib != nullptr && err.empty(); }
};
static NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // an already loaded libnccl.so.2 (torch's) is reused: same soname
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) {
      api.err = std::string("dlopen(libnccl.so.2): ") + dlerror();
      return;
    }
#define RK_NCCL_SYM(field, name)                                         \
  api.field = (decltype(api.field))dlsym(api.lib, name);                 \
  if (!api.field) api.err = std::string("libnccl lacks ") + name;
    RK_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    RK_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    RK_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    RK_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    RK_NCCL_SYM(Send, "ncclSend")
    RK_NCCL_SYM(Recv, "ncclRecv")
    RK_NCCL_SYM(AllGather, "ncclAllGather")
    RK_NCCL_SYM(GroupStart, "ncclGroupStart")
    RK_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    RK_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef RK_NCCL_SYM
  });
  return api;
}

// ---------------------------------------------------------------------------------------------------------------------
// transports
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SMALL_WORDS = 96;  // per-rank row of the count exchanges
constexpr int W_CNT_A = 0;       // [nr+1] destination counts (exchange 1 / X halo / output exchange)
constexpr int W_CNT_B = 17;      // [nr+1] destination counts of the Y exchange
constexpr int W_ERR = 40;        // device error word of the rank
constexpr int W_X0 = 41;         // roots on the rank
constexpr int W_X1 = 42;         // total groups
constexpr int W_X2 = 43;         // forest: fragments of the rank whose chain is still waiting for an answer from another rank
constexpr int W_CUTSX = 64;      // [2][nr] the first X super-bucket key every rank owns, per strand class
constexpr int W_CUTS = 44;       // [nr+1] the cuts of the exchange (identical on every rank; the host sizes the local sort keys from them)

struct Transport {
  int rank = 0, world = 1;
  std::string err;
  u64 bytes_sent = 0;  // payload this rank handed to other ranks
  virtual ~Transport() {}
  // every rank contributes SMALL_WORDS u32 from device memory; all rows land in h_all[world][SMALL_WORDS] (pinned).
  // Synchronises the stream.
  virtual int gather_small(const u32 *d_row, u32 *d_all, u32 *h_all, cudaStream_t st) = 0;
  // The same in two halves: begin() leaves all rows in DEVICE memory d_all (stream-ordered: kernels queued behind it may
  // read the other ranks' counts), end() brings them to the host and synchronises.
  virtual int gather_small_begin(const u32 *d_row, u32 *d_all, u32 *h_all, cudaStream_t st) = 0;
  virtual int gather_small_end(const u32 *d_all, u32 *h_all, cudaStream_t st) = 0;
  // all ranks' streams meet: work queued behind it starts after every rank's work queued before it is complete
  virtual int barrier(cudaStream_t st) = 0;
  // recv[world][bytes] <- every rank's send[bytes]; stream-ordered; doubles as a barrier between the ranks' streams
  virtual int all_gather(const void *send, void *recv, size_t bytes, cudaStream_t st) = 0;
  // variable all-to-all; offsets and counts in elements of `elem` bytes.  poff[p] = where this rank's block starts in
  // rank p's receive buffer (used by the transports that write into the peer's memory).
  virtual int all_to_all(const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *roff, const u64 *rcnt,
                         const u64 *poff, size_t elem, cudaStream_t st) = 0;
  // The same in two halves, so that rows can travel while kernels run: begin() enqueues the transfers on `side`, ordered
  // after everything queued on st so far; end() makes st wait for them and for every other rank's.  (Default: all in begin.)
  virtual int a2a_begin(const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *roff, const u64 *rcnt,
                        const u64 *poff, size_t elem, cudaStream_t st, cudaStream_t) {
    return all_to_all(send, soff, scnt, recv, roff, rcnt, poff, elem, st);
  }
  virtual int a2a_end(cudaStream_t) { return RK_OK; }
  // peer memory: every rank's partition arena has the same layout, so a local pointer translates to any peer's copy
  const u8 *my_base = nullptr;
  u8 *peer_base[DIST_MAX_RANKS] = {nullptr};
  bool peers_mapped = false;
  template <class T>
  T *on_peer(int p, T *local) const { return (T *)(peer_base[p] + ((const u8 *)local - my_base)); }
};

struct NcclTransport : Transport {
  ncclComm_t comm = nullptr;
  bool own_comm = true;
  int check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return RK_OK;
    err = std::string(what) + ": " + nccl_api().GetErrorString(r);
    return RK_ERR_CUDA;
  }
  ~NcclTransport() override {
    if (comm && own_comm) nccl_api().CommDestroy(comm);
    if (ev_a) cudaEventDestroy(ev_a);
    if (ev_b) cudaEventDestroy(ev_b);
    for (auto &a : aux) if (a) cudaStreamDestroy(a);
    for (auto &e : aux_ev) if (e) cudaEventDestroy(e);
  }
  int gather_small(const u32 *d_row, u32 *d_all, u32 *h_all, cudaStream_t st) override {
    int rc = check(nccl_api().AllGather(d_row, d_all, SMALL_WORDS * sizeof(u32), ncclUint8, comm, st), "ncclAllGather");
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(h_all, d_all, (size_t)world * SMALL_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      err = std::string("count exchange: ") + cudaGetErrorString(e);
      return RK_ERR_CUDA;
    }
    return RK_OK;
  }
  int gather_small_begin(const u32 *d_row, u32 *d_all, u32 *, cudaStream_t st) override {
    return check(nccl_api().AllGather(d_row, d_all, SMALL_WORDS * sizeof(u32), ncclUint8, comm, st), "ncclAllGather");
  }
  int gather_small_end(const u32 *d_all, u32 *h_all, cudaStream_t st) override {
    cudaError_t e = cudaMemcpyAsync(h_all, d_all, (size_t)world * SMALL_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      err = std::string("count exchange: ") + cudaGetErrorString(e);
      return RK_ERR_CUDA;
    }
    return RK_OK;
  }
  int barrier(cudaStream_t st) override { return check(nccl_api().AllGather(d_flag, d_flag_all, 4, ncclUint8, comm, st), "ncclAllGather (barrier)"); }
  int all_gather(const void *send, void *recv, size_t bytes, cudaStream_t st) override {
    bytes_sent += bytes * (u64)(world - 1);
    return check(nccl_api().AllGather(send, recv, bytes, ncclUint8, comm, st), "ncclAllGather");
  }
  // Bulk rows do not go through NCCL's send/recv kernels: every rank PUSHES its blocks into the peers' receive buffers
  // with copy-engine peer copies over NVLink (the arenas are mapped into each other: cudaIpc between processes), peers
  // in staggered order so that no receiver is hit by everybody at once; a 4-byte ncclAllGather behind the copies is the
  // barrier that tells a rank that all blocks addressed to it have landed.
  int all_to_all(const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *roff, const u64 *rcnt,
                 const u64 *poff, size_t elem, cudaStream_t st) override {
    const NcclApi &N = nccl_api();
    if (peers_mapped && !getenv_nccl_rows()) {
      const int rc = push_blocks(send, soff, scnt, recv, poff, elem, st, st);
      if (rc) return rc;
      return check(N.AllGather(d_flag, d_flag_all, 4, ncclUint8, comm, st), "ncclAllGather (barrier)");
    }
    if (scnt[rank]) {
      if (cudaMemcpyAsync((u8 *)recv + roff[rank] * elem, (const u8 *)send + soff[rank] * elem, scnt[rank] * elem,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        err = "all_to_all: local copy failed";
        return RK_ERR_CUDA;
      }
    }
    int rc = check(N.GroupStart(), "ncclGroupStart");
    for (int p = 0; p < world && !rc; ++p) {
      if (p == rank) continue;
      if (scnt[p]) {
        rc = check(N.Send((const u8 *)send + soff[p] * elem, scnt[p] * elem, ncclUint8, p, comm, st), "ncclSend");
        bytes_sent += scnt[p] * elem;
      }
      if (rcnt[p] && !rc) rc = check(N.Recv((u8 *)recv + roff[p] * elem, rcnt[p] * elem, ncclUint8, p, comm, st), "ncclRecv");
    }
    const int rc2 = check(N.GroupEnd(), "ncclGroupEnd");
    return rc ? rc : rc2;
  }
  int a2a_begin(const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *roff, const u64 *rcnt, const u64 *poff,
                size_t elem, cudaStream_t st, cudaStream_t side) override {
    if (!(peers_mapped && !getenv_nccl_rows())) return all_to_all(send, soff, scnt, recv, roff, rcnt, poff, elem, st);
    if (!ev_a && (cudaEventCreateWithFlags(&ev_a, cudaEventDisableTiming) != cudaSuccess ||
                  cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming) != cudaSuccess)) {
      err = "a2a_begin: cudaEventCreate failed";
      return RK_ERR_CUDA;
    }
    cudaEventRecord(ev_a, st);
    cudaStreamWaitEvent(side, ev_a, 0);
    const int rc = push_blocks(send, soff, scnt, recv, poff, elem, side, side);
    if (rc) return rc;
    cudaEventRecord(ev_b, side);
    split_pending = true;
    return RK_OK;
  }
  int a2a_end(cudaStream_t st) override {
    if (!split_pending) return RK_OK;
    split_pending = false;
    cudaStreamWaitEvent(st, ev_b, 0);
    return check(nccl_api().AllGather(d_flag, d_flag_all, 4, ncclUint8, comm, st), "ncclAllGather (barrier)");
  }
  // This rank's blocks into the peers' receive buffers, peers in staggered order.  Large exchanges are spread over a few
  // copy streams (one copy engine each), forked from `from` and joined into `into`: a single peer copy does not fill the
  // links of a GPU (measured 550 GB/s for one copy at a time).
  static constexpr int NAUX = 3;
  cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr};
  cudaEvent_t aux_ev[NAUX + 1] = {nullptr, nullptr, nullptr, nullptr};
  int push_blocks(const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *poff, size_t elem, cudaStream_t from,
                  cudaStream_t into) {
    u64 total = 0;
    int peers = 0;
    for (int p = 0; p < world; ++p)
      if (p != rank && scnt[p]) total += scnt[p] * elem, ++peers;
    const bool spread = peers > 1 && total >= (8u << 20);
    if (spread && !aux[0]) {
      for (int a = 0; a < NAUX; ++a) cudaStreamCreateWithFlags(&aux[a], cudaStreamNonBlocking);
      for (int a = 0; a <= NAUX; ++a) cudaEventCreateWithFlags(&aux_ev[a], cudaEventDisableTiming);
    }
    if (spread) {
      cudaEventRecord(aux_ev[NAUX], from);
      for (int a = 0; a < NAUX; ++a) cudaStreamWaitEvent(aux[a], aux_ev[NAUX], 0);
    }
    int turn = 0;
    for (int k = 0; k < world; ++k) {
      const int p = (rank + k) % world;
      if (!scnt[p]) continue;
      u8 *dst = (p == rank ? (u8 *)recv : on_peer(p, (u8 *)recv)) + poff[p] * elem;
      cudaStream_t cs = from;
      if (spread && p != rank) cs = (turn % (NAUX + 1)) == NAUX ? from : aux[turn % (NAUX + 1)], ++turn;
      if (cudaMemcpyAsync(dst, (const u8 *)send + soff[p] * elem, scnt[p] * elem, cudaMemcpyDeviceToDevice, cs) != cudaSuccess) {
        err = std::string("all_to_all: peer copy failed: ") + cudaGetErrorString(cudaGetLastError());
        return RK_ERR_CUDA;
      }
      if (p != rank) bytes_sent += scnt[p] * elem;
    }
    if (spread)
      for (int a = 0; a < NAUX; ++a) {
        cudaEventRecord(aux_ev[a], aux[a]);
        cudaStreamWaitEvent(into, aux_ev[a], 0);
      }
    return RK_OK;
  }
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  bool split_pending = false;
  u32 *d_flag = nullptr, *d_flag_all = nullptr;  // barrier payload (set by the owner of the transport)
  static bool getenv_nccl_rows() {  // RK_DIST_NCCL_ROWS=1: rows through ncclSend/ncclRecv (tuning reference)
    static const bool v = getenv("RK_DIST_NCCL_ROWS") != nullptr;
    return v;
  }
};

// ranks = threads of one process (any devices, also the same one): device copies between host barriers
struct LocalGroup {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  u64 generation = 0;
  const void *ptr[DIST_MAX_RANKS] = {nullptr};
  u64 soff[DIST_MAX_RANKS][DIST_MAX_RANKS] = {{0}};
  u64 scnt[DIST_MAX_RANKS][DIST_MAX_RANKS] = {{0}};
  u32 small[DIST_MAX_RANKS][SMALL_WORDS] = {{0}};
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const u64 gen = generation;
    if (++waiting == n) {
      waiting = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

struct LocalTransport : Transport {
  LocalGroup *g = nullptr;
  int cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return RK_OK;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return RK_ERR_CUDA;
  }
  int gather_small(const u32 *d_row, u32 *, u32 *h_all, cudaStream_t st) override {
    cudaError_t e = cudaMemcpyAsync(h_all + (size_t)rank * SMALL_WORDS, d_row, SMALL_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) memcpy(g->small[rank], h_all + (size_t)rank * SMALL_WORDS, SMALL_WORDS * sizeof(u32));
    g->barrier();
    for (int r = 0; r < world; ++r) memcpy(h_all + (size_t)r * SMALL_WORDS, g->small[r], SMALL_WORDS * sizeof(u32));
    g->barrier();
    return cuda(e, "count exchange");
  }
  int gather_small_begin(const u32 *d_row, u32 *d_all, u32 *h_all, cudaStream_t st) override {
    const int rc = gather_small(d_row, d_all, h_all, st);  // the host exchange, then the matrix goes back to the device
    if (rc) return rc;
    return cuda(cudaMemcpyAsync(d_all, h_all, (size_t)world * SMALL_WORDS * sizeof(u32), cudaMemcpyHostToDevice, st), "count upload");
  }
  int gather_small_end(const u32 *, u32 *, cudaStream_t st) override { return cuda(cudaStreamSynchronize(st), "count exchange"); }
  int barrier(cudaStream_t st) override {
    const cudaError_t e = cudaStreamSynchronize(st);
    g->barrier();
    return cuda(e, "barrier");
  }
  int all_gather(const void *send, void *recv, size_t bytes, cudaStream_t st) override {
    cudaError_t e = cudaStreamSynchronize(st);
    g->ptr[rank] = send;
    g->barrier();
    for (int r = 0; r < world && e == cudaSuccess; ++r)
      e = cudaMemcpyAsync((u8 *)recv + (size_t)r * bytes, g->ptr[r], bytes, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    g->barrier();
    bytes_sent += bytes * (u64)(world - 1);
    return cuda(e, "all_gather");
  }
  int all_to_all(const void *send, const u64 *soff, const u64 *scnt, void *recv, const u64 *roff, const u64 *rcnt, const u64 *,
                 size_t elem, cudaStream_t st) override {
    cudaError_t e = cudaStreamSynchronize(st);
    g->ptr[rank] = send;
    for (int p = 0; p < world; ++p) g->soff[rank][p] = soff[p], g->scnt[rank][p] = scnt[p];
    g->barrier();
    bool mismatch = false;
    for (int r = 0; r < world && e == cudaSuccess; ++r) {
      const u64 cnt = g->scnt[r][rank];
      if (cnt != rcnt[r]) mismatch = true;
      if (cnt && !mismatch)
        e = cudaMemcpyAsync((u8 *)recv + roff[r] * elem, (const u8 *)g->ptr[r] + g->soff[r][rank] * elem, cnt * elem, cudaMemcpyDefault, st);
      if (r != rank) bytes_sent += scnt[r] * elem;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    g->barrier();
    if (mismatch) {
      err = "all_to_all: send and receive counts disagree";
      return RK_ERR_INTERNAL;
    }
    return cuda(e, "all_to_all");
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// per-rank state
// ---------------------------------------------------------------------------------------------------------------------
// RK_DIST_TRACE=1: CUDA-event pairs around every transport call and the sections between them, printed per rank after
// the final synchronisation of rk_dist_group (a tuning aid: where a step spends its time)
struct Trace {
  bool on = false;
  struct Rec { const char *what; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void begin(const char *what, cudaStream_t st) {
    if (!on) return;
    Rec r{what, get(), get()};
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) {
    if (on && !recs.empty()) cudaEventRecord(recs.back().b, st);
  }
  void dump(int rank, const std::string &counts) {
    if (!on) return;
    std::string line = "[rk_dist rank " + std::to_string(rank) + "] " + counts;
    float total = 0.f;
    for (auto &r : recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) cudaGetLastError();
      // label: the transport function and its first argument, e.g. all_to_all(D.send_rows
      std::string w(r.what);
      const size_t p0 = w.find("->"), p1 = w.find(',');
      w = w.substr(p0 == std::string::npos ? 0 : p0 + 2, p1 == std::string::npos ? std::string::npos : p1 - (p0 == std::string::npos ? 0 : p0 + 2));
      char buf[160];
      snprintf(buf, sizeof buf, " %s)=%.3f", w.c_str(), ms);
      line += buf;
      total += ms;
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.clear();
    fprintf(stderr, "%s | comm total %.3f ms\n", line.c_str(), total);
  }
};

struct Exchange {  // counts of one variable all-to-all, as this rank sees it (elements)
  u64 soff[DIST_MAX_RANKS], scnt[DIST_MAX_RANKS], roff[DIST_MAX_RANKS], rcnt[DIST_MAX_RANKS];
  u64 poff[DIST_MAX_RANKS];   // forward: start of my block in rank p's receive buffer (= p's roff[me])
  u64 rpoff[DIST_MAX_RANKS];  // the way back: start of my block in rank p's send-order buffer (= p's soff[me])
  u64 n_send = 0, n_recv = 0;
  // from the gathered count matrix: row s = what rank s sends to every destination
  void from_matrix(const u32 *h_all, int word0, int nr, int me) {
    n_send = n_recv = 0;
    for (int d = 0; d < nr; ++d) {
      soff[d] = n_send;
      scnt[d] = h_all[(size_t)me * SMALL_WORDS + word0 + d];
      n_send += scnt[d];
    }
    for (int s = 0; s < nr; ++s) {
      roff[s] = n_recv;
      rcnt[s] = h_all[(size_t)s * SMALL_WORDS + word0 + me];
      n_recv += rcnt[s];
    }
    for (int p = 0; p < nr; ++p) {
      u64 f = 0, r = 0;
      for (int s = 0; s < me; ++s) f += h_all[(size_t)s * SMALL_WORDS + word0 + p];
      for (int d = 0; d < me; ++d) r += h_all[(size_t)p * SMALL_WORDS + word0 + d];
      poff[p] = f, rpoff[p] = r;
    }
  }
};

struct DistBlob {  // what a rank publishes so that its peers can read its parent / root-scan arrays
  u64 pid, ptr, cap;
  int device, pad;
  cudaIpcMemHandle_t handle;
};
static_assert(sizeof(DistBlob) <= RK_DIST_BLOB_BYTES, "blob size");

struct Dist {
  int rank = 0, world = 1;
  Transport *tr = nullptr;
  u64 cap = 0, hcap = 0;
  void *arena = nullptr;
  void *peer_open[DIST_MAX_RANKS] = {nullptr};
  bool peers_ready = false;
  u32 *h_small = nullptr;  // pinned [world][SMALL_WORDS]
  PeerTable pt{};
  Trace trace;
  cudaStream_t side = nullptr;  // rows that travel while kernels run (the Y exchange beside the X sort)
  // grow-only side buffers whose size depends on the input, not on cap
  u8 *aos_buf = nullptr;
  u64 aos_bytes = 0;
  u32 *link_buf = nullptr;  // [hist0][linkx_loc][linky_loc] (one all-gather) [linkx][linky][gather: world * the first three]
  u64 link_words = 0;
  void *h_res = nullptr;
  u64 h_res_cap = 0;

  // state of the last load
  bool loaded = false;
  Geometry g{};
  int bits_rank = 1, bits_y = 1;
  u32 m_loc = 0, rank_off = 0, n_halo = 0, n_away = 0, m_y = 0;
  u64 m_total = 0, n_total_loaded = 0;
  Exchange ex1, exh, exy;

  // carved device pointers
  Counters *cnt = nullptr;
  u32 *d_small = nullptr, *d_small_all = nullptr, *nroots_mat = nullptr;
  u32 *prehist = nullptr;  // 4 x [4][256]: digit counts of the rank / X / Y / gid sort keys, gathered by their producers
  u32 *hist = nullptr, *hist_all = nullptr, *cuts0 = nullptr, *cuts_y = nullptr, *cuts_x = nullptr, *cuts_g = nullptr;
  uint4 *rec4_loc = nullptr, *send_rows = nullptr, *recv_rows = nullptr;
  uint2 *rec6_arr = nullptr;  // the records this rank owns, in arrival order: 24-byte rows stored by the peers' push kernels
  u32 *key0_loc = nullptr, *tile_cnt = nullptr, *perm_x = nullptr, *perm_y = nullptr;
  u32 *key0a = nullptr, *k0_r = nullptr, *aidx_r = nullptr, *tmp_k = nullptr, *tmp_v = nullptr;
  void *sort_work = nullptr;
  uint2 *xl = nullptr, *yl_r = nullptr, *yl_a = nullptr;
  u32 *ys_r = nullptr, *kx2 = nullptr, *ky = nullptr, *gfidx_r = nullptr;
  float *identity_r = nullptr;
  uint4 *hfi_r = nullptr;
  uint4 *halo_send = nullptr, *halo_recv = nullptr;
  u32 *halo_grank = nullptr, *away_res = nullptr;
  u32 *skx = nullptr, *rx = nullptr, *ky_a = nullptr, *grank_a = nullptr, *sky_a = nullptr, *ry_a = nullptr;
  u32 *parent_x = nullptr, *xm_bits = nullptr, *parent = nullptr, *gidscan = nullptr, *parent_y = nullptr, *yo_s = nullptr;
  u8 *xm_a = nullptr;
  u32 *ent_rank = nullptr, *ent_c = nullptr, *ent_len = nullptr, *worklist = nullptr;
  u32 work_cap = 0;
  u32 *gid_rank = nullptr, *flag = nullptr, *flag_all = nullptr, *lroot = nullptr, *gid_l = nullptr;
  u32 *pend_idx[2] = {nullptr, nullptr}, *pend_key[2] = {nullptr, nullptr}, *n_pend = nullptr;  // forest: pending lists (double buffered)
  u32 *apos = nullptr, *roff_dev = nullptr;
  uint4 *queries = nullptr;  // forest bulk rounds: the ranks other GPUs ask this one about (stored by their push kernels)
  u64 *answers = nullptr;    // ... and the answers to this rank's questions, in the order it sent them (stored by the owners)
  u64 *res = nullptr;  // per fragment: what a walker from another GPU needs (k_chase_local); the peers read it
  void *scan_work = nullptr;
  u32 *gid_a = nullptr, *sgid = nullptr, *srank_g = nullptr;
  void *order_scratch = nullptr;
  u32 *out_order = nullptr, *out_gid = nullptr;
  u8 *out_repval = nullptr;
  float *out_identity = nullptr;
};

static u64 dist_carve(Dist &D, u8 *base) {
  u64 off = 0;
  auto take = [&](u64 bytes) -> u8 * {
    u8 *p = base ? base + off : nullptr;
    off += align_up(bytes ? bytes : 16, 256);
    return p;
  };
  const u64 M = D.cap, H = D.hcap, MC = M + H;
  const int nr = D.world;
  D.cnt = (Counters *)take(sizeof(Counters));
  D.d_small = (u32 *)take(SMALL_WORDS * 4);
  D.d_small_all = (u32 *)take((u64)nr * SMALL_WORDS * 4);
  D.nroots_mat = (u32 *)take((u64)nr * SMALL_WORDS * 4);
  D.prehist = (u32 *)take(4 * 4 * 256 * 4);
  D.hist = (u32 *)take(DIST_BINS * 4);
  D.hist_all = (u32 *)take((u64)nr * DIST_BINS * 4);
  D.cuts0 = (u32 *)take((DIST_MAX_RANKS + 1) * 4);
  D.cuts_y = (u32 *)take((DIST_MAX_RANKS + 1) * 4);
  D.cuts_x = (u32 *)take(2 * DIST_MAX_RANKS * 4);
  D.cuts_g = (u32 *)take((DIST_MAX_RANKS + 1) * 4);
  D.rec4_loc = (uint4 *)take(M * 32);
  D.send_rows = (uint4 *)take(M * 16);
  D.rec6_arr = (uint2 *)take(M * 24);
  D.recv_rows = (uint4 *)take(M * 16);
  D.key0_loc = (u32 *)take(M * 4);
  D.tile_cnt = (u32 *)take(dist_split_work_bytes(M));
  D.perm_x = (u32 *)take(H * 4);
  D.perm_y = (u32 *)take(M * 4);
  D.key0a = (u32 *)take(M * 4);
  D.k0_r = (u32 *)take(M * 4);
  D.aidx_r = (u32 *)take(M * 4);
  D.tmp_k = (u32 *)take(MC * 4);
  D.tmp_v = (u32 *)take(MC * 4);
  D.sort_work = take(sort_work_bytes(MC));
  D.xl = (uint2 *)take(MC * 8);
  D.yl_r = (uint2 *)take(M * 8);
  D.yl_a = (uint2 *)take(M * 8);
  D.ys_r = (u32 *)take(M * 4);
  D.kx2 = (u32 *)take(MC * 4);
  D.ky = (u32 *)take(M * 4);
  D.gfidx_r = (u32 *)take(M * 4);
  D.identity_r = (float *)take(M * 4);
  D.hfi_r = (uint4 *)take(M * 16);
  D.halo_send = (uint4 *)take(H * 16);
  D.halo_recv = (uint4 *)take(H * 16);
  D.halo_grank = (u32 *)take(H * 4);
  D.away_res = (u32 *)take(H * 4);
  D.skx = (u32 *)take(MC * 4);
  D.rx = (u32 *)take(MC * 4);
  D.ky_a = (u32 *)take(M * 4);
  D.grank_a = (u32 *)take(M * 4);
  D.sky_a = (u32 *)take(M * 4);
  D.ry_a = (u32 *)take(M * 4);
  D.parent_x = (u32 *)take(MC * 4);
  D.xm_bits = (u32 *)take((MC + 31) / 32 * 4);
  D.parent = (u32 *)take(M * 4);
  D.gidscan = (u32 *)take(M * 4);
  D.lroot = (u32 *)take(M * 4);
  D.gid_l = (u32 *)take(M * 4);
  for (int b = 0; b < 2; ++b) D.pend_idx[b] = (u32 *)take(M * 4), D.pend_key[b] = (u32 *)take(M * 4);
  D.n_pend = (u32 *)take(16);
  D.apos = (u32 *)take(M * 4);
  D.roff_dev = (u32 *)take((DIST_MAX_RANKS + 1) * 4);
  D.queries = (uint4 *)take(M * 16);
  D.answers = (u64 *)take(M * 8);
  D.res = (u64 *)take(M * 8);
  D.flag = (u32 *)take(16);
  D.flag_all = (u32 *)take(DIST_MAX_RANKS * 4);
  D.parent_y = (u32 *)take(M * 4);
  D.yo_s = (u32 *)take(M * 4);
  D.xm_a = take(M);
  D.ent_rank = (u32 *)take(MC * 4);
  D.ent_c = (u32 *)take(MC * 4);
  D.ent_len = (u32 *)take(MC * 4);
  D.work_cap = (u32)(MC / 32 + 2);
  D.worklist = (u32 *)take((u64)D.work_cap * 4);
  D.gid_rank = (u32 *)take(M * 4);
  D.scan_work = take(dist_scan_work_bytes((u32)M));
  D.gid_a = (u32 *)take(M * 4);
  D.sgid = (u32 *)take(M * 4);
  D.srank_g = (u32 *)take(M * 4);
  D.order_scratch = take(order_scratch_bytes(M));
  D.out_order = (u32 *)take(M * 4);
  D.out_gid = (u32 *)take(M * 4);
  D.out_repval = take(M);
  D.out_identity = (float *)take(M * 4);
  return off;
}

static void dist_release_buffers(Dist &D) {
  for (int r = 0; r < DIST_MAX_RANKS; ++r) {
    if (D.peer_open[r]) cudaIpcCloseMemHandle(D.peer_open[r]);
    D.peer_open[r] = nullptr;
  }
  D.peers_ready = false;
  if (D.arena) cudaFree(D.arena);
  D.arena = nullptr;
  if (D.tr) D.tr->peers_mapped = false;
  D.cap = D.hcap = 0;
  D.loaded = false;
}

// (re)allocates the partition buffers for cap rows per rank; the peers must import again afterwards
static int dist_allocate(rk_ctx *ctx, u64 cap) {
  Dist &D = *ctx->dist;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  dist_release_buffers(D);
  if (cap < 1024) cap = 1024;
  if (cap >= 0xFFFFFF00ull) return fail(ctx, RK_ERR_ARG, "capacity per rank must stay below 2^32");
  D.cap = cap;
  D.hcap = cap / 8 + 65536;
  const u64 need = dist_carve(D, nullptr);
  cudaError_t e = cudaMalloc(&D.arena, need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    dist_release_buffers(D);
    return fail(ctx, RK_ERR_NOMEM, "cudaMalloc(%llu bytes for %llu rows per rank): %s", (unsigned long long)need,
                (unsigned long long)cap, cudaGetErrorString(e));
  }
  dist_carve(D, (u8 *)D.arena);
  return RK_OK;
}

static int tr_fail(rk_ctx *ctx, int rc) { return fail(ctx, rc, "%s", ctx->dist->tr->err.c_str()); }
#define TR(call)                                  \
  do {                                            \
    D.trace.begin(#call, ctx->stream);            \
    const int rc_ = (call);                       \
    D.trace.end(ctx->stream);                     \
    if (rc_ != RK_OK) return tr_fail(ctx, rc_);   \
  } while (0)

static int dist_init_common(rk_ctx *ctx, int rank, int world, Transport *tr, u64 cap) {
  if (world < 1 || world > DIST_MAX_RANKS || rank < 0 || rank >= world) {
    delete tr;
    return fail(ctx, RK_ERR_ARG, "rank %d of %d: at most %d ranks", rank, world, DIST_MAX_RANKS);
  }
  dist_destroy(ctx);
  Dist *D = new Dist;
  D->rank = rank, D->world = world, D->tr = tr;
  tr->rank = rank, tr->world = world;
  ctx->dist = D;
  D->trace.on = getenv("RK_DIST_TRACE") != nullptr;
  CK(cudaSetDevice(ctx->device));
  CK(cudaHostAlloc((void **)&D->h_small, (size_t)world * SMALL_WORDS * sizeof(u32), cudaHostAllocDefault));
  CK(cudaStreamCreateWithFlags(&D->side, cudaStreamNonBlocking));
  return dist_allocate(ctx, cap);
}

static int dist_export(rk_ctx *ctx, DistBlob *b) {
  Dist &D = *ctx->dist;
  memset(b, 0, sizeof *b);
  b->pid = (u64)getpid(), b->ptr = (u64)(uintptr_t)D.arena, b->cap = D.cap, b->device = ctx->device;
  CK(cudaSetDevice(ctx->device));
  CK(cudaIpcGetMemHandle(&b->handle, D.arena));
  return RK_OK;
}

static int dist_import(rk_ctx *ctx, const u8 *blobs, size_t stride) {
  Dist &D = *ctx->dist;
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < D.world; ++r) {
    DistBlob b;
    memcpy(&b, blobs + (size_t)r * stride, sizeof b);
    if (b.cap != D.cap) return fail(ctx, RK_ERR_ARG, "rank %d was initialised with capacity %llu, this rank with %llu", r,
                                    (unsigned long long)b.cap, (unsigned long long)D.cap);
    u8 *base = nullptr;
    if (r == D.rank) {
      base = (u8 *)D.arena;
    } else if (b.pid == (u64)getpid()) {
      base = (u8 *)(uintptr_t)b.ptr;
      if (b.device != ctx->device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(ctx, RK_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctx->device, b.device, cudaGetErrorString(e));
        cudaGetLastError();
      }
    } else {
      if (D.peer_open[r]) cudaIpcCloseMemHandle(D.peer_open[r]);
      D.peer_open[r] = nullptr;
      CK(cudaIpcOpenMemHandle((void **)&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
      D.peer_open[r] = base;
    }
    D.tr->peer_base[r] = base;
    // the arenas are carved identically (same capacity, same number of ranks): local offsets hold on every peer
    D.pt.res[r] = (const u64 *)(base + ((const u8 *)D.res - (const u8 *)D.arena));
  }
  D.tr->my_base = (const u8 *)D.arena;
  D.tr->peers_mapped = true;
  if (NcclTransport *nt = dynamic_cast<NcclTransport *>(D.tr)) nt->d_flag = D.flag, nt->d_flag_all = D.flag_all;
  D.pt.nr = D.world, D.pt.me = D.rank;
  D.peers_ready = true;
  return RK_OK;
}

// the count exchange with the matrix left on the device as well (d_small_all) and on the host (h_small); synchronises
static int dist_gather_counts_dev(rk_ctx *ctx) {
  Dist &D = *ctx->dist;
  CK(cudaMemcpyAsync(D.d_small + W_ERR, &D.cnt->err, sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
  TR(D.tr->gather_small_begin(D.d_small, D.d_small_all, D.h_small, ctx->stream));
  TR(D.tr->gather_small_end(D.d_small_all, D.h_small, ctx->stream));
  return RK_OK;
}

// error words of all ranks after a count exchange: every rank sees the same matrix, so every rank returns together
static int dist_check_small(rk_ctx *ctx, bool range_errors) {
  Dist &D = *ctx->dist;
  u32 e = 0;
  for (int r = 0; r < D.world; ++r) e |= D.h_small[(size_t)r * SMALL_WORDS + W_ERR];
  if (!e) return RK_OK;
  return fail(ctx, (range_errors && !(e & (ERR_WORKLIST | ERR_SPIN))) ? RK_ERR_RANGE : RK_ERR_INTERNAL, "%s", err_bits_text(e));
}

// every rank's copy of one of this rank's arena buffers (the rows of the big exchanges are stored there by the kernels)
static void peer_rows(const Dist &D, uint4 *local, uint4 **outs);

// block table of a per-pair exchange whose values the producing kernel stores straight into the peers' buffers:
// forward = in send order (blocks by destination, landing where the rows of the load-time exchange landed),
// back = in arrival order (blocks by source, landing in the order this rank's rows were sent)
template <class T>
static ScatterTable scatter_table(const Dist &D, const Exchange &ex, bool back, T *local_buf) {
  ScatterTable t{};
  t.nr = D.world;
  u64 run = 0;
  for (int d = 0; d < D.world; ++d) {
    t.start[d] = (u32)run;
    run += back ? ex.rcnt[d] : ex.scnt[d];
    t.dst_off[d] = (u32)(back ? ex.rpoff[d] : ex.poff[d]);
    t.out[d] = d == D.rank ? (void *)local_buf : (void *)D.tr->on_peer(d, local_buf);
  }
  t.start[D.world] = (u32)run;
  return t;
}

static int dist_gather_counts(rk_ctx *ctx) {
  Dist &D = *ctx->dist;
  CK(cudaMemcpyAsync(D.d_small + W_ERR, &D.cnt->err, sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
  TR(D.tr->gather_small(D.d_small, D.d_small_all, D.h_small, ctx->stream));
  return RK_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// load: FragmentsDatabase's constructor over the ranks (src/FragmentsDatabase.cpp:84-100)
// ---------------------------------------------------------------------------------------------------------------------
static int dist_load(rk_ctx *ctx, const void *frags, u64 n_loc, u64 file_off, u64 seqx_len, u64 seqy_len, unsigned flags,
                     rk_load_stats *stats) {
  Dist &D = *ctx->dist;
  const int nr = D.world, me = D.rank;
  if (!D.peers_ready) return fail(ctx, RK_ERR_STATE, "rk_dist_load_aos before rk_dist_import");
  if (!frags && n_loc) return fail(ctx, RK_ERR_ARG, "null record pointer");
  if (seqx_len >= (1ull << 32) || seqy_len >= (1ull << 32)) return fail(ctx, RK_ERR_RANGE, "sequence length does not fit in 32 bits");
  if (file_off + n_loc >= 0xFFFFFFF0ull) return fail(ctx, RK_ERR_ARG, "more than 2^32-16 records in one comparison");
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  cudaStream_t st = ctx->stream;
  D.loaded = false;
  ctx->have_group = false;
  // NOTE: argument errors that only ONE rank can see (n_loc > cap) are turned into an error word all ranks receive with
  // the first count exchange, so that no rank is left waiting in a collective.
  const bool too_many = n_loc > D.cap;
  const u64 n_use = too_many ? 0 : n_loc;

  const Geometry g = make_geometry(seqx_len, seqy_len);
  D.g = g;
  D.bits_rank = ceil_log2(g.vsize) < 1 ? 1 : ceil_log2(g.vsize);
  D.bits_y = ceil_log2(2ull * g.nby);
  const u64 lxw = (2ull * g.nbx + 31) / 32 + 1, lyw = (2ull * g.nby + 31) / 32 + 1, lmax = lxw > lyw ? lxw : lyw;
  const u64 pub_words = DIST_BINS + lxw + lyw;  // what a rank publishes after K1: its xStart/10 histogram and its link maps
  const u64 link_need = pub_words + lxw + lyw + (u64)nr * pub_words + 64;
  (void)lmax;
  if (link_need > D.link_words) {
    CK(cudaStreamSynchronize(st));
    if (D.link_buf) cudaFree(D.link_buf);
    D.link_buf = nullptr, D.link_words = 0;
    CK(cudaMalloc((void **)&D.link_buf, link_need * 4));
    D.link_words = link_need;
  }
  u32 *hist0 = D.link_buf, *linkx_loc = hist0 + DIST_BINS, *linky_loc = linkx_loc + lxw, *linkx = linky_loc + lyw, *linky = linkx + lxw,
      *pub_all = linky + lyw;

  bool on_device = false;
  if (n_use) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, frags) == cudaSuccess) on_device = pa.type == cudaMemoryTypeDevice;
    else cudaGetLastError();
  }
  if (on_device && ((uintptr_t)frags & 15)) return fail(ctx, RK_ERR_ARG, "device record pointer must be 16-byte aligned");
  const u8 *aos = (const u8 *)frags;
  cudaEvent_t *ev = ctx->ev;
  CK(cudaEventRecord(ev[0], st));
  if (!on_device && n_use) {
    const u64 need = align_up(n_use * RK_FRAG_BYTES, 16) + 16;
    if (need > D.aos_bytes) {
      CK(cudaStreamSynchronize(st));
      if (D.aos_buf) cudaFree(D.aos_buf);
      D.aos_buf = nullptr, D.aos_bytes = 0;
      CK(cudaMalloc((void **)&D.aos_buf, need));
      D.aos_bytes = need;
    }
    CK(cudaMemcpyAsync(D.aos_buf, frags, n_use * RK_FRAG_BYTES, cudaMemcpyHostToDevice, st));
    aos = D.aos_buf;
  }
  CK(cudaEventRecord(ev[1], st));
  u64 launches = 0;

  // K1 on the local slice of the file
  CK(cudaMemsetAsync(D.cnt, 0, sizeof(Counters), st));
  CK(cudaMemsetAsync(D.d_small, 0, SMALL_WORDS * 4, st));
  CK(cudaMemsetAsync(D.prehist, 0, 3 * 4 * 256 * 4, st));
  CK(cudaMemsetAsync(linkx_loc, 0, (lxw + lyw) * 4, st));
  auto hist_of = [&](int which, int bits) { return HistOut{D.prehist + which * 1024, (bits + 7) / 8, bits}; };
  if (too_many) {
    const u32 e = ERR_WORKLIST;  // reported below as a capacity error
    CK(cudaMemcpyAsync(&D.cnt->err, &e, 4, cudaMemcpyHostToDevice, st));
  }
  launches += launch_decode(aos, n_use, g, nullptr, nullptr, nullptr, nullptr, nullptr, D.key0_loc, linkx_loc, linky_loc,
                            &D.cnt->n_dropped, &D.cnt->err, st, D.rec4_loc, HistOut{nullptr, 0, 0}, (u32)file_off);
  // exchange 1: to the owner of the xStart/10 range
  const int shift0 = D.bits_rank > 12 ? D.bits_rank - 12 : 0;
  launches += dist_coarse_hist(D.key0_loc, (u32)n_use, shift0, 0, g.vsize - 1, hist0, st);
  // one all-gather for everything K1 produced that the other ranks need: the histogram the cuts are made from and the link
  // maps (OR-ed over the ranks below, so that a run of linked buckets has one key everywhere)
  TR(D.tr->all_gather(hist0, pub_all, pub_words * 4, st));
  launches += dist_cuts_from_hist(pub_all, pub_words, nr, shift0, D.cuts0, st);
  launches += dist_or_rows(pub_all + DIST_BINS, pub_words, nr, lxw, linkx, st);
  launches += dist_or_rows(pub_all + DIST_BINS + lxw, pub_words, nr, lyw, linky, st);
  // count per destination -> all ranks' counts on the device -> every tile stores its records straight into the owners'
  // receive buffers (peer memory over NVLink) -> barrier; the host reads the count matrix while the rows travel
  uint2 *outs[DIST_MAX_RANKS];
  for (int d = 0; d < nr; ++d) outs[d] = d == me ? D.rec6_arr : D.tr->on_peer(d, D.rec6_arr);
  launches += dist_count_plain(D.key0_loc, (u32)n_use, D.cuts0, nr, g.vsize - 1, D.tile_cnt, D.d_small + W_CNT_A, st);
  CK(cudaMemcpyAsync(D.d_small + W_ERR, &D.cnt->err, sizeof(u32), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(D.d_small + W_CUTS, D.cuts0, (nr + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  TR(D.tr->gather_small_begin(D.d_small, D.d_small_all, D.h_small, st));
  launches += dist_push_records(D.key0_loc, (u32)n_use, D.cuts0, nr, g.vsize - 1, D.rec4_loc, outs, (u32)D.cap, D.tile_cnt, D.d_small_all,
                                SMALL_WORDS, me, st);
  TR(D.tr->barrier(st));
  TR(D.tr->gather_small_end(D.d_small_all, D.h_small, st));  // host sync 1
  if (too_many || dist_check_small(ctx, true) != RK_OK) {
    u32 e = 0;
    for (int r = 0; r < nr; ++r) e |= D.h_small[(size_t)r * SMALL_WORDS + W_ERR];
    if (e & ERR_WORKLIST) return fail(ctx, RK_ERR_NOMEM, "a rank holds more records than the capacity of %llu rows per rank given to rk_dist_init",
                                      (unsigned long long)D.cap);
    return dist_check_small(ctx, true);
  }
  D.ex1.from_matrix(D.h_small, W_CNT_A, nr, me);
  D.tr->bytes_sent += (D.ex1.n_send - D.ex1.scnt[me]) * 24;  // (the rows this rank's kernels stored into peer memory)
  u64 m_total = 0, loaded_total = 0;
  u64 m_of[DIST_MAX_RANKS] = {0};
  for (int s = 0; s < nr; ++s) {
    for (int d = 0; d < nr; ++d) m_of[d] += D.h_small[(size_t)s * SMALL_WORDS + W_CNT_A + d];
    for (int d = 0; d <= nr; ++d) loaded_total += D.h_small[(size_t)s * SMALL_WORDS + W_CNT_A + d];
  }
  u64 off = 0;
  for (int d = 0; d < nr; ++d) {
    D.pt.roff[d] = (u32)off;
    if (d == me) D.rank_off = (u32)off;
    off += m_of[d];
    if (m_of[d] > D.cap) return fail(ctx, RK_ERR_NOMEM, "rank %d would own %llu fragments, capacity is %llu rows per rank", d,
                                     (unsigned long long)m_of[d], (unsigned long long)D.cap);
  }
  m_total = off;
  D.pt.roff[nr] = (u32)m_total;
  for (int d = nr + 1; d <= DIST_MAX_RANKS; ++d) D.pt.roff[d] = 0xFFFFFFFFu;
  CK(cudaMemcpyAsync(D.roff_dev, D.pt.roff, (nr + 1) * sizeof(u32), cudaMemcpyHostToDevice, st));  // (pageable source: staged at once)
  D.m_total = m_total, D.n_total_loaded = loaded_total;
  const u32 m = (u32)m_of[me];
  D.m_loc = m;

  CK(cudaEventRecord(ev[2], st));
  // processing order: sources arrive in file order, so a stable sort by xStart/10 is the global order
  // the sort keys are taken relative to the rank's first xStart/10 bucket: log2(range of the rank) bits instead of log2(vsize)
  const u32 *cuts0_h = D.h_small + (size_t)me * SMALL_WORDS + W_CUTS;
  const u32 key0_base = cuts0_h[me];
  const u64 key0_end = (me + 1 < nr && cuts0_h[me + 1] < g.vsize) ? cuts0_h[me + 1] : g.vsize;
  const int bits_rank_l = ceil_log2(key0_end > key0_base ? key0_end - key0_base : 1) < 1 ? 1 : ceil_log2(key0_end - key0_base);
  launches += dist_key0_of_rec(D.rec6_arr, m, key0_base, D.key0a, hist_of(0, bits_rank_l), st);
  launches += launch_sort_pairs(D.key0a, nullptr, D.k0_r, D.aidx_r, D.tmp_k, D.tmp_v, m, bits_rank_l, D.sort_work, st, &D.cnt->err,
                                m ? D.prehist : nullptr);
  CK(cudaEventRecord(ev[3], st));
  launches += launch_keys(D.aidx_r, m, g, nullptr, linkx, linky, D.xl, D.yl_r, D.ys_r, D.kx2, D.ky, D.identity_r, st,
                          HistOut{nullptr, 0, 0}, HistOut{nullptr, 0, 0}, D.gfidx_r, 1u, D.rec6_arr);
  launches += launch_hkey(D.k0_r, D.ys_r, m, nullptr, st, D.gfidx_r, D.identity_r, D.hfi_r);
  CK(cudaEventRecord(ev[4], st));
  // X halo: fragments whose X super-bucket belongs to a higher rank
  launches += dist_cuts_x(D.cuts0, nr, g, linkx, D.cuts_x, st);
  launches += dist_split_halo(D.kx2, D.xl, m, D.cuts_x, nr, g.nbx, me, D.rank_off, D.halo_send, (u32)D.hcap, D.perm_x, D.tile_cnt,
                              D.d_small + W_CNT_A, st);
  // Y: to the owner of the Y super-bucket range
  const int shift_y = D.bits_y > 12 ? D.bits_y - 12 : 0;
  launches += dist_coarse_hist(D.ky, m, shift_y, 0, 0xFFFFFFFFu, D.hist, st);
  TR(D.tr->all_gather(D.hist, D.hist_all, DIST_BINS * 4, st));
  launches += dist_cuts_from_hist(D.hist_all, DIST_BINS, nr, shift_y, D.cuts_y, st);
  launches += dist_split_axis(D.ky, D.yl_r, m, D.cuts_y, nr, D.rank_off, D.send_rows, D.perm_y, D.tile_cnt, D.d_small + W_CNT_B, st);
  CK(cudaMemcpyAsync(D.d_small + W_CUTS, D.cuts_y, (nr + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(D.d_small + W_CUTSX, D.cuts_x, 2 * nr * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  {
    const int rc = dist_gather_counts(ctx);  // host sync 2
    if (rc) return rc;
    const int rc2 = dist_check_small(ctx, true);
    if (rc2) return rc2;
  }
  D.exh.from_matrix(D.h_small, W_CNT_A, nr, me);
  D.exy.from_matrix(D.h_small, W_CNT_B, nr, me);
  for (int d = 0; d < nr; ++d) {
    u64 halo_in = 0, halo_out = 0, y_in = 0;
    for (int s = 0; s < nr; ++s) {
      halo_in += D.h_small[(size_t)s * SMALL_WORDS + W_CNT_A + d];
      y_in += D.h_small[(size_t)s * SMALL_WORDS + W_CNT_B + d];
    }
    for (int t = 0; t < nr; ++t) halo_out += D.h_small[(size_t)d * SMALL_WORDS + W_CNT_A + t];
    if (halo_in > D.hcap || halo_out > D.hcap)
      return fail(ctx, RK_ERR_NOMEM, "rank %d: %llu halo fragments in / %llu out, the halo capacity is %llu", d,
                  (unsigned long long)halo_in, (unsigned long long)halo_out, (unsigned long long)D.hcap);
    if (y_in > D.cap) return fail(ctx, RK_ERR_NOMEM, "rank %d would own %llu fragments in the Y pass, capacity is %llu rows per rank", d,
                                  (unsigned long long)y_in, (unsigned long long)D.cap);
  }
  D.n_away = (u32)D.exh.n_send, D.n_halo = (u32)D.exh.n_recv, D.m_y = (u32)D.exy.n_recv;
  // Y sort keys relative to the first key of the rank's Y range
  const u32 *cutsy_h = D.h_small + (size_t)me * SMALL_WORDS + W_CUTS;
  const u32 ky_base = cutsy_h[me];
  const u64 ky_end = (me + 1 < nr && cutsy_h[me + 1] < 2ull * g.nby) ? cutsy_h[me + 1] : 2ull * g.nby;
  const int bits_y_l = ceil_log2(ky_end > ky_base ? ky_end - ky_base : 1) < 1 ? 1 : ceil_log2(ky_end - ky_base);

  TR(D.tr->all_to_all(D.halo_send, D.exh.soff, D.exh.scnt, D.halo_recv, D.exh.roff, D.exh.rcnt, D.exh.poff, 16, st));
  // the Y rows travel (copy engines, NVLink) while this GPU unpacks its halo and sorts its X buckets
  TR(D.tr->a2a_begin(D.send_rows, D.exy.soff, D.exy.scnt, D.recv_rows, D.exy.roff, D.exy.rcnt, D.exy.poff, 16, st, D.side));
  // X sort keys relative to the first super-bucket of the rank's range (per strand class)
  const u32 *cutsx_h = D.h_small + (size_t)me * SMALL_WORDS + W_CUTSX;
  u32 x_base[2], x_range[2];
  for (int sc = 0; sc < 2; ++sc) {
    x_base[sc] = cutsx_h[sc * nr + me];
    const u32 end = me + 1 < nr ? cutsx_h[sc * nr + me + 1] : (u32)(sc + 1) * g.nbx;
    x_range[sc] = end > x_base[sc] ? end - x_base[sc] : 1u;
  }
  const u32 x_R = std::max(x_range[0], x_range[1]);
  const int bits_x_l = ceil_log2(4ull * ((u64)x_R + 1));
  launches += dist_x_local_keys(D.kx2, m, g.nbx, x_base, x_range, hist_of(1, bits_x_l), st);
  launches += dist_unpack_halo_rows(D.halo_recv, D.n_halo, g.nbx, x_base, x_range, D.kx2 + m, D.xl + m, D.halo_grank, hist_of(1, bits_x_l), st);
  CK(cudaEventRecord(ev[5], st));
  launches += launch_sort_pairs(D.kx2, nullptr, D.skx, D.rx, D.tmp_k, D.tmp_v, (u64)m + D.n_halo, bits_x_l, D.sort_work, st, &D.cnt->err,
                                (m + D.n_halo) ? D.prehist + 1024 : nullptr);
  CK(cudaEventRecord(ev[6], st));
  TR(D.tr->a2a_end(st));
  launches += dist_unpack_axis_rows(D.recv_rows, D.m_y, ky_base, D.ky_a, D.yl_a, D.grank_a, hist_of(2, bits_y_l), st);
  launches += launch_sort_pairs(D.ky_a, nullptr, D.sky_a, D.ry_a, D.tmp_k, D.tmp_v, D.m_y, bits_y_l, D.sort_work, st, &D.cnt->err,
                                D.m_y ? D.prehist + 2048 : nullptr);
  CK(cudaEventRecord(ev[7], st));
  D.loaded = true;

  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->n_loaded = loaded_total;
    stats->n_kept = m_total;
    stats->vsize = g.vsize;
    stats->n_launches = launches;
    if (flags & RK_F_TIMING) {
      CK(cudaStreamSynchronize(st));
      stats->ms_stage[RK_ST_H2D] = ev_ms(ev[0], ev[1]);
      stats->ms_stage[RK_ST_DECODE] = ev_ms(ev[1], ev[2]);  // K1 + exchange 1
      stats->ms_stage[RK_ST_RANKSORT] = ev_ms(ev[2], ev[3]);
      stats->ms_stage[RK_ST_KEYS] = ev_ms(ev[3], ev[4]);
      stats->ms_stage[RK_ST_XSORT] = ev_ms(ev[4], ev[6]);   // routing of both axes, halo exchange, X sort
      stats->ms_stage[RK_ST_YSORT] = ev_ms(ev[6], ev[7]);   // Y exchange + Y sort
      stats->ms_device = ev_ms(ev[1], ev[7]);
    }
  }
  return RK_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// group: generate_fragment_groups + generate_diagonal_func + sort_groups over the ranks
// ---------------------------------------------------------------------------------------------------------------------
static int dist_group(rk_ctx *ctx, double len_ratio, double pos_ratio, unsigned flags, rk_result *out, rk_dist_info *info) {
  Dist &D = *ctx->dist;
  const int nr = D.world, me = D.rank;
  if (!D.loaded) return fail(ctx, RK_ERR_STATE, "rk_dist_group before a successful rk_dist_load_aos");
  if (!(len_ratio > 0)) return fail(ctx, RK_ERR_ARG, "Ratio between length and position must be greater than zero");
  if (!(pos_ratio > 0)) return fail(ctx, RK_ERR_ARG, "Position proximity must be greater than zero");
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  memset(out, 0, sizeof *out);
  cudaStream_t st = ctx->stream;
  cudaEvent_t *ev = ctx->ev;
  const u32 m = D.m_loc, nh = D.n_halo, mc = m + nh;
  u64 launches = 0;
  const u64 sent0 = D.tr->bytes_sent;
  CK(cudaMemsetAsync(D.d_small, 0, SMALL_WORDS * 4, st));

  CK(cudaEventRecord(ev[0], st));
  // X pass at home over [own fragments ++ halo]
  MatchArgs mx{};
  mx.skey = D.skx, mx.srank = D.rx, mx.cl_r = D.xl, mx.parent = D.parent_x, mx.xm_bits = D.xm_bits;
  mx.m = mc, mx.max_index = D.g.mx, mx.len_ratio = len_ratio, mx.pos_ratio = pos_ratio, mx.is_y = 0;
  mx.worklist = D.worklist, mx.work_count = D.cnt->work_x, mx.work_cap = D.work_cap;
  mx.ent_rank = D.ent_rank, mx.ent_c = D.ent_c, mx.ent_len = D.ent_len, mx.err = &D.cnt->err;
  mx.key_shift = 1;
  launches += launch_match(mx, st);
  launches += dist_x_owners(D.parent_x, m, nh, D.rank_off, D.halo_grank, D.parent, scatter_table(D, D.exh, true, D.away_res), st);
  TR(D.tr->barrier(st));  // the halo owners are stored by the kernel above straight into their home ranks' away_res
  launches += dist_apply_away(D.away_res, D.perm_x, D.n_away, D.parent, st);
  CK(cudaEventRecord(ev[1], st));
  // Y pass on the owners of the Y ranges: X-matched fragments insert without a query (commonFunctions.cpp:59)
  launches += dist_pack_xm(D.parent, D.perm_y, m, scatter_table(D, D.exy, false, D.xm_a), st);
  if (m) CK(cudaMemsetAsync(D.yo_s, 0xFF, (size_t)m * 4, st));  // "no Y owner" until a peer stores one (after the barrier)
  TR(D.tr->barrier(st));
  if (D.m_y) CK(cudaMemsetAsync(D.parent_y, 0xFF, (size_t)D.m_y * 4, st));
  MatchArgs my = mx;
  my.skey = D.sky_a, my.srank = D.ry_a, my.cl_r = D.yl_a, my.parent = D.parent_y, my.m = D.m_y, my.max_index = D.g.my, my.is_y = 1;
  my.work_count = D.cnt->work_y, my.key_shift = 0, my.xm_bytes = D.xm_a;
  launches += launch_match(my, st);
  launches += dist_y_owners(D.parent_y, D.grank_a, D.m_y, scatter_table(D, D.exy, true, D.yo_s), st);
  TR(D.tr->barrier(st));
  launches += dist_merge_y(D.yo_s, D.perm_y, m, D.parent, st);
  CK(cudaEventRecord(ev[2], st));
  // forest: roots per rank; local chains; the chains that leave the GPU are resolved in bulk rounds while there are many
  // of them anywhere, the rest by walking through the peers' memory (k7_dist.cu)
  launches += dist_root_scan(D.parent, m, D.gidscan, D.d_small + W_X0, D.scan_work, st);
  launches += dist_chase_local(D.parent, D.gidscan, m, D.rank_off, D.lroot, D.res, D.pend_idx[0], D.pend_key[0], D.n_pend, st);
  CK(cudaMemcpyAsync(D.d_small + W_X2, D.n_pend, sizeof(u32), cudaMemcpyDeviceToDevice, st));
  {
    const int rc = dist_gather_counts_dev(ctx);  // host sync: roots and pending chains of every rank (also: every rank's res[] is final)
    if (rc) return rc;
  }
  // the root counts stay on the device as a column of the gathered matrix, in a copy that later gathers do not overwrite
  CK(cudaMemcpyAsync(D.nroots_mat, D.d_small_all, (size_t)nr * SMALL_WORDS * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  const u32 *nroots = D.nroots_mat + W_X0;
  // (RK_DIST_BULK_MIN: tuning and tests — 1 resolves every chain in bulk rounds, a huge value only by walking)
  const u64 bulk_min = getenv("RK_DIST_BULK_MIN") ? strtoull(getenv("RK_DIST_BULK_MIN"), nullptr, 10) : (1ull << 20);
  int cur = 0;
  for (int round = 0; round < nr; ++round) {
    u64 most = 0;
    for (int r = 0; r < nr; ++r) most = std::max<u64>(most, D.h_small[(size_t)r * SMALL_WORDS + W_X2]);
    if (most == 0 || most < bulk_min) break;
    const u32 bound = (u32)std::min<u64>(D.h_small[(size_t)me * SMALL_WORDS + W_X2], m);  // this rank's pending chains
    // ask: the ranks in question go to their owners
    uint4 *qouts[DIST_MAX_RANKS];
    peer_rows(D, D.queries, qouts);
    launches += dist_chase_ask_count(D.pend_key[cur], bound, D.n_pend, D.roff_dev, nr, D.tile_cnt, D.d_small + W_CNT_A, st);
    CK(cudaMemcpyAsync(D.d_small + W_ERR, &D.cnt->err, sizeof(u32), cudaMemcpyDeviceToDevice, st));
    TR(D.tr->gather_small_begin(D.d_small, D.d_small_all, D.h_small, st));
    launches += dist_chase_ask_push(D.pend_key[cur], bound, D.n_pend, D.roff_dev, nr, qouts, (u32)D.cap, D.apos, D.tile_cnt, D.d_small_all,
                                    SMALL_WORDS, me, st);
    TR(D.tr->barrier(st));
    TR(D.tr->gather_small_end(D.d_small_all, D.h_small, st));
    Exchange exq;
    exq.from_matrix(D.h_small, W_CNT_A, nr, me);
    for (int d = 0; d < nr; ++d) {
      u64 in = 0;
      for (int sr = 0; sr < nr; ++sr) in += D.h_small[(size_t)sr * SMALL_WORDS + W_CNT_A + d];
      if (in > D.cap) return fail(ctx, RK_ERR_NOMEM, "rank %d is asked about %llu chain ends, capacity is %llu rows per rank", d,
                                  (unsigned long long)in, (unsigned long long)D.cap);
    }
    D.tr->bytes_sent += (exq.n_send - exq.scnt[me]) * 16 + (exq.n_recv - exq.rcnt[me]) * 8;
    // answer: every owner returns its res[] words, block by block, into the asking ranks' answer buffers
    launches += dist_chase_answer(D.queries, (u32)exq.n_recv, D.res, D.rank_off, scatter_table(D, exq, true, D.answers), st);
    TR(D.tr->barrier(st));
    // apply: roots end a chain, anything else is asked about in the next round
    launches += dist_chase_apply(D.pt, nroots, SMALL_WORDS, bound, D.pend_idx[cur], D.pend_key[cur], D.apos, D.answers, D.n_pend, D.gid_l,
                                 D.pend_idx[cur ^ 1], D.pend_key[cur ^ 1], D.n_pend + 1, st);
    CK(cudaMemcpyAsync(D.n_pend, D.n_pend + 1, sizeof(u32), cudaMemcpyDeviceToDevice, st));
    cur ^= 1;
    CK(cudaMemcpyAsync(D.d_small + W_X2, D.n_pend, sizeof(u32), cudaMemcpyDeviceToDevice, st));
    const int rc = dist_gather_counts_dev(ctx);  // how many chains are still open, everywhere
    if (rc) return rc;
  }
  launches += dist_chase_finish(D.pt, nroots, SMALL_WORDS, m, D.res, D.lroot, D.pend_idx[cur], D.pend_key[cur], D.n_pend, D.gid_l, D.gid_rank, st);
  // (no second barrier: the count exchange of the output stage below completes on a rank only after every rank has
  // entered it, i.e. finished chasing, and parent[] is not written again before the next rk_dist_group)
  CK(cudaEventRecord(ev[3], st));
  // output exchange: to the owner of the group-id range.  The ranges are cut for about equal sort_groups WORK: lines, with
  // the lines of groups with several members weighted up (groups founded early are the large ones: equal numbers of
  // groups gave rank 0 far more lines, equal numbers of lines still far more work — see k_cuts_from_hist).
  const int bits_gid = ceil_log2(D.m_total) < 1 ? 1 : ceil_log2(D.m_total);  // group ids are < number of fragments
  const int shift_g = bits_gid > 12 ? bits_gid - 12 : 0;
  launches += dist_cuts_gid(nroots, SMALL_WORDS, nr, D.cuts_g, D.d_small + W_X1, st);  // (the total; the cuts are replaced below)
  launches += dist_coarse_hist(D.gid_rank, m, shift_g, 0, 0xFFFFFFFFu, D.hist, st);
  TR(D.tr->all_gather(D.hist, D.hist_all, DIST_BINS * 4, st));
  // no range may hold more lines than a rank has rows (RK_DIST_LINE_CAP: tests — a tighter bound than the capacity)
  const u64 line_cap = getenv("RK_DIST_LINE_CAP") ? std::min<u64>(D.cap, strtoull(getenv("RK_DIST_LINE_CAP"), nullptr, 10)) : D.cap;
  launches += dist_cuts_from_hist(D.hist_all, DIST_BINS, nr, shift_g, D.cuts_g, st, D.d_small + W_X1, line_cap);
  CK(cudaMemcpyAsync(D.d_small + W_CNT_B, D.cuts_g, (nr + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, st));
  uint4 *outs[DIST_MAX_RANKS];
  peer_rows(D, D.recv_rows, outs);
  launches += dist_count_plain(D.gid_rank, m, D.cuts_g, nr, 0xFFFFFFFFu, D.tile_cnt, D.d_small + W_CNT_A, st);
  CK(cudaMemcpyAsync(D.d_small + W_ERR, &D.cnt->err, sizeof(u32), cudaMemcpyDeviceToDevice, st));
  TR(D.tr->gather_small_begin(D.d_small, D.d_small_all, D.h_small, st));
  launches += dist_push_gid(D.gid_rank, D.hfi_r, m, D.cuts_g, nr, outs, (u32)D.cap, D.tile_cnt, D.d_small_all, SMALL_WORDS, me, st);
  TR(D.tr->barrier(st));
  TR(D.tr->gather_small_end(D.d_small_all, D.h_small, st));  // host sync 3
  {
    const int rc2 = dist_check_small(ctx, false);
    if (rc2) return rc2;
  }
  Exchange exg;
  exg.from_matrix(D.h_small, W_CNT_A, nr, me);
  D.tr->bytes_sent += (exg.n_send - exg.scnt[me]) * 16;
  const u64 total_groups = D.h_small[(size_t)me * SMALL_WORDS + W_X1];
  u64 line_off = 0;
  for (int d = 0; d < nr; ++d) {
    u64 in = 0;
    for (int s = 0; s < nr; ++s) in += D.h_small[(size_t)s * SMALL_WORDS + W_CNT_A + d];
    if (in > D.cap) return fail(ctx, RK_ERR_NOMEM, "rank %d would own %llu output lines, capacity is %llu rows per rank", d,
                                (unsigned long long)in, (unsigned long long)D.cap);
    if (d < me) line_off += in;
  }
  const u32 mg = (u32)exg.n_recv;
  const u32 *cuts_h = D.h_small + (size_t)me * SMALL_WORDS + W_CNT_B;
  const u32 gid_base = cuts_h[me];
  const u64 gid_end = (u64)cuts_h[me + 1] < total_groups ? (u64)cuts_h[me + 1] : total_groups;
  const u64 local_groups = gid_end > gid_base ? gid_end - gid_base : 1;
  const int bits_g = ceil_log2(local_groups) < 1 ? 1 : ceil_log2(local_groups);
  CK(cudaMemsetAsync(D.prehist + 3072, 0, 4 * 256 * 4, st));
  launches += dist_gid_keys(D.recv_rows, mg, gid_base, D.gid_a, HistOut{D.prehist + 3072, (bits_g + 7) / 8, bits_g}, st);
  CK(cudaEventRecord(ev[4], st));
  // members arrive in processing order (sources in rank order): a stable sort by group id is push_back order
  launches += launch_sort_pairs(D.gid_a, nullptr, D.sgid, D.srank_g, D.tmp_k, D.tmp_v, mg, bits_g, D.sort_work, st, &D.cnt->err,
                                mg ? D.prehist + 3072 : nullptr);
  OrderArgs oa{};
  oa.sgid = D.sgid, oa.srank = D.srank_g, oa.hfi_r = D.recv_rows;
  oa.m = mg, oa.do_sort = (flags & RK_F_NO_SORT) ? 0 : 1, oa.gid_base = gid_base;
  order_carve(oa, D.order_scratch, D.cap);
  oa.work_count = D.cnt->work_g;
  oa.out_order = D.out_order, oa.out_gid = D.out_gid, oa.out_repval = D.out_repval, oa.out_identity = D.out_identity;
  oa.err = &D.cnt->err;
  launches += launch_order(oa, st);
  CK(cudaEventRecord(ev[5], st));

  u32 *h_order = nullptr, *h_gid = nullptr;
  float *h_ident = nullptr;
  u8 *h_rep = nullptr;
  if ((flags & RK_F_HOST_RESULT) && mg) {
    const u64 need = 3 * align_up((u64)mg * 4, 256) + align_up((u64)mg, 256);
    if (need > D.h_res_cap) {
      CK(cudaStreamSynchronize(st));
      if (D.h_res) cudaFreeHost(D.h_res);
      D.h_res = nullptr, D.h_res_cap = 0;
      CK(cudaHostAlloc(&D.h_res, need + need / 4, cudaHostAllocDefault));
      D.h_res_cap = need + need / 4;
    }
    u8 *hb = (u8 *)D.h_res;
    h_order = (u32 *)hb;
    h_gid = (u32 *)(hb + align_up((u64)mg * 4, 256));
    h_ident = (float *)(hb + 2 * align_up((u64)mg * 4, 256));
    h_rep = hb + 3 * align_up((u64)mg * 4, 256);
    CK(cudaEventRecord(ev[6], st));
    CK(cudaMemcpyAsync(h_order, D.out_order, (u64)mg * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_gid, D.out_gid, (u64)mg * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_ident, D.out_identity, (u64)mg * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_rep, D.out_repval, (u64)mg, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ev[7], st));
  }
  CK(cudaMemcpyAsync(ctx->h_cnt, D.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (ctx->h_cnt->err) return fail(ctx, RK_ERR_INTERNAL, "%s", err_bits_text(ctx->h_cnt->err));
  if (D.trace.on) {
    char cb[200];
    snprintf(cb, sizeof cb, "ranked=%u halo_in=%u halo_out=%u y=%u lines=%u groups=%llu stages x=%.3f y=%.3f forest=%.3f exch=%.3f order=%.3f |", m, nh,
             D.n_away, D.m_y, mg, (unsigned long long)total_groups, ev_ms(ev[0], ev[1]), ev_ms(ev[1], ev[2]), ev_ms(ev[2], ev[3]),
             ev_ms(ev[3], ev[4]), ev_ms(ev[4], ev[5]));
    D.trace.dump(me, cb);
  }

  out->n_kept = mg;
  out->n_groups = total_groups;
  out->order = h_order, out->gid = h_gid, out->repval = h_rep, out->identity = h_ident;
  out->d_order = D.out_order, out->d_gid = D.out_gid, out->d_repval = D.out_repval, out->d_identity = D.out_identity;
  out->n_launches = launches;
  if (flags & RK_F_TIMING) {
    out->ms_stage[RK_ST_XMATCH] = ev_ms(ev[0], ev[1]);
    out->ms_stage[RK_ST_YMATCH] = ev_ms(ev[1], ev[2]);
    out->ms_stage[RK_ST_FOREST] = ev_ms(ev[2], ev[3]);
    out->ms_stage[RK_ST_HKEY] = ev_ms(ev[3], ev[4]);   // output exchange
    out->ms_stage[RK_ST_GSORT] = ev_ms(ev[4], ev[5]);
    if (h_order) out->ms_stage[RK_ST_D2H] = ev_ms(ev[6], ev[7]);
    out->ms_device = ev_ms(ev[0], ev[5]);
  }
  if (info) {
    memset(info, 0, sizeof *info);
    info->rank = me, info->world = nr;
    info->total_loaded = D.n_total_loaded;
    info->total_kept = D.m_total;
    info->total_groups = total_groups;
    info->line_offset = line_off;
    info->n_lines = mg;
    info->rank_offset = D.rank_off;
    info->n_ranked = m;
    info->n_halo_in = nh;
    info->n_halo_out = D.n_away;
    info->n_y = D.m_y;
    info->bytes_sent = D.tr->bytes_sent - sent0;
  }
  return RK_OK;
}

static void peer_rows(const Dist &D, uint4 *local, uint4 **outs) {
  for (int d = 0; d < D.world; ++d) outs[d] = d == D.rank ? local : D.tr->on_peer(d, local);
}

void dist_destroy(rk_ctx *c) {
  if (!c || !c->dist) return;
  Dist *D = c->dist;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  dist_release_buffers(*D);
  if (D->aos_buf) cudaFree(D->aos_buf);
  if (D->link_buf) cudaFree(D->link_buf);
  if (D->h_res) cudaFreeHost(D->h_res);
  if (D->h_small) cudaFreeHost(D->h_small);
  if (D->side) cudaStreamSynchronize(D->side), cudaStreamDestroy(D->side);
  delete D->tr;
  delete D;
  c->dist = nullptr;
}

}  // namespace rk

// ---------------------------------------------------------------------------------------------------------------------
// C ABI: one process per GPU
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int rk_dist_unique_id(void *id128) {
  if (!id128) return RK_ERR_ARG;
  NcclApi &N = nccl_api();
  if (!N.ok()) return RK_ERR_CUDA;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == RK_DIST_ID_BYTES, "unique id size");
  if (N.GetUniqueId(&id) != ncclSuccess) return RK_ERR_CUDA;
  memcpy(id128, &id, sizeof id);
  return RK_OK;
}

int rk_dist_init(rk_ctx *ctx, int rank, int nranks, const void *id128, uint64_t cap_per_rank) {
  if (!ctx || !id128) return RK_ERR_ARG;
  NcclApi &N = nccl_api();
  if (!N.ok()) return fail(ctx, RK_ERR_CUDA, "NCCL is not available: %s", N.err.c_str());
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NcclTransport *tr = new NcclTransport;
  const ncclResult_t r = N.CommInitRank(&tr->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    const int rc = fail(ctx, RK_ERR_CUDA, "ncclCommInitRank: %s", N.GetErrorString(r));
    tr->comm = nullptr;
    delete tr;
    return rc;
  }
  return dist_init_common(ctx, rank, nranks, tr, cap_per_rank);
}

int rk_dist_export(rk_ctx *ctx, void *blob) {
  if (!ctx || !blob) return RK_ERR_ARG;
  if (!ctx->dist) return fail(ctx, RK_ERR_STATE, "rk_dist_export before rk_dist_init");
  memset(blob, 0, RK_DIST_BLOB_BYTES);
  return dist_export(ctx, (DistBlob *)blob);
}

int rk_dist_import(rk_ctx *ctx, const void *blobs) {
  if (!ctx || !blobs) return RK_ERR_ARG;
  if (!ctx->dist) return fail(ctx, RK_ERR_STATE, "rk_dist_import before rk_dist_init");
  return dist_import(ctx, (const u8 *)blobs, RK_DIST_BLOB_BYTES);
}

int rk_dist_load_aos(rk_ctx *ctx, const void *frags, uint64_t n_local, uint64_t file_offset, uint64_t seqx_len, uint64_t seqy_len,
                     unsigned flags, rk_load_stats *stats) {
  if (!ctx) return RK_ERR_ARG;
  if (!ctx->dist) return fail(ctx, RK_ERR_STATE, "rk_dist_load_aos before rk_dist_init");
  return dist_load(ctx, frags, n_local, file_offset, seqx_len, seqy_len, flags, stats);
}

int rk_dist_group(rk_ctx *ctx, double len_ratio, double pos_ratio, unsigned flags, rk_result *out, rk_dist_info *info) {
  if (!ctx || !out) return RK_ERR_ARG;
  if (!ctx->dist) return fail(ctx, RK_ERR_STATE, "rk_dist_group before rk_dist_init");
  return dist_group(ctx, len_ratio, pos_ratio, flags, out, info);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// C ABI: one process, one host thread per GPU
// ---------------------------------------------------------------------------------------------------------------------
struct rk_multi {
  int n = 0;
  std::vector<rk_ctx *> ctx;
  std::vector<int> devices;
  LocalGroup local;
  bool use_nccl = false;
  std::vector<ncclComm_t> comms;
  std::string err;
  u64 cap = 0;
  bool have_load = false;
  // merged host result of the last rk_multi_group
  void *h_res = nullptr;
  u64 h_res_cap = 0;
  std::vector<rk_dist_info> infos;
};

namespace {

template <class F>
int run_ranks(rk_multi *mg, F fn) {  // fn(rank) on one thread per rank; first failing rank's code and message
  std::vector<int> rc(mg->n, RK_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < mg->n; ++r) th.emplace_back([&, r] { rc[r] = fn(r); });
  for (auto &t : th) t.join();
  for (int r = 0; r < mg->n; ++r)
    if (rc[r] != RK_OK) {
      mg->err = "rank " + std::to_string(r) + ": " + rk_last_error(mg->ctx[r]);
      return rc[r];
    }
  return RK_OK;
}

int multi_set_capacity(rk_multi *mg, u64 cap) {
  int rc = run_ranks(mg, [&](int r) {
    rk_ctx *c = mg->ctx[r];
    if (!c->dist) {
      Transport *tr;
      if (mg->use_nccl) {
        NcclTransport *t = new NcclTransport;
        t->comm = mg->comms[r];
        t->own_comm = false;
        tr = t;
      } else {
        LocalTransport *t = new LocalTransport;
        t->g = &mg->local;
        tr = t;
      }
      return dist_init_common(c, r, mg->n, tr, cap);
    }
    return dist_allocate(c, cap);
  });
  if (rc) return rc;
  std::vector<u8> blobs((size_t)mg->n * RK_DIST_BLOB_BYTES, 0);
  rc = run_ranks(mg, [&](int r) { return dist_export(mg->ctx[r], (DistBlob *)(blobs.data() + (size_t)r * RK_DIST_BLOB_BYTES)); });
  if (rc) return rc;
  rc = run_ranks(mg, [&](int r) { return dist_import(mg->ctx[r], blobs.data(), RK_DIST_BLOB_BYTES); });
  if (rc) return rc;
  mg->cap = cap;
  return RK_OK;
}

}  // namespace

extern "C" {

rk_multi *rk_create_multi(const int *devices, int ndev) {
  if (!devices || ndev < 1 || ndev > DIST_MAX_RANKS) return nullptr;
  rk_multi *mg = new rk_multi;
  mg->n = ndev;
  mg->local.n = ndev;
  mg->devices.assign(devices, devices + ndev);
  bool distinct = true;
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j) distinct = distinct && devices[i] != devices[j];
  const char *force = getenv("RK_MULTI_TRANSPORT");  // "local" or "nccl"
  mg->use_nccl = distinct && ndev > 1 && !(force && !strcmp(force, "local"));
  for (int r = 0; r < ndev; ++r) {
    rk_ctx *c = rk_create(devices[r]);
    if (!c) {
      for (rk_ctx *p : mg->ctx) rk_destroy(p);
      delete mg;
      return nullptr;
    }
    mg->ctx.push_back(c);
  }
  if (mg->use_nccl) {
    NcclApi &N = nccl_api();
    mg->comms.resize(ndev);
    if (!N.ok() || N.CommInitAll(mg->comms.data(), ndev, devices) != ncclSuccess) {
      mg->comms.clear();
      mg->use_nccl = false;  // ranks are threads of this process: the copy transport always works
    }
  }
  return mg;
}

void rk_destroy_multi(rk_multi *mg) {
  if (!mg) return;
  for (rk_ctx *c : mg->ctx) rk_destroy(c);
  for (ncclComm_t c : mg->comms) nccl_api().CommDestroy(c);
  if (mg->h_res) cudaFreeHost(mg->h_res);
  delete mg;
}

const char *rk_multi_last_error(const rk_multi *mg) { return mg ? mg->err.c_str() : "null rk_multi"; }
int rk_multi_ranks(const rk_multi *mg) { return mg ? mg->n : 0; }
rk_ctx *rk_multi_ctx(rk_multi *mg, int rank) { return (mg && rank >= 0 && rank < mg->n) ? mg->ctx[rank] : nullptr; }
const char *rk_multi_transport(const rk_multi *mg) { return !mg ? "" : (mg->use_nccl ? "nccl" : "local"); }

int rk_multi_load_aos(rk_multi *mg, const void *frags, uint64_t n, uint64_t seqx_len, uint64_t seqy_len, unsigned flags,
                      rk_load_stats *stats) {
  if (!mg || (!frags && n)) return RK_ERR_ARG;
  mg->have_load = false;
  // every rank starts with a contiguous slice of the file (slices start on a 16-record boundary: 16-byte aligned records)
  std::vector<u64> lo(mg->n + 1);
  for (int r = 0; r <= mg->n; ++r) {
    u64 b = n * (u64)r / (u64)mg->n;
    if (r < mg->n) b -= b % 16;
    lo[r] = r == mg->n ? n : b;
  }
  u64 biggest = 0;
  for (int r = 0; r < mg->n; ++r) biggest = std::max(biggest, lo[r + 1] - lo[r]);
  const u64 want = biggest + biggest / 2 + (1u << 16);
  if (want > mg->cap) {
    const int rc = multi_set_capacity(mg, want);
    if (rc) return rc;
  }
  std::vector<rk_load_stats> st(mg->n);
  const int rc = run_ranks(mg, [&](int r) {
    return dist_load(mg->ctx[r], (const u8 *)frags + lo[r] * RK_FRAG_BYTES, lo[r + 1] - lo[r], lo[r], seqx_len, seqy_len, flags, &st[r]);
  });
  if (rc) return rc;
  if (stats) {
    *stats = st[0];
    for (int r = 1; r < mg->n; ++r) stats->n_launches += st[r].n_launches;
  }
  mg->have_load = true;
  return RK_OK;
}

int rk_multi_group(rk_multi *mg, double len_ratio, double pos_ratio, unsigned flags, rk_result *out) {
  if (!mg || !out) return RK_ERR_ARG;
  if (!mg->have_load) {
    mg->err = "rk_multi_group before a successful rk_multi_load_aos";
    return RK_ERR_STATE;
  }
  std::vector<rk_result> res(mg->n);
  mg->infos.assign(mg->n, rk_dist_info{});
  const int rc = run_ranks(mg, [&](int r) { return dist_group(mg->ctx[r], len_ratio, pos_ratio, flags | RK_F_HOST_RESULT, &res[r], &mg->infos[r]); });
  if (rc) return rc;
  // the ranks hold consecutive ranges of output lines: concatenate them
  u64 total = 0;
  for (int r = 0; r < mg->n; ++r) total += res[r].n_kept;
  const u64 need = 3 * align_up(total * 4, 256) + align_up(total, 256) + 256;
  if (need > mg->h_res_cap) {
    if (mg->h_res) cudaFreeHost(mg->h_res);
    mg->h_res = nullptr, mg->h_res_cap = 0;
    if (cudaHostAlloc(&mg->h_res, need, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      mg->err = "cudaHostAlloc for the merged result failed";
      return RK_ERR_NOMEM;
    }
    mg->h_res_cap = need;
  }
  u8 *hb = (u8 *)mg->h_res;
  u32 *h_order = (u32 *)hb, *h_gid = (u32 *)(hb + align_up(total * 4, 256));
  float *h_ident = (float *)(hb + 2 * align_up(total * 4, 256));
  u8 *h_rep = hb + 3 * align_up(total * 4, 256);
  std::vector<std::thread> th;
  for (int r = 0; r < mg->n; ++r)
    th.emplace_back([&, r] {
      const u64 off = mg->infos[r].line_offset, k = res[r].n_kept;
      if (!k) return;
      memcpy(h_order + off, res[r].order, k * 4);
      memcpy(h_gid + off, res[r].gid, k * 4);
      memcpy(h_ident + off, res[r].identity, k * 4);
      memcpy(h_rep + off, res[r].repval, k);
    });
  for (auto &t : th) t.join();
  memset(out, 0, sizeof *out);
  out->n_kept = total;
  out->n_groups = res[0].n_groups;
  out->order = h_order, out->gid = h_gid, out->repval = h_rep, out->identity = h_ident;
  for (int r = 0; r < mg->n; ++r) {
    out->n_launches += res[r].n_launches;
    out->ms_device = std::max(out->ms_device, res[r].ms_device);
    for (int s = 0; s < RK_NSTAGES; ++s) out->ms_stage[s] = std::max(out->ms_stage[s], res[r].ms_stage[s]);
  }
  return RK_OK;
}

int rk_multi_info(const rk_multi *mg, int rank, rk_dist_info *info) {
  if (!mg || !info || rank < 0 || rank >= (int)mg->infos.size()) return RK_ERR_ARG;
  *info = mg->infos[rank];
  return RK_OK;
}

}  // extern "C"
