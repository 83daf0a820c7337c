// C ABI of librk_b200.so (include/rk_b200.h): context, device workspace, and the launch sequences that stand
// in for FragmentsDatabase's constructor (rk_load_aos) and for generate_fragment_groups +
// generate_diagonal_func + sort_groups (rk_group).  Reference call sites: /root/reference/src/repkiller.cpp:52,84-91.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "rk_ctx.cuh"

using namespace rk;

namespace {

std::mutex g_create_mu;        // rk_create may run on several threads (one context per GPU)
std::string g_create_error;   // guarded by g_create_mu; rk_create_error() returns a per-thread copy
thread_local std::string tl_create_error;
void set_create_error(const std::string &s) {
  std::lock_guard<std::mutex> lk(g_create_mu);
  g_create_error = s;
}

const char *const kKernelNames[KID_COUNT] = {"k_decode", "k_radix_hist", "k_scan", "k_radix_scatter", "k_keys", "k_match_small",
                                             "k_match_long", "k_chase", "k_hkey", "k_pack", "k_order_tile",
                                             "k_groupsort_large", "k_finalize", "k_diag_table", "k_groupsort_warp", "k_format", "k_dist_rows",
                                             "k_group_stats", "k_chase_exits"};

thread_local Profiler *tl_prof = nullptr;

}  // namespace

namespace {

// carve the workspace for n records; returns bytes needed.  With base == nullptr only sizes are computed.
u64 carve(rk_ctx *c, u8 *base, u64 n, u64 stage_bytes, u64 lxw, u64 lyw) {
  u64 off = 0;
  auto take = [&](u64 bytes) -> u8 * {
    u8 *p = base ? base + off : nullptr;
    off += align_up(bytes ? bytes : 16, 256);
    return p;
  };
  const u64 n1 = n ? n : 1;
  c->d_cnt = (Counters *)take(sizeof(Counters));
  c->d_aos = stage_bytes ? take(stage_bytes) : nullptr;  // host records staged on the device (109 B or 33 B each)
  c->rec4 = (uint4 *)take(n1 * 32);
  c->key0 = (u32 *)take(n1 * 4);
  c->identity_r = (float *)take(n1 * 4);
  c->hfi_r = (uint4 *)take(n1 * 16);
  c->link_x = (u32 *)take(lxw * 4);
  c->link_y = (u32 *)take(lyw * 4);
  c->k0_r = (u32 *)take(n1 * 4);
  c->fidx_r = (u32 *)take(n1 * 4);
  c->tmp_k = (u32 *)take(n1 * 4);
  c->tmp_v = (u32 *)take(n1 * 4);
  c->xl_r = (uint2 *)take(n1 * 8);
  c->yl_r = (uint2 *)take(n1 * 8);
  c->ys_r = (u32 *)take(n1 * 4);
  c->kx = (u32 *)take(n1 * 4);
  c->ky = (u32 *)take(n1 * 4);
  c->skx = (u32 *)take(n1 * 4);
  c->rx = (u32 *)take(n1 * 4);
  c->sky = (u32 *)take(n1 * 4);
  c->ry = (u32 *)take(n1 * 4);
  c->sort_work = take(sort_work_bytes(n1));
  c->prehist = (u32 *)take(4 * 4 * 256 * 4);
  c->parent = (u32 *)take(n1 * 4);
  c->xm_bits = (u32 *)take((n1 + 31) / 32 * 4);
  c->gid_rank = (u32 *)take(n1 * 4);
  c->h = (u32 *)take(n1 * 4);
  c->sgid = (u32 *)take(n1 * 4);
  c->srank = (u32 *)take(n1 * 4);
  c->forest_work = take(forest_work_bytes((u32)n1));
  c->order_scratch = take(order_scratch_bytes(n1));
  c->work_cap = (u32)(n1 / 32 + 2);   // segments of more than 32 fragments (K3)
  c->worklist = (u32 *)take((u64)c->work_cap * 4);
  c->ent_rank = (u32 *)take(n1 * 4);
  c->ent_c = (u32 *)take(n1 * 4);
  c->ent_len = (u32 *)take(n1 * 4);
  c->out_order = (u32 *)take(n1 * 4);
  c->out_gid = (u32 *)take(n1 * 4);
  c->out_repval = take(n1);
  c->out_identity = (float *)take(n1 * 4);
  return off;
}

}  // namespace

namespace rk {
Geometry make_geometry(u64 seqx_len, u64 seqy_len) {
  Geometry g{};
  g.lx = seqx_len;
  g.ly = seqy_len;
  g.vsize = (u32)(1 + seqx_len / XBUCKET);  // FragmentsDatabase.cpp:84
  g.mx = (u32)(seqx_len / DIVISOR);         // SequenceOcupationList.cpp:4
  g.my = (u32)(seqy_len / DIVISOR);
  g.nbx = g.mx + 2;
  g.nby = g.my + 2;
  return g;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

const char *err_bits_text(u32 e) {
  if (e & ERR_XBUCKET) return "a fragment has xStart/10 >= vsize (the reference writes out of bounds there)";
  if (e & ERR_CENTER) return "a fragment center lies beyond its sequence's occupation list (xStart+length/2 > sequence length)";
  if (e & ERR_COORD) return "a coordinate, length or center does not fit in 32 bits";
  if (e & ERR_WORKLIST) return "internal: segment worklist overflow";
  if (e & ERR_SPIN) return "internal: bounded wait expired";
  return "unknown device error";
}

void prof_route(Profiler *p) { tl_prof = p; }
void prof_begin(int kid, cudaStream_t st, unsigned long long units) {
  if (!tl_prof) return;
  Profiler::Rec r{kid, tl_prof->get(), tl_prof->get(), (u64)units};
  cudaEventRecord(r.a, st);
  tl_prof->open_recs.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!tl_prof) return;
  cudaEventRecord(tl_prof->open_recs.back().b, st);
}
}  // namespace rk


namespace {
// K5b tail + K5c on the state the last rk_group left on the device
u64 run_order(rk_ctx *ctx, unsigned flags) {
  OrderArgs oa{};
  oa.sgid = ctx->sgid, oa.srank = ctx->srank, oa.hfi_r = ctx->hfi_r;
  oa.m = ctx->m, oa.do_sort = (flags & RK_F_NO_SORT) ? 0 : 1;
  order_carve(oa, ctx->order_scratch, ctx->n);
  oa.work_count = ctx->d_cnt->work_g;
  oa.out_order = ctx->out_order, oa.out_gid = ctx->out_gid, oa.out_repval = ctx->out_repval, oa.out_identity = ctx->out_identity;
  oa.err = &ctx->d_cnt->err;
  return (u64)launch_order(oa, ctx->stream);
}

// optional D2H of the four output arrays, final synchronisation, error word, result struct (events 0..5 recorded)
int finish_group(rk_ctx *ctx, unsigned flags, rk_result *out, u64 launches) {
  const u32 m = ctx->m;
  cudaStream_t st = ctx->stream;
  cudaEvent_t *ev = ctx->ev;
  if ((flags & RK_F_HOST_RESULT) && m) {
    const u64 need = 3 * align_up((u64)m * 4, 256) + align_up((u64)m, 256);
    if (need > ctx->h_res_cap) {
      CK(cudaStreamSynchronize(st));
      if (ctx->h_res) cudaFreeHost(ctx->h_res);
      ctx->h_res = nullptr;
      ctx->h_res_cap = 0;
      CK(cudaHostAlloc(&ctx->h_res, need, cudaHostAllocDefault));
      ctx->h_res_cap = need;
    }
  }
  CK(cudaEventRecord(ev[6], st));
  u8 *hb = (u8 *)ctx->h_res;
  u32 *h_order = nullptr, *h_gid = nullptr;
  float *h_ident = nullptr;
  u8 *h_rep = nullptr;
  if ((flags & RK_F_HOST_RESULT) && m) {
    h_order = (u32 *)hb;
    h_gid = (u32 *)(hb + align_up((u64)m * 4, 256));
    h_ident = (float *)(hb + 2 * align_up((u64)m * 4, 256));
    h_rep = hb + 3 * align_up((u64)m * 4, 256);
    CK(cudaMemcpyAsync(h_order, ctx->out_order, (u64)m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_gid, ctx->out_gid, (u64)m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_ident, ctx->out_identity, (u64)m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_rep, ctx->out_repval, (u64)m, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(ev[7], st));
  CK(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  if (ctx->copy_pending) {  // the formatter-only part of a compact load: done by the time a grouping is handed back
    CK(cudaStreamWaitEvent(st, ctx->copy_done, 0));
    ctx->copy_pending = false;
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (ctx->h_cnt->err) return fail(ctx, RK_ERR_INTERNAL, "%s", err_bits_text(ctx->h_cnt->err));
  ctx->have_group = true;
  ctx->n_groups_last = ctx->h_cnt->n_groups;

  out->n_kept = m;
  out->n_groups = ctx->h_cnt->n_groups;
  out->order = h_order;
  out->gid = h_gid;
  out->repval = h_rep;
  out->identity = h_ident;
  out->d_order = ctx->out_order;
  out->d_gid = ctx->out_gid;
  out->d_repval = ctx->out_repval;
  out->d_identity = ctx->out_identity;
  out->n_launches = launches;
  if (flags & RK_F_TIMING) {
    out->ms_stage[RK_ST_XMATCH] = ev_ms(ev[0], ev[1]);
    out->ms_stage[RK_ST_YMATCH] = ev_ms(ev[1], ev[2]);
    out->ms_stage[RK_ST_FOREST] = ev_ms(ev[2], ev[3]);
    out->ms_stage[RK_ST_HKEY] = ev_ms(ev[3], ev[4]);
    out->ms_stage[RK_ST_GSORT] = ev_ms(ev[4], ev[5]);
    out->ms_stage[RK_ST_D2H] = ev_ms(ev[6], ev[7]);
    out->ms_device = ev_ms(ev[0], ev[5]);
  }
  return RK_OK;
}
}  // namespace

extern "C" {

const char *rk_version(void) { return "repkiller-b200 0.1 (sm_100a)"; }
const char *rk_create_error(void) {
  std::lock_guard<std::mutex> lk(g_create_mu);
  tl_create_error = g_create_error;
  return tl_create_error.c_str();
}

rk_ctx *rk_create(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_create_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= count) {
    set_create_error("device index out of range");
    return nullptr;
  }
  rk_ctx *c = new rk_ctx;
  c->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaHostAlloc((void **)&c->h_cnt, sizeof(Counters), cudaHostAllocDefault)) != cudaSuccess) {
    set_create_error(cudaGetErrorString(e));
    delete c;
    return nullptr;
  }
  c->own_stream = true;
  if (const char *gr = getenv("RK_L2_GRAN")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(gr));  // tuning switch
  for (auto &ev : c->ev) cudaEventCreate(&ev);
  cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming);
  if (cudaMalloc((void **)&c->st_cnt, sizeof(Counters)) != cudaSuccess) c->st_cnt = nullptr;
  // function attributes are per device: every context sets them for its own device (a process-wide flag would leave
  // the second device of a process without the shared-memory opt-in)
  if ((e = decode_init_device()) != cudaSuccess || (e = sort_init_device()) != cudaSuccess || (e = order_init_device()) != cudaSuccess ||
      (e = dist_init_device()) != cudaSuccess) {
    set_create_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    rk_destroy(c);
    return nullptr;
  }
  return c;
}

void rk_destroy(rk_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  dist_destroy(c);
  if (c->arena) cudaFree(c->arena);
  if (c->st_cnt) cudaFree(c->st_cnt);
  if (c->st_scratch) cudaFree(c->st_scratch);
  if (c->d_text) cudaFree(c->d_text);
  if (c->d_stats) cudaFree(c->d_stats);
  if (c->h_stats) cudaFreeHost(c->h_stats);
  for (char *t : c->h_text) if (t) cudaFreeHost(t);
  if (c->h_res) cudaFreeHost(c->h_res);
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  for (auto &ev : c->ev) cudaEventDestroy(ev);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream), cudaStreamDestroy(c->copy_stream);
  if (c->copy_done) cudaEventDestroy(c->copy_done);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *rk_last_error(const rk_ctx *c) { return c ? c->err.c_str() : rk_create_error(); }

int rk_set_stream(rk_ctx *ctx, void *cuda_stream) {
  if (!ctx) return RK_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return RK_OK;
}

}  // extern "C"

namespace {

bool is_device_pointer(const void *p) {
  cudaPointerAttributes pa;
  if (cudaPointerGetAttributes(&pa, p) == cudaSuccess) return pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  return false;
}

// The two ingest forms: 109-byte records (rk_load_aos) or the compact arrays of rk_load_packed.
struct LoadSource {
  const void *aos = nullptr;
  const void *key4 = nullptr, *strand = nullptr, *rest4 = nullptr;
  bool packed() const { return aos == nullptr; }
};

// FragmentsDatabase's constructor after parsing (src/FragmentsDatabase.cpp:84-100) on the device: K1, the processing
// order (K2a), the occupation-list buckets of both axes (K2 keys, K2b/c) and generate_diagonal_func per fragment (K5a).
int load_common(rk_ctx *ctx, const LoadSource &src, uint64_t n, uint64_t seqx_len, uint64_t seqy_len, unsigned flags, rk_load_stats *stats) {
  if (n >= 0xFFFFFFF0ull) return fail(ctx, RK_ERR_ARG, "more than 2^32-16 records per context; partition the input");
  if (seqx_len >= (1ull << 32) || seqy_len >= (1ull << 32))
    return fail(ctx, RK_ERR_RANGE, "sequence length does not fit in 32 bits");
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  ctx->loaded = false;
  ctx->have_group = false;

  const void *first = src.packed() ? src.key4 : src.aos;
  const bool on_device = n ? is_device_pointer(first) : false;
  if (on_device && (((uintptr_t)first & 15) || (src.rest4 && ((uintptr_t)src.rest4 & 15))))
    return fail(ctx, RK_ERR_ARG, "device record pointers must be 16-byte aligned");
  if (src.packed() && n && (on_device != is_device_pointer(src.strand) || (src.rest4 && on_device != is_device_pointer(src.rest4))))
    return fail(ctx, RK_ERR_ARG, "the arrays of rk_load_packed must all be host or all be device memory");

  const Geometry g = make_geometry(seqx_len, seqy_len);
  const u64 lxw = (2ull * g.nbx + 31) / 32 + 1, lyw = (2ull * g.nby + 31) / 32 + 1;

  // (re)carve the workspace; the staging area holds the host records on the device: 109 B or 33 B per record
  const u64 stage_bytes = on_device ? 0 : (src.packed() ? align_up(n * 16, 256) + align_up(n, 256) + (src.rest4 ? align_up(n * 16, 256) : 0)
                                                        : align_up(n * RK_FRAG_BYTES, 16) + 16);
  const u64 need = carve(ctx, nullptr, n, stage_bytes, lxw, lyw);
  if (need > ctx->arena_bytes) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->arena, need);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, RK_ERR_NOMEM, "cudaMalloc(%llu bytes): %s", (unsigned long long)need, cudaGetErrorString(e));
    }
    ctx->arena_bytes = need;
  }
  carve(ctx, (u8 *)ctx->arena, n, stage_bytes, lxw, lyw);
  ctx->link_x_words = lxw;
  ctx->link_y_words = lyw;
  ctx->n = n;
  ctx->g = g;
  ctx->bits_rank = ceil_log2(g.vsize) < 1 ? 1 : ceil_log2(g.vsize);  // (the sort clamps to >= 1 bit as well)
  ctx->bits_x = ceil_log2(2ull * g.nbx);                             // >= 2: nbx >= 2
  ctx->bits_y = ceil_log2(2ull * g.nby);

  cudaStream_t st = ctx->stream;
  cudaEvent_t *ev = ctx->ev;
  CK(cudaEventRecord(ev[0], st));
  const u8 *aos = (const u8 *)src.aos;
  const uint4 *key4 = (const uint4 *)src.key4, *rest4 = (const uint4 *)src.rest4;
  const u8 *strand = (const u8 *)src.strand;
  if (!on_device && n) {
    if (src.packed()) {
      u8 *p = ctx->d_aos;
      CK(cudaMemcpyAsync(p, src.key4, n * 16, cudaMemcpyHostToDevice, st));
      key4 = (const uint4 *)p, p += align_up(n * 16, 256);
      CK(cudaMemcpyAsync(p, src.strand, n, cudaMemcpyHostToDevice, st));
      strand = p, p += align_up(n, 256);
      if (src.rest4) {  // only rk_format_lines reads these: the copy runs on its own stream while the kernels work
        CK(cudaStreamSynchronize(ctx->copy_stream));           // (an earlier load's copy into the old arena)
        CK(cudaEventRecord(ctx->copy_done, st));               // not before the work queued on st is done with the arena
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done, 0));
        CK(cudaMemcpyAsync(p, src.rest4, n * 16, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
        ctx->copy_pending = true;
        rest4 = (const uint4 *)p;
      }
    } else {
      CK(cudaMemcpyAsync(ctx->d_aos, src.aos, n * RK_FRAG_BYTES, cudaMemcpyHostToDevice, st));
      aos = ctx->d_aos;
    }
  }
  ctx->aos_dev = aos;
  ctx->pk_key = key4, ctx->pk_rest = rest4, ctx->pk_strand = strand;
  CK(cudaEventRecord(ev[1], st));
  CK(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(Counters), st));
  CK(cudaMemsetAsync(ctx->link_x, 0, lxw * 4, st));
  CK(cudaMemsetAsync(ctx->link_y, 0, lyw * 4, st));
  u64 launches = 0;
  CK(cudaMemsetAsync(ctx->prehist, 0, 3 * 4 * 256 * 4, st));
  auto hist_of = [&](int which, int bits) { return HistOut{ctx->prehist + which * 1024, (bits + 7) / 8, bits}; };
  if (src.packed())
    launches += launch_decode_packed(key4, strand, n, g, ctx->key0, ctx->link_x, ctx->link_y, &ctx->d_cnt->n_dropped, &ctx->d_cnt->err, st,
                                     ctx->rec4, hist_of(0, ctx->bits_rank));
  else
    launches += launch_decode(aos, n, g, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->key0, ctx->link_x, ctx->link_y,
                              &ctx->d_cnt->n_dropped, &ctx->d_cnt->err, st, ctx->rec4, hist_of(0, ctx->bits_rank));
  CK(cudaEventRecord(ev[2], st));
  // the rank sort does not depend on the number of dropped records: they carry the largest key and sort last
  launches += launch_sort_pairs(ctx->key0, nullptr, ctx->k0_r, ctx->fidx_r, ctx->tmp_k, ctx->tmp_v, n, ctx->bits_rank, ctx->sort_work, st, &ctx->d_cnt->err,
                                ctx->prehist);
  CK(cudaEventRecord(ev[3], st));
  CK(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (ctx->h_cnt->err) return fail(ctx, (ctx->h_cnt->err & (ERR_WORKLIST | ERR_SPIN)) ? RK_ERR_INTERNAL : RK_ERR_RANGE, "%s",
                                   err_bits_text(ctx->h_cnt->err));
  const u32 m = (u32)(n - ctx->h_cnt->n_dropped);
  ctx->m = m;

  CK(cudaEventRecord(ev[4], st));
  launches += launch_keys(ctx->fidx_r, m, g, ctx->rec4, ctx->link_x, ctx->link_y, ctx->xl_r, ctx->yl_r, ctx->ys_r, ctx->kx, ctx->ky,
                          ctx->identity_r, st, hist_of(1, ctx->bits_x), hist_of(2, ctx->bits_y));
  // generate_diagonal_func does not depend on the ratios: h and the per-rank output record are load-time work
  launches += launch_hkey(ctx->k0_r, ctx->ys_r, m, ctx->h, st, ctx->fidx_r, ctx->identity_r, ctx->hfi_r);
  CK(cudaEventRecord(ev[5], st));
  launches += launch_sort_pairs(ctx->kx, nullptr, ctx->skx, ctx->rx, ctx->tmp_k, ctx->tmp_v, m, ctx->bits_x, ctx->sort_work, st, &ctx->d_cnt->err,
                                ctx->prehist + 1024);
  CK(cudaEventRecord(ev[6], st));
  launches += launch_sort_pairs(ctx->ky, nullptr, ctx->sky, ctx->ry, ctx->tmp_k, ctx->tmp_v, m, ctx->bits_y, ctx->sort_work, st, &ctx->d_cnt->err,
                                ctx->prehist + 2048);
  CK(cudaEventRecord(ev[7], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  ctx->loaded = true;

  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->n_loaded = n;
    stats->n_kept = m;
    stats->vsize = g.vsize;
    stats->n_launches = launches;
    if (flags & RK_F_TIMING) {
      stats->ms_stage[RK_ST_H2D] = ev_ms(ev[0], ev[1]);
      stats->ms_stage[RK_ST_DECODE] = ev_ms(ev[1], ev[2]);
      stats->ms_stage[RK_ST_RANKSORT] = ev_ms(ev[2], ev[3]);
      stats->ms_stage[RK_ST_KEYS] = ev_ms(ev[4], ev[5]);
      stats->ms_stage[RK_ST_XSORT] = ev_ms(ev[5], ev[6]);
      stats->ms_stage[RK_ST_YSORT] = ev_ms(ev[6], ev[7]);
      stats->ms_device = ev_ms(ev[1], ev[3]) + ev_ms(ev[4], ev[7]);
    }
  }
  return RK_OK;
}

}  // namespace

extern "C" {

int rk_load_aos(rk_ctx *ctx, const void *frags, uint64_t n, uint64_t seqx_len, uint64_t seqy_len, unsigned flags,
                rk_load_stats *stats) {
  if (!ctx) return RK_ERR_ARG;
  if (!frags && n) return fail(ctx, RK_ERR_ARG, "null record pointer");
  LoadSource src;
  src.aos = frags ? frags : (const void *)"";  // n == 0: any non-null marker of the AoS form
  return load_common(ctx, src, n, seqx_len, seqy_len, flags, stats);
}

int rk_load_packed(rk_ctx *ctx, const uint32_t *key4, const uint8_t *strand, const uint32_t *rest4, uint64_t n, uint64_t seqx_len,
                   uint64_t seqy_len, unsigned flags, rk_load_stats *stats) {
  if (!ctx) return RK_ERR_ARG;
  if (n && (!key4 || !strand)) return fail(ctx, RK_ERR_ARG, "null array");
  LoadSource src;
  src.key4 = key4, src.strand = strand, src.rest4 = rest4;
  return load_common(ctx, src, n, seqx_len, seqy_len, flags, stats);
}

int rk_group(rk_ctx *ctx, double len_ratio, double pos_ratio, unsigned flags, rk_result *out) {
  if (!ctx || !out) return RK_ERR_ARG;
  if (!ctx->loaded) return fail(ctx, RK_ERR_STATE, "rk_group before a successful rk_load_aos");
  if (!(len_ratio > 0)) return fail(ctx, RK_ERR_ARG, "Ratio between length and position must be greater than zero");
  if (!(pos_ratio > 0)) return fail(ctx, RK_ERR_ARG, "Position proximity must be greater than zero");
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  memset(out, 0, sizeof *out);
  const u32 m = ctx->m;
  cudaStream_t st = ctx->stream;
  cudaEvent_t *ev = ctx->ev;
  u64 launches = 0;

  CK(cudaEventRecord(ev[0], st));
  MatchArgs mx{};
  mx.skey = ctx->skx, mx.srank = ctx->rx, mx.cl_r = ctx->xl_r, mx.parent = ctx->parent, mx.xm_bits = ctx->xm_bits;
  mx.m = m, mx.max_index = ctx->g.mx, mx.len_ratio = len_ratio, mx.pos_ratio = pos_ratio, mx.is_y = 0;
  mx.worklist = ctx->worklist, mx.work_count = ctx->d_cnt->work_x, mx.work_cap = ctx->work_cap;
  mx.ent_rank = ctx->ent_rank, mx.ent_c = ctx->ent_c, mx.ent_len = ctx->ent_len, mx.err = &ctx->d_cnt->err;
  launches += launch_match(mx, st);
  CK(cudaEventRecord(ev[1], st));
  MatchArgs my = mx;
  my.skey = ctx->sky, my.srank = ctx->ry, my.cl_r = ctx->yl_r, my.max_index = ctx->g.my, my.is_y = 1;
  my.work_count = ctx->d_cnt->work_y;
  launches += launch_match(my, st);
  CK(cudaEventRecord(ev[2], st));
  const int bits_g = ceil_log2(m) < 1 ? 1 : ceil_log2(m);
  CK(cudaMemsetAsync(ctx->prehist + 3072, 0, 4 * 256 * 4, st));
  launches += launch_forest(ctx->parent, m, ctx->gid_rank, &ctx->d_cnt->n_groups, ctx->forest_work, st, 0, 0xFFFFFFFFu,
                            HistOut{ctx->prehist + 3072, (bits_g + 7) / 8, bits_g});
  CK(cudaEventRecord(ev[3], st));
  CK(cudaEventRecord(ev[4], st));
  // gids are < number of groups <= m; sorting by ceil_log2(m) bits avoids a host round trip for the count
  launches += launch_sort_pairs(ctx->gid_rank, nullptr, ctx->sgid, ctx->srank, ctx->tmp_k, ctx->tmp_v, m, bits_g,
                                ctx->sort_work, st, &ctx->d_cnt->err, m ? ctx->prehist + 3072 : nullptr);
  launches += run_order(ctx, flags);
  CK(cudaEventRecord(ev[5], st));
  return finish_group(ctx, flags, out, launches);
}

int rk_sort_groups(rk_ctx *ctx, unsigned flags, rk_result *out) {
  if (!ctx || !out) return RK_ERR_ARG;
  if (!ctx->loaded || !ctx->have_group) return fail(ctx, RK_ERR_STATE, "rk_sort_groups before rk_group");
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  memset(out, 0, sizeof *out);
  cudaEvent_t *ev = ctx->ev;
  for (int i = 0; i < 5; ++i) CK(cudaEventRecord(ev[i], ctx->stream));
  const u64 launches = run_order(ctx, flags & ~RK_F_NO_SORT);
  CK(cudaEventRecord(ev[5], ctx->stream));
  return finish_group(ctx, flags, out, launches);
}

int rk_group_statistics(rk_ctx *ctx, unsigned flags, const rk_group_stats **stats, const rk_group_stats **d_stats, uint64_t *n_groups) {
  static_assert(sizeof(rk_group_stats) == sizeof(rk_group_stats_dev) && sizeof(rk_group_stats) == 40, "stats layout");
  if (!ctx) return RK_ERR_ARG;
  if (!ctx->loaded || !ctx->have_group) return fail(ctx, RK_ERR_STATE, "rk_group_statistics before rk_group");
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  cudaStream_t st = ctx->stream;
  const u64 ng = ctx->n_groups_last;
  if (ng > ctx->stats_cap) {
    CK(cudaStreamSynchronize(st));
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    ctx->d_stats = nullptr, ctx->stats_cap = 0;
    const u64 cap = ng + ng / 8 + 1024;
    cudaError_t e = cudaMalloc(&ctx->d_stats, cap * sizeof(rk_group_stats));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, RK_ERR_NOMEM, "cudaMalloc(%llu bytes): %s", (unsigned long long)(cap * sizeof(rk_group_stats)), cudaGetErrorString(e));
    }
    ctx->stats_cap = cap;
  }
  launch_group_stats(ctx->out_order, ctx->out_gid, ctx->out_identity, ctx->rec4, ctx->m, (u32)ng, (rk_group_stats_dev *)ctx->d_stats, st);
  if ((flags & RK_F_HOST_RESULT) && ng) {
    if (ng > ctx->h_stats_cap) {
      CK(cudaStreamSynchronize(st));
      if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
      ctx->h_stats = nullptr, ctx->h_stats_cap = 0;
      const u64 cap = ng + ng / 8 + 1024;
      CK(cudaHostAlloc(&ctx->h_stats, cap * sizeof(rk_group_stats), cudaHostAllocDefault));
      ctx->h_stats_cap = cap;
    }
    CK(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, ng * sizeof(rk_group_stats), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (stats) *stats = ((flags & RK_F_HOST_RESULT) && ng) ? (const rk_group_stats *)ctx->h_stats : nullptr;
  if (d_stats) *d_stats = (const rk_group_stats *)ctx->d_stats;
  if (n_groups) *n_groups = ng;
  return RK_OK;
}

int rk_diagonal_func(rk_ctx *ctx, uint64_t *diag_func) {
  if (!ctx || !diag_func) return RK_ERR_ARG;
  if (!ctx->loaded) return fail(ctx, RK_ERR_STATE, "rk_diagonal_func before rk_load_aos");
  CK(cudaSetDevice(ctx->device));
  const u32 nb = ctx->g.vsize - 1;
  if (nb == 0) return RK_OK;
  u64 *d = nullptr;
  CK(cudaMalloc(&d, (u64)nb * 8));
  launch_diag_table(ctx->k0_r, ctx->ys_r, ctx->m, ctx->g.vsize, d, nullptr, ctx->stream);
  cudaError_t e = cudaMemcpyAsync(diag_func, d, (u64)nb * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, RK_ERR_CUDA, "rk_diagonal_func: %s", cudaGetErrorString(e));
  return RK_OK;
}

int rk_format_lines(rk_ctx *ctx, uint64_t first_line, uint64_t n_lines, rk_text *out) {
  if (!ctx || !out) return RK_ERR_ARG;
  if (!ctx->loaded || !ctx->have_group) return fail(ctx, RK_ERR_STATE, "rk_format_lines before rk_group");
  if (!ctx->aos_dev && !ctx->pk_rest)
    return fail(ctx, RK_ERR_STATE, "the database was loaded with rk_load_packed without the xEnd/yEnd/score/similarity array");
  if (first_line > ctx->m || n_lines > ctx->m - first_line) return fail(ctx, RK_ERR_ARG, "line range beyond the %u output lines", ctx->m);
  if (n_lines > RK_FORMAT_MAX_LINES) return fail(ctx, RK_ERR_ARG, "at most %llu lines per call", (unsigned long long)RK_FORMAT_MAX_LINES);
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  memset(out, 0, sizeof *out);
  cudaStream_t st = ctx->stream;
  const u64 text_cap = align_up(n_lines * RK_FORMAT_MAX_LINE + 16, 256);
  const u64 need = text_cap + format_work_bytes((u32)n_lines);
  if (need > ctx->d_text_bytes) {
    CK(cudaStreamSynchronize(st));
    if (ctx->d_text) cudaFree(ctx->d_text);
    ctx->d_text = nullptr, ctx->d_text_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->d_text, need);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, RK_ERR_NOMEM, "cudaMalloc(%llu bytes): %s", (unsigned long long)need, cudaGetErrorString(e));
    }
    ctx->d_text_bytes = need;
  }
  FormatArgs fa{};
  fa.aos = ctx->aos_dev, fa.pk_key = ctx->pk_key, fa.pk_rest = ctx->pk_rest, fa.pk_strand = ctx->pk_strand;
  fa.order = ctx->out_order, fa.gid = ctx->out_gid, fa.repval = ctx->out_repval, fa.identity = ctx->out_identity;
  fa.first_line = (u32)first_line, fa.n_lines = (u32)n_lines;
  fa.text = (char *)ctx->d_text;
  u32 *work = (u32 *)((u8 *)ctx->d_text + text_cap);
  fa.total_bytes = work;  // first word of the work area; launch_format carves the rest after it
  CK(cudaEventRecord(ctx->ev[0], st));
  launch_format(fa, work + 8, st);
  CK(cudaMemcpyAsync(&ctx->h_cnt->pad, fa.total_bytes, sizeof(u32), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  const u64 nbytes = ctx->h_cnt->pad;
  const int hb = ctx->h_text_next;
  ctx->h_text_next ^= 1;
  if (nbytes + 1 > ctx->h_text_bytes[hb]) {
    if (ctx->h_text[hb]) cudaFreeHost(ctx->h_text[hb]);
    ctx->h_text[hb] = nullptr, ctx->h_text_bytes[hb] = 0;
    const u64 cap = align_up(nbytes + nbytes / 8 + 4096, 4096);
    CK(cudaHostAlloc((void **)&ctx->h_text[hb], cap, cudaHostAllocDefault));
    ctx->h_text_bytes[hb] = cap;
  }
  if (nbytes) CK(cudaMemcpyAsync(ctx->h_text[hb], ctx->d_text, nbytes, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[1], st));
  CK(cudaStreamSynchronize(st));
  out->text = ctx->h_text[hb];
  out->n_bytes = nbytes;
  out->ms_device = ev_ms(ctx->ev[0], ctx->ev[1]);
  return RK_OK;
}

int64_t rk_debug_fetch(rk_ctx *ctx, const char *name, void *host, uint64_t bytes) {
  if (!ctx || !name) return RK_ERR_ARG;
  if (!ctx->loaded) return fail(ctx, RK_ERR_STATE, "nothing loaded");
  const void *src = nullptr;
  u64 sz = (u64)ctx->m * 4;
  if (!strcmp(name, "rank_fidx")) src = ctx->fidx_r;
  else if (!strcmp(name, "k0_r")) src = ctx->k0_r;
  else if (!strcmp(name, "skx")) src = ctx->skx;
  else if (!strcmp(name, "rx")) src = ctx->rx;
  else if (!strcmp(name, "sky")) src = ctx->sky;
  else if (!strcmp(name, "ry")) src = ctx->ry;
  else if (!strcmp(name, "parent")) src = ctx->parent;
  else if (!strcmp(name, "gid_rank")) src = ctx->gid_rank;
  else if (!strcmp(name, "hkey")) src = ctx->h;
  else return fail(ctx, RK_ERR_ARG, "unknown array %s", name);
  if ((!strcmp(name, "parent") || !strcmp(name, "gid_rank")) && !ctx->have_group)
    return fail(ctx, RK_ERR_STATE, "no rk_group result yet");
  if (host && bytes >= sz && sz) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host, src, sz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return (int64_t)sz;
}

int rk_profile_enable(rk_ctx *ctx, int on) {
  if (!ctx) return RK_ERR_ARG;
  ctx->prof.on = on != 0;
  return RK_OK;
}

int rk_profile_read(rk_ctx *ctx, rk_kernel_time *out, int cap, int reset) {
  if (!ctx) return RK_ERR_ARG;
  int k = 0;
  for (int i = 0; i < KID_COUNT && out && k < cap; ++i) {
    if (!ctx->prof.launches[i]) continue;
    out[k].name = kKernelNames[i];
    out[k].launches = ctx->prof.launches[i];
    out[k].ms_total = ctx->prof.ms[i];
    out[k].units = ctx->prof.units[i];
    ++k;
  }
  if (reset) {
    for (int i = 0; i < KID_COUNT; ++i) ctx->prof.ms[i] = 0, ctx->prof.launches[i] = 0, ctx->prof.units[i] = 0;
  }
  return k;
}

void *rk_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void rk_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

uint64_t rk_sort_pairs_work_bytes(uint64_t n) { return sort_work_bytes(n ? n : 1); }

int rk_sort_pairs(rk_ctx *ctx, const uint32_t *keys_in, const uint32_t *values_in, uint32_t *keys_out, uint32_t *values_out,
                  uint32_t *keys_tmp, uint32_t *values_tmp, uint64_t n, int key_bits, void *work) {
  if (!ctx || !keys_in || !keys_out || !values_out || !keys_tmp || !values_tmp || !work) return RK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  launch_sort_pairs(keys_in, values_in, keys_out, values_out, keys_tmp, values_tmp, n, key_bits, work, ctx->stream);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return RK_OK;
}

// ---- scratch and error word of the calls that work on caller data (rk_sort_members) --------------------------

static int st_scratch(rk_ctx *ctx, u64 bytes, void **out) {
  if (bytes > ctx->st_scratch_bytes) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->st_scratch) cudaFree(ctx->st_scratch);
    ctx->st_scratch = nullptr;
    ctx->st_scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->st_scratch, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, RK_ERR_NOMEM, "cudaMalloc(%llu bytes): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    ctx->st_scratch_bytes = bytes;
  }
  *out = ctx->st_scratch;
  return RK_OK;
}

static int st_check_errors(rk_ctx *ctx, bool range_errors) {
  CK(cudaMemcpyAsync(ctx->h_cnt, ctx->st_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  const u32 e = ctx->h_cnt->err;
  if (e) {
    cudaMemsetAsync(&ctx->st_cnt->err, 0, sizeof(u32), ctx->stream);
    return fail(ctx, (range_errors && !(e & (ERR_WORKLIST | ERR_SPIN))) ? RK_ERR_RANGE : RK_ERR_INTERNAL, "%s", err_bits_text(e));
  }
  return RK_OK;
}

// sort_groups (src/commonFunctions.cpp:148-159) as the pure function the reference has: the order of the members of
// every group under std::sort by h = |yStart - diag_func[xStart/10]|, from the arguments alone — no state of an earlier
// rk_group is used.  The caller gathers y = f.yStart and d = diag_func[f.xStart/10] per member (loads, no arithmetic).
int rk_sort_members(rk_ctx *ctx, uint64_t m, const uint32_t *gid, const uint64_t *y, const uint64_t *d, uint32_t *perm) {
  if (!ctx || !ctx->st_cnt || (m && (!gid || !y || !d || !perm))) return RK_ERR_ARG;
  if (m >= 0xFFFFFFF0ull) return fail(ctx, RK_ERR_ARG, "more than 2^32-16 members");
  if (m == 0) return RK_OK;
  CK(cudaSetDevice(ctx->device));
  ProfGuard pg(ctx);
  cudaStream_t st = ctx->stream;
  const u64 a4 = align_up(m * 4, 256), a8 = align_up(m * 8, 256);
  void *scr = nullptr;
  const int rc = st_scratch(ctx, 2 * a8 + 7 * a4 + align_up(m, 256) + align_up(order_scratch_bytes(m), 256) + 256, &scr);
  if (rc != RK_OK) return rc;
  u8 *p = (u8 *)scr;
  u64 *d_y = (u64 *)p; p += a8;
  u64 *d_d = (u64 *)p; p += a8;
  u32 *d_gid = (u32 *)p; p += a4;
  u32 *d_h = (u32 *)p; p += a4;
  u32 *d_idx = (u32 *)p; p += a4;
  float *d_zero = (float *)p; p += a4;
  u32 *d_order = (u32 *)p; p += a4;
  u32 *d_ogid = (u32 *)p; p += a4;
  float *d_oident = (float *)p; p += a4;  // (identity is not part of this call: zeros in, zeros out)
  u8 *d_rep = p; p += align_up(m, 256);
  void *oscr = p;
  CK(cudaMemcpyAsync(d_y, y, m * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_d, d, m * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_gid, gid, m * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->st_cnt, 0, sizeof(Counters), st));
  launch_member_keys(d_y, d_d, (u32)m, d_h, d_idx, d_zero, &ctx->st_cnt->err, st);
  OrderArgs oa{};
  oa.sgid = d_gid, oa.srank = nullptr, oa.hfi_r = nullptr, oa.h = d_h, oa.fidx_r = d_idx, oa.identity_r = d_zero;
  oa.m = (u32)m, oa.do_sort = 1;
  order_carve(oa, oscr, m);
  oa.work_count = ctx->st_cnt->work_g;
  oa.out_order = d_order, oa.out_gid = d_ogid, oa.out_repval = d_rep, oa.out_identity = d_oident;
  oa.err = &ctx->st_cnt->err;
  launch_order(oa, st);
  CK(cudaMemcpyAsync(perm, d_order, m * 4, cudaMemcpyDeviceToHost, st));
  const int rc2 = st_check_errors(ctx, true);
  return rc2;
}

int rk_gen_workload(rk_ctx *ctx, uint64_t seed, uint64_t lx, uint64_t ly, double p_rep, uint64_t families, uint64_t ax, uint64_t ay,
                    uint64_t tandem_every, uint64_t start, uint64_t count, void *out_device) {
  if (!ctx || (count && !out_device)) return RK_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  launch_gen(seed, lx, ly, p_rep, families, ax, ay, tandem_every, start, count, (u8 *)out_device, ctx->stream);
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  return RK_OK;
}

}  // extern "C"
