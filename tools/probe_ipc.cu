// Probe for the multi-GPU design (tooling, not product): two PROCESSES, one GPU each.
//  1. cudaIpc: rank 0 exports a device buffer, rank 1 opens it and reads/writes it from a kernel (peer loads over NVLink);
//  2. NCCL through dlopen("libnccl.so.2"): ncclCommInitRank with a unique id passed over a pipe, send/recv both ways.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 tools/probe_ipc.cu -o gpurun_out/probe_ipc -ldl
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

static int g_rank = 0;
#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("[%d] %s failed: %s\n", g_rank, #x, cudaGetErrorString(e_));              \
      fflush(stdout);                                                                  \
      _exit(3);                                                                        \
    }                                                                                  \
  } while (0)

__global__ void k_fill(unsigned *p, unsigned n, unsigned v) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v + i;
}
__global__ void k_peer_sum(unsigned *peer, unsigned n, unsigned long long *out) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    atomicAdd(out, (unsigned long long)peer[i]);
    if (i < 4) peer[n + i] = 0xABCD0000u + i;  // write into the peer's buffer
  }
}

struct Id { char b[128]; };
typedef int (*fn_uid)(Id *);
typedef int (*fn_init)(void **, int, Id, int);
typedef int (*fn_sr)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_rr)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_v)();

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  int p01[2], p10[2];
  if (pipe(p01) || pipe(p10)) return 1;
  const pid_t pid = fork();
  g_rank = pid == 0 ? 1 : 0;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  printf("[%d] devices: %d\n", g_rank, ndev);
  if (ndev < 2) { printf("[%d] needs 2 GPUs\n", g_rank); return 0; }
  CK(cudaSetDevice(g_rank));
  const unsigned n = 1 << 20;
  void *nccl = dlopen("libnccl.so.2", RTLD_NOW);
  if (!nccl) printf("[%d] dlopen libnccl.so.2: %s\n", g_rank, dlerror());
  Id id{};
  unsigned *buf = nullptr;
  if (g_rank == 0) {
    CK(cudaMalloc(&buf, (n + 16) * 4));
    k_fill<<<n / 256, 256>>>(buf, n, 7);
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, buf));
    if (nccl) printf("[0] ncclGetUniqueId rc=%d\n", ((fn_uid)dlsym(nccl, "ncclGetUniqueId"))(&id));
    if (write(p01[1], &h, sizeof h) != sizeof h || write(p01[1], &id, sizeof id) != sizeof id) return 1;
    char done;
    if (read(p10[0], &done, 1) != 1) return 1;
    unsigned tail[4];
    CK(cudaMemcpy(tail, buf + n, 16, cudaMemcpyDeviceToHost));
    printf("[0] peer wrote %08x %08x %08x %08x (want abcd0000..3)\n", tail[0], tail[1], tail[2], tail[3]);
  } else {
    cudaIpcMemHandle_t h;
    if (read(p01[0], &h, sizeof h) != sizeof h || read(p01[0], &id, sizeof id) != sizeof id) return 1;
    unsigned *peer = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle((void **)&peer, h, cudaIpcMemLazyEnablePeerAccess);
    printf("[1] cudaIpcOpenMemHandle: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
      unsigned long long *out;
      CK(cudaMalloc(&out, 8));
      CK(cudaMemset(out, 0, 8));
      cudaEvent_t a, b;
      cudaEventCreate(&a), cudaEventCreate(&b);
      k_peer_sum<<<n / 256, 256>>>(peer, n, out);  // warm-up (maps the peer memory)
      CK(cudaDeviceSynchronize());
      CK(cudaMemset(out, 0, 8));
      cudaEventRecord(a);
      k_peer_sum<<<n / 256, 256>>>(peer, n, out);
      cudaEventRecord(b);
      CK(cudaDeviceSynchronize());
      unsigned long long s = 0;
      CK(cudaMemcpy(&s, out, 8, cudaMemcpyDeviceToHost));
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      const unsigned long long want = 7ull * n + (unsigned long long)n * (n - 1) / 2;
      printf("[1] peer sum %llu want %llu %s (%.3f ms for 4 MB of peer loads)\n", s, want, s == want ? "OK" : "MISMATCH", ms);
    } else {
      cudaGetLastError();
    }
    char done = 1;
    if (write(p10[1], &done, 1) != 1) return 1;
  }
  if (nccl) {
    void *comm = nullptr;
    int rc = ((fn_init)dlsym(nccl, "ncclCommInitRank"))(&comm, 2, id, g_rank);
    printf("[%d] ncclCommInitRank rc=%d\n", g_rank, rc);
    if (rc == 0) {
      unsigned *s, *r;
      CK(cudaMalloc(&s, 4096));
      CK(cudaMalloc(&r, 4096));
      k_fill<<<4, 256>>>(s, 1024, 1000 * g_rank);
      ((fn_v)dlsym(nccl, "ncclGroupStart"))();
      ((fn_sr)dlsym(nccl, "ncclSend"))(s, 4096, 0 /*ncclInt8*/, 1 - g_rank, comm, 0);
      ((fn_rr)dlsym(nccl, "ncclRecv"))(r, 4096, 0, 1 - g_rank, comm, 0);
      rc = ((fn_v)dlsym(nccl, "ncclGroupEnd"))();
      CK(cudaDeviceSynchronize());
      unsigned v[2];
      CK(cudaMemcpy(v, r, 8, cudaMemcpyDeviceToHost));
      printf("[%d] nccl send/recv rc=%d got %u %u (want %u %u)\n", g_rank, rc, v[0], v[1], 1000 * (1 - g_rank), 1000 * (1 - g_rank) + 1);
    }
  }
  if (g_rank == 0) {
    int st = 0;
    waitpid(pid, &st, 0);
    printf("[0] child exit %d\n", WEXITSTATUS(st));
  }
  return 0;
}
