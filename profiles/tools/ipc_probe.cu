// ipc_probe.cu — can two PROCESSES (one per GPU, as torchrun launches them) write into each other's device memory on this
// box, and what does a flag round trip cost?  (Design input for the slab path's peer-memory exchange)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ipc_probe ipc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("[%d] %s -> %s\n", me, #x, cudaGetErrorString(e)); exit(2); } } while (0)
static int me = 0;

__global__ void fill(uint32_t *dst, uint32_t n, uint32_t v) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = v + i;
}
__global__ void signal(volatile uint32_t *flag, uint32_t v) { __threadfence_system(); *flag = v; }
__global__ void wait_for(volatile uint32_t *flag, uint32_t v, uint32_t *timed_out) {
  const long long t0 = clock64();
  while (*flag < v) if (clock64() - t0 > 4000000000ll) { *timed_out = 1; return; }
}
__global__ void check(const uint32_t *src, uint32_t n, uint32_t v, uint32_t *bad) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (src[i] != v + i) atomicAdd(bad, 1u);
}

int main() {
  int up[2], down[2];
  if (pipe(up) || pipe(down)) return 1;
  const pid_t child = fork();
  me = child == 0 ? 1 : 0;
  int rd = me ? down[0] : up[0], wr = me ? up[1] : down[1];
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const int dev = ndev > 1 ? me : 0;
  CK(cudaSetDevice(dev));
  const uint32_t N = 1u << 22;  // 16 MB
  uint32_t *buf, *flags, *bad;
  CK(cudaMalloc(&buf, N * 4)); CK(cudaMalloc(&flags, 256)); CK(cudaMalloc(&bad, 8));
  CK(cudaMemset(flags, 0, 256)); CK(cudaMemset(bad, 0, 8)); CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mine[2], theirs[2];
  CK(cudaIpcGetMemHandle(&mine[0], buf)); CK(cudaIpcGetMemHandle(&mine[1], flags));
  if (write(wr, mine, sizeof(mine)) != sizeof(mine) || read(rd, theirs, sizeof(theirs)) != sizeof(theirs)) return 3;
  uint32_t *pbuf, *pflags;
  CK(cudaIpcOpenMemHandle((void **)&pbuf, theirs[0], cudaIpcMemLazyEnablePeerAccess));
  CK(cudaIpcOpenMemHandle((void **)&pflags, theirs[1], cudaIpcMemLazyEnablePeerAccess));
  printf("[%d] device %d of %d: peer buffers mapped\n", me, dev, ndev);
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // 1. push 16 MB into the peer, signal, peer waits and checks
  CK(cudaEventRecord(e0, s));
  fill<<<592, 256, 0, s>>>(pbuf, N, 1000u * (me + 1));
  signal<<<1, 1, 0, s>>>(pflags + 0, 1u);
  wait_for<<<1, 1, 0, s>>>(flags + 0, 1u, bad + 1);
  check<<<592, 256, 0, s>>>(buf, N, 1000u * (2 - me), bad);
  CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
  uint32_t h[2]; float ms;
  CK(cudaMemcpy(h, bad, 8, cudaMemcpyDeviceToHost)); CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("[%d] push 16 MB + flag + check: %.3f ms, mismatches %u, timed out %u\n", me, ms, h[0], h[1]);
  // 2. flag ping-pong: 200 round trips with one-thread kernels (signal peer, wait for peer)
  CK(cudaEventRecord(e0, s));
  for (uint32_t k = 2; k < 202; ++k) {
    signal<<<1, 1, 0, s>>>(pflags + 1, k);
    wait_for<<<1, 1, 0, s>>>(flags + 1, k, bad + 1);
  }
  CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("[%d] kernel-flag barrier: %.2f us per round (signal kernel + wait kernel)\n", me, ms * 1000.f / 200.f);
  // 3. the same with stream memory operations (no kernels)
  CUresult r = CUDA_SUCCESS;
  CK(cudaEventRecord(e0, s));
  for (uint32_t k = 2; k < 202 && r == CUDA_SUCCESS; ++k) {
    r = cuStreamWriteValue32((CUstream)s, (CUdeviceptr)(pflags + 2), k, 0);
    if (r == CUDA_SUCCESS) r = cuStreamWaitValue32((CUstream)s, (CUdeviceptr)(flags + 2), k, CU_STREAM_WAIT_VALUE_GEQ);
  }
  CK(cudaEventRecord(e1, s));
  if (r == CUDA_SUCCESS) {
    CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("[%d] stream-memop barrier: %.2f us per round\n", me, ms * 1000.f / 200.f);
  } else {
    const char *n = nullptr; cuGetErrorName(r, &n);
    printf("[%d] stream memops on IPC memory: %s\n", me, n ? n : "?");
    CK(cudaStreamSynchronize(s));
  }
  // keep the mappings alive until both sides are done
  char c = 'x';
  if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 4;
  cudaIpcCloseMemHandle(pbuf); cudaIpcCloseMemHandle(pflags);
  if (me == 0) { int st; waitpid(child, &st, 0); printf("child exit %d\n", WEXITSTATUS(st)); }
  return 0;
}
