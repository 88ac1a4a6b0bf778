// cub_sort.cu — YARDSTICK ONLY (never linked into the product): cub::DeviceRadixSort::SortPairs on the key/value
// shape of the PBF step (u32 Morton key, u32 index, bits [0, 30)), timed with CUDA events, beside which
// profiles/README.md quotes the hand-written sort of csrc/sort_scan.cu ("CUB 2.8.2, comparison only").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cub_sort cub_sort.cu
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv) {
  for (int n : {1000000, 8000000}) {
    const int G = n == 1000000 ? 488063 : 3735478;  // cell-table sizes of dam-1m / dam-8m
    std::vector<unsigned> hk(n), hv(n);
    unsigned s = 12345u;
    for (int i = 0; i < n; ++i) {
      s = s * 1664525u + 1013904223u;
      hk[i] = (s >> 4) % G;
      if ((s & 0xFFFu) == 0) hk[i] |= (s >> 2) & 0x3FF00000u;  // a few particles predicted outside the grid: keys up to 30 bits
      hv[i] = i;
    }
    unsigned *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    cudaMemcpy(k0, hk.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, v0, v1, n, 0, 30);
    void *tmp; cudaMalloc(&tmp, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bits : {30, 20}) {
      for (int w = 0; w < 5; ++w) cub::DeviceRadixSort::SortPairs(tmp, bytes, k0, k1, v0, v1, n, 0, bits);
      cudaEventRecord(e0);
      const int reps = 50;
      for (int r = 0; r < reps; ++r) cub::DeviceRadixSort::SortPairs(tmp, bytes, k0, k1, v0, v1, n, 0, bits);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("cub::DeviceRadixSort::SortPairs n=%d bits [0,%d): %.1f us per sort (CUB %d.%d.%d)\n", n, bits, ms * 1000.f / reps,
             CUB_MAJOR_VERSION, CUB_MINOR_VERSION, CUB_SUBMINOR_VERSION);
    }
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
  }
  return 0;
}
