// Link probe: how fast do SM stores into pinned host memory go, as a function of the size of the
// contiguous runs (32 / 64 / 128 B chunks) and of their density?  Compared with a copy-engine D2H.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/pcie_store_probe tools/probes/pcie_store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
template <int MODE>
__global__ void k_store(const uint4* __restrict__ src, uint4* __restrict__ dst, long long quads, int g4, unsigned thresh, unsigned salt) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  const unsigned chunk = (unsigned)(q / g4);
  const uint4 v = __ldcs(src + q);
  if (hash32(chunk ^ salt) <= thresh) {
    if (MODE == 0) dst[q] = v;
    if (MODE == 1) __stcs(dst + q, v);
    if (MODE == 2) __stwt(dst + q, v);
  }
}

int main() {
  const long long bytes = 48ll << 20, quads = bytes / 16;
  uint4 *src, *host, *hdev;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 1, bytes));
  CK(cudaHostAlloc(&host, bytes, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer(&hdev, host, 0));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a));
    for (int i = 0; i < 10; ++i) CK(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost));
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
  }
  printf("copy engine dense 48 MB: %.3f ms  %.1f GB/s\n", ms / 10, bytes / (ms / 10 * 1e-3) / 1e9);
  const int grans[3] = {2, 4, 8};
  const double dens[5] = {1.0, 0.5, 0.25, 0.125, 0.0625};
  for (int mode = 0; mode < 3; ++mode)
    for (int gi = 0; gi < 3; ++gi)
      for (int di = 0; di < 5; ++di) {
        const unsigned thresh = dens[di] >= 1.0 ? 0xffffffffu : (unsigned)(dens[di] * 4294967296.0);
        const int blocks = (int)((quads + 255) / 256);
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaEventRecord(a));
          for (int i = 0; i < 10; ++i) {
            if (mode == 0) k_store<0><<<blocks, 256>>>(src, hdev, quads, grans[gi], thresh, 77u * i);
            if (mode == 1) k_store<1><<<blocks, 256>>>(src, hdev, quads, grans[gi], thresh, 77u * i);
            if (mode == 2) k_store<2><<<blocks, 256>>>(src, hdev, quads, grans[gi], thresh, 77u * i);
          }
          CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        }
        const double shipped = bytes * dens[di];
        printf("mode %d chunk %3d B density %.4f: %.3f ms  shipped %.1f MB  %.1f GB/s\n", mode, grans[gi] * 16, dens[di], ms / 10,
               shipped / 1e6, shipped / (ms / 10 * 1e-3) / 1e9);
      }
  return 0;
}
