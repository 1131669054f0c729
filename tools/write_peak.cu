// Write-only HBM bandwidth of this GPU: cudaMemset and two fill kernels (plain and streaming 16-byte stores)
// over a 16 GiB buffer — the practical ceiling of a store-only kernel such as cov_build_kernel, next to the
// copy bandwidth in MEASURED_PEAKS.json.   nvcc -arch=sm_100a -O3 -o tools/write_peak tools/write_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int STREAMING>
__global__ void __launch_bounds__(256) fill_kernel(double2* p, size_t n, double v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (STREAMING)
      asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p + i), "d"(v), "d"(v) : "memory");
    else
      p[i] = make_double2(v, v);
  }
}

int main() {
  const size_t bytes = 16ull << 30;
  double2* buf;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t n = bytes / sizeof(double2);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) cudaMemsetAsync(buf, 0, bytes);
      if (mode == 1) fill_kernel<0><<<148 * 8, 256>>>(buf, n, 1.0);
      if (mode == 2) fill_kernel<1><<<148 * 8, 256>>>(buf, n, 1.0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const char* names[3] = {"cudaMemset", "fill kernel, st.global.v2.f64", "fill kernel, st.global.cs.v2.f64"};
    printf("%-36s %8.1f GB/s (best of 5, 16 GiB)\n", names[mode], bytes / best * 1e-6);
  }
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
