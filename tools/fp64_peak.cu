// fp64 roofline denominators for B200: DFMA (vector) and DMMA (mma.sync f64) peak, measured.
// MEASURED_PEAKS.json has no fp64 entry so the build measures its own.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per thread. NACC independent accumulators.
template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs per thread
template <int NACC>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  double a0 = a, a1 = a + 1, a2 = a + 2, a3 = a + 3, a4 = a, a5 = a, a6 = a, a7 = a;
  double b0 = b, b1 = b + 1, b2 = b, b3 = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7),
                     "d"(b0), "d"(b1), "d"(b2), "d"(b3));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs
template <int NACC>
__global__ void __launch_bounds__(256) dmma1684_kernel(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  double a0 = a, a1 = a + 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 20000;
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps;
    {
      float ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 16 * iters * 256.0 * grid;
      printf("DFMA  ilp16 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;  // 8 warps/block
      printf("DMMA m8n8k4 x8 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma1684_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 512 * 8 * iters * 8.0 * grid;
      printf("DMMA m16n8k4 x8 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma16816_kernel<4><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 2048 * 4 * (iters / 4) * 8.0 * grid;
      printf("DMMA m16n8k16 x4 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    }
  }
  // sustained: 2 s of DFMA back to back
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * 2; int n = 0;
    cudaEventRecord(e0);
    for (n = 0; n < 400; ++n) dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * 256.0 * grid * n;
    printf("DFMA sustained %d launches: %.1f ms  %.2f TFLOP/s\n", n, ms, fl / ms * 1e-9);
  }
  return 0;
}
