// tcgen05.mma kind::i8 probe for sm_100a (B200): settles, in ONE box visit, the facts the int8 (Ozaki) trailing
// update depends on and that cannot be checked without a GPU:
//   1. which shared-memory descriptor encodings (no swizzle / 32B / 64B / 128B swizzle, K-major, arbitrary SBO)
//      produce D = A·Bᵀ exactly (checked against a CPU int32 GEMM),
//   2. the dispatch pace of 128×N×32 int8 MMAs from shared memory for N = 64, 128, 256 (is 128×64 smem-bound?),
//   3. the TMEM→register read rate of the epilogue (tcgen05.ld 32x32b).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/umma_i8_probe tools/exp/umma_i8_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) {
  }
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ inline uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// correctness: one CTA, operands written to shared memory by the threads in the layout under test
// ------------------------------------------------------------------------------------------------
struct Variant {
  int layout;      // descriptor layout code: 0 none, 6 = 32B, 4 = 64B, 2 = 128B
  int row_bytes;   // bytes of K per row held in one slab (16-byte interleave for layout 0)
  int swz_bits;    // XOR bits [4,4+swz_bits) with bits [7,7+swz_bits)
  int sbo;         // bytes between 8-row groups
  int lbo;         // bytes between the two 16-byte K chunks (no swizzle only)
  int ksteps;      // number of K=32 MMAs (start address advanced by 32 B each)
};

// byte offset of element (row r, byte k) inside a slab
__host__ __device__ inline uint32_t slab_off(const Variant& v, int r, int k) {
  if (v.layout == 0) {  // core matrices of 8 rows × 16 B
    return (uint32_t)((r >> 3) * v.sbo + (k >> 4) * v.lbo + (r & 7) * 16 + (k & 15));
  }
  uint32_t off = (uint32_t)((r >> 3) * v.sbo + (r & 7) * v.row_bytes + k);
  const uint32_t mask = (1u << v.swz_bits) - 1;
  off ^= ((off >> 7) & mask) << 4;
  return off;
}

constexpr int M_ = 128;

__global__ void __launch_bounds__(128, 1)
    probe_kernel(const int8_t* A, const int8_t* B, int N, int K, Variant v, int32_t* D, int a_bytes, int b_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  for (int i = tid; i < a_bytes + b_bytes; i += blockDim.x) smem[i] = 0;
  __syncthreads();
  for (int i = tid; i < M_ * K; i += blockDim.x) sA[slab_off(v, i / K, i % K)] = (uint8_t)A[i];
  for (int i = tid; i < N * K; i += blockDim.x) sB[slab_off(v, i / K, i % K)] = (uint8_t)B[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 256);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_i8(M_, N);
    for (int ks = 0; ks < v.ksteps; ++ks) {
      const uint32_t adv = (v.layout == 0) ? ks * 2 * v.lbo : ks * 32;
      const uint64_t ad = make_desc(smem_u32(sA) + adv, v.lbo, v.sbo, v.layout);
      const uint64_t bd = make_desc(smem_u32(sB) + adv, v.lbo, v.sbo, v.layout);
      umma_i8(tb, ad, bd, idesc, ks > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

// ------------------------------------------------------------------------------------------------
// pace: the issue loop of the real kernel (ozaki.cu): per chunk 26 MMAs of 128×N×32 over six A and six B slices
// laid out [row group][slice][8 rows][32 B] (SW32, SBO = 1536), NACC accumulators, warp-uniform loop with one
// elected lane; cycles by clock64 around the loop + final commit.  Then the epilogue read of 448 TMEM columns.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
__host__ __device__ inline uint64_t desc_sw32(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1536 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)6 << 61);
}
template <int N>
__global__ void __launch_bounds__(128, 1) pace_kernel(int chunks, int nstage, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int NACC = (N == 64) ? 7 : (N == 128 ? 3 : 2);
  constexpr int A_BYTES = 16 * 1536, B_BYTES = (N / 8) * 1536, STAGE = A_BYTES + B_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < nstage * STAGE; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_i8(M_, N);
    const long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      const uint32_t a0 = smem_u32(smem) + (c % nstage) * STAGE, b0 = a0 + A_BYTES;
      const uint64_t ad0 = desc_sw32(a0), bd0 = desc_sw32(b0);
      if (leader) {
#pragma unroll
        for (int sa = 0; sa < 6; ++sa)
#pragma unroll
          for (int sb = 0; sb < 6; ++sb) {
            if (sa + sb >= 7) continue;
            umma_i8(tb + (uint32_t)((sa + sb) % NACC) * N, ad0 + sa * 16, bd0 + sb * 16, idesc, 1);
          }
      }
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (leader) out[0] = clock64() - t0;
  } else {
    mbar_wait(smem_u32(&bar), 0);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  long long e0 = clock64();
  uint32_t sink = 0;
  for (int c0 = 0; c0 < 448; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) sink ^= r[j];
  }
  long long e1 = clock64();
  if (sink == 0x12345678u) out[3] = 1;
  __syncthreads();
  if (tid == 0) out[1] = e1 - e0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N>
void run_pace(int sms) {
  constexpr int STAGE = 16 * 1536 + (N / 8) * 1536;
  const int nstage = (N == 256) ? 2 : 4;
  long long* dout;
  CK(cudaMalloc(&dout, 64));
  CK(cudaFuncSetAttribute(pace_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int chunks = 256;
  CK(cudaMemset(dout, 0, 64));
  pace_kernel<N><<<1, 128, nstage * STAGE + 1024>>>(chunks, nstage, dout);
  CK(cudaDeviceSynchronize());
  long long h[4];
  CK(cudaMemcpy(h, dout, 32, cudaMemcpyDeviceToHost));
  printf("pace 128x%dx32 i8, real issue loop (26 MMAs/chunk): %.1f cycles per MMA (floor %d) -> %.0f MAC/clk/SM; epilogue read 448 cols: %lld cycles\n",
         N, (double)h[0] / (chunks * 26.0), 128 * N / 256, 128.0 * N * 32 * chunks * 26.0 / (double)h[0], h[1]);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int big = 40000;
  pace_kernel<N><<<sms, 128, nstage * STAGE + 1024>>>(100, nstage, dout);
  CK(cudaEventRecord(a));
  pace_kernel<N><<<sms, 128, nstage * STAGE + 1024>>>(big, nstage, dout);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double ops = 2.0 * 128 * N * 32 * 26.0 * (double)big * sms;
  printf("chip 128x%dx32 i8 on %d SMs: %.2f ms -> %.1f TOP/s\n", N, sms, ms, ops / ms * 1e-9);
  cudaFree(dout);
}


// ------------------------------------------------------------------------------------------------
// A operand from TMEM: tcgen05.cp (smem -> TMEM, 128 lanes x 256 bit) of a K-major SW32 A slab, read back with
// tcgen05.ld (layout check), then the .ts form of the MMA (A in TMEM, B in smem) against the CPU GEMM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) ts_kernel(const int8_t* A, const int8_t* B, int N, int32_t* D, uint32_t* Aback) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const Variant v{6, 32, 1, 256, 16, 1};
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8192;
  for (int i = tid; i < 16384; i += blockDim.x) smem[i] = 0;
  __syncthreads();
  for (int i = tid; i < M_ * 32; i += blockDim.x) sA[slab_off(v, i / 32, i % 32)] = (uint8_t)A[i];
  for (int i = tid; i < N * 32; i += blockDim.x) sB[slab_off(v, i / 32, i % 32)] = (uint8_t)B[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t ta = tb + 256;   // A operand columns
  if (tid == 0) {
    utccp_128x256b(ta, make_desc(smem_u32(sA), 16, 256, 6));
    umma_i8_ts(tb, ta, make_desc(smem_u32(sB), 16, 256, 6), make_idesc_i8(M_, N), 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[8];
    tmem_ld8(ta + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) Aback[tid * 8 + j] = r[j];
  }
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// pace of the TS form: per chunk 6 tcgen05.cp (A slices -> 48 TMEM columns) + 26 MMAs with A from TMEM
__global__ void __launch_bounds__(128, 1) pace_ts_kernel(int chunks, int nstage, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int N = 64, A_BYTES = 16 * 1536, B_BYTES = (N / 8) * 1536, STAGE = A_BYTES + B_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < nstage * STAGE; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_i8(M_, N);
    const long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      const uint32_t a0 = smem_u32(smem) + (c % nstage) * STAGE, b0 = a0 + A_BYTES;
      const uint64_t ad0 = desc_sw32(a0), bd0 = desc_sw32(b0);
      if (leader) {
#pragma unroll
        for (int sa = 0; sa < 6; ++sa) utccp_128x256b(tb + 448 + sa * 8, ad0 + sa * 16);
#pragma unroll
        for (int sa = 0; sa < 6; ++sa)
#pragma unroll
          for (int sb = 0; sb < 6; ++sb) {
            if (sa + sb >= 7) continue;
            umma_i8_ts(tb + (uint32_t)(sa + sb) * N, tb + 448 + sa * 8, bd0 + sb * 16, idesc, 1);
          }
      }
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (leader) out[0] = clock64() - t0;
  } else {
    mbar_wait(smem_u32(&bar), 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

void run_ts(int sms) {
  const int N = 64;
  std::vector<int8_t> A(M_ * 32), B(N * 32);
  srand(77);
  for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB;
  int32_t* dD;
  uint32_t* dAb;
  CK(cudaMalloc(&dA, A.size()));
  CK(cudaMalloc(&dB, B.size()));
  CK(cudaMalloc(&dD, sizeof(int32_t) * M_ * N));
  CK(cudaMalloc(&dAb, sizeof(uint32_t) * M_ * 8));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, sizeof(int32_t) * M_ * N));
  CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  ts_kernel<<<1, 128, 20 * 1024>>>(dA, dB, N, dD, dAb);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("TS form: CUDA error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  std::vector<int32_t> D(M_ * N);
  std::vector<uint32_t> Ab(M_ * 8);
  CK(cudaMemcpy(D.data(), dD, sizeof(int32_t) * M_ * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Ab.data(), dAb, sizeof(uint32_t) * M_ * 8, cudaMemcpyDeviceToHost));
  long long badA = 0;
  for (int r = 0; r < M_; ++r)
    for (int j = 0; j < 8; ++j) {
      uint32_t w = 0;
      for (int b = 0; b < 4; ++b) w |= (uint32_t)(uint8_t)A[r * 32 + 4 * j + b] << (8 * b);
      if (w != Ab[r * 8 + j]) ++badA;
    }
  printf("tcgen05.cp 128x256b (SW32 K-major slab -> TMEM): lane r / column j = bytes 4j..4j+3 of row r: %s (%lld of %d words differ)\n",
         badA ? "NO" : "yes", badA, M_ * 8);
  if (badA) {
    printf("  row 0 got: ");
    for (int j = 0; j < 8; ++j) printf("%08x ", Ab[j]);
    printf("\n  row 0 exp: ");
    for (int j = 0; j < 8; ++j) {
      uint32_t w = 0;
      for (int b = 0; b < 4; ++b) w |= (uint32_t)(uint8_t)A[4 * j + b] << (8 * b);
      printf("%08x ", w);
    }
    printf("\n");
  }
  long long bad = 0;
  for (int i = 0; i < M_; ++i)
    for (int j = 0; j < N; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < 32; ++k) ref += (int32_t)A[i * 32 + k] * (int32_t)B[j * 32 + k];
      if (ref != D[i * N + j]) ++bad;
    }
  printf("tcgen05.mma .ts form (A from TMEM): %s (%lld mismatches of %d)\n", bad ? "WRONG" : "exact", bad, M_ * N);
  long long* dout;
  CK(cudaMalloc(&dout, 64));
  CK(cudaFuncSetAttribute(pace_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int STAGE = 16 * 1536 + 8 * 1536, nstage = 4, chunks = 256;
  pace_ts_kernel<<<1, 128, nstage * STAGE + 1024>>>(chunks, nstage, dout);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
  printf("pace TS 128x64x32 i8 (6 cp + 26 MMAs per chunk): %.1f cycles per MMA incl. copies (SS form: 52.4; floor 32)\n",
         (double)h[0] / (chunks * 26.0));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int big = 40000;
  pace_ts_kernel<<<sms, 128, nstage * STAGE + 1024>>>(100, nstage, dout);
  CK(cudaEventRecord(a));
  pace_ts_kernel<<<sms, 128, nstage * STAGE + 1024>>>(big, nstage, dout);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("chip TS 128x64x32 i8 on %d SMs: %.2f ms -> %.1f TOP/s\n", sms, ms, 2.0 * 128 * N * 32 * 26.0 * big * sms / ms * 1e-9);
}


// ------------------------------------------------------------------------------------------------
// CTA pair (cta_group::2): one MMA of M = 256 (128 rows from each CTA's shared memory) x N = 64 (32 B rows from each
// CTA) x K = 32, issued by the leader CTA; accumulators in each CTA's own TMEM.  Correctness against the CPU GEMM and
// the dispatch pace of the 26-pair issue pattern.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    pair_kernel(const int8_t* A /*256x32*/, const int8_t* B /*64x32*/, int32_t* D /*256x64*/, int chunks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const Variant v{6, 32, 1, 256, 16, 1};
  // per CTA: A rows 128*rank.., B rows 32*rank..  (same shared-memory offsets in both CTAs)
  uint8_t* sA = smem;            // 6 slabs of 4 KB (only slab 0 carries the test data; the others feed the pace loop)
  uint8_t* sB = smem + 6 * 4096; // 6 slabs of 1 KB
  for (int i = tid; i < 6 * 4096 + 6 * 1024; i += blockDim.x) smem[i] = (uint8_t)(i * 5 + 1);
  __syncthreads();
  for (int i = tid; i < 128 * 32; i += blockDim.x) sA[slab_off(v, i / 32, i % 32)] = (uint8_t)A[(128 * rank + i / 32) * 32 + i % 32];
  for (int i = tid; i < 32 * 32; i += blockDim.x) sB[slab_off(v, i / 32, i % 32)] = (uint8_t)B[(32 * rank + i / 32) * 32 + i % 32];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc2(smem_u32(&tmem_base), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t idesc = make_idesc_i8(256, 64);
  long long t0 = 0;
  if (rank == 0 && warp == 1) {
    const bool leader = elect_one();
    if (leader) umma_i8_2cta(tb, make_desc(smem_u32(sA), 16, 256, 6), make_desc(smem_u32(sB), 16, 256, 6), idesc, 0);
    t0 = clock64();
    for (int c = 0; c < chunks; ++c) {   // pace: 26-pair pattern into accumulators 1..6 (accumulator 0 keeps the test result)
      if (leader) {
#pragma unroll
        for (int sa = 0; sa < 6; ++sa)
#pragma unroll
          for (int sb = 0; sb < 6; ++sb) {
            if (sa + sb >= 7) continue;
            umma_i8_2cta(tb + (uint32_t)(1 + (sa + sb) % 6) * 64, make_desc(smem_u32(sA) + sa * 4096, 16, 256, 6),
                         make_desc(smem_u32(sB) + sb * 1024, 16, 256, 6), idesc, 1);
          }
      }
      __syncwarp();
    }
    if (leader) umma_commit_2cta(smem_u32(&bar));
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  if (rank == 0 && tid == 32) out[0] = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)(128 * rank + tid) * 64 + c0 + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tb, 512);
}

void run_pair(int sms) {
  std::vector<int8_t> A(256 * 32), B(64 * 32);
  srand(99);
  for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
  for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB;
  int32_t* dD;
  long long* dout;
  CK(cudaMalloc(&dA, A.size()));
  CK(cudaMalloc(&dB, B.size()));
  CK(cudaMalloc(&dD, sizeof(int32_t) * 256 * 64));
  CK(cudaMalloc(&dout, 64));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, sizeof(int32_t) * 256 * 64));
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int chunks = 256;
  pair_kernel<<<2, 128, 32 * 1024>>>(dA, dB, dD, chunks, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("cta_group::2: CUDA error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  std::vector<int32_t> D(256 * 64);
  long long h[1];
  CK(cudaMemcpy(D.data(), dD, sizeof(int32_t) * 256 * 64, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < 256; ++i)
    for (int j = 0; j < 64; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < 32; ++k) ref += (int32_t)A[i * 32 + k] * (int32_t)B[j * 32 + k];
      if (ref != D[i * 64 + j]) ++bad;
    }
  printf("cta_group::2 MMA 256x64x32 i8 (A rows 128/CTA, B rows 32/CTA): %s (%lld mismatches of %d)\n", bad ? "WRONG" : "exact", bad,
         256 * 64);
  printf("pace cta_group::2 256x64x32 i8, 26-pair pattern: %.1f cycles per MMA = %.1f per 128x64x32 equivalent (1-CTA SS form: 52.4)\n",
         (double)h[0] / (chunks * 26.0), (double)h[0] / (chunks * 26.0) / 2);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int big = 40000, nblk = sms & ~1;
  pair_kernel<<<nblk, 128, 32 * 1024>>>(dA, dB, dD, 100, dout);
  CK(cudaEventRecord(a));
  pair_kernel<<<nblk, 128, 32 * 1024>>>(dA, dB, dD, big, dout);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("chip cta_group::2 256x64x32 i8 on %d SMs: %.2f ms -> %.1f TOP/s\n", nblk, ms,
         2.0 * 256 * 64 * 32 * 26.0 * big * (nblk / 2) / ms * 1e-9);
}

// ------------------------------------------------------------------------------------------------
// What bounds the TS form once the A copies leave the tensor pipe's issue stream?  A resident in TMEM, B from shared
// memory, the 26-pair pattern per chunk, one CTA per SM, cycles per MMA by clock64 around the whole loop:
//   MODE 0  MMAs only, accumulate flag a literal
//   MODE 1  MMAs only, accumulate flag from a run-time bit mask (the kernel's `touched` logic)
//   MODE 2  MODE 0 + per chunk four mbarrier try_waits on completed barriers and four commits (the handshake traffic
//           of syrk_i8_st_kernel without anybody on the other side)
//   MODE 3  two issuing warps, alternating chunks, literal flag
//   MODE 4  MODE 0 + four loader warps streaming LDS.128 -> tcgen05.st into the A columns (no handshake): does the
//           register -> TMEM path disturb the MMAs?
//   MODE 5  MODE 2 with the try_waits issued one step ahead (result consumed after the MMAs)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
      "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int MODE>
__global__ void __launch_bounds__(320, 1) issue_kernel(int chunks, int nstage, uint32_t touched0, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, done[8], stop;
  __shared__ uint32_t tmem_base;
  constexpr int N = 64, A_BYTES = 16 * 1536, B_BYTES = (N / 8) * 1536, STAGE = A_BYTES + B_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < nstage * STAGE; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), MODE == 3 ? 2 : 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&done[i]), 1);
    mbar_init(smem_u32(&stop), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t idesc = make_idesc_i8(M_, N);
  if (warp == 1 || (MODE == 3 && warp == 2)) {
    const bool leader = elect_one();
    uint32_t touched = touched0;
    const long long t0 = clock64();
    uint32_t hs = 0;  // handshake counter (MODE 2/5)
    bool ok_next = true;
    for (int c = (MODE == 3 ? warp - 1 : 0); c < chunks; c += (MODE == 3 ? 2 : 1)) {
      const uint32_t b0 = smem_u32(smem) + (c % nstage) * STAGE + A_BYTES;
      const uint64_t bd0 = desc_sw32(b0);
      if (MODE == 2) {
        // nobody ever arrives on done[0..4]: a wait for parity 1 of a fresh barrier succeeds at once
        mbar_wait(smem_u32(&done[4]), 1);  // parity 1 of a fresh barrier: complete
      }
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        if (MODE == 2) {
          mbar_wait(smem_u32(&done[(hs++) & 3]), 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (MODE == 5) {
          while (!ok_next) ok_next = mbar_test(smem_u32(&done[hs & 3]), 1);
          ++hs;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          ok_next = mbar_test(smem_u32(&done[hs & 3]), 1);
        }
        if (leader) {
          uint32_t local = 0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int sa = h ? 5 - pr : pr;
#pragma unroll
            for (int sb = 0; sb < 6; ++sb) {
              if (sa + sb >= 7) continue;
              const uint32_t d = sa + sb;
              uint32_t acc = 1;
              if (MODE == 1) { acc = ((touched | local) >> d) & 1u; local |= 1u << d; }
              umma_i8_ts(tb + d * N, tb + 448 + ((c * 3 + pr) & 3) * 16 + h * 8, bd0 + sb * 16, idesc, acc);
            }
          }
          touched |= local;
          if (MODE == 2 || MODE == 5) umma_commit(smem_u32(&done[5 + ((hs) & 1)]));
        }
        __syncwarp();
      }
      if (MODE == 2 || MODE == 5) {
        if (leader) umma_commit(smem_u32(&done[7]));
        __syncwarp();
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (leader && warp == 1) out[0] = clock64() - t0;
    if (leader && warp == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&stop)) : "memory");
  } else if (MODE == 4 && warp >= 6) {
    // loader warps: free-running LDS -> tcgen05.st into the A columns until the MMA warp is done
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t aoff = (uint32_t)(row >> 3) * 1536 + (uint32_t)(row & 7) * 32;
    const uint32_t h0 = (uint32_t)((row & 7) >> 2) * 16, h1 = h0 ^ 16;
    const uint32_t tl = tb + ((uint32_t)(q * 32) << 16) + 448;
    long long n = 0;
    for (int c = 0; !mbar_test(smem_u32(&stop), 0); ++c) {
      const uint32_t base = smem_u32(smem) + (c % nstage) * STAGE + aoff;
      uint4 lo[6], hi[6];
#pragma unroll
      for (int sa = 0; sa < 6; ++sa) {
        lo[sa] = lds128(base + sa * 256 + h0);
        hi[sa] = lds128(base + sa * 256 + h1);
      }
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        tmem_st16(tl + 16 * ((c * 3 + pr) & 3), lo[pr], hi[pr], lo[5 - pr], hi[5 - pr]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      ++n;
    }
    if (tid == 6 * 32) out[1] = n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int MODE>
void run_issue_mode(int sms, const char* what) {
  long long* dout;
  CK(cudaMalloc(&dout, 64));
  CK(cudaMemset(dout, 0, 64));
  CK(cudaFuncSetAttribute(issue_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int STAGE = 16 * 1536 + 8 * 1536, nstage = 4, chunks = 512;
  issue_kernel<MODE><<<1, 320, nstage * STAGE + 1024>>>(chunks, nstage, 0u, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("issue mode %d: CUDA error %s\n", MODE, cudaGetErrorString(e)); exit(1); }
  long long h[2];
  CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int big = 20000;
  issue_kernel<MODE><<<sms, 320, nstage * STAGE + 1024>>>(100, nstage, 0u, dout);
  CK(cudaEventRecord(a));
  issue_kernel<MODE><<<sms, 320, nstage * STAGE + 1024>>>(big, nstage, 0u, dout);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("issue pace mode %d (%s): %.1f cycles per MMA on one SM", MODE, what, (double)h[0] / (chunks * 26.0));
  if (MODE == 4) printf(" (loaders: %.2f chunks of A stored per MMA chunk)", (double)h[1] / chunks);
  printf("; chip %.1f TOP/s\n", 2.0 * 128 * 64 * 32 * 26.0 * big * sms / ms * 1e-9);
  cudaFree(dout);
}
void run_issue(int sms) {
  run_issue_mode<0>(sms, "TS MMAs only, literal accumulate flag");
  run_issue_mode<1>(sms, "TS MMAs only, run-time accumulate flag");
  run_issue_mode<2>(sms, "+ 4 try_waits and 4 commits per chunk");
  run_issue_mode<5>(sms, "+ 4 try_waits one step ahead and 4 commits per chunk");
  run_issue_mode<3>(sms, "two issuing warps, alternating chunks");
  run_issue_mode<4>(sms, "+ free-running loader warps LDS -> tcgen05.st");
}

// ------------------------------------------------------------------------------------------------
// Wide-run schedule of syrk_i8_kernel: per chunk 9 SS MMAs — A slab sa (4 KB) against runs of consecutive B slabs
// (3+3, 3+3, 3+2, 4, 3, 2 slabs -> N = 128..256), operands resident in shared memory.
//   MODE 0  MMAs only
//   MODE 1  + a producer warp streaming 36 KB of bulk copies per chunk from global memory into the ring (full/empty
//           barriers as in the kernel): what the shared-memory writes of the operand stream cost the MMAs
//   MODE 2  MODE 0 with single-slab MMAs only (26 per chunk, N = 64) for reference
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_sw32_sbo256(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)16 << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128, 1) wide_kernel(int chunks, const uint8_t* src, int fstages, long long* out, int gap = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NST = 5, STAGE = 36864, A_BYTES = 24576;
  __shared__ uint64_t full[NST], empty[NST], bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < NST * STAGE; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < NST; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (warp == 2 && MODE == 1) {
    if (elect_one()) {
      for (int c = 0; c < chunks; ++c) {
        const int st = c % NST;
        if (c >= NST) mbar_wait(smem_u32(&empty[st]), ((c / NST) - 1) & 1);
        const uint32_t fb = smem_u32(&full[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(STAGE) : "memory");
        const uint8_t* g = src + (size_t)(((size_t)blockIdx.x * 37 + c) % fstages) * STAGE;
        for (int sl = 0; sl < 6; ++sl) {
          bulk_g2s(smem_u32(smem) + st * STAGE + sl * 4096, g + sl * 4096, 4096, fb);
          bulk_g2s(smem_u32(smem) + st * STAGE + A_BYTES + sl * 2048, g + A_BYTES + sl * 2048, 2048, fb);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      const int st = c % NST;
      if (MODE == 1) {
        mbar_wait(smem_u32(&full[st]), (c / NST) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t a0 = smem_u32(smem) + st * STAGE, b0 = a0 + A_BYTES;
      const uint64_t ad0 = desc_sw32_sbo256(a0), bd0 = desc_sw32_sbo256(b0);
      if (leader) {
        if (MODE == 2) {
#pragma unroll
          for (int sa = 0; sa < 6; ++sa)
#pragma unroll
            for (int sb = 0; sb < 6; ++sb) {
              if (sa + sb >= 7) continue;
              umma_i8(tb + (uint32_t)(sa + sb) * 64, ad0 + sa * 256, bd0 + sb * 128, make_idesc_i8(128, 64), 1);
            }
        } else {
          // (sa, sb0, len)
          constexpr int R[9][3] = {{0, 0, 3}, {0, 3, 3}, {1, 0, 3}, {1, 3, 3}, {2, 0, 3}, {2, 3, 2}, {3, 0, 4}, {4, 0, 3}, {5, 0, 2}};
#pragma unroll
          for (int i = 0; i < 9; ++i)
            umma_i8(tb + (uint32_t)(R[i][0] + R[i][1]) * 64, ad0 + R[i][0] * 256, bd0 + R[i][1] * 128, make_idesc_i8(128, R[i][2] * 64), 1);
        }
        if (MODE == 1) umma_commit(smem_u32(&empty[st]));
      }
      __syncwarp();
      if (gap > 0) {  // issue-thread work between two chunks' MMAs (how much does the tensor pipe's queue hide?)
        const long long g0 = clock64();
        while (clock64() - g0 < gap) {
        }
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (leader) out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}
template <int MODE>
void run_wide_mode(int sms, const uint8_t* src, int fstages, const char* what) {
  long long* dout;
  CK(cudaMalloc(&dout, 64));
  CK(cudaMemset(dout, 0, 64));
  const int smem_bytes = 5 * 36864 + 1024;
  CK(cudaFuncSetAttribute(wide_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int chunks = 2048;
  wide_kernel<MODE><<<1, 128, smem_bytes>>>(chunks, src, fstages, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("wide mode %d: CUDA error %s\n", MODE, cudaGetErrorString(e)); exit(1); }
  long long h1[1], h2[1];
  CK(cudaMemcpy(h1, dout, 8, cudaMemcpyDeviceToHost));
  wide_kernel<MODE><<<sms, 128, smem_bytes>>>(chunks, src, fstages, dout);
  CK(cudaDeviceSynchronize());
  wide_kernel<MODE><<<sms, 128, smem_bytes>>>(chunks, src, fstages, dout);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h2, dout, 8, cudaMemcpyDeviceToHost));
  printf("wide-run schedule mode %d (%s; operand footprint %.0f MB): %.0f cycles per chunk on one SM, %.0f with all %d SMs running (floor 832)\n",
         MODE, what, fstages * 36864e-6, (double)h1[0] / chunks, (double)h2[0] / chunks, sms);
  cudaFree(dout);
}
void run_wide_gap(int sms, const uint8_t* src) {
  long long* dout;
  CK(cudaMalloc(&dout, 64));
  const int smem_bytes = 5 * 36864 + 1024;
  CK(cudaFuncSetAttribute(wide_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int chunks = 2048;
  const int gaps[] = {0, 50, 100, 150, 200, 300, 400, 600};
  printf("wide-run schedule, a busy-wait of G cycles between two chunks' MMAs (cycles per chunk; 832 + G would mean nothing is hidden):");
  for (int g : gaps) {
    wide_kernel<0><<<sms, 128, smem_bytes>>>(chunks, src, 1024, dout, g);
    CK(cudaDeviceSynchronize());
    long long h[1];
    CK(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
    printf("  G=%d: %.0f", g, (double)h[0] / chunks);
  }
  printf("\n");
  cudaFree(dout);
}
void run_wide(int sms) {
  uint8_t* src;
  const int fmax = 148 * 64;
  const size_t bytes = (size_t)fmax * 36864;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 1, bytes));
  run_wide_mode<0>(sms, src, fmax, "9 wide SS MMAs, operands resident");
  run_wide_mode<2>(sms, src, fmax, "26 single-slab SS MMAs, operands resident");
  run_wide_mode<1>(sms, src, fmax, "9 wide SS MMAs + 36 KB of bulk copies per chunk");   // 350 MB: from HBM
  run_wide_mode<1>(sms, src, 1024, "9 wide SS MMAs + 36 KB of bulk copies per chunk");   // 38 MB: from L2
  run_wide_mode<1>(sms, src, 256, "9 wide SS MMAs + 36 KB of bulk copies per chunk");    // 9 MB: from L2, many readers per line
  run_wide_gap(sms, src);
  cudaFree(src);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));

  const int N = 64;
  struct Named { const char* name; Variant v; };
  const Named vars[] = {
      {"none  K=32  lbo=128 sbo=256", {0, 16, 0, 256, 128, 1}},
      {"none  K=64  lbo=128 sbo=512 (2 steps)", {0, 16, 0, 512, 128, 2}},
      {"sw32  K=32  sbo=256", {6, 32, 1, 256, 16, 1}},
      {"sw32  K=32  sbo=1536 (slices interleaved)", {6, 32, 1, 1536, 16, 1}},
      {"sw64  K=64  sbo=512  (2 steps)", {4, 64, 2, 512, 16, 2}},
      {"sw128 K=128 sbo=1024 (4 steps)", {2, 128, 3, 1024, 16, 4}},
      {"sw128 K=128 sbo=1024 lbo=0 (4 steps)", {2, 128, 3, 1024, 0, 4}},
  };
  for (const Named& nv : vars) {
    const Variant v = nv.v;
    const int K = 32 * v.ksteps;
    std::vector<int8_t> A(M_ * K), B(N * K);
    srand(1234);
    for (auto& x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto& x : B) x = (int8_t)(rand() % 256 - 128);
    const int a_bytes = ((M_ / 8) * v.sbo + 1023) / 1024 * 1024 + 1024;
    const int b_bytes = ((N / 8) * v.sbo + 1023) / 1024 * 1024 + 1024;
    int8_t *dA, *dB;
    int32_t* dD;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, sizeof(int32_t) * M_ * N));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, sizeof(int32_t) * M_ * N));
    probe_kernel<<<1, 128, a_bytes + b_bytes>>>(dA, dB, N, K, v, dD, a_bytes, b_bytes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-45s : CUDA error %s\n", nv.name, cudaGetErrorString(e));
      return 1;
    }
    std::vector<int32_t> D(M_ * N);
    CK(cudaMemcpy(D.data(), dD, sizeof(int32_t) * M_ * N, cudaMemcpyDeviceToHost));
    long long bad = 0, maxd = 0;
    for (int i = 0; i < M_; ++i)
      for (int j = 0; j < N; ++j) {
        int32_t ref = 0;
        for (int k = 0; k < K; ++k) ref += (int32_t)A[i * K + k] * (int32_t)B[j * K + k];
        long long d = llabs((long long)ref - D[i * N + j]);
        if (d) ++bad;
        if (d > maxd) maxd = d;
      }
    printf("%-45s : %s (mismatches %lld of %d, max |diff| %lld)\n", nv.name, bad ? "WRONG" : "exact", bad, M_ * N, maxd);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  run_wide(prop.multiProcessorCount);
  run_issue(prop.multiProcessorCount);
  run_pair(prop.multiProcessorCount);
  run_ts(prop.multiProcessorCount);
  run_pace<64>(prop.multiProcessorCount);
  run_pace<128>(prop.multiProcessorCount);
  run_pace<256>(prop.multiProcessorCount);
  return 0;
}
