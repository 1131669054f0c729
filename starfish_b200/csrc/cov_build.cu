// Fused residual-covariance build for B walkers (sm_100a).
//
//   C[b] = X[b]ᵀ·A[b]·X[b] + diag(σ² + jitter) + K_global(a_b, ℓ_b) + Σ_k K_local(A_bk, μ_bk, σ_bk)
//
// One pass, one HBM write per element (8·N² B per walker; 4·N² in lower-only mode) — replaces the
// reference's ~20 dense N² numpy temporaries:
//   K_global  Starfish/models/kernels.py:27-40   r=(c/2)|λj−λi|/(λj+λi), r0=6ℓ, Matérn-3/2 × Hann
//   K_local   Starfish/models/kernels.py:70-80   m=(c/μ)|λ−μ|, r0=4σ, Gaussian × Hann on max(m_i,m_j)
//   XᵀAX, σ²  Starfish/models/spectrum_model.py:334-338;  sums :347-363;  jitter :399
//
// Work decomposition: a CTA owns one ROW of 128×128 tiles of one walker and sweeps the column tiles it has
// to produce (0..row for the lower-triangular factorisation workspace, all of them for a user matrix).
// The row slices of λ, σ², X and the local-kernel metrics are staged once; the column slices (λ and the M
// rows of X, 7 KB per tile) are prefetched one tile ahead into a double buffer with 1-D bulk async copies
// (cp.async.bulk → mbarrier, SASS UBLKCP) so global-load latency never sits between two tiles' stores.
// Per tile every thread keeps its 4 columns' λ and Y = A·X in registers and sweeps 16 rows, writing 2×16 B
// per row: a warp store covers 512 contiguous bytes.  Matérn/Hann transcendentals are evaluated only in
// tiles that intersect the band r<=r0 (wave is checked once for monotonicity; unsorted grids fall back to
// per-element tests), local kernels only in tiles whose rows AND columns intersect the block m<=4σ.
// The kernel is HBM-write-bound: 128 KB out per tile against ≈0.1 MFLOP of fp64.
#include <cstdint>

#include "sfb_internal.cuh"

namespace sfb {

namespace {

constexpr int BT = 128;        // tile edge
constexpr int NTHREADS = 256;  // 8 warps
constexpr double kPi = 3.141592653589793;  // == numpy.pi

struct LocalK {
  double amp, mu, sigma;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int MT>
struct ColBuf {  // one stage of the column double buffer
  double wc[BT];
  double Xc[(MT > 0 ? MT : 1) * BT];
};

// MT = register/unroll width of the rank-M term; for MT <= 8 the kernel is instantiated with MT == M exactly
// (no predication in the hot loop), wider models use MT = 12 or 16 with the tail predicated.
template <int MT>
__global__ void __launch_bounds__(NTHREADS) cov_build_kernel(BuildParams p) {
  const int nt = (p.padN + BT - 1) / BT;
  const int ti = nt - 1 - (int)blockIdx.x;  // longest rows first (lower-only mode: row ti has ti+1 tiles)
  const int b = blockIdx.y;
  const int i0 = ti * BT;
  const int ntj = p.lower_only ? ti + 1 : nt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  const int M = (MT <= 8) ? MT : p.M;  // compile-time constant on the common path

  extern __shared__ __align__(16) uint8_t smem_raw[];
  ColBuf<MT>* cbuf = reinterpret_cast<ColBuf<MT>*>(smem_raw);          // [2]
  double* wr = reinterpret_cast<double*>(cbuf + 2);                      // [BT] row wavelengths
  double* s2 = wr + BT;                                                  // [BT] σ² of rows
  double* Xr = s2 + BT;                                                  // [MT][BT] X rows
  double* Yc = Xr + MT * BT;                                             // [MT][BT] Y = A·X of the current columns
  double* Am = Yc + MT * BT;                                             // [MT*MT]
  double* rowm = Am + MT * MT;                                           // [Kmax][BT] local metric of rows
  double* colm = rowm + p.Kmax * BT;                                     // [Kmax][BT] ... of current columns
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ LocalK lk[kMaxK];
  __shared__ int lk_row_any[kMaxK];
  __shared__ int lk_active[kMaxK];

  const int hb = b * p.hyper_stride;
  const double g_amp = p.glob ? p.glob[2 * hb] : 0.0;
  const double g_ls = p.glob ? p.glob[2 * hb + 1] : 1.0;
  const int nloc = p.nloc ? min(p.nloc[hb], p.Kmax) : 0;
  const double* Xb = (MT > 0) ? p.X + (long long)b * M * N : nullptr;
  // bulk copies need 16-byte aligned sources: every row offset (m·N + j0)·8 is, iff N is even and the
  // base pointers are 16-byte aligned (checked on the host and passed in p.vec2-like flag bulk_ok)
  const bool bulk_ok = p.bulk_ok != 0;
  const uint32_t bar0 = smem_u32(&bars[0]);

  // ---- prefetch of one column tile's slices into stage `st`
  auto prefetch = [&](int tj, int st) {
    const int j0 = tj * BT;
    const int ncol = min(BT, N - j0);  // may be <= 0 for pure padding tiles
    ColBuf<MT>& cb = cbuf[st];
    if (ncol <= 0) return;
    if (bulk_ok) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)ncol * 8u;
        const uint32_t bar = bar0 + 8 * st;
        mbar_arrive_expect_tx(bar, bytes * (1 + (MT > 0 ? M : 0)));
        bulk_g2s(smem_u32(cb.wc), p.wave + j0, bytes, bar);
        if (MT > 0)
          for (int m = 0; m < M; ++m) bulk_g2s(smem_u32(cb.Xc + m * BT), Xb + (long long)m * N + j0, bytes, bar);
      }
    } else {
      for (int t = tid; t < ncol; t += NTHREADS) cb.wc[t] = p.wave[j0 + t];
      if (MT > 0)
        for (int t = tid; t < M * BT; t += NTHREADS) {
          const int m = t / BT, c = t % BT;
          if (c < ncol) cb.Xc[m * BT + c] = Xb[(long long)m * N + j0 + c];
        }
    }
  };

  // ---- one-off staging of the row slices -------------------------------------------------------
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = tid; t < BT; t += NTHREADS) {
    const int i = i0 + t;
    wr[t] = (i < N) ? p.wave[i] : 0.0;
    const double s = (i < N) ? p.sigma[i] : 0.0;
    s2[t] = s * s;
  }
  if (MT > 0) {
    for (int t = tid; t < M * BT; t += NTHREADS) {
      const int m = t / BT, r = t % BT;
      Xr[m * BT + r] = (i0 + r < N) ? Xb[(long long)m * N + i0 + r] : 0.0;
    }
    for (int t = tid; t < M * M; t += NTHREADS) Am[t] = p.A[(long long)b * M * M + t];
  }
  if (tid < nloc) {
    const double* l = p.loc + ((long long)hb * p.Kmax + tid) * 3;
    lk[tid].amp = l[0];
    lk[tid].mu = l[1];
    lk[tid].sigma = l[2];
  }
  __syncthreads();  // barriers initialised, lk visible
  prefetch(0, 0);
  for (int k = 0; k < nloc; ++k) {
    const double mu = lk[k].mu, r0 = 4 * lk[k].sigma, f = kC_KMS / mu;
    int any_r = 0;
    for (int t = tid; t < BT; t += NTHREADS) {
      const double mr = f * fabs(wr[t] - mu);
      const bool in_r = (i0 + t < N) && (mr <= r0);
      rowm[k * BT + t] = in_r ? mr : -1.0;  // -1 marks "outside"
      any_r |= in_r;
    }
    any_r = __syncthreads_or(any_r);
    if (tid == 0) lk_row_any[k] = any_r;
  }
  __syncthreads();

  const double r0g = 6 * g_ls;
  const double sqrt3 = sqrt(3.0);
  const bool sorted = (*p.sorted != 0);
  const int cA = 2 * lane, cB = 64 + 2 * lane;  // this thread's column pairs (cA, cA+1) and (cB, cB+1)
  auto col_of = [&](int q) { return (q < 2 ? cA : cB) + (q & 1); };
  double* Cb = p.C + (long long)b * p.strideC;
  const int padN = p.padN;

  // ---- sweep over the column tiles ---------------------------------------------------------------
  for (int tj = 0; tj < ntj; ++tj) {
    const int st = tj & 1;
    const int j0 = tj * BT;
    const int ncol = min(BT, N - j0);
    ColBuf<MT>& cb = cbuf[st];
    if (tj + 1 < ntj) prefetch(tj + 1, st ^ 1);  // stage st^1 was released by the barrier ending tile tj-1
    if (ncol > 0 && bulk_ok) mbar_wait(bar0 + 8 * st, (tj >> 1) & 1);
    if (!bulk_ok) __syncthreads();               // plain-load fallback: make the stage visible

    // Y = A·Xc for this tile's columns, local-kernel column metrics
    if (MT > 0) {
      for (int c = tid; c < BT; c += NTHREADS) {
        double xc[MT > 0 ? MT : 1];
#pragma unroll
        for (int m = 0; m < MT; ++m) xc[m] = (m < M && c < ncol) ? cb.Xc[m * BT + c] : 0.0;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < MT; ++q)
            if (m < M && q < M) acc = fma(Am[m * M + q], xc[q], acc);
          if (m < M) Yc[m * BT + c] = acc;
        }
      }
    }
    int n_active = 0;
    for (int k = 0; k < nloc; ++k) {
      int any_c = 0;
      if (lk_row_any[k]) {  // uniform
        const double mu = lk[k].mu, r0 = 4 * lk[k].sigma, f = kC_KMS / mu;
        for (int t = tid; t < BT; t += NTHREADS) {
          const double mc = (t < ncol) ? f * fabs(cb.wc[t] - mu) : 0.0;
          const bool in_c = (t < ncol) && (mc <= r0);
          colm[k * BT + t] = in_c ? mc : -1.0;
          any_c |= in_c;
        }
        any_c = __syncthreads_or(any_c);
      }
      if (tid == 0) lk_active[k] = any_c;
      n_active += any_c;  // uniform: __syncthreads_or returns the same value to every thread
    }
    __syncthreads();  // Yc, colm, lk_active visible

    // does the tile intersect the Matérn band?
    bool band = g_amp > 0.0;
    if (band && sorted && i0 != j0) {
      // sorted ascending: the closest pair is (last row, first col) for tiles right of the diagonal and
      // (first row, last col) for tiles below it
      int ilo, jhi;
      if (j0 > i0) { ilo = min(i0 + BT, N) - 1; jhi = j0; } else { ilo = i0; jhi = min(j0 + BT, N) - 1; }
      if (ilo < N && jhi < N && ilo >= 0 && jhi >= 0) {
        const double a = p.wave[ilo], c = p.wave[jhi];
        band = (kC_KMS / 2 * fabs((c - a) / (c + a))) <= r0g;
      }
    }
    const bool diag_tile = (i0 == j0);
    const bool interior = p.vec2 && (i0 + BT <= N) && (j0 + BT <= N);

    double wj[4];
    double yj[MT > 0 ? MT : 1][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      wj[q] = (col_of(q) < ncol) ? cb.wc[col_of(q)] : 0.0;
#pragma unroll
      for (int m = 0; m < MT; ++m) yj[m][q] = (m < M) ? Yc[m * BT + col_of(q)] : 0.0;
    }

    // Fast path — the overwhelming majority of tiles: fully inside the N×N block, away from the diagonal,
    // outside the Matérn band and every local block.  Only the rank-M term is left: M shared loads, 4·M
    // DFMA and two 16-byte stores per row, with the row pointer carried as an induction variable.
    const bool fast = interior && !band && !n_active && !diag_tile;
    if (fast) {
      double* rowp = Cb + (long long)(i0 + warp) * p.ldc + j0;
      const long long step = 8 * p.ldc;
#pragma unroll 4
      for (int rr = 0; rr < BT / 8; ++rr) {
        const int r = warp + 8 * rr;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m < M) {
            const double x = Xr[m * BT + r];
            v0 = fma(x, yj[m][0], v0);
            v1 = fma(x, yj[m][1], v1);
            v2 = fma(x, yj[m][2], v2);
            v3 = fma(x, yj[m][3], v3);
          }
        }
        *reinterpret_cast<double2*>(rowp + cA) = make_double2(v0, v1);
        *reinterpret_cast<double2*>(rowp + cB) = make_double2(v2, v3);
        rowp += step;
      }
    } else
#pragma unroll 2
    for (int rr = 0; rr < BT / 8; ++rr) {
      const int r = warp + 8 * rr;
      const int i = i0 + r;
      if (i >= padN) break;
      const double wi = wr[r];
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      if (MT > 0) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m < M) {
            const double x = Xr[m * BT + r];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = fma(x, yj[m][q], v[q]);
          }
        }
      }
      if (diag_tile) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (col_of(q) == r) v[q] += s2[r];
      }
      if (band) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double rv = kC_KMS / 2 * fabs((wj[q] - wi) / (wj[q] + wi));
          if (rv <= r0g) {
            const double taper = 0.5 + 0.5 * cos(kPi * rv / r0g);
            v[q] += taper * g_amp * (1 + sqrt3 * rv / g_ls) * exp(-sqrt3 * rv / g_ls);
          }
        }
      }
      if (n_active) {
        double lsum[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < nloc; ++k) {
          if (!lk_active[k]) continue;
          const double mi = rowm[k * BT + r];
          if (mi < 0.0) continue;
          const double r0 = 4 * lk[k].sigma, sg2 = lk[k].sigma * lk[k].sigma, amp = lk[k].amp;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double mj = colm[k * BT + col_of(q)];
            if (mj >= 0.0) {
              const double rt = fmax(mi, mj);
              const double taper = 0.5 + 0.5 * cos(kPi * rt / r0);
              lsum[q] += taper * amp * exp(-0.5 * (mi * mi + mj * mj) / sg2);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += lsum[q];
      }
      if (diag_tile) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (col_of(q) == r) v[q] += p.jitter;
      }
      if (i >= N) {  // identity padding of the factorisation workspace
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (j0 + col_of(q) == i) ? 1.0 : 0.0;
      }
      double* row = Cb + (long long)i * p.ldc + j0;
      if (interior) {  // whole tile inside the N×N block and 16-byte aligned rows: two unconditional 16 B stores
        *reinterpret_cast<double2*>(row + cA) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(row + cB) = make_double2(v[2], v[3]);
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = h ? cB : cA;
          const int j = j0 + c;
          double a = v[2 * h], c1 = v[2 * h + 1];
          if (i < N) {  // columns beyond N inside a padded workspace are zero
            if (j >= N) a = 0.0;
            if (j + 1 >= N) c1 = 0.0;
          }
          if (j + 1 < padN && p.vec2) {
            *reinterpret_cast<double2*>(row + c) = make_double2(a, c1);
          } else {
            if (j < padN) row[c] = a;
            if (j + 1 < padN) row[c + 1] = c1;
          }
        }
      }
    }
    __syncthreads();  // everyone is done with stage st, Yc and colm before they are overwritten
  }
}

__global__ void check_sorted_kernel(const double* wave, int N, int* flag) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < N; i += gridDim.x * blockDim.x)
    bad |= !(wave[i + 1] > wave[i]) || !(wave[i] > 0.0);
  if (bad) atomicAnd(flag, 0);
}

__global__ void set_flag_kernel(int* flag, int v) { *flag = v; }

template <int MT>
cudaError_t launch_build_t(const BuildParams& p, int B, cudaStream_t st) {
  const size_t smem = 2 * sizeof(ColBuf<MT>) +
                      sizeof(double) * (2 * BT + 2 * MT * BT + MT * MT + 2 * (size_t)p.Kmax * BT);
  cudaError_t e =
      cudaFuncSetAttribute(cov_build_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  const int nt = (p.padN + BT - 1) / BT;
  dim3 grid(nt, B);
  cov_build_kernel<MT><<<grid, NTHREADS, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_cov_build(const BuildParams& p, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  BuildParams q = p;
  // bulk (TMA) prefetch needs 16-byte aligned row starts: N even and 16-byte aligned base pointers
  q.bulk_ok = ((p.N % 2) == 0) && ((reinterpret_cast<uintptr_t>(p.wave) & 15) == 0) &&
              (p.X == nullptr || (reinterpret_cast<uintptr_t>(p.X) & 15) == 0);
  if (p.M == 0 || p.X == nullptr) {
    q.M = 0;
    q.X = nullptr;
    return launch_build_t<0>(q, B, st);
  }
  switch (p.M) {
    case 1: return launch_build_t<1>(q, B, st);
    case 2: return launch_build_t<2>(q, B, st);
    case 3: return launch_build_t<3>(q, B, st);
    case 4: return launch_build_t<4>(q, B, st);
    case 5: return launch_build_t<5>(q, B, st);
    case 6: return launch_build_t<6>(q, B, st);
    case 7: return launch_build_t<7>(q, B, st);
    case 8: return launch_build_t<8>(q, B, st);
    default: break;
  }
  if (p.M <= 12) return launch_build_t<12>(q, B, st);
  return launch_build_t<kMaxM>(q, B, st);
}

cudaError_t launch_check_sorted(const double* wave, int N, int* flag, cudaStream_t st) {
  set_flag_kernel<<<1, 1, 0, st>>>(flag, 1);
  check_sorted_kernel<<<32, 256, 0, st>>>(wave, N, flag);
  return cudaGetLastError();
}

}  // namespace sfb
