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
//
// Thread layout inside a 128×128 tile: a warp owns 64 consecutive columns (2 per lane, one 16-byte store per
// lane = 512 contiguous bytes per warp store) and every fourth row, so a thread carries only 2·M column
// factors in registers.  That keeps the kernel at <= 64 registers and 4 CTAs (32 warps) per SM: the
// transcendental-heavy tiles on the band (two per tile row, but two thirds of the instructions) of one CTA
// then overlap the store streams of the other three.  One __syncthreads per tile: Y = A·X of the columns and
// the local-kernel column metrics are double-buffered like the column slices themselves.
template <int MT>
__global__ void __launch_bounds__(NTHREADS, (MT <= 8 ? 4 : 2)) cov_build_kernel(BuildParams p) {
  constexpr int MA = (MT > 0) ? MT : 1;            // array extent
  constexpr int MP = (MT > 0) ? ((MT + 1) & ~1) : 2;  // row stride of Xr: even, so rows are 16-byte aligned
  const int nt = (p.padN + BT - 1) / BT;
  const int ti = nt - 1 - (int)blockIdx.x;  // longest rows first (lower-only mode: row ti has ti+1 tiles)
  const int b = blockIdx.y;
  const int i0 = ti * BT;
  const int ntj = p.lower_only ? ti + 1 : nt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = warp & 1, rph = warp >> 1;  // column half of the tile, row phase (rows rph, rph+4, ...)
  const int c0 = 64 * half + 2 * lane;         // this thread's columns c0, c0+1
  const int N = p.N;
  const int M = (MT <= 8) ? MT : p.M;  // compile-time constant on the common path
  const int Kmax = p.Kmax;

  extern __shared__ __align__(16) uint8_t smem_raw[];
  ColBuf<MT>* cbuf = reinterpret_cast<ColBuf<MT>*>(smem_raw);          // [2]
  double* wr = reinterpret_cast<double*>(cbuf + 2);                      // [BT] row wavelengths
  double* s2 = wr + BT;                                                  // [BT] σ² of rows
  double* Xr = s2 + BT;                                                  // [BT][MP] X of the rows, row-major
  double* Yc = Xr + BT * MP;                                             // [2][MA][BT] Y = A·X of the columns
  double* Am = Yc + 2 * MA * BT;                                         // [MA*MA]
  double* rowm = Am + MA * MA;                                           // [Kmax][BT] local metric of rows
  double* colm = rowm + Kmax * BT;                                       // [2][Kmax][BT] ... of the columns
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ LocalK lk[kMaxK];
  __shared__ int lk_row_any[kMaxK];

  const int hb = b * p.hyper_stride;
  const double g_amp = p.glob ? p.glob[2 * hb] : 0.0;
  const double g_ls = p.glob ? p.glob[2 * hb + 1] : 1.0;
  const int nloc = p.nloc ? min(p.nloc[hb], Kmax) : 0;
  const double* Xb = (MT > 0) ? p.X + (long long)b * M * N : nullptr;
  // bulk copies need 16-byte aligned sources: every row offset (m·N + j0)·8 is, iff N is even and the
  // base pointers are 16-byte aligned (checked on the host and passed in p.bulk_ok)
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
    for (int t = tid; t < MP * BT; t += NTHREADS) {
      const int m = t / BT, r = t % BT;  // coalesced reads along the row of X, transposed into [r][m]
      Xr[r * MP + m] = (m < M && i0 + r < N) ? Xb[(long long)m * N + i0 + r] : 0.0;
    }
    for (int t = tid; t < M * M; t += NTHREADS) Am[t] = p.A[(long long)b * M * M + t];
  }
  if (tid < nloc) {
    const double* l = p.loc + ((long long)hb * Kmax + tid) * 3;
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
  double* Cb = p.C + (long long)b * p.strideC;
  const int padN = p.padN;
  const int nrow = min(BT, N - i0);  // rows of this tile row inside the N×N block (may be <= 0)

  // ---- sweep over the column tiles ---------------------------------------------------------------
  for (int tj = 0; tj < ntj; ++tj) {
    const int st = tj & 1;
    const int j0 = tj * BT;
    const int ncol = min(BT, N - j0);
    ColBuf<MT>& cb = cbuf[st];
    double* Ycs = Yc + st * MA * BT;
    double* colms = colm + st * Kmax * BT;
    if (ncol > 0 && bulk_ok) mbar_wait(bar0 + 8 * st, (tj >> 1) & 1);
    if (!bulk_ok) __syncthreads();  // plain-load fallback: make the stage visible

    // Y = A·Xc for this tile's columns, local-kernel column metrics
    if (MT > 0) {
      for (int c = tid; c < BT; c += NTHREADS) {
        double xc[MA];
#pragma unroll
        for (int m = 0; m < MT; ++m) xc[m] = (m < M && c < ncol) ? cb.Xc[m * BT + c] : 0.0;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < MT; ++q)
            if (m < M && q < M) acc = fma(Am[m * M + q], xc[q], acc);
          if (m < M) Ycs[m * BT + c] = acc;
        }
      }
    }
    unsigned active = 0;  // bit k: local kernel k intersects this tile (uniform across the CTA)
    for (int k = 0; k < nloc; ++k) {
      if (lk_row_any[k]) {  // uniform
        int any_c = 0;
        const double mu = lk[k].mu, r0 = 4 * lk[k].sigma, f = kC_KMS / mu;
        for (int t = tid; t < BT; t += NTHREADS) {
          const double mc = (t < ncol) ? f * fabs(cb.wc[t] - mu) : 0.0;
          const bool in_c = (t < ncol) && (mc <= r0);
          colms[k * BT + t] = in_c ? mc : -1.0;
          any_c |= in_c;
        }
        if (__syncthreads_or(any_c)) active |= 1u << k;
      }
    }
    // Ycs / colms of this tile visible; every thread has left the row loop of tile tj-1, so the other stage of
    // the column buffer is free: the next tile's slices are fetched while this one is computed and stored
    __syncthreads();
    if (tj + 1 < ntj) prefetch(tj + 1, st ^ 1);

    // does the tile intersect the Matérn band?
    bool band = g_amp > 0.0;
    if (band && sorted && i0 != j0 && nrow > 0 && ncol > 0) {
      // sorted ascending: the closest pair is (last row, first col) for tiles right of the diagonal and
      // (first row, last col) for tiles below it — both already staged in shared memory
      const double a = (j0 > i0) ? wr[nrow - 1] : wr[0];
      const double c = (j0 > i0) ? cb.wc[0] : cb.wc[ncol - 1];
      band = (kC_KMS / 2 * fabs((c - a) / (c + a))) <= r0g;
    }
    const bool diag_tile = (i0 == j0);
    const bool interior = p.vec2 && (i0 + BT <= N) && (j0 + BT <= N);

    // Fast path — the overwhelming majority of tiles: fully inside the N×N block, away from the diagonal,
    // outside the Matérn band and every local block.  Only the rank-M term is left: M/2 broadcast 16-byte
    // shared loads, 2·M DFMA and one 16-byte store per row, with the row pointer carried as an induction variable.
    const bool fast = interior && !band && !active && !diag_tile;
    if (fast) {
      double y0[MA], y1[MA];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const double2 y = *reinterpret_cast<const double2*>(Ycs + m * BT + c0);
        y0[m] = (m < M) ? y.x : 0.0;
        y1[m] = (m < M) ? y.y : 0.0;
      }
      double* rowp = Cb + (long long)(i0 + rph) * p.ldc + j0 + c0;
      const long long step = 4 * p.ldc;
      const double* xr = Xr + rph * MP;
#pragma unroll 4
      for (int rr = 0; rr < BT / 4; ++rr) {
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int m2 = 0; m2 < MP / 2; ++m2) {
          if (MT > 0 && 2 * m2 < M) {
            const double2 x = *reinterpret_cast<const double2*>(xr + 2 * m2);
            v0 = fma(x.x, y0[2 * m2], v0);
            v1 = fma(x.x, y1[2 * m2], v1);
            if (2 * m2 + 1 < MT && 2 * m2 + 1 < M) {
              v0 = fma(x.y, y0[2 * m2 + 1], v0);
              v1 = fma(x.y, y1[2 * m2 + 1], v1);
            }
          }
        }
        *reinterpret_cast<double2*>(rowp) = make_double2(v0, v1);
        rowp += step;
        xr += 4 * MP;
      }
    } else {
      const double wj0 = (c0 < ncol) ? cb.wc[c0] : 0.0;
      const double wj1 = (c0 + 1 < ncol) ? cb.wc[c0 + 1] : 0.0;
#pragma unroll 1
      for (int rr = 0; rr < BT / 4; ++rr) {
        const int r = rph + 4 * rr;
        const int i = i0 + r;
        if (i >= padN) break;
        const double wi = wr[r];
        double v[2] = {0.0, 0.0};
        if (MT > 0) {
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (m < M) {
              const double x = Xr[r * MP + m];
              const double2 y = *reinterpret_cast<const double2*>(Ycs + m * BT + c0);
              v[0] = fma(x, y.x, v[0]);
              v[1] = fma(x, y.y, v[1]);
            }
          }
        }
        if (diag_tile) {
          if (c0 == r) v[0] += s2[r];
          if (c0 + 1 == r) v[1] += s2[r];
        }
        // Lower-only workspace: nothing reads the strict upper triangle of a diagonal tile (potrf_diag_kernel
        // loads c <= r only), so a warp whose 64 columns all lie right of row r skips the kernel functions
        const bool kernels_on = !(p.lower_only && diag_tile && 64 * half > r);  // warp-uniform
        if (band && kernels_on) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const double wj = q ? wj1 : wj0;
            const double rv = kC_KMS / 2 * fabs((wj - wi) / (wj + wi));
            if (rv <= r0g) {
              const double taper = 0.5 + 0.5 * cos(kPi * rv / r0g);
              v[q] += taper * g_amp * (1 + sqrt3 * rv / g_ls) * exp(-sqrt3 * rv / g_ls);
            }
          }
        }
        if (active && kernels_on) {
          double lsum[2] = {0.0, 0.0};
          for (int k = 0; k < nloc; ++k) {
            if (!((active >> k) & 1u)) continue;
            const double mi = rowm[k * BT + r];
            if (mi < 0.0) continue;
            const double r0 = 4 * lk[k].sigma, sg2 = lk[k].sigma * lk[k].sigma, amp = lk[k].amp;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const double mj = colms[k * BT + c0 + q];
              if (mj >= 0.0) {
                const double rt = fmax(mi, mj);
                const double taper = 0.5 + 0.5 * cos(kPi * rt / r0);
                lsum[q] += taper * amp * exp(-0.5 * (mi * mi + mj * mj) / sg2);
              }
            }
          }
          v[0] += lsum[0];
          v[1] += lsum[1];
        }
        if (diag_tile) {
          if (c0 == r) v[0] += p.jitter;
          if (c0 + 1 == r) v[1] += p.jitter;
        }
        if (i >= N) {  // identity padding of the factorisation workspace
          v[0] = (j0 + c0 == i) ? 1.0 : 0.0;
          v[1] = (j0 + c0 + 1 == i) ? 1.0 : 0.0;
        }
        double* row = Cb + (long long)i * p.ldc + j0;
        if (interior) {  // whole tile inside the N×N block and 16-byte aligned rows: one unconditional 16 B store
          *reinterpret_cast<double2*>(row + c0) = make_double2(v[0], v[1]);
        } else {
          const int j = j0 + c0;
          double a = v[0], c1 = v[1];
          if (i < N) {  // columns beyond N inside a padded workspace are zero
            if (j >= N) a = 0.0;
            if (j + 1 >= N) c1 = 0.0;
          }
          if (j + 1 < padN && p.vec2) {
            *reinterpret_cast<double2*>(row + c0) = make_double2(a, c1);
          } else {
            if (j < padN) row[c0] = a;
            if (j + 1 < padN) row[c0 + 1] = c1;
          }
        }
      }
    }
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
  constexpr int MA = (MT > 0) ? MT : 1, MP = (MT > 0) ? ((MT + 1) & ~1) : 2;
  const size_t smem = 2 * sizeof(ColBuf<MT>) +
                      sizeof(double) * (2 * BT + MP * BT + 2 * MA * BT + MA * MA + 3 * (size_t)p.Kmax * BT);
  cudaError_t e =
      cudaFuncSetAttribute(cov_build_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  // 4 CTAs of ~40 KB per SM: ask for the shared-memory-heavy L1 split (the kernel's global traffic is stores)
  e = cudaFuncSetAttribute(cov_build_kernel<MT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
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
