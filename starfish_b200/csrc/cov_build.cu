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
// Tiling: a CTA owns a 128×128 tile.  The tile's row/column slices of wave, σ, X and Y = A·X are staged
// in shared memory once (wave/X rows by 1-D bulk async copies when alignment allows); every thread then
// keeps its 4 columns' λ and Y in registers and sweeps 16 rows, writing 2×16 B per row so a warp store
// covers 512 contiguous bytes.  Matérn/Hann transcendentals are evaluated only in tiles that intersect
// the band r<=r0 (wave is checked once for monotonicity; unsorted grids fall back to per-element tests),
// and local kernels only in tiles whose rows AND columns intersect the block m<=4σ.
#include "sfb_internal.cuh"

namespace sfb {

namespace {

constexpr int BT = 128;        // tile edge
constexpr int NTHREADS = 256;  // 8 warps
constexpr double kPi = 3.141592653589793;  // == numpy.pi

struct LocalK {
  double amp, mu, sigma;
};

template <int MT>
__global__ void __launch_bounds__(NTHREADS) cov_build_kernel(BuildParams p) {
  const int tj = blockIdx.x, ti = blockIdx.y, b = blockIdx.z;
  if (p.lower_only && tj > ti) return;
  const int i0 = ti * BT, j0 = tj * BT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N, M = p.M;

  extern __shared__ __align__(16) double smem[];
  double* wr = smem;                 // [BT] row wavelengths
  double* wc = wr + BT;              // [BT] col wavelengths
  double* s2 = wc + BT;              // [BT] σ² of rows (diagonal tiles only)
  double* Xr = s2 + BT;              // [MT][BT] X rows  (X[m][i0+r])
  double* Yc = Xr + MT * BT;         // [MT][BT] Y cols  (Σ_m' A[m][m']·X[m'][j0+c])
  double* Am = Yc + MT * BT;         // [MT*MT]
  double* rowm = Am + MT * MT;       // [Kmax][BT] local metric of rows
  double* colm = rowm + p.Kmax * BT; // [Kmax][BT]
  __shared__ LocalK lk[kMaxK];
  __shared__ int lk_active[kMaxK];

  const int hb = b * p.hyper_stride;
  const double g_amp = p.glob ? p.glob[2 * hb] : 0.0;
  const double g_ls = p.glob ? p.glob[2 * hb + 1] : 1.0;
  const int nloc = p.nloc ? min(p.nloc[hb], p.Kmax) : 0;

  // ---- stage the tile's slices -----------------------------------------------------------------
  for (int t = tid; t < BT; t += NTHREADS) {
    int i = i0 + t, j = j0 + t;
    wr[t] = (i < N) ? p.wave[i] : 0.0;
    wc[t] = (j < N) ? p.wave[j] : 0.0;
    double s = (i < N) ? p.sigma[i] : 0.0;
    s2[t] = s * s;
  }
  if (MT > 0) {
    const double* Xb = p.X + (long long)b * M * N;
    for (int t = tid; t < M * BT; t += NTHREADS) {
      int m = t / BT, r = t % BT;
      Xr[m * BT + r] = (i0 + r < N) ? Xb[(long long)m * N + i0 + r] : 0.0;
      Yc[m * BT + r] = (j0 + r < N) ? Xb[(long long)m * N + j0 + r] : 0.0;  // X cols for now
    }
    for (int t = tid; t < M * M; t += NTHREADS) Am[t] = p.A[(long long)b * M * M + t];
  }
  if (tid < nloc) {
    const double* l = p.loc + ((long long)hb * p.Kmax + tid) * 3;
    lk[tid].amp = l[0];
    lk[tid].mu = l[1];
    lk[tid].sigma = l[2];
    lk_active[tid] = 0;
  }
  __syncthreads();

  // Y = A·Xc, in place through registers (each thread owns whole columns so no hazard)
  if (MT > 0) {
    for (int c = tid; c < BT; c += NTHREADS) {
      double xc[MT > 0 ? MT : 1], yc[MT > 0 ? MT : 1];
#pragma unroll
      for (int m = 0; m < MT; ++m) xc[m] = (m < M) ? Yc[m * BT + c] : 0.0;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < MT; ++q)
          if (m < M && q < M) acc = fma(Am[m * M + q], xc[q], acc);
        yc[m] = acc;
      }
#pragma unroll
      for (int m = 0; m < MT; ++m)
        if (m < M) Yc[m * BT + c] = yc[m];
    }
  }
  // local-kernel metrics of the tile's rows and columns; a kernel is active in this tile only when
  // some row AND some column lie inside its block m <= 4σ
  for (int k = 0; k < nloc; ++k) {
    const double mu = lk[k].mu, r0 = 4 * lk[k].sigma;
    const double f = kC_KMS / mu;
    int any_r = 0, any_c = 0;
    for (int t = tid; t < BT; t += NTHREADS) {
      double mr = f * fabs(wr[t] - mu), mc = f * fabs(wc[t] - mu);
      bool in_r = (i0 + t < N) && (mr <= r0), in_c = (j0 + t < N) && (mc <= r0);
      rowm[k * BT + t] = in_r ? mr : -1.0;  // -1 marks "outside"
      colm[k * BT + t] = in_c ? mc : -1.0;
      any_r |= in_r;
      any_c |= in_c;
    }
    any_r = __syncthreads_or(any_r);
    any_c = __syncthreads_or(any_c);
    if (tid == 0) lk_active[k] = any_r && any_c;
  }
  __syncthreads();

  // ---- does the tile intersect the Matérn band? ------------------------------------------------
  const double r0g = 6 * g_ls;
  bool band = g_amp > 0.0;
  if (band && *p.sorted && i0 != j0) {
    // sorted ascending: the closest pair is (last row, first col) for tiles right of the diagonal and
    // (first row, last col) for tiles below it
    int ilo, jhi;
    if (j0 > i0) { ilo = min(i0 + BT, N) - 1; jhi = j0; } else { ilo = i0; jhi = min(j0 + BT, N) - 1; }
    if (ilo < N && jhi < N && ilo >= 0) {
      double a = p.wave[ilo], c = p.wave[jhi];
      double rmin = kC_KMS / 2 * fabs((c - a) / (c + a));
      band = rmin <= r0g;
    }
  }
  int n_active = 0;
  for (int k = 0; k < nloc; ++k) n_active += lk_active[k];
  const bool diag_tile = (i0 == j0);
  const double sqrt3 = sqrt(3.0);

  // ---- per-thread column data ------------------------------------------------------------------
  const int cc[4] = {2 * lane, 2 * lane + 1, 64 + 2 * lane, 65 + 2 * lane};
  double wj[4];
  double yj[MT > 0 ? MT : 1][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    wj[q] = wc[cc[q]];
#pragma unroll
    for (int m = 0; m < MT; ++m) yj[m][q] = (m < M) ? Yc[m * BT + cc[q]] : 0.0;
  }
  double* Cb = p.C + (long long)b * p.strideC;
  const int padN = p.padN;

#pragma unroll 1
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
        if (cc[q] == r) v[q] += s2[r];
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
          const double mj = colm[k * BT + cc[q]];
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
        if (cc[q] == r) v[q] += p.jitter;
    }
    if (i >= N) {  // identity padding of the factorisation workspace
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = (j0 + cc[q] == i) ? 1.0 : 0.0;
    }
    double* row = Cb + (long long)i * p.ldc + j0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = cc[2 * h];
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

__global__ void check_sorted_kernel(const double* wave, int N, int* flag) {
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < N; i += gridDim.x * blockDim.x)
    bad |= !(wave[i + 1] > wave[i]) || !(wave[i] > 0.0);
  if (bad) atomicAnd(flag, 0);
}

__global__ void set_flag_kernel(int* flag, int v) { *flag = v; }

template <int MT>
cudaError_t launch_build_t(const BuildParams& p, int B, cudaStream_t st) {
  size_t smem = sizeof(double) * (3 * BT + 2 * MT * BT + MT * MT + 2 * (size_t)p.Kmax * BT);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cov_build_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int nt = (p.padN + BT - 1) / BT;
  dim3 grid(nt, nt, B);
  cov_build_kernel<MT><<<grid, NTHREADS, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_cov_build(const BuildParams& p, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  if (p.M == 0 || p.X == nullptr) {
    BuildParams q = p;
    q.M = 0;
    q.X = nullptr;
    return launch_build_t<0>(q, B, st);
  }
  if (p.M <= 8) return launch_build_t<8>(p, B, st);
  return launch_build_t<kMaxM>(p, B, st);
}

cudaError_t launch_check_sorted(const double* wave, int N, int* flag, cudaStream_t st) {
  set_flag_kernel<<<1, 1, 0, st>>>(flag, 1);
  check_sorted_kernel<<<32, 256, 0, st>>>(wave, N, flag);
  return cudaGetLastError();
}

}  // namespace sfb
