// Structure-exploiting solver: the same log-likelihood as the dense path, computed from
//     C = S + XᵀAX,   S = diag(σ² + jitter) + K_global + Σ K_local
// On a strictly increasing wavelength grid S is banded (half-width b ≈ 24ℓ/dv pixels; the local blocks lie
// inside the band), so
//     log det C = log det S + log det(I + A·G),          G = X S⁻¹ Xᵀ  (M×M)
//     RᵀC⁻¹R   = RᵀS⁻¹R − uᵀ (I + A·G)⁻¹ A u,            u = X S⁻¹ R
// needs one banded Cholesky S = LLᵀ (N·b² FLOP instead of N³/3) and the forward solves Z = L⁻¹[R | Xᵀ]:
// everything above is a Gram matrix of Z.  It replaces Starfish/models/spectrum_model.py:334-363 + :399-405
// exactly like the dense path; it is a separate mode with its own roofline (fp64 FMA issue of the window
// update), reported separately by bench.py.
//
//   band_build_kernel   S in band storage Sb[i][d] = S[i, i−d], d < WD — the support masks of cov_build_kernel
//                       bit for bit, values within 2 ulp, 8·N·WD bytes per walker instead of 8·N²
//   band_chol_kernel    ONE persistent CTA per walker.  The active window of the factorisation (rows/columns
//                       j..j+WD−1) lives in REGISTERS, addressed circularly (index mod WD), 8×(WD/32) elements
//                       per thread; per pivot: the owners publish column j through a double-buffered
//                       shared-memory column, one __syncthreads, every thread applies the rank-1 update to
//                       its registers, and the row that enters the window replaces the one that retires.  The
//                       M+1 right-hand sides ride along in registers, the Gram matrix ZᵀZ is accumulated on
//                       the fly, so L is never written anywhere: HBM traffic is one read of Sb and of X.
//                       Entering rows are staged 16 pivots ahead into a shared-memory ring with cp.async.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "sfb_internal.cuh"

namespace sfb {

namespace {

constexpr double kPi = 3.141592653589793;  // == numpy.pi
constexpr int NRP = kMaxM + 2;             // padded right-hand-side columns per ring row (even)

// ------------------------------------------------------------------------------------------------
// band build
// ------------------------------------------------------------------------------------------------
constexpr int BB_ROWS = 8;  // rows per CTA (one per warp)

__global__ void __launch_bounds__(BB_ROWS * 32)
band_build_kernel(BandBuildParams p) {
  __shared__ double l_amp[kMaxK], l_mu[kMaxK], l_sig[kMaxK], l_ir0[kMaxK], l_mh[kMaxK], l_f[kMaxK];
  const int b = p.rowmap ? p.rowmap[blockIdx.y] : blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hb = b * p.hyper_stride;
  const double g_amp = p.glob[2 * hb], g_ls = p.glob[2 * hb + 1];
  const int nloc = min(p.nloc[hb], p.Kmax);
  if (threadIdx.x < nloc) {
    const double* l = p.loc + ((long long)hb * p.Kmax + threadIdx.x) * 3;
    l_amp[threadIdx.x] = l[0];
    l_mu[threadIdx.x] = l[1];
    l_sig[threadIdx.x] = l[2];
    l_ir0[threadIdx.x] = 1.0 / (4 * l[2]);     // 1/r0
    l_mh[threadIdx.x] = -0.5 / (l[2] * l[2]);   // −1/(2σ²)
    l_f[threadIdx.x] = kC_KMS / l[1];            // c/μ, the factor of the metric m = (c/μ)|λ − μ|
  }
  __syncthreads();
  const int i = blockIdx.x * BB_ROWS + warp;
  if (i >= p.N) return;
  const int WD = p.WD;
  const double wi = p.wave[i];
  const double r0g = 6 * g_ls, sqrt3 = sqrt(3.0);
  // The support tests (r <= r0) use exactly the expressions of cov_build_kernel, so both solvers agree on
  // which entries exist; inside the support the divisions by per-walker constants become multiplications by
  // their reciprocals and cos(π·x) is evaluated as cospi(x) — values within 2 ulp of the dense build's, at
  // half the instruction count (this kernel is issue-bound on fp64 division / cos / exp sequences).
  const double inv_r0g = 1.0 / r0g, s3_ls = sqrt3 / g_ls;
  double* out = p.Sb + (long long)b * p.strideSb + (long long)i * WD;
  // local kernels whose block contains THIS row (m_i <= 4σ): almost always none, so the per-element loop over
  // the kernels (a metric and two compares each) runs only for the few rows inside a local block
  unsigned rowmask = 0;
  for (int q = 0; q < nloc; ++q)
    if (l_f[q] * fabs(wi - l_mu[q]) <= 4 * l_sig[q]) rowmask |= 1u << q;
  int over = 0;
  for (int d = lane; d <= WD; d += 32) {  // d == WD is only the "does the band fit" probe
    const int k = i - d;
    double v = 0.0;
    bool nz = false;
    if (k >= 0) {
      const double wk = p.wave[k];
      if (d == 0) {
        const double s = p.sigma[i];
        v = s * s;
      }
      if (g_amp > 0.0) {
        const double rv = kC_KMS / 2 * fabs((wk - wi) / (wk + wi));
        if (rv <= r0g) {
          const double taper = 0.5 + 0.5 * cospi(rv * inv_r0g);
          const double y = s3_ls * rv;
          v += taper * g_amp * (1 + y) * exp(-y);
          nz = true;
        }
      }
      double lsum = 0.0;
      bool any = false;
      for (unsigned rem = rowmask; rem; rem &= rem - 1) {
        const int q = __ffs(rem) - 1;
        const double mu = l_mu[q], r0 = 4 * l_sig[q], f = l_f[q];
        const double mi = f * fabs(wi - mu), mj = f * fabs(wk - mu);
        if (mi <= r0 && mj <= r0) {
          const double rt = fmax(mi, mj);
          const double taper = 0.5 + 0.5 * cospi(rt * l_ir0[q]);
          lsum += taper * l_amp[q] * exp((mi * mi + mj * mj) * l_mh[q]);
          any = true;
        }
      }
      if (any) {
        v += lsum;
        nz = true;
      }
      if (d == 0) v += p.jitter;
    }
    if (d < WD)
      out[d] = v;
    else if (nz)
      over = 1;
  }
  if (over) atomicOr(p.overflow + b, 1);
}

// ------------------------------------------------------------------------------------------------
// exact half-bandwidth of S per walker (strictly increasing grid): the largest i − k with S[i,k] != 0.
// Global kernel: its support test is monotone in k, so row i's first in-support column is found by bisection
// with the very same r <= r0 expression the build uses; local kernel: its support is the contiguous pixel
// block with m <= 4σ, whose width is a count.  One CTA per walker.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
band_width_kernel(int N, int Kmax, int hyper_stride, const double* __restrict__ wave,
                  const double* __restrict__ glob, const int* __restrict__ nloc, const double* __restrict__ loc,
                  int* __restrict__ bw) {
  __shared__ int s_max;
  __shared__ int s_cnt[kMaxK];
  const int b = blockIdx.x, hb = b * hyper_stride, tid = threadIdx.x;
  const double g_amp = glob[2 * hb], g_ls = glob[2 * hb + 1];
  const int nl = min(nloc[hb], Kmax);
  if (tid == 0) s_max = 0;
  if (tid < kMaxK) s_cnt[tid] = 0;
  __syncthreads();
  const double r0g = 6 * g_ls;
  int mx = 0;
  if (g_amp > 0.0) {
    for (int i = tid; i < N; i += blockDim.x) {
      const double wi = wave[i];
      int lo = 0, hi = i;  // smallest k in [0, i] with r(i,k) <= r0g (k = i always qualifies: r = 0)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double wk = wave[mid];
        if (kC_KMS / 2 * fabs((wk - wi) / (wk + wi)) <= r0g) hi = mid; else lo = mid + 1;
      }
      mx = max(mx, i - lo);
    }
  }
  for (int q = 0; q < nl; ++q) {
    const double* l = loc + ((long long)hb * Kmax + q) * 3;
    const double mu = l[1], r0 = 4 * l[2], f = kC_KMS / mu;
    int c = 0;
    for (int i = tid; i < N; i += blockDim.x) c += (f * fabs(wave[i] - mu) <= r0) ? 1 : 0;
    if (c) atomicAdd(&s_cnt[q], c);
  }
  atomicMax(&s_max, mx);
  __syncthreads();
  if (tid == 0) {
    int m = s_max;
    for (int q = 0; q < nl; ++q) m = max(m, s_cnt[q] - 1);
    bw[b] = m;
  }
}

// ------------------------------------------------------------------------------------------------
// banded Cholesky + forward solves + Gram matrix + capacitance epilogue
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Thread-0 epilogue shared by the band kernels: the M×M capacitance system on the Gram matrix of
// Z = L⁻¹[R | Xᵀ] (gram[p·NR+q], NR = M+1), lnL = −½ (log det S + log det(I + A·G) + RᵀS⁻¹R − uᵀ(I + A·G)⁻¹A u).
__device__ __noinline__ void band_epilogue(const BandCholParams& p, int b, int M, const double* gram, double logdet,
                                           int info) {
  const int N = p.N, NR = M + 1;
  {
    double quad = gram[0], ld = logdet;
    if (M > 0 && info == 0) {
      double Kc[kMaxM][kMaxM], v[kMaxM], u[kMaxM];
      const double* A = p.A + (long long)b * M * M;
      for (int i = 0; i < M; ++i) {
        u[i] = gram[(i + 1) * NR];
        for (int k = 0; k < M; ++k) {
          double s = (i == k) ? 1.0 : 0.0;
          for (int q = 0; q < M; ++q) s = fma(A[i * M + q], gram[(q + 1) * NR + (k + 1)], s);
          Kc[i][k] = s;
        }
      }
      for (int i = 0; i < M; ++i) {
        double s = 0.0;
        for (int q = 0; q < M; ++q) s = fma(A[i * M + q], u[q], s);
        v[i] = s;
      }
      // LU with partial pivoting on the M×M capacitance matrix, solving K y = v alongside
      double sign = 1.0, ldk = 0.0;
      for (int c = 0; c < M; ++c) {
        int piv = c;
        for (int r = c + 1; r < M; ++r)
          if (fabs(Kc[r][c]) > fabs(Kc[piv][c])) piv = r;
        if (piv != c) {
          for (int q = 0; q < M; ++q) { const double t = Kc[c][q]; Kc[c][q] = Kc[piv][q]; Kc[piv][q] = t; }
          const double t = v[c]; v[c] = v[piv]; v[piv] = t;
          sign = -sign;
        }
        const double d = Kc[c][c];
        if (d < 0.0) sign = -sign;
        ldk += log(fabs(d));
        for (int r = c + 1; r < M; ++r) {
          const double f = Kc[r][c] / d;
          for (int q = c + 1; q < M; ++q) Kc[r][q] = fma(-f, Kc[c][q], Kc[r][q]);
          v[r] = fma(-f, v[c], v[r]);
        }
      }
      for (int c = M - 1; c >= 0; --c) {
        double s = v[c];
        for (int q = c + 1; q < M; ++q) s = fma(-Kc[c][q], v[q], s);
        v[c] = s / Kc[c][c];
      }
      // Positive definiteness of C.  With S = LLᵀ positive definite, C is PD iff every eigenvalue of I + A·G is
      // positive (they are real: A·G is similar to the symmetric GcᵀA·Gc with G = Gc·Gcᵀ); the sign of the
      // determinant alone would miss an even number of negative ones.  Gc: semi-definite Cholesky of the Gram
      // matrix (a numerically zero pivot zeroes its column), then a plain Cholesky of T = I + GcᵀA·Gc decides.
      {
        double Gc[kMaxM][kMaxM], T[kMaxM][kMaxM];
        double gmax = 0.0;
        for (int i = 0; i < M; ++i) gmax = fmax(gmax, gram[(i + 1) * NR + (i + 1)]);
        for (int c = 0; c < M; ++c) {
          double dd = gram[(c + 1) * NR + (c + 1)];
          for (int q = 0; q < c; ++q) dd -= Gc[c][q] * Gc[c][q];
          const bool zero = !(dd > 1e-14 * gmax);
          const double piv = zero ? 0.0 : sqrt(dd);
          for (int r = 0; r < M; ++r) {
            if (r < c) { Gc[r][c] = 0.0; continue; }
            if (r == c) { Gc[r][c] = piv; continue; }
            double sacc = gram[(r + 1) * NR + (c + 1)];
            for (int q = 0; q < c; ++q) sacc -= Gc[r][q] * Gc[c][q];
            Gc[r][c] = zero ? 0.0 : sacc / piv;
          }
        }
        for (int i = 0; i < M; ++i)
          for (int k = 0; k < M; ++k) {  // T = I + GcᵀA·Gc
            double sacc = (i == k) ? 1.0 : 0.0;
            for (int r = 0; r < M; ++r) {
              double t = 0.0;
              for (int q = 0; q < M; ++q) t = fma(A[r * M + q], Gc[q][k], t);
              sacc = fma(Gc[r][i], t, sacc);
            }
            T[i][k] = sacc;
          }
        bool pd = true;
        for (int c = 0; c < M && pd; ++c) {
          double dd = 0.5 * (T[c][c] + T[c][c]);
          for (int q = 0; q < c; ++q) dd -= T[c][q] * T[c][q];
          if (!(dd > 0.0)) { pd = false; break; }
          dd = sqrt(dd);
          T[c][c] = dd;
          for (int r = c + 1; r < M; ++r) {
            double sacc = 0.5 * (T[r][c] + T[c][r]);
            for (int q = 0; q < c; ++q) sacc -= T[r][q] * T[c][q];
            T[r][c] = sacc / dd;
          }
        }
        if (!pd || !(sign > 0.0) || !(ldk == ldk)) info = N;  // C is not positive definite
      }
      ld += ldk;
      for (int i = 0; i < M; ++i) quad = fma(-u[i], v[i], quad);
    }
    p.info[b] = info;
    p.lnL[b] = info == 0 ? -(ld + quad) / 2 : nan("");
  }
}

__host__ __device__ constexpr int band_lcm(int WD, int ER) { return (ER % (WD / 32) == 0) ? ER : ER * (WD / 32); }
__host__ __device__ constexpr int band_batch(int WD, int ER) {
  const int l = band_lcm(WD, ER);
  return (l <= 24 && WD % l == 0) ? (l < 16 ? 2 * l : l) : 16;
}

// ER = window rows per thread (8 → WD·4 threads; the 160-pixel window uses 10 → 512 threads so that its
// 10×5 register tile fits a 128-register budget), MAXNR = right-hand sides the instantiation can carry (M + 1 <= MAXNR)
template <int WD, int ER, int MAXNR>
__global__ void __launch_bounds__(WD * 32 / ER, 1)
band_chol_kernel(BandCholParams p) {
  constexpr int EC = WD / 32, NT = WD * 32 / ER, ROWLEN = WD + NRP;
  constexpr int BATCH = band_batch(WD, ER);  // pivots per staging batch (a multiple of the unroll length)
  constexpr int NE = (WD * MAXNR + NT - 1) / NT;          // right-hand-side registers per thread
  __shared__ double colbuf[2][WD];
  __shared__ double zbuf[2][NRP];
  __shared__ double invbuf[2];
  __shared__ double gram[(kMaxM + 1) * (kMaxM + 1)];
  extern __shared__ double ring[];  // [2][BATCH][ROWLEN]

  const int b = p.rowmap ? p.rowmap[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, tr = tid >> 5;
  const int N = p.N, M = p.M, NR = M + 1;
  if (p.overflow[b] != 0 || *p.sorted == 0) {  // band wider than the window / grid not increasing: not ours
    if (tid == 0) {
      p.info[b] = p.overflow[b] != 0 ? -2 : -3;
      p.lnL[b] = nan("");
    }
    return;
  }
  const double* Sb = p.Sb + (long long)b * p.strideSb;
  const double* Xb = (M > 0) ? p.X + (long long)b * M * N : nullptr;
  const double* Fb = p.model_flux + (long long)b * N;

  auto rhs_at = [&](int i, int q) -> double {
    if (i >= N) return 0.0;
    return q == 0 ? Fb[i] - p.data_flux[i] : Xb[(long long)(q - 1) * N + i];
  };
  // fill one half of the ring with the rows that enter during the batch whose first pivot is j0 (rows
  // j0+WD ...): band rows and X columns by 8-byte cp.async (fire and forget), the residual column and the
  // identity padding past the end of the matrix by plain stores
  auto stage_issue = [&](int j0, double* dst) {
    for (int sidx = tid; sidx < BATCH * ROWLEN; sidx += NT) {
      const int rb = sidx / ROWLEN, col = sidx - rb * ROWLEN;
      const int i = j0 + WD + rb;
      if (col < WD) {
        if (i < N) cp_async8(dst + sidx, Sb + (long long)i * WD + col);
        else dst[sidx] = (col == 0) ? 1.0 : 0.0;
      } else {
        const int q = col - WD;
        if (q >= 1 && q < NR && i < N) cp_async8(dst + sidx, Xb + (long long)(q - 1) * N + i);
        else dst[sidx] = (q == 0) ? rhs_at(i, 0) : 0.0;
      }
    }
    cp_async_commit();
  };

  // ---- initial window: slot (r, c) = element (i = r, k = c), symmetric fill is not needed (k <= i only)
  double a[ER][EC];
#pragma unroll
  for (int er = 0; er < ER; ++er) {
    const int i = tr * ER + er;
#pragma unroll
    for (int ec = 0; ec < EC; ++ec) {
      const int k = lane * EC + ec;
      double v = 0.0;
      if (k <= i) v = (i < N) ? Sb[(long long)i * WD + (i - k)] : (k == i ? 1.0 : 0.0);
      a[er][ec] = v;
    }
  }
  double rv[NE];
  int rres[NE], rq[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int idx = tid + NT * e;
    const bool ok = idx < WD * NR;
    rres[e] = ok ? idx / NR : -1;
    rq[e] = ok ? idx - (idx / NR) * NR : 0;
    rv[e] = ok ? rhs_at(rres[e], rq[e]) : 0.0;
  }
  stage_issue(0, ring);

  double logdet = 0.0, gacc = 0.0;
  int info = 0;
  const int gp_ = tid / NR, gq_ = tid - (tid / NR) * NR;  // Gram element of this thread (tid < NR²)
  const bool gram_on = tid < NR * NR;

  // log det S = Σ log(pivot): the pivots are multiplied up as mantissa × 2^exponent by warp 0 (a handful of
  // integer/DMUL instructions per pivot instead of a log() on the critical path) and logged once at the end.
  double mant = 1.0;
  long long expo = 0;

  // The pivot loop is unrolled by UN = lcm(ER, EC) so that "which register holds column j / row j" is a
  // compile-time fact inside the body (j ≡ u mod UN, UN | WD): publishing the pivot column and replacing
  // the retiring row are plain register moves, no selects.
  constexpr int UN = band_lcm(WD, ER);                // ER, EC as used here: 8|{1,2,4}, 8·3, 10|5
  static_assert(UN <= 24 && WD % UN == 0 && UN % ER == 0 && UN % EC == 0 && UN % 2 == 0 && BATCH % UN == 0,
                "unroll length must divide the window and the staging batch");
  // 1/pivot is taken off the critical path: the owner of the NEXT diagonal element updates it first thing
  // after the barrier, starts its reciprocal and publishes it for the following pivot, so that nobody waits
  // for a division between the barrier and the FMAs.
  if (tid == 0) invbuf[0] = 1.0 / a[0][0];
  for (int j0 = 0, jr0 = 0; j0 < N; j0 += UN, jr0 = (jr0 + UN == WD) ? 0 : jr0 + UN) {
    const int own_lane0 = jr0 / EC, own_warp0 = jr0 / ER;
    const int jb0 = j0 % BATCH, half = (j0 / BATCH) & 1;
    const double* rowbase = ring + half * (BATCH * ROWLEN) + jb0 * ROWLEN;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u, jr = jr0 + u;
      if (j >= N) break;
      const int buf = u & 1;
      const bool boundary = (u == 0) && (jb0 == 0);
      // staging boundary: this batch's rows (issued one boundary ago) must have landed before the barrier
      if (boundary) cp_async_wait_all();
      // ---- phase A: owners publish column j of the window and the pivot row of the right-hand sides
      if (lane == own_lane0 + u / EC) {
#pragma unroll
        for (int er = 0; er < ER; ++er) colbuf[buf][tr * ER + er] = a[er][u % EC];
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rres[e] == jr) zbuf[buf][rq[e]] = rv[e];
      __syncthreads();
      // past the barrier nobody reads the other ring half any more (its last reader was the previous
      // pivot): start filling it with the next batch
      if (boundary) stage_issue(j + BATCH, ring + (half ^ 1) * (BATCH * ROWLEN));
      // ---- phase C: rank-1 update of the window, right-hand sides, Gram matrix
      const double inv = invbuf[buf];
      {
        const int jn = (jr + 1 == WD) ? 0 : jr + 1;
        if (tr == jn / ER && lane == jn / EC) {  // next pivot: same arithmetic as the bulk update below
          const double aj1 = colbuf[buf][jn];
          invbuf[buf ^ 1] = 1.0 / fma(-aj1, aj1 * inv, a[(u + 1) % ER][(u + 1) % EC]);
        }
      }
      if (tr == 0) {  // warp-uniform bookkeeping
        const double pj = colbuf[buf][jr];
        if (!(pj > 0.0) && info == 0) info = j + 1;
        const int hi = __double2hiint(pj);
        expo += ((hi >> 20) & 0x7ff) - 1022;
        mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(pj));
        if ((j & 511) == 511) {
          const int h2 = __double2hiint(mant);
          expo += ((h2 >> 20) & 0x7ff) - 1022;
          mant = __hiloint2double((h2 & 0x800fffff) | 0x3fe00000, __double2loint(mant));
        }
      }
      double ak[EC];
#pragma unroll
      for (int ec = 0; ec < EC; ++ec) ak[ec] = colbuf[buf][lane * EC + ec] * inv;
#pragma unroll
      for (int er = 0; er < ER; ++er) {
        const double ai = -colbuf[buf][tr * ER + er];
#pragma unroll
        for (int ec = 0; ec < EC; ++ec) a[er][ec] = fma(ai, ak[ec], a[er][ec]);
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rres[e] >= 0) rv[e] = fma(-colbuf[buf][rres[e]], zbuf[buf][rq[e]] * inv, rv[e]);
      if (gram_on) gacc = fma(zbuf[buf][gp_] * inv, zbuf[buf][gq_], gacc);
      // ---- the row that enters the window (index j + WD) takes the slots of the retiring index j
      const double* row = rowbase + u * ROWLEN;
      if (tr == own_warp0 + u / ER) {
#pragma unroll
        for (int ec = 0; ec < EC; ++ec) {
          int t = lane * EC + ec - jr - 1;
          if (t < 0) t += WD;
          a[u % ER][ec] = row[WD - 1 - t];
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rres[e] == jr) rv[e] = row[WD + rq[e]];
    }
  }
  if (tid == 0) logdet = log(mant) + (double)expo * 0.6931471805599453;

  // ---- epilogue: lnL = −½ (log det S + log det(I + A·G) + RᵀS⁻¹R − uᵀ(I + A·G)⁻¹A u)
  if (gram_on) gram[tid] = gacc;
  __syncthreads();
  if (tid == 0) band_epilogue(p, b, M, gram, logdet, info);
}

// ------------------------------------------------------------------------------------------------
// Rank-4 window update on the fp64 tensor path (mma.sync m8n8k4, SASS DMMA).  Same circular band window in
// registers as band_chol_kernel, but laid out as DMMA accumulator fragments — 16 warps as a 4×4 grid, a warp
// owns (WD/4)×(WD/4) slots as TW×TW tiles of 8×8, lane (g = lane/4, t = lane%4) holds C[g][2t], C[g][2t+1] of
// every tile — and FOUR pivots are eliminated per step:
//   1. the owners publish the raw columns of the four pivots (Praw[row][0..3]) and the four pivot rows of the
//      right-hand sides;                                                                      __syncthreads
//   2. panel: every thread below WD + NR factors the 4×4 diagonal block (LDLᵀ, redundantly — no communication),
//      thread x then solves its row, W[x][t] = P[x][t] − Σ_{k<t} W[x][k]·L[j+t][k], L = W/d, and the threads
//      WD.. do the same for the right-hand-side columns; −W, L, z go to shared memory            __syncthreads
//   3. every warp loads 2·TW fragment values per lane and issues TW² DMMAs: C[r][c] −= Σ_t W[r][t]·L[c][t];
//      the right-hand sides and the Gram matrix take their four rank-1 terms with DFMA;
//   4. the four rows j+WD … j+WD+3 enter together; a row that reaches pivots of the block just eliminated (only
//      when b > WD − 4) receives their updates as it enters, so b <= WD − 1 suffices like for the rank-1 kernel.
// Per pivot this moves 2·TW/4 doubles per thread through shared memory instead of 12 and needs half a barrier
// instead of one; the arithmetic order of every slot equals the rank-1 kernel's, so results agree to rounding.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void band_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__host__ __device__ constexpr int band_gcd(int x, int y) { return y == 0 ? x : band_gcd(y, x % y); }

// WRG × WCG = warp grid over the window (rows × columns); a warp owns TR × TC tiles of 8×8.  The 160-pixel
// window (200 KB) leaves a 512-thread CTA no registers to work with, so it runs as 4×2 warps of 5×10 tiles
// (200 accumulator registers per thread, 256 threads); the narrower ones as 4×4 warps.
// 1/x for a positive normal x: MUFU.RCP64H seed (>= 20 bits) and two Newton steps — within 1 ulp, no special-case
// branches; four of these are chained in every panel, so their latency is on the critical path
__device__ __forceinline__ double band_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// ROLLED = true (used for the 160-pixel window): one copy of the block body instead of 2·lcm(TR,TC); the tile that
// holds the pivot columns / the retiring rows is then a run-time index and only the publish and the entering-row
// code select it (a chain of TC resp. TR predicated copies).  It keeps the 160-pixel kernel (30 k SASS instructions
// unrolled) inside the instruction cache (measured 6.89 vs 7.46 ms).
template <int WD, int WRG, int WCG, int MAXNR, bool ROLLED>
__global__ void __launch_bounds__(32 * WRG * WCG, 1)
band_mma_kernel(BandCholParams p) {
  constexpr int NT = 32 * WRG * WCG, TR = WD / 8 / WRG, TC = WD / 8 / WCG, ROWLEN = WD + NRP, BATCH = 16;
  constexpr int TL = TR * TC / band_gcd(TR, TC);  // tiles per unrolled sweep: lcm(TR, TC)
  constexpr int NE = (WD * MAXNR + NT - 1) / NT;  // right-hand-side registers per thread
  static_assert(TR <= TC && WD % (8 * WRG) == 0 && WD % (8 * WCG) == 0 && (WD / 8) % TL == 0 && WD + MAXNR <= NT &&
                    MAXNR * MAXNR <= NT && BATCH % (NT / 32) == 0,
                "warp grid must tile the window; panel rows, Gram elements and ring rows need enough threads");
  __shared__ __align__(16) double Praw[WD * 4];  // raw columns of the four pivots, [row residue][pivot]
  __shared__ __align__(16) double Wn[WD * 4];    // −W (unscaled panel): the DMMA A operand
  __shared__ __align__(16) double Lp[WD * 4];    // L = W/d: the DMMA B operand
  __shared__ double zraw[4][NRP], zW[4][NRP], zL[4][NRP];
  __shared__ double facL[4];  // L21, L31, L32 of the current block (late-row correction)
  __shared__ double gram[(kMaxM + 1) * (kMaxM + 1)];
  extern __shared__ __align__(16) double ring[];  // [2][BATCH][ROWLEN]

  const int b = p.rowmap ? p.rowmap[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp / WCG, wc = warp % WCG, g = lane >> 2, t = lane & 3;
  const int N = p.N, M = p.M, NR = M + 1;
  if (p.overflow[b] != 0 || *p.sorted == 0) {  // band wider than the window / grid not increasing: not ours
    if (tid == 0) {
      p.info[b] = p.overflow[b] != 0 ? -2 : -3;
      p.lnL[b] = nan("");
    }
    return;
  }
  const double* Sb = p.Sb + (long long)b * p.strideSb;
  const double* Xb = (M > 0) ? p.X + (long long)b * M * N : nullptr;
  const double* Fb = p.model_flux + (long long)b * N;

  auto rhs_at = [&](int i, int q) -> double {
    if (i >= N) return 0.0;
    return q == 0 ? Fb[i] - p.data_flux[i] : Xb[(long long)(q - 1) * N + i];
  };
  // Entering rows j0+WD … j0+WD+BATCH−1, one warp per ring row (BATCH == 16 warps): the band by 16-byte
  // cp.async, X by 8-byte cp.async; the residual column is staged as its two terms (model flux in column WD,
  // data flux in the spare column WD+NRP−1) and subtracted when the row is consumed, so nothing here waits for
  // a global load; rows past the end of the matrix are the identity padding.
  auto stage_issue = [&](int j0, double* dst) {
   for (int rb = warp; rb < BATCH; rb += NT / 32) {
    const int i = j0 + WD + rb;
    double* drow = dst + rb * ROWLEN;
    if (i < N) {
      const double* srow = Sb + (long long)i * WD;
      for (int c2 = lane; c2 < WD / 2; c2 += 32) cp_async16(drow + 2 * c2, srow + 2 * c2);
      if (lane == 0) cp_async8(drow + WD, Fb + i);
      else if (lane < NR) cp_async8(drow + WD + lane, Xb + (long long)(lane - 1) * N + i);
      else if (lane == NRP - 1) cp_async8(drow + WD + NRP - 1, p.data_flux + i);
    } else {
      for (int c = lane; c < ROWLEN; c += 32) drow[c] = (c == 0) ? 1.0 : 0.0;
    }
   }
    cp_async_commit();
  };
  static_assert(NRP <= 32, "one lane per right-hand-side column");

  // ---- initial window: slot (r, c) = element (i = r, k = c) for k <= i
  double a[TR][TC][2];
#pragma unroll
  for (int tr = 0; tr < TR; ++tr)
#pragma unroll
    for (int tc = 0; tc < TC; ++tc)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 8 * (wr * TR + tr) + g, k = 8 * (wc * TC + tc) + 2 * t + h;
        double v = 0.0;
        if (k <= i) v = (i < N) ? Sb[(long long)i * WD + (i - k)] : (k == i ? 1.0 : 0.0);
        a[tr][tc][h] = v;
      }
  double rv[NE];
  int rcode[NE];  // residue·32 + column of the right-hand-side element held in rv[e]; negative: none
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int idx = tid + NT * e;
    const bool ok = idx < WD * NR;
    const int res = idx / NR, q = idx - res * NR;
    rcode[e] = ok ? res * 32 + q : -32;
    rv[e] = ok ? rhs_at(res, q) : 0.0;
  }
  stage_issue(0, ring);

  // Four rows enter together.  If the band comes within three pixels of the window width (b > WD − 4), an entering
  // row j+WD+s reaches pivots j+s+1 … j+3 of the block that has just been eliminated without it; those updates are
  // then applied as the row enters (below).  Whether this walker needs that is a CTA-uniform fact of its band.
  int late_any = 0;
  for (int idx = tid; idx < 3 * N; idx += NT) {
    const int i = idx / 3, d = WD - 1 - (idx - 3 * i);
    late_any |= (Sb[(long long)i * WD + d] != 0.0) ? 1 : 0;
  }
  const bool late = __syncthreads_or(late_any) != 0;

  double logdet = 0.0, gacc = 0.0, mant = 1.0;
  long long expo = 0;
  int info = 0;
  const int gp_ = tid / NR, gq_ = tid - (tid / NR) * NR;  // Gram element of this thread (tid < NR²)
  const bool gram_on = tid < NR * NR;

  // The block loop is unrolled over lcm(TR, TC) tiles (two blocks of four pivots each), so that the tile-in-warp
  // index of the pivots' columns / of the retiring rows is a compile-time register index.
  constexpr int UBN = ROLLED ? 1 : 2 * TL;  // blocks per trip of the outer loop
  for (int J = 0, jrJ = 0; J < N; J += 4 * UBN, jrJ = (jrJ + 4 * UBN == WD) ? 0 : jrJ + 4 * UBN) {
#pragma unroll
    for (int ub = 0; ub < UBN; ++ub) {
      const int j = J + 4 * ub, jr = jrJ + 4 * ub;
      if (j >= N) break;
      int tci, tri, pp;  // tile inside the warp (columns / rows), half of the tile
      int ownc, ownr;    // warp column holding these pivot columns, warp row holding the retiring rows
      if constexpr (ROLLED) {
        const int ct = jr >> 3;
        tci = ct % TC; tri = ct % TR; pp = (jr >> 2) & 1;
        ownc = ct / TC; ownr = ct / TR;
      } else {
        tci = (ub >> 1) % TC; tri = (ub >> 1) % TR; pp = ub & 1;
        ownc = (jrJ >> 3) / TC + (ub >> 1) / TC;
        ownr = (jrJ >> 3) / TR + (ub >> 1) / TR;
      }
      const bool boundary = (j & (BATCH - 1)) == 0;
      if (boundary) cp_async_wait_all();  // this batch's entering rows have landed (visible after the barrier)
      // ---- 1. publish the raw pivot columns and the pivot rows of the right-hand sides
      if (wc == ownc && (t >> 1) == pp) {
        if constexpr (ROLLED) {
#pragma unroll
          for (int q = 0; q < TC; ++q)
            if (q == tci) {
#pragma unroll
              for (int tr = 0; tr < TR; ++tr)
                *reinterpret_cast<double2*>(&Praw[(8 * (wr * TR + tr) + g) * 4 + 2 * (t & 1)]) =
                    make_double2(a[tr][q][0], a[tr][q][1]);
            }
        } else {
#pragma unroll
          for (int tr = 0; tr < TR; ++tr)
            *reinterpret_cast<double2*>(&Praw[(8 * (wr * TR + tr) + g) * 4 + 2 * (t & 1)]) =
                make_double2(a[tr][tci][0], a[tr][tci][1]);
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int rs = rcode[e] >> 5;
        if (rs >= jr && rs < jr + 4) zraw[rs - jr][rcode[e] & 31] = rv[e];
      }
      __syncthreads();
      // nobody reads the other ring half any more (its last readers were the previous block's entering rows)
      if (boundary) stage_issue(j + BATCH, ring + (((j / BATCH) & 1) ^ 1) * (BATCH * ROWLEN));
      // ---- 2. panel: 4×4 LDLᵀ of the diagonal block (redundantly), then one row / one rhs column per thread
      if (tid < WD + NR) {
        const double* Dp = Praw + jr * 4;
        const double D00 = Dp[0], D10 = Dp[4], D11 = Dp[5], D20 = Dp[8], D21 = Dp[9], D22 = Dp[10];
        const double D30 = Dp[12], D31 = Dp[13], D32 = Dp[14], D33 = Dp[15];
        const double inv0 = band_rcp(D00);
        const double L10 = D10 * inv0, L20 = D20 * inv0, L30 = D30 * inv0;
        const double W11 = fma(-D10, L10, D11);
        const double inv1 = band_rcp(W11);
        const double W21 = fma(-D20, L10, D21), W31 = fma(-D30, L10, D31);
        const double L21 = W21 * inv1, L31 = W31 * inv1;
        const double W22 = fma(-W21, L21, fma(-D20, L20, D22));
        const double inv2 = band_rcp(W22);
        const double W32 = fma(-W31, L21, fma(-D30, L20, D32));
        const double L32 = W32 * inv2;
        const double W33 = fma(-W32, L32, fma(-W31, L31, fma(-D30, L30, D33)));
        const double inv3 = band_rcp(W33);
        if (tid < WD) {
          const double2 p01 = *reinterpret_cast<const double2*>(&Praw[tid * 4]);
          const double2 p23 = *reinterpret_cast<const double2*>(&Praw[tid * 4 + 2]);
          const double w0 = p01.x;
          const double w1 = fma(-w0, L10, p01.y);
          const double w2 = fma(-w1, L21, fma(-w0, L20, p23.x));
          const double w3 = fma(-w2, L32, fma(-w1, L31, fma(-w0, L30, p23.y)));
          *reinterpret_cast<double2*>(&Wn[tid * 4]) = make_double2(-w0, -w1);
          *reinterpret_cast<double2*>(&Wn[tid * 4 + 2]) = make_double2(-w2, -w3);
          *reinterpret_cast<double2*>(&Lp[tid * 4]) = make_double2(w0 * inv0, w1 * inv1);
          *reinterpret_cast<double2*>(&Lp[tid * 4 + 2]) = make_double2(w2 * inv2, w3 * inv3);
        } else {
          const int q = tid - WD;
          const double z0 = zraw[0][q], zl0 = z0 * inv0;
          const double z1 = fma(-D10, zl0, zraw[1][q]), zl1 = z1 * inv1;
          const double z2 = fma(-W21, zl1, fma(-D20, zl0, zraw[2][q])), zl2 = z2 * inv2;
          const double z3 = fma(-W32, zl2, fma(-W31, zl1, fma(-D30, zl0, zraw[3][q]))), zl3 = z3 * inv3;
          zW[0][q] = z0; zW[1][q] = z1; zW[2][q] = z2; zW[3][q] = z3;
          zL[0][q] = zl0; zL[1][q] = zl1; zL[2][q] = zl2; zL[3][q] = zl3;
        }
        if (tid == 0) {
          facL[0] = L21;
          facL[1] = L31;
          facL[2] = L32;
        }
        if (tid == 0) {  // log det S = Σ log(pivot) as mantissa × 2^exponent (see band_chol_kernel), info
          const double piv[4] = {D00, W11, W22, W33};
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const double pj = piv[s4];
            if (!(pj > 0.0) && info == 0) info = j + s4 + 1;
            const int hi = __double2hiint(pj);
            expo += ((hi >> 20) & 0x7ff) - 1022;
            mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(pj));
          }
          if ((j & 511) == 508) {
            const int h2 = __double2hiint(mant);
            expo += ((h2 >> 20) & 0x7ff) - 1022;
            mant = __hiloint2double((h2 & 0x800fffff) | 0x3fe00000, __double2loint(mant));
          }
        }
      }
      __syncthreads();
      // ---- 3. rank-4 update of the window (DMMA), of the right-hand sides and of the Gram matrix (DFMA)
      {
        double af[TR];  // TR <= TC: the shorter fragment set is held, the longer one streamed
#pragma unroll
        for (int tr = 0; tr < TR; ++tr) af[tr] = Wn[(8 * (wr * TR + tr) + g) * 4 + t];
#pragma unroll
        for (int tc = 0; tc < TC; ++tc) {
          const double bf = Lp[(8 * (wc * TC + tc) + g) * 4 + t];
#pragma unroll
          for (int tr = 0; tr < TR; ++tr) band_dmma(a[tr][tc][0], a[tr][tc][1], af[tr], bf);
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rcode[e] >= 0) {
          const int rs = rcode[e] >> 5, q = rcode[e] & 31;
          const double2 w01 = *reinterpret_cast<const double2*>(&Wn[rs * 4]);
          const double2 w23 = *reinterpret_cast<const double2*>(&Wn[rs * 4 + 2]);
          double r = rv[e];
          r = fma(w01.x, zL[0][q], r);
          r = fma(w01.y, zL[1][q], r);
          r = fma(w23.x, zL[2][q], r);
          r = fma(w23.y, zL[3][q], r);
          rv[e] = r;
        }
      if (gram_on) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) gacc = fma(zL[s4][gp_], zW[s4][gq_], gacc);
      }
      // ---- 4. indices j+WD … j+WD+3 take over the residues jr … jr+3: their rows of S replace the retiring rows
      const double* rowbase = ring + ((j / BATCH) & 1) * (BATCH * ROWLEN) + (j & (BATCH - 1)) * ROWLEN;
      // panel values of a late row (s = its position in the block): W'[t] = S(j+WD+s, j+t) − Σ_{s<k<t} W'[k]·L[j+t][k]
      auto late_w = [&](const double* row, int sblk, double& w1, double& w2, double& w3) {
        const double r1 = (sblk < 1) ? row[WD + sblk - 1] : 0.0;
        const double r2 = (sblk < 2) ? row[WD + sblk - 2] : 0.0;
        const double r3 = (sblk < 3) ? row[WD + sblk - 3] : 0.0;
        w1 = r1;
        w2 = fma(-w1, facL[0], r2);
        w3 = fma(-w2, facL[2], fma(-w1, facL[1], r3));
      };
      if (wr == ownr && (g >> 2) == pp) {
        const int rres_new = jr + (g & 3);
        const double* row = rowbase + (g & 3) * ROWLEN;
        double w1 = 0.0, w2 = 0.0, w3 = 0.0;
        if (late) late_w(row, g & 3, w1, w2, w3);
#pragma unroll
        for (int tc = 0; tc < TC; ++tc)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = 8 * (wc * TC + tc) + 2 * t + h;
            int dd = rres_new - c;  // band offset of the slot's column
            if (dd < 0) dd += WD;
            double v = row[dd];
            if (late && (unsigned)(c - jr) >= 4u) {  // not the block's own residues (new / dead columns)
              const double l1 = Lp[c * 4 + 1];
              const double2 l23 = *reinterpret_cast<const double2*>(&Lp[c * 4 + 2]);
              v = fma(-w3, l23.y, fma(-w2, l23.x, fma(-w1, l1, v)));
            }
            if constexpr (ROLLED) {
#pragma unroll
              for (int q = 0; q < TR; ++q)
                if (q == tri) a[q][tc][h] = v;
            } else {
              a[tri][tc][h] = v;
            }
          }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int rs = rcode[e] >> 5, q = rcode[e] & 31;
        if (rs >= jr && rs < jr + 4) {
          const double* row = rowbase + (rs - jr) * ROWLEN;
          const double* rr = row + WD;
          double v = (q == 0) ? rr[0] - rr[NRP - 1] : rr[q];  // residual = model flux − data flux
          if (late) {
            double w1, w2, w3;
            late_w(row, rs - jr, w1, w2, w3);
            v = fma(-w3, zL[3][q], fma(-w2, zL[2][q], fma(-w1, zL[1][q], v)));
          }
          rv[e] = v;
        }
      }
    }
  }
  if (tid == 0) logdet = log(mant) + (double)expo * 0.6931471805599453;

  if (gram_on) gram[tid] = gacc;
  __syncthreads();
  if (tid == 0) band_epilogue(p, b, M, gram, logdet, info);
}

template <int WD, int WRG, int WCG, bool ROLLED>
cudaError_t launch_band_mma_tr(const BandCholParams& p, int B, cudaStream_t st) {
  const size_t smem = sizeof(double) * 2 * 16 * (WD + NRP);
  constexpr int NT = 32 * WRG * WCG;
  if (p.M + 1 <= 8) {
    cudaError_t e = cudaFuncSetAttribute(band_mma_kernel<WD, WRG, WCG, 8, ROLLED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    band_mma_kernel<WD, WRG, WCG, 8, ROLLED><<<B, NT, smem, st>>>(p);
  } else if constexpr (WD + kMaxM + 1 <= NT && (kMaxM + 1) * (kMaxM + 1) <= NT) {
    cudaError_t e = cudaFuncSetAttribute(band_mma_kernel<WD, WRG, WCG, kMaxM + 1, ROLLED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    band_mma_kernel<WD, WRG, WCG, kMaxM + 1, ROLLED><<<B, NT, smem, st>>>(p);
  } else {
    return cudaErrorInvalidValue;  // launch_band_chol routes these to the rank-1 kernel
  }
  return cudaGetLastError();
}

// The unrolled body (tile indices are compile-time register indices) is the faster one while it fits the instruction
// cache: 96 px 2.76 ms vs 2.99 rolled, 128 px 4.22 vs 4.41.  At 160 px the unrolled loop is 30 k SASS instructions and
// misses the cache; the rolled form (one copy of the block body, run-time tile index) runs 6.89 ms vs 7.46
// (profiles/r2a_band_ab.txt, one wave of CTAs each).
template <int WD, int WRG, int WCG>
cudaError_t launch_band_mma_t(const BandCholParams& p, int B, cudaStream_t st) {
  return launch_band_mma_tr<WD, WRG, WCG, (WD >= 160)>(p, B, st);
}

// ------------------------------------------------------------------------------------------------
// Symmetric window.  The active window W[r][c] = S(index(r), index(c)) (residues mod WD) is a symmetric
// matrix and the rank-1 update W −= v·vᵀ/pivot is symmetric as well, so only one of every pair of T×T tiles
// {(I,J), (J,I)} needs to exist: tile (I, J) is kept iff (J − I) mod 32 ≤ 16.  Warp δ ∈ [0,16], lane I owns
// tile (I, (I+δ) mod 32): 17 warps, T² registers per thread (WD = 32·T) — half the FMAs and half the
// registers of the full square, which lets two walkers share an SM for T ≤ 4, and windows up to 256 pixels.
// The pivot column v (all pairs {r, jr}) is scattered over the tiles that touch residue jr: in warp δ the
// lane (Jp − δ) mod 32 holds a column piece of it and — for δ ≥ 1 — lane Jp a row piece (Jp = jr / T); the same
// threads take the entering row's values when index j + WD replaces j.  The pivot loop is unrolled by T, so
// jr mod T is a compile-time register index.
// ------------------------------------------------------------------------------------------------
template <int T, int MAXNR>
__global__ void __launch_bounds__(17 * 32, (T <= 4) ? 2 : 1)
band_sym_kernel(BandCholParams p) {
  constexpr int WD = 32 * T, NT = 17 * 32, ROWLEN = WD + NRP;
  constexpr int BATCH = ((16 + T - 1) / T) * T;             // pivots per staging batch (a multiple of T)
  constexpr int NE = (WD * MAXNR + NT - 1) / NT;            // right-hand-side registers per thread
  __shared__ double colbuf[2][WD];
  __shared__ double zbuf[2][NRP];
  __shared__ double invbuf[2];
  __shared__ double gram[(kMaxM + 1) * (kMaxM + 1)];
  extern __shared__ double ring[];  // [2][BATCH][ROWLEN]

  const int b = p.rowmap ? p.rowmap[blockIdx.x] : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, dl = tid >> 5;  // dl = tile offset δ of this warp
  const int N = p.N, M = p.M, NR = M + 1;
  if (p.overflow[b] != 0 || *p.sorted == 0) {
    if (tid == 0) {
      p.info[b] = p.overflow[b] != 0 ? -2 : -3;
      p.lnL[b] = nan("");
    }
    return;
  }
  const double* Sb = p.Sb + (long long)b * p.strideSb;
  const double* Xb = (M > 0) ? p.X + (long long)b * M * N : nullptr;
  const double* Fb = p.model_flux + (long long)b * N;
  const int r0 = T * lane, c0 = T * ((lane + dl) & 31);     // first row / column residue of this thread's tile

  auto rhs_at = [&](int i, int q) -> double {
    if (i >= N) return 0.0;
    return q == 0 ? Fb[i] - p.data_flux[i] : Xb[(long long)(q - 1) * N + i];
  };
  auto stage_issue = [&](int j0, double* dst) {
    for (int sidx = tid; sidx < BATCH * ROWLEN; sidx += NT) {
      const int rb = sidx / ROWLEN, col = sidx - rb * ROWLEN;
      const int i = j0 + WD + rb;
      if (col < WD) {
        if (i < N) cp_async8(dst + sidx, Sb + (long long)i * WD + col);
        else dst[sidx] = (col == 0) ? 1.0 : 0.0;
      } else {
        const int q = col - WD;
        if (q >= 1 && q < NR && i < N) cp_async8(dst + sidx, Xb + (long long)(q - 1) * N + i);
        else dst[sidx] = (q == 0) ? rhs_at(i, 0) : 0.0;
      }
    }
    cp_async_commit();
  };

  // ---- initial window: residue = index for the first WD indices
  double a[T][T];
#pragma unroll
  for (int rr = 0; rr < T; ++rr)
#pragma unroll
    for (int cc = 0; cc < T; ++cc) {
      const int i = max(r0 + rr, c0 + cc), k = min(r0 + rr, c0 + cc);
      a[rr][cc] = (i < N) ? Sb[(long long)i * WD + (i - k)] : (i == k ? 1.0 : 0.0);
    }
  double rv[NE];
  int rres[NE], rq[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const int idx = tid + NT * e;
    const bool ok = idx < WD * NR;
    rres[e] = ok ? idx / NR : -1;
    rq[e] = ok ? idx - (idx / NR) * NR : 0;
    rv[e] = ok ? rhs_at(rres[e], rq[e]) : 0.0;
  }
  stage_issue(0, ring);

  double logdet = 0.0, gacc = 0.0, mant = 1.0;
  long long expo = 0;
  int info = 0;
  const int gp_ = tid / NR, gq_ = tid - (tid / NR) * NR;
  const bool gram_on = tid < NR * NR;

  if (tid == 0) invbuf[0] = 1.0 / a[0][0];
  for (int j0 = 0, jr0 = 0; j0 < N; j0 += T, jr0 = (jr0 + T == WD) ? 0 : jr0 + T) {
    const int Jp = jr0 / T;                                   // tile index of the pivots of this block
    const bool colpiece = lane == ((Jp - dl) & 31);           // holds W[r0..r0+T)[jr]   (its column jr mod T)
    const bool rowpiece = (dl >= 1) && (lane == Jp);          // holds W[jr][c0..c0+T)   (its row jr mod T)
    const bool diagtile = (dl == 0) && (lane == Jp);
    const int jb0 = j0 % BATCH, half = (j0 / BATCH) & 1;
    const double* rowbase = ring + half * (BATCH * ROWLEN) + jb0 * ROWLEN;
#pragma unroll
    for (int u = 0; u < T; ++u) {
      const int j = j0 + u, jr = jr0 + u;
      if (j >= N) break;
      const int buf = ((T % 2) == 0) ? (u & 1) : (j & 1);
      const bool boundary = (u == 0) && (jb0 == 0);
      if (boundary) cp_async_wait_all();
      // ---- phase A: publish column jr of the symmetric window and the pivot row of the right-hand sides
      if (colpiece) {
#pragma unroll
        for (int rr = 0; rr < T; ++rr) colbuf[buf][r0 + rr] = a[rr][u];
      }
      if (rowpiece && dl <= 15) {
#pragma unroll
        for (int cc = 0; cc < T; ++cc) colbuf[buf][c0 + cc] = a[u][cc];
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rres[e] == jr) zbuf[buf][rq[e]] = rv[e];
      __syncthreads();
      if (boundary) stage_issue(j + BATCH, ring + (half ^ 1) * (BATCH * ROWLEN));
      // ---- phase C: symmetric rank-1 update, right-hand sides, Gram matrix
      const double inv = invbuf[buf];
      if (diagtile || (dl == 0 && u == T - 1)) {  // reciprocal of the next pivot, one step ahead
        const int jn = (jr + 1 == WD) ? 0 : jr + 1;
        if (lane == jn / T) {
          const double x = colbuf[buf][jn];
          invbuf[buf ^ 1] = 1.0 / fma(-x, x * inv, a[(u + 1) % T][(u + 1) % T]);
        }
      }
      if (dl == 0) {  // warp-uniform bookkeeping
        const double pj = colbuf[buf][jr];
        if (!(pj > 0.0) && info == 0) info = j + 1;
        const int hi = __double2hiint(pj);
        expo += ((hi >> 20) & 0x7ff) - 1022;
        mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(pj));
        if ((j & 511) == 511) {
          const int h2 = __double2hiint(mant);
          expo += ((h2 >> 20) & 0x7ff) - 1022;
          mant = __hiloint2double((h2 & 0x800fffff) | 0x3fe00000, __double2loint(mant));
        }
      }
      double vc[T];
#pragma unroll
      for (int cc = 0; cc < T; ++cc) vc[cc] = colbuf[buf][c0 + cc] * inv;
#pragma unroll
      for (int rr = 0; rr < T; ++rr) {
        const double vr = -colbuf[buf][r0 + rr];
#pragma unroll
        for (int cc = 0; cc < T; ++cc) a[rr][cc] = fma(vr, vc[cc], a[rr][cc]);
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rres[e] >= 0) rv[e] = fma(-colbuf[buf][rres[e]], zbuf[buf][rq[e]] * inv, rv[e]);
      if (gram_on) gacc = fma(zbuf[buf][gp_] * inv, zbuf[buf][gq_], gacc);
      // ---- index j + WD takes over residue jr: its row of S replaces every pair {·, jr}
      const double* row = rowbase + u * ROWLEN;
      auto entering = [&](int c) -> double {  // S(j+WD, index of residue c in the new window)
        int t = c - jr - 1;
        if (t < 0) t += WD;
        return row[WD - 1 - t];
      };
      if (colpiece) {
#pragma unroll
        for (int rr = 0; rr < T; ++rr) a[rr][u] = entering(r0 + rr);
      }
      if (rowpiece || diagtile) {
#pragma unroll
        for (int cc = 0; cc < T; ++cc) a[u][cc] = entering(c0 + cc);
      }
#pragma unroll
      for (int e = 0; e < NE; ++e)
        if (rres[e] == jr) rv[e] = row[WD + rq[e]];
    }
  }
  if (tid == 0) logdet = log(mant) + (double)expo * 0.6931471805599453;
  if (gram_on) gram[tid] = gacc;
  __syncthreads();
  if (tid == 0) band_epilogue(p, b, M, gram, logdet, info);
}

template <int T>
cudaError_t launch_band_sym_t(const BandCholParams& p, int B, cudaStream_t st) {
  constexpr int WD = 32 * T, BATCH = ((16 + T - 1) / T) * T;
  const size_t smem = sizeof(double) * 2 * BATCH * (WD + NRP);
  // the opt-in is a per-device function attribute: set it on every launch (a handle may live on any device)
  if (p.M + 1 <= 8) {
    cudaError_t e = cudaFuncSetAttribute(band_sym_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    band_sym_kernel<T, 8><<<B, 17 * 32, smem, st>>>(p);
  } else {
    cudaError_t e =
        cudaFuncSetAttribute(band_sym_kernel<T, kMaxM + 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    band_sym_kernel<T, kMaxM + 1><<<B, 17 * 32, smem, st>>>(p);
  }
  return cudaGetLastError();
}

template <int WD, int ER>
cudaError_t launch_band_t(const BandCholParams& p, int B, cudaStream_t st) {
  const size_t smem = sizeof(double) * 2 * band_batch(WD, ER) * (WD + NRP);
  // static + dynamic shared memory exceeds 48 KB for the widest window; the opt-in is a per-device function
  // attribute, so it is set on every launch (a handle may live on any device)
  if (p.M + 1 <= 8) {
    cudaError_t e = cudaFuncSetAttribute(band_chol_kernel<WD, ER, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    band_chol_kernel<WD, ER, 8><<<B, WD * 32 / ER, smem, st>>>(p);
  } else {
    cudaError_t e = cudaFuncSetAttribute(band_chol_kernel<WD, ER, kMaxM + 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    band_chol_kernel<WD, ER, kMaxM + 1><<<B, WD * 32 / ER, smem, st>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace

const int kBandWidths[] = {64, 96, 128, 160, 192, 256};
const int kNumBandWidths = 6;

__global__ void residual_only_kernel(const double* __restrict__ F, const double* __restrict__ data, int N,
                                     double* __restrict__ resid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i < N) resid[(long long)b * N + i] = F[(long long)b * N + i] - data[i];
}

// ------------------------------------------------------------------------------------------------
// Shared-factor path (frozen kernels): the right-hand sides of walker b, already solved against the shared factor
// (rows b·(M+1) .. b·(M+1)+M of Zt), give its Gram matrix; band_epilogue turns it into lnL with the shared logdet.
// One CTA per walker, one warp per Gram entry at a time, fixed summation order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_capacitance_kernel(const double* __restrict__ Zt, int ldz, int N, int M,
                                                               const double* A, const double* logdet_S,
                                                               const int* info_S, double* lnL, int* info) {
  __shared__ double gram[(kMaxM + 1) * (kMaxM + 1)];
  const int b = blockIdx.x, NR = M + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* Zb = Zt + (long long)b * NR * ldz;
  for (int pq = warp; pq < NR * NR; pq += 8) {
    const int i = pq / NR, j = pq % NR;
    if (j > i) continue;
    const double* zi = Zb + (long long)i * ldz;
    const double* zj = Zb + (long long)j * ldz;
    double acc = 0.0;
    for (int c = lane; c < N; c += 32) acc = fma(zi[c], zj[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) { gram[i * NR + j] = acc; gram[j * NR + i] = acc; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    BandCholParams p;
    p.N = N; p.M = M; p.A = A; p.lnL = lnL; p.info = info;
    band_epilogue(p, b, M, gram, *logdet_S, *info_S);
  }
}

cudaError_t launch_gram_capacitance(const double* Zt, int ldz, int N, int M, int B, const double* A,
                                    const double* logdet_S, const int* info_S, double* lnL, int* info, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  gram_capacitance_kernel<<<B, 256, 0, st>>>(Zt, ldz, N, M, A, logdet_S, info_S, lnL, info);
  return cudaGetLastError();
}

cudaError_t launch_residual_only(const double* model_flux, const double* data_flux, int N, int B, double* resid,
                                 cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  residual_only_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(model_flux, data_flux, N, resid);
  return cudaGetLastError();
}

cudaError_t launch_band_width(int N, int Kmax, int hyper_stride, const double* wave, const double* glob,
                              const int* nloc, const double* loc, int* bw, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  band_width_kernel<<<B, 256, 0, st>>>(N, Kmax, hyper_stride, wave, glob, nloc, loc, bw);
  return cudaGetLastError();
}

cudaError_t launch_band_build(const BandBuildParams& p, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  band_build_kernel<<<dim3((p.N + BB_ROWS - 1) / BB_ROWS, B), BB_ROWS * 32, 0, st>>>(p);
  return cudaGetLastError();
}

// Windows up to 160 pixels use the full-square register window: the rank-4 DMMA kernel, or the rank-1 kernel for
// more than 8 right-hand sides at 160 pixels (the 256-thread 160-pixel DMMA kernel carries at most 8).  The
// symmetric window serves the 192- and 256-pixel classes, which do not fit the register file as a square.
// pixels of slack a window needs beyond the half-bandwidth: b + slack <= WD
int band_slack(int WD) {
  (void)WD;
  return 1;
}

cudaError_t launch_band_chol(const BandCholParams& p, int WD, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  switch (WD) {
    case 64: return launch_band_mma_t<64, 4, 4>(p, B, st);
    case 96: return launch_band_mma_t<96, 4, 4>(p, B, st);
    case 128: return launch_band_mma_t<128, 4, 4>(p, B, st);
    case 160: return (p.M + 1 <= 8) ? launch_band_mma_t<160, 4, 2>(p, B, st) : launch_band_t<160, 10>(p, B, st);
    case 192: return launch_band_sym_t<6>(p, B, st);
    case 256: return launch_band_sym_t<8>(p, B, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sfb
