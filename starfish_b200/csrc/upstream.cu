// Upstream of the covariance: everything SpectrumModel.__call__ does before the rank-M term
// (Starfish/models/spectrum_model.py:287-332), batched over walkers, on the device.
//
//   gp_predict_kernel   Emulator.__call__                     Starfish/emulator/emulator.py:382-388
//                       (+ batch_kernel / rbf_kernel          Starfish/emulator/kernels.py:5-49)
//                       + Σ_w⁻¹ of spectrum_model.py:334-335
//   rot_transfer_kernel rotational_broaden's transfer fn      Starfish/transforms.py:124-131
//   broaden_kernel      rfft · sb → irfft                     Starfish/transforms.py:126-133
//   resample_kernel     doppler_shift + resample (k=5 spline) Starfish/transforms.py:11-42, 137-158
//   combine_kernel      chebyshev_correct, X = eig·std, flux = w·X + mean, rescale / renorm
//                                                            Starfish/transforms.py:209-304,
//                                                            spectrum_model.py:301-332
//
// What is static is hoisted to set-up (ModelState): the spectrum of the bulk fluxes (the reference
// re-runs rfft on the same arrays every call), the Cholesky factor of v11 and L⁻¹ŵ (the reference
// re-solves v11 twice per call), and the spline collocation problem.  The Doppler shift only rescales
// the knot vector, and B-spline values are invariant under a common scaling of knots and abscissae, so
// the collocation matrix never changes; its inverse is banded to working precision (entries decay by
// 0.4306 per knot for quintic splines), which turns the per-call spline fit into a (2W+1)-tap filter
// that is embarrassingly parallel — no sequential banded solve on the device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "sfb_internal.cuh"

namespace sfb {

namespace {

constexpr int kDeg = 5;
constexpr double kTwoPi = 6.283185307179586;  // 2.0 * numpy.pi

// ------------------------------------------------------------------------------------------------
// quintic B-spline helpers shared by host set-up and device evaluation.
// Knot vector of InterpolatedUnivariateSpline(x, y, k=5) (FITPACK fpcurf, s = 0): t[0..5] = x[0],
// t[6+j] = x[3+j] (j = 0..n-7), t[n..n+5] = x[n-1]; n coefficients.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ double knot_at(const double* __restrict__ fw, int nf, int l) {
  int i = l - 3;
  if (l < 6) i = 0;
  if (l >= nf) i = nf - 1;
  return fw[i];
}

// FITPACK splev's interval search: t[l] <= x < t[l+1], clamped to [k, n-1] (end pieces extrapolate).
__host__ __device__ __forceinline__ int find_interval(const double* __restrict__ fw, int nf, double x,
                                                      double scale) {
  int lo = kDeg, hi = nf - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (knot_at(fw, nf, mid) * scale <= x)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// FITPACK fpbspl: the k+1 non-zero B-splines of degree k at x, t[l] <= x < t[l+1].
__host__ __device__ __forceinline__ void bspl6(const double* __restrict__ fw, int nf, double x, int l,
                                               double scale, double* h) {
  double hh[kDeg];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= kDeg; ++j) {
#pragma unroll
    for (int i = 0; i < kDeg; ++i)
      if (i < j) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll
    for (int i = 0; i < kDeg; ++i) {
      if (i < j) {
        const double tli = knot_at(fw, nf, l + i + 1) * scale;
        const double tlj = knot_at(fw, nf, l + i + 1 - j) * scale;
        const double f = hh[i] / (tli - tlj);
        h[i] = h[i] + f * (tli - x);
        h[i + 1] = f * (x - tlj);
      }
    }
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// exp(+2πi j/nf) for any j >= 0 from the half table T[0 .. nf/2)
__device__ __forceinline__ double2 twiddle(const double2* __restrict__ T, int nf, long long j) {
  int jj = (int)(j & (long long)(nf - 1));
  const int n2 = nf >> 1;
  if (jj >= n2) {
    const double2 t = T[jj - n2];
    return make_double2(-t.x, -t.y);
  }
  return T[jj];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// (1) emulator GP predictive, one CTA per walker.
//     v12 (block diagonal: column m lives in rows [mG,(m+1)G)) → u = L⁻¹v12 as a matrix-vector product
//     with the explicit inverse of the Cholesky factor → mu = uᵀ(L⁻¹ŵ), Σ_w = diag(var) − uᵀu → A = Σ_w⁻¹.
//     The u-form (rather than an explicit v11⁻¹) keeps the cancellation 1e4 → O(1) as accurate as the
//     reference's LU solves (measured against 50-digit arithmetic: 2.7e-12 vs 2.9e-12 relative; an
//     explicit v11⁻¹ gives 2e-8).
// ------------------------------------------------------------------------------------------------
constexpr int GP_THREADS = 256;

__global__ void __launch_bounds__(GP_THREADS)
gp_predict_kernel(int M, int G, int D, int n, int ld, const double* __restrict__ grid,
                  const double* __restrict__ var, const double* __restrict__ ls, const double* __restrict__ L,
                  const double* __restrict__ zw, const double* __restrict__ theta, int ntheta, int paper,
                  double* __restrict__ mu_out, double* __restrict__ wcov_out, double* __restrict__ A_out,
                  int* __restrict__ status) {
  extern __shared__ double u[];  // ld × M, row-major
  __shared__ double red[kMaxM + kMaxM * kMaxM];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = GP_THREADS / 32;
  const double* p = theta + (long long)b * ntheta;

  for (int idx = tid; idx < ld * M; idx += GP_THREADS) {
    const int row = idx / M, m = idx - row * M;
    double v = 0.0;
    if (row < n && row / G == m) {
      const int g = row - m * G;
      double s = 0.0;
      for (int d = 0; d < D; ++d) {
        const double l = ls[m * D + d];
        const double diff = grid[g * D + d] / l - p[d] / l;
        s += diff * diff;
      }
      v = var[m] * exp(-0.5 * s);
    }
    u[idx] = v;
  }
  __syncthreads();

  // u = L⁻¹·v12 with the explicit inverse of the Cholesky factor (set-up): column m of v12 is non-zero only
  // in rows [mG,(m+1)G), so u[r][m] = Σ_g Linv[r][mG+g]·k_m[g] — one warp per row, lanes over g, no
  // sequential dependency.  (Inverting the triangular factor is benign — its condition number is the
  // square root of v11's; it is the explicit v11⁻¹ that loses the 1e4 → O(1) cancellation.)
  {
    double* kk = u + (size_t)ld * M;  // [n] kernel values, block m at offset mG
    for (int idx = tid; idx < n; idx += GP_THREADS) {
      const int m = idx / G;
      kk[idx] = u[idx * M + m];
    }
    __syncthreads();
    for (int r = warp; r < n; r += NW) {
      const double* Lr = L + (long long)r * ld;
      for (int m = 0; m < M; ++m) {
        double acc = 0.0;
        const int c1 = min((m + 1) * G, r + 1);
        for (int c = m * G + lane; c < c1; c += 32) acc = fma(Lr[c], kk[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) u[r * M + m] = acc;
      }
    }
    __syncthreads();
  }

  // reductions: mu_m = Σ_r u[r][m]·zw[r];  S[m][m'] = Σ_r u[r][m]·u[r][m']
  const int nred = M + M * M;
  for (int q = warp; q < nred; q += NW) {
    double acc = 0.0;
    if (q < M) {
      for (int r = lane; r < n; r += 32) acc = fma(u[r * M + q], zw[r], acc);
    } else {
      const int m1 = (q - M) / M, m2 = (q - M) - m1 * M;
      for (int r = lane; r < n; r += 32) acc = fma(u[r * M + m1], u[r * M + m2], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) red[q] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    double S[kMaxM][kMaxM], Li[kMaxM][kMaxM];
    int bad = 0;
    for (int i = 0; i < M; ++i) {
      mu_out[(long long)b * M + i] = red[i];
      for (int j = 0; j < M; ++j) {
        const double s = 0.5 * (red[M + i * M + j] + red[M + j * M + i]);
        S[i][j] = (i == j ? var[i] : 0.0) - s;
        wcov_out[((long long)b * M + i) * M + j] = S[i][j];
      }
    }
    if (paper) {
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) A_out[((long long)b * M + i) * M + j] = S[i][j];
    } else {
      // Cholesky S = C·Cᵀ (lower, in place), Li = C⁻¹, A = LiᵀLi  (= cho_solve(cho_factor(Σ_w), I))
      for (int j = 0; j < M && !bad; ++j) {
        double d = S[j][j];
        for (int k = 0; k < j; ++k) d -= S[j][k] * S[j][k];
        if (!(d > 0.0)) { bad = 1; break; }
        d = sqrt(d);
        S[j][j] = d;
        for (int i = j + 1; i < M; ++i) {
          double s = S[i][j];
          for (int k = 0; k < j; ++k) s -= S[i][k] * S[j][k];
          S[i][j] = s / d;
        }
      }
      if (!bad) {
        for (int c = 0; c < M; ++c) {
          for (int i = 0; i < M; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= S[i][k] * Li[k][c];
            Li[i][c] = (i < c) ? 0.0 : s / S[i][i];
          }
        }
      }
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
          double s = 0.0;
          if (!bad)
            for (int k = (i > j ? i : j); k < M; ++k) s += Li[k][i] * Li[k][j];
          A_out[((long long)b * M + i) * M + j] = bad ? nan("") : s;
        }
    }
    status[b] = bad;
  }
}

// ------------------------------------------------------------------------------------------------
// (2a) Gray's rotational transfer function per walker:  sb[0] = 1,
//      sb[k] = j1(ub)/ub − 3cos(ub)/(2ub²) + 3sin(ub)/(2ub³),  ub = 2π·vsini·k/(nf·dv)
// ------------------------------------------------------------------------------------------------
__global__ void rot_transfer_kernel(int n2p1, double freq_val, const double* __restrict__ theta, int ntheta,
                                    int col_vsini, double* __restrict__ sb) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (k >= n2p1) return;
  double out = 1.0;
  if (k > 0) {
    const double vsini = theta[(long long)b * ntheta + col_vsini];
    const double freq = (double)k * freq_val;
    const double ub = kTwoPi * vsini * freq;
    double s, c;
    sincos(ub, &s, &c);
    const double ub2 = ub * ub;
    out = j1(ub) / ub - 3.0 * c / (2.0 * ub2) + 3.0 * s / (2.0 * (ub2 * ub));
  }
  sb[(long long)b * n2p1 + k] = out;
}

// ------------------------------------------------------------------------------------------------
// (2b) y = irfft(F · sb): real inverse FFT of length nf as ONE complex transform of nf/2 points held in
//      shared memory (even/odd packing).  Transforms longer than kFftMaxPoints are decimated S ways
//      (S a power of two): CTA h produces the samples z[S·j+h] from the S-fold folded input, so no pass
//      through global memory is needed.  Radix-2 decimation in time after a bit-reversed store.
// ------------------------------------------------------------------------------------------------
constexpr int FFT_THREADS = 1024;

__global__ void __launch_bounds__(FFT_THREADS)
broaden_kernel(int nf, int S, int log2np, int R, const double2* __restrict__ F, const double2* __restrict__ T,
               const double* __restrict__ sb, double* __restrict__ y) {
  extern __shared__ double2 sm[];
  const int n2 = nf >> 1, np = n2 / S;
  const int h = blockIdx.x, r = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const double2* Fr = F + (long long)r * (n2 + 1);
  const double* sbb = sb + (long long)b * (n2 + 1);
  for (int kp = tid; kp < np; kp += FFT_THREADS) {
    double2 acc = make_double2(0.0, 0.0);
    for (int q = 0; q < S; ++q) {
      const int k = kp + q * np;
      double2 a = Fr[k];
      const double sa = sbb[k];
      a.x *= sa; a.y *= sa;
      double2 c = Fr[n2 - k];
      const double sc = sbb[n2 - k];
      c.x *= sc; c.y *= -sc;                    // conj(Xs[n2-k])
      if (k == 0) { a.y = 0.0; c.y = 0.0; }     // C2R ignores the imaginary parts of DC and Nyquist
      const double2 E = make_double2(0.5 * (a.x + c.x), 0.5 * (a.y + c.y));
      const double2 Dd = make_double2(0.5 * (a.x - c.x), 0.5 * (a.y - c.y));
      const double2 O = cmul(Dd, T[k]);
      double2 Z = make_double2(E.x - O.y, E.y + O.x);  // E + i·O
      if (S > 1) Z = cmul(Z, twiddle(T, nf, (long long)((h * q) & (S - 1)) * (nf / S)));
      acc.x += Z.x; acc.y += Z.y;
    }
    if (S > 1) acc = cmul(acc, twiddle(T, nf, 2LL * h * kp));
    const int dst = (log2np > 0) ? (int)(__brev((unsigned)kp) >> (32 - log2np)) : 0;
    sm[dst] = acc;
  }
  __syncthreads();
  for (int len = 2; len <= np; len <<= 1) {
    const int half = len >> 1, tstep = nf / len;
    for (int t = tid; t < (np >> 1); t += FFT_THREADS) {
      const int pos = t & (half - 1);
      const int i = ((t - pos) << 1) + pos, j = i + half;
      const double2 w = T[pos * tstep];
      const double2 a = sm[i], bb = cmul(sm[j], w);
      sm[i] = make_double2(a.x + bb.x, a.y + bb.y);
      sm[j] = make_double2(a.x - bb.x, a.y - bb.y);
    }
    __syncthreads();
  }
  const double scale = 1.0 / (double)n2;
  double2* yo = reinterpret_cast<double2*>(y + ((long long)b * R + r) * nf);
  for (int jp = tid; jp < np; jp += FFT_THREADS) {
    const double2 v = sm[jp];
    yo[(long long)S * jp + h] = make_double2(v.x * scale, v.y * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// (3) Doppler shift + quintic-spline resampling onto the data pixels.  A CTA owns a tile of 256 output
//     pixels of one (walker, row): it finds the knot intervals, computes the spline coefficients its
//     pixels need with the banded-inverse filter (fine-grid samples staged in shared memory) and
//     evaluates de Boor's recurrence with the knots scaled by the walker's Doppler factor.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_SPAN = 512;   // coefficients a tile may stage; wider tiles (masked gaps) go pixel by pixel
constexpr int RS_ROWS = 4;     // bulk-flux rows handled together: one read of the filter taps serves all of them

__device__ __forceinline__ double spline_coef_direct(const double* __restrict__ GinvT, const double* __restrict__ yb,
                                                     int nf, int j) {
  double acc = 0.0;
  for (int d = 0; d <= 2 * kSplineW; ++d) {
    const int i = j - kSplineW + d;
    if (i >= 0 && i < nf) acc = fma(GinvT[(long long)d * nf + j], yb[i], acc);
  }
  return acc;
}

__global__ void __launch_bounds__(RS_THREADS)
resample_kernel(int nf, int R, int N, const double* __restrict__ fw, const double* __restrict__ GinvT,
                const double* __restrict__ wave, const double* __restrict__ theta, int ntheta, int col_vz,
                const double* __restrict__ ysrc, long long y_stride_b, double* __restrict__ Y) {
  __shared__ double ys[RS_ROWS][RS_SPAN + 2 * kSplineW];
  __shared__ double cs[RS_ROWS][RS_SPAN];
  __shared__ int s_min[RS_THREADS / 32], s_max[RS_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x * RS_THREADS + tid, r0 = blockIdx.y * RS_ROWS, b = blockIdx.z;
  const int nr = min(RS_ROWS, R - r0);
  double scale = 1.0;
  if (col_vz >= 0) {
    const double vz = theta[(long long)b * ntheta + col_vz];
    scale = sqrt((kC_KMS + vz) / (kC_KMS - vz));
  }
  const bool act = p < N;
  const double x = act ? wave[p] : 0.0;
  const int l = act ? find_interval(fw, nf, x, scale) : -1;
  int lmin = act ? l : 0x7fffffff, lmax = act ? l : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if (lane == 0) { s_min[warp] = lmin; s_max[warp] = lmax; }
  __syncthreads();
  lmin = s_min[0]; lmax = s_max[0];
#pragma unroll
  for (int w = 1; w < RS_THREADS / 32; ++w) { lmin = min(lmin, s_min[w]); lmax = max(lmax, s_max[w]); }
  const int jlo = lmin - kDeg, span = lmax - jlo + 1;
  const double* yb = ysrc + (long long)b * y_stride_b + (long long)r0 * nf;
  double hb[kDeg + 1];
  if (act) bspl6(fw, nf, x, l, scale, hb);
  double out[RS_ROWS];
#pragma unroll
  for (int q = 0; q < RS_ROWS; ++q) out[q] = 0.0;
  if (span <= RS_SPAN) {
    for (int q = 0; q < nr; ++q)
      for (int i = tid; i < span + 2 * kSplineW; i += RS_THREADS) {
        const int gi = jlo - kSplineW + i;
        ys[q][i] = (gi >= 0 && gi < nf) ? yb[(long long)q * nf + gi] : 0.0;
      }
    for (int q = nr; q < RS_ROWS; ++q)
      for (int i = tid; i < span + 2 * kSplineW; i += RS_THREADS) ys[q][i] = 0.0;
    __syncthreads();
    for (int jj = tid; jj < span; jj += RS_THREADS) {
      const double* g = GinvT + (jlo + jj);
      double acc[RS_ROWS];
#pragma unroll
      for (int q = 0; q < RS_ROWS; ++q) acc[q] = 0.0;
#pragma unroll 4
      for (int d = 0; d <= 2 * kSplineW; ++d) {
        const double gv = g[(long long)d * nf];
#pragma unroll
        for (int q = 0; q < RS_ROWS; ++q) acc[q] = fma(gv, ys[q][jj + d], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < RS_ROWS; ++q) cs[q][jj] = acc[q];
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int q = 0; q < RS_ROWS; ++q)
#pragma unroll
        for (int m = 0; m <= kDeg; ++m) out[q] = fma(hb[m], cs[q][l - kDeg - jlo + m], out[q]);
    }
  } else if (act) {
    for (int q = 0; q < nr; ++q)
      for (int m = 0; m <= kDeg; ++m)
        out[q] = fma(hb[m], spline_coef_direct(GinvT, yb + (long long)q * nf, nf, l - kDeg + m), out[q]);
  }
  if (act)
    for (int q = 0; q < nr; ++q) Y[((long long)b * R + r0 + q) * N + p] = out[q];
}

// ------------------------------------------------------------------------------------------------
// (4) per-walker combination, one CTA per walker.
//     p(λ) = chebval(λ/λmax, [1, c1, ...]) multiplies every row; X_m = (E_m p)(S p); flux = Σ w_m X_m + F̄ p;
//     scale = exp(log_scale)·norm, or — without log_scale — ∫data / ∫(flux·norm) (trapezoid) times norm.
// ------------------------------------------------------------------------------------------------
constexpr int CB_THREADS = 512;

__device__ __forceinline__ double chebval1(double x, const double* __restrict__ c, int nc) {
  // numpy.polynomial.chebyshev.chebval with coefficients [1, c[0], ..., c[nc-1]] (Clenshaw)
  const int len = nc + 1;
  auto coef = [&](int i) { return i == 0 ? 1.0 : c[i - 1]; };
  if (len == 1) return 1.0;
  double c0, c1;
  if (len == 2) {
    c0 = coef(0); c1 = coef(1);
  } else {
    const double x2 = 2.0 * x;
    c0 = coef(len - 2); c1 = coef(len - 1);
    for (int i = 3; i <= len; ++i) {
      const double tmp = c0;
      c0 = coef(len - i) - c1;
      c1 = tmp + c1 * x2;
    }
  }
  return c0 + c1 * x;
}

__device__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < CB_THREADS / 32; ++w) s += red[w];
  return s;
}

// A_λ/A_V of Cardelli, Clayton & Mathis (1989) for R_V = 3.1 (the reference calls extinct() with its defaults,
// spectrum_model.py:298-299): a(x) + b(x)/R_V, x = 1e4/λ[Å].  Same Horner evaluation as the oracle restatement.
__device__ __forceinline__ double ccm89_curve(double wave_aa) {
  const double x = 1e4 / wave_aa;
  double a, b;
  if (x < 1.1) {
    const double p = pow(x, 1.61);
    a = 0.574 * p;
    b = -0.527 * p;
  } else if (x < 3.3) {
    const double y = x - 1.82;
    a = ((((((0.32999 * y - 0.77530) * y + 0.01979) * y + 0.72085) * y - 0.02427) * y - 0.50447) * y + 0.17699) * y + 1.0;
    b = ((((((-2.09002 * y + 5.30260) * y - 0.62251) * y - 5.38434) * y + 1.07233) * y + 2.28305) * y + 1.41338) * y;
  } else if (x < 8.0) {
    const double d = x >= 5.9 ? x - 5.9 : 0.0;
    a = 1.752 - 0.316 * x - 0.104 / ((x - 4.67) * (x - 4.67) + 0.341) - 0.04473 * d * d - 0.009779 * d * d * d;
    b = -3.090 + 1.825 * x + 1.206 / ((x - 4.62) * (x - 4.62) + 0.263) + 0.2130 * d * d + 0.1207 * d * d * d;
  } else {
    const double z = x - 8.0;
    a = -1.073 - 0.628 * z + 0.137 * z * z - 0.070 * z * z * z;
    b = 13.670 + 4.257 * z - 0.420 * z * z + 0.374 * z * z * z;
  }
  return a + b / 3.1;
}

__global__ void __launch_bounds__(CB_THREADS)
combine_kernel(int N, int M, int R, int ncheb, int flags, const double* __restrict__ wave,
               const double* __restrict__ wave_max, const double* __restrict__ data_flux,
               const double* __restrict__ Y, const double* __restrict__ mu, const double* __restrict__ theta,
               int ntheta, int D, double* __restrict__ X, double* __restrict__ flux,
               double* __restrict__ log_scale_out) {
  __shared__ double red[CB_THREADS / 32];
  __shared__ double s_scale;
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* th = theta + (long long)b * ntheta;
  const double* cheb = th + D + 4;
  const double* Yb = Y + (long long)b * R * N;
  const double* w = mu + (long long)b * M;
  const double wmax = *wave_max;
  const double norm = (flags & SFB_MODEL_NORM) ? th[D + 3] : 1.0;
  const bool fit = !(flags & SFB_MODEL_LOG_SCALE);
  double* fb = flux + (long long)b * N;

  // one resampled row at pixel p: (Y · extinction) · Chebyshev, in the reference's order (spectrum_model.py:296-304)
  const bool has_av = (flags & SFB_MODEL_AV) != 0;
  const double av = has_av ? th[D + 4 + ncheb] : 0.0;
  auto row = [&](int r, int p, double ext, double pc) {
    double v = Yb[(long long)r * N + p];
    if (has_av) v *= ext;
    if (ncheb > 0) v *= pc;
    return v;
  };
  for (int p = tid; p < N; p += CB_THREADS) {
    const double pc = ncheb > 0 ? chebval1(wave[p] / wmax, cheb, ncheb) : 1.0;
    const double ext = has_av ? exp10(-0.4 * (av * ccm89_curve(wave[p]))) : 1.0;
    const double sd = row(M + 1, p, ext, pc);
    const double mean = row(M, p, ext, pc);
    double f = 0.0;
    for (int m = 0; m < M; ++m) f = fma(w[m], row(m, p, ext, pc) * sd, f);
    fb[p] = f + mean;
  }
  __syncthreads();
  double scale;
  if (fit) {
    double a_ref = 0.0, a_mod = 0.0;
    for (int p = tid; p + 1 < N; p += CB_THREADS) {
      const double d = wave[p + 1] - wave[p];
      a_ref += d * (data_flux[p + 1] + data_flux[p]) / 2.0;
      a_mod += d * (fb[p + 1] * norm + fb[p] * norm) / 2.0;
    }
    a_ref = block_sum(a_ref, red);
    a_mod = block_sum(a_mod, red);
    scale = a_ref / a_mod;
    if (tid == 0) log_scale_out[b] = log(scale);
    scale *= norm;
  } else {
    const double ls = th[D + 2];
    if (tid == 0) log_scale_out[b] = ls;
    scale = exp(ls) * norm;
  }
  if (tid == 0) s_scale = scale;
  __syncthreads();
  scale = s_scale;
  for (int p = tid; p < N; p += CB_THREADS) {
    const double pc = ncheb > 0 ? chebval1(wave[p] / wmax, cheb, ncheb) : 1.0;
    const double ext = has_av ? exp10(-0.4 * (av * ccm89_curve(wave[p]))) : 1.0;
    const double sd = row(M + 1, p, ext, pc);
    for (int m = 0; m < M; ++m) X[((long long)b * M + m) * N + p] = (row(m, p, ext, pc) * sd) * scale;
    fb[p] = fb[p] * scale;
  }
}

__global__ void wave_max_kernel(const double* __restrict__ wave, int N, double* out) {
  __shared__ double red[32];
  double m = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmax(m, wave[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    *out = m;
  }
}

// rows whose emulator weight covariance was not positive definite: info = -1, lnL = NaN
__global__ void merge_status_kernel(const int* __restrict__ status, int* __restrict__ info,
                                    double* __restrict__ lnL, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && status[b] != 0) {
    info[b] = -1;
    lnL[b] = nan("");
  }
}

template <typename T>
cudaError_t upload(T** dst, const T* src_h, size_t count) {
  cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(sizeof(T) * count, 16));
  if (e != cudaSuccess) return e;
  if (src_h) return cudaMemcpy(*dst, src_h, sizeof(T) * count, cudaMemcpyHostToDevice);
  return cudaSuccess;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host helpers (set-up time)
// ------------------------------------------------------------------------------------------------
void host_rfft(int n, const double* x, double* out) {
  // radix-2 complex FFT in extended precision (set-up only): out[2k], out[2k+1] = Re, Im of X[k], k <= n/2
  std::vector<long double> re(n), im(n, 0.0L);
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  for (int i = 0; i < n; ++i) {
    unsigned rv = 0;
    for (int bb = 0; bb < bits; ++bb)
      if (i & (1 << bb)) rv |= 1u << (bits - 1 - bb);
    re[rv] = x[i];
  }
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1;
    for (int pos = 0; pos < half; ++pos) {
      const long double ang = -2.0L * pi * pos / len;
      const long double wr = cosl(ang), wi = sinl(ang);
      for (int i = pos; i < n; i += len) {
        const int j = i + half;
        const long double br = re[j] * wr - im[j] * wi, bi = re[j] * wi + im[j] * wr;
        re[j] = re[i] - br; im[j] = im[i] - bi;
        re[i] += br; im[i] += bi;
      }
    }
  }
  for (int k = 0; k <= n / 2; ++k) {
    out[2 * k] = (double)re[k];
    out[2 * k + 1] = (double)im[k];
  }
}

int host_spline_inverse_band(int nf, const double* fw, int W, double* out) {
  // out[d*nf + j] = (B⁻¹)[j, j-W+d], B[i,c] = B_c(fw[i]) the quintic collocation matrix.
  // Row j of B⁻¹ solves Bᵀg = e_j; because g decays geometrically it is computed on the window
  // [j-2W, j+2W] (truncation 0.43^(2W) relative), by banded elimination without pivoting (B is
  // totally positive).
  if (nf < 2 * (kDeg + 1)) return -1;
  const int bw = kDeg;  // |i - c| <= 5
  std::vector<double> band((size_t)nf * (2 * bw + 1), 0.0);  // band[i*(11) + (c-i+5)] = B[i,c]
  for (int i = 0; i < nf; ++i) {
    const int l = find_interval(fw, nf, fw[i], 1.0);
    double h[kDeg + 1];
    bspl6(fw, nf, fw[i], l, 1.0, h);
    for (int m = 0; m <= kDeg; ++m) {
      const int c = l - kDeg + m;
      const int off = c - i + bw;
      if (off < 0 || off > 2 * bw) {
        if (h[m] != 0.0) return -2;  // bandwidth assumption violated
        continue;
      }
      band[(size_t)i * (2 * bw + 1) + off] = h[m];
    }
  }
  const int WW = 2 * W;
  std::vector<double> Tb, g;
  for (size_t q = 0; q < (size_t)nf * (2 * W + 1); ++q) out[q] = 0.0;
  for (int j = 0; j < nf; ++j) {
    const int a = std::max(0, j - WW), b = std::min(nf, j + WW + 1), n = b - a;
    // T = (Bᵀ)[a:b, a:b] in band storage: Tb[r*11 + (c-r+5)] = T[r,c] = B[a+c, a+r]
    Tb.assign((size_t)n * (2 * bw + 1), 0.0);
    for (int r = 0; r < n; ++r)
      for (int off = 0; off <= 2 * bw; ++off) {
        const int c = r + off - bw;
        if (c < 0 || c >= n) continue;
        // B[a+c, a+r] sits at band[(a+c)*11 + ((a+r)-(a+c)+5)]
        Tb[(size_t)r * (2 * bw + 1) + off] = band[(size_t)(a + c) * (2 * bw + 1) + (r - c + bw)];
      }
    g.assign(n, 0.0);
    g[j - a] = 1.0;
    for (int k = 0; k < n; ++k) {  // forward elimination
      const double piv = Tb[(size_t)k * (2 * bw + 1) + bw];
      if (piv == 0.0) return -3;
      for (int r = k + 1; r < std::min(n, k + bw + 1); ++r) {
        const double f = Tb[(size_t)r * (2 * bw + 1) + (k - r + bw)] / piv;
        if (f == 0.0) continue;
        for (int c = k; c < std::min(n, k + bw + 1); ++c)
          Tb[(size_t)r * (2 * bw + 1) + (c - r + bw)] -= f * Tb[(size_t)k * (2 * bw + 1) + (c - k + bw)];
        g[r] -= f * g[k];
      }
    }
    for (int k = n - 1; k >= 0; --k) {  // back substitution
      double s = g[k];
      for (int c = k + 1; c < std::min(n, k + bw + 1); ++c) s -= Tb[(size_t)k * (2 * bw + 1) + (c - k + bw)] * g[c];
      g[k] = s / Tb[(size_t)k * (2 * bw + 1) + bw];
    }
    for (int d = 0; d <= 2 * W; ++d) {
      const int i = j - W + d;
      if (i >= a && i < b) out[(size_t)d * nf + j] = g[i - a];
    }
  }
  return 0;
}

int host_cholesky_lower(int n, double* a, int lda) {
  // row-oriented (dot-product) Cholesky; lower triangle in place; returns LAPACK-style info
  for (int j = 0; j < n; ++j) {
    double* aj = a + (size_t)j * lda;
    for (int i = 0; i < j; ++i) {  // L[j][i]
      const double* ai = a + (size_t)i * lda;
      double s = aj[i];
      for (int k = 0; k < i; ++k) s -= aj[k] * ai[k];
      aj[i] = s / ai[i];
    }
    double d = aj[j];
    for (int k = 0; k < j; ++k) d -= aj[k] * aj[k];
    if (!(d > 0.0)) return j + 1;
    aj[j] = std::sqrt(d);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// set-up / tear-down
// ------------------------------------------------------------------------------------------------
cudaError_t upstream_init() {
  cudaError_t e = cudaFuncSetAttribute(broaden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(double2) * kFftMaxPoints));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gp_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

void model_free(ModelState* ms) {
  void* ptrs[] = {ms->fw, ms->bulk, ms->F, ms->T, ms->GinvT, ms->gp_grid, ms->gp_var, ms->gp_ls, ms->L, ms->zw,
                  ms->theta, ms->sb, ms->y, ms->Y, ms->mu, ms->wcov, ms->X, ms->A, ms->flux, ms->log_scale,
                  ms->status, ms->wave_max};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  *ms = ModelState();
}

cudaError_t model_setup(ModelState* ms, int N, int M, int Bmax, int nf, const double* fine_wave_h,
                        const double* bulk_h, int G, int D, const double* grid_h, const double* var_h,
                        const double* ls_h, const double* v11_h, const double* what_h, int ncheb_max,
                        int flags, std::string* err) {
  model_free(ms);
  const int R = M + 2, n = M * G, ld = ((n + 31) / 32) * 32, n2 = nf / 2;
  if ((size_t)ld * (M + 1) * sizeof(double) > 200 * 1024) {
    *err = "emulator too large for the shared-memory GP solve (M·G·M·8 B > 200 KB)";
    return cudaErrorInvalidValue;
  }
  ms->nf = nf; ms->R = R; ms->M = M; ms->G = G; ms->D = D; ms->n = n; ms->ld = ld; ms->N = N; ms->Bmax = Bmax;
  ms->ncheb_max = ncheb_max; ms->flags = flags;
  // calculate_dv of the fine grid (Starfish/utils.py:8-22) and numpy.fft.rfftfreq's 1/(n·d)
  double mn = INFINITY;
  for (int i = 0; i + 1 < nf; ++i) mn = std::min(mn, (fine_wave_h[i + 1] - fine_wave_h[i]) / fine_wave_h[i]);
  ms->dv_fine = kC_KMS * mn;
  ms->freq_val = 1.0 / ((double)nf * ms->dv_fine);
  // spectrum of the bulk fluxes, twiddles
  std::vector<double> F((size_t)R * (n2 + 1) * 2), T((size_t)n2 * 2);
  for (int r = 0; r < R; ++r) host_rfft(nf, bulk_h + (size_t)r * nf, F.data() + (size_t)r * (n2 + 1) * 2);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int j = 0; j < n2; ++j) {
    const long double ang = 2.0L * pi * j / nf;
    T[2 * j] = (double)cosl(ang);
    T[2 * j + 1] = (double)sinl(ang);
  }
  // spline: banded inverse of the collocation matrix
  std::vector<double> Gi((size_t)nf * (2 * kSplineW + 1));
  const int rc = host_spline_inverse_band(nf, fine_wave_h, kSplineW, Gi.data());
  if (rc != 0) {
    *err = "spline set-up failed (fine wavelength grid must be strictly increasing)";
    return cudaErrorInvalidValue;
  }
  // emulator: Cholesky of v11 (identity padded to a multiple of 32) and L⁻¹ŵ
  std::vector<double> L((size_t)ld * ld, 0.0), zw(ld, 0.0);
  for (int i = 0; i < ld; ++i) {
    if (i < n)
      std::memcpy(&L[(size_t)i * ld], v11_h + (size_t)i * n, sizeof(double) * (i + 1));
    else
      L[(size_t)i * ld + i] = 1.0;
  }
  if (host_cholesky_lower(n, L.data(), ld) != 0) {
    *err = "emulator v11 is not positive definite";
    return cudaErrorInvalidValue;
  }
  for (int i = 0; i < n; ++i) {
    double s = what_h[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * ld + k] * zw[k];
    zw[i] = s / L[(size_t)i * ld + i];
  }
  // explicit inverse of the triangular factor, row by row: Linv[i][:] = (e_i − Σ_{k<i} L[i][k]·Linv[k][:]) / L[i][i]
  std::vector<double> Linv((size_t)ld * ld, 0.0);
  for (int i = 0; i < ld; ++i) {
    double* out = &Linv[(size_t)i * ld];
    if (i >= n) { out[i] = 1.0; continue; }
    const double* Li = &L[(size_t)i * ld];
    for (int k = 0; k < i; ++k) {
      const double lik = Li[k];
      if (lik == 0.0) continue;
      const double* rk = &Linv[(size_t)k * ld];
      for (int c = 0; c <= k; ++c) out[c] -= lik * rk[c];
    }
    out[i] += 1.0;
    const double d = 1.0 / Li[i];
    for (int c = 0; c <= i; ++c) out[c] *= d;
  }
  L.swap(Linv);
  cudaError_t e;
#define UP(call) if ((e = (call)) != cudaSuccess) { *err = #call; return e; }
  UP(upload(&ms->fw, fine_wave_h, (size_t)nf));
  UP(upload(&ms->bulk, bulk_h, (size_t)R * nf));
  UP(upload(&ms->F, F.data(), F.size()));
  UP(upload(&ms->T, T.data(), T.size()));
  UP(upload(&ms->GinvT, Gi.data(), Gi.size()));
  UP(upload(&ms->gp_grid, grid_h, (size_t)G * D));
  UP(upload(&ms->gp_var, var_h, (size_t)M));
  UP(upload(&ms->gp_ls, ls_h, (size_t)M * D));
  UP(upload(&ms->L, L.data(), L.size()));
  UP(upload(&ms->zw, zw.data(), zw.size()));
  const size_t nth = (size_t)D + 5 + ncheb_max;  // [grid | vsini | vz | log_scale | norm | cheb.. | Av]
  UP(upload(&ms->theta, (const double*)nullptr, (size_t)Bmax * nth));
  if (flags & SFB_MODEL_VSINI) {
    UP(upload(&ms->sb, (const double*)nullptr, (size_t)Bmax * (n2 + 1)));
    UP(upload(&ms->y, (const double*)nullptr, (size_t)Bmax * R * nf));
  }
  UP(upload(&ms->Y, (const double*)nullptr, (size_t)Bmax * R * N));
  UP(upload(&ms->mu, (const double*)nullptr, (size_t)Bmax * M));
  UP(upload(&ms->wcov, (const double*)nullptr, (size_t)Bmax * M * M));
  UP(upload(&ms->X, (const double*)nullptr, (size_t)Bmax * M * N));
  UP(upload(&ms->A, (const double*)nullptr, (size_t)Bmax * M * M));
  UP(upload(&ms->flux, (const double*)nullptr, (size_t)Bmax * N));
  UP(upload(&ms->log_scale, (const double*)nullptr, (size_t)Bmax));
  UP(upload(&ms->status, (const int*)nullptr, (size_t)Bmax));
  UP(upload(&ms->wave_max, (const double*)nullptr, 1));
#undef UP
  return cudaSuccess;
}

cudaError_t launch_wave_max(const double* wave, int N, double* out, cudaStream_t st) {
  wave_max_kernel<<<1, 1024, 0, st>>>(wave, N, out);
  return cudaGetLastError();
}

cudaError_t launch_merge_status(const int* status, int* info, double* lnL, int B, cudaStream_t st) {
  merge_status_kernel<<<(B + 255) / 256, 256, 0, st>>>(status, info, lnL, B);
  return cudaGetLastError();
}

cudaError_t launch_upstream(const ModelState& ms, const UpstreamArgs& a, cudaStream_t st, long long* launches) {
  if (a.B <= 0) return cudaSuccess;
  const int B = a.B, D = ms.D, M = ms.M, R = ms.R, nf = ms.nf, n2 = nf / 2, N = ms.N;
  cudaError_t e;
  wave_max_kernel<<<1, 1024, 0, st>>>(a.wave, N, ms.wave_max);
  ++*launches;
  // (1) emulator
  gp_predict_kernel<<<B, GP_THREADS, sizeof(double) * ((size_t)ms.ld * M + ms.ld), st>>>(
      M, ms.G, D, ms.n, ms.ld, ms.gp_grid, ms.gp_var, ms.gp_ls, ms.L, ms.zw, a.theta, a.ntheta,
      (ms.flags & SFB_MODEL_PAPER_TERM) ? 1 : 0, ms.mu, ms.wcov, a.A, a.status);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*launches;
  // (2) rotational broadening on the fine grid
  const double* ysrc = ms.bulk;
  long long ystride = 0;
  if (ms.flags & SFB_MODEL_VSINI) {
    rot_transfer_kernel<<<dim3((n2 + 1 + 255) / 256, B), 256, 0, st>>>(n2 + 1, ms.freq_val, a.theta, a.ntheta, D,
                                                                     ms.sb);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    int S = 1;
    while (n2 / S > kFftMaxPoints) S *= 2;
    const int np = n2 / S;
    int log2np = 0;
    while ((1 << log2np) < np) ++log2np;
    broaden_kernel<<<dim3(S, R, B), FFT_THREADS, sizeof(double2) * (size_t)np, st>>>(
        nf, S, log2np, R, reinterpret_cast<const double2*>(ms.F), reinterpret_cast<const double2*>(ms.T), ms.sb,
        ms.y);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *launches += 2;
    ysrc = ms.y;
    ystride = (long long)R * nf;
  }
  // (3) Doppler shift + spline resampling onto the data pixels
  resample_kernel<<<dim3((N + RS_THREADS - 1) / RS_THREADS, (R + RS_ROWS - 1) / RS_ROWS, B), RS_THREADS, 0, st>>>(
      nf, R, N, ms.fw, ms.GinvT, a.wave, a.theta, a.ntheta, (ms.flags & SFB_MODEL_VZ) ? D + 1 : -1, ysrc, ystride,
      ms.Y);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  // (4) Chebyshev correction, reconstruction, scaling
  combine_kernel<<<B, CB_THREADS, 0, st>>>(N, M, R, a.ncheb, ms.flags, a.wave, ms.wave_max, a.data_flux, ms.Y, ms.mu,
                                           a.theta, a.ntheta, D, a.X, a.flux, a.log_scale_out);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  *launches += 2;
  return cudaSuccess;
}

}  // namespace sfb
