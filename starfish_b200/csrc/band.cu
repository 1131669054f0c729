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
//   band_build_kernel   S in band storage Sb[i][d] = S[i, i−d], d < WD — same element arithmetic (operation
//                       order, support masks) as cov_build_kernel, 8·N·WD bytes per walker instead of 8·N²
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
  __shared__ double l_amp[kMaxK], l_mu[kMaxK], l_sig[kMaxK];
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
  }
  __syncthreads();
  const int i = blockIdx.x * BB_ROWS + warp;
  if (i >= p.N) return;
  const int WD = p.WD;
  const double wi = p.wave[i];
  const double r0g = 6 * g_ls, sqrt3 = sqrt(3.0);
  double* out = p.Sb + (long long)b * p.strideSb + (long long)i * WD;
  int over = 0;
  // local kernels: this row's metric per kernel, and the overflow test (block wider than the window)
  for (int d = lane; d <= WD; d += 32) {  // d == WD is only a probe; the factorisation needs a band <= WD−2
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
          const double taper = 0.5 + 0.5 * cos(kPi * rv / r0g);
          v += taper * g_amp * (1 + sqrt3 * rv / g_ls) * exp(-sqrt3 * rv / g_ls);
          nz = true;
        }
      }
      double lsum = 0.0;
      bool any = false;
      for (int q = 0; q < nloc; ++q) {
        const double mu = l_mu[q], r0 = 4 * l_sig[q], f = kC_KMS / mu;
        const double mi = f * fabs(wi - mu), mj = f * fabs(wk - mu);
        if (mi <= r0 && mj <= r0) {
          const double rt = fmax(mi, mj);
          const double taper = 0.5 + 0.5 * cos(kPi * rt / r0);
          lsum += taper * l_amp[q] * exp(-0.5 * (mi * mi + mj * mj) / (l_sig[q] * l_sig[q]));
          any = true;
        }
      }
      if (any) {
        v += lsum;
        nz = true;
      }
      if (d == 0) v += p.jitter;
    }
    if (d < WD) out[d] = v;
    if (d >= WD - 1 && nz) over = 1;
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__host__ __device__ constexpr int band_batch(int ER) { return ER < 16 ? 2 * ER : ER; }

// One CTA per walker; thread (warp tr, lane) keeps the window slots of rows [tr·ER, tr·ER+ER) × columns
// {lane + 32·ec} in registers.  ER = rows per thread: 16 → WD·2 threads (8 warps for the 128-pixel window);
// the 160-pixel window uses 10 → 512 threads so that its 10×5 tile fits a 128-register budget.
// MAXNR = right-hand sides the instantiation can carry (M + 1 <= MAXNR).
//
// What bounds this kernel is the shared-memory pipe (every warp re-reads the pivot columns), then the fp64
// pipe; hence: few, fat warps (a warp reads a column once for ER rows), column ownership strided by 32 so
// that a warp's read of a whole column is conflict-free, and two pivots per barrier.
template <int WD, int ER, int MAXNR>
__global__ void __launch_bounds__(WD * 32 / ER, 1)
band_chol_kernel(BandCholParams p) {
  constexpr int EC = WD / 32, NT = WD * 32 / ER, ROWLEN = WD + NRP;
  constexpr int BATCH = band_batch(ER);                   // pivots per staging batch (a multiple of ER)
  constexpr int NE = (WD * MAXNR + NT - 1) / NT;          // right-hand-side registers per thread
  constexpr int NG = (MAXNR * MAXNR + NT - 1) / NT;       // Gram-matrix accumulators per thread
  static_assert(WD % 32 == 0 && WD % ER == 0 && ER % 2 == 0 && BATCH % ER == 0, "window tiling");
  __shared__ double colbuf[2][2][WD];
  __shared__ double zbuf[2][2][NRP];
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

  // ---- initial window: slot (r, c) = element (i = r, k = c), k <= i (the other half is not live yet)
  double a[ER][EC];
#pragma unroll
  for (int er = 0; er < ER; ++er) {
    const int i = tr * ER + er;
#pragma unroll
    for (int ec = 0; ec < EC; ++ec) {
      const int k = lane + 32 * ec;
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
  double gacc[NG];
  int gp_[NG], gq_[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int idx = tid + NT * g;
    const bool ok = idx < NR * NR;
    gp_[g] = ok ? idx / NR : -1;
    gq_[g] = ok ? idx - (idx / NR) * NR : 0;
    gacc[g] = 0.0;
  }
  stage_issue(0, ring);

  double logdet = 0.0;
  int info = 0;
  // log det S = Σ log(pivot): the pivots are multiplied up as mantissa × 2^exponent by warp 0 (a handful of
  // integer/DMUL instructions per pivot instead of a log() on the critical path) and logged once at the end.
  double mant = 1.0;
  long long expo = 0;
  auto track_pivot = [&](double pj, int j) {  // LAPACK-style info, log det as mantissa × 2^exponent
    if (!(pj > 0.0) && info == 0) info = j + 1;
    const int hi = __double2hiint(pj);
    expo += ((hi >> 20) & 0x7ff) - 1022;
    mant *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(pj));
  };
  // register of this thread that holds column-slot c of its row er (c / 32 is uniform over the CTA, so the
  // branch is uniform and the register index compile-time inside each arm)
  auto col_reg = [&](int er_static_row, int ecj, const double (&row)[EC]) -> double {
    (void)er_static_row;
    double x = row[0];
#pragma unroll
    for (int ec = 1; ec < EC; ++ec)
      if (ecj == ec) x = row[ec];
    return x;
  };

  // The pivot loop is unrolled by ER so that "which register row holds index j" is a compile-time fact inside
  // the body (j ≡ u mod ER, ER | WD).  Two pivots per barrier: the owners publish the raw columns j and j+1
  // (both as left by the pivots before j); every thread corrects the entries of column j+1 it needs with
  // pivot j itself (one FMA each) and then applies the rank-2 update to its registers.  1/pivot_j is taken off
  // the critical path: the owner of the diagonal element of the NEXT pair updates it first thing after the
  // barrier (same arithmetic as the bulk update) and publishes its reciprocal for the following step.  Rows
  // j+WD and j+1+WD enter afterwards; row j+WD misses pivot j+1's update, which is exactly zero because the
  // band is at most WD−2 wide.
  if (tid == 0) invbuf[0] = 1.0 / a[0][0];
  for (int j0 = 0, jr0 = 0; j0 < N; j0 += ER, jr0 = (jr0 + ER == WD) ? 0 : jr0 + ER) {
    const int own_warp = jr0 / ER;
    const int jb0 = j0 % BATCH, half = (j0 / BATCH) & 1;
    const double* rowbase = ring + half * (BATCH * ROWLEN) + jb0 * ROWLEN;
#pragma unroll
    for (int u = 0; u < ER; u += 2) {
      const int j = j0 + u, jr = jr0 + u, jr1 = jr + 1;  // ER is even and divides WD: no wrap inside a pair
      if (j >= N) break;
      // double buffer by pair parity (static when a block holds an even number of pairs)
      const int buf = ((ER / 2) % 2 == 0) ? ((u >> 1) & 1) : ((j >> 1) & 1);
      const bool boundary = (u == 0) && (jb0 == 0);
      // staging boundary: this batch's rows (issued one boundary ago) must have landed before the barrier
      if (boundary) cp_async_wait_all();
      // ---- phase A: owners publish columns j, j+1 of the window and rows j, j+1 of the right-hand sides
      {
        const int ec0 = jr >> 5, ec1 = jr1 >> 5;
        if (lane == (jr & 31)) {
#pragma unroll
          for (int er = 0; er < ER; ++er) colbuf[buf][0][tr * ER + er] = col_reg(er, ec0, a[er]);
        }
        if (lane == (jr1 & 31)) {
#pragma unroll
          for (int er = 0; er < ER; ++er) colbuf[buf][1][tr * ER + er] = col_reg(er, ec1, a[er]);
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (rres[e] == jr) zbuf[buf][0][rq[e]] = rv[e];
        if (rres[e] == jr1) zbuf[buf][1][rq[e]] = rv[e];
      }
      __syncthreads();
      // past the barrier nobody reads the other ring half any more (its last reader was the previous
      // pair): start filling it with the next batch
      if (boundary) stage_issue(j + BATCH, ring + (half ^ 1) * (BATCH * ROWLEN));
      // ---- phase C
      const double* cb0 = colbuf[buf][0];
      const double* cb1 = colbuf[buf][1];
      const double inv0 = invbuf[buf];
      const double m = cb0[jr1] * inv0;                 // a(j+1,j)/pivot_j
      const double p1 = fma(-cb0[jr1], m, cb1[jr1]);    // pivot j+1 after pivot j's update
      const double inv1 = 1.0 / p1;
      {
        const int jn = (jr + 2 == WD) ? 0 : jr + 2;
        if (tr == jn / ER && lane == (jn & 31)) {  // first pivot of the next pair
          const double x0 = cb0[jn], x1 = fma(-x0, m, cb1[jn]);
          double d = fma(-x0, x0 * inv0, col_reg(0, jn >> 5, a[(u + 2) % ER]));
          d = fma(-x1, x1 * inv1, d);
          invbuf[buf ^ 1] = 1.0 / d;
        }
      }
      if (tr == 0) {  // warp-uniform bookkeeping
        track_pivot(cb0[jr], j);
        if (j + 1 < N) track_pivot(p1, j + 1);
        if ((j & 510) == 510) {
          const int h2 = __double2hiint(mant);
          expo += ((h2 >> 20) & 0x7ff) - 1022;
          mant = __hiloint2double((h2 & 0x800fffff) | 0x3fe00000, __double2loint(mant));
        }
      }
      double ak0[EC], ak1[EC];
#pragma unroll
      for (int ec = 0; ec < EC; ++ec) {
        const double c0 = cb0[lane + 32 * ec];
        ak0[ec] = c0 * inv0;
        ak1[ec] = fma(-c0, m, cb1[lane + 32 * ec]) * inv1;
      }
#pragma unroll
      for (int er = 0; er < ER; ++er) {
        const double c0 = cb0[tr * ER + er];
        const double ai0 = -c0, ai1 = -fma(-c0, m, cb1[tr * ER + er]);
#pragma unroll
        for (int ec = 0; ec < EC; ++ec) a[er][ec] = fma(ai1, ak1[ec], fma(ai0, ak0[ec], a[er][ec]));
      }
      const double* zb0 = zbuf[buf][0];
      const double* zb1 = zbuf[buf][1];
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (rres[e] >= 0) {
          const double c0 = cb0[rres[e]], c1 = fma(-c0, m, cb1[rres[e]]);
          const double z0 = zb0[rq[e]], z1 = fma(-m, z0, zb1[rq[e]]);
          rv[e] = fma(-c1, z1 * inv1, fma(-c0, z0 * inv0, rv[e]));
        }
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (gp_[g] >= 0) {
          const double z0p = zb0[gp_[g]], z0q = zb0[gq_[g]];
          gacc[g] = fma(z0p * inv0, z0q, gacc[g]);
          gacc[g] = fma(fma(-m, z0p, zb1[gp_[g]]) * inv1, fma(-m, z0q, zb1[gq_[g]]), gacc[g]);
        }
      }
      // ---- rows j+WD and j+1+WD enter the window in the slots of the retiring indices j and j+1
      if (tr == own_warp) {
        const double* row0 = rowbase + u * ROWLEN;
        const double* row1 = row0 + ROWLEN;
#pragma unroll
        for (int ec = 0; ec < EC; ++ec) {
          int t0 = lane + 32 * ec - jr - 1;
          if (t0 < 0) t0 += WD;
          a[u][ec] = row0[WD - 1 - t0];
          int t1 = lane + 32 * ec - jr1 - 1;
          if (t1 < 0) t1 += WD;
          a[u + 1][ec] = row1[WD - 1 - t1];
        }
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (rres[e] == jr) rv[e] = (rowbase + u * ROWLEN)[WD + rq[e]];
        if (rres[e] == jr1) rv[e] = (rowbase + (u + 1) * ROWLEN)[WD + rq[e]];
      }
    }
  }
  if (tid == 0) logdet = log(mant) + (double)expo * 0.6931471805599453;

  // ---- epilogue: lnL = −½ (log det S + log det(I + A·G) + RᵀS⁻¹R − uᵀ(I + A·G)⁻¹A u)
#pragma unroll
  for (int g = 0; g < NG; ++g)
    if (gp_[g] >= 0) gram[tid + NT * g] = gacc[g];
  __syncthreads();
  if (tid == 0) {
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

template <int WD, int ER>
cudaError_t launch_band_t(const BandCholParams& p, int B, cudaStream_t st) {
  const size_t smem = sizeof(double) * 2 * band_batch(ER) * (WD + NRP);
  static bool opted_in = false;  // static + dynamic shared memory exceeds 48 KB for the widest window
  if (!opted_in) {
    cudaError_t e = cudaFuncSetAttribute(band_chol_kernel<WD, ER, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(band_chol_kernel<WD, ER, kMaxM + 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               64 * 1024);
    if (e != cudaSuccess) return e;
    opted_in = true;
  }
  if (p.M + 1 <= 8)
    band_chol_kernel<WD, ER, 8><<<B, WD * 32 / ER, smem, st>>>(p);
  else
    band_chol_kernel<WD, ER, kMaxM + 1><<<B, WD * 32 / ER, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

const int kBandWidths[] = {64, 96, 128, 160};
const int kNumBandWidths = 4;

__global__ void residual_only_kernel(const double* __restrict__ F, const double* __restrict__ data, int N,
                                     double* __restrict__ resid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i < N) resid[(long long)b * N + i] = F[(long long)b * N + i] - data[i];
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

cudaError_t launch_band_chol(const BandCholParams& p, int WD, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  switch (WD) {
    case 64: return launch_band_t<64, 16>(p, B, st);
    case 96: return launch_band_t<96, 16>(p, B, st);
    case 128: return launch_band_t<128, 16>(p, B, st);
    case 160: return launch_band_t<160, 10>(p, B, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sfb
