// Internal declarations shared by the translation units of libsfb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sfb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsfb200 is written for sm_100a (B200) only"
#endif

struct CUtensorMap_st;  // <cuda.h>

namespace sfb {

constexpr double kC_KMS = 2.99792458e5;  // Starfish/constants.py:7
constexpr int kTile = 128;               // factorisation panel width / tile edge
constexpr int kOuterTiles = 8;           // tile columns per outer block: trailing updates run with K = 1024 (measured: 4 -> 180.5, 6 -> 181.7, 8 -> 182.3 evals/s)
constexpr int kMaxM = 16;                // max eigenspectra handled by the fused build kernel
constexpr int kMaxK = 32;                // max local kernels per walker

// ---------------------------------------------------------------------------------------------
// kernel parameter blocks
// ---------------------------------------------------------------------------------------------
struct BuildParams {
  int N;               // pixels
  int M;               // eigenspectra (0 => no emulator term)
  int Kmax;            // row length of loc (per walker) in kernels
  long long ldc;       // leading dimension of C (elements)
  long long strideC;   // elements between walkers in C
  int padN;            // rows/cols written: N (user buffer) or padded Np (workspace, identity padding)
  int lower_only;      // 1: only tiles on/below the diagonal are produced
  int hyper_stride;    // 1: per-walker glob/nloc/loc rows, 0: one shared row
  int vec2;            // 1: C rows are 16 B aligned (ldc, strideC even and base aligned) -> double2 stores
  int bulk_ok;         // 1: wave/X rows are 16 B aligned -> column slices prefetched with cp.async.bulk (set by the launcher)
  double jitter;
  const double* wave;
  const double* sigma;
  const double* X;     // B×M×N or nullptr
  const double* A;     // B×M×M
  const double* glob;  // B×2
  const int* nloc;     // B
  const double* loc;   // B×Kmax×3
  const int* sorted;   // device flag: 1 when wave is strictly increasing (enables tile rejection)
  double* C;
};

struct CholParams {
  double* W;           // workspace base: slot s at W + s*strideW
  long long strideW;   // Np*Np
  int Np;              // padded size (multiple of kTile), also leading dimension
  int k0;              // first row/col of the current panel
  double* Minv;        // per-slot 128×128 inverse of the diagonal factor (row-major, ld 128)
  double* rhs;         // per-slot residual / solution vector (Np)
  double* zk;          // per-slot 128 doubles: z of the current panel
  double* logdet;      // per-slot accumulators
  double* sqmah;
  int* info;           // per-slot LAPACK-style info
};

// ---------------------------------------------------------------------------------------------
// launchers (defined in the .cu files)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_cov_build(const BuildParams& p, int B, cudaStream_t st);
cudaError_t launch_check_sorted(const double* wave, int N, int* flag, cudaStream_t st);
cudaError_t launch_residual(const double* model_flux, const double* data_flux, int N, int Np, int B,
                            double* rhs, double* resid_out, double* logdet, double* sqmah, int* info,
                            cudaStream_t st);
cudaError_t launch_potrf_diag(const CholParams& p, int B, int last, double* lnL_out, int* info_out,
                              cudaStream_t st);
struct GemmMaps {  // TMA tensor maps over the factorisation workspace and the per-slot L_kk⁻¹ buffers
  CUtensorMap_st* W = nullptr;
  CUtensorMap_st* Minv = nullptr;
};
cudaError_t make_gemm_maps(GemmMaps* out, double* W, int Np, double* Minv, int slots);
void free_gemm_maps(GemmMaps* m);
cudaError_t launch_trsm(const CholParams& p, const GemmMaps& m, int slot0, int B, cudaStream_t st);
cudaError_t launch_syrk_strip(const CholParams& p, const GemmMaps& m, int slot0, int kb, int K, int jt0, int njt,
                              int B, cudaStream_t st);
cudaError_t launch_syrk_tri(const CholParams& p, const GemmMaps& m, int slot0, int kb, int K, int jt0, int B,
                            cudaStream_t st);
cudaError_t launch_copy_in_lower(const double* C, int N, double* W, int Np, long long strideW, int B,
                                 cudaStream_t st);
cudaError_t launch_copy_out_lower(double* C, int N, const double* W, int Np, long long strideW, int B,
                                  cudaStream_t st);
cudaError_t launch_solve_lower(const double* L, long long strideL, int ldl, const double* r, double* z,
                               int N, int B, cudaStream_t st);
cudaError_t kernels_init();  // sets max dynamic shared memory attributes
void potrf_set_blocked(bool on);  // experiments: blocked shared-memory potrf_diag (default) or the one-sweep register kernel

// ---- shared-factor path (frozen kernel groups): one factorisation of S, right-hand sides of all walkers as rows ----
struct FwdMaps {  // TMA tensor maps over the right-hand-side rows and over the per-panel L_kk⁻¹ blocks
  CUtensorMap_st* Z = nullptr;
  CUtensorMap_st* Mall = nullptr;
};
cudaError_t make_fwd_maps(FwdMaps* out, double* Zt, int Np, int Jp, double* MinvAll, int panels);
void free_fwd_maps(FwdMaps* m);
cudaError_t launch_pack_rhs(const double* model_flux, const double* data_flux, const double* X, int N, int Np, int M,
                            int J, int Jp, double* Zt, double* resid_out, cudaStream_t st);
cudaError_t launch_forward_rows(const FwdMaps& fm, const GemmMaps& gm, double* Zt, int Np, int Jp, int slotL,
                                cudaStream_t st, long long* launches);
// Gram matrix of every walker's solved right-hand sides + the M×M capacitance system (band.cu's epilogue):
// lnL_b = −½[logdet S + logdet(I + A_b G_b) + ‖z_R‖² − uᵀ(I + A_b G_b)⁻¹A_b u]
cudaError_t launch_gram_capacitance(const double* Zt, int ldz, int N, int M, int B, const double* A,
                                    const double* logdet_S, const int* info_S, double* lnL, int* info, cudaStream_t st);

// ---- int8 tensor-core trailing update (SFB_SOLVER_DENSE_I8, ozaki.cu) ----
constexpr int kOzSlices = 6;   // balanced radix-256 digits per fp64 operand element (48-bit fixed point per row)
constexpr int kOzChunk = 32;   // k per MMA / pipeline stage (bytes per row per slice)
struct OzParams {
  int8_t* P;             // sliced panels of the current outer block: slot s at P + s*strideP
  long long strideP;     // bytes per slot
  double* rscale;        // per slot, Np doubles: 2^(e_i - 7)
  uint8_t* F;            // per slot: [chunk][row/64] bit t = digit slab t of that 64-row × 32-k block is not all zero
  long long strideF;     // bytes per slot
  unsigned long long* stats;  // optional [2]: int8 MMAs issued / MMAs a dense digit pattern would issue
  int dbg;               // experiments build only: 1 = treat every digit slab as non-zero (timing of the dense schedule)
};
void ozaki_set_debug(int d);
cudaError_t ozaki_init();
void ozaki_set_tpc(int n);   // most tiles per CTA (experiments)
void ozaki_set_pair(bool on);  // CTA pairs sharing the A operand by multicast (default) or independent CTAs (experiments)
size_t oz_panel_bytes_per_slot(int Np, int outer_tiles);
size_t oz_flag_bytes_per_slot(int Np, int outer_tiles);
cudaError_t launch_oz_rowscale(const CholParams& p, const OzParams& oz, int B, cudaStream_t st);
cudaError_t launch_oz_syrk_strip(const CholParams& p, const OzParams& oz, int K, int jt0, int njt, int B,
                                 cudaStream_t st);
cudaError_t launch_oz_syrk_tri(const CholParams& p, const OzParams& oz, int K, int jt0, int B, cudaStream_t st);

#ifdef __CUDACC__
// Fixed-point slices of one 64-row × 128-column block of the panel (rows row_base.., four 32-k chunks ch0..ch0+3) for
// the int8 trailing update (layout and arithmetic: ozaki.cu).  256 threads: a warp takes 8 rows; per row lane l owns
// k = 4l..4l+3 (32 contiguous bytes of `src`, which may be global memory — the panel as trsm left it — or the shared
// memory tile trsm still holds), six packed 4-byte stores.  Per chunk the CTA also records which digit slabs are not
// identically zero (F): |L_ik| is usually far below its row's scale away from the band, so the leading digit slab of
// most operand blocks is all zero and the update kernel skips every copy and product with it.
__device__ __forceinline__ void oz_slice_block(const OzParams& oz, int s, int Np, int row_base, int ch0, const double* src,
                                               long long src_ld, uint32_t (*wmask)[4]) {
  constexpr int OZ_S = kOzSlices, OZ_KC = kOzChunk, OZ_GROUP_BYTES = 8 * kOzChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = ch0 + (lane >> 3);
  const int kb = (4 * lane) & 31;                       // byte inside the 32-byte row
  uint32_t orw[OZ_S];                                   // OR of this lane's packed digits per slab (zero <=> all digits zero)
#pragma unroll
  for (int t = 0; t < OZ_S; ++t) orw[t] = 0u;
  int rs_hi8[8];                                        // exponent words of the eight rows' scales, fetched together
#pragma unroll
  for (int i = 0; i < 8; ++i) rs_hi8[i] = __double2hiint(__ldg(oz.rscale + (long long)s * Np + row_base + warp * 8 + i));
  // The warp's eight rows are ONE 8-row group of P (row_base is a multiple of 64): per slab one pointer, per row an
  // immediate offset.  32-byte swizzle: the 16-byte halves of a row are swapped for rows 4-7 of the group.
  const long long slice_bytes = (long long)Np * OZ_KC;
  int8_t* dst[OZ_S];
  {
    int8_t* d0 = oz.P + (long long)s * oz.strideP + (long long)chunk * OZ_S * slice_bytes +
                 (long long)((row_base + warp * 8) >> 3) * OZ_GROUP_BYTES + (kb & 15) + 16 * ((kb >> 4) & 1);
#pragma unroll
    for (int t = 0; t < OZ_S; ++t) dst[t] = d0 + t * slice_bytes;
  }
  const int flip = 16 - 32 * ((kb >> 4) & 1);           // rows 4-7: the other half (+16 or −16 bytes)
  const double* sp0 = src + (long long)(warp * 8) * src_ld + 4 * lane;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i == 4) {
#pragma unroll
      for (int t = 0; t < OZ_S; ++t) dst[t] += flip;
    }
    const double* sp = sp0 + (long long)i * src_ld;
    const double2 v01 = *reinterpret_cast<const double2*>(sp);
    const double2 v23 = *reinterpret_cast<const double2*>(sp + 2);
    // 2^40 / rscale = 2^(47−e), exact: rscale is a power of two, so only its exponent field is needed
    const double inv = __hiloint2double((2 * 1023 + 40 - ((rs_hi8[i] >> 20) & 0x7ff)) << 20, 0);
    // q = rint(L·2^(47−e)) by the magic-number add: the integer sits in the mantissa of x + 1.5·2^52.
    // Balanced radix-256 digits b_t ∈ [−128, 127] of q = Σ b_t·256^t, all six at once: u = q + Σ 128·256^t has the
    // UNSIGNED bytes b_t + 128 (the representation is unique), and x − 128 = x XOR 0x80 as an int8.  The bias is
    // folded into the magic constant (1.5·2^52 + 0x808080808080 is an integer below 2^53, hence exact), so the low
    // 48 bits of the sum's bit pattern ARE u — no integer arithmetic at all; the XOR is applied to the packed words.
    // Representable range |q| <= 0x7f7f7f7f7f7f, and no clamp: the pivots of an accepted factorisation satisfy
    // Σ_k L_ik² <= C_ii(1+ε), so |L_ik| < 2^(e−1) and |q| < 2^46(1+ε) — a matrix that violates it has been flagged by
    // potrf_diag before this panel is sliced (every kernel returns on info != 0), and whatever bit pattern an
    // out-of-range or non-finite value produces is still six int8 digits (fmin/fmax cost more than everything else here).
    const double magic = 6755399441055744.0 + 141289400074368.0;                // 1.5·2^52 + 0x808080808080
    unsigned long long u[4];
    u[0] = (unsigned long long)__double_as_longlong(fma(v01.x, inv, magic));
    u[1] = (unsigned long long)__double_as_longlong(fma(v01.y, inv, magic));
    u[2] = (unsigned long long)__double_as_longlong(fma(v23.x, inv, magic));
    u[3] = (unsigned long long)__double_as_longlong(fma(v23.y, inv, magic));
#pragma unroll
    for (int t = 0; t < OZ_S; ++t) {       // slab t = digit 5 − t (t = 0 most significant): byte 5 − t of each element
      const int bt = OZ_S - 1 - t;
      uint32_t x0, x1, x2, x3;
      if (bt < 4) { x0 = (uint32_t)u[0]; x1 = (uint32_t)u[1]; x2 = (uint32_t)u[2]; x3 = (uint32_t)u[3]; }
      else { x0 = (uint32_t)(u[0] >> 32); x1 = (uint32_t)(u[1] >> 32); x2 = (uint32_t)(u[2] >> 32); x3 = (uint32_t)(u[3] >> 32); }
      const uint32_t sel = (uint32_t)(bt & 3) | ((uint32_t)(4 + (bt & 3)) << 4);
      const uint32_t w = __byte_perm(__byte_perm(x0, x1, sel), __byte_perm(x2, x3, sel), 0x5410) ^ 0x80808080u;
      orw[t] |= w;
      *reinterpret_cast<uint32_t*>(dst[t] + i * OZ_KC) = w;
    }
  }
  // bit t: slab t of this lane's chunk has a non-zero digit; OR over the 8 lanes of a chunk, then over the CTA's 8 warps
  uint32_t nz = 0;
#pragma unroll
  for (int t = 0; t < OZ_S; ++t) nz |= (orw[t] != 0u) << t;
  nz |= __shfl_xor_sync(0xffffffffu, nz, 1);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 2);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 4);
  if ((lane & 7) == 0) wmask[warp][lane >> 3] = nz;
  __syncthreads();
  if (threadIdx.x < 4) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) m |= wmask[w][threadIdx.x];
    oz.F[(long long)s * oz.strideF + (long long)(ch0 + threadIdx.x) * (Np / 64) + row_base / 64] = (uint8_t)m;
  }
}
#endif
cudaError_t launch_trsm_slice(const CholParams& p, const GemmMaps& m, const OzParams& oz, int chunk0, int slot0, int B,
                              cudaStream_t st);

}  // namespace sfb

// ---------------------------------------------------------------------------------------------
// upstream of the covariance: emulator GP predictive + spectral transforms
// ---------------------------------------------------------------------------------------------
namespace sfb {

constexpr int kSplineW = 48;        // half-width of the banded inverse of the quintic collocation matrix
                                    // (entries decay by 0.4306 per knot: 0.43^48 = 2.7e-18)
constexpr int kFftMaxPoints = 8192; // complex points of one in-shared-memory transform (128 KB)

struct ModelState {
  // sizes
  int nf = 0, R = 0, M = 0, G = 0, D = 0, n = 0, ld = 0, N = 0, Bmax = 0, ncheb_max = 0;
  int flags = 0;
  double dv_fine = 0.0, freq_val = 0.0;
  // static tables (device)
  double* fw = nullptr;       // nf          fine log-λ grid
  double* bulk = nullptr;     // R×nf        bulk fluxes on the fine grid (used when there is no vsini)
  double* F = nullptr;        // R×(nf/2+1)  complex spectrum of the bulk fluxes
  double* T = nullptr;        // nf/2        complex twiddles exp(+2πi j/nf)
  double* GinvT = nullptr;    // (2W+1)×nf   banded inverse of the collocation matrix, tap-major
  double* gp_grid = nullptr;  // G×D
  double* gp_var = nullptr;   // M
  double* gp_ls = nullptr;    // M×D
  double* L = nullptr;        // ld×ld INVERSE of the lower Cholesky factor of v11 (row-major, identity padding)
  double* zw = nullptr;       // ld  L⁻¹·ŵ
  // per-call scratch (device), sized for Bmax walkers
  double* theta = nullptr;    // Bmax×ntheta staging for the host-buffer entry
  double* sb = nullptr;       // Bmax×(nf/2+1) rotational transfer function
  double* y = nullptr;        // Bmax×R×nf   broadened bulk fluxes on the fine grid
  double* Y = nullptr;        // Bmax×R×N    resampled onto the data pixels
  double* mu = nullptr;       // Bmax×M      emulator weights
  double* wcov = nullptr;     // Bmax×M×M    emulator weight covariance Σ_w
  double* X = nullptr;        // Bmax×M×N
  double* A = nullptr;        // Bmax×M×M
  double* flux = nullptr;     // Bmax×N
  double* log_scale = nullptr;  // Bmax
  int* status = nullptr;      // Bmax
  double* wave_max = nullptr; // device scalar: max of the data wavelengths
};

struct UpstreamArgs {
  int B = 0;
  int ntheta = 0;             // row length of theta
  int ncheb = 0;              // Chebyshev coefficients c1..c_ncheb per walker (c0 is pinned to 1)
  const double* theta = nullptr;      // B×ntheta: [grid params (D) | vsini | vz | log_scale | norm | cheb...]
  const double* wave = nullptr;       // N  data wavelengths
  const double* data_flux = nullptr;  // N
  double* X = nullptr;        // outputs
  double* A = nullptr;
  double* flux = nullptr;
  double* log_scale_out = nullptr;
  int* status = nullptr;
};

// pure host helpers (also exported through the C ABI for CPU tests)
void host_rfft(int n, const double* x, double* out_complex);                      // n power of two
int host_spline_inverse_band(int nf, const double* fw, int W, double* out_tapmajor);
int host_cholesky_lower(int n, double* a, int lda);                                 // in place, returns LAPACK info

cudaError_t model_setup(ModelState* ms, int N, int M, int Bmax, int nf, const double* fine_wave_h,
                        const double* bulk_h, int G, int D, const double* grid_h, const double* var_h,
                        const double* ls_h, const double* v11_h, const double* what_h, int ncheb_max,
                        int flags, std::string* err);
void model_free(ModelState* ms);
cudaError_t launch_wave_max(const double* wave, int N, double* out, cudaStream_t st);
cudaError_t launch_upstream(const ModelState& ms, const UpstreamArgs& a, cudaStream_t st, long long* launches);
cudaError_t launch_merge_status(const int* status, int* info, double* lnL, int B, cudaStream_t st);
cudaError_t upstream_init();

}  // namespace sfb

// ---------------------------------------------------------------------------------------------
// structure-exploiting solver: banded Cholesky of S + rank-M capacitance
// ---------------------------------------------------------------------------------------------
namespace sfb {

struct BandBuildParams {
  int N, WD, Kmax, hyper_stride;
  double jitter;
  const double* wave;
  const double* sigma;
  const double* glob;
  const int* nloc;
  const double* loc;
  double* Sb;            // walker b at Sb + b*strideSb, N×WD row-major
  long long strideSb;
  int* overflow;         // per walker: set when the band does not fit in WD (caller zeroes it)
  const int* rowmap;     // optional: CTA y -> walker index (classes of equal window width)
};

struct BandCholParams {
  int N, M;
  const double* Sb;
  long long strideSb;
  const int* rowmap;
  const double* X;           // B×M×N (may be null when M == 0)
  const double* A;           // B×M×M
  const double* model_flux;  // B×N
  const double* data_flux;   // N
  const int* overflow;       // B
  const int* sorted;         // device flag from sfb_set_static
  double* lnL;
  int* info;
};

extern const int kBandWidths[];
extern const int kNumBandWidths;
cudaError_t launch_band_width(int N, int Kmax, int hyper_stride, const double* wave, const double* glob,
                              const int* nloc, const double* loc, int* bw, int B, cudaStream_t st);
cudaError_t launch_band_build(const BandBuildParams& p, int B, cudaStream_t st);
cudaError_t launch_band_chol(const BandCholParams& p, int WD, int B, cudaStream_t st);
int band_slack(int WD);  // b + band_slack(WD) <= WD must hold for a walker to use window WD
cudaError_t launch_residual_only(const double* model_flux, const double* data_flux, int N, int B, double* resid,
                                 cudaStream_t st);

}  // namespace sfb
