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

}  // namespace sfb
