// Batched blocked in-place fp64 Cholesky with the forward solve folded in (sm_100a).
//
// Replaces scipy.linalg.cho_factor / cho_solve at Starfish/models/spectrum_model.py:400-404 for a batch
// of walkers.  Right-looking, panel width 128, lower triangle, row-major, leading dimension Np
// (N padded to a multiple of 128 with an identity block so no kernel has ragged edges).  Per panel k:
//
//   potrf_diag  one CTA per walker: factor the 128×128 diagonal tile in shared memory, AND its inverse
//               M = L_kk⁻¹ (kept in the unused upper triangle of the same tile), z_k = M·r_k,
//               logdet += 2Σlog L_ii, sqmah += ‖z_k‖²;  reports LAPACK-style info on a non-positive pivot
//   trsm        L_ik = A_ik·Mᵀ as a DMMA GEMM (64×128 tiles), and r_i −= L_ik·z_k in its epilogue
//   syrk        A_ij −= L_ik·L_jkᵀ for i ≥ j > k, DMMA GEMM on 128×64 tiles — the dominant kernel
//
// so the residual vector rides along as an extra right-hand side: when the last panel is done
// rhs = L⁻¹R and lnL = −(logdet + sqmah)/2 without a separate triangular-solve pass over L.
//
// Tensor cores: fp64 has no tcgen05 kind; the B200 fp64 tensor path is mma.sync m8n8k4 (SASS
// DMMA.8x8x4), measured at 37.1 TFLOP/s = the nominal fp64 peak (tools/fp64_peak.cu) while plain DFMA
// tops out at 33.8 — hence DMMA.  Operands are K-contiguous in memory for both factors (rows of L), so
// tiles go global→shared by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-stage mbarrier ring; a
// row permutation of the fragments makes the swizzled 16-byte loads bank-conflict-free.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "sfb_internal.cuh"

namespace sfb {

namespace {

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) {
  }
}
// Explicit shared-space 16-byte load.  Must be a real LDS: shared loads and mbarrier arrives go through
// the same in-order pipe, which is what makes "arrive on the empty barrier after the last read" safe; a
// generic-address load (which the compiler emits for pointers that went through integer arithmetic) can
// still be in flight when the arrive lands and the slot is refilled by TMA.
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
// TMA: 3-D tiled bulk tensor load global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// NT GEMM core:  acc[r][c] = Σ_k Aop[r][k]·Bop[c][k]   (both operands K-contiguous rows of L or L_kk⁻¹)
//
//   CTA tile BM×BN with BM+BN = 192, 8 warps, warp tile 32×32 = 4×4 DMMA tiles.
//   Operands arrive by TMA: per k-slab of 16 doubles one elected thread issues three 64-row × 128-byte
//   boxes (SWIZZLE_128B) into a 4-deep ring of 24 KB stages; `full` mbarriers count the bytes, `empty`
//   mbarriers count the 8 consumer warps.  No __syncthreads in the loop: warps drift by up to a slab.
//
//   Shared-memory layout of a stage: row r (128 B) of the box at r·128, its 16-byte chunk c at
//   ((c ^ (r & 7)) << 4).  Thread (g = lane/4, t = lane%4) reads, for each of its 8-row tiles, chunk t and
//   chunk t+4 of row ρ(g) = (g>>1)|((g&1)<<2) with one LDS.128 each: within every quarter-warp the two
//   rows differ in bit 2, so the eight 16-byte chunks are distinct -> conflict-free.  A chunk holds
//   k = 2c, 2c+1; DMMA "k-group" (h, e) therefore uses k = 2(t+4h)+e for both operands.
//   Consequence for the accumulators: fragment row g is tile row ρ(g), fragment columns 2t, 2t+1 are
//   tile columns t and t+4.
// ------------------------------------------------------------------------------------------------
constexpr int BK = 16;
constexpr int ROW_BYTES = 128;
constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 192 * ROW_BYTES;  // 24 KB
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;  // ring + mbarriers, no slack

struct GemmOperand {
  const CUtensorMap* map;
  int col0;   // first k column
  int row0;   // first row of the operand tile
  int slot;   // walker slot (3rd tensor coordinate)
};

// TRI = false: every warp consumes `warp_slabs` slabs.  TRI = true (trsm: B = L_kk⁻¹ is lower triangular, so output
// column c only needs k <= c): the warp's four 8-column tiles are columns 16·wn, 16·wn+8 (need slabs 0..wn) and
// 16·(7−wn), 16·(7−wn)+8 (need slabs 0..7−wn) — pairing column group g with 7−g gives every warp the same 9
// half-slabs of work (the contiguous 32-column assignment left the warps with 2, 4, 6 and 8 slabs: the CTA ran at the
// pace of the slowest and the tensor pipe at 62 %).
template <int BM, int BN, bool TRI = false>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[4][4][2], uint8_t* smem_raw, const GemmOperand A,
                                              const GemmOperand B, int K, int warp_slabs) {
  static_assert(BM + BN == 192 && BM % 64 == 0 && BN % 64 == 0, "tile shape");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WN_CNT = BN / 32;
  const int wm = warp / WN_CNT, wn = warp % WN_CNT;
  const int g = lane >> 2, t = lane & 3;
  const int KT = K / BK;

  // SWIZZLE_128B needs the ring 1024-byte aligned.  The kernels have no static shared memory, so the
  // dynamic window starts at the CTA's (1 KB-granular) allocation base; verified rather than padded,
  // because every KB counts for co-residency with the panel kernels.
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[s] at +8s, empty[s] at +8(STAGES+s)

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (STAGES + s), GEMM_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int slab) {
    const int slot = slab % STAGES;
    const uint32_t dst = smem_base + slot * STAGE_BYTES;
    const uint32_t bar = bar_base + 8 * slot;
    mbar_arrive_expect_tx(bar, STAGE_BYTES);
#pragma unroll
    for (int rb = 0; rb < BM / 64; ++rb)
      tma_load_3d(dst + rb * 64 * ROW_BYTES, A.map, bar, A.col0 + slab * BK, A.row0 + rb * 64, A.slot);
#pragma unroll
    for (int rb = 0; rb < BN / 64; ++rb)
      tma_load_3d(dst + (BM + rb * 64) * ROW_BYTES, B.map, bar, B.col0 + slab * BK, B.row0 + rb * 64, B.slot);
  };

  int next_issue = 0;  // only meaningful in thread 0
  if (tid == 0) {
    for (; next_issue < STAGES - 1 && next_issue < KT; ++next_issue) issue(next_issue);
  }
  // slab j may be (re)issued into its slot once slab j-STAGES has been released by all 8 warps
  auto try_issue = [&](bool block) {
    if (next_issue >= KT) return;
    if (next_issue >= STAGES) {
      const int prev = next_issue - STAGES;
      const uint32_t bar = bar_base + 8 * (STAGES + prev % STAGES);
      const uint32_t par = (prev / STAGES) & 1;
      if (block) mbar_wait(bar, par);
      else if (!mbar_test(bar, par)) return;
    }
    issue(next_issue);
    ++next_issue;
  };

  const int rho = (g >> 1) | ((g & 1) << 2);
  const uint32_t a_off = (wm * 32 + rho) * ROW_BYTES;
  uint32_t b_off[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    b_off[i] = TRI ? (BM + (i < 2 ? 16 * wn : 16 * (7 - wn)) + 8 * (i & 1) + rho) * ROW_BYTES
                   : (BM + wn * 32 + 8 * i + rho) * ROW_BYTES;
  const uint32_t x0 = (t ^ rho) << 4, x1 = ((t + 4) ^ rho) << 4;

  for (int it = 0; it < KT; ++it) {
    const int slot = it % STAGES;
    if (tid == 0) {
      while (next_issue <= it) try_issue(true);  // the slab about to be consumed must be in flight
      try_issue(false);
    }
    mbar_wait(bar_base + 8 * slot, (it / STAGES) & 1);
    if (TRI ? (it <= 7 - wn) : (it < warp_slabs)) {
      const uint32_t st = smem_base + slot * STAGE_BYTES;
      const int nt0 = (TRI && it > wn) ? 2 : 0;   // first active column tile of this warp for this slab
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t x = h ? x1 : x0;
        double2 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = lds128(st + a_off + i * 8 * ROW_BYTES + x);
          if (!TRI || i >= 2 || nt0 == 0) b[i] = lds128(st + b_off[i] + x);
        }
        if (!TRI || nt0 == 0) {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
        } else {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 2; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 2; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
        }
        if (h == 0 && tid == 0) try_issue(false);
      }
    }
    // Release the slot.  The fence is load-bearing: without it ptxas hoists the arrive above the last
    // DMMA group (they have no register dependence), i.e. right behind the *issue* of the last LDS, and an
    // arrive can then overtake shared-memory reads that are still queued in the LSU — the producer refills
    // the slot by TMA and those reads return the next slab's data (seen as sporadic 1e-6…1e-1 errors once
    // the workspace no longer fits L2 and warps drift apart).  The CTA-scope fence completes this thread's
    // outstanding loads first; it costs nothing measurable (syrk went from 88 % to 92 % of DMMA peak with
    // the TMA ring including it).
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + slot));
  }
  __syncthreads();  // every warp is past its last shared-memory read: the ring may be reused by the caller
}

// ------------------------------------------------------------------------------------------------
// syrk: A_ij −= Σ_{k in [kb, kb+K)} L_ik·L_jkᵀ on 128×64 tiles.  Two grid shapes:
//   strip    (update of a few whole tile columns jt0.. by the K columns [kb, kb+K)): grid (2·ncols, rows, B)
//   triangle (trailing update of every tile column >= jt0):  grid (T(T+1), 1, B), T = nt − jt0, the
//            linear block index is decoded to (row tile, 64-column block) so no CTA is launched for the
//            upper triangle
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    syrk_kernel(const __grid_constant__ CUtensorMap tmW, CholParams p, int slot0, int kb, int K, int jt0, int strip) {
  const int s = blockIdx.z;
  if (p.info[s] != 0) return;
  int it, c0;  // row tile (absolute), first column (absolute)
  if (strip) {
    const int jt = jt0 + (blockIdx.x >> 1);
    it = jt0 + blockIdx.y;
    if (it < jt) return;  // above the diagonal of this tile column
    c0 = jt * kTile + (blockIdx.x & 1) * 64;
  } else {
    const int L = blockIdx.x;
    int t = (int)((sqrtf(4.0f * (float)L + 1.0f) - 1.0f) * 0.5f);
    while (t * (t + 1) > L) --t;
    while ((t + 1) * (t + 2) <= L) ++t;
    it = jt0 + t;
    c0 = jt0 * kTile + (L - t * (t + 1)) * 64;
  }
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  double* Wm = p.W + (long long)s * p.strideW;
  const long long ld = p.Np;
  const int r0 = it * kTile;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const GemmOperand opA{&tmW, kb, r0, slot0 + s}, opB{&tmW, kb, c0, slot0 + s};
  gemm_mainloop<128, 64>(acc, smem_raw, opA, opB, K, K / BK);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / 2, wn = warp % 2, g = lane >> 2, t = lane & 3;
  const int rho = (g >> 1) | ((g & 1) << 2);
  double* Cg = Wm + (long long)(r0 + wm * 32 + rho) * ld + c0 + wn * 32 + t;
  double cv[4][4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      cv[mt][nt][0] = Cg[(long long)mt * 8 * ld + nt * 8];
      cv[mt][nt][1] = Cg[(long long)mt * 8 * ld + nt * 8 + 4];
    }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      Cg[(long long)mt * 8 * ld + nt * 8] = cv[mt][nt][0] - acc[mt][nt][0];
      Cg[(long long)mt * 8 * ld + nt * 8 + 4] = cv[mt][nt][1] - acc[mt][nt][1];
    }
}

// ------------------------------------------------------------------------------------------------
// trsm: L_ik = A_ik·Mᵀ (M = L_kk⁻¹, lower), tile 64 rows × 128 cols (whole panel width, so the in-place
// overwrite is private to the CTA), then rhs_i −= L_ik·z_k.   grid (rows/64, B).
// M is lower triangular: output columns [32·wn, 32·wn+32) only need k < 32·(wn+1), so each warp skips the
// k-slabs beyond its columns.
// ------------------------------------------------------------------------------------------------
// SLICE (int8 trailing update): the CTA also writes the fixed-point slices and digit-slab flags of its 64 × 128
// block (oz_slice_block) from a shared-memory copy of the tile, so the panel is not read back from global memory by
// a separate kernel.
constexpr int TRSM_TILE_LD = 128 + 4;            // doubles per row of the staged tile (row stride = 8 words mod 32 banks)
constexpr int TRSM_TILE_OFF = 2048;              // behind the 4 × 64 partial sums of the rhs reduction
static_assert(TRSM_TILE_OFF + 64 * TRSM_TILE_LD * 8 <= STAGES * STAGE_BYTES, "staged tile must fit the idle ring");
template <bool SLICE>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    trsm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmM, CholParams p,
                int slot0, OzParams oz, int ch0) {
  const int rb = blockIdx.x, s = blockIdx.y;
  if (p.info[s] != 0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t wmask[8][4];
  double* Wm = p.W + (long long)s * p.strideW;
  const long long ld = p.Np;
  const int r0 = p.k0 + kTile + rb * 64;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / 4, wn = warp % 4, g = lane >> 2, t = lane & 3;
  const GemmOperand opA{&tmW, p.k0, r0, slot0 + s}, opB{&tmM, 0, 0, slot0 + s};
  gemm_mainloop<64, 128, true>(acc, smem_raw, opA, opB, kTile, 0);

  const int rho = (g >> 1) | ((g & 1) << 2);
  double* Cg = Wm + (long long)(r0 + wm * 32 + rho) * ld + p.k0 + t;
  const double* zk = p.zk + (long long)s * kTile + t;
  double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int cb = (nt < 2 ? 16 * wn : 16 * (7 - wn)) + 8 * (nt & 1);   // first column of this warp's tile nt
    const double z0 = zk[cb], z1 = zk[cb + 4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      Cg[(long long)mt * 8 * ld + cb] = acc[mt][nt][0];
      Cg[(long long)mt * 8 * ld + cb + 4] = acc[mt][nt][1];
      part[mt] = fma(acc[mt][nt][0], z0, part[mt]);
      part[mt] = fma(acc[mt][nt][1], z1, part[mt]);
      if (SLICE) {
        double* tl = reinterpret_cast<double*>(smem_raw + TRSM_TILE_OFF) + (wm * 32 + mt * 8 + rho) * TRSM_TILE_LD + cb + t;
        tl[0] = acc[mt][nt][0];
        tl[4] = acc[mt][nt][1];
      }
    }
  }
  // reduce over the 4 lanes sharing a row, then over the 4 column-warps in a fixed order
  double* red = reinterpret_cast<double*>(smem_raw);  // [4 wn][64 rows], the ring is idle now
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    double v = part[mt];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) red[wn * 64 + wm * 32 + mt * 8 + rho] = v;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int r = threadIdx.x;
    double sum = ((red[r] + red[64 + r]) + red[128 + r]) + red[192 + r];
    p.rhs[(long long)s * p.Np + r0 + r] -= sum;
  }
  if (SLICE)  // the barrier above also covers the staged tile
    oz_slice_block(oz, s, p.Np, r0, ch0, reinterpret_cast<const double*>(smem_raw + TRSM_TILE_OFF), TRSM_TILE_LD, wmask);
}

// ------------------------------------------------------------------------------------------------
// Shared-factor path (frozen kernel groups, Starfish/models/spectrum_model.py:341-363): every walker has the same
// S = diag + K_global + ΣK_local, so S is factorised ONCE and the walkers only differ in their right-hand sides
// [R_b | X_bᵀ].  Zt holds all of them as rows (row j = one right-hand side, Np contiguous doubles); the forward
// substitution Z = L⁻¹·RHS runs panel by panel with the two GEMM shapes of the factorisation itself:
//   fwd_diag    Zt[j][k0..k0+127] = Σ_c Zt[j][k0+c]·M_k[r][c]       (M_k = L_kk⁻¹ kept per panel)   == trsm_kernel
//   fwd_update  Zt[j][i] −= Σ_c Zt[j][k0+c]·L[i][k0+c]   for i ≥ k0+128                           == syrk_kernel
// Both operands are K-contiguous rows again (right-hand sides are stored as rows for exactly that reason).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
    fwd_diag_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmMall, double* Zt,
                    int ldz, int k0, int panel) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int j0 = blockIdx.x * 64;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / 4, wn = warp % 4, g = lane >> 2, t = lane & 3;
  const GemmOperand opA{&tmZ, k0, j0, 0}, opB{&tmMall, 0, 0, panel};
  gemm_mainloop<64, 128, true>(acc, smem_raw, opA, opB, kTile, 0);  // M_k is lower triangular
  const int rho = (g >> 1) | ((g & 1) << 2);
  double* Cg = Zt + (long long)(j0 + wm * 32 + rho) * ldz + k0 + t;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int cb = (nt < 2 ? 16 * wn : 16 * (7 - wn)) + 8 * (nt & 1);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      Cg[(long long)mt * 8 * ldz + cb] = acc[mt][nt][0];
      Cg[(long long)mt * 8 * ldz + cb + 4] = acc[mt][nt][1];
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 2)
    fwd_update_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW, double* Zt, int ldz,
                      int k0, int slotL) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int j0 = blockIdx.x * 128;
  const int i0 = k0 + kTile + blockIdx.y * 64;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const GemmOperand opA{&tmZ, k0, j0, 0}, opB{&tmW, k0, i0, slotL};
  gemm_mainloop<128, 64>(acc, smem_raw, opA, opB, kTile, kTile / BK);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / 2, wn = warp % 2, g = lane >> 2, t = lane & 3;
  const int rho = (g >> 1) | ((g & 1) << 2);
  double* Cg = Zt + (long long)(j0 + wm * 32 + rho) * ldz + i0 + wn * 32 + t;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      Cg[(long long)mt * 8 * ldz + nt * 8] -= acc[mt][nt][0];
      Cg[(long long)mt * 8 * ldz + nt * 8 + 4] -= acc[mt][nt][1];
    }
}

// Right-hand sides as rows: row b·(M+1) = model_flux_b − data_flux (also written to resid_out), rows b·(M+1)+1+m =
// X_b[m]; columns N..Np−1 and rows J..Jp−1 are zero (the identity padding of the factor leaves them zero).
__global__ void pack_rhs_kernel(const double* __restrict__ model_flux, const double* __restrict__ data_flux,
                                const double* __restrict__ X, int N, int Np, int M, int J, double* Zt, double* resid_out) {
  const int j = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  double v = 0.0;
  if (j < J && i < N) {
    const int b = j / (M + 1), m = j % (M + 1);
    if (m == 0) {
      v = model_flux[(long long)b * N + i] - data_flux[i];
      if (resid_out) resid_out[(long long)b * N + i] = v;
    } else {
      v = X[((long long)b * M + (m - 1)) * N + i];
    }
  }
  Zt[(long long)j * Np + i] = v;
}

// ------------------------------------------------------------------------------------------------
// potrf_diag: factor + invert the diagonal tile, solve the panel's slice of the right-hand side
// ------------------------------------------------------------------------------------------------
constexpr int PD_THREADS = 512;

// Register-blocked: the 128×128 tile lives in REGISTERS, 32 elements per thread — thread (ty = tid/16,
// tx = tid%16) owns S[ty+32a][tx+16b], a<4, b<8.  The lower triangle holds A→L, the strict upper triangle
// holds (L⁻¹)ᵀ under construction (S[r][c], c>r, is T[c][r] of the forward substitution L·M = I), so factor
// and inverse come out of ONE sweep over the columns: per column j its 32 owners publish column j of S
// (both halves) through a double-buffered 1 KB shared buffer, one __syncthreads, and every thread updates
// its 32 registers with at most 12 shared loads.  ≈13 µs per tile instead of 335 µs for the previous
// shared-memory-resident version.
__global__ void __launch_bounds__(PD_THREADS, 1) potrf_diag_kernel(CholParams p, int last, double* lnL_out,
                                                                    int* info_out) {
  const int s = blockIdx.x, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  __shared__ double colbuf[2][kTile];
  __shared__ double ddiag[kTile];
  __shared__ double rk[kTile];
  __shared__ double zpart[PD_THREADS / 32][kTile];
  __shared__ double red[8];
  __shared__ int fail_col;

  if (p.info[s] != 0) {
    if (last && tid == 0) {
      if (lnL_out) lnL_out[s] = __longlong_as_double(0x7ff8000000000000LL);
      if (info_out) info_out[s] = p.info[s];
    }
    return;
  }
  double* Wm = p.W + (long long)s * p.strideW;
  const long long ld = p.Np;
  double* Ag = Wm + (long long)p.k0 * ld + p.k0;

  double v[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = ty + 32 * a, c = tx + 16 * b;
      v[a][b] = (c <= r) ? Ag[(long long)r * ld + c] : 0.0;
    }
  if (tid < kTile) rk[tid] = p.rhs[(long long)s * p.Np + p.k0 + tid];
  if (tid == 0) fail_col = -1;
  // (visibility of rk / fail_col is covered by the first __syncthreads of the sweep)

  bool failed = false;
#pragma unroll
  for (int b0 = 0; b0 < 8; ++b0) {
    if (failed) break;
    for (int jj = 0; jj < 16; ++jj) {
      const int j = 16 * b0 + jj;
      double* cb = colbuf[j & 1];
      if (tx == jj) {
#pragma unroll
        for (int a = 0; a < 4; ++a) cb[ty + 32 * a] = v[a][b0];
      }
      __syncthreads();
      const double piv = cb[j];
      if (!(piv > 0.0) || isinf(piv)) {  // also catches NaN; uniform across the CTA
        if (tid == 0) fail_col = j;
        failed = true;
        break;
      }
      const double d = sqrt(piv);
      const double rinv = 1.0 / d;
      if (tid == 0) ddiag[j] = d;
      double cr[4], cc[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) cr[a] = cb[ty + 32 * a] * rinv;
#pragma unroll
      for (int b = 0; b < 8; ++b) cc[b] = (b >= b0) ? cb[tx + 16 * b] * rinv : 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int r = ty + 32 * a;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (b < b0) continue;
          const int c = tx + 16 * b;
          const bool colok = (b > b0) || (tx > jj);      // c > j
          if (b == b0 && tx == jj) {                       // column j itself: scaled L column / M row, diagonal
            v[a][b] = (r == j) ? d : cr[a];
          } else if (colok) {
            if (r == j) v[a][b] = -cc[b] * rinv;           // T[c][j] = -L[c][j]/d
            else if (r >= c || r < j) v[a][b] = fma(-cr[a], cc[b], v[a][b]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (fail_col >= 0) {
    if (tid == 0) {
      const int code = p.k0 + fail_col + 1;
      p.info[s] = code;
      if (last) {
        if (lnL_out) lnL_out[s] = __longlong_as_double(0x7ff8000000000000LL);
        if (info_out) info_out[s] = code;
      }
    }
    return;
  }

  // write L (lower incl. diagonal) back; M = L⁻¹ to the per-slot buffer (row-major, ld 128): the strict upper
  // element S[r][c] is M[c][r]; the owner of the lower element (r,c) also zeroes M[c][r]'s mirror M[c'][r'] above
  // the diagonal.  z_k = M·r_k is accumulated from the same registers.
  double* Mg = p.Minv + (long long)s * kTile * kTile;
  double zacc[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) zacc[b] = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = ty + 32 * a;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int c = tx + 16 * b;
      if (c < r) {
        Ag[(long long)r * ld + c] = v[a][b];
        Mg[c * kTile + r] = 0.0;                 // M is lower triangular
      } else if (c == r) {
        Ag[(long long)r * ld + c] = v[a][b];
        const double mi = 1.0 / v[a][b];
        Mg[r * kTile + r] = mi;
        zacc[b] = fma(mi, rk[r], zacc[b]);
      } else {
        Mg[c * kTile + r] = v[a][b];             // M[c][r]
        zacc[b] = fma(v[a][b], rk[r], zacc[b]);  // contributes to z[c]
      }
    }
  }
  // z[c] = Σ over rows r: reduce the two ty values inside the warp, then the 16 warps in fixed order
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    double zb = zacc[b] + __shfl_xor_sync(0xffffffffu, zacc[b], 16);
    if (lane < 16) zpart[warp][tx + 16 * b] = zb;
  }
  __syncthreads();
  double zz = 0.0, lg = 0.0;
  if (tid < kTile) {
    double z = 0.0;
#pragma unroll
    for (int w = 0; w < PD_THREADS / 32; ++w) z += zpart[w][tid];
    zz = z * z;
    lg = log(ddiag[tid]);
    p.zk[(long long)s * kTile + tid] = z;
    p.rhs[(long long)s * p.Np + p.k0 + tid] = z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    zz += __shfl_xor_sync(0xffffffffu, zz, o);
    lg += __shfl_xor_sync(0xffffffffu, lg, o);
  }
  if (tid < kTile && lane == 0) {
    red[warp] = zz;
    red[4 + warp] = lg;
  }
  __syncthreads();
  if (tid == 0) {
    const double zsum = ((red[0] + red[1]) + red[2]) + red[3];
    const double lsum = ((red[4] + red[5]) + red[6]) + red[7];
    const double sq = p.sqmah[s] + zsum;
    const double ldt = p.logdet[s] + 2.0 * lsum;
    p.sqmah[s] = sq;
    p.logdet[s] = ldt;
    if (last) {
      if (lnL_out) lnL_out[s] = -(ldt + sq) / 2;
      if (info_out) info_out[s] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// potrf_diag, blocked (the default): the tile lives in SHARED memory and is factorised right-looking in 32-column
// blocks — the only sequential part left is the 32-column sweep of a 32×32 diagonal block, done by ONE warp in
// registers with shuffles (no barrier per column); panel substitution, trailing update and the assembly of the
// explicit inverse are 4×4-register-tile GEMM sweeps by the whole CTA.  The register-resident one-sweep kernel above
// spends ≈ 2000 cycles per column on a barrier, a square root, a division and a full-square predicated update
// (131 µs per tile); here the column chain is shuffle → rsqrt → 31−c shuffled FMAs.
//   T (128 × P2_LD doubles): lower triangle A → L; the strict upper triangle receives (L⁻¹)ᵀ (M[r][c], r > c, at
//   T[c][r]) as in the one-sweep kernel, the diagonal of L⁻¹ is dinv.
// ------------------------------------------------------------------------------------------------
constexpr int P2_THREADS = 256;
constexpr int P2_LD = 129;
constexpr int P2_SMEM_DOUBLES = kTile * P2_LD + 3 * kTile + 64 * 65;
constexpr size_t kPotrf2Smem = sizeof(double) * P2_SMEM_DOUBLES;

// acc[4][4] += Σ_{q<K} A[(4·ty+a)·lda + q] · B[(4·tx+b)·ldb + q]   (both operands contiguous in q)
__device__ __forceinline__ void nt_tile_4x4(double (&acc)[4][4], const double* __restrict__ A, int lda,
                                            const double* __restrict__ B, int ldb, int K) {
#pragma unroll 4
  for (int q = 0; q < K; ++q) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = A[i * lda + q];
      b[i] = B[i * ldb + q];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

#ifdef SFB_EXPERIMENTS   // phase timing of the first tile of a factorisation (experiments build only)
#define P2_MARK(i) do { if (tid == 0) tph[i] = clock64(); } while (0)
#else
#define P2_MARK(i) do { } while (0)
#endif

__global__ void __launch_bounds__(P2_THREADS, 1) potrf_diag2_kernel(CholParams p, int last, double* lnL_out,
                                                                     int* info_out) {
#ifdef SFB_EXPERIMENTS
  long long tph[16];
#endif
  extern __shared__ __align__(16) double p2_smem[];
  double* T = p2_smem;                   // [128][P2_LD]
  double* rk = T + kTile * P2_LD;        // right-hand-side slice of this panel
  double* dinv = rk + kTile;             // 1 / L_jj
  double* dg = dinv + kTile;             // L_jj
  double* Wb = dg + kTile;               // scratch: column buffers of S1, then [64][65] for the inverse assembly
  __shared__ int fail_col;
  __shared__ double red[16];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (p.info[s] != 0) {
    if (last && tid == 0) {
      if (lnL_out) lnL_out[s] = __longlong_as_double(0x7ff8000000000000LL);
      if (info_out) info_out[s] = p.info[s];
    }
    return;
  }
  double* Wm = p.W + (long long)s * p.strideW;
  const long long ld = p.Np;
  double* Ag = Wm + (long long)p.k0 * ld + p.k0;

  P2_MARK(0);
  // thread -> (column pair, rows r0, r0+2, ...): 16 independent 16-byte loads in flight per batch
  {
    const int c2 = (tid & 63) * 2, rq = tid >> 6;          // columns c2, c2+1; rows rq, rq+4, ...
#pragma unroll
    for (int batch = 0; batch < 2; ++batch) {
      double2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = rq + 4 * (batch * 16 + i);
        v[i] = (c2 <= r) ? *reinterpret_cast<const double2*>(Ag + (long long)r * ld + c2) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = rq + 4 * (batch * 16 + i);
        T[r * P2_LD + c2] = v[i].x;
        T[r * P2_LD + c2 + 1] = (c2 + 1 <= r) ? v[i].y : 0.0;
      }
    }
  }
  if (tid < kTile) rk[tid] = p.rhs[(long long)s * p.Np + p.k0 + tid];
  if (tid == 0) fail_col = -1;
  __syncthreads();
  P2_MARK(1);

  for (int bj = 0; bj < 4; ++bj) {
    const int j0 = 32 * bj;
    // ---- S1: the 32×32 diagonal block, one warp, row `lane` in registers, columns fully unrolled
    if (warp == 0) {
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = T[(j0 + lane) * P2_LD + j0 + c];
      int bad = -1;
      double* colb = Wb;   // two 32-double column buffers (parity of c): column c of L_d, published for the warp
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double piv = __shfl_sync(0xffffffffu, a[c], c);
        if (bad < 0 && (!(piv > 0.0) || isinf(piv))) bad = c;   // also catches NaN; uniform across the warp
        const double rs = rsqrt(piv);                            // 1/√piv for the scaling (one MUFU chain), and
        double d = piv * rs;                                     // √piv from it with one Newton correction
        d = fma(0.5 * rs, fma(-d, d, piv), d);
        const double lc = (lane == c) ? d : a[c] * rs;           // column c of L (rows >= c meaningful)
        a[c] = lc;
        if (lane == c) {
          dinv[j0 + c] = rs;
          dg[j0 + c] = d;
        }
        // L[c2][c] for every c2 > c: one shared-memory broadcast read each (the shuffle version issued 62 SHFL per
        // column and was bound by their throughput: 312 cycles per column)
        double* cb = colb + 32 * (c & 1);
        cb[lane] = lc;
        __syncwarp();
#pragma unroll
        for (int c2 = c + 1; c2 < 32; ++c2) a[c2] = fma(-lc, cb[c2], a[c2]);   // meaningful for lane >= c2
      }
      if (bad >= 0) {
        if (lane == 0) fail_col = j0 + bad;
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c <= lane) T[(j0 + lane) * P2_LD + j0 + c] = a[c];
      }
    }
    __syncthreads();
    if (bj == 0) P2_MARK(2);
    if (fail_col >= 0) break;
    const int nrows = kTile - j0 - 32;  // rows below the diagonal block
    // ---- S2: L[i][j0..j0+31] for the rows below, one thread per row (forward substitution against L_d)
    if (tid < nrows) {
      double* row = T + (j0 + 32 + tid) * P2_LD + j0;
      double x[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = row[c];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        double v = x[c];
#pragma unroll
        for (int q = 0; q < c; ++q) v = fma(-x[q], T[(j0 + c) * P2_LD + j0 + q], v);   // broadcast reads of L_d
        x[c] = v * dinv[j0 + c];
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) row[c] = x[c];
    }
    __syncthreads();
    if (bj == 0) P2_MARK(3);
    // ---- S3: trailing update T[i][k] −= Σ_c L[i][j0+c]·L[k][j0+c] on the 32×32 blocks (bi, bk), bi >= bk > bj.
    // 64 threads per block, each a 4×4 register tile with rows ty+8i and columns tx+8j: a warp's loads then touch
    // consecutive rows of T (stride 129 doubles -> distinct banks); 4 consecutive rows per thread were 8-way conflicts.
    {
      const int nb = 3 - bj;                       // trailing blocks per side
      const int nblk = nb * (nb + 1) / 2;
      const int grp = tid >> 6, t64 = tid & 63, ty = t64 >> 3, tx = t64 & 7;
      for (int bl = grp; bl < nblk; bl += 4) {
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= bl) ++bi;
        const int bk = bl - bi * (bi + 1) / 2;
        const int i0 = j0 + 32 * (bi + 1), k0 = j0 + 32 * (bk + 1);
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        nt_tile_4x4(acc, T + (i0 + ty) * P2_LD + j0, 8 * P2_LD, T + (k0 + tx) * P2_LD + j0, 8 * P2_LD, 32);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k0 + tx + 8 * j <= i0 + ty + 8 * i) T[(i0 + ty + 8 * i) * P2_LD + k0 + tx + 8 * j] -= acc[i][j];
      }
    }
    __syncthreads();
    if (bj == 0) P2_MARK(4);
  }
  P2_MARK(5);
  if (fail_col >= 0) {
    if (tid == 0) {
      const int code = p.k0 + fail_col + 1;
      p.info[s] = code;
      if (last) {
        if (lnL_out) lnL_out[s] = __longlong_as_double(0x7ff8000000000000LL);
        if (info_out) info_out[s] = code;
      }
    }
    return;
  }

  // ---- inverse, diagonal blocks: warp b inverts L_d of block b (row `lane` of X = L_d⁻¹ in registers):
  // X[l][c] = −(Σ_{m=c+1..l} X[l][m]·L[m][c])·dinv[c] for c = l−1 … 0, X[l][l] = dinv[l]; stored transposed in the
  // strict upper triangle of T
  if (warp < 4) {
    const int j0 = 32 * warp;
    double X[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) X[c] = 0.0;
#pragma unroll
    for (int c = 31; c >= 0; --c) {
      double sum = 0.0;
#pragma unroll
      for (int m = c + 1; m < 32; ++m)
        sum = fma(X[m], T[(j0 + m) * P2_LD + j0 + c], sum);     // L[m][c]: broadcast read; X[l][m] is zero for m > l
      const double di = dinv[j0 + c];
      X[c] = (lane == c) ? di : ((lane > c) ? -sum * di : 0.0);
    }
    __syncwarp();   // every lane has read L_d's column entries it needs before the transposed stores below
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < lane) T[(j0 + c) * P2_LD + j0 + lane] = X[c];
  }
  __syncthreads();
  P2_MARK(6);
  // ---- inverse, off-diagonal part, by recursive halving (every level uses the whole CTA):
  //   L = [[L1, 0], [B, L2]]  =>  L⁻¹ = [[M1, 0], [−M2·B·M1, M2]]
  // level 1: inside each 64×64 half, the 32×32 block below the diagonal (two halves side by side, 128 threads each);
  // level 2: the 64×64 block M[64..127][0..63] = −M2·(B·M1), all 256 threads, 4×4 register tiles.
  // Register tiles are interleaved (rows ty + 8i or 16i, columns tx + 16j) so that a warp's shared-memory loads
  // touch consecutive rows / words: no bank conflicts.  M[r][c] (r > c) lives at T[c][r]; its diagonal is dinv.
  {
    const int h = tid >> 7, t128 = tid & 127, ty = t128 >> 4, tx = t128 & 15;
    const int j0 = 64 * h, i0 = j0 + 32;
    double* Wh = Wb + h * (32 * 33);
    double acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {            // W = L[i0.., j0..]·M[j0.., j0..]   (M lower triangular)
      double a[4], bb[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = T[(i0 + ty + 8 * i) * P2_LD + j0 + q];
      const double dq = dinv[j0 + q];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int y = tx + 16 * j;
        const double v = T[(j0 + y) * P2_LD + j0 + q];
        bb[j] = (q > y) ? v : ((q == y) ? dq : 0.0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fma(a[i], bb[0], acc[i][0]);
        acc[i][1] = fma(a[i], bb[1], acc[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Wh[(ty + 8 * i) * 33 + tx] = acc[i][0];
      Wh[(ty + 8 * i) * 33 + tx + 16] = acc[i][1];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {            // M[i0+x][j0+y] = −Σ_{q<=x} M[i0+x][i0+q]·W[q][y]
      double a[4];
      const double dq = dinv[i0 + q];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = ty + 8 * i;
        const double v = T[(i0 + q) * P2_LD + i0 + x];
        a[i] = (x > q) ? v : ((x == q) ? dq : 0.0);
      }
      const double b0 = Wh[q * 33 + tx], b1 = Wh[q * 33 + tx + 16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fma(a[i], b0, acc[i][0]);
        acc[i][1] = fma(a[i], b1, acc[i][1]);
      }
    }
    __syncthreads();   // all reads of the diagonal blocks' stored inverses done before the stores below? (disjoint blocks: kept for the scratch reuse)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      T[(j0 + tx) * P2_LD + i0 + ty + 8 * i] = -acc[i][0];
      T[(j0 + tx + 16) * P2_LD + i0 + ty + 8 * i] = -acc[i][1];
    }
  }
  __syncthreads();
  {
    const int ty = tid >> 4, tx = tid & 15;     // rows x = ty + 16i, columns y = tx + 16j of the 64×64 block
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
    for (int q = 0; q < 64; ++q) {              // W2 = B·M1,  B = L[64.., 0..63],  M1 lower triangular
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = T[(64 + ty + 16 * i) * P2_LD + q];
      const double dq = dinv[q];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int y = tx + 16 * j;
        const double v = T[y * P2_LD + q];
        bb[j] = (q > y) ? v : ((q == y) ? dq : 0.0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Wb[(ty + 16 * i) * 65 + tx + 16 * j] = acc[i][j];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
    for (int q = 0; q < 64; ++q) {              // M[64+x][y] = −Σ_{q<=x} M2[x][q]·W2[q][y]
      double a[4], bb[4];
      const double dq = dinv[64 + q];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = ty + 16 * i;
        const double v = T[(64 + q) * P2_LD + 64 + x];
        a[i] = (x > q) ? v : ((x == q) ? dq : 0.0);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Wb[q * 65 + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) T[(tx + 16 * j) * P2_LD + 64 + ty + 16 * i] = -acc[i][j];   // transposed store
  }
  __syncthreads();

  P2_MARK(7);
  // ---- outputs: L back to the workspace, M = L⁻¹ to the per-slot buffer, z = M·r, logdet, sqmah
  double* Mg = p.Minv + (long long)s * kTile * kTile;
  for (int idx = tid; idx < kTile * kTile; idx += P2_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    if (c <= r) Ag[(long long)r * ld + c] = T[r * P2_LD + c];
    Mg[idx] = (c < r) ? T[c * P2_LD + r] : ((c == r) ? dinv[r] : 0.0);
  }
  P2_MARK(8);
  double zz = 0.0, lg = 0.0;
  if (tid < kTile) {
    double z = dinv[tid] * rk[tid];
    for (int q = 0; q < tid; ++q) z = fma(T[q * P2_LD + tid], rk[q], z);   // M[tid][q], consecutive threads -> consecutive words
    zz = z * z;
    lg = log(dg[tid]);
    p.zk[(long long)s * kTile + tid] = z;
    p.rhs[(long long)s * p.Np + p.k0 + tid] = z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    zz += __shfl_xor_sync(0xffffffffu, zz, o);
    lg += __shfl_xor_sync(0xffffffffu, lg, o);
  }
  if (tid < kTile && lane == 0) {
    red[warp] = zz;
    red[4 + warp] = lg;
  }
  __syncthreads();
  if (tid == 0) {
    const double zsum = ((red[0] + red[1]) + red[2]) + red[3];
    const double lsum = ((red[4] + red[5]) + red[6]) + red[7];
    const double sq = p.sqmah[s] + zsum;
    const double ldt = p.logdet[s] + 2.0 * lsum;
    p.sqmah[s] = sq;
    p.logdet[s] = ldt;
    if (last) {
      if (lnL_out) lnL_out[s] = -(ldt + sq) / 2;
      if (info_out) info_out[s] = 0;
    }
#ifdef SFB_EXPERIMENTS
    tph[9] = clock64();
    if (s == 0 && p.k0 == 0)
      printf("potrf_diag2 cycles: load %lld | S1(b0) %lld | S2(b0) %lld | S3(b0) %lld | blocks1-3 %lld | inv diag %lld | inv off %lld | "
             "store %lld | z,logdet %lld | total %lld\n",
             tph[1] - tph[0], tph[2] - tph[1], tph[3] - tph[2], tph[4] - tph[3], tph[5] - tph[4], tph[6] - tph[5],
             tph[7] - tph[6], tph[8] - tph[7], tph[9] - tph[8], tph[9] - tph[0]);
#endif
  }
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void residual_kernel(const double* __restrict__ model_flux, const double* __restrict__ data_flux,
                                int N, int Np, double* rhs, double* resid_out, double* logdet, double* sqmah,
                                int* info) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) {
    double r = 0.0;
    if (i < N) {
      r = model_flux[(long long)b * N + i] - data_flux[i];
      if (resid_out) resid_out[(long long)b * N + i] = r;
    }
    rhs[(long long)b * Np + i] = r;
  }
  if (i == 0) {
    logdet[b] = 0.0;
    sqmah[b] = 0.0;
    info[b] = 0;
  }
}

__global__ void copy_in_lower_kernel(const double* __restrict__ C, int N, double* W, int Np, long long strideW) {
  const int b = blockIdx.z;
  const int i = blockIdx.y;
  const double* src = C + (long long)b * N * N + (long long)i * N;
  double* dst = W + (long long)b * strideW + (long long)i * Np;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= i; j += gridDim.x * blockDim.x)
    dst[j] = (i < N) ? src[j] : ((j == i) ? 1.0 : 0.0);
}

__global__ void copy_out_lower_kernel(double* C, int N, const double* __restrict__ W, int Np, long long strideW) {
  const int b = blockIdx.z;
  const int i = blockIdx.y;
  double* dst = C + (long long)b * N * N + (long long)i * N;
  const double* src = W + (long long)b * strideW + (long long)i * Np;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= i; j += gridDim.x * blockDim.x) dst[j] = src[j];
}

__global__ void zero_rhs_kernel(double* rhs, int Np, double* logdet, double* sqmah, int* info) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) rhs[(long long)b * Np + i] = 0.0;
  if (i == 0) { logdet[b] = 0.0; sqmah[b] = 0.0; info[b] = 0; }
}

// Forward substitution z = L⁻¹ r, one CTA per matrix (the cho_solve seam; not on the fused path).
// Blocks of 64 rows: all warps accumulate the dot products with the already-solved prefix (coalesced row
// reads), then warp 0 solves the 64×64 triangle.
__global__ void __launch_bounds__(256) solve_lower_kernel(const double* __restrict__ L, long long strideL, int ldl,
                                                          const double* __restrict__ r, double* z, int N) {
  extern __shared__ __align__(16) double zs[];  // [N] solution so far
  __shared__ double partial[64];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Lb = L + (long long)b * strideL;
  for (int i0 = 0; i0 < N; i0 += 64) {
    const int nb = min(64, N - i0);
    for (int rr = warp; rr < nb; rr += 8) {
      const double* row = Lb + (long long)(i0 + rr) * ldl;
      double acc = 0.0;
      for (int c = lane; c < i0; c += 32) acc = fma(row[c], zs[c], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) partial[rr] = r[(long long)b * N + i0 + rr] - acc;
    }
    __syncthreads();
    if (warp == 0) {
      for (int rr = 0; rr < nb; ++rr) {
        const double* row = Lb + (long long)(i0 + rr) * ldl + i0;
        double acc = 0.0;
        for (int c = lane; c < rr; c += 32) acc = fma(row[c], zs[i0 + c], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) zs[i0 + rr] = (partial[rr] - acc) / row[rr];
        __syncwarp();
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < N; i += 256) z[(long long)b * N + i] = zs[i];
}

constexpr size_t kPotrfSmem = 0;  // static shared memory only
bool g_potrf_blocked = true;      // potrf_diag2_kernel (blocked, shared-memory resident); experiments build can switch

}  // namespace

cudaError_t kernels_init() {
  // Opt in to > 48 KB of dynamic shared memory.  (Forcing the maximum carve-out so that panel CTAs can
  // co-reside with syrk CTAs was measured to cost the syrk kernel 3.5 % — its epilogue's C tiles go through
  // L1 — and is therefore not done.)
  struct Item { const void* fn; int bytes; };
  const Item items[] = {{(const void*)syrk_kernel, GEMM_SMEM_BYTES},
                        {(const void*)trsm_kernel<false>, GEMM_SMEM_BYTES},
                        {(const void*)trsm_kernel<true>, GEMM_SMEM_BYTES},
                        {(const void*)potrf_diag2_kernel, (int)kPotrf2Smem},
                        {(const void*)solve_lower_kernel, 200 * 1024}};
  for (const Item& it : items) {
    cudaError_t e = cudaFuncSetAttribute(it.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, it.bytes);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

void potrf_set_blocked(bool on) { g_potrf_blocked = on; }

cudaError_t launch_residual(const double* model_flux, const double* data_flux, int N, int Np, int B,
                            double* rhs, double* resid_out, double* logdet, double* sqmah, int* info,
                            cudaStream_t st) {
  dim3 grid((Np + 255) / 256, B);
  if (model_flux == nullptr) {
    zero_rhs_kernel<<<grid, 256, 0, st>>>(rhs, Np, logdet, sqmah, info);
  } else {
    residual_kernel<<<grid, 256, 0, st>>>(model_flux, data_flux, N, Np, rhs, resid_out, logdet, sqmah, info);
  }
  return cudaGetLastError();
}

cudaError_t launch_potrf_diag(const CholParams& p, int B, int last, double* lnL_out, int* info_out,
                              cudaStream_t st) {
  if (g_potrf_blocked)
    potrf_diag2_kernel<<<B, P2_THREADS, kPotrf2Smem, st>>>(p, last, lnL_out, info_out);
  else
    potrf_diag_kernel<<<B, PD_THREADS, kPotrfSmem, st>>>(p, last, lnL_out, info_out);
  return cudaGetLastError();
}

cudaError_t launch_trsm(const CholParams& p, const GemmMaps& m, int slot0, int B, cudaStream_t st) {
  const int rows = p.Np - p.k0 - kTile;
  if (rows <= 0) return cudaSuccess;
  trsm_kernel<false><<<dim3(rows / 64, B), GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(*m.W, *m.Minv, p, slot0, OzParams{}, 0);
  return cudaGetLastError();
}

// trsm + fixed-point slices of the panel it produces (chunks [chunk0, chunk0+4) of oz.P, flags in oz.F)
cudaError_t launch_trsm_slice(const CholParams& p, const GemmMaps& m, const OzParams& oz, int chunk0, int slot0, int B,
                              cudaStream_t st) {
  const int rows = p.Np - p.k0 - kTile;
  if (rows <= 0) return cudaSuccess;
  trsm_kernel<true><<<dim3(rows / 64, B), GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(*m.W, *m.Minv, p, slot0, oz, chunk0);
  return cudaGetLastError();
}

// update of tile columns [jt0, jt0+njt) (each from its diagonal tile down) by columns [kb, kb+K)
cudaError_t launch_syrk_strip(const CholParams& p, const GemmMaps& m, int slot0, int kb, int K, int jt0, int njt,
                              int B, cudaStream_t st) {
  const int rows = p.Np / kTile - jt0;
  if (rows <= 0 || K <= 0 || njt <= 0) return cudaSuccess;
  syrk_kernel<<<dim3(2 * njt, rows, B), GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(*m.W, p, slot0, kb, K, jt0, 1);
  return cudaGetLastError();
}

// trailing update of every tile column >= jt0 by columns [kb, kb+K)
cudaError_t launch_syrk_tri(const CholParams& p, const GemmMaps& m, int slot0, int kb, int K, int jt0, int B,
                            cudaStream_t st) {
  const int T = p.Np / kTile - jt0;
  if (T <= 0 || K <= 0) return cudaSuccess;
  syrk_kernel<<<dim3(T * (T + 1), 1, B), GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(*m.W, p, slot0, kb, K, jt0, 0);
  return cudaGetLastError();
}

// Tensor maps for TMA: 3-D {k (inner), row, slot}, box {16 doubles, 64 rows, 1}, 128-byte swizzle.
cudaError_t make_gemm_maps(GemmMaps* out, double* W, int Np, double* Minv, int slots) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return e;
  if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  out->W = new CUtensorMap;
  out->Minv = new CUtensorMap;
  const cuuint32_t box[3] = {BK, 64, 1}, estr[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {(cuuint64_t)Np, (cuuint64_t)Np, (cuuint64_t)slots};
    const cuuint64_t strides[2] = {(cuuint64_t)Np * 8, (cuuint64_t)Np * Np * 8};
    CUresult r = encode(out->W, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, W, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)kTile, (cuuint64_t)kTile, (cuuint64_t)slots};
    const cuuint64_t strides[2] = {(cuuint64_t)kTile * 8, (cuuint64_t)kTile * kTile * 8};
    CUresult r = encode(out->Minv, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, Minv, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

void free_gemm_maps(GemmMaps* m) {
  delete m->W;
  delete m->Minv;
  m->W = m->Minv = nullptr;
}

// ---- shared-factor path -----------------------------------------------------------------------------
cudaError_t make_fwd_maps(FwdMaps* out, double* Zt, int Np, int Jp, double* MinvAll, int panels) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return e;
  if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  out->Z = new CUtensorMap;
  out->Mall = new CUtensorMap;
  const cuuint32_t box[3] = {BK, 64, 1}, estr[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {(cuuint64_t)Np, (cuuint64_t)Jp, 1};
    const cuuint64_t strides[2] = {(cuuint64_t)Np * 8, (cuuint64_t)Np * Jp * 8};
    if (encode(out->Z, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, Zt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)kTile, (cuuint64_t)kTile, (cuuint64_t)panels};
    const cuuint64_t strides[2] = {(cuuint64_t)kTile * 8, (cuuint64_t)kTile * kTile * 8};
    if (encode(out->Mall, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, MinvAll, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

void free_fwd_maps(FwdMaps* m) {
  delete m->Z;
  delete m->Mall;
  m->Z = m->Mall = nullptr;
}

cudaError_t launch_pack_rhs(const double* model_flux, const double* data_flux, const double* X, int N, int Np, int M,
                            int J, int Jp, double* Zt, double* resid_out, cudaStream_t st) {
  pack_rhs_kernel<<<dim3((Np + 255) / 256, Jp), 256, 0, st>>>(model_flux, data_flux, X, N, Np, M, J, Zt, resid_out);
  return cudaGetLastError();
}

// Z = L⁻¹·RHS for the Jp rows of Zt, L = the factor in workspace slot `slotL`, M_k = MinvAll[k]
cudaError_t launch_forward_rows(const FwdMaps& fm, const GemmMaps& gm, double* Zt, int Np, int Jp, int slotL,
                                cudaStream_t st, long long* launches) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute((const void*)fwd_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute((const void*)fwd_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int nt = Np / kTile;
  for (int k = 0; k < nt; ++k) {
    const int k0 = k * kTile;
    fwd_diag_kernel<<<Jp / 64, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(*fm.Z, *fm.Mall, Zt, Np, k0, k);
    if (k + 1 < nt)
      fwd_update_kernel<<<dim3(Jp / 128, (Np - k0 - kTile) / 64), GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(*fm.Z, *gm.W, Zt, Np,
                                                                                                    k0, slotL);
    if (launches) *launches += (k + 1 < nt) ? 2 : 1;
  }
  return cudaGetLastError();
}

cudaError_t launch_copy_in_lower(const double* C, int N, double* W, int Np, long long strideW, int B,
                                 cudaStream_t st) {
  dim3 grid(min(64, (Np + 255) / 256), Np, B);
  copy_in_lower_kernel<<<grid, 256, 0, st>>>(C, N, W, Np, strideW);
  return cudaGetLastError();
}

cudaError_t launch_copy_out_lower(double* C, int N, const double* W, int Np, long long strideW, int B,
                                  cudaStream_t st) {
  dim3 grid(min(64, (N + 255) / 256), N, B);
  copy_out_lower_kernel<<<grid, 256, 0, st>>>(C, N, W, Np, strideW);
  return cudaGetLastError();
}

cudaError_t launch_solve_lower(const double* L, long long strideL, int ldl, const double* r, double* z, int N,
                               int B, cudaStream_t st) {
  solve_lower_kernel<<<B, 256, sizeof(double) * N, st>>>(L, strideL, ldl, r, z, N);
  return cudaGetLastError();
}

}  // namespace sfb
