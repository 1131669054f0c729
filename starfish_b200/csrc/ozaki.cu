// fp64-equivalent trailing update on the int8 tensor path (tcgen05.mma kind::i8, TMEM accumulators) — the
// SFB_SOLVER_DENSE_I8 mode of the dense factorisation (sm_100a).
//
// The dense path is fp64-bound: A_ij −= L_ik·L_jkᵀ is N³/3 of its FLOPs and the fp64 tensor rate (DMMA,
// 37.1 TFLOP/s) caps it at 201 evals/s per GPU (N=8192).  tcgen05 has no fp64 kind, but it multiplies int8
// exactly into int32.  This file restates the update's OPERANDS in fixed point and keeps every product exact
// (an Ozaki-style split; it is not iterative refinement — potrf_diag, trsm, the forward solve and logdet stay
// true fp64, chol.cu):
//
//   row scale  2^e_i ∈ (2√C_ii, 4√C_ii]   (C_ii = the diagonal before the factorisation; |L_ik| ≤ √C_ii for an
//                                          SPD matrix, so |L_ik|/2^e_i < ½ for every panel, fixed up front)
//   q_ik     = rint(L_ik · 2^(47−e_i))     a 48-bit signed integer
//   q_ik     = Σ_t b_ikt·256^(5−t)         six balanced radix-256 digits, b ∈ [−128, 127]  (int8 "slices")
//   L_ik·L_jk ≈ 2^(e_i+e_j−14) · Σ_{d=0..6} 256^(−d) · S_d ,   S_d = Σ_k Σ_{s+t=d} b_iks·b_jkt
//
// Every S_d is exact in int32 (|S_d| ≤ 6·2^14·1024 < 2^27 for K ≤ 1024), so one TMEM accumulator per
// anti-diagonal d: 7 accumulators × 64 columns = 448 of the 512 TMEM columns for a 128×64 tile, fed by 26 int8
// MMAs per 32-deep k-chunk (pairs s,t ≤ 5, s+t ≤ 6).  What is dropped is the anti-diagonals d ≥ 7 and the
// rounding of q: measured |ΔlnL|/|lnL| ≤ 1.6e-12 against the dense oracle up to cond 7e5 (tools/ozaki_experiment.py,
// tests/test_gpu_ozaki.py; 5 accumulators give 1e-7 — the anti-diagonal count, not the digit count, is what
// matters).  The epilogue recombines Σ_d 256^(−d)·S_d by Horner in fp64 and applies C −= 2^(e_i−7)·2^(e_j−7)·(…).
//
// Data layout.  The sliced panels of the current outer block live in P (per slot), pre-tiled so that an operand
// tile is ONE contiguous block and already in the canonical K-major shared-memory layout of the MMA:
//     P[chunk c (32 k)][row group g = row/8][slice t][row%8][32 bytes]        (NCH × Np/8 × 6 × 256 B)
// -> the A operand (128 rows, all slices) of a chunk is 24 KB contiguous, the B operand (64 rows) 12 KB: two 1-D
// bulk async copies (cp.async.bulk, SASS UBLKCP) per pipeline stage, no tensor map.  Inside a 256-byte row group
// the two 16-byte halves of a row are XOR-swapped by bit 2 of the row (the 32-byte swizzle pattern, applied by
// the slicing kernel), and the 8-row groups of one slice are 6·256 = 1536 B apart: UMMA descriptor
// {SWIZZLE_32B, SBO = 1536}.
//
// Kernel shape: 192 threads = warp 0 bulk-copy producer (one lane), warp 1 MMA issuer (one elected lane) + TMEM
// allocator, warps 2-5 epilogue (TMEM lane quarter = warp%4).  5-stage ring of 36 KB, full/empty mbarriers,
// tcgen05.commit releases a stage / signals the epilogue.  The epilogue warps prefetch the C tile (coalesced,
// into registers) while the main loop runs, transpose the recombined update through the idle ring and finish
// the read-modify-write with coalesced streaming stores.
#include <algorithm>
#include <cstdint>

#include "sfb_internal.cuh"

namespace sfb {

namespace {

constexpr int OZ_S = kOzSlices;      // 6
constexpr int OZ_NACC = 7;           // anti-diagonals 0..6
constexpr int OZ_KC = kOzChunk;      // 32
constexpr int OZ_BM = 128, OZ_BN = 64;
constexpr int OZ_STAGES = 4;
constexpr int OZ_GROUP_BYTES = 8 * OZ_KC;                    // 256: 8 rows × 32 B of one slice
constexpr int OZ_ROWGROUP_BYTES = OZ_S * OZ_GROUP_BYTES;     // 1536: all slices of 8 rows
constexpr int OZ_A_BYTES = (OZ_BM / 8) * OZ_ROWGROUP_BYTES;  // 24576
constexpr int OZ_B_BYTES = (OZ_BN / 8) * OZ_ROWGROUP_BYTES;  // 12288
constexpr int OZ_STAGE_BYTES = OZ_A_BYTES + OZ_B_BYTES;      // 36864
constexpr int OZ_THREADS = 192;
constexpr uint32_t OZ_TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem]·B[smem]ᵀ, int8 × int8 -> int32, one thread issues for the CTA (SASS UTCIMMA)
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync)
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// shared memory -> TMEM, 128 lanes × 256 bits (one K-major 128×32-byte operand slab; SASS UTCCP)
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// D[tmem] (+)= A[tmem]·B[smem]ᵀ
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// exact int32 -> double on the full-rate fp64 add pipe: (2^52 + 2^31 + v) − (2^52 + 2^31)
__device__ __forceinline__ double i2d(uint32_t v) {
  return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: lane i of the warp writes 16 consecutive 32-bit columns of TMEM lane (quarter base + i)  (SASS STTM)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
      "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_32B (layout code 6), descriptor version 1:
// start address, LBO (unused for swizzled K-major; canonical value 1) and SBO in 16-byte units.
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(OZ_ROWGROUP_BYTES >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
// instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N >> 3, M >> 4
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) |
                              ((uint32_t)(OZ_BM >> 4) << 24);

// ------------------------------------------------------------------------------------------------
// syrk on the int8 path: C[r0.., c0..] −= Σ_{chunks} (sliced L rows r0..) · (sliced L rows c0..)ᵀ.
//
// A CTA works through `tpc` consecutive tiles of one walker (semi-persistent: long enough to amortise the TMEM
// allocation / barrier set-up and to overlap a tile's epilogue with the next tile's operand loads, short enough —
// a few hundred µs — that the high-priority panel kernels of the look-ahead still find free SMs).
// Tile enumeration (`l` = linear tile index inside the walker):
//   strip    update of tile columns [jt0, jt0+njt) from their diagonal tile down: l -> (x = l % (2·njt), y = l / (2·njt)),
//            row tile jt0+y, 64-column block x; tiles above the diagonal of their tile column (y < x/2) are skipped
//   triangle trailing update of every tile column >= jt0: l in [0, T(T+1)), decoded to (row tile t, 64-column block)
// Roles: warp 0 = bulk-copy producer (runs ahead across tile boundaries), warp 1 = MMA issuer (waits for the previous
// tile's accumulators to be drained), warps 2-5 = epilogue.
// ------------------------------------------------------------------------------------------------
struct OzTile { int r0, c0, live; };

__device__ __forceinline__ OzTile oz_tile(int l, int jt0, int njt, int strip) {
  OzTile t;
  if (strip) {
    const int x = l % (2 * njt), y = l / (2 * njt);
    t.live = (y >= (x >> 1));
    t.r0 = (jt0 + y) * kTile;
    t.c0 = (jt0 + (x >> 1)) * kTile + (x & 1) * 64;
  } else {
    int q = (int)((sqrtf(4.0f * (float)l + 1.0f) - 1.0f) * 0.5f);
    while (q * (q + 1) > l) --q;
    while ((q + 1) * (q + 2) <= l) ++q;
    t.live = 1;
    t.r0 = (jt0 + q) * kTile;
    t.c0 = jt0 * kTile + (l - q * (q + 1)) * 64;
  }
  return t;
}

constexpr uint32_t OZ_TROW = 66 * 8;                 // padded row of the transpose buffer (528 B: conflict-free 16-byte accesses)
constexpr int OZ_TBUF_BYTES = 4 * 32 * OZ_TROW;     // 4 epilogue warps × 32 rows
constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + OZ_TBUF_BYTES + 1024;  // + alignment slack
constexpr uint32_t OZ_TMEM_A = OZ_NACC * OZ_BN;     // first TMEM column of the A operand (TS form): 448..495

// One 32-deep chunk: the six A slices shared memory -> TMEM once (TS form; the SS form re-reads each A slice for every
// pair), then one int8 MMA per digit pair (sa, sb), sa + sb <= 6, into the accumulator of anti-diagonal sa + sb.
// fa / fb: bit t set = digit slab t of the A / B operand block has a non-zero entry; a product with an all-zero slab
// contributes nothing and is skipped (exactly).  Called with literal masks the tests fold away at compile time.
// Copies and MMAs of one thread execute in issue order, so this chunk's copies follow the previous chunk's MMAs
// without a wait.  Returns the number of MMAs issued.
template <bool TS>
__device__ __forceinline__ uint32_t oz_issue_chunk(uint32_t fa, uint32_t fb, uint32_t& touched, uint32_t tmem,
                                                   uint64_t ad0, uint64_t bd0) {
  if constexpr (TS) {
#pragma unroll
    for (int sa = 0; sa < OZ_S; ++sa)
      if ((fa >> sa) & 1u)
        utccp_128x256b(tmem + OZ_TMEM_A + sa * (OZ_KC / 4), ad0 + (uint64_t)(sa * (OZ_GROUP_BYTES >> 4)));
  }
  uint32_t local = 0, n = 0;  // accumulators touched by this chunk so far (compile-time known for literal masks)
#pragma unroll
  for (int sa = 0; sa < OZ_S; ++sa) {
#pragma unroll
    for (int sb = 0; sb < OZ_S; ++sb) {
      if (sa + sb >= OZ_NACC) continue;
      if (!(((fa >> sa) & 1u) && ((fb >> sb) & 1u))) continue;
      // the start-address field counts 16-byte units and never leaves its 14-bit range inside the ring, so the
      // other slices' descriptors are the stage's plus a constant
      const uint64_t bd = bd0 + (uint64_t)(sb * (OZ_GROUP_BYTES >> 4));
      const uint32_t d = sa + sb;
      const uint32_t acc = ((touched | local) >> d) & 1u;   // the first product into an accumulator overwrites it
      if constexpr (TS) {
        umma_i8_ts(tmem + d * OZ_BN, tmem + OZ_TMEM_A + sa * (OZ_KC / 4), bd, OZ_IDESC, acc);
      } else {
        umma_i8(tmem + d * OZ_BN, ad0 + (uint64_t)(sa * (OZ_GROUP_BYTES >> 4)), bd, OZ_IDESC, acc);
      }
      local |= 1u << d;
      ++n;
    }
  }
  touched |= local;
  return n;
}

// ---- epilogue (4 warps; warp q owns tile rows 32q..32q+31 = its TMEM lane quarter).
// Two thread mappings: TMEM hands a thread ONE ROW (lane = row, 16 columns per load), global memory wants a
// warp on one row (lane = column pair, 512 contiguous bytes).  The C tile is therefore fetched in the global
// mapping BEFORE the accumulators are ready (32 independent 16-byte loads per thread, in flight under the
// main loop), the recombined update goes through a warp-private transpose buffer, the accumulators are handed
// back to the MMA warp, and the read-modify-write finishes with coalesced streaming stores while the next
// tile's MMAs already run.
__device__ __forceinline__ void oz_epilogue(const CholParams& p, const OzParams& oz, int s, int l0, int l1, int jt0, int njt,
                                            int strip, uint32_t tmem, uint32_t tbuf0, int q, int lane, uint32_t meta,
                                            uint32_t accfull, uint32_t tmem_empty, const volatile uint32_t* touched_p) {
  const double* rs = oz.rscale + (long long)s * p.Np;
  const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t tbuf = tbuf0 + (uint32_t)q * (32 * OZ_TROW);
  int k = 0;
  for (int l = l0; l < l1; ++l) {
    const OzTile t = oz_tile(l, jt0, njt, strip);
    if (!t.live) continue;
    const double ri = rs[t.r0 + q * 32 + lane];                                         // row scale, TMEM mapping
    const double2 rj = *reinterpret_cast<const double2*>(rs + t.c0 + 2 * lane);        // column scales, global mapping
    double* Cw = p.W + (long long)s * p.strideW + (long long)(t.r0 + q * 32) * p.Np + t.c0 + 2 * lane;
    double2 creg[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) creg[r] = __ldcs(reinterpret_cast<const double2*>(Cw + (long long)r * p.Np));
    mbar_wait(meta, k & 1);
    const uint32_t touched = *touched_p;
    mbar_wait(accfull, k & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int cb = 0; cb < OZ_BN / 16; ++cb) {
      double acc[16];
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll
      for (int d = OZ_NACC - 1; d >= 0; --d) {   // Horner from the least significant anti-diagonal
        if ((touched >> d) & 1u) {
          tmem_ld16(tlane + d * OZ_BN + cb * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], 0.00390625, i2d(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] *= 0.00390625;
        }
      }
      const uint32_t dst = tbuf + (uint32_t)lane * OZ_TROW + (uint32_t)cb * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dst + 16 * j), "d"(acc[2 * j] * ri), "d"(acc[2 * j + 1] * ri)
                     : "memory");
    }
    // accumulators drained: hand TMEM back to the MMA warp (which orders its next MMAs after this arrive)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      double tx, ty;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tx), "=d"(ty) : "r"(tbuf + (uint32_t)r * OZ_TROW + 16 * lane) : "memory");
      double2 c = creg[r];
      c.x = fma(-tx, rj.x, c.x);
      c.y = fma(-ty, rj.y, c.y);
      __stcs(reinterpret_cast<double2*>(Cw + (long long)r * p.Np), c);
    }
    __syncwarp();  // the transpose buffer is rewritten by the next tile
    ++k;
  }

}

template <bool TS>
__global__ void __launch_bounds__(OZ_THREADS, 1)
    syrk_i8_kernel(CholParams p, OzParams oz, int nch, int jt0, int njt, int strip, int ntiles, int tpc) {
  const int s = blockIdx.z;
  if (p.info[s] != 0) return;
  const int l0 = blockIdx.x * tpc;
  const int l1 = min(ntiles, l0 + tpc);

  extern __shared__ uint8_t oz_smem_raw[];
  __shared__ uint64_t bars[2 * OZ_STAGES + 3];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t touched_s;
  const uint32_t ring = (smem_u32(oz_smem_raw) + 1023u) & ~1023u;
  const uint32_t tbuf0 = ring + OZ_STAGES * OZ_STAGE_BYTES;
  const uint32_t bar0 = smem_u32(bars);  // full[s] at +8s, empty[s] at +8(STAGES+s), then accfull, tmem_empty
  const uint32_t accfull = bar0 + 16 * OZ_STAGES, tmem_empty = accfull + 8, meta = accfull + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2 * OZ_STAGES + 1; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init(tmem_empty, 4);  // one arrival per epilogue warp
    mbar_init(meta, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), OZ_TMEM_COLS);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const int8_t* Ps = oz.P + (long long)s * oz.strideP;
  const long long chunk_bytes = (long long)(p.Np / 8) * OZ_ROWGROUP_BYTES;

  if (warp == 0) {
    if (lane == 0) {  // ---- producer: two contiguous bulk copies per chunk, continuous over the CTA's tiles
      int g = 0;      // chunks issued so far (ring position)
      for (int l = l0; l < l1; ++l) {
        const OzTile t = oz_tile(l, jt0, njt, strip);
        if (!t.live) continue;
        const int8_t* srcA = Ps + (long long)(t.r0 / 8) * OZ_ROWGROUP_BYTES;
        const int8_t* srcB = Ps + (long long)(t.c0 / 8) * OZ_ROWGROUP_BYTES;
        for (int c = 0; c < nch; ++c, ++g) {
          const int st = g % OZ_STAGES;
          if (g >= OZ_STAGES) mbar_wait(bar0 + 8 * (OZ_STAGES + st), ((g / OZ_STAGES) - 1) & 1);
          const uint32_t full = bar0 + 8 * st;
          const uint32_t dst = ring + st * OZ_STAGE_BYTES;
          mbar_arrive_expect_tx(full, OZ_STAGE_BYTES);
          bulk_g2s(dst, srcA + c * chunk_bytes, OZ_A_BYTES, full);
          bulk_g2s(dst + OZ_A_BYTES, srcB + c * chunk_bytes, OZ_B_BYTES, full);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: 26 int8 MMAs per chunk into 7 accumulators.  The whole warp runs the loop (descriptor
    // arithmetic stays warp-uniform, i.e. in uniform registers); one elected lane issues.
    const bool leader = elect_one();
    const uint8_t* Fs = oz.F + (long long)s * oz.strideF;
    const int nrb = p.Np / 64;
    unsigned long long issued = 0;
    int g = 0, k = 0;  // ring position, live tiles done
    for (int l = l0; l < l1; ++l) {
      const OzTile t = oz_tile(l, jt0, njt, strip);
      if (!t.live) continue;
      // which digit slabs of this tile's operands are not identically zero, chunk by chunk (written by the slicing
      // kernel; lane j holds chunks j and j+32) — fetched while the previous tile's accumulators are being drained
      uint32_t fa0 = 0, fb0 = 0, fa1 = 0, fb1 = 0;
      if (lane < nch) {
        const uint8_t* f = Fs + (long long)lane * nrb;
        fa0 = f[t.r0 / 64] | f[t.r0 / 64 + 1];
        fb0 = f[t.c0 / 64];
      }
      if (lane + 32 < nch) {
        const uint8_t* f = Fs + (long long)(lane + 32) * nrb;
        fa1 = f[t.r0 / 64] | f[t.r0 / 64 + 1];
        fb1 = f[t.c0 / 64];
      }
      if (k > 0) {  // the epilogue must have drained the accumulators of the previous tile
        mbar_wait(tmem_empty, (k - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      uint32_t touched = 0;  // accumulators that have received a product in this tile
      for (int c = 0; c < nch; ++c, ++g) {
        const int st = g % OZ_STAGES;
        const uint32_t fa = __shfl_sync(0xffffffffu, c < 32 ? fa0 : fa1, c & 31);
        const uint32_t fb = __shfl_sync(0xffffffffu, c < 32 ? fb0 : fb1, c & 31);
        mbar_wait(bar0 + 8 * st, (g / OZ_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = ring + st * OZ_STAGE_BYTES, b0 = a0 + OZ_A_BYTES;
        const uint64_t ad0 = oz_desc(a0), bd0 = oz_desc(b0);
        if (leader) {
          // The two digit patterns that dominate (nothing zero / leading slab zero), for either operand, get fully
          // unrolled code with compile-time masks: a single thread issues every MMA of the CTA, so per-product
          // run-time tests would make the issue loop the bottleneck.  Anything else takes the generic path.
          if (fa == 0x3fu && fb == 0x3fu) issued += oz_issue_chunk<TS>(0x3fu, 0x3fu, touched, tmem, ad0, bd0);
          else if (fa == 0x3eu && fb == 0x3eu) issued += oz_issue_chunk<TS>(0x3eu, 0x3eu, touched, tmem, ad0, bd0);
          else if (fa == 0x3eu && fb == 0x3fu) issued += oz_issue_chunk<TS>(0x3eu, 0x3fu, touched, tmem, ad0, bd0);
          else if (fa == 0x3fu && fb == 0x3eu) issued += oz_issue_chunk<TS>(0x3fu, 0x3eu, touched, tmem, ad0, bd0);
          else issued += oz_issue_chunk<TS>(fa, fb, touched, tmem, ad0, bd0);
          umma_commit(bar0 + 8 * (OZ_STAGES + st));  // stage free once these MMAs (and copies) have read it
        }
        __syncwarp();
      }
      if (leader) {
        touched_s = touched;               // which accumulators hold a sum (the others are stale: treated as zero)
        mbar_arrive(meta);                 // release: the epilogue reads touched_s after acquiring this barrier
        umma_commit(accfull);
      }
      __syncwarp();
      ++k;
    }
    if (leader && oz.stats) {
      atomicAdd(oz.stats, issued);
      atomicAdd(oz.stats + 1, (unsigned long long)k * nch * 26ull);
    }
  } else {
    oz_epilogue(p, oz, s, l0, l1, jt0, njt, strip, tmem, tbuf0, warp & 3, lane, meta, accfull, tmem_empty, &touched_s);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, OZ_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// The same update with the A operand fed to TMEM by LOADER WARPS (registers -> TMEM, tcgen05.st) instead of
// tcgen05.cp.  Why: the 128×64×32 int8 MMA needs 32 tensor cycles, but with both operands in shared memory it reads
// 6 KB per MMA (48 cycles at 128 B/clk), and the TS form of syrk_i8_kernel pays 6 × 4 KB of tcgen05.cp per chunk at
// 64 B/clk IN the tensor pipe's issue order (26·32 + 384 cycles per chunk = 46.8 per MMA; measured 48.2,
// tools/exp/umma_i8_probe.cu).  A register -> TMEM store runs beside the MMAs (256 B/clk), so the issue stream holds
// MMAs only and the tensor pipe sees its 32-cycle floor; shared memory carries B (2 KB per MMA) plus one LDS pass
// over A (24 KB per chunk): 76 of the 106 KB a chunk's 832 cycles can deliver.
//
// TMEM: 7 accumulators × 64 columns, then a ring of four 16-column A buffers (448..511); a buffer holds TWO digit
// slabs (8 columns each).  The slabs of a chunk are paired {0,5}, {1,4}, {2,3} — 8, 9 and 9 of the 26 MMAs — so the
// MMA warp consumes a buffer every ≈ 280 cycles and the loaders may run three buffers ahead.
// Warp groups (setmaxnreg needs aligned groups of four warps): 0-3 epilogue (240 registers), 4-7 loaders — thread =
// operand row = TMEM lane, 12 conflict-free 16-byte shared loads and three 16-column stores per chunk —, 8 = bulk-copy
// producer, 9 = MMA issuer, 10-11 idle.
// Barriers: full/empty per stage (empty = 1 commit of the MMA warp + the 4 loader warps), afull/aempty per A buffer.
// ------------------------------------------------------------------------------------------------
constexpr int OZ_ST_THREADS = 384;
constexpr int OZ_ABUF = 4;                       // TMEM A buffers of 16 columns
constexpr int OZ_NBARS_ST = 2 * OZ_STAGES + 2 * OZ_ABUF + 3;

// MMAs of one slab pair PR = {PR, 5−PR}; A from the TMEM buffer at column `acol`
template <int PR>
__device__ __forceinline__ uint32_t oz_issue_pair(uint32_t fa, uint32_t fb, uint32_t& touched, uint32_t tmem, uint32_t acol,
                                                  uint64_t bd0) {
  uint32_t local = 0, n = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int sa = h ? 5 - PR : PR;
    if (!((fa >> sa) & 1u)) continue;
#pragma unroll
    for (int sb = 0; sb < OZ_S; ++sb) {
      if (sa + sb >= OZ_NACC) continue;
      if (!((fb >> sb) & 1u)) continue;
      const uint32_t d = sa + sb;
      const uint32_t acc = ((touched | local) >> d) & 1u;
      umma_i8_ts(tmem + d * OZ_BN, acol + h * (OZ_KC / 4), bd0 + (uint64_t)(sb * (OZ_GROUP_BYTES >> 4)), OZ_IDESC, acc);
      local |= 1u << d;
      ++n;
    }
  }
  touched |= local;
  return n;
}
template <int PR>
__device__ __forceinline__ uint32_t oz_issue_pair_dispatch(uint32_t fa, uint32_t fb, uint32_t& touched, uint32_t tmem,
                                                           uint32_t acol, uint64_t bd0) {
  // compile-time masks for the dominant digit patterns (see syrk_i8_kernel)
  if (fa == 0x3fu && fb == 0x3fu) return oz_issue_pair<PR>(0x3fu, 0x3fu, touched, tmem, acol, bd0);
  if (fa == 0x3eu && fb == 0x3eu) return oz_issue_pair<PR>(0x3eu, 0x3eu, touched, tmem, acol, bd0);
  if (fa == 0x3eu && fb == 0x3fu) return oz_issue_pair<PR>(0x3eu, 0x3fu, touched, tmem, acol, bd0);
  if (fa == 0x3fu && fb == 0x3eu) return oz_issue_pair<PR>(0x3fu, 0x3eu, touched, tmem, acol, bd0);
  return oz_issue_pair<PR>(fa, fb, touched, tmem, acol, bd0);
}

__global__ void __launch_bounds__(OZ_ST_THREADS, 1)
    syrk_i8_st_kernel(CholParams p, OzParams oz, int nch, int jt0, int njt, int strip, int ntiles, int tpc) {
  const int s = blockIdx.z;
  if (p.info[s] != 0) return;
  const int l0 = blockIdx.x * tpc;
  const int l1 = min(ntiles, l0 + tpc);

  extern __shared__ uint8_t oz_smem_raw[];
  __shared__ uint64_t bars[OZ_NBARS_ST];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t touched_s;
  const uint32_t ring = (smem_u32(oz_smem_raw) + 1023u) & ~1023u;
  const uint32_t tbuf0 = ring + OZ_STAGES * OZ_STAGE_BYTES;
  const uint32_t bar0 = smem_u32(bars);                       // full[s] at +8s, empty[s] at +8(STAGES+s)
  const uint32_t afull0 = bar0 + 16 * OZ_STAGES;              // afull[i] at +8i, aempty[i] at +8(ABUF+i)
  const uint32_t aempty0 = afull0 + 8 * OZ_ABUF;
  const uint32_t accfull = aempty0 + 8 * OZ_ABUF, tmem_empty = accfull + 8, meta = accfull + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < OZ_STAGES; ++i) {
      mbar_init(bar0 + 8 * i, 1);
      mbar_init(bar0 + 8 * (OZ_STAGES + i), 5);  // the MMA warp's commit + the four loader warps
    }
#pragma unroll
    for (int i = 0; i < OZ_ABUF; ++i) {
      mbar_init(afull0 + 8 * i, 4);              // one arrival per loader warp
      mbar_init(aempty0 + 8 * i, 1);
    }
    mbar_init(accfull, 1);
    mbar_init(tmem_empty, 4);
    mbar_init(meta, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(smem_u32(&tmem_base_s), OZ_TMEM_COLS);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const int8_t* Ps = oz.P + (long long)s * oz.strideP;
  const long long chunk_bytes = (long long)(p.Np / 8) * OZ_ROWGROUP_BYTES;

  if (warp < 4) {
    setmaxnreg_inc<240>();
    oz_epilogue(p, oz, s, l0, l1, jt0, njt, strip, tmem, tbuf0, warp, lane, meta, accfull, tmem_empty, &touched_s);
  } else if (warp < 8) {
    setmaxnreg_dec<96>();
    // ---- loaders: thread = row of the A tile = TMEM lane.  The two 16-byte halves of a row sit swapped in shared
    // memory for rows 4-7 of a group (32-byte swizzle); reading the LOGICAL halves in order makes the eight lanes of
    // a quarter-warp cover all 32 banks (the physical order would put rows r and r+4 on the same banks).
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t aoff = (uint32_t)(row >> 3) * OZ_ROWGROUP_BYTES + (uint32_t)(row & 7) * OZ_KC;
    const uint32_t h0 = (uint32_t)((row & 7) >> 2) * 16, h1 = h0 ^ 16;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + OZ_TMEM_A;
    int g = 0, ga = 0;  // chunks, A buffers handled so far
    for (int l = l0; l < l1; ++l) {
      const OzTile t = oz_tile(l, jt0, njt, strip);
      if (!t.live) continue;
      for (int c = 0; c < nch; ++c, ++g) {
        const int st = g % OZ_STAGES;
        mbar_wait(bar0 + 8 * st, (g / OZ_STAGES) & 1);
        const uint32_t base = ring + st * OZ_STAGE_BYTES + aoff;
        uint4 lo[OZ_S], hi[OZ_S];
#pragma unroll
        for (int sa = 0; sa < OZ_S; ++sa) {
          lo[sa] = lds128(base + sa * OZ_GROUP_BYTES + h0);
          hi[sa] = lds128(base + sa * OZ_GROUP_BYTES + h1);
        }
#pragma unroll
        for (int pr = 0; pr < 3; ++pr, ++ga) {
          const int slot = ga % OZ_ABUF;
          if (ga >= OZ_ABUF) mbar_wait(aempty0 + 8 * slot, ((ga / OZ_ABUF) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          tmem_st16(tl + 16 * slot, lo[pr], hi[pr], lo[5 - pr], hi[5 - pr]);
          tmem_st_wait();
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(afull0 + 8 * slot);
        }
        // every shared-memory read of this stage has been consumed by a store above
        if (lane == 0) mbar_arrive(bar0 + 8 * (OZ_STAGES + st));
      }
    }
  } else {
    setmaxnreg_dec<96>();
    if (warp == 8) {
      if (lane == 0) {  // ---- producer (as in syrk_i8_kernel)
        int g = 0;
        for (int l = l0; l < l1; ++l) {
          const OzTile t = oz_tile(l, jt0, njt, strip);
          if (!t.live) continue;
          const int8_t* srcA = Ps + (long long)(t.r0 / 8) * OZ_ROWGROUP_BYTES;
          const int8_t* srcB = Ps + (long long)(t.c0 / 8) * OZ_ROWGROUP_BYTES;
          for (int c = 0; c < nch; ++c, ++g) {
            const int st = g % OZ_STAGES;
            if (g >= OZ_STAGES) mbar_wait(bar0 + 8 * (OZ_STAGES + st), ((g / OZ_STAGES) - 1) & 1);
            const uint32_t full = bar0 + 8 * st;
            const uint32_t dst = ring + st * OZ_STAGE_BYTES;
            mbar_arrive_expect_tx(full, OZ_STAGE_BYTES);
            bulk_g2s(dst, srcA + c * chunk_bytes, OZ_A_BYTES, full);
            bulk_g2s(dst + OZ_A_BYTES, srcB + c * chunk_bytes, OZ_B_BYTES, full);
          }
        }
      }
    } else if (warp == 9) {
      // ---- MMA issuer: per chunk three A buffers, 8 + 9 + 9 MMAs, B from the shared-memory stage
      const bool leader = elect_one();
      const uint8_t* Fs = oz.F + (long long)s * oz.strideF;
      const int nrb = p.Np / 64;
      unsigned long long issued = 0;
      int g = 0, ga = 0, k = 0;
      for (int l = l0; l < l1; ++l) {
        const OzTile t = oz_tile(l, jt0, njt, strip);
        if (!t.live) continue;
        uint32_t fa0 = 0, fb0 = 0, fa1 = 0, fb1 = 0;
        if (lane < nch) {
          const uint8_t* f = Fs + (long long)lane * nrb;
          fa0 = f[t.r0 / 64] | f[t.r0 / 64 + 1];
          fb0 = f[t.c0 / 64];
        }
        if (lane + 32 < nch) {
          const uint8_t* f = Fs + (long long)(lane + 32) * nrb;
          fa1 = f[t.r0 / 64] | f[t.r0 / 64 + 1];
          fb1 = f[t.c0 / 64];
        }
        if (k > 0) {
          mbar_wait(tmem_empty, (k - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        uint32_t touched = 0;
        for (int c = 0; c < nch; ++c, ++g) {
          const int st = g % OZ_STAGES;
          const uint32_t fa = __shfl_sync(0xffffffffu, c < 32 ? fa0 : fa1, c & 31);
          const uint32_t fb = __shfl_sync(0xffffffffu, c < 32 ? fb0 : fb1, c & 31);
          mbar_wait(bar0 + 8 * st, (g / OZ_STAGES) & 1);
          const uint64_t bd0 = oz_desc(ring + st * OZ_STAGE_BYTES + OZ_A_BYTES);
#pragma unroll
          for (int pr = 0; pr < 3; ++pr, ++ga) {
            const int slot = ga % OZ_ABUF;
            mbar_wait(afull0 + 8 * slot, (ga / OZ_ABUF) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acol = tmem + OZ_TMEM_A + 16 * slot;
            if (leader) {
              if (pr == 0) issued += oz_issue_pair_dispatch<0>(fa, fb, touched, tmem, acol, bd0);
              else if (pr == 1) issued += oz_issue_pair_dispatch<1>(fa, fb, touched, tmem, acol, bd0);
              else issued += oz_issue_pair_dispatch<2>(fa, fb, touched, tmem, acol, bd0);
              umma_commit(aempty0 + 8 * slot);
            }
            __syncwarp();
          }
          if (leader) umma_commit(bar0 + 8 * (OZ_STAGES + st));
          __syncwarp();
        }
        if (leader) {
          touched_s = touched;
          mbar_arrive(meta);
          umma_commit(accfull);
        }
        __syncwarp();
        ++k;
      }
      if (leader && oz.stats) {
        atomicAdd(oz.stats, issued);
        atomicAdd(oz.stats + 1, (unsigned long long)k * nch * 26ull);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, OZ_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// row scales from the diagonal of the (not yet factorised) matrix: rscale_i = 2^(e_i − 7)
// ------------------------------------------------------------------------------------------------
__global__ void oz_rowscale_kernel(CholParams p, OzParams oz) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.Np) return;
  const double d = p.W[(long long)s * p.strideW + (long long)i * p.Np + i];
  double r = 1.0;
  if (d > 0.0 && d < 1e300) {
    int ex;
    frexp(sqrt(d), &ex);        // sqrt(d) = m·2^ex, m ∈ [0.5, 1)  ->  2^(ex+1) ∈ (2√d, 4√d]
    r = ldexp(1.0, ex + 1 - 7);
  }
  oz.rscale[(long long)s * p.Np + i] = r;
}

// ------------------------------------------------------------------------------------------------
// slice the panel just produced by trsm (rows k0+128.., columns k0..k0+127) into chunks [ch0, ch0+4) of P.
// A CTA takes 64 rows (one flag block), a warp 8 of them; per row lane l owns k = 4l..4l+3 (one coalesced 1 KB row
// read), six packed 4-byte stores.  Per (64-row block, chunk) the CTA also records which digit slabs are not
// identically zero (F): |L_ik| is usually far below its row's scale away from the band, so the leading digit slab
// of most operand blocks is all zero and the update kernel skips every product with it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) oz_slice_kernel(CholParams p, OzParams oz, int ch0) {
  const int s = blockIdx.y;
  if (p.info[s] != 0) return;
  __shared__ uint32_t wmask[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row_base = p.k0 + kTile + blockIdx.x * 64;
  const int chunk = ch0 + (lane >> 3);
  const int kb = (4 * lane) & 31;                       // byte inside the 32-byte row
  uint32_t nz = 0;                                      // bit t: slab t of this lane's chunk has a non-zero digit
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + warp * 8 + i;
    const double* src = p.W + (long long)s * p.strideW + (long long)row * p.Np + p.k0 + 4 * lane;
    const double2 v01 = *reinterpret_cast<const double2*>(src);
    const double2 v23 = *reinterpret_cast<const double2*>(src + 2);
    const double inv = 1099511627776.0 / oz.rscale[(long long)s * p.Np + row];  // 2^40 / 2^(e−7) = 2^(47−e)
    const double lim = 140737488355327.0;                                       // 2^47 − 1
    long long qv[4];
    qv[0] = __double2ll_rn(fmin(fmax(v01.x * inv, -lim), lim));
    qv[1] = __double2ll_rn(fmin(fmax(v01.y * inv, -lim), lim));
    qv[2] = __double2ll_rn(fmin(fmax(v23.x * inv, -lim), lim));
    qv[3] = __double2ll_rn(fmin(fmax(v23.y * inv, -lim), lim));
    uint32_t packed[OZ_S];
#pragma unroll
    for (int t = OZ_S - 1; t >= 0; --t) {  // least significant digit first
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long dgt = (long long)(int8_t)(qv[e] & 0xff);
        qv[e] = (qv[e] - dgt) >> 8;
        w |= ((uint32_t)dgt & 0xffu) << (8 * e);
      }
      packed[t] = w;
      nz |= (w != 0u) << t;
    }
    const int r8 = row & 7;
    const int half = ((kb >> 4) ^ (r8 >> 2)) & 1;         // 32-byte swizzle: 16-byte halves swapped for rows 4-7
    int8_t* dst = oz.P + (long long)s * oz.strideP +
                  ((long long)chunk * (p.Np / 8) + (row >> 3)) * OZ_ROWGROUP_BYTES + r8 * OZ_KC + half * 16 + (kb & 15);
#pragma unroll
    for (int t = 0; t < OZ_S; ++t) *reinterpret_cast<uint32_t*>(dst + t * OZ_GROUP_BYTES) = packed[t];
  }
  // OR over the 8 lanes of a chunk, then over the CTA's 8 warps
  nz |= __shfl_xor_sync(0xffffffffu, nz, 1);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 2);
  nz |= __shfl_xor_sync(0xffffffffu, nz, 4);
  if ((lane & 7) == 0) wmask[warp][lane >> 3] = nz;
  __syncthreads();
  if (threadIdx.x < 4) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) m |= wmask[w][threadIdx.x];
    oz.F[(long long)s * oz.strideF + (long long)(ch0 + threadIdx.x) * (p.Np / 64) + row_base / 64] = (uint8_t)m;
  }
}

}  // namespace

namespace {
// Where the A operand of the int8 MMAs comes from: 0 = shared memory (SS form), 1 = TMEM filled by tcgen05.cp
// (418 vs 398 evals/s for the SS form, profiles/r2g_i8_ss_vs_ts.txt), 2 = TMEM filled by loader warps with tcgen05.st
// (syrk_i8_st_kernel).  Modes 0 and 1 stay selectable in an experiments build.
int g_oz_mode = 2;
int g_oz_tpc = 8;      // most tiles a CTA works through
}
void ozaki_set_mode(int m) { g_oz_mode = m; }
void ozaki_set_tpc(int n) { g_oz_tpc = n < 1 ? 1 : n; }

cudaError_t ozaki_init() {
  cudaError_t e = cudaFuncSetAttribute((const void*)syrk_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       OZ_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((const void*)syrk_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)syrk_i8_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
}

size_t oz_panel_bytes_per_slot(int Np, int outer_tiles) {
  return (size_t)(outer_tiles * (kTile / OZ_KC)) * (size_t)(Np / 8) * OZ_ROWGROUP_BYTES;
}
size_t oz_flag_bytes_per_slot(int Np, int outer_tiles) {
  return (size_t)(outer_tiles * (kTile / OZ_KC)) * (size_t)(Np / 64);
}

namespace {
// tiles per CTA: enough CTAs for ~4 waves over the SMs, at most 8 tiles each
cudaError_t launch_syrk_i8(const CholParams& p, const OzParams& oz, int K, int jt0, int njt, int strip, int ntiles, int B,
                           cudaStream_t st) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  long long total = (long long)ntiles * B;
  int tpc = (int)std::min<long long>(g_oz_tpc, std::max<long long>(1, total / (4LL * sms)));
  tpc = std::min(tpc, ntiles);
  const dim3 grid((ntiles + tpc - 1) / tpc, 1, B);
  if (g_oz_mode == 2)
    syrk_i8_st_kernel<<<grid, OZ_ST_THREADS, OZ_SMEM_BYTES, st>>>(p, oz, K / OZ_KC, jt0, njt, strip, ntiles, tpc);
  else if (g_oz_mode == 1)
    syrk_i8_kernel<true><<<grid, OZ_THREADS, OZ_SMEM_BYTES, st>>>(p, oz, K / OZ_KC, jt0, njt, strip, ntiles, tpc);
  else
    syrk_i8_kernel<false><<<grid, OZ_THREADS, OZ_SMEM_BYTES, st>>>(p, oz, K / OZ_KC, jt0, njt, strip, ntiles, tpc);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_oz_rowscale(const CholParams& p, const OzParams& oz, int B, cudaStream_t st) {
  oz_rowscale_kernel<<<dim3((p.Np + 255) / 256, B), 256, 0, st>>>(p, oz);
  return cudaGetLastError();
}

cudaError_t launch_oz_slice(const CholParams& p, const OzParams& oz, int chunk0, int B, cudaStream_t st) {
  const int rows = p.Np - p.k0 - kTile;
  if (rows <= 0) return cudaSuccess;
  oz_slice_kernel<<<dim3(rows / 64, B), 256, 0, st>>>(p, oz, chunk0);
  return cudaGetLastError();
}

cudaError_t launch_oz_syrk_strip(const CholParams& p, const OzParams& oz, int K, int jt0, int njt, int B,
                                 cudaStream_t st) {
  const int rows = p.Np / kTile - jt0;
  if (rows <= 0 || K <= 0 || njt <= 0) return cudaSuccess;
  return launch_syrk_i8(p, oz, K, jt0, njt, 1, 2 * njt * rows, B, st);
}

cudaError_t launch_oz_syrk_tri(const CholParams& p, const OzParams& oz, int K, int jt0, int B, cudaStream_t st) {
  const int T = p.Np / kTile - jt0;
  if (T <= 0 || K <= 0) return cudaSuccess;
  return launch_syrk_i8(p, oz, K, jt0, 1, 0, T * (T + 1), B, st);
}

}  // namespace sfb
