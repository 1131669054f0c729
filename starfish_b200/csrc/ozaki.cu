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
// anti-diagonal d: 7 accumulators × 64 columns = 448 of the 512 TMEM columns for a 128×64 tile, fed by 26 digit-slab
// products per 32-deep k-chunk (pairs s,t ≤ 5, s+t ≤ 6; issued as 9 wide MMAs, oz_issue_chunk).  What is dropped is the anti-diagonals d ≥ 7 and the
// rounding of q: measured |ΔlnL|/|lnL| ≤ 1.6e-12 against the dense oracle up to cond 7e5 (tools/ozaki_experiment.py,
// tests/test_gpu_ozaki.py; 5 accumulators give 1e-7 — the anti-diagonal count, not the digit count, is what
// matters).  The epilogue recombines Σ_d 256^(−d)·S_d by Horner in fp64 and applies C −= 2^(e_i−7)·2^(e_j−7)·(…).
//
// Data layout.  The sliced panels of the current outer block live in P (per slot), pre-tiled so that every digit
// slab of an operand tile is ONE contiguous block, already in the canonical K-major shared-memory layout of the MMA:
//     P[chunk c (32 k)][slice t][row group g = row/8][row%8][32 bytes]        (NCH × 6 × Np/8 × 256 B)
// -> one digit slab of the A operand (128 rows) of a chunk is 4 KB contiguous, of the B operand (64 rows) 2 KB: one
// 1-D bulk async copy (cp.async.bulk, SASS UBLKCP) per NON-ZERO slab and pipeline stage, no tensor map.  Inside a
// 256-byte row group the two 16-byte halves of a row are XOR-swapped by bit 2 of the row (the 32-byte swizzle pattern,
// applied where the slices are written: oz_slice_block in sfb_internal.cuh, called from trsm_kernel), consecutive
// 8-row groups are 256 B apart: UMMA descriptor {SWIZZLE_32B, SBO = 256}.
// In shared memory the six B slabs of a stage lie back to back, so a run of consecutive B slabs is one operand of
// N = 64·len rows (oz_issue_chunk).
//
// Kernel shape: 352 threads = warp 0 bulk-copy producer (A), warp 1 MMA issuer (one elected lane) + TMEM allocator,
// warps 2-9 epilogue (TMEM lane quarter = warp%4, two warps per quarter), warp 10 bulk-copy producer (B); launched as
// clusters of two CTAs that share the A operand.  5-stage ring of 36 KB, full/empty mbarriers, tcgen05.commit
// releases a stage / signals the epilogue.  The epilogue warps prefetch the C tile (coalesced, into registers) while
// the main loop runs, transpose the recombined update through a padded buffer and finish the read-modify-write with
// coalesced streaming stores.
#include <algorithm>
#include <cstdint>

#include "sfb_internal.cuh"

namespace sfb {

namespace {

constexpr int OZ_S = kOzSlices;      // 6
constexpr int OZ_NACC = 7;           // anti-diagonals 0..6
constexpr int OZ_KC = kOzChunk;      // 32
constexpr int OZ_BM = 128, OZ_BN = 64;
constexpr int OZ_STAGES = 5;
constexpr int OZ_GROUP_BYTES = 8 * OZ_KC;                    // 256: 8 rows × 32 B of one slice
constexpr int OZ_ROWGROUP_BYTES = OZ_S * OZ_GROUP_BYTES;     // 1536: all slices of 8 rows
constexpr int OZ_A_BYTES = (OZ_BM / 8) * OZ_ROWGROUP_BYTES;  // 24576
constexpr int OZ_B_BYTES = (OZ_BN / 8) * OZ_ROWGROUP_BYTES;  // 12288
constexpr int OZ_STAGE_BYTES = OZ_A_BYTES + OZ_B_BYTES;      // 36864
constexpr int OZ_THREADS = 352;      // producer warp (A), MMA warp, eight epilogue warps, second producer warp (B)
constexpr int OZ_PRODB_WARP = 10;
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
// one look, never suspends (mbarrier.test_wait): used to learn a barrier's state ahead of time
__device__ __forceinline__ bool mbar_test_nb(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
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
// the same copy delivered to the same shared-memory offset of every CTA in `mask` of the cluster; each destination
// CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
// the same arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
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


// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_32B (layout code 6), descriptor version 1:
// start address, LBO (unused for swizzled K-major; canonical value 1) and SBO in 16-byte units.
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(OZ_GROUP_BYTES >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
// instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, M >> 4 at bit 24; the
// N >> 3 field (bit 17) is added per MMA (oz_idesc)
constexpr uint32_t OZ_IDESC_BASE = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);

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
// tile's accumulators to be drained, clears them with two MMAs against an all-zero A slab), warps 2-9 = epilogue.
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

constexpr int OZ_SLAB_A = OZ_BM * OZ_KC;             // 4096: one digit slab of the A operand (128 rows × 32 B)
constexpr int OZ_SLAB_B = OZ_BN * OZ_KC;             // 2048
constexpr uint32_t OZ_TROW = 18 * 8;                 // padded row of a transpose buffer (16 doubles + 16 B: conflict-free 16-byte accesses)
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_TBUF_BYTES = OZ_EPI_WARPS * 32 * OZ_TROW;     // one 32 × 16 buffer per epilogue warp
constexpr int OZ_ZERO_BYTES = OZ_SLAB_A;            // an all-zero A slab (clears the accumulators at the start of a tile)
constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + OZ_TBUF_BYTES + OZ_ZERO_BYTES + 1024;  // + alignment slack

__device__ __forceinline__ uint32_t oz_idesc(uint32_t n) { return OZ_IDESC_BASE | ((n >> 3) << 17); }

// One 32-deep chunk.  A digit slab of A (128 rows) meets every digit slab of B (64 rows) whose anti-diagonal sa + sb
// stays <= 6; the B slabs of a stage lie back to back (2 KB each, 8-row groups 256 B apart) and the accumulators of
// consecutive anti-diagonals are adjacent in TMEM, so a RUN of consecutive B slabs is ONE MMA of N = 64·len columns
// against accumulators sa+sb0 … sa+sb0+len−1.  Why it matters: a 128×64×32 int8 MMA needs 32 tensor cycles but reads
// 6 KB of shared memory (48 cycles at 128 B/clk; 52 measured), and feeding A through TMEM (tcgen05.cp, 64 B/clk, in
// the tensor pipe's issue order) costs 384 cycles per chunk on top of 26·32.  At N >= 128 the shared-memory traffic
// per MMA (4 KB + 2 KB·len) fits under its 32·len tensor cycles: 9 MMAs per chunk (3+3, 3+3, 3+2, 4, 3, 2 slabs) at the
// floor of 832 cycles, both operands from shared memory, no TMEM operand at all (tools/exp/umma_i8_probe.cu: 65.5 /
// 128.1 cycles for N = 128 / 256).  Runs of 5 and 6 slabs are split 3+2 / 3+3 so that no single-slab MMA remains.
// fa / fb: bit t set = digit slab t of the A / B operand block has a non-zero entry (and was loaded); products with
// an all-zero slab are skipped — exactly.  Every MMA accumulates (the accumulators are cleared per tile), so with
// literal masks the whole schedule folds at compile time.  Returns the number of slab products issued.
// PHASE 0 issues the MMAs of 1 and 2 slabs, PHASE 1 those of 3 and 4, each in ascending length: the tensor pipe's queue
// holds about two MMAs beyond the one executing (profiles/r3t_umma_i8_queue_slack_probe.txt: 165 cycles of issue-thread
// work between two chunks are hidden when a chunk ends with a 2- and a 3-slab MMA), so a chunk should END with its
// longest MMAs — that is the time the issue thread has to get from one chunk's last MMA to the next one's first.
template <int PHASE>
__device__ __forceinline__ void oz_issue_chunk(uint32_t fa, uint32_t fb, uint32_t tmem, uint64_t ad0, uint64_t bd0) {
#pragma unroll
  for (int want = 2 * PHASE + 1; want <= 2 * PHASE + 2; ++want) {
#pragma unroll
    for (int sa = 0; sa < OZ_S; ++sa) {
      if (!((fa >> sa) & 1u)) continue;
      const uint32_t m = fb & ((1u << (OZ_NACC - sa)) - 1u) & 0x3fu;
      const uint64_t ad = ad0 + (uint64_t)(sa * (OZ_SLAB_A >> 4));
#pragma unroll
      for (int s = 0; s < OZ_S; ++s) {
        const uint32_t b0 = (m >> s) & 1u, prev = s ? (m >> (s - 1)) & 1u : 0u;
        if (!(b0 && !prev)) continue;             // a run of set bits starts at s
        uint32_t len = 1, run = 1;
#pragma unroll
        for (int t = s + 1; t < OZ_S; ++t) {
          run &= (m >> t) & 1u;
          len += run;
        }
        const uint64_t bd = bd0 + (uint64_t)(s * (OZ_SLAB_B >> 4));
        const uint32_t d = tmem + (uint32_t)(sa + s) * OZ_BN;
        if (len <= 4) {
          if (len == (uint32_t)want) umma_i8(d, ad, bd, oz_idesc(len * OZ_BN), 1);
        } else {   // 5 -> 2 + 3, 6 -> 3 + 3
          if (len - 3 == (uint32_t)want) umma_i8(d, ad, bd, oz_idesc((len - 3) * OZ_BN), 1);
          if (want == 3)
            umma_i8(d + (len - 3) * OZ_BN, ad, bd + (uint64_t)((len - 3) * (OZ_SLAB_B >> 4)), oz_idesc(3 * OZ_BN), 1);
        }
      }
    }
  }
}
// the four patterns that make up the bench ensemble with compile-time masks, anything else generic
template <int PHASE>
__device__ __forceinline__ void oz_issue_dispatch(uint32_t fa, uint32_t fb, uint32_t tmem, uint64_t ad0, uint64_t bd0) {
  if (fa == 0u || fb == 0u) return;   // an all-zero operand block (exact zeros outside a band): nothing to issue
  if (fa == 0x3fu && fb == 0x3fu) oz_issue_chunk<PHASE>(0x3fu, 0x3fu, tmem, ad0, bd0);
  else if (fa == 0x3eu && fb == 0x3eu) oz_issue_chunk<PHASE>(0x3eu, 0x3eu, tmem, ad0, bd0);
  else if (fa == 0x3eu && fb == 0x3fu) oz_issue_chunk<PHASE>(0x3eu, 0x3fu, tmem, ad0, bd0);
  else if (fa == 0x3fu && fb == 0x3eu) oz_issue_chunk<PHASE>(0x3fu, 0x3eu, tmem, ad0, bd0);
  else oz_issue_chunk<PHASE>(fa, fb, tmem, ad0, bd0);
}
// number of slab products for digit-slab masks (fa, fb)
__device__ __forceinline__ uint32_t oz_products(uint32_t fa, uint32_t fb) {
  uint32_t n = 0;
#pragma unroll
  for (int sa = 0; sa < OZ_S; ++sa)
    if ((fa >> sa) & 1u) n += __popc(fb & ((1u << (OZ_NACC - sa)) - 1u) & 0x3fu);
  return n;
}
// accumulators that receive a product for digit-slab masks (fa, fb)
__device__ __forceinline__ uint32_t oz_touched(uint32_t fa, uint32_t fb) {
  uint32_t t = 0;
#pragma unroll
  for (int sa = 0; sa < OZ_S; ++sa)
    if ((fa >> sa) & 1u) t |= (fb & ((1u << (OZ_NACC - sa)) - 1u)) << sa;
  return t & 0x7fu;
}

// ---- epilogue (8 warps; warp (q, hh) owns tile rows 32q..32q+31 — its TMEM lane quarter — and columns 32hh..32hh+31).
// Two thread mappings: TMEM hands a thread ONE ROW (lane = row, 16 columns per load), global memory wants
// neighbouring lanes on one row (8 lanes = 128 contiguous bytes of a row, four rows per instruction).  The C tile is
// therefore fetched in the global mapping BEFORE the accumulators are ready (16 independent 16-byte loads per thread,
// in flight under the main loop), the recombined update goes through a warp-private padded transpose buffer in two
// rounds of 16 columns (a 32 × 16 buffer per warp keeps shared memory free for a fifth pipeline stage), the
// accumulators are handed back to the MMA warp, and the read-modify-write finishes with coalesced streaming stores
// while the next tile's MMAs already run.  The drain (accfull -> tmem_empty) is the serial part of a tile — the MMA
// warp waits for it — hence eight warps on it and the next accumulator's TMEM load in flight under the Horner step.
__device__ __forceinline__ void oz_drain16(uint32_t taddr, uint32_t touched, double ri, uint32_t dst) {
  double acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0;
  if (touched == 0x7fu) {
    // every accumulator holds a sum (the common case): software-pipelined loads
    uint32_t va[16], vb[16];
    tmem_ld16(taddr + (OZ_NACC - 1) * OZ_BN, va);
#pragma unroll
    for (int d = OZ_NACC - 1; d >= 0; --d) {   // Horner from the least significant anti-diagonal
      tmem_ld_wait();
      if (((OZ_NACC - 1 - d) & 1) == 0) {
        if (d > 0) tmem_ld16(taddr + (d - 1) * OZ_BN, vb);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], 0.00390625, i2d(va[j]));
      } else {
        if (d > 0) tmem_ld16(taddr + (d - 1) * OZ_BN, va);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], 0.00390625, i2d(vb[j]));
      }
    }
  } else {
    uint32_t v[16];
#pragma unroll
    for (int d = OZ_NACC - 1; d >= 0; --d) {
      if ((touched >> d) & 1u) {   // an accumulator without a product holds zeros: skip its load
        tmem_ld16(taddr + d * OZ_BN, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], 0.00390625, i2d(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] *= 0.00390625;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(dst + 16 * j), "d"(acc[2 * j] * ri), "d"(acc[2 * j + 1] * ri) : "memory");
}

__device__ __forceinline__ void oz_epilogue(const CholParams& p, const OzParams& oz, int s, int l0, int l1, int lstep, int jt0, int njt,
                                            int strip, uint32_t tmem, uint32_t tbuf0, int ew, int lane, uint32_t meta,
                                            uint32_t accfull, uint32_t tmem_empty, const volatile uint32_t* touched_p) {
  const int q = ew & 3, hh = ew >> 2;   // TMEM lane quarter (must equal warp id % 4), column half
  const double* rs = oz.rscale + (long long)s * p.Np;
  const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hh * 32);
  const uint32_t tbuf = tbuf0 + (uint32_t)ew * (32 * OZ_TROW);
  const int grow = lane >> 3, gcol = 2 * (lane & 7);   // global mapping: instruction g of a round covers rows 4g..4g+3
  const uint32_t tdst = tbuf + (uint32_t)lane * OZ_TROW;                       // TMEM mapping: my row
  const uint32_t tsrc = tbuf + (uint32_t)grow * OZ_TROW + 8 * (uint32_t)gcol;  // global mapping
  int k = 0;
  for (int l = l0; l < l1; l += lstep) {
    const OzTile t = oz_tile(l, jt0, njt, strip);
    if (!t.live) continue;
    const double ri = rs[t.r0 + q * 32 + lane];                                                 // row scale, TMEM mapping
    double* Cw = p.W + (long long)s * p.strideW + (long long)(t.r0 + q * 32 + grow) * p.Np + t.c0 + hh * 32 + gcol;
    const double2 rj0 = *reinterpret_cast<const double2*>(rs + t.c0 + hh * 32 + gcol);         // column scales, global mapping
    const double2 rj1 = *reinterpret_cast<const double2*>(rs + t.c0 + hh * 32 + 16 + gcol);
    double2 creg[16];   // [round cb][row group g]: rows 4g + grow, columns 16·cb + gcol
#pragma unroll
    for (int i = 0; i < 16; ++i)
      creg[i] = __ldcs(reinterpret_cast<const double2*>(Cw + (long long)(4 * (i & 7)) * p.Np + 16 * (i >> 3)));
    mbar_wait(meta, k & 1);
    const uint32_t touched = *touched_p;
    mbar_wait(accfull, k & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // round 0: columns 0..15 of this warp's 32
    oz_drain16(tlane, touched, ri, tdst);
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      double tx, ty;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tx), "=d"(ty) : "r"(tsrc + (uint32_t)(4 * g) * OZ_TROW) : "memory");
      creg[g].x = fma(-tx, rj0.x, creg[g].x);
      creg[g].y = fma(-ty, rj0.y, creg[g].y);
    }
    __syncwarp();
    // round 1: columns 16..31; afterwards the accumulators are drained
    oz_drain16(tlane + 16, touched, ri, tdst);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty);   // TMEM back to the MMA warp (which orders its next MMAs after this arrive)
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      double tx, ty;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tx), "=d"(ty) : "r"(tsrc + (uint32_t)(4 * g) * OZ_TROW) : "memory");
      creg[8 + g].x = fma(-tx, rj1.x, creg[8 + g].x);
      creg[8 + g].y = fma(-ty, rj1.y, creg[8 + g].y);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      __stcs(reinterpret_cast<double2*>(Cw + (long long)(4 * (i & 7)) * p.Np + 16 * (i >> 3)), creg[i]);
    __syncwarp();  // the transpose buffer is rewritten by the next tile
    ++k;
  }
}

// PAIR: launched as clusters of two CTAs that work on the two 64-column halves (l = 2m, 2m+1) of the same 128×128 tile
// in lockstep.  They need the SAME A operand, so each CTA fetches every other non-zero A slab and the copy is delivered
// to both (bulk copy with .multicast::cluster): 24 instead of 36 KB per CTA and chunk cross the L2 -> SM fabric, which
// is what bounds this kernel once the MMAs run at their floor (8.7 TB/s of operand traffic, ring of 4 stages starved:
// profiles/r2x).  A stage is refilled only after BOTH CTAs have read it: the MMA warps commit to the `empty` barrier
// of both CTAs (count 2).  tpc counts tile PAIRS per cluster then.
template <bool PAIR>
__global__ void __launch_bounds__(OZ_THREADS, 1)
    syrk_i8_kernel(CholParams p, OzParams oz, int nch, int jt0, int njt, int strip, int ntiles, int tpc) {
  const int s = blockIdx.z;
  if (p.info[s] != 0) return;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int lstep = PAIR ? 2 : 1;
  const int l0 = PAIR ? 2 * ((int)(blockIdx.x >> 1) * tpc) + (int)rank : blockIdx.x * tpc;
  const int l1 = PAIR ? min(ntiles, 2 * ((int)(blockIdx.x >> 1) * tpc + tpc)) : min(ntiles, l0 + tpc);

  extern __shared__ uint8_t oz_smem_raw[];
  __shared__ uint64_t bars[2 * OZ_STAGES + 3];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t touched_s;
  const uint32_t ring = (smem_u32(oz_smem_raw) + 1023u) & ~1023u;
  const uint32_t tbuf0 = ring + OZ_STAGES * OZ_STAGE_BYTES;
  const uint32_t zero0 = tbuf0 + OZ_TBUF_BYTES;
  const uint32_t bar0 = smem_u32(bars);  // full[s] at +8s, empty[s] at +8(STAGES+s), then accfull, tmem_empty, meta
  const uint32_t accfull = bar0 + 16 * OZ_STAGES, tmem_empty = accfull + 8, meta = accfull + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < OZ_STAGES; ++i) {
      // BOTH producers arrive on `full` every round, each with its own byte count — also when it has nothing to fetch.
      // (With a single arrival, a producer whose byte count is zero for a while — operand blocks of exact zeros in a
      // banded matrix — is not needed for the phase to complete, falls two phases behind on a stage and then misreads
      // the parity of `empty`: a deadlock that only timing had hidden.)
      mbar_init(bar0 + 8 * i, 2);
      mbar_init(bar0 + 8 * (OZ_STAGES + i), PAIR ? 2 : 1);  // a commit from the MMA warp of every CTA that reads the stage
    }
    mbar_init(accfull, 1);
    mbar_init(tmem_empty, OZ_EPI_WARPS);  // one arrival per epilogue warp
    mbar_init(meta, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < OZ_ZERO_BYTES / 16; i += OZ_THREADS)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(zero0 + 16 * i), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zero slab is read by the tensor pipe (async proxy)
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), OZ_TMEM_COLS);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before a multicast copy or commit can reach them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const int8_t* Ps = oz.P + (long long)s * oz.strideP;
  const uint8_t* Fs = oz.F + (long long)s * oz.strideF;
  const int nrb = p.Np / 64;
  const long long slice_bytes = (long long)p.Np * OZ_KC;       // one digit slab of one chunk, all rows
  const long long chunk_bytes = OZ_S * slice_bytes;

  if (warp <= 1 || warp == OZ_PRODB_WARP) {
    // the producers (warp 0: A slabs, warp 10: B slabs) and the MMA issuer (warp 1) walk the same tiles and chunks with the same digit-slab flags:
    // lane j holds the flags of chunks j and j+32 of the current tile (written with the slices by trsm_kernel), fetched once
    // per tile; whole warps run the loops (uniform control flow), one elected lane issues.
    const bool leader = elect_one();
    unsigned long long issued = 0;
#ifdef SFB_EXPERIMENTS
    long long t_full = 0, t_drain = 0;
    const long long t_begin = clock64();
#endif
    int g = 0, k = 0;  // ring position (producer), live tiles done
    int st = 0;        // ring position of the MMA warp: stage and the parity of its `full` barrier
    uint32_t par = 0;
    const uint64_t desc_ring = oz_desc(ring);
    for (int l = l0; l < l1; l += lstep) {
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
      if (warp != 1) {
        // ---- producers: one bulk copy per non-zero digit slab (4 KB of A, 2 KB of B), continuous over the CTA's tiles.
        // Issuing a copy costs the lone thread ~10 dependent instructions; twelve per chunk made ONE producer the
        // pace-setter of the whole kernel (it waited for a free stage only 20 % of its time while the MMA warp waited
        // for operands 24 % of its own).  Hence two: warp 0 fetches A, warp 10 fetches B; each posts its own byte count
        // on the stage's `full` barrier (two arrivals per phase).
        const bool prod_a = (warp == 0);
        const int8_t* src = Ps + (long long)((prod_a ? t.r0 : t.c0) / 8) * OZ_GROUP_BYTES;
        for (int c = 0; c < nch; ++c, ++g) {
          uint32_t fa = __shfl_sync(0xffffffffu, c < 32 ? fa0 : fa1, c & 31);
          uint32_t fb = __shfl_sync(0xffffffffu, c < 32 ? fb0 : fb1, c & 31);
#ifdef SFB_EXPERIMENTS
          if (oz.dbg & 1) fa = fb = 0x3fu;
          const long long te0 = clock64();
#endif
          if (g >= OZ_STAGES) mbar_wait(bar0 + 8 * (OZ_STAGES + st), par ^ 1u);
#ifdef SFB_EXPERIMENTS
          t_full += clock64() - te0;   // producer: time waiting for a free stage
#endif
          if (leader) {
            const uint32_t full = bar0 + 8 * st;
            const uint32_t dst = ring + st * OZ_STAGE_BYTES;
            if (prod_a) {
              mbar_arrive_expect_tx(full, (uint32_t)__popc(fa) * OZ_SLAB_A);
#pragma unroll
              for (int sl = 0; sl < OZ_S; ++sl) {
                if (!((fa >> sl) & 1u)) continue;
                if (!PAIR) bulk_g2s(dst + sl * OZ_SLAB_A, src + sl * slice_bytes, OZ_SLAB_A, full);
                else if ((uint32_t)(sl & 1) == rank) bulk_g2s_mc(dst + sl * OZ_SLAB_A, src + sl * slice_bytes, OZ_SLAB_A, full, 3);
              }
            } else {
              mbar_arrive_expect_tx(full, (uint32_t)__popc(fb) * OZ_SLAB_B);
#pragma unroll
              for (int sl = 0; sl < OZ_S; ++sl)
                if ((fb >> sl) & 1u) bulk_g2s(dst + OZ_A_BYTES + sl * OZ_SLAB_B, src + sl * slice_bytes, OZ_SLAB_B, full);
            }
          }
          __syncwarp();
          src += chunk_bytes;
          if (++st == OZ_STAGES) { st = 0; par ^= 1u; }
        }
      } else {
        // ---- MMA issuer
#ifdef SFB_EXPERIMENTS
        const long long tt0 = clock64();
#endif
        if (k > 0) {  // the epilogue must have drained the accumulators of the previous tile
          mbar_wait(tmem_empty, (k - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#ifdef SFB_EXPERIMENTS
        t_drain += clock64() - tt0;
#endif
        if (leader) {  // clear the seven accumulators: zero A slab × whatever the ring holds (448 columns)
          umma_i8(tmem, oz_desc(zero0), oz_desc(ring), oz_idesc(256), 0);
          umma_i8(tmem + 256, oz_desc(zero0), oz_desc(ring), oz_idesc(192), 0);
        }
        // accumulators that receive a product in this tile (the epilogue skips the others)
        uint32_t touched = __reduce_or_sync(0xffffffffu, oz_touched(fa0, fb0) | oz_touched(fa1, fb1));
        if (oz.stats) issued += __reduce_add_sync(0xffffffffu, oz_products(fa0, fb0) + oz_products(fa1, fb1));
#ifdef SFB_EXPERIMENTS
        if (oz.dbg & 1) touched = 0x7fu;
#endif
        // The loop is software-pipelined around the single issuing thread: the flags of chunk c+1 and a first,
        // non-blocking look at its `full` barrier are taken BETWEEN the short and the long MMAs of chunk c, i.e. while
        // the issue is back-pressured anyway, instead of standing between two chunks' MMAs, where the tensor pipe's
        // short queue would run dry.
        uint32_t fa = __shfl_sync(0xffffffffu, fa0, 0), fb = __shfl_sync(0xffffffffu, fb0, 0);
        bool ready = mbar_test_nb(bar0 + 8 * st, par);
        for (int c = 0; c < nch; ++c) {
#ifdef SFB_EXPERIMENTS
          if (oz.dbg & 1) fa = fb = 0x3fu;
          const long long tf0 = clock64();
#endif
          while (!ready) ready = mbar_test(bar0 + 8 * st, par);
#ifdef SFB_EXPERIMENTS
          t_full += clock64() - tf0;
#endif
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // descriptors of the stage: start-address field = (ring + st·STAGE) >> 4, never leaves its 14 bits
          const uint64_t ad0 = desc_ring + (uint64_t)((uint32_t)st * (OZ_STAGE_BYTES >> 4));
          const uint64_t bd0 = ad0 + (uint64_t)(OZ_A_BYTES >> 4);
          const uint32_t empty_bar = bar0 + 8 * (OZ_STAGES + st);
          // The digit patterns that dominate (nothing zero / leading slab zero, for either operand) get fully unrolled
          // code with compile-time masks: a single thread issues every MMA of the CTA, so run-time tests per product
          // would make the issue loop the bottleneck.  Anything else takes the generic path.
          if (leader) oz_issue_dispatch<0>(fa, fb, tmem, ad0, bd0);   // short MMAs first
          // ring position, flags and barrier state of the next chunk — while the queue is full (no division in the loop)
          const int c1 = c + 1;
          const int st1 = (st + 1 == OZ_STAGES) ? 0 : st + 1;
          const uint32_t par1 = (st + 1 == OZ_STAGES) ? par ^ 1u : par;
          uint32_t fan = 0, fbn = 0;
          bool readyn = false;
          if (c1 < nch) {
            fan = __shfl_sync(0xffffffffu, c1 < 32 ? fa0 : fa1, c1 & 31);
            fbn = __shfl_sync(0xffffffffu, c1 < 32 ? fb0 : fb1, c1 & 31);
            readyn = mbar_test_nb(bar0 + 8 * st1, par1);
          }
          if (leader) {
            oz_issue_dispatch<1>(fa, fb, tmem, ad0, bd0);               // the chunk ends with its longest MMAs
            // stage free once these MMAs have read it (in both CTAs of a pair: the peer's copies land here too)
            if (PAIR) umma_commit_mc(empty_bar, 3);
            else umma_commit(empty_bar);
          }
          __syncwarp();
          fa = fan; fb = fbn; ready = readyn; st = st1; par = par1;
        }
        if (leader) {
          touched_s = touched;
          mbar_arrive(meta);                 // release: the epilogue reads touched_s after acquiring this barrier
          umma_commit(accfull);
        }
        __syncwarp();
      }
      ++k;
    }
#ifdef SFB_EXPERIMENTS
    if (warp == OZ_PRODB_WARP && leader && oz.stats) {
      atomicAdd(oz.stats + 6, (unsigned long long)t_full);
      atomicAdd(oz.stats + 7, (unsigned long long)(clock64() - t_begin));
    }
#endif
    if (warp == 1 && leader && oz.stats) {
      atomicAdd(oz.stats, issued);
      atomicAdd(oz.stats + 1, (unsigned long long)k * nch * 26ull);
#ifdef SFB_EXPERIMENTS
      atomicAdd(oz.stats + 2, (unsigned long long)(clock64() - t_begin));
      atomicAdd(oz.stats + 3, (unsigned long long)t_full);
      atomicAdd(oz.stats + 4, (unsigned long long)t_drain);
      atomicAdd(oz.stats + 5, (unsigned long long)k);
#endif
    }
  } else {
    // warps 2..9: TMEM lane quarter = warp % 4 (hardware rule), column half = (warp − 2) / 4
    oz_epilogue(p, oz, s, l0, l1, lstep, jt0, njt, strip, tmem, tbuf0, (warp & 3) + 4 * ((warp - 2) >> 2), lane, meta, accfull,
                tmem_empty, &touched_s);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no exit while the peer may still signal this CTA's barriers
  if (warp == 1) tmem_dealloc(tmem, OZ_TMEM_COLS);
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

}  // namespace

namespace {
int g_oz_tpc = 8;      // most tiles a CTA works through
bool g_oz_pair = true; // clusters of two CTAs sharing the A operand by multicast (off: experiments build, A/B)
}
void ozaki_set_pair(bool on) { g_oz_pair = on; }
namespace { int g_oz_dbg = 0; }
void ozaki_set_debug(int d) { g_oz_dbg = d; }
void ozaki_set_tpc(int n) { g_oz_tpc = n < 1 ? 1 : n; }

cudaError_t ozaki_init() {
  cudaError_t e = cudaFuncSetAttribute((const void*)syrk_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       OZ_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)syrk_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OZ_SMEM_BYTES);
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
  OzParams ozd = oz;
  ozd.dbg = g_oz_dbg;
  long long total = (long long)ntiles * B;
  int tpc = (int)std::min<long long>(g_oz_tpc, std::max<long long>(1, total / (4LL * sms)));
  tpc = std::min(tpc, ntiles);
  if (!g_oz_pair) {
    const dim3 grid((ntiles + tpc - 1) / tpc, 1, B);
    syrk_i8_kernel<false><<<grid, OZ_THREADS, OZ_SMEM_BYTES, st>>>(p, ozd, K / OZ_KC, jt0, njt, strip, ntiles, tpc);
    return cudaGetLastError();
  }
  // clusters of two: tile l = 2m + rank; ntiles is even in both enumerations and the halves of a 128×128 tile are
  // adjacent.  tpc tile pairs per cluster = tpc tiles per CTA.
  const int npairs = ntiles / 2;
  tpc = std::min(tpc, npairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ((npairs + tpc - 1) / tpc), 1, B);
  cfg.blockDim = dim3(OZ_THREADS, 1, 1);
  cfg.dynamicSmemBytes = OZ_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, syrk_i8_kernel<true>, p, ozd, K / OZ_KC, jt0, njt, strip, ntiles, tpc);
}
}  // namespace

cudaError_t launch_oz_rowscale(const CholParams& p, const OzParams& oz, int B, cudaStream_t st) {
  oz_rowscale_kernel<<<dim3((p.Np + 255) / 256, B), 256, 0, st>>>(p, oz);
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
