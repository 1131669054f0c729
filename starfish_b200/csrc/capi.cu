// extern "C" boundary of libsfb200.so (see include/sfb200.h for the contract and the reference lines
// each entry point replaces).  Host-side orchestration only: workspace, streams, chunking of the walker
// batch over factorisation slots, event timing.  No torch types cross this boundary.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>  // header-only (NVTX v3): ranges cost nothing unless a profiler attaches

#include "sfb_internal.cuh"

using namespace sfb;

constexpr int kMaxLanes = 4;

struct ProfEvent {
  cudaEvent_t a, b;
  int cls;
  double work;
};

struct sfb_ctx {
  int device = 0, N = 0, Np = 0, M = 0, Kmax = 0, Bmax = 0, slots = 0;
  // factorisation workspace (per slot)
  double* W = nullptr;
  double* Minv = nullptr;
  double* rhs = nullptr;
  double* zk = nullptr;
  double* logdet = nullptr;
  double* sqmah = nullptr;
  int* info_ws = nullptr;
  // static data
  double* wave = nullptr;
  double* sigma = nullptr;
  double* data_flux = nullptr;
  int* sorted = nullptr;
  bool have_static = false;
  GemmMaps maps;
  // device staging for the host-buffer path
  double *dX = nullptr, *dA = nullptr, *dflux = nullptr, *dglob = nullptr, *dloc = nullptr, *dlnL = nullptr,
         *dresid = nullptr;
  int *dnloc = nullptr, *dinfo = nullptr;
  // two lanes; each lane = a main stream (bulk trailing updates, copies) + a high-priority stream for the
  // panel work that is on the critical path (look-ahead)
  cudaStream_t streams[kMaxLanes] = {};
  cudaStream_t hi[kMaxLanes] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes] = {};
  cudaEvent_t ev_hi[kMaxLanes] = {}, ev_lo[kMaxLanes] = {};
  int nlanes = 2;  // lanes the walker chunks of one call alternate over (experiments build: SFB_LANES)
  bool profile = false;
  int outer_tiles = kOuterTiles;  // (experiments build only: SFB_OUTER_TILES overrides)
  int debug_mode = 0;  // (experiments build only: SFB_DEBUG_MODE, bit0 = no high-priority stream, bit1 = single lane)
  std::vector<ProfEvent> prof;
  std::vector<cudaEvent_t> ev_pool;
  long long launches = 0;
  // structure-exploiting solver (row f4): band storage of S and the per-walker window classes
  int solver = SFB_SOLVER_DENSE;
  double* Sb = nullptr;
  int *bw_d = nullptr, *overflow = nullptr, *rowmap_d = nullptr;
  int *bw_h = nullptr, *rowmap_h = nullptr;  // pinned
  cudaEvent_t ev_band = nullptr, ev_up = nullptr;
  long long band_rows[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // walkers per window class (last = dense fallback) since creation
  // int8 tensor-core trailing update (SFB_SOLVER_DENSE_I8): sliced panels (two buffers per slot: the look-ahead
  // slices block J+1 while the bulk update of block J still reads its own) and the per-row scales
  int8_t* ozP = nullptr;
  double* oz_rscale = nullptr;
  uint8_t* ozF = nullptr;                 // non-zero digit-slab flags, same two-buffer scheme as ozP
  unsigned long long* oz_stats = nullptr; // [issued, dense-pattern] int8 MMA counts since the last read
  size_t oz_bytes = 0;    // bytes of ONE panel buffer of one slot
  size_t oz_fbytes = 0;   // bytes of ONE flag buffer of one slot
  // shared-factor path (frozen kernel groups): right-hand sides of all walkers as rows, per-panel L_kk^-1 blocks
  bool shared_factor = true;   // sfb_set_shared_factor
  double* Zt = nullptr;
  double* MinvAll = nullptr;
  int Jp_max = 0;
  FwdMaps fwd_maps;
  long long shared_factor_calls = 0;
  // multi-GPU: one NCCL communicator per handle, created by sfb_comm_init; NCCL is bound at run time
  // (dlopen) so that a host that never calls sfb_comm_init needs no NCCL, and a torch host shares torch's copy
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  ModelState model;       // upstream of the covariance (rows f1/f2); empty until sfb_set_model_host
  bool have_model = false;
  std::string err;
};

namespace {

int fail(sfb_ctx* h, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (h) {
    h->err = what;
    if (e != cudaSuccess) {
      h->err += ": ";
      h->err += cudaGetErrorString(e);
    }
  }
  return code;
}

#define SFB_CUDA(h, call)                                              \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return fail(h, SFB_ERR_CUDA, #call, e__); \
  } while (0)

struct NvtxRange {  // one named range per C-ABI call (visible in Nsight Systems / ncu --nvtx)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

cudaEvent_t get_event(sfb_ctx* h) {
  if (!h->ev_pool.empty()) {
    cudaEvent_t e = h->ev_pool.back();
    h->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {  // brackets one launch with events when profiling is on
  sfb_ctx* h;
  cudaStream_t st;
  ProfEvent pe;
  bool on;
  ProfScope(sfb_ctx* h_, cudaStream_t st_, int cls, double work) : h(h_), st(st_), on(h_->profile) {
    if (on) {
      pe.a = get_event(h);
      pe.b = get_event(h);
      pe.cls = cls;
      pe.work = work;
      cudaEventRecord(pe.a, st);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(pe.b, st);
      h->prof.push_back(pe);
    }
  }
};

// A new call may arrive on a different caller stream while the previous call's work is still running on the
// handle's lanes; per-handle scratch (the upstream stage's X/A/flux, staging buffers) is shared by the lanes, so
// lane 0 must not start the new call before lane 1 has finished the old one (and vice versa).
int order_after_previous_call(sfb_ctx* h) {
  for (int i = 0; i < kMaxLanes; ++i) SFB_CUDA(h, cudaEventRecord(h->ev_join[i], h->streams[i]));
  for (int i = 0; i < kMaxLanes; ++i)
    for (int j = 0; j < kMaxLanes; ++j)
      if (i != j) SFB_CUDA(h, cudaStreamWaitEvent(h->streams[i], h->ev_join[j], 0));
  return SFB_OK;
}

// fork: internal streams wait for everything queued on the caller's stream (and for the handle's previous call)
int fork_streams(sfb_ctx* h, cudaStream_t caller, int nstreams) {
  int rc = order_after_previous_call(h);
  if (rc != SFB_OK) return rc;
  SFB_CUDA(h, cudaEventRecord(h->ev_fork, caller));
  for (int i = 0; i < nstreams; ++i) SFB_CUDA(h, cudaStreamWaitEvent(h->streams[i], h->ev_fork, 0));
  return SFB_OK;
}
int join_streams(sfb_ctx* h, cudaStream_t caller, int nstreams) {
  for (int i = 0; i < nstreams; ++i) {
    SFB_CUDA(h, cudaEventRecord(h->ev_join[i], h->streams[i]));
    SFB_CUDA(h, cudaStreamWaitEvent(caller, h->ev_join[i], 0));
  }
  return SFB_OK;
}

// Factor `nb` matrices sitting in slots [slot0, slot0+nb); rhs/logdet/sqmah/info already initialised.
//
// Two-level blocking with one block of look-ahead.  Tile columns are grouped into outer blocks of
// kOuterTiles.  PANEL(J): inside block J each column is brought up to date left-looking (strip update by
// the block's earlier columns), factored (potrf_diag) and solved (trsm).  The trailing matrix is updated
// once per block with K = kOuterTiles·128 — split into NEXT(J), the tile columns of block J+1, and REST(J),
// everything right of them.  PANEL and NEXT run on the lane's high-priority stream, REST on its main stream:
//
//     hi :  PANEL(J) ─┬─ wait REST(J-1) ─ NEXT(J) ─ PANEL(J+1) ─┬─ ...
//     lo :            └─ REST(J) ────────────────────────────────┴─ REST(J+1) ...
//
// so the latency-bound panel work of block J+1 hides under the bulk update of block J (they touch
// disjoint tile columns).  On entry and exit all ordering is expressed on the main stream `lo`.
// Factorisation workspace: `slots` N×N fp64 matrices + the per-slot panel buffers + the TMA tensor maps over them.
// sfb_create allocates it unless asked not to (workspace_walkers < 0: a build-only handle, e.g. the kernel-builder
// seam); every entry point that factorises calls ensure_workspace first.
int alloc_workspace(sfb_ctx* h, long long slots) {
  const size_t per_slot = sizeof(double) * (size_t)h->Np * h->Np;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if (slots <= 0) {
    const size_t budget = std::min<size_t>((size_t)48 << 30, (size_t)(0.6 * (double)free_b));
    slots = (long long)std::max<size_t>(2, std::min<size_t>(1024, budget / per_slot));
    slots = std::min<long long>(slots, std::max(h->Bmax, 1));
  }
  if (slots > 1) slots -= slots % 2;  // two equal halves, one per stream
  if ((size_t)slots * per_slot > (size_t)(0.9 * (double)free_b))
    return fail(h, SFB_ERR_NOMEM, "factorisation workspace does not fit in free device memory");
  auto alloc = [&](void** p, size_t bytes) { return cudaMalloc(p, std::max<size_t>(bytes, 16)) == cudaSuccess; };
  bool ok = true;
  ok &= alloc((void**)&h->W, per_slot * slots);
  ok &= alloc((void**)&h->Minv, sizeof(double) * kTile * kTile * slots);
  ok &= alloc((void**)&h->rhs, sizeof(double) * h->Np * slots);
  ok &= alloc((void**)&h->zk, sizeof(double) * kTile * slots);
  ok &= alloc((void**)&h->logdet, sizeof(double) * slots);
  ok &= alloc((void**)&h->sqmah, sizeof(double) * slots);
  ok &= alloc((void**)&h->info_ws, sizeof(int) * slots);
  if (!ok) return fail(h, SFB_ERR_NOMEM, "factorisation workspace allocation failed");
  h->slots = (int)slots;
  if (make_gemm_maps(&h->maps, h->W, h->Np, h->Minv, h->slots) != cudaSuccess)
    return fail(h, SFB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return SFB_OK;
}

int ensure_workspace(sfb_ctx* h) { return h->W ? SFB_OK : alloc_workspace(h, 0); }

int ensure_ozaki(sfb_ctx* h) {
  if (h->ozP) return SFB_OK;
  h->oz_bytes = oz_panel_bytes_per_slot(h->Np, h->outer_tiles);
  h->oz_fbytes = oz_flag_bytes_per_slot(h->Np, h->outer_tiles);
  if (cudaMalloc((void**)&h->ozP, 2 * h->oz_bytes * (size_t)h->slots) != cudaSuccess ||
      cudaMalloc((void**)&h->oz_rscale, sizeof(double) * (size_t)h->Np * h->slots) != cudaSuccess ||
      cudaMalloc((void**)&h->ozF, 2 * h->oz_fbytes * (size_t)h->slots) != cudaSuccess ||
      cudaMalloc((void**)&h->oz_stats, 8 * sizeof(unsigned long long)) != cudaSuccess)
    return fail(h, SFB_ERR_NOMEM, "int8 solver: sliced-panel workspace allocation failed");
  SFB_CUDA(h, cudaMemset(h->oz_stats, 0, 8 * sizeof(unsigned long long)));
  SFB_CUDA(h, ozaki_init());
  return SFB_OK;
}

int run_cholesky(sfb_ctx* h, int lane, int slot0, int nb, double* lnL_out, int* info_out, double* minv_all = nullptr) {
  cudaStream_t lo = h->streams[lane];
  cudaStream_t hi = (h->profile || (h->debug_mode & 1)) ? lo : h->hi[lane];
  const bool two = (hi != lo);
  const bool i8 = (h->solver == SFB_SOLVER_DENSE_I8);
  CholParams p;
  p.Np = h->Np;
  p.strideW = (long long)h->Np * h->Np;
  p.W = h->W + (long long)slot0 * p.strideW;
  p.Minv = h->Minv + (long long)slot0 * kTile * kTile;
  p.rhs = h->rhs + (long long)slot0 * h->Np;
  p.zk = h->zk + (long long)slot0 * kTile;
  p.logdet = h->logdet + slot0;
  p.sqmah = h->sqmah + slot0;
  p.info = h->info_ws + slot0;
  p.k0 = 0;
  const int nt = h->Np / kTile;
  OzParams oz{nullptr, 0, nullptr, nullptr, 0, nullptr, 0};
  if (i8) {
    int rc = ensure_ozaki(h);
    if (rc != SFB_OK) return rc;
    oz.strideP = (long long)(2 * h->oz_bytes);
    oz.strideF = (long long)(2 * h->oz_fbytes);
    oz.stats = h->oz_stats;
    oz.rscale = h->oz_rscale + (long long)slot0 * h->Np;
    SFB_CUDA(h, launch_oz_rowscale(p, oz, nb, lo));  // from the diagonal of the matrix as built
    h->launches++;
  }
  if (two) {  // hi starts after whatever produced the matrices on lo (build / copy-in)
    SFB_CUDA(h, cudaEventRecord(h->ev_lo[lane], lo));
    SFB_CUDA(h, cudaStreamWaitEvent(hi, h->ev_lo[lane], 0));
  }
  const int OT = h->outer_tiles;
  // update of tile columns [jt0, jt0+njt) by the K columns starting at kb (strip), or of everything right of jt0 (tri)
  auto strip = [&](int kb, int K, int jt0, int njt, cudaStream_t st) {
    return i8 ? launch_oz_syrk_strip(p, oz, K, jt0, njt, nb, st)
              : launch_syrk_strip(p, h->maps, slot0, kb, K, jt0, njt, nb, st);
  };
  auto tri = [&](int kb, int K, int jt0, cudaStream_t st) {
    return i8 ? launch_oz_syrk_tri(p, oz, K, jt0, nb, st) : launch_syrk_tri(p, h->maps, slot0, kb, K, jt0, nb, st);
  };
  for (int J0 = 0, blk = 0; J0 < nt; J0 += OT, ++blk) {
    const int q = std::min(OT, nt - J0);
    const int kb = J0 * kTile;
    if (i8) {
      oz.P = h->ozP + (size_t)slot0 * 2 * h->oz_bytes + (size_t)(blk & 1) * h->oz_bytes;
      oz.F = h->ozF + (size_t)slot0 * 2 * h->oz_fbytes + (size_t)(blk & 1) * h->oz_fbytes;
    }
    // ---- PANEL(J) on hi
    for (int c = 0; c < q; ++c) {
      const int jt = J0 + c;
      p.k0 = jt * kTile;
      const int last = (jt == nt - 1);
      const double rem = (double)(h->Np - p.k0 - kTile);
      if (c > 0) {
        const double rows = (double)(h->Np - p.k0), K = (double)c * kTile;
        ProfScope ps(h, hi, SFB_K_SYRK, nb * (2.0 * K * kTile * rows - K * kTile * (kTile - 1.0)));
        SFB_CUDA(h, strip(kb, c * kTile, jt, 1, hi));
        h->launches++;
      }
      {
        ProfScope ps(h, hi, SFB_K_POTRF_DIAG, nb * (2.0 * kTile * kTile * kTile / 3.0));
        SFB_CUDA(h, launch_potrf_diag(p, nb, last, lnL_out, info_out, hi));
        h->launches++;
        if (minv_all)  // shared-factor path (nb == 1): keep every panel's L_kk^-1 for the forward substitution
          SFB_CUDA(h, cudaMemcpyAsync(minv_all + (size_t)jt * kTile * kTile, p.Minv, sizeof(double) * kTile * kTile,
                                      cudaMemcpyDeviceToDevice, hi));
      }
      if (!last) {
        {
          // trsm; in int8 mode the same kernel writes the fixed-point restatement of the panel it just solved ->
          // chunks [4c, 4c+4) of this block's buffer
          ProfScope ps(h, hi, SFB_K_TRSM, nb * (rem * kTile * kTile));  // useful FLOPs of a triangular solve
          if (i8) SFB_CUDA(h, launch_trsm_slice(p, h->maps, oz, c * (kTile / kOzChunk), slot0, nb, hi));
          else SFB_CUDA(h, launch_trsm(p, h->maps, slot0, nb, hi));
          h->launches++;
        }
      }
    }
    const int jn = J0 + q;  // first tile column right of this block
    if (jn >= nt) break;
    const int qn = std::min(OT, nt - jn);  // width of the next block
    const double K = (double)q * kTile;
    if (two) {
      SFB_CUDA(h, cudaEventRecord(h->ev_hi[lane], hi));        // PANEL(J) done
      SFB_CUDA(h, cudaStreamWaitEvent(lo, h->ev_hi[lane], 0));
      SFB_CUDA(h, cudaStreamWaitEvent(hi, h->ev_lo[lane], 0));  // REST(J-1) done (or the build, for J=0)
    }
    {  // ---- NEXT(J): tile columns [jn, jn+qn), on hi
      const double rows = (double)(h->Np - jn * kTile), w = (double)qn * kTile;
      ProfScope ps(h, hi, SFB_K_SYRK, nb * (2.0 * K * w * rows - K * w * (w - 1.0)));
      SFB_CUDA(h, strip(kb, q * kTile, jn, qn, hi));
      h->launches++;
    }
    if (jn + qn < nt) {  // ---- REST(J): everything right of the next block, on lo
      const double n = (double)(h->Np - (jn + qn) * kTile);
      ProfScope ps(h, lo, SFB_K_SYRK, nb * (n * (n + 1.0) * K));  // algorithmic syrk FLOPs: n(n+1)k
      SFB_CUDA(h, tri(kb, q * kTile, jn + qn, lo));
      h->launches++;
    }
    if (two) SFB_CUDA(h, cudaEventRecord(h->ev_lo[lane], lo));
  }
  if (two) {  // fold hi back into lo
    SFB_CUDA(h, cudaEventRecord(h->ev_hi[lane], hi));
    SFB_CUDA(h, cudaStreamWaitEvent(lo, h->ev_hi[lane], 0));
  }
  return SFB_OK;
}

int check_batch(sfb_ctx* h, int B) {
  if (!h) return SFB_ERR_ARG;
  if (B < 0 || B > h->Bmax) return fail(h, SFB_ERR_ARG, "B out of range (0..Bmax)");
  return SFB_OK;
}

BuildParams make_build_params(sfb_ctx* h, const double* X, const double* A, const double* glob, const int* nloc,
                              const double* loc, int shared_hyper, double jitter) {
  BuildParams bp;
  bp.N = h->N;
  bp.M = (X != nullptr) ? h->M : 0;
  bp.Kmax = h->Kmax;
  bp.hyper_stride = shared_hyper ? 0 : 1;
  bp.jitter = jitter;
  bp.wave = h->wave;
  bp.sigma = h->sigma;
  bp.X = X;
  bp.A = A;
  bp.glob = glob;
  bp.nloc = nloc;
  bp.loc = loc;
  bp.sorted = h->sorted;
  bp.bulk_ok = 0;
  return bp;
}

// the device-resident log-likelihood over walkers [0,B), chunked over the slots of `nstreams` streams
int loglike_device(sfb_ctx* h, int B, const double* X, const double* A, const double* model_flux,
                   const double* glob, const int* nloc, const double* loc, int shared_hyper, double* lnL,
                   int* info, double* resid, int nstreams,
                   // optional host staging (host path): copies are issued per chunk on the chunk's stream
                   const double* X_h, const double* A_h, const double* flux_h, const double* glob_h,
                   const int* nloc_h, const double* loc_h, double* lnL_h, int* info_h, double* resid_h) {
  const int N = h->N, M = h->M, K = h->Kmax;
  const int per_max = std::max(1, h->slots / nstreams);
  const int nchunks = (B + per_max - 1) / per_max;
  const int per = (B + nchunks - 1) / nchunks;  // balanced chunks: no small tail chunk
  const bool host = (flux_h != nullptr);
  if (host && shared_hyper) {  // one shared hyper-parameter row: copy once, up front, on stream 0
    SFB_CUDA(h, cudaMemcpyAsync((void*)glob, glob_h, sizeof(double) * 2, cudaMemcpyHostToDevice, h->streams[0]));
    SFB_CUDA(h, cudaMemcpyAsync((void*)nloc, nloc_h, sizeof(int), cudaMemcpyHostToDevice, h->streams[0]));
    if (K > 0)
      SFB_CUDA(h, cudaMemcpyAsync((void*)loc, loc_h, sizeof(double) * 3 * K, cudaMemcpyHostToDevice, h->streams[0]));
    SFB_CUDA(h, cudaEventRecord(h->ev_fork, h->streams[0]));
    for (int i = 1; i < nstreams; ++i) SFB_CUDA(h, cudaStreamWaitEvent(h->streams[i], h->ev_fork, 0));
  }
  int chunk = 0;
  for (int b0 = 0; b0 < B; b0 += per, ++chunk) {
    const int nb = std::min(per, B - b0);
    const int si = chunk % nstreams;
    cudaStream_t st = h->streams[si];
    const int slot0 = si * per_max;
    const int hb = shared_hyper ? 0 : b0;
    if (host) {
      if (X_h)
        SFB_CUDA(h, cudaMemcpyAsync((void*)(X + (long long)b0 * M * N), X_h + (long long)b0 * M * N,
                                    sizeof(double) * (size_t)nb * M * N, cudaMemcpyHostToDevice, st));
      if (X_h)
        SFB_CUDA(h, cudaMemcpyAsync((void*)(A + (long long)b0 * M * M), A_h + (long long)b0 * M * M,
                                    sizeof(double) * (size_t)nb * M * M, cudaMemcpyHostToDevice, st));
      SFB_CUDA(h, cudaMemcpyAsync((void*)(model_flux + (long long)b0 * N), flux_h + (long long)b0 * N,
                                  sizeof(double) * (size_t)nb * N, cudaMemcpyHostToDevice, st));
      if (!shared_hyper) {
        SFB_CUDA(h, cudaMemcpyAsync((void*)(glob + 2LL * b0), glob_h + 2LL * b0, sizeof(double) * 2 * nb,
                                    cudaMemcpyHostToDevice, st));
        SFB_CUDA(h, cudaMemcpyAsync((void*)(nloc + b0), nloc_h + b0, sizeof(int) * nb, cudaMemcpyHostToDevice, st));
        if (K > 0)
          SFB_CUDA(h, cudaMemcpyAsync((void*)(loc + 3LL * K * b0), loc_h + 3LL * K * b0,
                                      sizeof(double) * 3 * K * nb, cudaMemcpyHostToDevice, st));
      }
    }
    SFB_CUDA(h, launch_residual(model_flux + (long long)b0 * N, h->data_flux, N, h->Np, nb,
                                h->rhs + (long long)slot0 * h->Np, resid ? resid + (long long)b0 * N : nullptr,
                                h->logdet + slot0, h->sqmah + slot0, h->info_ws + slot0, st));
    h->launches++;
    BuildParams bp = make_build_params(h, X ? X + (long long)b0 * M * N : nullptr,
                                       A ? A + (long long)b0 * M * M : nullptr, glob + 2LL * hb, nloc + hb,
                                       loc + 3LL * K * hb, shared_hyper, 1e-10);
    bp.ldc = h->Np;
    bp.strideC = (long long)h->Np * h->Np;
    bp.padN = h->Np;
    bp.lower_only = 1;
    bp.vec2 = 1;
    bp.C = h->W + (long long)slot0 * bp.strideC;
    {
      ProfScope ps(h, st, SFB_K_BUILD, nb * (4.0 * h->Np * ((double)h->Np + kTile)));  // lower tiles, 8 B each
      SFB_CUDA(h, launch_cov_build(bp, nb, st));
      h->launches++;
    }
    int rc = run_cholesky(h, si, slot0, nb, lnL + b0, info + b0);
    if (rc != SFB_OK) return rc;
    if (host) {
      SFB_CUDA(h, cudaMemcpyAsync(lnL_h + b0, lnL + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
      SFB_CUDA(h, cudaMemcpyAsync(info_h + b0, info + b0, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
      if (resid_h)
        SFB_CUDA(h, cudaMemcpyAsync(resid_h + (long long)b0 * N, resid + (long long)b0 * N,
                                    sizeof(double) * (size_t)nb * N, cudaMemcpyDeviceToHost, st));
    }
  }
  return SFB_OK;
}


// ---------------------------------------------------------------------------------------------------
// structure-exploiting solver (row f4).  One host round trip of B ints (the exact half-bandwidth of every
// walker's S) decides each walker's window class; every class is one band_build + one band_chol launch
// over a row map; walkers whose band exceeds the widest window, or an unsorted grid, take the dense path.
// ---------------------------------------------------------------------------------------------------
struct UpstreamJob {  // parameters whose upstream stage should run concurrently with the band set-up
  const double* theta;
  int ncheb;
  double* log_scale_out;
};
int run_upstream(sfb_ctx* h, int B, const double* theta, int ncheb, double* X, double* A, double* flux,
                 double* log_scale_out, int* status, cudaStream_t st);

int structured_device(sfb_ctx* h, int B, const double* X, const double* A, const double* model_flux,
                      const double* glob, const int* nloc, const double* loc, int shared_hyper, double* lnL,
                      int* info, double* resid, cudaStream_t caller, const UpstreamJob* up = nullptr) {
  const int N = h->N, K = h->Kmax, Bm = h->Bmax;
  const int WDmax = kBandWidths[kNumBandWidths - 1];
  if (!h->Sb) {
    bool ok = cudaMalloc((void**)&h->Sb, sizeof(double) * (size_t)Bm * N * WDmax) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->bw_d, sizeof(int) * (Bm + 1)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->overflow, sizeof(int) * Bm) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->rowmap_d, sizeof(int) * Bm) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&h->bw_h, sizeof(int) * (Bm + 1)) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&h->rowmap_h, sizeof(int) * Bm) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_band, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return fail(h, SFB_ERR_NOMEM, "structured solver: band storage allocation failed");
  }
  cudaStream_t s0 = h->streams[0], s1 = h->streams[1];
  int rc = fork_streams(h, caller, 2);
  if (rc != SFB_OK) return rc;
  const int hs = shared_hyper ? 0 : 1;
  // 0. the transforms and the emulator (when asked for) run on lane 1 while lane 0 and the host size up the
  //    bands: band widths and band builds need only the kernel hyper-parameters, not X or the model flux
  if (up) {
    rc = run_upstream(h, B, up->theta, up->ncheb, (double*)X, (double*)A, (double*)model_flux, up->log_scale_out,
                      h->model.status, s1);
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, cudaEventRecord(h->ev_up, s1));
  }
  // 1. exact half-bandwidths + the sortedness flag, one small D2H
  SFB_CUDA(h, launch_band_width(N, K, hs, h->wave, glob, nloc, loc, h->bw_d, B, s0));
  h->launches++;
  SFB_CUDA(h, cudaMemcpyAsync(h->bw_h, h->bw_d, sizeof(int) * B, cudaMemcpyDeviceToHost, s0));
  SFB_CUDA(h, cudaMemcpyAsync(h->bw_h + Bm, h->sorted, sizeof(int), cudaMemcpyDeviceToHost, s0));
  if (resid) SFB_CUDA(h, launch_residual_only(model_flux, h->data_flux, N, B, resid, s1));
  SFB_CUDA(h, cudaStreamSynchronize(s0));
  const bool sorted = h->bw_h[Bm] != 0;
  // 2. classes -> row map
  int count[8] = {0}, start[8] = {0};
  const int nclass = kNumBandWidths;  // class nclass = dense
  auto cls = [&](int b) {
    if (!sorted) return nclass;
    const int nr = (X ? h->M : 0) + 1;
    for (int c = 0; c < nclass; ++c) {
      if (kBandWidths[c] == 64 && nr * nr > 256) continue;  // the 64-pixel window's 256 threads hold the Gram matrix
      if (h->bw_h[b] + band_slack(kBandWidths[c]) <= kBandWidths[c]) return c;
    }
    return nclass;
  };
  for (int b = 0; b < B; ++b) count[cls(b)]++;
  for (int c = 1; c <= nclass; ++c) start[c] = start[c - 1] + count[c - 1];
  {
    int pos[8];
    for (int c = 0; c <= nclass; ++c) pos[c] = start[c];
    for (int b = 0; b < B; ++b) h->rowmap_h[pos[cls(b)]++] = b;
  }
  SFB_CUDA(h, cudaMemcpyAsync(h->rowmap_d, h->rowmap_h, sizeof(int) * B, cudaMemcpyHostToDevice, s0));
  SFB_CUDA(h, cudaMemsetAsync(h->overflow, 0, sizeof(int) * B, s0));
  SFB_CUDA(h, cudaEventRecord(h->ev_band, s0));
  // 3. one build + one factorisation launch per class, widest window first (its CTAs run longest), each
  //    class on its own stream so that the classes' CTAs fill the SMs together
  cudaStream_t lanes[4] = {h->hi[0], s0, s1, h->hi[1]};
  for (int i = 0; i < 4; ++i)
    if (lanes[i] != s0) SFB_CUDA(h, cudaStreamWaitEvent(lanes[i], h->ev_band, 0));
  int lane = 0;
  bool used[4] = {false, false, false, false};
  for (int c = nclass - 1; c >= 0; --c) {
    if (!count[c]) continue;
    h->band_rows[c] += count[c];
    const int li = h->profile ? 1 : (lane++ & 3);
    cudaStream_t st = lanes[li];
    used[li] = true;
    const int WD = kBandWidths[c];
    BandBuildParams bp;
    bp.N = N; bp.WD = WD; bp.Kmax = K; bp.hyper_stride = hs; bp.jitter = 1e-10;
    bp.wave = h->wave; bp.sigma = h->sigma; bp.glob = glob; bp.nloc = nloc; bp.loc = loc;
    bp.Sb = h->Sb; bp.strideSb = (long long)N * WDmax; bp.overflow = h->overflow;
    bp.rowmap = h->rowmap_d + start[c];
    {
      ProfScope ps(h, st, SFB_K_BAND_BUILD, (double)count[c] * 8.0 * N * WD);
      SFB_CUDA(h, launch_band_build(bp, count[c], st));
    }
    BandCholParams cp;
    cp.N = N; cp.M = X ? h->M : 0; cp.Sb = h->Sb; cp.strideSb = bp.strideSb; cp.rowmap = bp.rowmap;
    cp.X = X; cp.A = A; cp.model_flux = model_flux; cp.data_flux = h->data_flux; cp.overflow = h->overflow;
    cp.sorted = h->sorted; cp.lnL = lnL; cp.info = info;
    double flops = 0.0;
    for (int q = 0; q < count[c]; ++q) {
      const double bwq = h->bw_h[h->rowmap_h[start[c] + q]];
      flops += (double)N * (bwq * bwq + 2.0 * bwq * (cp.M + 1));
    }
    if (up && st != s1) SFB_CUDA(h, cudaStreamWaitEvent(st, h->ev_up, 0));  // X, A, model flux ready
    {
      ProfScope ps(h, st, SFB_K_BAND_CHOL, flops);
      SFB_CUDA(h, launch_band_chol(cp, WD, count[c], st));
    }
    h->launches += 2;
  }
  // the two high-priority streams fold back into the lanes the caller is joined with
  if (used[0]) { SFB_CUDA(h, cudaEventRecord(h->ev_hi[0], h->hi[0])); SFB_CUDA(h, cudaStreamWaitEvent(s0, h->ev_hi[0], 0)); }
  if (used[3]) { SFB_CUDA(h, cudaEventRecord(h->ev_hi[1], h->hi[1])); SFB_CUDA(h, cudaStreamWaitEvent(s1, h->ev_hi[1], 0)); }
  // 4. dense path for what does not fit a window (contiguous runs of the original order)
  if (count[nclass] && up) SFB_CUDA(h, cudaStreamWaitEvent(s0, h->ev_up, 0));
  if (count[nclass]) {
    h->band_rows[nclass] += count[nclass];
    const int nstreams = (h->profile || h->slots < 2 || (h->debug_mode & 2)) ? 1 : std::min(h->nlanes, h->slots);
    const int M = h->M;
    int q = 0;
    while (q < count[nclass]) {
      const int b0 = h->rowmap_h[start[nclass] + q];
      int len = 1;
      while (q + len < count[nclass] && h->rowmap_h[start[nclass] + q + len] == b0 + len) ++len;
      const int hb = shared_hyper ? 0 : b0;
      rc = loglike_device(h, len, X ? X + (long long)b0 * M * N : nullptr, A ? A + (long long)b0 * M * M : nullptr,
                          model_flux + (long long)b0 * N, glob + 2LL * hb, nloc + hb, loc + 3LL * K * hb,
                          shared_hyper, lnL + b0, info + b0, nullptr, nstreams, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr, nullptr, nullptr);
      if (rc != SFB_OK) return rc;
      q += len;
    }
  }
  return join_streams(h, caller, 2);
}

// ---------------------------------------------------------------------------------------------------
// Shared-factor path: all walkers share the kernel hyper-parameters (the reference's frozen-group cache,
// spectrum_model.py:341-363), so S = diag(σ²+jitter) + K_global + ΣK_local is built and factorised ONCE;
// per walker only Z_b = L⁻¹[R_b | X_bᵀ] (N²(M+1) FLOP instead of N³/3) and the M×M capacitance system remain:
//   log det C_b = log det S + log det(I + A_b G_b),   R_bᵀC_b⁻¹R_b = ‖z_R‖² − uᵀ(I + A_b G_b)⁻¹A_b u,
//   G_b = Z_XᵀZ_X,  u = Z_Xᵀz_R   (determinant lemma / Woodbury; the same identity the structured solver uses).
// ---------------------------------------------------------------------------------------------------
int shared_factor_device(sfb_ctx* h, int B, const double* X, const double* A, const double* model_flux,
                         const double* glob, const int* nloc, const double* loc, double* lnL, int* info, double* resid,
                         cudaStream_t caller) {
  const int N = h->N, Np = h->Np, M = X ? h->M : 0;
  const int J = B * (M + 1), Jp = ((J + 127) / 128) * 128;
  const int nt = Np / kTile;
  if (!h->Zt || Jp > h->Jp_max) {
    if (h->Zt) { cudaFree(h->Zt); free_fwd_maps(&h->fwd_maps); h->Zt = nullptr; }
    const int want = ((h->Bmax * (h->M + 1) + 127) / 128) * 128;
    if (cudaMalloc((void**)&h->Zt, sizeof(double) * (size_t)want * Np) != cudaSuccess)
      return fail(h, SFB_ERR_NOMEM, "shared-factor path: right-hand-side workspace allocation failed");
    if (!h->MinvAll && cudaMalloc((void**)&h->MinvAll, sizeof(double) * (size_t)nt * kTile * kTile) != cudaSuccess)
      return fail(h, SFB_ERR_NOMEM, "shared-factor path: L_kk^-1 workspace allocation failed");
    h->Jp_max = want;
    SFB_CUDA(h, make_fwd_maps(&h->fwd_maps, h->Zt, Np, want, h->MinvAll, nt));
  }
  int rc = fork_streams(h, caller, 1);
  if (rc != SFB_OK) return rc;
  cudaStream_t st = h->streams[0];
  // 1. S into slot 0 (no emulator term), factorised with the handle's dense solver (fp64 or int8 trailing update)
  BuildParams bp = make_build_params(h, nullptr, nullptr, glob, nloc, loc, 1, 1e-10);
  bp.ldc = Np;
  bp.strideC = (long long)Np * Np;
  bp.padN = Np;
  bp.lower_only = 1;
  bp.vec2 = 1;
  bp.C = h->W;
  SFB_CUDA(h, launch_residual(nullptr, nullptr, N, Np, 1, h->rhs, nullptr, h->logdet, h->sqmah, h->info_ws, st));
  {
    ProfScope ps(h, st, SFB_K_BUILD, 4.0 * Np * ((double)Np + kTile));
    SFB_CUDA(h, launch_cov_build(bp, 1, st));
  }
  h->launches += 2;
  if ((rc = run_cholesky(h, 0, 0, 1, nullptr, nullptr, h->MinvAll)) != SFB_OK) return rc;
  // 2. right-hand sides of every walker as rows, 3. Z = L^-1 RHS panel by panel
  SFB_CUDA(h, launch_pack_rhs(model_flux, h->data_flux, X, N, Np, M, J, Jp, h->Zt, resid, st));
  h->launches++;
  {
    ProfScope ps(h, st, SFB_K_FWD_ROWS, (double)J * Np * Np);
    SFB_CUDA(h, launch_forward_rows(h->fwd_maps, h->maps, h->Zt, Np, Jp, 0, st, &h->launches));
  }
  // 4. Gram matrices + capacitance systems -> lnL, info
  SFB_CUDA(h, launch_gram_capacitance(h->Zt, Np, N, M, B, A, h->logdet, h->info_ws, lnL, info, st));
  h->launches++;
  h->shared_factor_calls++;
  return join_streams(h, caller, 1);
}

bool use_shared_factor(const sfb_ctx* h, int B, int shared_hyper) {
  return shared_hyper && h->shared_factor && B >= 2 && h->solver != SFB_SOLVER_STRUCTURED;
}

}  // namespace

extern "C" {

int sfb_abi_version(void) { return 5; }

int sfb_create(int device, int N, int M, int Kmax, int Bmax, int workspace_walkers, sfb_t** out) {
  if (!out) return SFB_ERR_ARG;
  *out = nullptr;
  if (N < 1 || M < 0 || M > kMaxM || Kmax < 0 || Kmax > kMaxK || Bmax < 1) return SFB_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SFB_ERR_CUDA;
  sfb_ctx* h = new sfb_ctx();
  h->device = device;
  h->N = N;
  h->Np = ((N + kTile - 1) / kTile) * kTile;
  h->M = M;
  h->Kmax = std::max(Kmax, 1);
  h->Bmax = Bmax;
#ifdef SFB_EXPERIMENTS  // A/B knobs exist only in an experiments build (python -m starfish_b200.build --experiments)
  if (const char* dbg = getenv("SFB_DEBUG_MODE")) h->debug_mode = atoi(dbg);
  if (const char* ot = getenv("SFB_OUTER_TILES")) h->outer_tiles = std::max(1, atoi(ot));
  if (const char* v = getenv("SFB_OZ_TPC")) ozaki_set_tpc(atoi(v));
  if (const char* v = getenv("SFB_OZ_PAIR")) ozaki_set_pair(atoi(v) != 0);
  if (const char* v = getenv("SFB_OZ_DEBUG")) ozaki_set_debug(atoi(v));
  if (const char* v = getenv("SFB_POTRF_BLOCKED")) potrf_set_blocked(atoi(v) != 0);
  if (const char* v = getenv("SFB_LANES")) h->nlanes = std::max(1, std::min(kMaxLanes, atoi(v)));
#endif
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    delete h;
    return SFB_ERR_CUDA;  // sm_100a only
  }
  if (workspace_walkers >= 0) {
    const int rc = alloc_workspace(h, workspace_walkers);
    if (rc != SFB_OK) {
      sfb_destroy(h);
      return rc;
    }
  }
  auto alloc = [&](void** p, size_t bytes) { return cudaMalloc(p, std::max<size_t>(bytes, 16)) == cudaSuccess; };
  bool ok = true;
  ok &= alloc((void**)&h->wave, sizeof(double) * N);
  ok &= alloc((void**)&h->sigma, sizeof(double) * N);
  ok &= alloc((void**)&h->data_flux, sizeof(double) * N);
  ok &= alloc((void**)&h->sorted, sizeof(int));
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i < kMaxLanes && ok; ++i) {
    ok &= cudaStreamCreateWithPriority(&h->streams[i], cudaStreamNonBlocking, prio_lo) == cudaSuccess;
    ok &= cudaStreamCreateWithPriority(&h->hi[i], cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    ok &= cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    ok &= cudaEventCreateWithFlags(&h->ev_hi[i], cudaEventDisableTiming) == cudaSuccess;
    ok &= cudaEventCreateWithFlags(&h->ev_lo[i], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && kernels_init() == cudaSuccess;
  ok = ok && upstream_init() == cudaSuccess;
  if (!ok) {
    sfb_destroy(h);
    return SFB_ERR_NOMEM;
  }
  *out = h;
  return SFB_OK;
}

int sfb_destroy(sfb_t* h) {
  if (!h) return SFB_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {h->W,     h->Minv, h->rhs,  h->zk,   h->logdet, h->sqmah, h->info_ws, h->wave,  h->sigma,
                  h->data_flux, h->sorted, h->dX, h->dA, h->dflux, h->dglob, h->dloc, h->dlnL, h->dresid,
                  h->dnloc, h->dinfo};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < kMaxLanes; ++i) {
    if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    if (h->hi[i]) cudaStreamDestroy(h->hi[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (h->ev_hi[i]) cudaEventDestroy(h->ev_hi[i]);
    if (h->ev_lo[i]) cudaEventDestroy(h->ev_lo[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  free_gemm_maps(&h->maps);
  model_free(&h->model);
  sfb_comm_destroy(h);
  if (h->Zt) cudaFree(h->Zt);
  if (h->MinvAll) cudaFree(h->MinvAll);
  free_fwd_maps(&h->fwd_maps);
  if (h->ozP) cudaFree(h->ozP);
  if (h->ozF) cudaFree(h->ozF);
  if (h->oz_stats) cudaFree(h->oz_stats);
  if (h->oz_rscale) cudaFree(h->oz_rscale);
  if (h->Sb) cudaFree(h->Sb);
  if (h->bw_d) cudaFree(h->bw_d);
  if (h->overflow) cudaFree(h->overflow);
  if (h->rowmap_d) cudaFree(h->rowmap_d);
  if (h->bw_h) cudaFreeHost(h->bw_h);
  if (h->rowmap_h) cudaFreeHost(h->rowmap_h);
  if (h->ev_band) cudaEventDestroy(h->ev_band);
  if (h->ev_up) cudaEventDestroy(h->ev_up);
  for (auto& pe : h->prof) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  delete h;
  return SFB_OK;
}

int sfb_set_static(sfb_t* h, const double* wave, const double* sigma, const double* data_flux, void* stream) {
  NvtxRange nvtx_range("sfb_set_static");
  if (!h || !wave || !sigma || !data_flux) return fail(h, SFB_ERR_ARG, "sfb_set_static: NULL argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  // earlier work on the handle's own streams may still read the old static data
  for (int i = 0; i < kMaxLanes; ++i) {
    SFB_CUDA(h, cudaEventRecord(h->ev_join[i], h->streams[i]));
    SFB_CUDA(h, cudaStreamWaitEvent(st, h->ev_join[i], 0));
  }
  const size_t nb = sizeof(double) * h->N;
  SFB_CUDA(h, cudaMemcpyAsync(h->wave, wave, nb, cudaMemcpyDeviceToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->sigma, sigma, nb, cudaMemcpyDeviceToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->data_flux, data_flux, nb, cudaMemcpyDeviceToDevice, st));
  SFB_CUDA(h, launch_check_sorted(h->wave, h->N, h->sorted, st));
  h->launches += 2;
  h->have_static = true;
  return SFB_OK;
}

int sfb_set_static_host(sfb_t* h, const double* wave_h, const double* sigma_h, const double* data_flux_h) {
  NvtxRange nvtx_range("sfb_set_static_host");
  if (!h || !wave_h || !sigma_h || !data_flux_h) return fail(h, SFB_ERR_ARG, "sfb_set_static_host: NULL argument");
  DeviceGuard guard(h->device);
  for (int i = 0; i < kMaxLanes; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
  const size_t nb = sizeof(double) * h->N;
  cudaStream_t st = h->streams[0];
  SFB_CUDA(h, cudaMemcpyAsync(h->wave, wave_h, nb, cudaMemcpyHostToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->sigma, sigma_h, nb, cudaMemcpyHostToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->data_flux, data_flux_h, nb, cudaMemcpyHostToDevice, st));
  SFB_CUDA(h, launch_check_sorted(h->wave, h->N, h->sorted, st));
  h->launches += 2;
  SFB_CUDA(h, cudaStreamSynchronize(st));
  h->have_static = true;
  return SFB_OK;
}

int sfb_build_cov(sfb_t* h, int B, const double* X, const double* A, const double* glob, const int* nloc,
                  const double* loc, int shared_hyper, double jitter, double* C, void* stream) {
  NvtxRange nvtx_range("sfb_build_cov");
  int rc = check_batch(h, B);
  if (rc != SFB_OK) return rc;
  if (!h->have_static) return fail(h, SFB_ERR_STATE, "sfb_build_cov: call sfb_set_static first");
  if (!C || !glob || !nloc || !loc || (X && !A)) return fail(h, SFB_ERR_ARG, "sfb_build_cov: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  cudaStream_t caller = (cudaStream_t)stream, st = h->streams[0];
  if ((rc = fork_streams(h, caller, 1)) != SFB_OK) return rc;
  BuildParams bp = make_build_params(h, X, A, glob, nloc, loc, shared_hyper, jitter);
  bp.ldc = h->N;
  bp.strideC = (long long)h->N * h->N;
  bp.padN = h->N;
  bp.lower_only = 0;
  bp.vec2 = ((h->N % 2) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  bp.C = C;
  {
    ProfScope ps(h, st, SFB_K_BUILD, (double)B * 8.0 * h->N * h->N);
    SFB_CUDA(h, launch_cov_build(bp, B, st));
    h->launches++;
  }
  return join_streams(h, caller, 1);
}

int sfb_potrf(sfb_t* h, int B, double* C, int* info, double* logdet, void* stream) {
  NvtxRange nvtx_range("sfb_potrf");
  int rc = check_batch(h, B);
  if (rc != SFB_OK) return rc;
  if (!C || !info) return fail(h, SFB_ERR_ARG, "sfb_potrf: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h)) != SFB_OK) return rc;
  cudaStream_t caller = (cudaStream_t)stream;
  const int nstreams = (h->profile || h->slots < 2 || (h->debug_mode & 2)) ? 1 : std::min(h->nlanes, h->slots);
  if ((rc = fork_streams(h, caller, nstreams)) != SFB_OK) return rc;
  const int per = std::max(1, h->slots / nstreams);
  const long long strideW = (long long)h->Np * h->Np;
  int chunk = 0;
  for (int b0 = 0; b0 < B; b0 += per, ++chunk) {
    const int nb = std::min(per, B - b0);
    const int si = chunk % nstreams;
    cudaStream_t st = h->streams[si];
    const int slot0 = si * per;
    double* Cb = C + (long long)b0 * h->N * h->N;
    SFB_CUDA(h, launch_copy_in_lower(Cb, h->N, h->W + slot0 * strideW, h->Np, strideW, nb, st));
    SFB_CUDA(h, launch_residual(nullptr, nullptr, h->N, h->Np, nb, h->rhs + (long long)slot0 * h->Np, nullptr,
                                h->logdet + slot0, h->sqmah + slot0, h->info_ws + slot0, st));
    h->launches += 2;
    if ((rc = run_cholesky(h, si, slot0, nb, nullptr, info + b0)) != SFB_OK) return rc;
    SFB_CUDA(h, launch_copy_out_lower(Cb, h->N, h->W + slot0 * strideW, h->Np, strideW, nb, st));
    h->launches++;
    if (logdet)
      SFB_CUDA(h, cudaMemcpyAsync(logdet + b0, h->logdet + slot0, sizeof(double) * nb, cudaMemcpyDeviceToDevice, st));
  }
  return join_streams(h, caller, nstreams);
}

int sfb_solve_lower(sfb_t* h, int B, const double* L, const double* r, double* z, void* stream) {
  NvtxRange nvtx_range("sfb_solve_lower");
  int rc = check_batch(h, B);
  if (rc != SFB_OK) return rc;
  if (!L || !r || !z) return fail(h, SFB_ERR_ARG, "sfb_solve_lower: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  cudaStream_t caller = (cudaStream_t)stream, st = h->streams[0];
  if ((rc = fork_streams(h, caller, 1)) != SFB_OK) return rc;
  SFB_CUDA(h, launch_solve_lower(L, (long long)h->N * h->N, h->N, r, z, h->N, B, st));
  h->launches++;
  return join_streams(h, caller, 1);
}

int sfb_loglike(sfb_t* h, int B, const double* X, const double* A, const double* model_flux, const double* glob,
                const int* nloc, const double* loc, int shared_hyper, double* lnL, int* info, double* resid,
                void* stream) {
  NvtxRange nvtx_range("sfb_loglike");
  int rc = check_batch(h, B);
  if (rc != SFB_OK) return rc;
  if (!h->have_static) return fail(h, SFB_ERR_STATE, "sfb_loglike: call sfb_set_static first");
  if (!model_flux || !glob || !nloc || !loc || !lnL || !info || (X && !A))
    return fail(h, SFB_ERR_ARG, "sfb_loglike: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h)) != SFB_OK) return rc;
  cudaStream_t caller = (cudaStream_t)stream;
  if (h->solver == SFB_SOLVER_STRUCTURED)
    return structured_device(h, B, X, A, model_flux, glob, nloc, loc, shared_hyper, lnL, info, resid, caller);
  if (use_shared_factor(h, B, shared_hyper))
    return shared_factor_device(h, B, X, A, model_flux, glob, nloc, loc, lnL, info, resid, caller);
  const int nstreams = (h->profile || h->slots < 2 || (h->debug_mode & 2)) ? 1 : std::min(h->nlanes, h->slots);
  if ((rc = fork_streams(h, caller, nstreams)) != SFB_OK) return rc;
  rc = loglike_device(h, B, X, A, model_flux, glob, nloc, loc, shared_hyper, lnL, info, resid, nstreams, nullptr,
                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != SFB_OK) return rc;
  return join_streams(h, caller, nstreams);
}

int sfb_loglike_host(sfb_t* h, int B, const double* X_h, const double* A_h, const double* model_flux_h,
                     const double* glob_h, const int* nloc_h, const double* loc_h, int shared_hyper,
                     double* lnL_h, int* info_h, double* resid_h) {
  NvtxRange nvtx_range("sfb_loglike_host");
  int rc = check_batch(h, B);
  if (rc != SFB_OK) return rc;
  if (!h->have_static) return fail(h, SFB_ERR_STATE, "sfb_loglike_host: call sfb_set_static first");
  if (!model_flux_h || !glob_h || !nloc_h || !loc_h || !lnL_h || !info_h || (X_h && !A_h))
    return fail(h, SFB_ERR_ARG, "sfb_loglike_host: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h)) != SFB_OK) return rc;
  const int N = h->N, M = h->M, K = h->Kmax, Bm = h->Bmax;
  auto lazy = [&](void** p, size_t bytes) {
    if (*p) return true;
    return cudaMalloc(p, std::max<size_t>(bytes, 16)) == cudaSuccess;
  };
  bool ok = true;
  ok &= lazy((void**)&h->dX, sizeof(double) * (size_t)Bm * std::max(M, 1) * N);
  ok &= lazy((void**)&h->dA, sizeof(double) * (size_t)Bm * std::max(M * M, 1));
  ok &= lazy((void**)&h->dflux, sizeof(double) * (size_t)Bm * N);
  ok &= lazy((void**)&h->dglob, sizeof(double) * 2 * Bm);
  ok &= lazy((void**)&h->dloc, sizeof(double) * 3 * (size_t)K * Bm);
  ok &= lazy((void**)&h->dnloc, sizeof(int) * Bm);
  ok &= lazy((void**)&h->dlnL, sizeof(double) * Bm);
  ok &= lazy((void**)&h->dinfo, sizeof(int) * Bm);
  if (resid_h) ok &= lazy((void**)&h->dresid, sizeof(double) * (size_t)Bm * N);
  if (!ok) return fail(h, SFB_ERR_NOMEM, "sfb_loglike_host: staging allocation failed");
  const bool shared_path = use_shared_factor(h, B, shared_hyper);
  if (h->solver == SFB_SOLVER_STRUCTURED || shared_path) {  // inputs up on lane 0, one device call, results down
    cudaStream_t st = h->streams[0];
    const int Bh = shared_hyper ? 1 : B;
    for (int i = 1; i < kMaxLanes; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
    if (X_h) {
      SFB_CUDA(h, cudaMemcpyAsync(h->dX, X_h, sizeof(double) * (size_t)B * M * N, cudaMemcpyHostToDevice, st));
      SFB_CUDA(h, cudaMemcpyAsync(h->dA, A_h, sizeof(double) * (size_t)B * M * M, cudaMemcpyHostToDevice, st));
    }
    SFB_CUDA(h, cudaMemcpyAsync(h->dflux, model_flux_h, sizeof(double) * (size_t)B * N, cudaMemcpyHostToDevice, st));
    SFB_CUDA(h, cudaMemcpyAsync(h->dglob, glob_h, sizeof(double) * 2 * Bh, cudaMemcpyHostToDevice, st));
    SFB_CUDA(h, cudaMemcpyAsync(h->dnloc, nloc_h, sizeof(int) * Bh, cudaMemcpyHostToDevice, st));
    SFB_CUDA(h, cudaMemcpyAsync(h->dloc, loc_h, sizeof(double) * 3 * (size_t)K * Bh, cudaMemcpyHostToDevice, st));
    rc = shared_path
             ? shared_factor_device(h, B, X_h ? h->dX : nullptr, X_h ? h->dA : nullptr, h->dflux, h->dglob, h->dnloc,
                                    h->dloc, h->dlnL, h->dinfo, resid_h ? h->dresid : nullptr, st)
             : structured_device(h, B, X_h ? h->dX : nullptr, X_h ? h->dA : nullptr, h->dflux, h->dglob, h->dnloc,
                                 h->dloc, shared_hyper, h->dlnL, h->dinfo, resid_h ? h->dresid : nullptr, st);
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, cudaMemcpyAsync(lnL_h, h->dlnL, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    SFB_CUDA(h, cudaMemcpyAsync(info_h, h->dinfo, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    if (resid_h)
      SFB_CUDA(h, cudaMemcpyAsync(resid_h, h->dresid, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, st));
    SFB_CUDA(h, cudaStreamSynchronize(st));
    return SFB_OK;
  }
  const int nstreams = (h->profile || h->slots < 2 || (h->debug_mode & 2)) ? 1 : std::min(h->nlanes, h->slots);
  rc = loglike_device(h, B, X_h ? h->dX : nullptr, X_h ? h->dA : nullptr, h->dflux, h->dglob, h->dnloc, h->dloc,
                      shared_hyper, h->dlnL, h->dinfo, resid_h ? h->dresid : nullptr, nstreams, X_h, A_h,
                      model_flux_h, glob_h, nloc_h, loc_h, lnL_h, info_h, resid_h);
  if (rc != SFB_OK) return rc;
  for (int i = 0; i < nstreams; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
  return SFB_OK;
}

// ---------------------------------------------------------------------------------------------------
// upstream of the covariance (rows f1/f2/f3)
// ---------------------------------------------------------------------------------------------------
int sfb_set_model_host(sfb_t* h, int nf, const double* fine_wave_h, const double* bulk_h, int G, int D,
                       const double* grid_points_h, const double* variances_h, const double* lengthscales_h,
                       const double* v11_h, const double* w_hat_h, int ncheb_max, int flags) {
  NvtxRange nvtx_range("sfb_set_model_host");
  if (!h) return SFB_ERR_ARG;
  if (!fine_wave_h || !bulk_h || !grid_points_h || !variances_h || !lengthscales_h || !v11_h || !w_hat_h)
    return fail(h, SFB_ERR_ARG, "sfb_set_model_host: NULL argument");
  if (nf < 16 || (nf & (nf - 1)) != 0) return fail(h, SFB_ERR_ARG, "sfb_set_model_host: nf must be a power of two >= 16");
  if (h->M < 1 || G < 1 || D < 1 || D > 16 || ncheb_max < 0 || ncheb_max > 32)
    return fail(h, SFB_ERR_ARG, "sfb_set_model_host: sizes out of range (needs M >= 1)");
  DeviceGuard guard(h->device);
  for (int i = 0; i < kMaxLanes; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
  h->have_model = false;
  std::string why;
  cudaError_t e = model_setup(&h->model, h->N, h->M, h->Bmax, nf, fine_wave_h, bulk_h, G, D, grid_points_h,
                              variances_h, lengthscales_h, v11_h, w_hat_h, ncheb_max, flags, &why);
  if (e != cudaSuccess) {
    model_free(&h->model);
    return fail(h, e == cudaErrorInvalidValue ? SFB_ERR_ARG : (e == cudaErrorMemoryAllocation ? SFB_ERR_NOMEM : SFB_ERR_CUDA),
                why.c_str(), e == cudaErrorInvalidValue ? cudaSuccess : e);
  }
  h->have_model = true;
  return SFB_OK;
}

extern "C++" {
namespace {
int check_upstream(sfb_ctx* h, int B, int ncheb, const char* who) {
  int rc = check_batch(h, B);
  if (rc != SFB_OK) return rc;
  if (!h->have_model) return fail(h, SFB_ERR_STATE, "call sfb_set_model_host first");
  if (!h->have_static) return fail(h, SFB_ERR_STATE, "call sfb_set_static first");
  if (ncheb < 0 || ncheb > h->model.ncheb_max) return fail(h, SFB_ERR_ARG, "ncheb exceeds ncheb_max of sfb_set_model_host");
  (void)who;
  return SFB_OK;
}

int run_upstream(sfb_ctx* h, int B, const double* theta, int ncheb, double* X, double* A, double* flux,
                 double* log_scale_out, int* status, cudaStream_t st) {
  UpstreamArgs a;
  a.B = B;
  a.ncheb = ncheb;
  a.ntheta = h->model.D + 4 + ncheb + ((h->model.flags & SFB_MODEL_AV) ? 1 : 0);
  a.theta = theta;
  a.wave = h->wave;
  a.data_flux = h->data_flux;
  a.X = X;
  a.A = A;
  a.flux = flux;
  a.log_scale_out = log_scale_out;
  a.status = status;
  const double bytes = (double)B * 8.0 * ((double)h->model.R * h->model.nf * 2.0 + (double)h->model.R * h->N * 2.0 +
                                          (double)(h->M + 1) * h->N);
  ProfScope ps(h, st, SFB_K_UPSTREAM, bytes);
  SFB_CUDA(h, launch_upstream(h->model, a, st, &h->launches));
  return SFB_OK;
}
}  // namespace
}  // extern "C++"

int sfb_upstream(sfb_t* h, int B, const double* theta, int ncheb, double* X, double* A, double* model_flux,
                 double* log_scale_out, int* status, double* weights, double* weights_cov, void* stream) {
  NvtxRange nvtx_range("sfb_upstream");
  int rc = check_upstream(h, B, ncheb, "sfb_upstream");
  if (rc != SFB_OK) return rc;
  if (!theta || !X || !A || !model_flux || !log_scale_out || !status)
    return fail(h, SFB_ERR_ARG, "sfb_upstream: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  cudaStream_t caller = (cudaStream_t)stream, st = h->streams[0];
  if ((rc = fork_streams(h, caller, 1)) != SFB_OK) return rc;
  if ((rc = run_upstream(h, B, theta, ncheb, X, A, model_flux, log_scale_out, status, st)) != SFB_OK) return rc;
  if (weights)
    SFB_CUDA(h, cudaMemcpyAsync(weights, h->model.mu, sizeof(double) * (size_t)B * h->M, cudaMemcpyDeviceToDevice, st));
  if (weights_cov)
    SFB_CUDA(h, cudaMemcpyAsync(weights_cov, h->model.wcov, sizeof(double) * (size_t)B * h->M * h->M,
                                cudaMemcpyDeviceToDevice, st));
  return join_streams(h, caller, 1);
}

int sfb_loglike_params(sfb_t* h, int B, const double* theta, int ncheb, const double* glob, const int* nloc,
                       const double* loc, int shared_hyper, double* lnL, int* info, double* resid,
                       double* log_scale_out, void* stream) {
  NvtxRange nvtx_range("sfb_loglike_params");
  int rc = check_upstream(h, B, ncheb, "sfb_loglike_params");
  if (rc != SFB_OK) return rc;
  if (!theta || !glob || !nloc || !loc || !lnL || !info) return fail(h, SFB_ERR_ARG, "sfb_loglike_params: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h)) != SFB_OK) return rc;
  cudaStream_t caller = (cudaStream_t)stream;
  const int nstreams = (h->profile || h->slots < 2 || (h->debug_mode & 2)) ? 1 : std::min(h->nlanes, h->slots);
  ModelState& ms = h->model;
  if (h->solver == SFB_SOLVER_STRUCTURED) {  // upstream on lane 1, overlapped with the band set-up on lane 0
    UpstreamJob job{theta, ncheb, log_scale_out ? log_scale_out : ms.log_scale};
    if ((rc = structured_device(h, B, ms.X, ms.A, ms.flux, glob, nloc, loc, shared_hyper, lnL, info, resid, caller,
                                &job)) != SFB_OK)
      return rc;
    SFB_CUDA(h, launch_merge_status(ms.status, info, lnL, B, caller));
    h->launches++;
    return SFB_OK;
  }
  // the upstream stage runs once for the whole batch on stream 0; the other lane waits for it
  if ((rc = order_after_previous_call(h)) != SFB_OK) return rc;
  SFB_CUDA(h, cudaEventRecord(h->ev_fork, caller));
  SFB_CUDA(h, cudaStreamWaitEvent(h->streams[0], h->ev_fork, 0));
  if ((rc = run_upstream(h, B, theta, ncheb, ms.X, ms.A, ms.flux, log_scale_out ? log_scale_out : ms.log_scale,
                         ms.status, h->streams[0])) != SFB_OK)
    return rc;
  SFB_CUDA(h, cudaEventRecord(h->ev_fork, h->streams[0]));
  for (int i = 1; i < nstreams; ++i) SFB_CUDA(h, cudaStreamWaitEvent(h->streams[i], h->ev_fork, 0));
  rc = loglike_device(h, B, ms.X, ms.A, ms.flux, glob, nloc, loc, shared_hyper, lnL, info, resid, nstreams, nullptr,
                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != SFB_OK) return rc;
  if ((rc = join_streams(h, caller, nstreams)) != SFB_OK) return rc;
  SFB_CUDA(h, launch_merge_status(ms.status, info, lnL, B, caller));
  h->launches++;
  return SFB_OK;
}

int sfb_loglike_params_host(sfb_t* h, int B, const double* theta_h, int ncheb, const double* glob_h,
                            const int* nloc_h, const double* loc_h, int shared_hyper, double* lnL_h, int* info_h,
                            double* resid_h, double* log_scale_h) {
  NvtxRange nvtx_range("sfb_loglike_params_host");
  int rc = check_upstream(h, B, ncheb, "sfb_loglike_params_host");
  if (rc != SFB_OK) return rc;
  if (!theta_h || !glob_h || !nloc_h || !loc_h || !lnL_h || !info_h)
    return fail(h, SFB_ERR_ARG, "sfb_loglike_params_host: NULL argument");
  if (B == 0) return SFB_OK;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h)) != SFB_OK) return rc;
  const int N = h->N, K = h->Kmax, Bm = h->Bmax;
  auto lazy = [&](void** p, size_t bytes) {
    if (*p) return true;
    return cudaMalloc(p, std::max<size_t>(bytes, 16)) == cudaSuccess;
  };
  bool ok = true;
  ok &= lazy((void**)&h->dglob, sizeof(double) * 2 * Bm);
  ok &= lazy((void**)&h->dloc, sizeof(double) * 3 * (size_t)K * Bm);
  ok &= lazy((void**)&h->dnloc, sizeof(int) * Bm);
  ok &= lazy((void**)&h->dlnL, sizeof(double) * Bm);
  ok &= lazy((void**)&h->dinfo, sizeof(int) * Bm);
  if (resid_h) ok &= lazy((void**)&h->dresid, sizeof(double) * (size_t)Bm * N);
  if (!ok) return fail(h, SFB_ERR_NOMEM, "sfb_loglike_params_host: staging allocation failed");
  ModelState& ms = h->model;
  cudaStream_t st = h->streams[0];
  const int Bh = shared_hyper ? 1 : B;
  const int ntheta = ms.D + 4 + ncheb + ((ms.flags & SFB_MODEL_AV) ? 1 : 0);
  for (int i = 1; i < kMaxLanes; ++i)
    SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));  // staging buffers may still be read by the other lanes
  SFB_CUDA(h, cudaMemcpyAsync(ms.theta, theta_h, sizeof(double) * (size_t)B * ntheta, cudaMemcpyHostToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->dglob, glob_h, sizeof(double) * 2 * Bh, cudaMemcpyHostToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->dnloc, nloc_h, sizeof(int) * Bh, cudaMemcpyHostToDevice, st));
  SFB_CUDA(h, cudaMemcpyAsync(h->dloc, loc_h, sizeof(double) * 3 * (size_t)K * Bh, cudaMemcpyHostToDevice, st));
  rc = sfb_loglike_params(h, B, ms.theta, ncheb, h->dglob, h->dnloc, h->dloc, shared_hyper, h->dlnL, h->dinfo,
                          resid_h ? h->dresid : nullptr, ms.log_scale, (void*)st);
  if (rc != SFB_OK) return rc;
  SFB_CUDA(h, cudaMemcpyAsync(lnL_h, h->dlnL, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  SFB_CUDA(h, cudaMemcpyAsync(info_h, h->dinfo, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
  if (log_scale_h)
    SFB_CUDA(h, cudaMemcpyAsync(log_scale_h, ms.log_scale, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  if (resid_h)
    SFB_CUDA(h, cudaMemcpyAsync(resid_h, h->dresid, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, st));
  SFB_CUDA(h, cudaStreamSynchronize(st));
  return SFB_OK;
}

int sfb_host_rfft(int n, const double* x_h, double* out_complex_h) {
  if (n < 2 || (n & (n - 1)) != 0 || !x_h || !out_complex_h) return SFB_ERR_ARG;
  host_rfft(n, x_h, out_complex_h);
  return SFB_OK;
}

int sfb_host_spline_inverse_band(int nf, const double* fine_wave_h, int W, double* out_h) {
  if (!fine_wave_h || !out_h || W < 1) return SFB_ERR_ARG;
  return host_spline_inverse_band(nf, fine_wave_h, W, out_h) == 0 ? SFB_OK : SFB_ERR_ARG;
}

int sfb_host_cholesky_lower(int n, double* a_h) {
  if (n < 1 || !a_h) return SFB_ERR_ARG;
  return host_cholesky_lower(n, a_h, n);
}

int sfb_spline_halfwidth(void) { return kSplineW; }

int sfb_set_solver(sfb_t* h, int solver) {
  if (!h) return SFB_ERR_ARG;
  if (solver != SFB_SOLVER_DENSE && solver != SFB_SOLVER_STRUCTURED && solver != SFB_SOLVER_DENSE_I8)
    return fail(h, SFB_ERR_ARG, "unknown solver");
  h->solver = solver;
  return SFB_OK;
}

int sfb_get_solver(const sfb_t* h) { return h ? h->solver : -1; }

int sfb_set_shared_factor(sfb_t* h, int on) {
  if (!h) return SFB_ERR_ARG;
  h->shared_factor = (on != 0);
  return SFB_OK;
}

long long sfb_shared_factor_calls(const sfb_t* h) { return h ? h->shared_factor_calls : -1; }

int sfb_i8_mma_counts(sfb_t* h, unsigned long long* issued, unsigned long long* dense) {
  if (!h || !issued || !dense) return SFB_ERR_ARG;
  *issued = *dense = 0;
  if (!h->oz_stats) return SFB_OK;
  DeviceGuard guard(h->device);
  for (int i = 0; i < kMaxLanes; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
  unsigned long long v[8];
  SFB_CUDA(h, cudaMemcpy(v, h->oz_stats, sizeof(v), cudaMemcpyDeviceToHost));
  SFB_CUDA(h, cudaMemset(h->oz_stats, 0, sizeof(v)));
  *issued = v[0];
  *dense = v[1];
#ifdef SFB_EXPERIMENTS  // where the MMA warp of the int8 update spends its time (clock64 sums over all CTAs)
  if (getenv("SFB_OZ_TIMING") && v[5])
    fprintf(stderr, "[sfb] int8 MMA warp: %.0f cycles per tile = %.1f %% waiting for operands + %.1f %% waiting for the epilogue "
            "to drain the accumulators + %.1f %% issuing; %llu tiles, %.1f slab products per chunk\n",
            (double)v[2] / v[5], 100.0 * v[3] / v[2], 100.0 * v[4] / v[2], 100.0 * (v[2] - v[3] - v[4]) / v[2], v[5],
            26.0 * v[0] / v[1]);
  if (getenv("SFB_OZ_TIMING") && v[7])
    fprintf(stderr, "[sfb] int8 producer warp: %.1f %% of its time waiting for a free stage\n", 100.0 * v[6] / v[7]);
#endif
  return SFB_OK;
}

int sfb_band_classes(const sfb_t* h, int* widths, long long* walkers, int n) {
  if (!h || !widths || !walkers || n < kNumBandWidths + 1) return SFB_ERR_ARG;
  for (int c = 0; c < kNumBandWidths; ++c) { widths[c] = kBandWidths[c]; walkers[c] = h->band_rows[c]; }
  widths[kNumBandWidths] = 0;  // dense fallback
  walkers[kNumBandWidths] = h->band_rows[kNumBandWidths];
  return kNumBandWidths + 1;
}

// ---- NCCL, bound lazily ---------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char b[128]; };  // layout of ncclUniqueId (nccl.h: struct { char internal[128]; })
struct NcclApi {
  // the few entry points of nccl.h this path needs (ncclResult_t / ncclDataType_t are ints)
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);   // the copy already in the process (torch's) if any
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
  api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
  return api.ok ? &api : nullptr;
}
int nccl_fail(sfb_ctx* h, NcclApi* n, const char* what, int code) {
  h->err = std::string(what) + ": " + n->GetErrorString(code);
  return SFB_ERR_CUDA;
}
}  // namespace

int sfb_comm_unique_id(void* id_h) {
  NcclApi* n = nccl_api();
  if (!n || !id_h) return SFB_ERR_STATE;
  return n->GetUniqueId(id_h) == 0 ? SFB_OK : SFB_ERR_CUDA;
}

int sfb_comm_init(sfb_t* h, int rank, int nranks, const void* unique_id_h) {
  NvtxRange nvtx_range("sfb_comm_init");
  if (!h || !unique_id_h || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, SFB_ERR_ARG, "sfb_comm_init: bad argument");
  NcclApi* n = nccl_api();
  if (!n) return fail(h, SFB_ERR_STATE, "sfb_comm_init: libnccl.so.2 not found");
  DeviceGuard guard(h->device);
  if (h->nccl_comm) { n->CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
  NcclUniqueId id;
  memcpy(id.b, unique_id_h, sizeof(id.b));
  const int rc = n->CommInitRank(&h->nccl_comm, nranks, id, rank);
  if (rc != 0) return nccl_fail(h, n, "ncclCommInitRank", rc);
  h->comm_rank = rank;
  h->comm_size = nranks;
  return SFB_OK;
}

int sfb_comm_destroy(sfb_t* h) {
  if (!h) return SFB_ERR_ARG;
  if (h->nccl_comm) {
    NcclApi* n = nccl_api();
    DeviceGuard guard(h->device);
    if (n) n->CommDestroy(h->nccl_comm);
    h->nccl_comm = nullptr;
    h->comm_size = 1;
    h->comm_rank = 0;
  }
  return SFB_OK;
}

int sfb_allgather_lnL(sfb_t* h, const double* lnL_local, int count, double* lnL_all, void* stream) {
  NvtxRange nvtx_range("sfb_allgather_lnL");
  if (!h || !lnL_local || !lnL_all || count < 0) return fail(h, SFB_ERR_ARG, "sfb_allgather_lnL: bad argument");
  if (!h->nccl_comm) return fail(h, SFB_ERR_STATE, "sfb_allgather_lnL: call sfb_comm_init first");
  if (count == 0) return SFB_OK;
  NcclApi* n = nccl_api();
  DeviceGuard guard(h->device);
  const int rc = n->AllGather(lnL_local, lnL_all, (size_t)count, /* ncclFloat64 */ 8, h->nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(h, n, "ncclAllGather", rc);
  return SFB_OK;
}

int sfb_sync(sfb_t* h) {
  if (!h) return SFB_ERR_ARG;
  DeviceGuard guard(h->device);
  for (int i = 0; i < kMaxLanes; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
  return SFB_OK;
}

int sfb_profile_enable(sfb_t* h, int on) {
  if (!h) return SFB_ERR_ARG;
  h->profile = (on != 0);
  return SFB_OK;
}

int sfb_profile_read(sfb_t* h, double* out, int n) {
  if (!h || !out || n < 3 * SFB_K_NCLASS) return fail(h, SFB_ERR_ARG, "sfb_profile_read: buffer too small");
  DeviceGuard guard(h->device);
  for (int i = 0; i < kMaxLanes; ++i) SFB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
  for (int i = 0; i < 3 * SFB_K_NCLASS; ++i) out[i] = 0.0;
  for (auto& pe : h->prof) {
    float ms = 0.f;
    SFB_CUDA(h, cudaEventElapsedTime(&ms, pe.a, pe.b));
    out[3 * pe.cls + 0] += 1.0;
    out[3 * pe.cls + 1] += ms;
    out[3 * pe.cls + 2] += pe.work;
    h->ev_pool.push_back(pe.a);
    h->ev_pool.push_back(pe.b);
  }
  h->prof.clear();
  return SFB_OK;
}

int sfb_workspace_walkers(const sfb_t* h) { return h ? h->slots : 0; }
int sfb_padded_n(const sfb_t* h) { return h ? h->Np : 0; }
long long sfb_launch_count(const sfb_t* h) { return h ? h->launches : 0; }
const char* sfb_last_error(const sfb_t* h) { return h ? h->err.c_str() : "null handle"; }

}  // extern "C"
