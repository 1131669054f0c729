/*
 * sfb200.h — C ABI of libsfb200.so: the B200-native (sm_100a) implementation of Starfish's per-step
 * log-likelihood hot path.
 *
 * The reference (Starfish v0.4.2, pure Python) has no FFI for this path: the seam is a handful of Python
 * call sites.  Each entry point below names the reference lines it replaces (paths relative to the
 * reference root):
 *
 *   sfb_build_cov      Starfish/models/kernels.py:7-41   global_covariance_matrix (Matérn-3/2 × Hann)
 *                      Starfish/models/kernels.py:44-81  local_covariance_matrix  (Gaussian × Hann)
 *                      Starfish/models/spectrum_model.py:334-363  XᵀAX + diag(σ²) + global + Σ local
 *   sfb_potrf          Starfish/models/spectrum_model.py:400  scipy.linalg.cho_factor  (LAPACK dpotrf)
 *   sfb_loglike(_host) Starfish/models/spectrum_model.py:334-363 + :399-405
 *                      (assembly, +1e-10·I, cho_factor, logdet, residual, cho_solve, quadratic form)
 *   sfb_solve_lower    Starfish/models/spectrum_model.py:404  scipy.linalg.cho_solve (forward half;
 *                      sqmah = ‖L⁻¹R‖²)
 *
 * Conventions
 *   - All arithmetic is IEEE fp64.  Matrices are row-major ("C order", what numpy hands over).
 *   - Pointers are DEVICE pointers unless the parameter name ends in _h (host; pinned recommended).
 *   - `stream` is the caller's CUDA stream (cudaStream_t cast to void*; NULL = legacy default stream).
 *     Work is ordered after everything already queued on `stream`, and `stream` waits for the result,
 *     so the caller may enqueue consumers (or record timing events) on `stream` right after the call.
 *     Calls return without waiting for the GPU (except the *_host variants' final copy, see below).
 *   - Every function returns 0 on success, a negative sfb_status otherwise; never throws.
 *     sfb_last_error() gives the text.  A handle is not thread-safe: serialise calls on it.
 *   - Per-walker `info` follows LAPACK dpotrf: 0 = ok, i > 0 = leading minor of order i is not positive
 *     definite (the reference raises numpy.linalg.LinAlgError there); lnL is NaN for such rows.
 */
#ifndef SFB200_H
#define SFB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfb_ctx sfb_t;

enum sfb_status {
  SFB_OK = 0,
  SFB_ERR_ARG = -1,     /* bad argument (sizes, NULL pointers)          */
  SFB_ERR_CUDA = -2,    /* a CUDA runtime call failed                   */
  SFB_ERR_NOMEM = -3,   /* workspace does not fit in device memory      */
  SFB_ERR_STATE = -4    /* call order (e.g. loglike before set_static)  */
};

/* ABI version of this header; bumped on any signature change. */
int sfb_abi_version(void);

/*
 * Create a handle on `device` for spectra of N pixels, up to M eigenspectra (rank of the emulator term),
 * up to Kmax local kernels per walker and batches of up to Bmax walkers per call.
 * `workspace_walkers` = how many N×N fp64 factorisation slots to allocate (0 = choose automatically:
 * enough to keep the GPU full, bounded by free memory; < 0 = none yet — a build-only handle for sfb_build_cov,
 * the workspace is then allocated, automatically sized, by the first call that factorises).
 */
int sfb_create(int device, int N, int M, int Kmax, int Bmax, int workspace_walkers, sfb_t** out);
int sfb_destroy(sfb_t* h);

/* Data shared by all walkers: wavelengths, per-pixel noise σ (NOT σ²) and observed flux (each N). */
int sfb_set_static(sfb_t* h, const double* wave, const double* sigma, const double* data_flux, void* stream);
int sfb_set_static_host(sfb_t* h, const double* wave_h, const double* sigma_h, const double* data_flux_h);

/*
 * Covariance only:  C[b] = X[b]ᵀ·A[b]·X[b] + diag(σ² + jitter) + K_global(glob[b]) + Σ_k K_local(loc[b,k])
 *   X     B×M×N   (NULL ⇒ no emulator term; then A is ignored)
 *   A     B×M×M   symmetric (parity mode: A = Σ_w⁻¹)
 *   glob  B×2     (amplitude, lengthscale), already exponentiated; amplitude <= 0 ⇒ no global kernel
 *   nloc  B       number of local kernels of walker b (<= Kmax);  loc  B×Kmax×3 (amplitude, mu, sigma)
 *   shared_hyper != 0 ⇒ glob/nloc/loc hold ONE row used by every walker (frozen-kernel sharing)
 *   C     B×N×N   row-major, both triangles written
 */
int sfb_build_cov(sfb_t* h, int B, const double* X, const double* A, const double* glob, const int* nloc,
                  const double* loc, int shared_hyper, double jitter, double* C, void* stream);

/*
 * Batched in-place Cholesky of B row-major N×N SPD matrices (lower triangle read; on return the lower
 * triangle holds L with C = L·Lᵀ, the strict upper triangle is left untouched).  logdet may be NULL;
 * otherwise logdet[b] = 2·Σ log L_ii.
 */
int sfb_potrf(sfb_t* h, int B, double* C, int* info, double* logdet, void* stream);

/* z[b] = L[b]⁻¹·r[b] for B lower-triangular row-major factors (forward substitution). */
int sfb_solve_lower(sfb_t* h, int B, const double* L, const double* r, double* z, void* stream);

/*
 * The whole stage boundary, B walkers:
 *   lnL[b] = −½·( log det(C_b + 1e-10·I) + R_bᵀ (C_b + 1e-10·I)⁻¹ R_b ),  R_b = model_flux[b] − data_flux
 * Inputs as sfb_build_cov plus model_flux (B×N).  Outputs lnL (B), info (B), resid (B×N or NULL).
 * No priors (they stay on the host, spectrum_model.py:387-395).
 */
int sfb_loglike(sfb_t* h, int B, const double* X, const double* A, const double* model_flux,
                const double* glob, const int* nloc, const double* loc, int shared_hyper,
                double* lnL, int* info, double* resid, void* stream);

/*
 * Same, with HOST buffers (the end-to-end path): inputs are copied host→device and lnL/info (and resid
 * if not NULL) device→host inside the call, chunk by chunk, overlapped with the factorisation of the
 * previous chunk.  Returns after the results are in the host buffers.
 */
int sfb_loglike_host(sfb_t* h, int B, const double* X_h, const double* A_h, const double* model_flux_h,
                     const double* glob_h, const int* nloc_h, const double* loc_h, int shared_hyper,
                     double* lnL_h, int* info_h, double* resid_h);

/*
 * ---- Upstream of the covariance (scope-table rows f1/f2/f3): parameters in, log-likelihood out ----
 *
 * sfb_set_model_host   what SpectrumModel.__init__ prepares          Starfish/models/spectrum_model.py:149-156
 *                      + the static part of Emulator.__call__        Starfish/emulator/emulator.py:382-388
 * sfb_upstream         SpectrumModel.__call__ up to the rank-M term  Starfish/models/spectrum_model.py:287-332:
 *                        rotational_broaden   Starfish/transforms.py:93-134
 *                        doppler_shift        Starfish/transforms.py:137-158
 *                        resample (k=5)       Starfish/transforms.py:11-42
 *                        extinct (ccm89)      Starfish/transforms.py:161-206
 *                        chebyshev_correct    Starfish/transforms.py:271-304
 *                        Emulator.__call__    Starfish/emulator/emulator.py:330-394 (weights, Σ_w)
 *                        X = eig·std, flux = w·X + mean, rescale / renorm   spectrum_model.py:306-332
 *                        A = Σ_w⁻¹            spectrum_model.py:334-335
 * sfb_loglike_params(_host)  sfb_upstream followed by sfb_loglike — one call per ensemble step; with the
 *                      _host variant only B×ntheta parameter values cross the PCIe bus.
 *
 * Model flags say which optional parameters the model has (the reference tests `"vsini" in self.params` ...).
 */
enum sfb_model_flags {
  SFB_MODEL_VSINI = 1,      /* rotational broadening (theta column D)                                  */
  SFB_MODEL_VZ = 2,         /* Doppler shift (theta column D+1)                                        */
  SFB_MODEL_LOG_SCALE = 4,  /* log_scale given (column D+2); otherwise renormalise to the data flux    */
  SFB_MODEL_NORM = 8,       /* multiply by the emulator's norm factor (column D+3, host-interpolated)   */
  SFB_MODEL_PAPER_TERM = 16,/* A = Σ_w (paper) instead of Σ_w⁻¹ (as the reference codes it)            */
  SFB_MODEL_AV = 32         /* interstellar extinction (LAST theta column): Starfish/transforms.py:161-206 with
                               the defaults SpectrumModel uses (law ccm89, R_V = 3.1), spectrum_model.py:298-299   */
};

/*
 * Static model data (host pointers; copied).  M (eigenspectra) and N come from sfb_create.
 *   nf            power-of-two length of the model's internal log-λ grid (create_log_lam_grid)
 *   fine_wave_h   nf         that grid (SpectrumModel.min_dv_wave)
 *   bulk_h        (M+2)×nf   eigenspectra, flux_mean, flux_std on it (SpectrumModel.bulk_fluxes)
 *   G, D          emulator grid points and their dimension;  grid_points_h G×D
 *   variances_h   M;  lengthscales_h M×D;  v11_h (M·G)×(M·G) row-major;  w_hat_h M·G
 *   ncheb_max     most Chebyshev coefficients (c1..) a call will pass
 */
int sfb_set_model_host(sfb_t* h, int nf, const double* fine_wave_h, const double* bulk_h, int G, int D,
                       const double* grid_points_h, const double* variances_h, const double* lengthscales_h,
                       const double* v11_h, const double* w_hat_h, int ncheb_max, int flags);

/*
 * theta: B×ntheta (ntheta = D+4+ncheb, +1 with SFB_MODEL_AV), row b = [grid params (D) | vsini | vz | log_scale | norm |
 * c1..c_ncheb | Av];
 * columns the model's flags do not use are ignored.  Outputs (device): X B×M×N, A B×M×M, model_flux B×N,
 * log_scale_out B (the fitted value when SFB_MODEL_LOG_SCALE is off), status B (0 ok, 1 = Σ_w not
 * positive definite — the reference raises LinAlgError there).  weights/weights_cov (B×M, B×M×M) may be NULL.
 */
int sfb_upstream(sfb_t* h, int B, const double* theta, int ncheb, double* X, double* A, double* model_flux,
                 double* log_scale_out, int* status, double* weights, double* weights_cov, void* stream);

/* sfb_upstream + sfb_loglike on device buffers.  info[b] = -1 where status[b] != 0. */
int sfb_loglike_params(sfb_t* h, int B, const double* theta, int ncheb, const double* glob, const int* nloc,
                       const double* loc, int shared_hyper, double* lnL, int* info, double* resid,
                       double* log_scale_out, void* stream);

/* The end-to-end entry: HOST buffers in and out (theta, hyper-parameters; lnL, info, log_scale, resid). */
int sfb_loglike_params_host(sfb_t* h, int B, const double* theta_h, int ncheb, const double* glob_h,
                            const int* nloc_h, const double* loc_h, int shared_hyper, double* lnL_h, int* info_h,
                            double* resid_h, double* log_scale_h);

/* Pure host helpers used at set-up, exported so that they can be checked without a GPU. */
int sfb_host_rfft(int n, const double* x_h, double* out_complex_h /* 2·(n/2+1) */);
int sfb_host_spline_inverse_band(int nf, const double* fine_wave_h, int W, double* out_h /* (2W+1)×nf */);
int sfb_host_cholesky_lower(int n, double* a_h /* n×n row-major, in place */);
int sfb_spline_halfwidth(void);

/*
 * ---- Structure-exploiting solver (scope-table row f4) ----
 * Same contract and results as the dense path (the lines of spectrum_model.py:334-363 + :399-405), computed
 * from C = S + XᵀAX with S = diag + K_global + ΣK_local banded on a strictly increasing wavelength grid:
 * banded Cholesky of S (N·b² FLOP instead of N³/3) + the M×M capacitance system (determinant lemma,
 * Woodbury).  SFB_SOLVER_STRUCTURED makes sfb_loglike / sfb_loglike_host / sfb_loglike_params(_host) use it;
 * walkers whose band does not fit the widest register window (160 pixels), or an unsorted grid, silently
 * take the dense path inside the same call, so results never depend on the choice beyond rounding
 * (|ΔlnL| <= 1e-10·|lnL|, tests/test_gpu_structured.py).  The default is SFB_SOLVER_DENSE.
 * In structured mode the call synchronises with the host once (B ints: each walker's half-bandwidth).
 * info[b] in structured mode: 0 = ok; i in [1, N] = pivot i of the banded factor of S is not positive (S itself is
 * not positive definite — the dense path reports the same leading minor); N = S is fine but S + XᵀAX is not
 * positive definite (the dense path reports whichever leading minor fails first — both mean LinAlgError upstream).
 */
enum sfb_solver { SFB_SOLVER_DENSE = 0, SFB_SOLVER_STRUCTURED = 1, SFB_SOLVER_DENSE_I8 = 2 };
int sfb_set_solver(sfb_t* h, int solver);
int sfb_get_solver(const sfb_t* h);
/*
 * Shared-factor path — the reference's frozen-kernel cache (Starfish/models/spectrum_model.py:341-363: with
 * "global_cov" and "local_cov" frozen the cached kernel matrices are reused, its production MCMC mode,
 * examples/single.ipynb:436).  With shared_hyper != 0 every walker has the same S = diag(σ²+1e-10) + K_global +
 * ΣK_local; the dense solvers then build and factorise S ONCE per call and solve L⁻¹[R_b | X_bᵀ] for all walkers
 * together (N³/3 + B·N²(M+1) FLOP instead of B·N³/3), finishing with the M×M capacitance system per walker.
 * Same lnL to rounding (|ΔlnL| <= 1e-10·|lnL|, tests/test_gpu_shared_factor.py).  On by default for B >= 2;
 * sfb_set_shared_factor(h, 0) makes shared_hyper calls factorise every walker's full covariance again.
 */
/*
 * SFB_SOLVER_DENSE_I8 skips every int8 product one of whose digit slabs is identically zero (exact: a product with
 * zero).  Counts since the last call: MMAs issued, and MMAs a dense digit pattern would have needed (26 per tile and
 * 32-deep chunk).  Resets the counters; synchronises the handle.
 */
int sfb_i8_mma_counts(sfb_t* h, unsigned long long* issued, unsigned long long* dense);

int sfb_set_shared_factor(sfb_t* h, int on);
long long sfb_shared_factor_calls(const sfb_t* h);

/* Walkers routed to each register-window width since creation; the last entry (width 0) is the dense
 * fallback.  Returns the number of entries written (n must be >= 5). */
int sfb_band_classes(const sfb_t* h, int* widths, long long* walkers, int n);

/*
 * ---- Multi-GPU: walkers are independent, so every rank (one handle, one GPU, one process or
 * thread) evaluates its own contiguous shard and the ONLY exchange per ensemble step is one all-gather of the
 * log-likelihood scalars — the reference evaluates the walkers serially (examples/single.ipynb:458-470) and has no
 * counterpart.  NCCL is bound at run time (dlopen of libnccl.so.2), so a host that never calls sfb_comm_init does
 * not need it.
 *   sfb_comm_unique_id   rank 0 creates the 128-byte NCCL unique id; the host distributes it (MPI, files, sockets …)
 *   sfb_comm_init        every rank, collectively: ncclCommInitRank on the handle's device
 *   sfb_allgather_lnL    lnL_all[r·count + i] = lnL_local[i] of rank r (ncclAllGather of `count` doubles per rank,
 *                        equal on every rank), enqueued on the caller's `stream` — after sfb_loglike on the same
 *                        stream it follows the last factorisation epilogue without a host round trip
 */
int sfb_comm_unique_id(void* unique_id_h /* 128 bytes out */);
int sfb_comm_init(sfb_t* h, int rank, int nranks, const void* unique_id_h);
int sfb_allgather_lnL(sfb_t* h, const double* lnL_local, int count, double* lnL_all, void* stream);
int sfb_comm_destroy(sfb_t* h);

/* Block the host until all work queued on the handle has finished. */
int sfb_sync(sfb_t* h);

/*
 * Kernel-level timing for roofline reports.  With profiling on, the handle runs its kernels on ONE
 * stream and brackets every launch with CUDA events.  sfb_profile_read fills, per kernel class,
 * out[3*c+0] = number of launches, out[3*c+1] = total device milliseconds, out[3*c+2] = algorithmic work
 * (bytes for SFB_K_BUILD, FLOPs otherwise), and resets the counters.  `n` = capacity of out in doubles.
 * SFB_K_OZ_SLICE stays in the enumeration for layout compatibility; it no longer receives launches (the fixed-point
 * slices of the int8 solver are written by the SFB_K_TRSM kernel).
 */
enum sfb_kernel_class { SFB_K_BUILD = 0, SFB_K_POTRF_DIAG = 1, SFB_K_TRSM = 2, SFB_K_SYRK = 3, SFB_K_UPSTREAM = 4,
                        SFB_K_BAND_BUILD = 5, SFB_K_BAND_CHOL = 6, SFB_K_OZ_SLICE = 7, SFB_K_FWD_ROWS = 8, SFB_K_NCLASS = 9 };
int sfb_profile_enable(sfb_t* h, int on);
int sfb_profile_read(sfb_t* h, double* out, int n);

/* Introspection: number of workspace slots, padded leading dimension, kernel launches since creation. */
int sfb_workspace_walkers(const sfb_t* h);
int sfb_padded_n(const sfb_t* h);
long long sfb_launch_count(const sfb_t* h);

const char* sfb_last_error(const sfb_t* h);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H */
