/*
 * artgpn_b200 -- C ABI of the B200-native likelihood / prediction hot path of artgpn.
 *
 * The reference (jdavidrcamacho/artgpn) has no FFI layer: its boundary is the Python class
 * artgpn.arttwo.network plus the node / weight / mean classes.  Every entry point below names
 * the reference code it replaces (file:line relative to the reference tree).  The Python
 * package artgpn_b200 binds these with ctypes (artgpn_b200/_lib.py); INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types (streams travel as void*).
 *   - every function returns 0 on success, non-zero on CUDA / usage errors; the message is
 *     available from agpn_last_error().  Numerical failure of a walker is NOT an error return:
 *     it is reported through out[w] / status[w] (see agpn_loglike_batch).
 *   - all floating point data is IEEE binary64, row-major.
 *   - one context per (process, device); calls on one context are serialised by the caller.
 *
 * Kernel kind ids (node.py / weight.py class -> id; Laplacian shares Exponential's formula,
 * weight.py:573-574 vs :618-619):
 *   0 Constant  1 WhiteNoise  2 SquaredExponential  3 Periodic  4 QuasiPeriodic
 *   5 RationalQuadratic  6 RQP  7 Cosine  8 Exponential|Laplacian  9 Matern32  10 Matern52
 *   11 Linear  12 GammaExp  13 Polynomial
 * Node kernels take their .pars as in node.py (1..4 doubles); weight kernels take the same
 * list with the amplitude `weight` first (weight.py), except Constant / WhiteNoise (1 double).
 * Mean kind ids (mean.py): 0 Constant 1 Linear 2 Parabola 3 Cubic 4 Sine 5 Cosine 6 Keplerian;
 * a mean.Sum is a list of terms.
 *
 * theta row layout (D = agpn_theta_dim doubles), SURVEY.md section 8(b):
 *   node_0.pars .. node_{q-1}.pars | weight_0.pars .. weight_{pq-1}.pars (list index j + q*i,
 *   arttwo.py:274) | mean terms of series 0 .. p-1 in order | jitter_0 .. jitter_{p-1}
 */
#ifndef ARTGPN_B200_H
#define ARTGPN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agpn_ctx agpn_ctx;

#define AGPN_FAMILY_NODE 0
#define AGPN_FAMILY_WEIGHT 1

/* per-walker status values written by agpn_loglike_batch */
#define AGPN_STATUS_OK 0
#define AGPN_STATUS_NOT_PD 1      /* scipy LinAlgError -> reference returns -inf (arttwo.py:288-289) */
#define AGPN_STATUS_NOT_FINITE 2  /* scipy check_finite ValueError (arttwo.py:284) */

int agpn_version(void);

/* Last error message of this thread's most recent failing call ("" if none). */
const char* agpn_last_error(void);

/* Number of .pars doubles of a kernel / mean kind; -1 if the id is unknown. */
int agpn_kernel_npars(int family, int kind);
int agpn_mean_npars(int kind);

/*
 * Replaces network.__init__ (arttwo.py:27-51): packs time[N], y[p][N] and the RAW sigma
 * yerr[p][N] (not squared, arttwo.py:40-48) and copies them to `device` once.
 */
int agpn_create(agpn_ctx** out, int device, int N, int p, int q,
                const double* time, const double* y, const double* yerr);
int agpn_destroy(agpn_ctx* ctx);

/*
 * Fixes which kernel / mean classes the parameter rows refer to (the reference re-reads
 * type(k) and k.pars on every call, arttwo.py:273-277).
 *   node_kind[q]; weight_kind[p*q] in list order j + q*i; mean_nterms[p] (0 == None,
 *   arttwo.py:112-113); mean_kind[sum(mean_nterms)].
 *   use_means == 0 reproduces `if means` being falsy (arttwo.py:263): nothing is subtracted
 *   and the theta row carries no mean parameters.
 */
int agpn_set_structure(agpn_ctx* ctx, const int* node_kind, const int* weight_kind,
                       const int* mean_nterms, const int* mean_kind, int use_means);
int agpn_theta_dim(const agpn_ctx* ctx);

/*
 * Accuracy / speed switch for every covariance entry the context produces afterwards.  enable = 0
 * (default): divisions hoisted per walker, sin(pi |r| / P) from per-tile angle-addition tables, fused
 * multiply-adds -- entries agree with the reference to ~1e-13 relative.  enable = 1: the reference's own
 * operation order with one rounding per numpy operation (node.py / weight.py formulas as written,
 * arttwo.py:279-281 accumulation) -- entries agree to the last ulp or two (libm differences only), ~4 x
 * slower assembly.  Matters only for ill-conditioned covariances: a relative difference d in K shows up
 * as ~ cond(K) * d in the log-likelihood.
 */
int agpn_set_exact(agpn_ctx* ctx, int enable);

/*
 * Replaces network.log_likelihood (arttwo.py:240-290) for W parameter rows at once.
 *   theta[W][D], out[W]; *_on_device says whether the pointer is device memory on ctx's device
 *   (else host memory: the copy happens inside the call on `stream`).
 *   status (host, may be NULL) receives AGPN_STATUS_* per walker.  out[w] is -inf for
 *   NOT_PD and NaN for NOT_FINITE.  The call returns after the results are in `out`
 *   (host out) or are stream-ordered on `stream` (device out, status == NULL).
 */
int agpn_loglike_batch(agpn_ctx* ctx, int W, const double* theta, int theta_on_device,
                       double* out, int out_on_device, int* status, void* stream);

/*
 * Walker sharding over the GPUs of one box (one process per GPU).  The reference's only parallelism is a process
 * pool over emcee walkers (examples/Sun/01_emcee.py:119-122); here the W rows of a batch are block-partitioned over
 * the ranks of an NCCL communicator (rows [lo, hi) of rank r: contiguous blocks, the first W % nranks ranks get one
 * more), every rank evaluates its rows on its own GPU with the result written straight into its slot of the gather
 * buffer, and ONE ncclAllGather on `stream` returns all W log-likelihoods on every rank.  No matrix is split across
 * GPUs: every value is bit-identical to a single-GPU agpn_loglike_batch.
 *   agpn_comm_unique_id: fills 128 bytes (an ncclUniqueId) -- call on one rank, distribute by any means.
 *   agpn_comm_init: collective over all ranks; binds the communicator to ctx (released by agpn_destroy).
 *   agpn_loglike_batch_sharded: collective; theta[W][D] is the FULL batch (same on every rank), out[W] / status[W]
 *   receive the full result on every rank.  NCCL is resolved at run time (dlopen of libnccl.so.2; an already loaded
 *   copy, e.g. torch's, is reused); the library has no link-time NCCL dependency.
 */
int agpn_comm_unique_id(void* id128_out);
int agpn_comm_init(agpn_ctx* ctx, const void* nccl_unique_id128, int rank, int nranks);
int agpn_comm_rank(const agpn_ctx* ctx, int* rank, int* nranks);
int agpn_loglike_batch_sharded(agpn_ctx* ctx, int W, const double* theta, int theta_on_device,
                               double* out, int out_on_device, int* status, void* stream);

/*
 * Replaces network._covariance_matrix (arttwo.py:151-192) and the K* loop of prediction
 * (arttwo.py:373-384): out[na][nb] = sum_j W_{series,j}(ta,tb) * F_j(ta,tb) for one theta row.
 *   ta == NULL / tb == NULL mean the training grid.  `square_quirk` != 0 makes WhiteNoise put
 *   wn^2 where row index == column index (node.py:68-72 fires whenever the lag array is square).
 *   add_errors != 0 adds yerr[series]^2 on the diagonal (needs na == nb == N, arttwo.py:190-191).
 *   series is 0-based.  theta, ta, tb: host.  out: host or device.
 */
int agpn_covariance_block(agpn_ctx* ctx, const double* theta, int series,
                          const double* ta, int na, const double* tb, int nb,
                          int square_quirk, int add_errors,
                          double* out, int out_on_device, void* stream);

/*
 * Replaces network._mean (arttwo.py:106-126): out[p][n] on the grid t (NULL = training grid).
 * Rows of series without a mean stay 0.  theta: full host row.  out: host.
 */
int agpn_mean_table(agpn_ctx* ctx, const double* theta, const double* t, int n, double* out);

/*
 * The same mean functions (mean.py:68-202, network._mean arttwo.py:106-126) without a context: the
 * classes and parameters travel with the call -- mean_nterms[p] terms per series (0 == None; a mean.Sum
 * is several terms), mean_kind[sum] and pars[] = the terms' .pars back to back -- so no context's
 * structure is touched.  Linear is centred on mean(t), Keplerian anchored at t[0] (mean.py:92,175).
 * out[p][n]; host pointers; runs on `device`.
 */
int agpn_mean_eval(int device, int p, const int* mean_nterms, const int* mean_kind, const double* pars,
                   const double* t, int n, double* out);

/*
 * Replaces network.prediction (arttwo.py:329-396) for one theta row and one series (0-based):
 * mean_out[M] = K* K^-1 r + m(t*), std_out[M] = sqrt(diag(K** + s^2 I - K* K^-1 K*^T)).
 * All pointers host.  *status receives AGPN_STATUS_* of the factorisation (the reference lets
 * LinAlgError escape here, arttwo.py:371).
 */
int agpn_predict(agpn_ctx* ctx, const double* theta, int series, int M, const double* tstar,
                 double* mean_out, double* std_out, int* status, void* stream);

/*
 * agpn_predict for a SLICE of a larger prediction grid (multi-GPU: the grid is split over ranks, SURVEY
 * section 8(e)): tstar[M] = whole_grid[grid_offset : grid_offset + M] of a grid of grid_M points.
 * mean.Linear is centred on the mean of the WHOLE grid and mean.Keplerian anchored at its first point
 * (mean.py:92,175), so both are passed in; WhiteNoise's square-lag quirk inside K* (node.py:68-72) is
 * decided by the whole grid too: wn^2 iff grid_M == N and (grid_offset + row) == training index, whatever
 * the slice length -- results do not depend on how the grid is split.  Not valid for a Keplerian mean whose
 * "any |M1| > 1e-10" test (mean.py:180) would differ between the slice and the whole grid.
 */
int agpn_predict_slice(agpn_ctx* ctx, const double* theta, int series, int M, const double* tstar,
                       int grid_M, int grid_offset, double grid_mean, double grid_t0,
                       double* mean_out, double* std_out, int* status, void* stream);

/*
 * Predictive mean AND full covariance (what y_cov of arttwo.py:393 / the legacy art.py:350-354 hold):
 * cov_out[M][M] = K** + s^2 I - K* K^-1 K*^T, mean_out[M] as agpn_predict.  Computed as the Schur
 * complement of a partial blocked Cholesky of the augmented matrix [K . ; K* K**] with the same
 * kernels as the likelihood.  train_jitter: NaN = the theta row's jitter is used for the training block
 * too; a number reproduces the legacy art.network quirk (art.py:320 uses the constructor's jitter for
 * the training block and the call's jitter for K**).  Dense (N + M)^2 workspace: meant for M up to a
 * few 10^4.  All pointers host.
 */
int agpn_predict_cov(agpn_ctx* ctx, const double* theta, int series, int M, const double* tstar,
                     double train_jitter, double* mean_out, double* cov_out, int* status, void* stream);

/*
 * Log-likelihood of one theta row AND its gradient with respect to every entry of the row (SURVEY.md section 8(f):
 * new capability -- the reference carries derivative kernel classes, weight.py:81-931, several of them wrong, and never
 * assembles a gradient).  dLL/dtheta = 1/2 sum_i tr((alpha_i alpha_i^T - K_i^-1) dK_i/dtheta), dLL/dphi = alpha_i^T dm_i/dphi;
 * K_i^-1 and alpha_i come from a partial blocked Cholesky of [[K_i, .], [I, 0]] on the factorisation kernels, the
 * derivative tiles are generated on the fly from closed forms (artgpn_b200/csrc/agpn_grad.cuh).  Supported classes: node /
 * weight kernels Constant, SquaredExponential, Periodic, QuasiPeriodic, Cosine, Exponential|Laplacian, Matern32, Matern52;
 * means Constant, Linear (and None); any other class is a usage error.  theta, ll_out, grad_out[D]: host.
 * *status as agpn_loglike_batch (ll = -inf / NaN, gradient undefined, when not OK).
 */
int agpn_loglike_grad(agpn_ctx* ctx, const double* theta, double* ll_out, double* grad_out, int* status, void* stream);

/*
 * Replaces one kernel object's __call__ (node.py / weight.py): element-wise value on a
 * [rows][cols] lag array r, or on t1[rows] x t2[cols] for Linear / Polynomial (r ignored).
 * Host pointers; runs on `device`.  square_quirk as above (1 iff rows == cols in the reference).
 */
int agpn_kernel_eval(int device, int family, int kind, const double* pars,
                     const double* r, const double* t1, const double* t2,
                     int rows, int cols, int square_quirk, double* out);

/*
 * Device-side building blocks, exported for tests / profiling of the Cholesky path alone:
 * in-place batched factorisation of nmat row-major [n_pad][n_pad] matrices already resident on
 * the device (lower triangle referenced; n_pad multiple of 128), optional right-hand sides
 * rhs[nmat][n_pad] overwritten with L^-1 rhs.  logdet[nmat] = sum log L_kk, quad[nmat] =
 * |L^-1 rhs|^2, info[nmat] = 0 / 1 (not PD).  All pointers device.
 */
int agpn_potrf_batched(int device, int nmat, int n_pad, double* A, double* rhs,
                       double* logdet, double* quad, int* info, void* stream);

/*
 * Replaces network.sample (arttwo.py:129-147: scipy multivariate_normal(0, K, allow_singular=True).rvs()):
 * out[nsamp][n] = L z_s with K (+ nugget I) = L L^T factorised by the batched Cholesky kernels and z[nsamp][n]
 * standard-normal draws supplied by the caller (the random stream stays on the host side, like the reference's).
 * allow_singular: a positive SEMI-definite K is shifted by the smallest nugget of 1e-14 max|K_ii| x 100^k (k <= 4) that makes
 * it factorisable; an indefinite K is an error (*nugget_used, may be NULL; 0 for a definite K).  Host pointers; runs on `device`.
 */
int agpn_mvn_sample(int device, int n, const double* K, int nsamp, const double* z, double* out, double* nugget_used);

/* Measured issue-rate peak of the FP64 tensor instruction (DMMA.8x8x4) on `device`, in TFLOP/s:
 * the denominator bench.py uses for the Cholesky roofline (MEASURED_PEAKS.json has no FP64 figure). */
int agpn_dmma_peak(int device, double* tflops);

/* Measured issue-rate peak of the INT8 tensor instruction (tcgen05.mma.kind::i8, M = 128, N = 256, operands resident
 * in shared memory) on `device`, in TOP/s: the denominator bench.py uses for the INT8-sliced trailing update
 * (MEASURED_PEAKS.json has no INT8 figure). */
int agpn_i8_peak(int device, double* tops);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
long long agpn_launch_count(void);

/* Last-call timing probe: device milliseconds spent per phase of the last
 * agpn_loglike_batch when profiling was enabled with agpn_set_profiling(ctx, 1).
 * phases: 0 mean residual, 1 assembly, 2 diagonal-block factor, 3 panel solve,
 * 4 trailing update, 5 finalise.  Returns 0 if unavailable. */
int agpn_set_profiling(agpn_ctx* ctx, int enable);
int agpn_phase_ms(agpn_ctx* ctx, double* ms6, long long* launches6);

/* Trailing-update path of this context (the reference has one path: LAPACK dpotrf inside scipy cho_factor,
 * arttwo.py:284): 0 = FP64 tensor instruction (DMMA.8x8x4) everywhere; bit 0 / bit 1 set = the narrow / wide
 * rank-k updates run as exact integer products of 7-bit slices on the INT8 tensor path (tcgen05.mma.kind::i8,
 * 56 bits below each row's scale -- FP64-equivalent, bit-reproducible).  Chosen at agpn_create from the
 * environment (AGPN_SYRK = dmma | i8 | i8_wide | i8_narrow).
 * agpn_slice_ms: device milliseconds / launches of the digit-slicing kernels of the last profiled call. */
int agpn_update_path(const agpn_ctx* ctx);
/* Path the last agpn_predict / agpn_predict_slice took: 1 = K* appended as block rows to the factorisation (INT8-sliced
 * updates; default), 0 = stand-alone predict_kernel on the FP64 tensor instruction (AGPN_PREDICT=dmma, or the automatic
 * fall-back when an entry of K* L^-T exceeds its row scale sqrt(k**_ii)), -1 = no prediction yet. */
int agpn_predict_path(const agpn_ctx* ctx);
int agpn_slice_ms(agpn_ctx* ctx, double* ms, long long* launches);
/* Diagnostics (context created with AGPN_TIMELINE=1 in the environment): CSV of start / end times, phase and sub-batch
 * stream of every launch of the last agpn_loglike_batch, measured WITHOUT serialising the streams. */
int agpn_timeline_dump(agpn_ctx* ctx, const char* path);
/* Work the last wave of agpn_loglike_batch executed on the INT8-sliced path.  The factor of a covariance whose correlation
 * length is short against the time span is numerically banded: panel entries below 2^-57 of their row scale have all eight
 * digits zero, and a K step in which either operand chunk is all-zero adds exact integer zeros -- the update leaves it out
 * (bit-identical results; AGPN_I8_SKIP=0 runs the dense schedule).  Panel-solve tile halves whose entries are bounded below
 * 2^-96 of their row scale are left out too (AGPN_TRSM_SKIP=0: never).
 * out[0] = K steps (128 x 128 x 32, 36 slice pairs each) of the dense schedule, out[1] = K steps executed,
 * out[2] = panel-solve tile halves, out[3] = halves left out, out[4] = the right (K = 128) halves among them (out: 5 doubles).
 * No reference counterpart (np.linalg / LAPACK is dense). */
int agpn_skip_stats(agpn_ctx* ctx, double* out);
/* Diagnostics (AGPN_MEGA_PROF=1): per-CTA means of the persistent task kernel's time slots of the last wave, in
 * microseconds (slots 1, 4, 7 are task counts): 0 update runs, 1 update tasks, 2 diagonal-block bodies, 3 their waits,
 * 4 diagonal-block tasks, 5 panel-solve bodies, 6 their waits, 7 panel-solve tasks (thread group 0), 8 whole kernel,
 * 9 cluster syncs before update runs, 10 producer waits for panels. */
int agpn_mega_prof(agpn_ctx* ctx, double* out16);

#ifdef __cplusplus
}
#endif
#endif /* ARTGPN_B200_H */
