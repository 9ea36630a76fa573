// C ABI of libartgpn_b200.so (see include/artgpn_b200.h for the contract and the reference
// code each entry point replaces).  Host-side orchestration only: context, workspace, kernel
// launches.  No torch types; streams arrive as void*.
#include "../../include/artgpn_b200.h"

#include <dlfcn.h>
#include <nccl.h>          // types only: the NCCL entry points are resolved with dlopen / dlsym at agpn_comm_init

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "agpn_assemble.cuh"
#include "agpn_chol.cuh"
#include "agpn_eval.cuh"
#include "agpn_predict.cuh"
#include "agpn_small.cuh"
#include "agpn_tma.cuh"
#include "agpn_i8.cuh"
#include "agpn_mega.cuh"
#include "agpn_grad.cuh"

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
// launches enqueued while a CUDA graph is being captured are counted here (they execute at every replay, not now)
static thread_local long long* t_capture_launches = nullptr;

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _b[512];                                                                       \
            snprintf(_b, sizeof(_b), "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(_e),   \
                     cudaGetErrorString(_e), __FILE__, __LINE__, #expr);                        \
            g_err = _b;                                                                         \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define USAGE_FAIL(msg)   \
    do {                  \
        g_err = (msg);    \
        return 2;         \
    } while (0)

#define LAUNCH_CHECK()                                        \
    do {                                                      \
        if (t_capture_launches) ++*t_capture_launches;        \
        else g_launches.fetch_add(1);                         \
        CUDA_TRY(cudaGetLastError());                         \
    } while (0)

// Every entry point runs on its context's device and puts the caller's current device back on every exit path
// (a network on cuda:1 must not silently switch the caller -- e.g. torch -- to cuda:1).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        err = cudaSetDevice(dev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ENTER_DEVICE(dev)          \
    DeviceGuard _dev_guard(dev);   \
    CUDA_TRY(_dev_guard.err)

enum { PH_MEAN = 0, PH_ASM = 1, PH_DIAG = 2, PH_TRSM = 3, PH_SYRK = 4, PH_FINAL = 5, PH_SLICE = 6, PH_COUNT = 7 };

struct agpn_ctx {
    int device = 0;
    int N = 0, p = 0, q = 0, nblk = 0, npad = 0;
    double* d_t = nullptr;
    double* d_y = nullptr;
    double* d_yerr = nullptr;
    double tmean = 0.0, t0 = 0.0;
    Structure S;
    bool has_structure = false;
    // batch workspace
    size_t cap_mats = 0;
    double* d_A = nullptr;
    double* d_Linv = nullptr;
    double* d_rhs = nullptr;
    double* d_logdet = nullptr;
    double* d_quad = nullptr;
    int* d_info = nullptr;     // info | kflag | rflag | refine, 4 * cap_mats
    double refine_ratio = 1.0e3;   // AGPN_REFINE_RATIO: pivot ratio of a diagonal block above which its panel is solved by substitution
    size_t cap_w = 0;
    double* d_theta = nullptr;
    double* d_out = nullptr;
    int* d_status = nullptr;
    size_t cap_D = 0;
    // prediction workspace
    size_t cap_linv_blocks = 0;
    double* d_LinvAll = nullptr;
    size_t cap_V = 0;
    double* d_V = nullptr;
    size_t cap_M = 0;
    double* d_pred = nullptr;  // tstar | mstar | kss | mean | std, 5 * cap_M
    // prediction on the factorisation kernels: [K ; K*] stacked (T x n_pad), its digit planes, rhs | row scales | flag
    bool predict_dmma = false;  // AGPN_PREDICT=dmma: the stand-alone predict_kernel (FP64 tensor instruction) only
    bool predict_test_wrap = false;   // AGPN_I8_TEST_WRAP=1 (fault injection for the tests): unit row scales for the K* rows
    int last_predict_path = -1; // 1: stacked factorisation path, 0: predict_kernel
    size_t cap_paug = 0, cap_ps8 = 0, cap_pvec = 0;
    double* d_paug = nullptr;
    signed char* d_ps8 = nullptr;
    double* d_pvec = nullptr;
    size_t cap_aug = 0;        // augmented [K . ; K* K**] matrix of agpn_predict_cov, doubles
    double* d_aug = nullptr;
    size_t cap_augv = 0;
    double* d_augv = nullptr;  // rhs of the augmented system | second theta row
    // profiling
    bool profiling = false;
    double phase_ms[PH_COUNT] = {0, 0, 0, 0, 0, 0, 0};
    long long phase_launches[PH_COUNT] = {0, 0, 0, 0, 0, 0, 0};
    struct Ev { cudaEvent_t a, b; int ph; int stream_id; };
    bool timeline = false;     // AGPN_TIMELINE=1: events around every launch WITHOUT serialising the sub-batch streams (agpn_timeline_dump)
    size_t timeline_used = 0;
    std::vector<Ev> evs;
    bool attrs_set = false;
    // TMA descriptors of the batch workspace d_A (box 128 x 16 and 64 x 16 doubles, 128-byte swizzle)
    bool use_tma = false;      // AGPN_SYRK=tma | tma_nomc selects the TMA trailing-update kernel (default: cp.async, faster by 1-4 %)
    bool maps_valid = false;
    int tma_mc = 1;            // multicast the shared A operand inside the CTA pair (AGPN_SYRK=tma_nomc: off)
    int tma_which = 3;         // bit 0: narrow updates, bit 1: wide updates (AGPN_SYRK=tma_narrow / tma_wide)
    CUtensorMap tmA, tmB;
    // INT8-sliced trailing update (agpn_i8.cuh), the default: bit 0 = narrow updates, bit 1 = wide updates
    // (AGPN_SYRK = i8 | i8_narrow | i8_wide; dmma | cpasync | tma* select the FP64 tensor instruction everywhere)
    int i8_which = 3;
    bool i8_persist = true;        // AGPN_I8_PERSIST=0: one CTA pair per tile instead of persistent clusters
    int i8_clusters = 74;          // persistent clusters per launch = SMs / 2
    int i8_clusters_max = 74;
    bool i8_clusters_fixed = false;   // AGPN_I8_CLUSTERS given
    bool i8_balance = false;          // AGPN_I8_BALANCE=1: equal tile counts per persistent cluster (measured: no gain at 16 walkers, -1.7 % at 128)
    bool fuse_slices = true;       // AGPN_I8_FUSE=0: separate i8_slice_kernel after every panel solve
    int i8_group = 8;              // group of the INT8 path: all block columns (pure left-looking) unless AGPN_GROUP is given
    signed char* d_S8 = nullptr;   // [cap_mats] digit planes of one group of block columns
    double* d_rscale = nullptr;    // [cap_mats][2][npad]
    int sgroup = 0;                // block columns the slice storage holds per matrix
    int group = 8;             // block columns per wide trailing update (AGPN_GROUP)
    bool group_fixed = false;  // AGPN_GROUP given: no adaptation to the batch size
    int group_eff = 8;         // group used by the current call (set from the TOTAL number of matrices)
    bool no_fast_asm = false;  // AGPN_GENERIC_ASM=1 forces the generic assembly kernel
    bool old_generic_asm = false;   // AGPN_GENERIC_ASM=2: the per-entry generic kernel (assemble_kernel) instead of assemble_block_kernel
    bool no_small = false;     // AGPN_NO_SMALL=1 disables the fused small-N kernel
    bool exact = false;        // agpn_set_exact: reference operation order in every kernel evaluation
    // sub-batch streams: the matrices of a batch are split over nsub streams so that one sub-batch's
    // small serial kernels (diagonal block, panel solve) overlap another's wide tensor update
    int nsub = 6;              // AGPN_STREAMS
    bool nsub_fixed = false;   // AGPN_STREAMS given: no adaptation to the batch size
    cudaStream_t sub[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr;
    // CUDA-graph replay of one wave (memsets, residual, assembly, all Cholesky launches on all sub-streams,
    // finalise): ~300 launches per step collapse into one graph launch.  AGPN_GRAPH=0 disables it.
    bool use_graph = true;
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_gin = nullptr, ev_gout = nullptr;
    unsigned long long struct_version = 0, ws_version = 0;
    struct GraphEntry { int Ww; unsigned long long sv, wv; int exact, group, nsub, tma, mc; cudaGraphExec_t exec; long long nlaunch; };
    std::vector<GraphEntry> graphs;
    cudaEvent_t ev_join[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // look-ahead of the INT8 left-looking factorisation (AGPN_LOOKAHEAD=0 disables): one side stream per sub-batch stream
    bool lookahead = true;
    bool lookahead_forced = false;     // AGPN_LOOKAHEAD=2: for every batch size (default: few matrices only)
    cudaStream_t side[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_la[8][2] = {};
    // zero-chunk masks of the digit planes (CholArgs::nzmask): the update skips K steps that would add exact zeros (AGPN_I8_SKIP=0: dense)
    bool i8_skip = true;
    unsigned long long* d_nzmask = nullptr;     // [cap_mats][nblk][nzwords]
    int nzwords = 0;
    int tile_order = -1;                        // AGPN_TILE_ORDER (CholArgs::tile_order); -1: by size
    bool tile_bounds = true;                    // AGPN_TILE_BOUNDS=0: the panel solve's bound test always scans the tile itself
    bool trsm_skip = true;                      // AGPN_TRSM_SKIP=0: panel solves never leave a negligible tile half out
    double* d_linvnorm = nullptr;               // [cap_mats][linv_blocks][2]
    double* d_tmaxA = nullptr;                  // [cap_mats][nblk (nblk + 1) / 2] bound on |K| per assembled tile (fast assembly kernel)
    unsigned long long* d_tmaxU = nullptr;      // [cap_mats][nblk][nblk] scaled maximum of the tiles the update has rewritten
    unsigned long long* d_pmask = nullptr;      // prediction: masks of the stacked factor | 2 counters | 2 inverse norms
    size_t cap_pmask = 0;
    int mask_nmat = 0;                          // matrices of the last wave that ran with masks (0: it ran dense)
    int wave_nmat = 0;                          // matrices of the last wave
    // panel solves with one CTA per 64-column half of a tile (AGPN_TRSM_SPLIT=0: one CTA per tile)
    bool trsm_split = true;
    bool trsm_split_fixed = false;  // AGPN_TRSM_SPLIT given: for every batch size (default: matrices x block columns < 2304 only)
    double* d_rslot = nullptr;     // [cap_mats][nblk][2][2][128] parked forward-solve sums of the split panel solves
    // persistent tile-task kernel (agpn_mega.cuh): one launch per wave instead of the per-column chain; opt-in
    bool use_mega = false;         // AGPN_MEGA=1 enables it (measured: 23.9 vs 21.1 ms at 128 walkers, 3.66 vs 3.24 ms at 16)
    int* d_mega = nullptr;         // fD | fP | fU | refine per block, ints
    size_t cap_mega = 0;
    int linv_blocks = 1;           // inverse blocks kept per matrix in d_Linv (nblk with the task kernel)
    unsigned long long* d_megaprof = nullptr;   // AGPN_MEGA_PROF=1: [CTAs][16] time / count slots of the task kernel (agpn_mega_prof)
    // walker sharding over the GPUs of one box (agpn_comm_init / agpn_loglike_batch_sharded)
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    size_t cap_gather = 0;
    double* d_gather = nullptr;    // [nranks][slot] log-likelihoods | same for the status words (as doubles' storage: ints)
    int* d_gstatus = nullptr;
};

// owning device allocation for the short-lived buffers of the test / API helpers (freed on every path)
struct DevBuf {
    double* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(double) * n); }
};

static double host_mean(const double* t, int n) {
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += t[i];
    return (double)(s / (long double)n);
}

extern "C" int agpn_version(void) { return 100; }
extern "C" const char* agpn_last_error(void) { return g_err.c_str(); }
extern "C" int agpn_kernel_npars(int family, int kind) { return kernel_npars(family, kind); }
extern "C" int agpn_mean_npars(int kind) { return mean_npars(kind); }
extern "C" long long agpn_launch_count(void) { return g_launches.load(); }

// ---- TMA tensor maps (driver entry point fetched through the runtime; libcuda is not linked) ----
typedef CUresult (*agpn_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static agpn_encode_tiled_fn get_encode_tiled() {
    static agpn_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (agpn_encode_tiled_fn)p;
    }
    return fn;
}

// maps over matrices [nmat][n_pad][ld] (row-major doubles): boxes of box_rows x 16 elements
static bool make_batch_map(CUtensorMap* out, double* base, size_t nmat, int n_pad, int ld, int box_rows) {
    agpn_encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)ld, (cuuint64_t)n_pad, (cuuint64_t)nmat};
    const cuuint64_t gstride[2] = {(cuuint64_t)ld * 8, (cuuint64_t)n_pad * ld * 8};
    const cuuint32_t box[3] = {GEMM_KC, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int i8_mode_from_env() {
    const char* e = getenv("AGPN_SYRK");
    if (!e || !*e) return 3;
    if (strncmp(e, "i8", 2) == 0) return strcmp(e, "i8_narrow") == 0 ? 1 : (strcmp(e, "i8_wide") == 0 ? 2 : 3);
    return 0;
}

static int set_kernel_attrs(agpn_ctx* c) {
    if (c->attrs_set) return 0;
    CUDA_TRY(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(syrk_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(syrk_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(syrk_i8_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(chol_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_SMEM_BYTES));
    // function attributes are per function (not per context): always allow the largest K0 configuration
    {
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, small_fused_kernel));
        CUDA_TRY(cudaFuncSetAttribute(small_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      232448 - (int)fa.sharedSizeBytes));
    }
    c->attrs_set = true;
    return 0;
}

extern "C" int agpn_create(agpn_ctx** out, int device, int N, int p, int q, const double* time,
                           const double* y, const double* yerr) {
    if (!out || !time || !y || !yerr) USAGE_FAIL("agpn_create: null pointer");
    if (N < 1 || p < 1 || p > AGPN_MAXP || q < 1 || q > AGPN_MAXQ)
        USAGE_FAIL("agpn_create: need N >= 1, 1 <= p <= 8 series, 1 <= q <= 8 nodes");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) USAGE_FAIL("agpn_create: no such CUDA device");
    ENTER_DEVICE(device);
    agpn_ctx* c = new agpn_ctx();
    c->device = device;
    c->N = N;
    c->p = p;
    c->q = q;
    c->nblk = (N + TB - 1) / TB;
    c->npad = c->nblk * TB;
    c->tmean = host_mean(time, N);
    c->t0 = time[0];
    if (const char* e = getenv("AGPN_GROUP")) { c->group = atoi(e) > 0 ? atoi(e) : 8; c->group_fixed = true; }
    if (const char* e = getenv("AGPN_GENERIC_ASM")) { c->no_fast_asm = atoi(e) != 0; c->old_generic_asm = atoi(e) == 2; }
    if (const char* e = getenv("AGPN_NO_SMALL")) c->no_small = atoi(e) != 0;
    if (const char* e = getenv("AGPN_SYRK")) { c->use_tma = strncmp(e, "tma", 3) == 0; c->tma_mc = strcmp(e, "tma_nomc") != 0;
        c->tma_which = strcmp(e, "tma_narrow") == 0 ? 1 : (strcmp(e, "tma_wide") == 0 ? 2 : 3); }
    c->i8_which = i8_mode_from_env();
    if (const char* e = getenv("AGPN_I8_FUSE")) c->fuse_slices = atoi(e) != 0;
    if (const char* e = getenv("AGPN_I8_PERSIST")) c->i8_persist = atoi(e) != 0;
    if (const char* e = getenv("AGPN_I8_BALANCE")) c->i8_balance = atoi(e) != 0;
    if (const char* e = getenv("AGPN_MEGA")) c->use_mega = atoi(e) != 0;
    if (const char* e = getenv("AGPN_TRSM_SPLIT")) { c->trsm_split = atoi(e) != 0; c->trsm_split_fixed = true; }
    if (const char* e = getenv("AGPN_I8_SKIP")) c->i8_skip = atoi(e) != 0;
    if (const char* e = getenv("AGPN_TRSM_SKIP")) c->trsm_skip = atoi(e) != 0;
    if (const char* e = getenv("AGPN_TILE_BOUNDS")) c->tile_bounds = atoi(e) != 0;
    if (const char* e = getenv("AGPN_TILE_ORDER")) c->tile_order = atoi(e);
    if (const char* e = getenv("AGPN_MEGA_PROF")) if (atoi(e) != 0) {
        if (cudaMalloc(&c->d_megaprof, sizeof(unsigned long long) * 16 * 2 * c->i8_clusters_max) != cudaSuccess) { cudaGetLastError(); c->d_megaprof = nullptr; }
    }
    if (const char* e = getenv("AGPN_PREDICT")) c->predict_dmma = strcmp(e, "dmma") == 0;
    if (const char* e = getenv("AGPN_I8_TEST_WRAP")) c->predict_test_wrap = atoi(e) != 0;
    if (const char* e = getenv("AGPN_REFINE_RATIO")) c->refine_ratio = atof(e);
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        c->i8_clusters = c->i8_clusters_max = sms / 2 > 0 ? sms / 2 : 1;
        if (const char* e = getenv("AGPN_I8_CLUSTERS")) if (atoi(e) > 0) { c->i8_clusters = atoi(e); c->i8_clusters_fixed = true; }
    }
    // the INT8 path is fastest purely left-looking (every C tile is visited once, measured: C3 25.1 ms with one group
    // of 16 vs 26.0 with groups of 8; C4 213 vs 249 ms); int32 accumulators bound the depth to 256 block columns
    c->i8_group = c->group_fixed ? c->group : c->nblk;
    if (c->i8_group > 256) c->i8_group = 256;
    memset(&c->S, 0, sizeof(Structure));
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_t, sizeof(double) * N)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_y, sizeof(double) * (size_t)N * p)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_yerr, sizeof(double) * (size_t)N * p)) != cudaSuccess ||
        (e = cudaMemcpy(c->d_t, time, sizeof(double) * N, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(c->d_y, y, sizeof(double) * (size_t)N * p, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(c->d_yerr, yerr, sizeof(double) * (size_t)N * p, cudaMemcpyHostToDevice)) != cudaSuccess) {
        g_err = std::string("agpn_create: ") + cudaGetErrorString(e);
        agpn_destroy(c);
        return 1;
    }
    if (set_kernel_attrs(c)) {
        agpn_destroy(c);
        return 1;
    }
    if (const char* e = getenv("AGPN_GRAPH")) c->use_graph = atoi(e) != 0;
    if (const char* e = getenv("AGPN_TIMELINE")) c->timeline = atoi(e) != 0;
    if (cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_gin, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_gout, cudaEventDisableTiming) != cudaSuccess)
        c->use_graph = false;
    if (const char* e = getenv("AGPN_STREAMS")) { c->nsub = atoi(e); c->nsub_fixed = true; }
    if (c->nsub < 1) c->nsub = 1;
    if (c->nsub > 8) c->nsub = 8;
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) c->nsub = 1;
    if (const char* e = getenv("AGPN_LOOKAHEAD")) { c->lookahead = atoi(e) != 0; c->lookahead_forced = atoi(e) == 2; }
    for (int i = 0; i < c->nsub && c->nsub > 1; ++i) {
        if (cudaStreamCreateWithFlags(&c->sub[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_la[i][0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_la[i][1], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
            g_err = "agpn_create: cannot create sub-batch streams";
            agpn_destroy(c);
            return 1;
        }
    }
    *out = c;
    return 0;
}

static void agpn_comm_release(agpn_ctx* c);

extern "C" int agpn_destroy(agpn_ctx* c) {
    if (!c) return 0;
    DeviceGuard _dev_guard(c->device);
    for (auto& e : c->evs) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (int i = 0; i < 8; ++i) {
        if (c->sub[i]) cudaStreamDestroy(c->sub[i]);
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->ev_la[i][0]) cudaEventDestroy(c->ev_la[i][0]);
        if (c->ev_la[i][1]) cudaEventDestroy(c->ev_la[i][1]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (auto& ge : c->graphs) cudaGraphExecDestroy(ge.exec);
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->ev_gin) cudaEventDestroy(c->ev_gin);
    if (c->ev_gout) cudaEventDestroy(c->ev_gout);
    cudaFree(c->d_t); cudaFree(c->d_y); cudaFree(c->d_yerr);
    cudaFree(c->d_A); cudaFree(c->d_Linv); cudaFree(c->d_rhs); cudaFree(c->d_logdet); cudaFree(c->d_quad);
    cudaFree(c->d_info); cudaFree(c->d_theta); cudaFree(c->d_out); cudaFree(c->d_status);
    cudaFree(c->d_S8); cudaFree(c->d_rscale);
    cudaFree(c->d_paug); cudaFree(c->d_ps8); cudaFree(c->d_pvec);
    cudaFree(c->d_LinvAll); cudaFree(c->d_V); cudaFree(c->d_pred); cudaFree(c->d_aug); cudaFree(c->d_augv);
    cudaFree(c->d_gather); cudaFree(c->d_gstatus); cudaFree(c->d_mega); cudaFree(c->d_megaprof); cudaFree(c->d_rslot); cudaFree(c->d_nzmask); cudaFree(c->d_linvnorm); cudaFree(c->d_pmask); cudaFree(c->d_tmaxA); cudaFree(c->d_tmaxU);
    agpn_comm_release(c);
    delete c;
    return 0;
}

extern "C" int agpn_set_structure(agpn_ctx* c, const int* node_kind, const int* weight_kind,
                                  const int* mean_nterms, const int* mean_kind, int use_means) {
    if (!c || !node_kind || !weight_kind) USAGE_FAIL("agpn_set_structure: null pointer");
    Structure S;
    memset(&S, 0, sizeof(S));
    S.p = c->p;
    S.q = c->q;
    S.use_means = use_means ? 1 : 0;
    int off = 0;
    for (int j = 0; j < c->q; ++j) {
        const int n = kernel_npars(0, node_kind[j]);
        if (n < 0) USAGE_FAIL("agpn_set_structure: unknown node kernel kind");
        S.node_kind[j] = node_kind[j];
        S.node_off[j] = off;
        off += n;
    }
    for (int i = 0; i < c->p * c->q; ++i) {
        const int n = kernel_npars(1, weight_kind[i]);
        if (n < 0) USAGE_FAIL("agpn_set_structure: unknown weight kernel kind");
        S.weight_kind[i] = weight_kind[i];
        S.weight_off[i] = off;
        off += n;
    }
    int nt = 0;
    for (int i = 0; i < c->p; ++i) {
        S.mean_first[i] = nt;
        if (use_means) {
            if (!mean_nterms) USAGE_FAIL("agpn_set_structure: mean_nterms is null");
            for (int u = 0; u < mean_nterms[i]; ++u) {
                if (nt >= AGPN_MAXTERMS) USAGE_FAIL("agpn_set_structure: more than 32 mean terms");
                const int n = mean_npars(mean_kind[nt]);
                if (n < 0) USAGE_FAIL("agpn_set_structure: unknown mean kind");
                S.mean_kind[nt] = mean_kind[nt];
                S.mean_off[nt] = off;
                off += n;
                ++nt;
            }
        }
    }
    S.mean_first[c->p] = nt;
    S.jit_off = off;
    off += c->p;
    S.D = off;
    c->S = S;
    c->has_structure = true;
    ++c->struct_version;
    return 0;
}

extern "C" int agpn_theta_dim(const agpn_ctx* c) { return (c && c->has_structure) ? c->S.D : -1; }

extern "C" int agpn_set_exact(agpn_ctx* c, int enable) {
    if (!c) USAGE_FAIL("agpn_set_exact: null context");
    c->exact = enable != 0;
    return 0;
}

extern "C" int agpn_set_profiling(agpn_ctx* c, int enable) {
    if (!c) USAGE_FAIL("agpn_set_profiling: null context");
    c->profiling = enable != 0;
    return 0;
}

extern "C" int agpn_phase_ms(agpn_ctx* c, double* ms6, long long* launches6) {
    if (!c) USAGE_FAIL("agpn_phase_ms: null context");
    for (int i = 0; i < 6; ++i) {
        if (ms6) ms6[i] = c->phase_ms[i];
        if (launches6) launches6[i] = c->phase_launches[i];
    }
    return 0;
}

extern "C" int agpn_slice_ms(agpn_ctx* c, double* ms, long long* launches) {
    if (!c) USAGE_FAIL("agpn_slice_ms: null context");
    if (ms) *ms = c->phase_ms[PH_SLICE];
    if (launches) *launches = c->phase_launches[PH_SLICE];
    return 0;
}

// AGPN_TIMELINE=1 diagnostics: start / end (ms after the first launch of the call) of every launch of the last
// agpn_loglike_batch, with its phase and sub-batch stream -- the concurrency picture of the sub-batch streams
// Work the last wave of agpn_loglike_batch actually executed on the INT8 path (read back from the zero-chunk masks):
// out[0] K steps (128 x 128 x 32, all 36 slice pairs) a dense schedule runs, out[1] K steps executed,
// out[2] panel-solve tile halves, out[3] of them left out by the bound test, out[4] the right (K = 128) halves among those.
// Dense run: out[1] = out[0], out[3] = out[4] = 0.
extern "C" int agpn_skip_stats(agpn_ctx* c, double* out) {
    if (!c || !out) USAGE_FAIL("agpn_skip_stats: null pointer");
    ENTER_DEVICE(c->device);
    const int nb = c->nblk, nm = c->mask_nmat;
    double nominal = 0.0, halves = 0.0;
    for (int k = 0; k < nb; ++k) {
        nominal += 4.0 * k * (nb - k);
        halves += 2.0 * (nb - k - 1);
    }
    if (nm == 0 || !c->d_nzmask) {
        out[0] = out[1] = nominal * c->wave_nmat; out[2] = halves * c->wave_nmat; out[3] = out[4] = 0.0;
        return 0;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    const size_t words = (size_t)nm * nb * c->nzwords;
    std::vector<unsigned long long> h(words + 2);
    CUDA_TRY(cudaMemcpy(h.data(), c->d_nzmask, sizeof(unsigned long long) * words, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&h[words], c->d_nzmask + (size_t)c->cap_mats * nb * c->nzwords, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double active = 0.0;
    for (int m = 0; m < nm; ++m)
        for (int k = 1; k < nb; ++k)
            for (int ti = k; ti < nb; ++ti)
                for (int w = 0; w * 64 < 4 * k; ++w) {
                    unsigned long long a = h[((size_t)m * nb + ti) * c->nzwords + w] & h[((size_t)m * nb + k) * c->nzwords + w];
                    const int rem = 4 * k - 64 * w;
                    if (rem < 64) a &= (1ull << rem) - 1ull;
                    active += (double)__builtin_popcountll(a);
                }
    out[0] = nominal * nm; out[1] = active; out[2] = halves * nm; out[3] = (double)(h[words] + h[words + 1]); out[4] = (double)h[words + 1];
    return 0;
}

extern "C" int agpn_timeline_dump(agpn_ctx* c, const char* path) {
    if (!c || !path) USAGE_FAIL("agpn_timeline_dump: null pointer");
    if (!c->timeline || c->timeline_used == 0) USAGE_FAIL("agpn_timeline_dump: no timeline (create the context with AGPN_TIMELINE=1)");
    FILE* f = fopen(path, "w");
    if (!f) USAGE_FAIL("agpn_timeline_dump: cannot open the output file");
    fprintf(f, "launch,phase,stream,start_ms,end_ms\n");
    for (size_t i = 0; i < c->timeline_used; ++i) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, c->evs[0].a, c->evs[i].a);
        cudaEventElapsedTime(&b, c->evs[0].a, c->evs[i].b);
        fprintf(f, "%zu,%d,%d,%.4f,%.4f\n", i, c->evs[i].ph, c->evs[i].stream_id, a, b);
    }
    fclose(f);
    return 0;
}

// AGPN_MEGA_PROF=1 diagnostics: per-CTA means of the task kernel's time slots (microseconds / counts) of the last wave:
// 0 U runs, 1 U tasks, 2 D body, 3 D wait, 4 D tasks, 5 P body, 6 P wait, 7 P tasks (group 0), 8 whole kernel,
// 9 cluster sync before U runs, 10 producer waits for panels
extern "C" int agpn_mega_prof(agpn_ctx* c, double* out16) {
    if (!c || !out16) USAGE_FAIL("agpn_mega_prof: null pointer");
    if (!c->d_megaprof) USAGE_FAIL("agpn_mega_prof: create the context with AGPN_MEGA_PROF=1");
    ENTER_DEVICE(c->device);
    const int nct = 2 * c->i8_clusters_max;
    std::vector<unsigned long long> h((size_t)16 * nct);
    CUDA_TRY(cudaMemcpy(h.data(), c->d_megaprof, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
    for (int s = 0; s < 16; ++s) {
        double acc = 0.0;
        for (int b = 0; b < nct; ++b) acc += (double)h[(size_t)b * 16 + s];
        const bool is_count = (s == 1 || s == 4 || s == 7);
        out16[s] = acc / nct / (is_count ? 1.0 : 1000.0);
    }
    return 0;
}

extern "C" int agpn_update_path(const agpn_ctx* c) { return c ? c->i8_which : -1; }
extern "C" int agpn_predict_path(const agpn_ctx* c) { return c ? c->last_predict_path : -1; }

// ---- profiling helpers ---------------------------------------------------------------------
struct PhaseScope {
    agpn_ctx* c;
    cudaStream_t st;
    size_t idx;
    bool on;
    PhaseScope(agpn_ctx* c_, cudaStream_t st_, int ph, size_t& cursor) : c(c_), st(st_), idx(0), on(c_->profiling || c_->timeline) {
        c->phase_launches[ph] += 1;
        if (!on) return;
        if (cursor >= c->evs.size()) {
            agpn_ctx::Ev e;
            cudaEventCreate(&e.a);
            cudaEventCreate(&e.b);
            e.ph = ph;
            c->evs.push_back(e);
        }
        idx = cursor++;
        c->evs[idx].ph = ph;
        c->evs[idx].stream_id = -1;
        for (int i = 0; i < 8; ++i) {
            if (st == c->sub[i] && st) c->evs[idx].stream_id = i;
            if (st == c->side[i] && st) c->evs[idx].stream_id = 8 + i;
        }
        cudaEventRecord(c->evs[idx].a, st);
    }
    ~PhaseScope() {
        if (on) cudaEventRecord(c->evs[idx].b, st);
    }
};

static void collect_phases(agpn_ctx* c, size_t used) {
    if (c->timeline) c->timeline_used = used;
    if (!c->profiling) return;
    for (size_t i = 0; i < used; ++i) {
        float ms = 0.f;
        cudaEventSynchronize(c->evs[i].b);
        cudaEventElapsedTime(&ms, c->evs[i].a, c->evs[i].b);
        c->phase_ms[c->evs[i].ph] += ms;
    }
}

// The persistent tile-task kernel covers the default configuration of the INT8-sliced path: purely left-looking (one group
// of all block columns), digit planes written by the panel solve, from three block columns on.
static bool mega_eligible(const agpn_ctx* c) {
    return c->use_mega && c->i8_which == 3 && c->i8_persist && c->fuse_slices && c->nblk > 2 && c->i8_group >= c->nblk &&
           !c->use_tma;
}

// one launch for the whole factorisation: flags = fD[nmat] | fP[nmat][nrows] | fU[nmat][nrows] (zeroed here);
// ncols block columns are eliminated (ncols < g.nblk: block rows appended below a square block, prediction)
static int launch_mega(agpn_ctx* c, CholArgs& g, int* flags, int ncols, cudaStream_t st) {
    const size_t nm = (size_t)g.nmat, nr = (size_t)g.nblk;
    CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int) * nm * (1 + 2 * nr), st));
    MegaArgs m;
    m.fD = flags; m.fP = flags + nm; m.fU = flags + nm + nm * nr;
    m.ncols = ncols;
    m.tflags = 1 | (g.keep_factor ? 0 : 2);
    m.prof = c->d_megaprof;
    if (m.prof) CUDA_TRY(cudaMemsetAsync(m.prof, 0, sizeof(unsigned long long) * 16 * 2 * c->i8_clusters_max, st));
    int ncl = c->i8_clusters_max;
    const long long work = (long long)g.nmat * g.nblk;           // no point in more clusters than the widest segment has tasks
    if (work < ncl) ncl = (int)work;
    if (ncl < 1) ncl = 1;
    chol_mega_kernel<<<2 * ncl, MEGA_THREADS, MEGA_SMEM_BYTES, st>>>(g, m);
    LAUNCH_CHECK();
    return 0;
}

// ---- workspace -------------------------------------------------------------------------------
static int ensure_batch_workspace(agpn_ctx* c, size_t nmat, size_t W) {
    if (nmat > c->cap_mats) {
        cudaFree(c->d_A); cudaFree(c->d_Linv); cudaFree(c->d_rhs); cudaFree(c->d_logdet); cudaFree(c->d_quad); cudaFree(c->d_info);
        cudaFree(c->d_S8); cudaFree(c->d_rscale);
        c->d_A = c->d_Linv = c->d_rhs = c->d_logdet = c->d_quad = c->d_rscale = nullptr;
        c->d_S8 = nullptr;
        c->d_info = nullptr;
        c->cap_mats = 0;
        const size_t np = (size_t)c->npad;
        CUDA_TRY(cudaMalloc(&c->d_A, sizeof(double) * nmat * np * np));
        c->maps_valid = c->use_tma && make_batch_map(&c->tmA, c->d_A, nmat, c->npad, c->npad, TB) &&
                        make_batch_map(&c->tmB, c->d_A, nmat, c->npad, c->npad, BN);
        c->linv_blocks = mega_eligible(c) ? c->nblk : 1;
        CUDA_TRY(cudaMalloc(&c->d_Linv, sizeof(double) * nmat * c->linv_blocks * TB * TB));
        // potrf_diag_kernel writes the lower block triangle of an inverse only: the zeros above it are put here once
        CUDA_TRY(cudaMemset(c->d_Linv, 0, sizeof(double) * nmat * c->linv_blocks * TB * TB));
        CUDA_TRY(cudaMalloc(&c->d_rhs, sizeof(double) * nmat * np));
        CUDA_TRY(cudaMalloc(&c->d_logdet, sizeof(double) * nmat));
        CUDA_TRY(cudaMalloc(&c->d_quad, sizeof(double) * nmat));
        CUDA_TRY(cudaMalloc(&c->d_info, sizeof(int) * 4 * nmat));
        cudaFree(c->d_nzmask);
        cudaFree(c->d_linvnorm);
        cudaFree(c->d_tmaxA);
        cudaFree(c->d_tmaxU);
        c->d_nzmask = nullptr;
        c->d_linvnorm = nullptr;
        c->d_tmaxA = nullptr;
        c->d_tmaxU = nullptr;
        if (c->i8_skip && c->fuse_slices && c->i8_which && c->nblk > 2) {
            const int sg = c->i8_group < 2 ? 2 : c->i8_group;
            c->nzwords = (4 * sg + 63) / 64;
            CUDA_TRY(cudaMalloc(&c->d_nzmask, sizeof(unsigned long long) * (nmat * c->nblk * c->nzwords + 2)));   // + 2 counters
            if (c->trsm_skip) {
                CUDA_TRY(cudaMalloc(&c->d_linvnorm, sizeof(double) * nmat * c->linv_blocks * 2));
                if (c->tile_bounds) {
                    CUDA_TRY(cudaMalloc(&c->d_tmaxA, sizeof(double) * nmat * ((size_t)c->nblk * (c->nblk + 1) / 2)));
                    CUDA_TRY(cudaMalloc(&c->d_tmaxU, sizeof(unsigned long long) * nmat * c->nblk * c->nblk));
                }
            }
        }
        cudaFree(c->d_rslot);
        c->d_rslot = nullptr;
        if (c->trsm_split && c->i8_which == 3 && c->nblk > 2)
            CUDA_TRY(cudaMalloc(&c->d_rslot, sizeof(double) * nmat * c->nblk * 4 * 128));
        if (mega_eligible(c)) {
            cudaFree(c->d_mega);
            c->d_mega = nullptr;
            c->cap_mega = 0;
            const size_t ni = nmat * (1 + 3 * (size_t)c->nblk);
            CUDA_TRY(cudaMalloc(&c->d_mega, sizeof(int) * ni));
            c->cap_mega = ni;
        }
        if (c->i8_which && c->nblk > 2) {
            c->sgroup = c->i8_group < 2 ? 2 : c->i8_group;
            CUDA_TRY(cudaMalloc(&c->d_S8, nmat * i8_slices_per_matrix(c->nblk, c->sgroup)));
            CUDA_TRY(cudaMalloc(&c->d_rscale, sizeof(double) * nmat * 2 * np));
        }
        c->cap_mats = nmat;
        ++c->ws_version;
    }
    if (W > c->cap_w || (size_t)c->S.D > c->cap_D) {
        cudaFree(c->d_theta); cudaFree(c->d_out); cudaFree(c->d_status);
        c->d_theta = c->d_out = nullptr;
        c->d_status = nullptr;
        c->cap_w = 0;
        const size_t D = (size_t)c->S.D;
        CUDA_TRY(cudaMalloc(&c->d_theta, sizeof(double) * W * D));
        CUDA_TRY(cudaMalloc(&c->d_out, sizeof(double) * W));
        CUDA_TRY(cudaMalloc(&c->d_status, sizeof(int) * W));
        c->cap_w = W;
        c->cap_D = D;
        ++c->ws_version;
    }
    return 0;
}

// factorise nmat matrices in place (+ optional rhs), all on `st`.
// Block columns are processed in groups of c->group: inside a group, block column k first takes
// the (narrow) update from the group's earlier columns, then its diagonal block is factorised and
// its panel solved; after the group the whole trailing matrix takes ONE update of depth
// 128 * group (wide).  group = 1 is the plain right-looking algorithm.
// look-ahead (side stream + two events, INT8 left-looking updates only): column k's diagonal tile is updated first and
// factorised on `st` while the update of the tiles below it runs on `side`; the panel solve joins both.
struct LookAhead { cudaStream_t side; cudaEvent_t fork, join; };
// persistent clusters of one update launch: every cluster walks the same number of tiles (+-1), so no cluster idles through
// a last partial wave -- the SMs a launch does not need stay free for the other sub-batch streams
// (AGPN_I8_BALANCE=1; default off = min(tiles, limit) clusters: measured 3.233 vs 3.245 ms at 16 walkers, 21.49 vs 21.12 ms at 128)
static int balanced_clusters(const agpn_ctx* c, int total) {
    const int lim = c->i8_clusters < 1 ? 1 : c->i8_clusters;
    if (total <= lim) return total;
    if (!c->i8_balance) return lim;
    const int waves = (total + lim - 1) / lim;
    return (total + waves - 1) / waves;
}
static int run_cholesky(agpn_ctx* c, CholArgs& g, cudaStream_t st, size_t* cursor, int kstop = -1, bool no_trailing = false,
                        const LookAhead* la = nullptr) {
    size_t local_cursor = 0;
    size_t& cur = cursor ? *cursor : local_cursor;
    // few small matrices: the left-looking narrow updates of a deep group have too few CTAs to fill the GPU
    // (measured: 12 matrices of 2048^2 -> group 2 is 7 % faster than 8; 48 of 2048^2 or 12 of 8192^2 -> 8);
    // callers set group_eff from matrices x block columns
    const int group = c->group_eff < 1 ? 1 : c->group_eff;
    // kstop < nblk: partial factorisation -- only the first kstop block columns are eliminated; the
    // remaining trailing block then holds the Schur complement (used for the predictive covariance)
    if (kstop < 0 || kstop > g.nblk) kstop = g.nblk;
    const int i8 = (g.S8 && group <= g.sgroup) ? c->i8_which : 0;
    // the zero-chunk masks are indexed from block column 0 and only the persistent update kernel reads them: one group,
    // fused digit extraction, persistent clusters -- otherwise everything runs dense (a panel tile left out by the bound
    // test would leave stale digit planes behind for a kernel that does not look at the masks)
    if (g.nzmask != nullptr && !(i8 != 0 && c->i8_persist && c->fuse_slices && group >= kstop)) {
        g.nzmask = nullptr;
        g.linvnorm = nullptr;
        g.tmaxA = nullptr;
        g.tmaxU = nullptr;
        g.nzstat = nullptr;
        c->mask_nmat = 0;
    }
    for (int k0 = 0; k0 < kstop; k0 += group) {
        const int kend = (k0 + group < kstop) ? k0 + group : kstop;
        for (int k = k0; k < kend; ++k) {
            if (k > k0) {
                PhaseScope ps(c, st, PH_SYRK, cur);
                if ((i8 & 1) && c->i8_persist && la != nullptr && g.nblk - k > 1) {
                    CUDA_TRY(cudaEventRecord(la->fork, st));
                    CUDA_TRY(cudaStreamWaitEvent(la->side, la->fork, 0));
                    const int ncl_d = balanced_clusters(c, g.nmat);
                    syrk_i8_persistent_kernel<<<2 * ncl_d, I8_PTHREADS, I8_SMEM_BYTES, st>>>(g, k - k0, k, 1, 1, g.nmat, 0);
                    LAUNCH_CHECK();
                    const int ntl = g.nblk - k - 1, total = ntl * g.nmat;
                    const int ncl = balanced_clusters(c, total);
                    syrk_i8_persistent_kernel<<<2 * ncl, I8_PTHREADS, I8_SMEM_BYTES, la->side>>>(g, k - k0, k, 1, ntl, total, 1);
                    CUDA_TRY(cudaEventRecord(la->join, la->side));
                } else if ((i8 & 1) && c->i8_persist) {
                    const int ntl = g.nblk - k, total = ntl * g.nmat;
                    const int ncl = balanced_clusters(c, total);
                    syrk_i8_persistent_kernel<<<2 * ncl, I8_PTHREADS, I8_SMEM_BYTES, st>>>(g, k - k0, k, 1, ntl, total, 0);
                } else if (i8 & 1)
                    syrk_i8_kernel<<<dim3(2 * (g.nblk - k), g.nmat), I8_THREADS, I8_SMEM_BYTES, st>>>(g, k0, k - k0, k, 1);
                else if (c->maps_valid && (c->tma_which & 1))
                    syrk_tma_kernel<<<dim3(2 * (g.nblk - k), g.nmat), TMA_THREADS, TMA_SMEM_BYTES, st>>>(c->tmA, c->tmB, g, k0, k - k0, k, 1, c->tma_mc);
                else
                    syrk_kernel<<<dim3(2 * (g.nblk - k), g.nmat), 256, GEMM_SMEM_BYTES, st>>>(g, k0, k - k0, k, 1);
                LAUNCH_CHECK();
            }
            {
                PhaseScope ps(c, st, PH_DIAG, cur);
                potrf_diag_kernel<<<g.nmat, 256, POTRF_SMEM_BYTES, st>>>(g, k);
                LAUNCH_CHECK();
            }
            const bool trailing = kend < g.nblk && !(no_trailing && kend == kstop);
            const bool need_slices = k < g.nblk - 1 && ((((i8 & 1) && k + 1 < kend)) || ((i8 & 2) && trailing));
            // the panel solve writes the digit planes itself; the FP64 panel is stored only if somebody reads it
            // (a DMMA update of a mixed mode, or a caller that wants the factor: g.keep_factor)
            const int tflags = (need_slices && c->fuse_slices) ? (1 | ((i8 == 3 && !g.keep_factor) ? 2 : 0)) : 0;
            if (k > k0 && (i8 & 1) && c->i8_persist && la != nullptr && g.nblk - k > 1)
                CUDA_TRY(cudaStreamWaitEvent(st, la->join, 0));
            if (k < g.nblk - 1) {
                PhaseScope ps(c, st, PH_TRSM, cur);
                // one CTA per 64-column half of a tile when the FP64 panel is not stored (no in-place overwrite to order) and
                // the forward-solve sums can be parked (g.rslot): half the latency per launch, finer-grained SM packing
                const bool split = (tflags & 2) && g.rslot != nullptr && g.rhs != nullptr;
                trsm_kernel<<<dim3((split ? 2 : 1) * (g.nblk - k - 1), g.nmat), 256, GEMM_SMEM_BYTES, st>>>(g, k, k0, tflags | (split ? 8 : 0));
                LAUNCH_CHECK();
            }
            // digit planes of the new panel, if a later update of this group reads them on the INT8 path
            if (need_slices && !c->fuse_slices) {
                PhaseScope ps(c, st, PH_SLICE, cur);
                i8_slice_kernel<<<dim3(g.nblk - k - 1, g.nmat), 256, 0, st>>>(g, k, k0);
                LAUNCH_CHECK();
            }
        }
        // no_trailing: the columns right of kstop do not exist (K* rows appended below a training block)
        const int nt = (no_trailing && kend == kstop) ? 0 : g.nblk - kend;
        if (nt > 0) {
            PhaseScope ps(c, st, PH_SYRK, cur);
            if ((i8 & 2) && c->i8_persist) {
                const int ntl = nt * (nt + 1) / 2, total = ntl * g.nmat;
                const int ncl = balanced_clusters(c, total);
                syrk_i8_persistent_kernel<<<2 * ncl, I8_PTHREADS, I8_SMEM_BYTES, st>>>(g, kend - k0, kend, 0, ntl, total, 0);
            } else if (i8 & 2)
                syrk_i8_kernel<<<dim3(nt * (nt + 1), g.nmat), I8_THREADS, I8_SMEM_BYTES, st>>>(g, k0, kend - k0, kend, 0);
            else if (c->maps_valid && (c->tma_which & 2))
                syrk_tma_kernel<<<dim3(nt * (nt + 1), g.nmat), TMA_THREADS, TMA_SMEM_BYTES, st>>>(c->tmA, c->tmB, g, k0, kend - k0, kend, 0, c->tma_mc);
            else
                syrk_kernel<<<dim3(nt * (nt + 1), g.nmat), 256, GEMM_SMEM_BYTES, st>>>(g, k0, kend - k0, kend, 0);
            LAUNCH_CHECK();
        }
    }
    return 0;
}

// the likelihood-path assembly has a specialised kernel when every node is SE / Periodic /
// QuasiPeriodic and every weight Constant / SE (the reference's example networks and the bench)
static bool fast_assembly_ok(const agpn_ctx* c, const AsmArgs& a, bool* wse) {
    if (!a.lower_only || !a.square || !a.add_err || !a.add_jit || !a.pad_identity || !a.vec2) return false;
    if (c->no_fast_asm || c->exact || c->S.q > ASMF_MAXQ || a.nser > 4 || a.rows_out % ASM_TILE || a.cols_out != a.rows_out) return false;
    *wse = false;
    for (int j = 0; j < c->S.q; ++j) {
        const int k = c->S.node_kind[j];
        if (k != K_SE && k != K_PERIODIC && k != K_QP) return false;
    }
    for (int s = 0; s < a.nser; ++s)
        for (int j = 0; j < c->S.q; ++j) {
            const int k = c->S.weight_kind[j + c->S.q * (a.series0 + s)];
            if (k == K_SE) *wse = true;
            else if (k != K_CONSTANT) return false;
        }
    return true;
}

template <int P>
static void launch_fast(const agpn_ctx* c, const AsmArgs& a, bool wse, dim3 grid, cudaStream_t st) {
    if (wse) assemble_fast_kernel<P, true><<<grid, 256, 0, st>>>(c->S, a);
    else assemble_fast_kernel<P, false><<<grid, 256, 0, st>>>(c->S, a);
}

static int launch_assemble(agpn_ctx* c, AsmArgs& a, int W, cudaStream_t st) {
    int tiles;
    if (a.lower_only) {
        const int nt = (a.rows_out + ASM_TILE - 1) / ASM_TILE;
        tiles = nt * (nt + 1) / 2;
    } else {
        tiles = ((a.rows_out + ASM_TILE - 1) / ASM_TILE) * ((a.cols_out + ASM_TILE - 1) / ASM_TILE);
    }
    bool wse = false;
    a.tmax_valid = 0;
    if (fast_assembly_ok(c, a, &wse)) {
        const dim3 grid(tiles, W);
        a.tmax_valid = (a.tmax != nullptr && a.lower_only) ? 1 : 0;
        switch (a.nser) {
            case 1: launch_fast<1>(c, a, wse, grid, st); break;
            case 2: launch_fast<2>(c, a, wse, grid, st); break;
            case 3: launch_fast<3>(c, a, wse, grid, st); break;
            default: launch_fast<4>(c, a, wse, grid, st); break;
        }
    } else if (!a.exact && a.nser <= 4 && !c->old_generic_asm) {
        // generic path, dispatch on the kernel class hoisted out of the entry loop
        const dim3 grid(tiles, W);
        switch (a.nser) {
            case 1: assemble_block_kernel<1><<<grid, 256, 0, st>>>(c->S, a); break;
            case 2: assemble_block_kernel<2><<<grid, 256, 0, st>>>(c->S, a); break;
            case 3: assemble_block_kernel<3><<<grid, 256, 0, st>>>(c->S, a); break;
            default: assemble_block_kernel<4><<<grid, 256, 0, st>>>(c->S, a); break;
        }
    } else {
        assemble_kernel<<<dim3(tiles, W), 256, 0, st>>>(c->S, a);
    }
    LAUNCH_CHECK();
    return 0;
}

// everything of one wave between "theta is on the device" and "results are in d_out / d_status", enqueued
// on `st` (plus the sub-batch streams, forked from and joined back into `st`)
static int enqueue_wave(agpn_ctx* c, int Ww, const double* d_theta, double* d_out, cudaStream_t st, size_t& cur) {
    const int p = c->p;
    const size_t np = (size_t)c->npad;
    const int nmat = Ww * p;
    const double half_n_log2pi = 0.5 * (double)c->N * std::log(2.0 * AGPN_PI);
    CUDA_TRY(cudaMemsetAsync(c->d_logdet, 0, sizeof(double) * nmat, st));
    CUDA_TRY(cudaMemsetAsync(c->d_quad, 0, sizeof(double) * nmat, st));
    CUDA_TRY(cudaMemsetAsync(c->d_info, 0, sizeof(int) * 4 * c->cap_mats, st));
    int* d_infoflag = c->d_info;
    int* d_kflag = c->d_info + c->cap_mats;
    int* d_rflag = c->d_info + 2 * c->cap_mats;
    {
        PhaseScope ps(c, st, PH_MEAN, cur);
        MeanArgs m;
        m.theta = d_theta; m.t = c->d_t; m.n = c->N; m.n_pad = c->npad; m.tmean = c->tmean; m.t0 = c->t0;
        m.y = c->d_y; m.out = c->d_rhs; m.series0 = 0; m.nser = p; m.subtract = 1; m.rflag = d_rflag;
        mean_kernel<<<dim3(p, Ww), 256, 0, st>>>(c->S, m);
        LAUNCH_CHECK();
    }
    const int small_nb = (c->N + 15) / 16;
    const bool use_small = !c->no_small && small_nb <= SM_MAXNB;
    bool tmax_valid = false;
    if (use_small) {
        // K0: assembly + factorisation + solve fused in one CTA's shared memory per (walker, series)
        PhaseScope ps(c, st, PH_ASM, cur);
        AsmArgs probe;
        probe.lower_only = 1; probe.square = 1; probe.add_err = 1; probe.add_jit = 1; probe.pad_identity = 1; probe.vec2 = 1;
        probe.nser = 1; probe.rows_out = c->npad; probe.cols_out = c->npad; probe.series0 = 0; probe.exact = 0;
        bool wse = false, fast = true;
        for (int s = 0; s < p && fast; ++s) {
            probe.series0 = s;
            fast = fast_assembly_ok(c, probe, &wse);
        }
        SmallArgs sa;
        sa.theta = d_theta; sa.t = c->d_t; sa.yerr = c->d_yerr; sa.rhs = c->d_rhs; sa.rhs_stride = c->npad;
        sa.N = c->N; sa.nb = small_nb; sa.fast = (fast && !c->exact) ? 1 : 0; sa.exact = c->exact ? 1 : 0; sa.torigin = c->t0;
        sa.logdet = c->d_logdet; sa.quad = c->d_quad; sa.info = d_infoflag; sa.kflag = d_kflag;
        const size_t smem = small_smem_bytes(small_nb, c->q);    // <= 227 KB for nb <= 12, q <= 8 (attribute set at create)
        small_fused_kernel<<<dim3(p, Ww), 256, smem, st>>>(c->S, sa);
        LAUNCH_CHECK();
    } else {
    {
        PhaseScope ps(c, st, PH_ASM, cur);
        AsmArgs a;
        a.theta = d_theta; a.ta = c->d_t; a.tb = c->d_t; a.na = c->N; a.nb = c->N;
        a.rows_out = c->npad; a.cols_out = c->npad; a.yerr = c->d_yerr; a.out = c->d_A;
        a.mat_stride = (long long)(np * np); a.ld = c->npad; a.series0 = 0; a.nser = p;
        a.lower_only = 1; a.square = 1; a.add_err = 1; a.add_jit = 1; a.pad_identity = 1;
        a.use_trig = 1; a.vec2 = 1; a.torigin = c->t0; a.kflag = d_kflag; a.exact = c->exact ? 1 : 0;
        a.tmax = c->d_tmaxA;
        if (launch_assemble(c, a, Ww, st)) return 1;
        tmax_valid = a.tmax_valid != 0;
    }
    CholArgs g;
    g.A = c->d_A; g.mat_stride = (long long)(np * np); g.ld = c->npad; g.nblk = c->nblk; g.nmat = nmat;
    g.Linv = c->d_Linv; g.linv_blocks = c->linv_blocks; g.rhs = c->d_rhs; g.logdet = c->d_logdet; g.quad = c->d_quad;
    g.info = d_infoflag;
    g.refine = c->d_info + 3 * c->cap_mats; g.refine_ratio = c->refine_ratio;
    g.A0 = c->d_A; g.mat0 = 0;
    g.store_diag = 0;                           // a wave's results are log-det and quadratic form; the next wave re-assembles A
    c->group_eff = (c->group_fixed || (long long)nmat * c->nblk >= 384) ? c->group : 2;
    // two block columns: a single depth-128 update, where the per-tile cost of the INT8 kernel still outweighs its
    // rate (measured at 3 x 256, 256 walkers: 0.77 vs 0.73 ms); from three block columns on it wins (3 x 384: 1.43 vs 1.45 ms)
    if (c->i8_which && c->d_S8 && c->nblk > 2) {
        c->group_eff = c->i8_group;
        // small batches: a launch that takes every SM serialises the sub-batch streams (their latency-bound diagonal
        // blocks and panel solves then wait for whole update launches); a third of the GPU per launch lets them
        // interleave (measured, 16 walkers of 3 x 2048: 3.45 -> 3.24 ms; 128 walkers: all SMs is fastest)
        if (!c->i8_clusters_fixed) {
            const long long work = (long long)nmat * c->nblk;
            const bool one_stream = c->profiling || c->nsub < 2 || nmat < 2 * c->nsub;      // nothing to interleave with
            c->i8_clusters = (one_stream || work >= 3072) ? c->i8_clusters_max
                                                          : (work >= 1536 ? (2 * c->i8_clusters_max) / 3 : c->i8_clusters_max / 3);
            if (c->i8_clusters < 1) c->i8_clusters = 1;
        }
        // row scales r_i = sqrt(K_ii) >= |L_ik| from the assembled diagonal, before the factorisation overwrites it
        g.S8 = c->d_S8; g.rscale = c->d_rscale; g.sgroup = c->sgroup;
        // split panel solves: on the pure INT8 likelihood path every panel solve of the wave leaves its FP64 panel unstored,
        // so every launch can be split (the task kernel keeps one CTA group per tile)
        const bool mega_wave = mega_eligible(c) && c->d_mega && !c->profiling && !c->timeline;
        // (measured with the banded schedule, 3 x 2048: split wins below ~ 40 walkers -- 16 walkers 2.11 vs 2.17 ms -- and loses
        // above, where half of its CTAs only find out that there is nothing to do: 128 walkers 11.29 vs 11.00 ms)
        // by tiles per launch, matrices x block columns: 3 x 8192, 32 walkers: 33.7 unsplit vs 34.8 ms split; 4 walkers: 7.96 vs 7.30)
        if (c->trsm_split && (c->trsm_split_fixed || (long long)nmat * c->nblk < 144 * 16) && c->d_rslot && c->i8_which == 3 && c->fuse_slices &&
            c->i8_group >= c->nblk && !mega_wave) g.rslot = c->d_rslot;
        c->mask_nmat = 0;
        c->wave_nmat = (int)nmat;
        if (c->d_nzmask && c->i8_group >= c->nblk && !mega_wave) {
            // masks are rebuilt by every wave's panel solves (one group of all block columns: kc32 = 4 k + chunk)
            g.nzmask = c->d_nzmask; g.nzwords = c->nzwords;
            // long block rows with few live tiles (3 x 8192: update 14.66 -> 13.56 ms) like the matrix-fastest walk, many short ones
            // do not (3 x 2048, 128 walkers: 3.63 -> 3.74 ms: the tiles of one matrix share their B planes in L2)
            // large batches of short matrices: tile-fastest with a skewed round-robin (3.61 -> 3.46 ms; no gain below ~ 50 walkers)
            g.tile_order = c->tile_order >= 0 ? c->tile_order : (c->nblk >= 32 ? 1 : (nmat >= 144 ? 2 : 0));
            g.linvnorm = c->d_linvnorm;
            if (c->d_linvnorm && c->d_tmaxU) {
                g.tmaxU = c->d_tmaxU;
                if (tmax_valid) g.tmaxA = c->d_tmaxA;
                CUDA_TRY(cudaMemsetAsync(c->d_tmaxU, 0, sizeof(unsigned long long) * (size_t)nmat * c->nblk * c->nblk, st));
            }
            g.nzstat = c->d_nzmask + (size_t)c->cap_mats * c->nblk * c->nzwords;
            c->mask_nmat = (int)nmat;
            CUDA_TRY(cudaMemsetAsync(c->d_nzmask, 0, sizeof(unsigned long long) * (size_t)nmat * c->nblk * c->nzwords, st));
            CUDA_TRY(cudaMemsetAsync(g.nzstat, 0, 2 * sizeof(unsigned long long), st));
        }
        PhaseScope ps(c, st, PH_SLICE, cur);
        i8_rowscale_kernel<<<dim3((c->npad + 255) / 256, nmat), 256, 0, st>>>(g, c->npad, nullptr, 0);
        LAUNCH_CHECK();
    }
    // sub-batch streams: six for small batches (latency-bound chains need neighbours to fill the GPU), fewer for large ones,
    // where every launch fills the GPU by itself and fewer, larger launches pay fewer tails (measured, 3 x 2048: 128 walkers
    // 20.60 / 20.66 / 20.80 ms with 3 / 4 / 6 streams, 64 walkers 10.56 vs 10.66 with 4 vs 6, 16 walkers 3.32 vs 3.20)
    int nsub_pref = c->nsub;
    // (banded schedule, shorter kernels: 128 walkers 10.53 / 10.78 ms with 2 / 3 streams, 64 walkers 5.71 / 5.76 with 3 / 4)
    if (!c->nsub_fixed) nsub_pref = nmat >= 288 ? 2 : (nmat >= 144 ? 3 : c->nsub);
    const int nsub = (c->profiling || nsub_pref < 2 || nmat < 2 * nsub_pref) ? 1 : nsub_pref;
    if (mega_eligible(c) && g.S8 && c->d_mega && !c->profiling && !c->timeline) {
        // the whole factorisation as ONE persistent launch of tile tasks (agpn_mega.cuh)
        g.refine = c->d_mega + (size_t)nmat * (1 + 2 * (size_t)c->nblk); g.refine_blocks = c->nblk;
        PhaseScope ps(c, st, PH_SYRK, cur);
        if (launch_mega(c, g, c->d_mega, c->nblk, st)) return 1;
    } else if (nsub == 1) {
        if (run_cholesky(c, g, st, &cur)) return 1;
    } else {
        // fork: every sub-stream waits for the assembly, factorises its slice of the matrices
        // (walker results do not depend on the split), and the caller's stream joins them all
        CUDA_TRY(cudaEventRecord(c->ev_fork, st));
        for (int si = 0; si < nsub; ++si) {
            const int m0 = (int)((long long)nmat * si / nsub), m1 = (int)((long long)nmat * (si + 1) / nsub);
            CholArgs gs = g;
            gs.A = g.A + (size_t)m0 * g.mat_stride;
            gs.Linv = g.Linv + (size_t)m0 * g.linv_blocks * TB * TB;
            gs.rhs = g.rhs + (size_t)m0 * np;
            gs.logdet = g.logdet + m0;
            gs.quad = g.quad + m0;
            gs.info = g.info + m0;
            gs.refine = g.refine + (size_t)m0 * g.refine_blocks;
            if (g.rslot) gs.rslot = g.rslot + rslot_off(g.nblk, m0, 0, 0, 0);
            if (g.nzmask) gs.nzmask = g.nzmask + (size_t)m0 * g.nblk * g.nzwords;
            if (g.linvnorm) gs.linvnorm = g.linvnorm + (size_t)m0 * g.linv_blocks * 2;
            if (g.tmaxA) gs.tmaxA = g.tmaxA + (size_t)m0 * ((size_t)g.nblk * (g.nblk + 1) / 2);
            if (g.tmaxU) gs.tmaxU = g.tmaxU + (size_t)m0 * g.nblk * g.nblk;
            gs.nmat = m1 - m0;
            gs.mat0 = m0;
            if (g.S8) {
                gs.S8 = g.S8 + (size_t)m0 * i8_slices_per_matrix(g.nblk, g.sgroup);
                gs.rscale = g.rscale + (size_t)m0 * 2 * np;
            }
            CUDA_TRY(cudaStreamWaitEvent(c->sub[si], c->ev_fork, 0));
            const LookAhead la = {c->side[si], c->ev_la[si][0], c->ev_la[si][1]};
            // measured (3 x 2048): +4 % at 8 walkers, nothing at 16, -2 % at 128; 3 x 8192, 4 walkers: +2 %
            const bool use_la = c->lookahead && gs.S8 && (c->lookahead_forced || nmat <= 48);
            if (run_cholesky(c, gs, c->sub[si], c->timeline ? &cur : nullptr, -1, false, use_la ? &la : nullptr)) return 1;
            CUDA_TRY(cudaEventRecord(c->ev_join[si], c->sub[si]));
            CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join[si], 0));
        }
    }
    }   // !use_small
    {
        PhaseScope ps(c, st, PH_FINAL, cur);
        finalize_kernel<<<(Ww + 127) / 128, 128, 0, st>>>(Ww, p, half_n_log2pi, c->d_quad, c->d_logdet, d_infoflag,
                                                         d_kflag, d_rflag, d_out, c->d_status);
        LAUNCH_CHECK();
    }
    return 0;
}

// Replay (or capture once) the CUDA graph of one wave.  theta is first copied into the context's own
// buffer and the results are produced in the context's d_out, so the graph's pointers never change; the
// caller's stream and the internal graph stream are ordered by two events.
static int run_wave_graph(agpn_ctx* c, int Ww, const double* theta, int theta_on_device, double* d_out_user,
                          cudaStream_t st, bool* replayed) {
    *replayed = false;
    const int D = c->S.D;
    const int tma = c->maps_valid ? 1 : 0;
    agpn_ctx::GraphEntry* hit = nullptr;
    for (auto& ge : c->graphs)
        if (ge.Ww == Ww && ge.sv == c->struct_version && ge.wv == c->ws_version && ge.exact == (int)c->exact &&
            ge.group == c->group && ge.nsub == c->nsub && ge.tma == tma && ge.mc == c->tma_mc)
            hit = &ge;
    if (!hit) {
        // stale entries (other structure / workspace) are dropped; at most a handful of batch sizes are kept
        for (size_t i = 0; i < c->graphs.size();) {
            if (c->graphs[i].sv != c->struct_version || c->graphs[i].wv != c->ws_version || c->graphs.size() > 8) {
                cudaGraphExecDestroy(c->graphs[i].exec);
                c->graphs.erase(c->graphs.begin() + i);
            } else ++i;
        }
        long long captured = 0;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(c->gstream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            return 0;                                  // cannot capture here: the caller falls back to plain launches
        }
        size_t cur = 0;
        t_capture_launches = &captured;
        const int rc = enqueue_wave(c, Ww, c->d_theta, c->d_out, c->gstream, cur);
        t_capture_launches = nullptr;
        const cudaError_t ee = cudaStreamEndCapture(c->gstream, &graph);
        if (rc || ee != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc) return rc;
            c->use_graph = false;
            return 0;
        }
        agpn_ctx::GraphEntry ge;
        ge.Ww = Ww; ge.sv = c->struct_version; ge.wv = c->ws_version; ge.exact = c->exact; ge.group = c->group;
        ge.nsub = c->nsub; ge.tma = tma; ge.mc = c->tma_mc; ge.exec = nullptr; ge.nlaunch = captured;
        const cudaError_t ie = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            cudaGetLastError();
            c->use_graph = false;
            return 0;
        }
        c->graphs.push_back(ge);
        hit = &c->graphs.back();
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_theta, theta, sizeof(double) * (size_t)Ww * D,
                             theta_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(c->ev_gin, st));
    CUDA_TRY(cudaStreamWaitEvent(c->gstream, c->ev_gin, 0));
    CUDA_TRY(cudaGraphLaunch(hit->exec, c->gstream));
    g_launches.fetch_add(hit->nlaunch);
    CUDA_TRY(cudaEventRecord(c->ev_gout, c->gstream));
    CUDA_TRY(cudaStreamWaitEvent(st, c->ev_gout, 0));
    if (d_out_user != c->d_out)
        CUDA_TRY(cudaMemcpyAsync(d_out_user, c->d_out, sizeof(double) * Ww, cudaMemcpyDeviceToDevice, st));
    *replayed = true;
    return 0;
}

extern "C" int agpn_loglike_batch(agpn_ctx* c, int W, const double* theta, int theta_on_device,
                                  double* out, int out_on_device, int* status, void* stream) {
    if (!c || !theta || !out) USAGE_FAIL("agpn_loglike_batch: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_loglike_batch: call agpn_set_structure first");
    if (W < 1) USAGE_FAIL("agpn_loglike_batch: W must be >= 1");
    ENTER_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int p = c->p, D = c->S.D;
    const size_t np = (size_t)c->npad;

    // walkers per wave: bounded by free memory and by grid.y.  cudaMemGetInfo is slow (it can take
    // milliseconds), so it is only consulted when the workspace has to grow.
    size_t wave = (size_t)W;
    if (wave * p > 65535) wave = 65535 / p;
    if (wave * p > c->cap_mats) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t have = c->cap_mats * np * np * sizeof(double);
        const size_t per_walker = (size_t)p * ((np * np + TB * TB + 3 * np) * sizeof(double) +
                                               ((c->i8_which && c->nblk > 2) ? i8_slices_per_matrix(c->nblk, c->i8_group < 2 ? 2 : c->i8_group) : 0)) + 64;
        const size_t fit = (size_t)((free_b + have) * 0.85) / per_walker;
        if (fit < 1) USAGE_FAIL("agpn_loglike_batch: not enough device memory for one walker");
        if (wave > fit) wave = fit;
    }
    if (ensure_batch_workspace(c, wave * p, wave)) return 1;

    for (int ph = 0; ph < PH_COUNT; ++ph) {
        c->phase_ms[ph] = 0;
        c->phase_launches[ph] = 0;
    }
    const double half_n_log2pi = 0.5 * (double)c->N * std::log(2.0 * AGPN_PI);

    for (size_t w0 = 0; w0 < (size_t)W; w0 += wave) {
        const int Ww = (int)std::min(wave, (size_t)W - w0);
        size_t cur = 0;
        double* d_out = out_on_device ? out + w0 : c->d_out;
        bool replayed = false;
        if (c->use_graph && !c->profiling && !c->timeline) {
            const int grc = run_wave_graph(c, Ww, theta + w0 * D, theta_on_device, d_out, st, &replayed);
            if (grc) return grc;
        }
        if (!replayed) {
            const double* d_theta = theta + w0 * D;
            if (!theta_on_device) {
                CUDA_TRY(cudaMemcpyAsync(c->d_theta, theta + w0 * D, sizeof(double) * Ww * D, cudaMemcpyHostToDevice, st));
                d_theta = c->d_theta;
            }
            if (enqueue_wave(c, Ww, d_theta, d_out, st, cur)) return 1;
        }
        if (!out_on_device)
            CUDA_TRY(cudaMemcpyAsync(out + w0, c->d_out, sizeof(double) * Ww, cudaMemcpyDeviceToHost, st));
        if (status) CUDA_TRY(cudaMemcpyAsync(status + w0, c->d_status, sizeof(int) * Ww, cudaMemcpyDeviceToHost, st));
        const bool more = w0 + wave < (size_t)W;
        if (!out_on_device || status || c->profiling || c->timeline || more) CUDA_TRY(cudaStreamSynchronize(st));
        collect_phases(c, cur);
    }
    return 0;
}

// ---- walker sharding over NCCL -------------------------------------------------------
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process has already loaded -- e.g. torch's -- is
// reused, so ids made by either side are compatible); the library itself has no link-time NCCL dependency.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
        }
    }
    return api.ok ? &api : nullptr;
}
#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            g_err = std::string("NCCL error at ") + #expr + ": " + nccl_api()->GetErrorString(_r);  \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

static void agpn_comm_release(agpn_ctx* c) {
    if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
    c->comm = nullptr;
    c->rank = 0;
    c->nranks = 1;
}

extern "C" int agpn_comm_unique_id(void* id_out) {
    if (!id_out) USAGE_FAIL("agpn_comm_unique_id: null pointer");
    NcclApi* n = nccl_api();
    if (!n) USAGE_FAIL("agpn_comm_unique_id: libnccl.so.2 cannot be loaded");
    ncclUniqueId id;
    NCCL_TRY(n->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

extern "C" int agpn_comm_init(agpn_ctx* c, const void* nccl_unique_id, int rank, int nranks) {
    if (!c || !nccl_unique_id) USAGE_FAIL("agpn_comm_init: null pointer");
    if (nranks < 1 || rank < 0 || rank >= nranks) USAGE_FAIL("agpn_comm_init: need 0 <= rank < nranks");
    NcclApi* n = nccl_api();
    if (!n) USAGE_FAIL("agpn_comm_init: libnccl.so.2 cannot be loaded");
    ENTER_DEVICE(c->device);
    agpn_comm_release(c);
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    NCCL_TRY(n->CommInitRank(&c->comm, nranks, id, rank));
    c->rank = rank;
    c->nranks = nranks;
    return 0;
}

extern "C" int agpn_comm_rank(const agpn_ctx* c, int* rank, int* nranks) {
    if (!c) USAGE_FAIL("agpn_comm_rank: null context");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return 0;
}

// rows [lo, hi) of rank r: contiguous blocks, the first W % nranks ranks get one more
static void shard_bounds(int W, int nranks, int r, int* lo, int* hi) {
    const int base = W / nranks, rem = W % nranks;
    *lo = r * base + (r < rem ? r : rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

// gathered [nranks][slot] -> contiguous [W] (uneven shards leave holes at the end of the short ranks' slots)
__global__ void shard_compact_kernel(const double* g, const int* gs, int W, int nranks, int slot, double* out, int* status) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const int base = W / nranks, rem = W % nranks;
    int r = (base + 1) * rem > w ? w / (base + 1) : rem + (base > 0 ? (w - (base + 1) * rem) / base : 0);
    const int lo = r * base + (r < rem ? r : rem);
    out[w] = g[(size_t)r * slot + (w - lo)];
    if (status) status[w] = gs[(size_t)r * slot + (w - lo)];
}

extern "C" int agpn_loglike_batch_sharded(agpn_ctx* c, int W, const double* theta, int theta_on_device, double* out,
                                          int out_on_device, int* status, void* stream) {
    if (!c || !theta || !out) USAGE_FAIL("agpn_loglike_batch_sharded: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_loglike_batch_sharded: call agpn_set_structure first");
    if (!c->comm) USAGE_FAIL("agpn_loglike_batch_sharded: call agpn_comm_init first");
    if (W < 1) USAGE_FAIL("agpn_loglike_batch_sharded: W must be >= 1");
    NcclApi* n = nccl_api();
    ENTER_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    int lo, hi;
    shard_bounds(W, c->nranks, c->rank, &lo, &hi);
    const int slot = (W + c->nranks - 1) / c->nranks;
    const size_t need = (size_t)slot * c->nranks + (size_t)W;
    if (need > c->cap_gather) {
        cudaFree(c->d_gather); cudaFree(c->d_gstatus); cudaFree(c->d_mega); cudaFree(c->d_megaprof);
        c->d_gather = nullptr; c->d_gstatus = nullptr; c->cap_gather = 0;
        CUDA_TRY(cudaMalloc(&c->d_gather, sizeof(double) * need));
        CUDA_TRY(cudaMalloc(&c->d_gstatus, sizeof(int) * need));
        c->cap_gather = need;
    }
    // steady state (even shards, workspace in place, device output, no status): the wave's own result buffer is the send
    // buffer and the caller's `out` the receive buffer -- no staging copies around the collective
    if (W % c->nranks == 0 && out_on_device && !status && (size_t)(hi - lo) * c->p <= c->cap_mats && (size_t)(hi - lo) <= c->cap_w &&
        (size_t)(hi - lo) * c->p <= 65535) {
        const int rc = agpn_loglike_batch(c, hi - lo, theta + (size_t)lo * c->S.D, theta_on_device, c->d_out, 1, nullptr, stream);
        if (rc) return rc;
        NCCL_TRY(n->AllGather(c->d_out, out, (size_t)slot, ncclDouble, c->comm, st));
        return 0;
    }
    double* mine = c->d_gather + (size_t)c->rank * slot;          // the likelihood kernels write straight into the send slot
    int* mine_s = c->d_gstatus + (size_t)c->rank * slot;
    if (hi > lo) {
        std::vector<int> hs(status ? hi - lo : 0);                 // a batch may run in several waves: statuses come back per call
        const int rc = agpn_loglike_batch(c, hi - lo, theta + (size_t)lo * c->S.D, theta_on_device, mine, 1,
                                          status ? hs.data() : nullptr, stream);
        if (rc) return rc;
        if (status) CUDA_TRY(cudaMemcpy(mine_s, hs.data(), sizeof(int) * (hi - lo), cudaMemcpyHostToDevice));
    }
    NCCL_TRY(n->AllGather(mine, c->d_gather, (size_t)slot, ncclDouble, c->comm, st));
    if (status) NCCL_TRY(n->AllGather(mine_s, c->d_gstatus, (size_t)slot, ncclInt32, c->comm, st));
    double* d_full = c->d_gather + (size_t)slot * c->nranks;
    int* d_fulls = c->d_gstatus + (size_t)slot * c->nranks;
    const bool even = (W % c->nranks) == 0;
    const double* src = c->d_gather;
    const int* srcs = c->d_gstatus;
    if (!even) {
        shard_compact_kernel<<<(W + 255) / 256, 256, 0, st>>>(c->d_gather, c->d_gstatus, W, c->nranks, slot, d_full,
                                                             status ? d_fulls : nullptr);
        LAUNCH_CHECK();
        src = d_full;
        srcs = d_fulls;
    }
    CUDA_TRY(cudaMemcpyAsync(out, src, sizeof(double) * W, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, srcs, sizeof(int) * W, cudaMemcpyDeviceToHost, st));
    if (!out_on_device || status) CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int agpn_covariance_block(agpn_ctx* c, const double* theta, int series, const double* ta, int na,
                                     const double* tb, int nb, int square_quirk, int add_errors, double* out,
                                     int out_on_device, void* stream) {
    if (!c || !theta || !out) USAGE_FAIL("agpn_covariance_block: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_covariance_block: call agpn_set_structure first");
    if (series < 0 || series >= c->p) USAGE_FAIL("agpn_covariance_block: series out of range");
    if (!ta) na = c->N;
    if (!tb) nb = c->N;
    if (na < 1 || nb < 1) USAGE_FAIL("agpn_covariance_block: empty grid");
    if (add_errors && (na != c->N || nb != c->N)) USAGE_FAIL("agpn_covariance_block: add_errors needs the training grid size");
    ENTER_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    double *d_th = nullptr, *d_ta = nullptr, *d_tb = nullptr, *d_out = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_th);
        if (ta) cudaFree(d_ta);
        if (tb) cudaFree(d_tb);
        if (!out_on_device) cudaFree(d_out);
    };
#define CB_TRY(expr)                                                                   \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            g_err = std::string("agpn_covariance_block: ") + cudaGetErrorString(_e);   \
            cleanup();                                                                 \
            return 1;                                                                  \
        }                                                                              \
    } while (0)
    CB_TRY(cudaMalloc(&d_th, sizeof(double) * c->S.D));
    CB_TRY(cudaMemcpyAsync(d_th, theta, sizeof(double) * c->S.D, cudaMemcpyHostToDevice, st));
    if (ta) {
        CB_TRY(cudaMalloc(&d_ta, sizeof(double) * na));
        CB_TRY(cudaMemcpyAsync(d_ta, ta, sizeof(double) * na, cudaMemcpyHostToDevice, st));
    } else d_ta = c->d_t;
    if (tb) {
        CB_TRY(cudaMalloc(&d_tb, sizeof(double) * nb));
        CB_TRY(cudaMemcpyAsync(d_tb, tb, sizeof(double) * nb, cudaMemcpyHostToDevice, st));
    } else d_tb = c->d_t;
    if (out_on_device) d_out = out;
    else CB_TRY(cudaMalloc(&d_out, sizeof(double) * (size_t)na * nb));
    AsmArgs a;
    a.theta = d_th; a.ta = d_ta; a.tb = d_tb; a.na = na; a.nb = nb; a.rows_out = na; a.cols_out = nb;
    a.yerr = c->d_yerr; a.out = d_out; a.mat_stride = (long long)na * nb; a.ld = nb; a.series0 = series; a.nser = 1;
    a.lower_only = 0; a.square = square_quirk ? 1 : 0; a.add_err = add_errors ? 1 : 0; a.add_jit = 0;
    a.pad_identity = 0; a.use_trig = 0; a.vec2 = ((nb % 2) == 0 && ((uintptr_t)d_out % 16) == 0) ? 1 : 0;
    a.torigin = c->t0; a.kflag = nullptr; a.exact = c->exact ? 1 : 0;
    rc = launch_assemble(c, a, 1, st);
    if (rc == 0 && !out_on_device)
        CB_TRY(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)na * nb, cudaMemcpyDeviceToHost, st));
    CB_TRY(cudaStreamSynchronize(st));
    cleanup();
#undef CB_TRY
    return rc;
}

extern "C" int agpn_mean_table(agpn_ctx* c, const double* theta, const double* t, int n, double* out) {
    if (!c || !theta || !out) USAGE_FAIL("agpn_mean_table: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_mean_table: call agpn_set_structure first");
    if (!t) n = c->N;
    if (n < 1) USAGE_FAIL("agpn_mean_table: empty grid");
    ENTER_DEVICE(c->device);
    DevBuf b_th, b_t, b_o;
    CUDA_TRY(b_th.alloc(c->S.D));
    CUDA_TRY(b_o.alloc((size_t)c->p * n));
    double *d_th = b_th.p, *d_o = b_o.p;
    CUDA_TRY(cudaMemcpy(d_th, theta, sizeof(double) * c->S.D, cudaMemcpyHostToDevice));
    MeanArgs m;
    if (t) {
        CUDA_TRY(b_t.alloc(n));
        CUDA_TRY(cudaMemcpy(b_t.p, t, sizeof(double) * n, cudaMemcpyHostToDevice));
        m.t = b_t.p; m.tmean = host_mean(t, n); m.t0 = t[0];
    } else {
        m.t = c->d_t; m.tmean = c->tmean; m.t0 = c->t0;
    }
    m.theta = d_th; m.n = n; m.n_pad = n; m.y = nullptr; m.out = d_o; m.series0 = 0; m.nser = c->p; m.subtract = 0;
    m.rflag = nullptr;
    mean_kernel<<<dim3(c->p, 1), 256>>>(c->S, m);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpy(out, d_o, sizeof(double) * (size_t)c->p * n, cudaMemcpyDeviceToHost));
    return 0;
}

// mean functions on a grid without a context: the kinds / parameters travel with the call, so neither a context's
// structure nor its cached CUDA graphs are touched (network._mean, MeanModel.__call__)
extern "C" int agpn_mean_eval(int device, int p, const int* mean_nterms, const int* mean_kind, const double* pars,
                              const double* t, int n, double* out) {
    if (!mean_nterms || !t || !out) USAGE_FAIL("agpn_mean_eval: null pointer");
    if (p < 1 || p > AGPN_MAXP) USAGE_FAIL("agpn_mean_eval: need 1 <= p <= 8 series");
    if (n < 1) USAGE_FAIL("agpn_mean_eval: empty grid");
    Structure S;
    memset(&S, 0, sizeof(S));
    S.p = p; S.q = 1; S.use_means = 1;
    int nt = 0, off = 0;
    for (int i = 0; i < p; ++i) {
        S.mean_first[i] = nt;
        for (int u = 0; u < mean_nterms[i]; ++u) {
            if (nt >= AGPN_MAXTERMS) USAGE_FAIL("agpn_mean_eval: more than 32 mean terms");
            if (!mean_kind || !pars) USAGE_FAIL("agpn_mean_eval: null pointer");
            const int np_ = mean_npars(mean_kind[nt]);
            if (np_ < 0) USAGE_FAIL("agpn_mean_eval: unknown mean kind");
            S.mean_kind[nt] = mean_kind[nt];
            S.mean_off[nt] = off;
            off += np_;
            ++nt;
        }
    }
    S.mean_first[p] = nt;
    S.jit_off = off;
    S.D = off > 0 ? off : 1;
    ENTER_DEVICE(device);
    DevBuf b_th, b_t, b_o;
    CUDA_TRY(b_th.alloc(S.D));
    CUDA_TRY(b_t.alloc(n));
    CUDA_TRY(b_o.alloc((size_t)p * n));
    if (off > 0) CUDA_TRY(cudaMemcpy(b_th.p, pars, sizeof(double) * off, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b_t.p, t, sizeof(double) * n, cudaMemcpyHostToDevice));
    MeanArgs m;
    m.theta = b_th.p; m.t = b_t.p; m.n = n; m.n_pad = n; m.tmean = host_mean(t, n); m.t0 = t[0]; m.y = nullptr; m.out = b_o.p;
    m.series0 = 0; m.nser = p; m.subtract = 0; m.rflag = nullptr;
    mean_kernel<<<dim3(p, 1), 256>>>(S, m);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpy(out, b_o.p, sizeof(double) * (size_t)p * n, cudaMemcpyDeviceToHost));
    return 0;
}

static int predict_impl(agpn_ctx* c, const double* theta, int series, int M, const double* tstar, bool have_grid,
                        int grid_M, int grid_off, double grid_mean, double grid_t0, double* mean_out, double* std_out,
                        int* status, void* stream);

extern "C" int agpn_predict(agpn_ctx* c, const double* theta, int series, int M, const double* tstar,
                            double* mean_out, double* std_out, int* status, void* stream) {
    return predict_impl(c, theta, series, M, tstar, false, M, 0, 0.0, 0.0, mean_out, std_out, status, stream);
}

extern "C" int agpn_predict_slice(agpn_ctx* c, const double* theta, int series, int M, const double* tstar,
                                  int grid_M, int grid_offset, double grid_mean, double grid_t0, double* mean_out,
                                  double* std_out, int* status, void* stream) {
    if (grid_offset < 0 || grid_M < M || grid_offset > grid_M - M) USAGE_FAIL("agpn_predict_slice: slice outside the grid");
    return predict_impl(c, theta, series, M, tstar, true, grid_M, grid_offset, grid_mean, grid_t0, mean_out, std_out, status,
                        stream);
}

// Prediction on the factorisation kernels.  V = K* L^-T is what the blocked Cholesky computes for block rows appended
// below K: [K ; K*] (T x n_pad, no columns right of the training block) is eliminated left-looking over the training
// block columns only; the K* rows take the same INT8-sliced updates (row scale sqrt(k**_ii) >= |V_i|) and the same
// panel solves, the folded forward solve leaves -K* K^-1 r in the appended part of the right-hand side, and one pass
// over V gives the row sums of squares (arttwo.py:371-395).  *wrapped: an entry exceeded its row scale (a covariance
// that is not positive semi-definite jointly, e.g. the WhiteNoise square-lag quirk) -> the caller uses predict_kernel.
static int predict_stacked(agpn_ctx* c, int series, int M, size_t Mpad, int grid_M, int grid_off, const double* d_ts,
                           const double* d_ms, double* d_kss, double* d_mu, double* d_sd, int* d_kflag, cudaStream_t st,
                           int* wrapped, int* unavailable) {
    const size_t np = (size_t)c->npad;
    const size_t T = np + Mpad;
    const int nrt = (int)(T / TB);
    const int sgroup = c->nblk < 2 ? 2 : c->nblk;
    const size_t s8_bytes = c->nblk > 1 ? i8_slices_per_matrix(nrt, sgroup) : 0;
    *wrapped = 0;
    *unavailable = 0;
    if (T * np > c->cap_paug) {
        cudaFree(c->d_paug);
        c->d_paug = nullptr;
        c->cap_paug = 0;
        if (cudaMalloc(&c->d_paug, sizeof(double) * T * np) != cudaSuccess) { cudaGetLastError(); *unavailable = 1; return 0; }
        c->cap_paug = T * np;
    }
    if (s8_bytes > c->cap_ps8) {
        cudaFree(c->d_ps8);
        c->d_ps8 = nullptr;
        c->cap_ps8 = 0;
        if (cudaMalloc(&c->d_ps8, s8_bytes) != cudaSuccess) { cudaGetLastError(); *unavailable = 1; return 0; }
        c->cap_ps8 = s8_bytes;
    }
    if (3 * T + 2 > c->cap_pvec) {
        cudaFree(c->d_pvec);
        c->d_pvec = nullptr;
        c->cap_pvec = 0;
        if (cudaMalloc(&c->d_pvec, sizeof(double) * (3 * T + 2)) != cudaSuccess) { cudaGetLastError(); *unavailable = 1; return 0; }
        c->cap_pvec = 3 * T + 2;
    }
    // zero-chunk masks of the stacked factor's row tiles (+ 2 counters) and the inverse norms of the current diagonal block
    const int nzwords = (4 * sgroup + 63) / 64;
    const size_t mask_words = (size_t)nrt * nzwords + 2 + 2;    // masks | counters | 2 doubles
    if (c->i8_skip && c->nblk > 1 && mask_words > c->cap_pmask) {
        cudaFree(c->d_pmask);
        c->d_pmask = nullptr;
        c->cap_pmask = 0;
        if (cudaMalloc(&c->d_pmask, sizeof(unsigned long long) * mask_words) != cudaSuccess) { cudaGetLastError(); *unavailable = 1; return 0; }
        c->cap_pmask = mask_words;
    }
    double* d_rhs = c->d_pvec;                  // [T]: residual | 0  ->  z | -K* K^-1 r
    double* d_rs = c->d_pvec + T;               // [2][T] row scales
    int* d_flag = reinterpret_cast<int*>(c->d_pvec + 3 * T);
    CUDA_TRY(cudaMemsetAsync(d_rhs, 0, sizeof(double) * T, st));
    CUDA_TRY(cudaMemcpyAsync(d_rhs, c->d_rhs, sizeof(double) * np, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    // K + sigma^2 + s^2 (arttwo.py:367-370) on top, K* (arttwo.py:373-384) below
    AsmArgs a;
    a.theta = c->d_theta; a.ta = c->d_t; a.tb = c->d_t; a.na = c->N; a.nb = c->N; a.rows_out = c->npad; a.cols_out = c->npad;
    a.yerr = c->d_yerr; a.out = c->d_paug; a.mat_stride = (long long)(T * np); a.ld = c->npad; a.series0 = series; a.nser = 1;
    a.lower_only = 1; a.square = 1; a.add_err = 1; a.add_jit = 1; a.pad_identity = 1; a.use_trig = 1; a.vec2 = 1;
    a.torigin = c->t0; a.kflag = d_kflag; a.exact = c->exact ? 1 : 0;
    if (launch_assemble(c, a, 1, st)) return 1;
    a.ta = d_ts; a.tb = c->d_t; a.na = M; a.nb = c->N; a.rows_out = (int)Mpad; a.cols_out = c->npad;
    // WhiteNoise's square-lag quirk (node.py:68-72) looks at the WHOLE prediction grid: wn^2 where the grid has N points
    // and (index in the grid) == (training index), whatever slice of it this call evaluates
    a.out = c->d_paug + np * np; a.lower_only = 0; a.square = (grid_M == c->N) ? 1 : 0; a.row_off = grid_off; a.add_err = 0; a.add_jit = 0;
    a.pad_identity = 0; a.use_trig = 0; a.kflag = nullptr;
    if (launch_assemble(c, a, 1, st)) return 1;
    kss_kernel<<<(M + 127) / 128, 128, 0, st>>>(c->S, c->d_theta, series, d_ts, M, d_kss);
    LAUNCH_CHECK();
    CholArgs g;
    g.A = c->d_paug; g.mat_stride = (long long)(T * np); g.ld = c->npad; g.nblk = nrt; g.nmat = 1;
    g.Linv = c->d_Linv; g.linv_blocks = 1; g.rhs = d_rhs; g.logdet = c->d_logdet; g.quad = c->d_quad; g.info = c->d_info;
    g.A0 = c->d_paug; g.mat0 = 0;
    g.refine = c->d_info + 3 * c->cap_mats; g.refine_ratio = c->refine_ratio;
    g.keep_factor = 1;                          // V is read back for the row sums of squares
    if (c->nblk > 1) {
        g.S8 = c->d_ps8; g.rscale = d_rs; g.sgroup = sgroup; g.i8flag = d_flag;
        if (c->i8_skip && c->d_pmask && c->fuse_slices) {
            // K* is banded too (a grid point far from every training time adds nothing): the same masks, the same bound test
            g.nzmask = c->d_pmask; g.nzwords = nzwords;
            g.nzstat = c->d_pmask + (size_t)nrt * nzwords;
            if (c->trsm_skip) g.linvnorm = reinterpret_cast<double*>(c->d_pmask + (size_t)nrt * nzwords + 2);
            CUDA_TRY(cudaMemsetAsync(c->d_pmask, 0, sizeof(unsigned long long) * mask_words, st));
        }
        i8_rowscale_kernel<<<dim3((unsigned)((T + 255) / 256), 1), 256, 0, st>>>(g, c->npad, d_kss, c->predict_test_wrap ? 0 : M);
        LAUNCH_CHECK();
    }
    const bool maps_were = c->maps_valid;
    c->maps_valid = false;                      // the tensor maps describe the batch workspace, not this matrix
    c->group_eff = c->nblk;                     // one group: purely left-looking over the training block columns
    if (!c->i8_clusters_fixed) c->i8_clusters = c->i8_clusters_max;      // one stream: every SM
    const int rc = run_cholesky(c, g, st, nullptr, c->nblk, true);
    c->maps_valid = maps_were;
    if (rc) return 1;
    predict_finish_kernel<<<(M + 7) / 8, 256, 0, st>>>(c->d_paug + np * np, c->npad, c->npad, M, d_rhs + np, d_ms, d_kss, c->d_theta,
                                                       c->S.jit_off + series, d_mu, d_sd);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(wrapped, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

static int predict_impl(agpn_ctx* c, const double* theta, int series, int M, const double* tstar, bool have_grid,
                        int grid_M, int grid_off, double grid_mean, double grid_t0, double* mean_out, double* std_out,
                        int* status, void* stream) {
    if (!c || !theta || !tstar || !mean_out || !std_out) USAGE_FAIL("agpn_predict: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_predict: call agpn_set_structure first");
    if (series < 0 || series >= c->p) USAGE_FAIL("agpn_predict: series out of range");
    if (M < 1) USAGE_FAIL("agpn_predict: empty prediction grid");
    ENTER_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t np = (size_t)c->npad;
    if (ensure_batch_workspace(c, 1, 1)) return 1;
    const size_t Mpad = ((size_t)M + TB - 1) / TB * TB;
    if ((size_t)M > c->cap_M) {
        cudaFree(c->d_pred);
        c->d_pred = nullptr;
        c->cap_M = 0;
        CUDA_TRY(cudaMalloc(&c->d_pred, sizeof(double) * 5 * (size_t)M));
        c->cap_M = M;
    }
    double* d_ts = c->d_pred;
    double* d_ms = c->d_pred + c->cap_M;
    double* d_kss = c->d_pred + 2 * c->cap_M;
    double* d_mu = c->d_pred + 3 * c->cap_M;
    double* d_sd = c->d_pred + 4 * c->cap_M;
    CUDA_TRY(cudaMemcpyAsync(c->d_theta, theta, sizeof(double) * c->S.D, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ts, tstar, sizeof(double) * M, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(c->d_logdet, 0, sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(c->d_quad, 0, sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(c->d_info, 0, sizeof(int) * 4 * c->cap_mats, st));
    int* d_kflag = c->d_info + c->cap_mats;
    int* d_rflag = c->d_info + 2 * c->cap_mats;
    // residual of the chosen series (arttwo.py:359-365)
    MeanArgs m;
    m.theta = c->d_theta; m.t = c->d_t; m.n = c->N; m.n_pad = c->npad; m.tmean = c->tmean; m.t0 = c->t0; m.y = c->d_y;
    m.out = c->d_rhs; m.series0 = series; m.nser = 1; m.subtract = 1; m.rflag = d_rflag;
    mean_kernel<<<dim3(1, 1), 256, 0, st>>>(c->S, m);
    LAUNCH_CHECK();
    // mean on the prediction grid (arttwo.py:388): centred on / anchored at the prediction grid itself
    m.t = d_ts; m.n = M; m.n_pad = M; m.tmean = have_grid ? grid_mean : host_mean(tstar, M);
    m.t0 = have_grid ? grid_t0 : tstar[0]; m.y = nullptr; m.out = d_ms;
    m.subtract = 0; m.rflag = nullptr;
    mean_kernel<<<dim3(1, 1), 256, 0, st>>>(c->S, m);
    LAUNCH_CHECK();
    // default: K* rides along as extra block rows of the factorisation (INT8-sliced left-looking updates)
    bool done_stacked = false;
    if (c->i8_which == 3 && !c->predict_dmma && c->nblk <= 256) {
        int wrapped = 0, unavailable = 0;
        if (predict_stacked(c, series, M, Mpad, grid_M, grid_off, d_ts, d_ms, d_kss, d_mu, d_sd, d_kflag, st, &wrapped, &unavailable)) return 1;
        done_stacked = !wrapped && !unavailable;
        if (!done_stacked) {
            CUDA_TRY(cudaMemsetAsync(c->d_logdet, 0, sizeof(double), st));
            CUDA_TRY(cudaMemsetAsync(c->d_quad, 0, sizeof(double), st));
            CUDA_TRY(cudaMemsetAsync(c->d_info, 0, sizeof(int) * 2 * c->cap_mats, st));      // info | kflag (rflag is the residual's)
        }
    }
    if (!done_stacked) {
    if ((size_t)c->nblk > c->cap_linv_blocks) {
        cudaFree(c->d_LinvAll);
        c->d_LinvAll = nullptr;
        c->cap_linv_blocks = 0;
        CUDA_TRY(cudaMalloc(&c->d_LinvAll, sizeof(double) * (size_t)c->nblk * TB * TB));
        CUDA_TRY(cudaMemset(c->d_LinvAll, 0, sizeof(double) * (size_t)c->nblk * TB * TB));
        c->cap_linv_blocks = c->nblk;
    }
    if (Mpad * np > c->cap_V) {
        cudaFree(c->d_V);
        c->d_V = nullptr;
        c->cap_V = 0;
        CUDA_TRY(cudaMalloc(&c->d_V, sizeof(double) * Mpad * np));
        c->cap_V = Mpad * np;
    }
    // K + sigma^2 + s^2 (arttwo.py:367-370), factor (arttwo.py:371)
    AsmArgs a;
    a.theta = c->d_theta; a.ta = c->d_t; a.tb = c->d_t; a.na = c->N; a.nb = c->N; a.rows_out = c->npad; a.cols_out = c->npad;
    a.yerr = c->d_yerr; a.out = c->d_A; a.mat_stride = (long long)(np * np); a.ld = c->npad; a.series0 = series; a.nser = 1;
    a.lower_only = 1; a.square = 1; a.add_err = 1; a.add_jit = 1; a.pad_identity = 1; a.use_trig = 1; a.vec2 = 1;
    a.torigin = c->t0; a.kflag = d_kflag; a.exact = c->exact ? 1 : 0;
    if (launch_assemble(c, a, 1, st)) return 1;
    CholArgs g;
    g.A = c->d_A; g.mat_stride = (long long)(np * np); g.ld = c->npad; g.nblk = c->nblk; g.nmat = 1;
    g.Linv = c->d_LinvAll; g.linv_blocks = c->nblk; g.rhs = c->d_rhs; g.logdet = c->d_logdet; g.quad = c->d_quad;
    g.info = c->d_info;
    g.A0 = c->d_A; g.mat0 = 0;
    c->group_eff = c->group_fixed ? c->group : 2;
    if (run_cholesky(c, g, st, nullptr)) return 1;
    kss_kernel<<<(M + 127) / 128, 128, 0, st>>>(c->S, c->d_theta, series, d_ts, M, d_kss);
    LAUNCH_CHECK();
    PredArgs pa;
    pa.theta = c->d_theta; pa.tstar = d_ts; pa.ttrain = c->d_t; pa.M = M; pa.N = c->N; pa.n_pad = c->npad; pa.nblk = c->nblk;
    pa.series = series; pa.square = (grid_M == c->N) ? 1 : 0; pa.row_off = grid_off; pa.exact = c->exact ? 1 : 0; pa.L = c->d_A; pa.ld = c->npad; pa.Linv = c->d_LinvAll;
    pa.z = c->d_rhs; pa.V = c->d_V; pa.mstar = d_ms; pa.kss = d_kss; pa.mean_out = d_mu; pa.std_out = d_sd;
    predict_kernel<<<(unsigned)(Mpad / TB), 256, GEMM_SMEM_BYTES, st>>>(c->S, pa);
    LAUNCH_CHECK();
    }   // !done_stacked
    c->last_predict_path = done_stacked ? 1 : 0;
    CUDA_TRY(cudaMemcpyAsync(mean_out, d_mu, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(std_out, d_sd, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    int h_flags[3] = {0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(&h_flags[0], c->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&h_flags[1], d_kflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&h_flags[2], d_rflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (status) *status = h_flags[1] ? AGPN_STATUS_NOT_FINITE : (h_flags[0] ? AGPN_STATUS_NOT_PD : (h_flags[2] ? AGPN_STATUS_NOT_FINITE : AGPN_STATUS_OK));
    return 0;
}

// mean[m] = mstar[m] - rhs_aug[Npad + m] ; cov = symmetrised trailing block of the augmented matrix
__global__ void extract_cov_kernel(const double* aug, int T, int npad, int M, const double* rhs_aug, const double* mstar,
                                   double* mean_out, double* cov_out) {
    const size_t n = (size_t)M * M;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / M), j = (int)(idx % M);
        const int r = i >= j ? i : j, cc = i >= j ? j : i;
        cov_out[idx] = aug[(size_t)(npad + r) * T + npad + cc];
        if (j == 0) mean_out[i] = mstar[i] - rhs_aug[npad + i];
    }
}

extern "C" int agpn_predict_cov(agpn_ctx* c, const double* theta, int series, int M, const double* tstar,
                                double train_jitter, double* mean_out, double* cov_out, int* status, void* stream) {
    if (!c || !theta || !tstar || !mean_out || !cov_out) USAGE_FAIL("agpn_predict_cov: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_predict_cov: call agpn_set_structure first");
    if (series < 0 || series >= c->p) USAGE_FAIL("agpn_predict_cov: series out of range");
    if (M < 1) USAGE_FAIL("agpn_predict_cov: empty prediction grid");
    ENTER_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = c->S.D;
    const size_t npad = (size_t)c->npad;
    const size_t Mpad = ((size_t)M + TB - 1) / TB * TB;
    const size_t T = npad + Mpad;
    if (T > 46000) USAGE_FAIL("agpn_predict_cov: N + M too large for a dense covariance (use agpn_predict)");
    if (ensure_batch_workspace(c, 1, 1)) return 1;
    if (T * T > c->cap_aug) {
        cudaFree(c->d_aug);
        c->d_aug = nullptr;
        c->cap_aug = 0;
        CUDA_TRY(cudaMalloc(&c->d_aug, sizeof(double) * T * T));
        c->cap_aug = T * T;
    }
    if (T + 2 * (size_t)D > c->cap_augv) {
        cudaFree(c->d_augv);
        c->d_augv = nullptr;
        c->cap_augv = 0;
        CUDA_TRY(cudaMalloc(&c->d_augv, sizeof(double) * (T + 2 * (size_t)D)));
        c->cap_augv = T + 2 * (size_t)D;
    }
    if ((size_t)M > c->cap_M) {
        cudaFree(c->d_pred);
        c->d_pred = nullptr;
        c->cap_M = 0;
        CUDA_TRY(cudaMalloc(&c->d_pred, sizeof(double) * 5 * (size_t)M));
        c->cap_M = M;
    }
    double* d_rhs = c->d_augv;                 // [T]
    double* d_th = c->d_augv + T;              // theta (prediction blocks) | theta with the training jitter
    double* d_ts = c->d_pred;
    double* d_ms = c->d_pred + c->cap_M;
    double* d_mu = c->d_pred + 3 * c->cap_M;
    std::vector<double> th2(theta, theta + D);
    if (train_jitter == train_jitter) th2[c->S.jit_off + series] = train_jitter;    // not NaN: legacy art.network quirk
    CUDA_TRY(cudaMemcpyAsync(d_th, theta, sizeof(double) * D, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_th + D, th2.data(), sizeof(double) * D, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ts, tstar, sizeof(double) * M, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));       // th2 is a local
    // The Schur update K** - (K* L^-T)(K* L^-T)^T runs on the INT8 tensor path like the likelihood's trailing updates (one group
    // of the training block columns, then ONE wide update of the trailing block); if an entry of K* L^-T exceeds its row scale
    // sqrt(K**_ii + s^2) -- a covariance that is not jointly positive semi-definite -- everything is redone on DMMA
    const int nrt = (int)(T / TB);
    const bool try_i8 = c->i8_which == 3 && !c->predict_dmma && c->nblk > 2 && c->nblk <= 256;
    const size_t s8_bytes = try_i8 ? i8_slices_per_matrix(nrt, c->nblk) : 0;
    if (try_i8) {
        if (s8_bytes > c->cap_ps8) {
            cudaFree(c->d_ps8);
            c->d_ps8 = nullptr;
            c->cap_ps8 = 0;
            CUDA_TRY(cudaMalloc(&c->d_ps8, s8_bytes));
            c->cap_ps8 = s8_bytes;
        }
        if (3 * T + 2 > c->cap_pvec) {
            cudaFree(c->d_pvec);
            c->d_pvec = nullptr;
            c->cap_pvec = 0;
            CUDA_TRY(cudaMalloc(&c->d_pvec, sizeof(double) * (3 * T + 2)));
            c->cap_pvec = 3 * T + 2;
        }
    }
    int* d_kflag = c->d_info + c->cap_mats;
    int* d_rflag = c->d_info + 2 * c->cap_mats;
    for (int attempt = try_i8 ? 0 : 1; attempt < 2; ++attempt) {
        CUDA_TRY(cudaMemsetAsync(d_rhs, 0, sizeof(double) * T, st));
        CUDA_TRY(cudaMemsetAsync(c->d_logdet, 0, sizeof(double), st));
        CUDA_TRY(cudaMemsetAsync(c->d_quad, 0, sizeof(double), st));
        CUDA_TRY(cudaMemsetAsync(c->d_info, 0, sizeof(int) * 4 * c->cap_mats, st));
        MeanArgs m;
        m.theta = d_th; m.t = c->d_t; m.n = c->N; m.n_pad = c->npad; m.tmean = c->tmean; m.t0 = c->t0; m.y = c->d_y;
        m.out = d_rhs; m.series0 = series; m.nser = 1; m.subtract = 1; m.rflag = d_rflag;
        mean_kernel<<<dim3(1, 1), 256, 0, st>>>(c->S, m);
        LAUNCH_CHECK();
        m.t = d_ts; m.n = M; m.n_pad = M; m.tmean = host_mean(tstar, M); m.t0 = tstar[0]; m.y = nullptr; m.out = d_ms;
        m.subtract = 0; m.rflag = nullptr;
        mean_kernel<<<dim3(1, 1), 256, 0, st>>>(c->S, m);
        LAUNCH_CHECK();
        // augmented matrix [K + sigma^2 + s^2 , . ; K* , K** + s^2] (arttwo.py:367-387), lower block triangle
        AsmArgs a;
        a.theta = d_th + D; a.ta = c->d_t; a.tb = c->d_t; a.na = c->N; a.nb = c->N; a.rows_out = c->npad; a.cols_out = c->npad;
        a.yerr = c->d_yerr; a.out = c->d_aug; a.mat_stride = (long long)(T * T); a.ld = (int)T; a.series0 = series; a.nser = 1;
        a.lower_only = 1; a.square = 1; a.add_err = 1; a.add_jit = 1; a.pad_identity = 1; a.use_trig = 1; a.vec2 = 1;
        a.torigin = c->t0; a.kflag = d_kflag; a.exact = c->exact ? 1 : 0;
        if (launch_assemble(c, a, 1, st)) return 1;
        a.theta = d_th; a.ta = d_ts; a.tb = c->d_t; a.na = M; a.nb = c->N; a.rows_out = (int)Mpad; a.cols_out = c->npad;
        a.out = c->d_aug + npad * T; a.lower_only = 0; a.square = (M == c->N) ? 1 : 0; a.add_err = 0; a.add_jit = 0;
        a.pad_identity = 0; a.use_trig = 0; a.kflag = nullptr;
        if (launch_assemble(c, a, 1, st)) return 1;
        a.ta = d_ts; a.tb = d_ts; a.na = M; a.nb = M; a.rows_out = (int)Mpad; a.cols_out = (int)Mpad;
        a.out = c->d_aug + npad * T + npad; a.lower_only = 1; a.square = 1; a.add_err = 0; a.add_jit = 1; a.pad_identity = 1;
        if (launch_assemble(c, a, 1, st)) return 1;
        // eliminate the training block columns only: the trailing block becomes K** + s^2 - K* K^-1 K*^T and
        // the trailing part of the right-hand side becomes -K* K^-1 r
        CholArgs g;
        g.A = c->d_aug; g.mat_stride = (long long)(T * T); g.ld = (int)T; g.nblk = (int)(T / TB); g.nmat = 1;
        g.Linv = c->d_Linv; g.linv_blocks = 1; g.rhs = d_rhs; g.logdet = c->d_logdet; g.quad = c->d_quad; g.info = c->d_info;
        g.A0 = c->d_aug; g.mat0 = 0;
        g.refine = c->d_info + 3 * c->cap_mats; g.refine_ratio = c->refine_ratio;
        const bool maps_were = c->maps_valid;
        c->maps_valid = false;                      // the tensor maps describe the batch workspace, not this matrix
        int* d_wrap = nullptr;
        if (attempt == 0) {
            double* d_rs = c->d_pvec;                              // [2][T] row scales | wrap flag
            d_wrap = reinterpret_cast<int*>(c->d_pvec + 2 * T);
            CUDA_TRY(cudaMemsetAsync(d_wrap, 0, sizeof(int), st));
            g.S8 = c->d_ps8; g.rscale = d_rs; g.sgroup = c->nblk; g.i8flag = d_wrap;
            i8_rowscale_kernel<<<dim3((unsigned)((T + 255) / 256), 1), 256, 0, st>>>(g, (int)T, nullptr, 0);
            LAUNCH_CHECK();
            c->group_eff = c->nblk;
            if (!c->i8_clusters_fixed) c->i8_clusters = c->i8_clusters_max;
        } else {
            c->group_eff = c->group_fixed ? c->group : 2;
        }
        const int rc = run_cholesky(c, g, st, nullptr, c->nblk);
        c->maps_valid = maps_were;
        if (rc) return 1;
        c->last_predict_path = attempt == 0 ? 1 : 0;
        if (attempt == 0) {
            int wrapped = 0;
            CUDA_TRY(cudaMemcpyAsync(&wrapped, d_wrap, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (!wrapped && !c->predict_test_wrap) break;
        }
    }
    double* d_cov = nullptr;
    CUDA_TRY(cudaMalloc(&d_cov, sizeof(double) * (size_t)M * M));
    extract_cov_kernel<<<1024, 256, 0, st>>>(c->d_aug, (int)T, c->npad, M, d_rhs, d_ms, d_mu, d_cov);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(mean_out, d_mu, sizeof(double) * M, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cov_out, d_cov, sizeof(double) * (size_t)M * M, cudaMemcpyDeviceToHost, st);
    int h_flags[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flags[0], c->d_info, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flags[1], d_kflag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flags[2], d_rflag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_cov);
    CUDA_TRY(e);
    if (status) *status = h_flags[1] ? AGPN_STATUS_NOT_FINITE : (h_flags[0] ? AGPN_STATUS_NOT_PD : (h_flags[2] ? AGPN_STATUS_NOT_FINITE : AGPN_STATUS_OK));
    return 0;
}

// log-likelihood and its analytic gradient for one theta row (agpn_grad.cuh)
extern "C" int agpn_loglike_grad(agpn_ctx* c, const double* theta, double* ll_out, double* grad_out, int* status, void* stream) {
    if (!c || !theta || !ll_out || !grad_out) USAGE_FAIL("agpn_loglike_grad: null pointer");
    if (!c->has_structure) USAGE_FAIL("agpn_loglike_grad: call agpn_set_structure first");
    const Structure& S = c->S;
    if (S.D > GRAD_MAXD) USAGE_FAIL("agpn_loglike_grad: more than 96 parameters per row");
    for (int j = 0; j < S.q; ++j)
        if (!grad_shape_supported(S.node_kind[j])) USAGE_FAIL("agpn_loglike_grad: node kernel class without a gradient (supported: Constant, SquaredExponential, Periodic, QuasiPeriodic, Cosine, Exponential, Matern32, Matern52)");
    for (int i = 0; i < S.p * S.q; ++i)
        if (!grad_shape_supported(S.weight_kind[i])) USAGE_FAIL("agpn_loglike_grad: weight kernel class without a gradient (supported: Constant, SquaredExponential, Periodic, QuasiPeriodic, Cosine, Exponential, Matern32, Matern52)");
    if (S.use_means)
        for (int t = 0; t < S.mean_first[S.p]; ++t)
            if (!grad_mean_supported(S.mean_kind[t])) USAGE_FAIL("agpn_loglike_grad: mean class without a gradient (supported: Constant, Linear)");
    ENTER_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = S.D, p = c->p;
    const size_t npad = (size_t)c->npad, T = 2 * npad;
    if (ensure_batch_workspace(c, 1, 1)) return 1;
    if (T * T > c->cap_aug) {
        cudaFree(c->d_aug);
        c->d_aug = nullptr;
        c->cap_aug = 0;
        CUDA_TRY(cudaMalloc(&c->d_aug, sizeof(double) * T * T));
        c->cap_aug = T * T;
    }
    const int nt64 = (c->N + 63) / 64, ntiles = nt64 * (nt64 + 1) / 2;
    DevBuf b_rhs, b_th, b_grad, b_part, b_res;
    CUDA_TRY(b_rhs.alloc(T));
    CUDA_TRY(b_th.alloc(D));
    CUDA_TRY(b_grad.alloc(D));
    CUDA_TRY(b_part.alloc((size_t)ntiles * GRAD_MAXD));
    CUDA_TRY(b_res.alloc((size_t)p * npad));
    CUDA_TRY(cudaMemcpyAsync(b_th.p, theta, sizeof(double) * D, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(b_grad.p, 0, sizeof(double) * D, st));
    CUDA_TRY(cudaMemsetAsync(c->d_logdet, 0, sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(c->d_quad, 0, sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(c->d_info, 0, sizeof(int) * 4 * c->cap_mats, st));
    int* d_kflag = c->d_info + c->cap_mats;
    // residuals of all series (arttwo.py:262-264)
    MeanArgs m;
    m.theta = b_th.p; m.t = c->d_t; m.n = c->N; m.n_pad = c->npad; m.tmean = c->tmean; m.t0 = c->t0; m.y = c->d_y;
    m.out = b_res.p; m.series0 = 0; m.nser = p; m.subtract = 1; m.rflag = nullptr;
    mean_kernel<<<dim3(p, 1), 256, 0, st>>>(S, m);
    LAUNCH_CHECK();
    const bool maps_were = c->maps_valid;
    int st_code = AGPN_STATUS_OK;
    for (int i = 0; i < p && st_code == AGPN_STATUS_OK; ++i) {
        // [[K_i + sigma^2 + s^2, .], [I, 0]] with right-hand side [r_i; 0]
        AsmArgs a;
        a.theta = b_th.p; a.ta = c->d_t; a.tb = c->d_t; a.na = c->N; a.nb = c->N; a.rows_out = c->npad; a.cols_out = c->npad;
        a.yerr = c->d_yerr; a.out = c->d_aug; a.mat_stride = (long long)(T * T); a.ld = (int)T; a.series0 = i; a.nser = 1;
        a.lower_only = 1; a.square = 1; a.add_err = 1; a.add_jit = 1; a.pad_identity = 1; a.use_trig = 1; a.vec2 = 1;
        a.torigin = c->t0; a.kflag = d_kflag; a.exact = c->exact ? 1 : 0;
        if (launch_assemble(c, a, 1, st)) return 1;
        grad_aug_fill_kernel<<<1024, 256, 0, st>>>(c->d_aug, (int)T, c->npad);
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemsetAsync(b_rhs.p, 0, sizeof(double) * T, st));
        CUDA_TRY(cudaMemcpyAsync(b_rhs.p, b_res.p + (size_t)i * npad, sizeof(double) * npad, cudaMemcpyDeviceToDevice, st));
        CholArgs g;
        g.A = c->d_aug; g.mat_stride = (long long)(T * T); g.ld = (int)T; g.nblk = (int)(T / TB); g.nmat = 1;
        g.Linv = c->d_Linv; g.linv_blocks = 1; g.rhs = b_rhs.p; g.logdet = c->d_logdet; g.quad = c->d_quad; g.info = c->d_info;
        g.A0 = c->d_aug; g.mat0 = 0;
        g.refine = c->d_info + 3 * c->cap_mats; g.refine_ratio = c->refine_ratio;
        c->maps_valid = false;
        c->group_eff = c->group_fixed ? c->group : 2;
        const int rc = run_cholesky(c, g, st, nullptr, c->nblk);
        c->maps_valid = maps_were;
        if (rc) return 1;
        GradArgs ga;
        ga.theta = b_th.p; ga.t = c->d_t; ga.N = c->N; ga.series = i; ga.Kinv_neg = c->d_aug + npad * T + npad; ga.ld = (int)T;
        ga.alpha_neg = b_rhs.p + npad; ga.tmean = c->tmean; ga.partial = b_part.p;
        grad_tile_kernel<<<ntiles, 256, 0, st>>>(S, ga);
        LAUNCH_CHECK();
        grad_reduce_kernel<<<(D + 63) / 64, 64, 0, st>>>(b_part.p, ntiles, D, b_grad.p);
        LAUNCH_CHECK();
        int h_flags[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(&h_flags[0], c->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&h_flags[1], d_kflag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_flags[1]) st_code = AGPN_STATUS_NOT_FINITE;
        else if (h_flags[0]) st_code = AGPN_STATUS_NOT_PD;
    }
    double h_ld = 0.0, h_q = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&h_ld, c->d_logdet, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&h_q, c->d_quad, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(grad_out, b_grad.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (st_code == AGPN_STATUS_OK && !(std::isfinite(h_q) && std::isfinite(h_ld))) st_code = AGPN_STATUS_NOT_FINITE;
    *ll_out = st_code == AGPN_STATUS_OK ? -0.5 * h_q - h_ld - 0.5 * (double)c->N * p * std::log(2.0 * AGPN_PI)
                                        : (st_code == AGPN_STATUS_NOT_PD ? -INFINITY : NAN);
    if (status) *status = st_code;
    return 0;
}

extern "C" int agpn_kernel_eval(int device, int family, int kind, const double* pars, const double* r,
                                const double* t1, const double* t2, int rows, int cols, int square_quirk, double* out) {
    if (!pars || !out) USAGE_FAIL("agpn_kernel_eval: null pointer");
    if (kernel_npars(family, kind) < 0) USAGE_FAIL("agpn_kernel_eval: unknown kernel kind");
    if (rows < 1 || cols < 1) USAGE_FAIL("agpn_kernel_eval: empty array");
    const bool nonstat = (kind == K_LINEAR || kind == K_POLY);
    if (nonstat && (!t1 || !t2)) USAGE_FAIL("agpn_kernel_eval: Linear / Polynomial need t1 and t2");
    if (!nonstat && !r) USAGE_FAIL("agpn_kernel_eval: stationary kernels need r");
    ENTER_DEVICE(device);
    const size_t n = (size_t)rows * cols;
    DevBuf b_r, b_t1, b_t2, b_o;
    CUDA_TRY(b_o.alloc(n));
    if (r) {
        CUDA_TRY(b_r.alloc(n));
        CUDA_TRY(cudaMemcpy(b_r.p, r, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    if (t1) {
        CUDA_TRY(b_t1.alloc(rows));
        CUDA_TRY(cudaMemcpy(b_t1.p, t1, sizeof(double) * rows, cudaMemcpyHostToDevice));
    }
    if (t2) {
        CUDA_TRY(b_t2.alloc(cols));
        CUDA_TRY(cudaMemcpy(b_t2.p, t2, sizeof(double) * cols, cudaMemcpyHostToDevice));
    }
    const PKern k = prep_kernel(family, kind, pars);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 4096);
    kernel_eval_kernel<<<blocks, 256>>>(k, b_r.p, b_t1.p, b_t2.p, rows, cols, square_quirk, b_o.p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpy(out, b_o.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int agpn_potrf_batched(int device, int nmat, int n_pad, double* A, double* rhs, double* logdet,
                                  double* quad, int* info, void* stream) {
    if (!A || !logdet || !info) USAGE_FAIL("agpn_potrf_batched: null pointer");
    if (rhs && !quad) USAGE_FAIL("agpn_potrf_batched: rhs needs quad");
    if (n_pad < TB || n_pad % TB) USAGE_FAIL("agpn_potrf_batched: n_pad must be a positive multiple of 128");
    if (nmat < 1 || nmat > 65535) USAGE_FAIL("agpn_potrf_batched: 1 <= nmat <= 65535");
    ENTER_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    agpn_ctx tmp;
    tmp.device = device;
    if (const char* e = getenv("AGPN_GROUP")) { tmp.group = atoi(e) > 0 ? atoi(e) : 8; tmp.group_fixed = true; }
    if (set_kernel_attrs(&tmp)) return 1;
    DevBuf b_linv, b_rs;
    struct ByteBuf { signed char* p = nullptr; ~ByteBuf() { if (p) cudaFree(p); } } b_s8;
    struct IntBuf { int* p = nullptr; ~IntBuf() { if (p) cudaFree(p); } } b_refine;
    CUDA_TRY(cudaMalloc(&b_refine.p, sizeof(int) * nmat));
    CUDA_TRY(cudaMemsetAsync(b_refine.p, 0, sizeof(int) * nmat, st));
    CUDA_TRY(b_linv.alloc((size_t)nmat * TB * TB));
    double* d_linv = b_linv.p;
    CUDA_TRY(cudaMemsetAsync(d_linv, 0, sizeof(double) * (size_t)nmat * TB * TB, st));     // zeros above the block diagonal of the inverse
    CUDA_TRY(cudaMemsetAsync(logdet, 0, sizeof(double) * nmat, st));
    if (quad) CUDA_TRY(cudaMemsetAsync(quad, 0, sizeof(double) * nmat, st));
    CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(int) * nmat, st));
    CholArgs g;
    g.A = A; g.mat_stride = (long long)n_pad * n_pad; g.ld = n_pad; g.nblk = n_pad / TB; g.nmat = nmat; g.Linv = d_linv;
    g.linv_blocks = 1; g.rhs = rhs; g.logdet = logdet; g.quad = quad; g.info = info;
    g.refine = b_refine.p;
    if (const char* e = getenv("AGPN_REFINE_RATIO")) g.refine_ratio = atof(e);
    g.A0 = A; g.mat0 = 0;
    if (const char* e = getenv("AGPN_SYRK")) { tmp.use_tma = strncmp(e, "tma", 3) == 0; tmp.tma_mc = strcmp(e, "tma_nomc") != 0; }
    tmp.maps_valid = tmp.use_tma && make_batch_map(&tmp.tmA, A, nmat, n_pad, n_pad, TB) && make_batch_map(&tmp.tmB, A, nmat, n_pad, n_pad, BN);
    tmp.group_eff = (tmp.group_fixed || (long long)nmat * (n_pad / TB) >= 384) ? tmp.group : 2;
    tmp.i8_which = i8_mode_from_env();
    if (const char* e = getenv("AGPN_I8_FUSE")) tmp.fuse_slices = atoi(e) != 0;
    if (const char* e = getenv("AGPN_I8_PERSIST")) tmp.i8_persist = atoi(e) != 0;
    g.keep_factor = 1;                                   // the caller reads L
    if (tmp.i8_which && g.nblk > 1) {
        int ig = tmp.group_fixed ? tmp.group : g.nblk;
        if (ig > 256) ig = 256;
        g.sgroup = ig < 2 ? 2 : ig;
        if (cudaMalloc(&b_s8.p, (size_t)nmat * i8_slices_per_matrix(g.nblk, g.sgroup)) != cudaSuccess ||
            b_rs.alloc((size_t)nmat * 2 * n_pad) != cudaSuccess) {
            cudaGetLastError();
            USAGE_FAIL("agpn_potrf_batched: not enough device memory for the digit planes");
        }
        g.S8 = b_s8.p; g.rscale = b_rs.p;
        tmp.group_eff = ig;
        i8_rowscale_kernel<<<dim3((n_pad + 255) / 256, nmat), 256, 0, st>>>(g, n_pad, nullptr, 0);
        LAUNCH_CHECK();
    }
    const int rc = run_cholesky(&tmp, g, st, nullptr);
    const cudaError_t e = cudaStreamSynchronize(st);       // the temporaries are freed (RAII) only after the work has drained
    if (rc) return rc;
    CUDA_TRY(e);
    return 0;
}

// ---- draws from N(0, K) (network.sample, arttwo.py:129-147) ---------------------------------------------------------
// A[npad][npad] = K + nugget I on [n][n], identity padding
__global__ void sample_pad_kernel(const double* __restrict__ K, int n, int npad, double nugget, double* __restrict__ A) {
    const size_t total = (size_t)npad * npad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / npad), j = (int)(idx % npad);
        double v = (i == j) ? 1.0 : 0.0;
        if (i < n && j < n) v = K[(size_t)i * n + j] + ((i == j) ? nugget : 0.0);
        A[idx] = v;
    }
}
// out[s][i] = sum_{j <= i} L[i][j] z[s][j]; grid (n, ceil(nsamp / 8)), one warp per (row, sample)
__global__ void __launch_bounds__(256) sample_trmv_kernel(const double* __restrict__ L, int ld, int n, const double* __restrict__ z,
                                                          int nsamp, double* __restrict__ out) {
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sidx = blockIdx.y * 8 + warp;
    if (sidx >= nsamp) return;
    const double* lr = L + (size_t)i * ld;
    const double* zr = z + (size_t)sidx * n;
    double acc = 0.0;
    for (int j = lane; j <= i; j += 32) acc = fma(lr[j], zr[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(size_t)sidx * n + i] = acc;
}

extern "C" int agpn_mvn_sample(int device, int n, const double* K, int nsamp, const double* z, double* out,
                               double* nugget_used) {
    if (!K || !z || !out) USAGE_FAIL("agpn_mvn_sample: null pointer");
    if (n < 1 || nsamp < 1) USAGE_FAIL("agpn_mvn_sample: need n >= 1 and nsamp >= 1");
    ENTER_DEVICE(device);
    const int npad = (n + TB - 1) / TB * TB;
    DevBuf b_K, b_A, b_z, b_o, b_ld;
    struct IntBuf { int* p = nullptr; ~IntBuf() { if (p) cudaFree(p); } } b_info;
    CUDA_TRY(b_K.alloc((size_t)n * n));
    CUDA_TRY(b_A.alloc((size_t)npad * npad));
    CUDA_TRY(b_z.alloc((size_t)n * nsamp));
    CUDA_TRY(b_o.alloc((size_t)n * nsamp));
    CUDA_TRY(b_ld.alloc(1));
    CUDA_TRY(cudaMalloc(&b_info.p, sizeof(int)));
    CUDA_TRY(cudaMemcpy(b_K.p, K, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(b_z.p, z, sizeof(double) * (size_t)n * nsamp, cudaMemcpyHostToDevice));
    double dmax = 0.0;
    for (int i = 0; i < n; ++i) dmax = std::max(dmax, std::fabs(K[(size_t)i * n + i]));
    if (!(dmax > 0.0) || !std::isfinite(dmax)) dmax = 1.0;
    // allow_singular semantics: a semi-definite K gets the smallest diagonal shift that makes it factorisable
    double nugget = 0.0;
    int info = 1;
    for (int attempt = 0; attempt < 6 && info; ++attempt) {      // nugget <= 1e-6 max|K_ii|: beyond that K is not PSD
        sample_pad_kernel<<<1024, 256>>>(b_K.p, n, npad, nugget, b_A.p);
        LAUNCH_CHECK();
        if (agpn_potrf_batched(device, 1, npad, b_A.p, nullptr, b_ld.p, nullptr, b_info.p, nullptr)) return 1;
        CUDA_TRY(cudaMemcpy(&info, b_info.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (info) nugget = (nugget == 0.0) ? 1e-14 * dmax : nugget * 100.0;
    }
    if (info) USAGE_FAIL("agpn_mvn_sample: covariance is not positive semi-definite");
    sample_trmv_kernel<<<dim3(n, (nsamp + 7) / 8), 256>>>(b_A.p, npad, n, b_z.p, nsamp, b_o.p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpy(out, b_o.p, sizeof(double) * (size_t)n * nsamp, cudaMemcpyDeviceToHost));
    if (nugget_used) *nugget_used = nugget;
    return 0;
}

// ---- INT8 tensor-path peak probe (roofline denominator of syrk_i8_kernel) --------------------
extern "C" int agpn_i8_peak(int device, double* tops) {
    if (!tops) USAGE_FAIL("agpn_i8_peak: null pointer");
    ENTER_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount, iters = 40000;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    i8_peak_kernel<<<grid, 128>>>(2000, nullptr);
    LAUNCH_CHECK();
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        i8_peak_kernel<<<grid, 128>>>(iters, nullptr);
        LAUNCH_CHECK();
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double t = 2.0 * 128.0 * 256.0 * 32.0 * (double)iters * grid / (ms * 1e-3) / 1e12;
        if (t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tops = best;
    return 0;
}

// ---- FP64 tensor-pipe peak probe (roofline denominator measured in the same run) -------------
template <int NACC>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int agpn_dmma_peak(int device, double* tflops) {
    if (!tflops) USAGE_FAIL("agpn_dmma_peak: null pointer");
    ENTER_DEVICE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 2, threads = 256, iters = 20000;
    constexpr int NACC = 8;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * grid * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    dmma_peak_kernel<NACC><<<grid, threads>>>(out, 1000);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        dmma_peak_kernel<NACC><<<grid, threads>>>(out, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 256.0 * NACC * (double)iters * (threads / 32) * grid;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return 0;
}
