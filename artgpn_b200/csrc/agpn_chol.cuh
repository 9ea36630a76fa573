// Batched blocked FP64 Cholesky for sm_100a: replaces scipy cho_factor / cho_solve of
// artgpn (arttwo.py:284-287, :371-372) for nmat independent matrices at once.
//
// Layout: each matrix is row-major [n_pad][ld], n_pad = nblk * 128; only tiles on or below the
// block diagonal are referenced ("lower": K = L L^T, L[i][j] at i*ld + j; the reference's
// upper factor U of cho_factor(lower=False) is L^T, same diagonal).  Padding rows carry the
// identity, so the factor of the padded matrix is [L 0; 0 I] and log-det / solves are unchanged.
//
// Tile algorithm over 128-wide block columns, right-looking between groups of `group` block
// columns and left-looking inside a group (see run_cholesky in agpn_api.cu): the wide trailing
// update applies a whole group at once (K = 128 * group).  Three kernels, each over ALL matrices
// of the batch (grid.y = matrix):
//   potrf_diag_kernel   one CTA per matrix: 128x128 diagonal block factorised in shared memory
//                       as 8 panels of 16 columns (serial 16x16 block by one warp with look-ahead,
//                       rank-16 updates on DMMA), its inverse by block forward substitution on
//                       DMMA, and -- folded in -- the forward-solve block z_k = L_kk^-1 r_k,
//                       |z_k|^2 and sum log L_jj (warp-shuffle reductions).
//   trsm_kernel         L_ik = A_ik * L_kk^-T as an NT GEMM with the inverted diagonal block
//                       on the FP64 tensor pipe (DMMA.8x8x4); also r_i -= L_ik z_k.
//   syrk_kernel         A_ij -= L_ik L_jk^T for all trailing tiles i >= j > k, DMMA.8x8x4,
//                       operands staged through a 4-stage cp.async shared-memory pipeline whose
//                       stages are handed over by mbarriers (no CTA-wide barrier per k-chunk), with
//                       an XOR swizzle that makes the 64-bit fragment loads bank-conflict free.
//                       CTA tile 128 x 64, two CTAs per SM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TB 128                 // tile edge (block column width of the factorisation)
#define BN 64                  // n extent of one GEMM CTA tile (128 x 64, two CTAs per SM)
#ifndef GEMM_KC
#define GEMM_KC 16             // k extent of one pipeline stage
#endif
#ifndef GEMM_STAGES
#define GEMM_STAGES 4
#endif
#define GEMM_STAGE_DOUBLES ((TB + BN) * GEMM_KC)
#define GEMM_SMEM_BYTES (GEMM_STAGES * GEMM_STAGE_DOUBLES * 8)   // 98304

struct CholArgs {
    double* A;              // [nmat][n_pad][ld]
    long long mat_stride;   // doubles between matrices
    int ld;
    int nblk;
    int nmat;
    double* Linv;           // [nmat][linv_blocks][TB][TB]
    int linv_blocks;        // 1 (reused every step) or nblk (kept, prediction path)
    double* rhs;            // [nmat][n_pad] or nullptr
    double* logdet;         // [nmat]  accumulated
    double* quad;           // [nmat]  accumulated |z|^2
    int* info;              // [nmat]  0 ok / 1 not positive definite
    double* A0;             // base of the whole batch allocation the tensor maps describe (TMA path)
    int mat0;               // index of this launch's first matrix inside that allocation
    // INT8-sliced trailing update (agpn_i8.cuh); S8 == nullptr: not in use
    double* rscale = nullptr;   // [nmat][2][n_pad]  Q / r_i  |  r_i / (126 2^24)
    signed char* S8 = nullptr;  // [nmat] digit planes of the current group's block columns
    int sgroup = 0;             // block columns the slice storage of one matrix holds
    int keep_factor = 0;        // 1: the FP64 panels of L are stored even when only the digit planes are read again
    int store_diag = 1;         // 0: the factored diagonal blocks are not written back (nobody reads them: likelihood waves)
    int* i8flag = nullptr;      // [nmat] or nullptr: set when a panel entry exceeds its row scale (digits would wrap)
    // ill-conditioned diagonal blocks: potrf_diag_kernel sets refine[mat] when max L_jj > refine_ratio * min L_jj for the
    // block it has just factorised; trsm_kernel then solves that block column by substitution (backward stable, like
    // LAPACK's dtrsm) instead of multiplying by the explicit inverse (whose error grows with cond(L_kk))
    int* refine = nullptr;      // [nmat][refine_blocks] or nullptr
    int refine_blocks = 1;      // 1: one flag per matrix, rewritten every block column (launch-ordered chain); nblk: one per
                                // diagonal block (the persistent task kernel, where panel solves of column k may still run
                                // when diagonal block k + 1 is factorised -- Linv is then kept per column too, linv_blocks = nblk)
    double refine_ratio = 1.0e3;
    // split panel solves (trsm_kernel with one CTA per 64-column half of a tile): the two CTAs' sums L_ik z_k of the folded
    // forward solve are parked here and subtracted from rhs by the next reader in a fixed order (bit-reproducible):
    // [nmat][nblk][2 column parity][2 halves][TB]; nullptr: one CTA per tile updates rhs itself
    double* rslot = nullptr;
    // which 32-column chunks of the digit planes hold a non-zero digit at all: bit kc32 of nzmask[mat][row tile][word].
    // The factor of a covariance whose correlation length is short against the time span is numerically banded -- entries
    // below 2^-57 of their row scale have ALL digits zero -- and a K step in which either operand chunk is all-zero adds
    // exact integer zeros: the persistent update skips it (loads, MMAs, and whole tiles without any active step).  Results
    // are bit-identical to the dense schedule.  nullptr: dense.
    unsigned long long* nzmask = nullptr;
    int nzwords = 0;
    // inf-norms of the two 64-row halves of Linv_kk, written by potrf_diag_kernel: [nmat][linv_blocks][2].  With the masks
    // above, trsm_kernel first bounds its result |X_ic| <= max_j |C_ij| * ||Linv half||_inf and leaves the tile half out
    // (no GEMM, no digits, no forward-solve sum, mask bits stay clear; zeros if the FP64 panel is kept) when every entry is
    // below 2^-40 of one digit unit -- below 2^-96 of its row scale, 2^-43 of an ulp of anything it could be added to.
    // nullptr: never.
    double* linvnorm = nullptr;
    // what the bound test knows about a tile without reading it: tmaxA[mat][lower tile index] bounds |K| of the assembled
    // tile (assemble_fast_kernel), tmaxU[mat][ti][k] holds max_i |C_ij| Q / r_i (bits of a non-negative double, 0 = not
    // recorded) of a tile the persistent update has just rewritten.  Either nullptr: the test scans the tile itself.
    const double* tmaxA = nullptr;
    unsigned long long* tmaxU = nullptr;
    int tile_order = 0;                         // persistent update: 0 = tile index fastest, 1 = matrix index fastest (narrow launches), 2 = 0 with a skewed round-robin
    unsigned long long* nzstat = nullptr;       // or nullptr: tile halves left out, [0] left (K = 64) and [1] right (K = 128) halves (agpn_skip_stats)
};
#define TRSM_SKIP_BOUND 9.094947017729282e-13      /* 2^-40 */
__host__ __device__ __forceinline__ size_t rslot_off(int nblk, int mat, int ti, int parity, int h) {
    return ((((size_t)mat * nblk + ti) * 2 + parity) * 2 + h) * 128;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Thread groups of 256: the FP64 bodies below (operand pipeline, panel solve, diagonal block) are written for 256
// threads.  In a 256-thread CTA the group is the CTA; the persistent task kernel (agpn_mega.cuh) runs 512-thread CTAs in
// which the two halves work on two tiles at once, each with its own named barrier.
// (G = false: plain 256-thread CTA, everything below compiles to what it was without groups)
template <bool G>
__device__ __forceinline__ int grp_tid() { return G ? (int)(threadIdx.x & 255u) : (int)threadIdx.x; }
template <bool G>
__device__ __forceinline__ int grp_id() { return G ? (int)(threadIdx.x >> 8) : 0; }
template <bool G>
__device__ __forceinline__ void grp_sync() {
    if (!G) __syncthreads();
    else if (threadIdx.x < 256) asm volatile("bar.sync 3, 256;\n" ::: "memory");
    else asm volatile("bar.sync 4, 256;\n" ::: "memory");
}

// ---- digit planes of the INT8-sliced trailing update (agpn_i8.cuh) ---------------------------------------------
// x = L_ik * Q / r_i, |x| <= 127 * 2^49.  q = rint(x) = sum_s a_s 128^(7-s) with balanced digits a_1 .. a_7 in
// [-64, 63] and a_0 = the signed rest: adding 64 * (128^0 + .. + 128^6) turns the balanced digits into plain 7-bit
// fields u = a + 64.  Returns the eight digits of two values as eight 16-bit pieces (value 0 in the low byte).
#define I8_NSL 8
#define I8_Q (126.0 * 562949953421312.0)                         // 126 * 2^49
#define I8_DIGIT_BIAS 0x0001020408102040ll                       // 64 * sum_{j<7} 128^j
#define I8_QMAXF (127.0 * 562949953421312.0)                     // largest magnitude the eight digits can hold
#define I8_TILE_BYTES (I8_NSL * TB * 32)                         // digit planes of 128 rows x 32 columns: 32 KB
// fields (x >> sh) & 127 of two values, two digits at a time, as bytes [v0 digit a, v1 digit a, v0 digit b, v1 digit b];
// the bias removal u - 64 (mod 256) on all four bytes at once: v = u ^ 0x40, byte = v | ((v & 0x40) << 1)
// (host versions of the two intrinsics so that tests/test_i8_digits_host.py can check the digit algebra without a GPU)
__host__ __device__ __forceinline__ unsigned i8_byte_perm(unsigned a, unsigned b, unsigned sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const unsigned long long ab = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((ab >> (8 * ((sel >> (4 * i)) & 7))) & 0xFFull) << (8 * i);
    return r;
#endif
}
__host__ __device__ __forceinline__ long long i8_rint(double x) {
#ifdef __CUDA_ARCH__
    return __double2ll_rn(x);
#else
    return llrint(x);          // round to nearest even, like cvt.rni
#endif
}
__host__ __device__ __forceinline__ unsigned i8_pack4(unsigned a0, unsigned a1, unsigned b0, unsigned b1, unsigned keep, unsigned flip, unsigned sign) {
    const unsigned lo = i8_byte_perm(a0, a1, 0x0040), hi = i8_byte_perm(b0, b1, 0x0040);
    const unsigned v = (i8_byte_perm(lo, hi, 0x5410) & keep) ^ flip;
    return v | ((v << 1) & sign);
}
__host__ __device__ __forceinline__ void i8_digits2(double x0, double x1, unsigned (&piece)[I8_NSL]) {
    // |x| <= 126 2^49 (1 + 1e-9) for every factorisable matrix (sum_k L_ik^2 = K_ii); a block that is not positive
    // definite or not finite is flagged by the factorisation and its digits only have to be harmless: NaN converts
    // to 0, overflow saturates / wraps
    const long long q0 = i8_rint(x0) + I8_DIGIT_BIAS, q1 = i8_rint(x1) + I8_DIGIT_BIAS;
    const unsigned l0 = (unsigned)q0, l1 = (unsigned)q1;
    const unsigned h0 = (unsigned)(q0 >> 28), h1 = (unsigned)(q1 >> 28);      // digits 4 .. 7 (bits 28 .. 56 + sign)
    // slice 7 - j holds digit j (weight 128^j); words: (slice 7 | slice 6 << 16), (5 | 4), (3 | 2), (1 | 0)
    const unsigned w76 = i8_pack4(l0, l1, l0 >> 7, l1 >> 7, 0x7F7F7F7Fu, 0x40404040u, 0x80808080u);
    const unsigned w54 = i8_pack4(l0 >> 14, l1 >> 14, l0 >> 21, l1 >> 21, 0x7F7F7F7Fu, 0x40404040u, 0x80808080u);
    const unsigned w32 = i8_pack4(h0, h1, h0 >> 7, h1 >> 7, 0x7F7F7F7Fu, 0x40404040u, 0x80808080u);
    // digit 6 and the signed rest (bits 49 ..: no bias on it)
    const unsigned w10 = i8_pack4(h0 >> 14, h1 >> 14, (unsigned)((int)h0 >> 21), (unsigned)((int)h1 >> 21), 0xFFFF7F7Fu, 0x00004040u, 0x00008080u);
    piece[7] = w76 & 0xFFFFu; piece[6] = w76 >> 16;
    piece[5] = w54 & 0xFFFFu; piece[4] = w54 >> 16;
    piece[3] = w32 & 0xFFFFu; piece[2] = w32 >> 16;
    piece[1] = w10 & 0xFFFFu; piece[0] = w10 >> 16;
}
__host__ __device__ __forceinline__ size_t i8_slices_per_matrix(int nblk, int sgroup) { return (size_t)(sgroup * 4) * nblk * I8_TILE_BYTES; }
__host__ __device__ __forceinline__ size_t i8_tile_off(int kc32, int rt, int nblk) { return ((size_t)kc32 * nblk + rt) * I8_TILE_BYTES; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// swizzled offset (in doubles) of element (row, k) inside a [rows][16] stage operand: the 16-byte
// chunk index is XORed with 2*(row & 3), which spreads the 8 rows x 4 k of one DMMA fragment load
// over all 32 banks (64-bit accesses, two half-warp phases).
__device__ __forceinline__ int swz_x(int row) {
#if GEMM_KC >= 16
    return (row & 3) << 1;             // 128-byte (or longer) rows: 4 rows x 2 chunks -> 8 distinct chunks
#else
    return ((row >> 1) & 1) << 1;      // 64-byte rows: two rows share a bank line
#endif
}
__device__ __forceinline__ int swz(int row, int k) {
    return row * GEMM_KC + ((((k >> 1) ^ swz_x(row)) << 1) | (k & 1));
}

// ---------------------------------------------------------------------------------------------
// acc[mi][ni][2] += A[128 x K] * B[64 x K]^T  (both row-major, k contiguous), 256 threads,
// warp grid 4 (m) x 2 (n), warp tile 32 x 32 = 4 x 4 DMMA tiles.  Two such CTAs share an SM so
// that one CTA's prologue / epilogue memory traffic overlaps the other's tensor work.
// Accumulator element (mi, ni, e) of lane l lives at
//     row = wm*32 + mi*8 + (l >> 2),  col = wn*32 + ni*8 + 2*(l & 3) + e.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_load_stage(double* stage, const double* __restrict__ Ag, int lda,
                                                const double* __restrict__ Bg, int ldb, int kc, int tid) {
    constexpr int CPR = GEMM_KC / 2;                 // 16-byte chunks per operand row
    constexpr int A_ITERS = TB * CPR / 256, B_ITERS = BN * CPR / 256;
#pragma unroll
    for (int u = 0; u < A_ITERS; ++u) {
        const int idx = tid + 256 * u;
        const int row = idx / CPR, ch = idx % CPR;
        cp_async16(stage + row * GEMM_KC + ((ch ^ swz_x(row)) << 1), Ag + (size_t)row * lda + kc * GEMM_KC + ch * 2);
    }
#pragma unroll
    for (int u = 0; u < B_ITERS; ++u) {
        const int idx = tid + 256 * u;
        const int row = idx / CPR, ch = idx % CPR;
        cp_async16(stage + TB * GEMM_KC + row * GEMM_KC + ((ch ^ swz_x(row)) << 1),
                   Bg + (size_t)row * ldb + kc * GEMM_KC + ch * 2);
    }
}

// NI0: first of the warp's four 8-column groups that takes part (2: the triangular operand's rows of groups 0 and 1 are zero
// in this k-chunk); NIH: group whose rows are zero in the second half of the chunk (its diagonal 8 x 8 block sits in the first
// half), -1: none
template <int NI0 = 0, int NIH = -1>
__device__ __forceinline__ void gemm_compute_stage(const double* stage, double (&acc)[4][4][2], int wm, int wn, int lane) {
    const double* As = stage;
    const double* Bs = stage + TB * GEMM_KC;
    const int r = lane >> 2;
    const int kq = lane & 3;
#pragma unroll
    for (int ks = 0; ks < GEMM_KC / 4; ++ks) {
        double a[4], b[4];
        const int k = ks * 4 + kq;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = As[swz(wm * 32 + mi * 8 + r, k)];
#pragma unroll
        for (int ni = NI0; ni < 4; ++ni)
            if (!(ni == NIH && ks >= GEMM_KC / 8)) b[ni] = Bs[swz(wn * 32 + ni * 8 + r, k)];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = NI0; ni < 4; ++ni)
                if (!(ni == NIH && ks >= GEMM_KC / 8)) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// ---- mbarrier-decoupled cp.async pipeline ----------------------------------------------------------
// A CTA-wide __syncthreads per k-chunk makes every warp wait for the slowest one (14 % of warp time
// in the K = 512 update, ncu r01c).  Instead each stage has two mbarriers:
//   full[s]   256 arrivals: every thread's cp.async group for that stage signals it on completion
//             (cp.async.mbarrier.arrive.noinc) -> a warp waits only for the DATA;
//   empty[s]  8 arrivals: one per warp after its last read of the stage -> a stage is refilled only
//             when all warps are done with it, checked one chunk later than needed (slack).
// `g` counts chunks over the whole kernel so the phase parities stay consistent across several GEMMs.
struct GemmPipe {
    unsigned full;    // shared-space address of full[GEMM_STAGES]
    unsigned empty;   // shared-space address of empty[GEMM_STAGES]
    unsigned g;       // chunks consumed so far by this thread (identical in all threads)
};

__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(unsigned addr) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    unsigned done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();     // a protocol bug must fault, not hang the GPU
    }
}

// call once per kernel by all 256 threads, before the first GEMM; bars = 2 * GEMM_STAGES uint64_t in shared
template <bool G = false>
__device__ __forceinline__ GemmPipe gemm_pipe_init(unsigned long long* bars) {
    GemmPipe p;
    p.full = (unsigned)__cvta_generic_to_shared(bars);
    p.empty = p.full + 8 * GEMM_STAGES;
    p.g = 0;
    if (grp_tid<G>() == 0) {
#pragma unroll
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(p.full + 8 * s, 256);
            mbar_init(p.empty + 8 * s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    grp_sync<G>();
    return p;
}

// issue the loads of local chunk `c` (global chunk index p.g + c) into its stage
__device__ __forceinline__ void gemm_issue_chunk(const GemmPipe& p, int c, const double* __restrict__ Ag, int lda,
                                                 const double* __restrict__ Bg, int ldb, double* smem, int tid) {
    const unsigned n = p.g + c;
    const unsigned s = n % GEMM_STAGES;
    if (n >= GEMM_STAGES) mbar_wait(p.empty + 8 * s, ((n / GEMM_STAGES) - 1) & 1);
    gemm_load_stage(smem + s * GEMM_STAGE_DOUBLES, Ag, lda, Bg, ldb, c, tid);
    mbar_cp_async_arrive(p.full + 8 * s);
}

// K must be a multiple of GEMM_KC.  All 256 threads must call both halves.  `active` = this warp
// issues the tensor work (warps whose output is not needed still take part in the loads and the
// barriers).  gemm_nt_prologue only ISSUES the first pipeline stages, so the caller can put its own
// global loads (the C tile) in flight behind them before anything waits.  When gemm_nt_mainloop
// returns, every cp.async issued by ANY thread of the CTA for this GEMM has completed (the last
// full-barrier needs all 256 arrivals), so global operands may be overwritten; shared memory may
// still be read by slower warps -- callers that reuse it by other means need their own barrier.
template <bool G = false>
__device__ __forceinline__ void gemm_nt_prologue(const GemmPipe& p, const double* __restrict__ Ag, int lda,
                                                 const double* __restrict__ Bg, int ldb, int K, double* smem) {
    const int tid = grp_tid<G>();
    const int nch = K / GEMM_KC;
#pragma unroll
    for (int c = 0; c < GEMM_STAGES - 1; ++c)
        if (c < nch) gemm_issue_chunk(p, c, Ag, lda, Bg, ldb, smem, tid);
}

// kc_limit: this warp's operand columns are exactly zero from k-chunk kc_limit on (triangular B): it skips the tensor
// work of those chunks but still takes part in the loads and the barriers.  tri: B is lower triangular with its diagonal
// running through the warp's last two chunks (16 columns each against the warp's 32 rows of B): in the last one the first two
// 8-row groups of B are zero as well, and in each of the two the group whose diagonal 8 x 8 block lies in the chunk's first half
// is zero in its second half.
template <bool G = false>
__device__ __forceinline__ void gemm_nt_mainloop(GemmPipe& p, const double* __restrict__ Ag, int lda,
                                                 const double* __restrict__ Bg, int ldb, int K, double (&acc)[4][4][2],
                                                 double* smem, bool active = true, int kc_limit = 1 << 30, bool tri = false) {
    const int tid = grp_tid<G>();
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp >> 1;
    const int wn = warp & 1;
    const int nch = K / GEMM_KC;
    for (int kc = 0; kc < nch; ++kc) {
        const unsigned n = p.g + kc;
        const unsigned s = n % GEMM_STAGES;
#ifdef GEMM_ISSUE_BEFORE
        const int nxt = kc + GEMM_STAGES - 1;
        if (nxt < nch) gemm_issue_chunk(p, nxt, Ag, lda, Bg, ldb, smem, tid);
        mbar_wait(p.full + 8 * s, (n / GEMM_STAGES) & 1);
        if (active) gemm_compute_stage(smem + s * GEMM_STAGE_DOUBLES, acc, wm, wn, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(p.empty + 8 * s);
#else
        mbar_wait(p.full + 8 * s, (n / GEMM_STAGES) & 1);
        if (active && kc < kc_limit) {
            if (tri && kc == kc_limit - 1) gemm_compute_stage<2, 2>(smem + s * GEMM_STAGE_DOUBLES, acc, wm, wn, lane);
            else if (tri && kc == kc_limit - 2) gemm_compute_stage<0, 0>(smem + s * GEMM_STAGE_DOUBLES, acc, wm, wn, lane);
            else gemm_compute_stage(smem + s * GEMM_STAGE_DOUBLES, acc, wm, wn, lane);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(p.empty + 8 * s);
        const int nxt = kc + GEMM_STAGES - 1;
        if (nxt < nch) gemm_issue_chunk(p, nxt, Ag, lda, Bg, ldb, smem, tid);
#endif
    }
    p.g += nch;
}

template <bool G = false>
__device__ __forceinline__ void gemm_nt_tile(GemmPipe& p, const double* __restrict__ Ag, int lda, const double* __restrict__ Bg,
                                             int ldb, int K, double (&acc)[4][4][2], double* smem, bool active = true,
                                             int kc_limit = 1 << 30, bool tri = false) {
    gemm_nt_prologue<G>(p, Ag, lda, Bg, ldb, K, smem);
    gemm_nt_mainloop<G>(p, Ag, lda, Bg, ldb, K, acc, smem, active, kc_limit, tri);
}

// ---------------------------------------------------------------------------------------------
// trailing update with `depth` already-factored block columns kfirst .. kfirst+depth-1 at once:
//     A_ij -= [L_i,kfirst .. L_i,kfirst+depth-1] [L_j,kfirst .. ]^T      (K = 128 * depth)
// The block columns of one block row are adjacent in memory, so a deeper update is the same
// NT GEMM with a longer k loop: the C tile is read and written once per `depth` panels and the
// pipeline prologue / epilogue is amortised over depth x the tensor work.
//   narrow = 1: only block column j0 (tiles (i, j0), i >= j0)      grid.x = 2 * (nblk - j0)
//   narrow = 0: all tiles i >= j >= j0                              grid.x = nt * (nt + 1), nt = nblk - j0
// one CTA per 128 x 64 half tile; grid.y = nmat
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) syrk_kernel(CholArgs g, int kfirst, int depth, int j0, int narrow) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long bars[2 * GEMM_STAGES];
    GemmPipe pipe = gemm_pipe_init(bars);
    const int t = blockIdx.x >> 1;
    const int h = blockIdx.x & 1;
    int bi, bj;
    if (narrow) {
        bi = t;
        bj = 0;
    } else {
        bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        while (bi * (bi + 1) / 2 > t) --bi;
        bj = t - bi * (bi + 1) / 2;
    }
    const int ti = j0 + bi, tj = j0 + bj;
    double* A = g.A + (size_t)blockIdx.y * g.mat_stride;
    const int ld = g.ld;
    const double* Ai = A + (size_t)ti * TB * ld + (size_t)kfirst * TB;
    const double* Aj = A + ((size_t)tj * TB + h * BN) * ld + (size_t)kfirst * TB;
    double* C = A + (size_t)ti * TB * ld + (size_t)tj * TB + h * BN;

    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    // right half of a diagonal tile: rows 0..63 lie strictly above the diagonal, never read
    const bool active = !(bi == bj && h == 1 && wm < 2);
    double acc[4][4][2];
    // operand pipeline first, the C tile loads go in flight behind it
    gemm_nt_prologue(pipe, Ai, ld, Aj, ld, depth * TB, smem);
    // acc = -C, so that -(acc + A B^T) = C - A B^T
    if (active) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int row = wm * 32 + mi * 8 + (lane >> 2);
                const int col = wn * 32 + ni * 8 + 2 * (lane & 3);
                const double2 c = *reinterpret_cast<const double2*>(C + (size_t)row * ld + col);
                acc[mi][ni][0] = -c.x;
                acc[mi][ni][1] = -c.y;
            }
    }
    gemm_nt_mainloop(pipe, Ai, ld, Aj, ld, depth * TB, acc, smem, active);
    if (active) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int row = wm * 32 + mi * 8 + (lane >> 2);
                const int col = wn * 32 + ni * 8 + 2 * (lane & 3);
                double2 c;
                c.x = -acc[mi][ni][0];
                c.y = -acc[mi][ni][1];
                *reinterpret_cast<double2*>(C + (size_t)row * ld + col) = c;
            }
    }
}

// K steps (32-column chunks ks < nks) of tile (ti, tj) in which BOTH operands hold a non-zero digit, 64 per word
// (CholArgs::nzmask; without masks every step is active)
__device__ __forceinline__ unsigned long long i8_active_word(const CholArgs& g, int mat, int ti, int tj, int word, int nks) {
    const int left = nks - word * 64;
    if (left <= 0) return 0ull;
    unsigned long long a = ~0ull;
    if (g.nzmask != nullptr)
        a = g.nzmask[((size_t)mat * g.nblk + ti) * g.nzwords + word] & g.nzmask[((size_t)mat * g.nblk + tj) * g.nzwords + word];
    if (left < 64) a &= (1ull << left) - 1ull;
    return a;
}
__device__ __forceinline__ bool i8_any_active(const CholArgs& g, int mat, int ti, int tj, int nks) {
    if (g.nzmask == nullptr) return nks > 0;
    for (int wd = 0; wd * 64 < nks; ++wd)
        if (i8_active_word(g, mat, ti, tj, wd, nks)) return true;
    return false;
}


// ---------------------------------------------------------------------------------------------
// panel solve: A_ik <- A_ik * Linv_kk^T (i > k), and rhs_i -= L_ik z_k
// grid.x = nblk-k-1, grid.y = nmat.  Two passes per CTA over the two 64-column halves; the right
// half goes first (it reads all 128 input columns), the left half needs only K = 64 because the
// inverse is lower triangular -- this also makes the in-place overwrite race free.
// ---------------------------------------------------------------------------------------------
// flags bit 0: also write the digit planes of the new panel (INT8-sliced update; k0 = first block column of the slice
// storage's group) -- digits are staged in the idle pipeline buffers in the tensor core's operand layout and leave
// as two contiguous 32 KB blocks per half;  bit 1: do not store the FP64 panel (nothing reads it when every later
// update runs on the digit planes).
// body of the panel solve for tile (ti, k) of matrix `mat`: all 256 threads of the CTA; `smem` = GEMM_SMEM_BYTES of
// dynamic shared memory, `pipe` = the CTA's operand pipeline (its chunk counter runs on across calls)
template <bool G = false>
__device__ __forceinline__ void trsm_body(const CholArgs& g, int k, int k0, int flags, int ti, int mat, double* smem, GemmPipe& pipe,
                                          int hsel = -1) {
    __shared__ double zs_[G ? 2 : 1][TB];
    __shared__ double qs_[G ? 2 : 1][TB];      // Q / r_i of this row tile (digit planes)
    __shared__ double red_[G ? 2 : 1][2][TB];
    __shared__ int s_nz_[G ? 2 : 1][2];
    double* zs = zs_[grp_id<G>()];
    double* qs = qs_[grp_id<G>()];
    double (*red)[TB] = red_[grp_id<G>()];
    int* s_nz = s_nz_[grp_id<G>()];
    double* A = g.A + (size_t)mat * g.mat_stride;
    const int ld = g.ld;
    double* Aik = A + (size_t)ti * TB * ld + (size_t)k * TB;
    const double* Linv = g.Linv + ((size_t)mat * g.linv_blocks + (g.linv_blocks > 1 ? k : 0)) * (TB * TB);
    double* rhs = g.rhs ? g.rhs + (size_t)mat * ((size_t)g.nblk * TB) : nullptr;

    const int tid = grp_tid<G>(), lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    if (rhs != nullptr && tid < TB) zs[tid] = rhs[(size_t)k * TB + tid];
    if ((flags & 1) && tid >= TB) qs[tid - TB] = g.rscale[(size_t)mat * 2 * ((size_t)g.nblk * TB) + (size_t)ti * TB + tid - TB];
    // ill-conditioned diagonal block (flag set by potrf_diag_kernel): X L_kk^T = A_ik by substitution, in place, one
    // thread per row, 16 columns at a time in registers; L_kk sits packed (lower triangle) in the pipeline buffers
    const bool refine = g.refine != nullptr && g.refine[(size_t)mat * g.refine_blocks + (g.refine_blocks > 1 ? k : 0)] != 0;   // CTA-uniform
    // hsel >= 0: this CTA solves only the 64-column half hsel of the tile (its partner the other one).  The substitution solve
    // is not split: the h = 1 CTA takes the whole tile, the h = 0 CTA only keeps the hand-over of the forward-solve sums going
    const bool split = hsel >= 0 && g.rslot != nullptr && rhs != nullptr;
    if (hsel >= 0 && refine) {
        if (hsel == 0) {
            if (split && tid < TB) {
                if (k > 0) {
                    const double* pp = g.rslot + rslot_off(g.nblk, mat, ti, (k - 1) & 1, 0);
                    rhs[(size_t)ti * TB + tid] -= pp[tid] + pp[TB + tid];
                }
                g.rslot[rslot_off(g.nblk, mat, ti, k & 1, 0) + tid] = 0.0;
            }
            return;
        }
        hsel = -1;
    }
    if (refine) {
        const double* Lkk = A + (size_t)k * TB * ld + (size_t)k * TB;
        for (int idx = tid; idx < TB * TB; idx += 256) {
            const int i = idx >> 7, j = idx & (TB - 1);
            if (j <= i) smem[i * (i + 1) / 2 + j] = Lkk[(size_t)i * ld + j];
        }
        grp_sync<G>();
        if (tid < TB) {
            double* xr = Aik + (size_t)tid * ld;
#pragma unroll 1
            for (int b = 0; b < TB / 16; ++b) {
                double x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = xr[16 * b + j];
#pragma unroll 1
                for (int m = 0; m < 16 * b; ++m) {
                    const double xm = xr[m];
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = fma(-xm, smem[(16 * b + j) * (16 * b + j + 1) / 2 + m], x[j]);
                }
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int gm = 16 * b + m;
                    x[m] = x[m] / smem[gm * (gm + 1) / 2 + gm];
#pragma unroll
                    for (int j = m + 1; j < 16; ++j) x[j] = fma(-x[m], smem[(16 * b + j) * (16 * b + j + 1) / 2 + gm], x[j]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) xr[16 * b + j] = x[j];
            }
        }
    }
    grp_sync<G>();
    // skip_h[h]: the 64-column half h of the tile is left out.  One CTA per half (hsel) decides for its half, a CTA that owns
    // the whole tile for both -- with the same bounds and the same arithmetic, so the results do not depend on the launch mode
    bool skip_h[2] = {false, false};
    if (g.linvnorm != nullptr && g.nzmask != nullptr && (flags & 1) && !refine) {
        // bound the half before touching the tensor pipe: |X_ic| <= max_j |C_ij| * ||rows of Linv of this half||_inf
        // everything the decision reads is fetched up front, independent of each other: one global-memory latency, not a chain
        // of three (this is all a CTA whose half is left out ever does, and it delays the pipeline fill of the others)
        const double* ln = g.linvnorm + ((size_t)mat * g.linv_blocks + (g.linv_blocks > 1 ? k : 0)) * 2;
        const double ln0 = ln[0], ln1 = ln[1];
        const unsigned long long bU = (g.tmaxU != nullptr) ? g.tmaxU[((size_t)mat * g.nblk + ti) * g.nblk + k] : 0ull;
        const double bA = (g.tmaxA != nullptr && k0 == 0)
            ? g.tmaxA[(size_t)mat * ((size_t)g.nblk * (g.nblk + 1) / 2) + (size_t)ti * (ti + 1) / 2 + k] : -1.0;
        unsigned long long act = 0ull;                          // any K step of the update on this tile active?  (k > k0)
        {
            const int nks = 4 * (k - k0);
            const unsigned long long* mi_ = g.nzmask + ((size_t)mat * g.nblk + ti) * g.nzwords;
            const unsigned long long* mk_ = g.nzmask + ((size_t)mat * g.nblk + k) * g.nzwords;
            for (int wd = 0; wd * 64 < nks; ++wd) {
                unsigned long long a = mi_[wd] & mk_[wd];
                const int left = nks - wd * 64;
                if (left < 64) a &= (1ull << left) - 1ull;
                act |= a;
            }
        }
        const double thr0 = TRSM_SKIP_BOUND / ln0, thr1 = TRSM_SKIP_BOUND / ln1;     // norm inf / NaN: nothing passes
        int bad0 = 0, bad1 = 0;
        // without reading the tile: the update has just rewritten it and recorded max |C_ij| Q / r_i, or no K step of it was
        // active (same mask test as the update kernel) and it still holds the assembled entries, bounded by tmaxA
        double known = -1.0;                                    // bound on |C_ij| (x Q / r_i if `scaled`), < 0: unknown
        bool scaled = false;
        if (act != 0ull) {
            if (bU != 0ull) { known = __longlong_as_double((long long)bU); scaled = true; }
        } else {
            known = bA;
        }
        if (known >= 0.0) {                                      // false for NaN
            if (tid < TB) {
                const double v = known * (scaled ? 1.0 : qs[tid]);
                bad0 = !(v < thr0);
                bad1 = !(v < thr1);
            }
        } else {
            // warp w scans rows 16 w .. 16 w + 15, a row's columns spread over the lanes (coalesced), eight rows in flight;
            // the left half depends on input columns 0 .. 63 only, the right half on all 128
            const bool want_right = hsel != 0;
#pragma unroll
            for (int r8 = 0; r8 < 2; ++r8) {
                double2 v[8][2];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double* src = Aik + (size_t)(warp * 16 + r8 * 8 + r) * ld;
                    v[r][0] = *reinterpret_cast<const double2*>(src + 2 * lane);
                    v[r][1] = want_right ? *reinterpret_cast<const double2*>(src + 64 + 2 * lane) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    // the sums bound the maxima and, unlike fmax, keep a NaN
                    const double q = qs[warp * 16 + r8 * 8 + r];
                    const double ml = fabs(v[r][0].x) + fabs(v[r][0].y);
                    const double ma = ml + (fabs(v[r][1].x) + fabs(v[r][1].y));
                    bad0 |= !(ml * q < thr0);                    // NaN / inf count as "solve it"
                    bad1 |= !(ma * q < thr1);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (hsel >= 0 && h != hsel) continue;
            const int bad = h ? bad1 : bad0;
            if constexpr (!G) {
                skip_h[h] = __syncthreads_or(bad) == 0;
            } else {
                if (tid == 0) s_nz[0] = 0;
                grp_sync<G>();
                if (bad) s_nz[0] = 1;
                grp_sync<G>();
                skip_h[h] = s_nz[0] == 0;
                grp_sync<G>();
            }
            if (skip_h[h] && tid == 0 && g.nzstat != nullptr) atomicAdd(g.nzstat + h, 1ull);
        }
    }
    // forward-solve sums: a CTA that owns both halves keeps them apart and adds them like the two CTAs of a split solve do
    const bool two_sums = rhs != nullptr && !refine && hsel < 0;
    double hsum[2] = {0.0, 0.0};
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    double acc[4][4][2];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int h = 1 - pass;
        if (hsel >= 0 && h != hsel) continue;
        if (skip_h[h]) continue;                                 // CTA-uniform
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        // rows h * 64 + wn * 32 .. + 31 of the (lower triangular) inverse are zero right of column h * 64 + wn * 32 + 31:
        // the wn = 0 warps stop 32 columns (two k-chunks) early
        if (refine) {
            // the panel has been solved in place above: fetch this half in accumulator layout
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int row = wm * 32 + mi * 8 + (lane >> 2);
                    const int col = h * BN + wn * 32 + ni * 8 + 2 * (lane & 3);
                    const double2 v = *reinterpret_cast<const double2*>(Aik + (size_t)row * ld + col);
                    acc[mi][ni][0] = v.x;
                    acc[mi][ni][1] = v.y;
                }
        } else
        gemm_nt_tile<G>(pipe, Aik, ld, Linv + (size_t)h * BN * TB, TB, (h + 1) * BN, acc, smem, true,
                     (h * BN + wn * 32 + 32) / GEMM_KC, GEMM_KC == 16);
        // every load of this GEMM (by any thread) has completed: the in-place store is safe
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int row = wm * 32 + mi * 8 + (lane >> 2);
                const int col = h * BN + wn * 32 + ni * 8 + 2 * (lane & 3);
                if (!(flags & 2) && !refine) *reinterpret_cast<double2*>(Aik + (size_t)row * ld + col) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                if (rhs != nullptr) {
                    part[mi] = fma(acc[mi][ni][0], zs[col], part[mi]);
                    part[mi] = fma(acc[mi][ni][1], zs[col + 1], part[mi]);
                }
            }
        if (flags & 1) {
            if (tid < 2) s_nz[tid] = 0;
            grp_sync<G>();                                       // slower warps may still read the last operand stages
            unsigned char* stg = reinterpret_cast<unsigned char*>(smem);      // [wn = 32-column chunk][slice][4096]
            int wrapped = 0;
            unsigned nzbits = 0;
            const bool check_wrap = g.i8flag != nullptr;        // only the prediction paths can exceed a row scale
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const double f = qs[wm * 32 + mi * 8 + (lane >> 2)];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    unsigned piece[I8_NSL];
                    const double x0 = acc[mi][ni][0] * f, x1 = acc[mi][ni][1] * f;
                    const double xm = fmax(fabs(x0), fabs(x1));
                    if (check_wrap && xm > I8_QMAXF) wrapped = 1;
                    if (xm > 0.5) nzbits = 1;                    // rint(x) != 0 (ties go to the even 0): some digit is non-zero
                    i8_digits2(x0, x1, piece);
                    unsigned char* dst = stg + wn * I8_TILE_BYTES + (wm * 4 + mi) * 256 + (ni >> 1) * 128 + (lane >> 2) * 16 +
                                         (ni & 1) * 8 + 2 * (lane & 3);
#pragma unroll
                    for (int sl = 0; sl < I8_NSL; ++sl) {
                        *reinterpret_cast<unsigned short*>(dst + sl * (TB * 32)) = (unsigned short)piece[sl];
                    }
                }
            }
            if (wrapped && g.i8flag != nullptr) atomicOr(g.i8flag + mat, 1);
            if (nzbits) s_nz[wn] = 1;                            // this 32-column chunk holds a non-zero digit
            grp_sync<G>();
            if (g.nzmask != nullptr && tid < 2 && s_nz[tid]) {
                const int kc32 = 4 * (k - k0) + 2 * h + tid;
                atomicOr(g.nzmask + ((size_t)mat * g.nblk + ti) * g.nzwords + (kc32 >> 6), 1ull << (kc32 & 63));
            }
            signed char* S = g.S8 + (size_t)mat * i8_slices_per_matrix(g.nblk, g.sgroup);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                // a chunk without a non-zero digit is never read when the update goes by the masks: its store is left out
                if (g.nzmask != nullptr && !s_nz[c]) continue;
                uint4* dstg = reinterpret_cast<uint4*>(S + i8_tile_off(4 * (k - k0) + 2 * h + c, ti, g.nblk));
                const uint4* src = reinterpret_cast<const uint4*>(stg + c * I8_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < I8_TILE_BYTES / 16 / 256; ++i) dstg[i * 256 + tid] = src[i * 256 + tid];
            }
            grp_sync<G>();                                       // the next pass refills the pipeline buffers
        }
        if (two_sums) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double s = part[mi];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if ((lane & 3) == 0) red[wn][wm * 32 + mi * 8 + (lane >> 2)] = s;
                part[mi] = 0.0;
            }
            grp_sync<G>();
            if (tid < TB) hsum[h] = red[0][tid] + red[1][tid];
            grp_sync<G>();
        }
    }
    if (!(flags & 2)) {
        // the FP64 panel is kept (prediction reads V back): a half that was left out -- entries below 2^-96 of their row
        // scale -- becomes zeros, after the other half's solve has read its input columns
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (skip_h[h])
                for (int idx = tid; idx < TB * (BN / 2); idx += 256) {
                    const int row = idx / (BN / 2), c2 = idx - row * (BN / 2);
                    *reinterpret_cast<double2*>(Aik + (size_t)row * ld + h * BN + 2 * c2) = make_double2(0.0, 0.0);
                }
    }
    if (two_sums) {
        if (tid < TB) rhs[(size_t)ti * TB + tid] -= hsum[0] + hsum[1];
    } else if (rhs != nullptr) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            double s = part[mi];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if ((lane & 3) == 0) red[wn][wm * 32 + mi * 8 + (lane >> 2)] = s;
        }
        grp_sync<G>();
        if (tid < TB) {
            if (split) {
                // the h = 0 CTA subtracts the previous column's two sums (both complete: the previous launch), then both park theirs;
                // the h = 1 CTA of an un-split (substitution) tile parks the whole sum
                const int hh = hsel >= 0 ? hsel : 1;
                if (hh == 0 && k > 0) {
                    const double* pp = g.rslot + rslot_off(g.nblk, mat, ti, (k - 1) & 1, 0);
                    rhs[(size_t)ti * TB + tid] -= pp[tid] + pp[TB + tid];
                }
                g.rslot[rslot_off(g.nblk, mat, ti, k & 1, hh) + tid] = red[0][tid] + red[1][tid];
            } else {
                rhs[(size_t)ti * TB + tid] -= red[0][tid] + red[1][tid];
            }
        }
    }
}

__global__ void __launch_bounds__(256, 2) trsm_kernel(CholArgs g, int k, int k0, int flags) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long bars[2 * GEMM_STAGES];
    GemmPipe pipe = gemm_pipe_init(bars);
    // flags bit 3: grid.x = 2 (nblk - k - 1), one CTA per 64-column half of a tile (the heavier right half first)
    if (flags & 8) trsm_body(g, k, k0, flags, k + 1 + (int)(blockIdx.x >> 1), (int)blockIdx.y, smem, pipe, 1 - (int)(blockIdx.x & 1));
    else trsm_body(g, k, k0, flags, k + 1 + (int)blockIdx.x, (int)blockIdx.y, smem, pipe);
}

// ---------------------------------------------------------------------------------------------
// diagonal block: factor, invert, forward-solve block, log-det.   grid.x = nmat, 256 threads.
//
// The 128 x 128 block lives in shared memory (row stride 132 doubles: conflict-free DMMA
// fragment loads).  It is processed as 8 panels of 16 columns:
//   A  warp 0 factors the 16 x 16 diagonal block in registers (lane r owns row r; pivots and
//      column values travel by warp shuffles) -- the only inherently serial part;
//   B  one thread per row below solves its 16 panel entries against that block;
//   C  the trailing 16 x 16 blocks take the rank-16 update on DMMA.8x8x4, one block per warp.
// The inverse (needed by trsm_kernel, which applies L_kk^-T as a GEMM) is built from the 16 x 16
// diagonal inverses by block forward substitution, again on DMMA, and kept transposed in the
// (otherwise unused) upper triangle of the same shared tile.
// ---------------------------------------------------------------------------------------------
#define PD_LD 132
#define PD_DV_LD 20
#define POTRF_SMEM_DOUBLES (TB * PD_LD + 3 * TB + 8 * 16 * PD_DV_LD + 8 * 16 * PD_DV_LD)
#define POTRF_SMEM_BYTES (POTRF_SMEM_DOUBLES * 8)

// 16 x 16 diagonal block of panel pb, one warp: lane r (and its mirror r + 16) owns row r in
// registers.  Per column the UNSCALED column is published once through shared memory; every lane
// reads the pivot and the entries it needs back as broadcasts, so the serial chain per column is
// one shared-memory round trip + rsqrt + two multiplies + one fma.
__device__ __forceinline__ void potrf16_warp(double* S, double* invd, double* ub, int c0, int lane, int* s_fail) {
    const int r = lane & 15;
    double d[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) d[c] = S[(c0 + r) * 132 + c0 + c];
    double myrs = 1.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        double* u = ub + (j & 1) * 16;
        u[r] = d[j];
        __syncwarp();
        double piv = u[j];
        if (!(piv > 0.0) || isinf(piv)) {          // LAPACK dpotrf: ajj <= 0 or NaN -> not PD
            bad = 1;
            piv = 1.0;
        }
        const double rs = rsqrt(piv);              // <= 1 ulp; keeps the serial chain short
        const double t = d[j] * (rs * rs);
#pragma unroll
        for (int c = j + 1; c < 16; ++c) d[c] = fma(-t, u[c], d[c]);
        d[j] = (j == r) ? piv * rs : d[j] * rs;
        if (j == r) myrs = rs;
    }
    if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c <= r) S[(c0 + r) * 132 + c0 + c] = d[c];
        invd[c0 + r] = myrs;
    }
    if (bad && lane == 0) *s_fail = 1;
}

// rank-16 update of one 16 x 16 block (rows r0.., cols q0..) with panel columns c0..c0+15, one warp
__device__ __forceinline__ void potrf_block_update(double* S, int r0, int q0, int c0, int lane) {
    double acc[2][2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const double2 c = *reinterpret_cast<const double2*>(S + (r0 + 8 * mi + (lane >> 2)) * 132 + q0 + 8 * ni + 2 * (lane & 3));
            acc[mi][ni][0] = c.x;
            acc[mi][ni][1] = c.y;
        }
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        double a[2], b[2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) a[mi] = -S[(r0 + 8 * mi + (lane >> 2)) * 132 + c0 + 4 * k4 + (lane & 3)];
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) b[ni] = S[(q0 + 8 * ni + (lane >> 2)) * 132 + c0 + 4 * k4 + (lane & 3)];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
            *reinterpret_cast<double2*>(S + (r0 + 8 * mi + (lane >> 2)) * 132 + q0 + 8 * ni + 2 * (lane & 3)) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// body for the diagonal block k of matrix `mat`: all 256 threads of the CTA, `smem` = POTRF_SMEM_BYTES
template <bool G = false>
__device__ __forceinline__ void potrf_diag_body(const CholArgs& g, int k, int mat, double* smem) {
    double* S = smem;                          // [128][132]: lower = A -> L ; strict upper = Linv^T
    double* invd = S + TB * PD_LD;             // [128] 1 / L_jj
    double* zs = invd + TB;                    // [128] rhs block
    double* zout = zs + TB;                    // [128] z block
    double* Dv = zout + TB;                    // [8][16][20] diagonal 16 x 16 inverses (full squares)
    double* scr = Dv + 8 * 16 * PD_DV_LD;      // [8][16][20] per-warp scratch
    __shared__ int s_fail;
    __shared__ int s_refine;
    __shared__ double s_ln_[G ? 2 : 1][4];
    double* s_ln = s_ln_[grp_id<G>()];

    const int tid = grp_tid<G>(), lane = tid & 31, warp = tid >> 5;
    const int ld = g.ld;
    double* Akk = g.A + (size_t)mat * g.mat_stride + (size_t)k * TB * ld + (size_t)k * TB;
    double* rhs = g.rhs ? g.rhs + (size_t)mat * ((size_t)g.nblk * TB) + (size_t)k * TB : nullptr;
    if (tid == 0) s_fail = 0;

    // ---- load the tile: warp w takes rows 16 w .. 16 w + 15, all 32 loads of a thread in flight ------
    {
        double2 v[16][2];
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const double* src = Akk + (size_t)(warp * 16 + it) * ld;
            v[it][0] = *reinterpret_cast<const double2*>(src + 2 * lane);
            v[it][1] = *reinterpret_cast<const double2*>(src + 64 + 2 * lane);
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int i = warp * 16 + it;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int c = hh * 64 + 2 * lane;
                double2 o = v[it][hh];
                if (c > i) o.x = 0.0;
                if (c + 1 > i) o.y = 0.0;
                *reinterpret_cast<double2*>(S + i * PD_LD + c) = o;
            }
        }
    }
    if (rhs != nullptr && tid < TB) {
        double r = rhs[tid];
        if (g.rslot != nullptr && k > 0) {
            // split panel solves: the last column's two half-tile sums L_k,k-1 z_k-1 are still parked (fixed order)
            const double* pp = g.rslot + rslot_off(g.nblk, mat, k, (k - 1) & 1, 0);
            r -= pp[tid] + pp[TB + tid];
        }
        zs[tid] = r;
    }
    grp_sync<G>();

    // ---- factor: panels of 16 columns; the diagonal block of panel pb+1 (serial, one warp) is
    //      factorised while the other warps apply panel pb to the rest of the trailing matrix --------
    if (warp == 0) potrf16_warp(S, invd, scr, 0, lane, &s_fail);
#ifdef WHATIF_SKIP_FACTOR
    for (int pb = 8; pb < 8; ++pb) {
#else
    for (int pb = 0; pb < 8; ++pb) {
#endif
        const int c0 = pb * 16;
        grp_sync<G>();
        // B: rows below the diagonal block, one thread per row: x L_D^T = b
        const int nrows = TB - c0 - 16;
        if (tid < nrows) {
            // right-looking in registers: once x[m] is known all later entries take their update at
            // once (independent fmas), so the dependent chain is 16 x (mul + fma), not 120 fmas
            double* rowp = S + (c0 + 16 + tid) * PD_LD + c0;
            double x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = rowp[j];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                x[m] *= invd[c0 + m];
#pragma unroll
                for (int j = m + 1; j < 16; ++j) x[j] = fma(-x[m], S[(c0 + j) * PD_LD + c0 + m], x[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) rowp[j] = x[j];
        }
        grp_sync<G>();
        if (pb == 7) break;
        // C: trailing blocks (bi >= bj > pb) -= X_bi X_bj^T
        const int nb = 7 - pb;
        const int nblocks = nb * (nb + 1) / 2;
        if (warp == 0) {
            potrf_block_update(S, c0 + 16, c0 + 16, c0, lane);      // next diagonal block first ...
            __syncwarp();
            potrf16_warp(S, invd, scr, c0 + 16, lane, &s_fail);    // ... and factor it right away
        } else {
            for (int t = warp; t < nblocks; t += 7) {
                int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
                while (bi * (bi + 1) / 2 > t) --bi;
                const int bj = t - bi * (bi + 1) / 2;
                potrf_block_update(S, (pb + 1 + bi) * 16, (pb + 1 + bj) * 16, c0, lane);
            }
        }
    }

    // ---- conditioning of the block: max L_jj / min L_jj (a lower bound of cond(L_kk)) decides whether the panel solve
    //      multiplies by the explicit inverse (error ~ eps cond(L_kk)) or substitutes (see CholArgs::refine) ----------
    if (warp == 0) {
        double mx = 0.0, mn = 1.0e300;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double v = invd[lane + 32 * u];            // 1 / L_jj
            mx = fmax(mx, v);
            mn = fmin(mn, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) {
            s_refine = (g.refine != nullptr && !(mx <= g.refine_ratio * mn)) ? 1 : 0;
            if (g.refine != nullptr) g.refine[(size_t)mat * g.refine_blocks + (g.refine_blocks > 1 ? k : 0)] = s_refine;
        }
    }

    double* Linv = g.Linv + ((size_t)mat * g.linv_blocks + (g.linv_blocks > 1 ? k : 0)) * (TB * TB);
    // ---- inverse, step 1: the eight 16 x 16 diagonal inverses, warp w -> block w, lane c -> column c --
    {
        const int c0 = warp * 16;
        const int c = lane & 15;
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < 16; ++m) {                 // right-looking: short dependent chain
            x[m] = (m >= c) ? x[m] * invd[c0 + m] : 0.0;
#pragma unroll
            for (int i = m + 1; i < 16; ++i) x[i] = fma(-S[(c0 + i) * PD_LD + c0 + m], x[m], x[i]);
        }
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) Dv[warp * 16 * PD_DV_LD + i * PD_DV_LD + c] = x[i];
        }
    }
    grp_sync<G>();
    // ---- inverse, step 2: warp j owns block column j of X = L^-1 and walks down the block rows:
    //      X_ij = -Dinv_i * sum_{m=j}^{i-1} L_im X_mj.  No block-wide barrier: a warp only reads
    //      the X blocks it produced itself.  Four accumulator sets keep the DMMA chains short. -------
    {
        const int j = warp;
        double* sc = scr + warp * 16 * PD_DV_LD;
#ifdef WHATIF_SKIP_INV2
        for (int i = 8; i < 8; ++i) {
#else
        for (int i = j + 1; i < 8; ++i) {
#endif
            double acc[4][2][2][2];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) acc[u][mi][ni][0] = acc[u][mi][ni][1] = 0.0;
            for (int m = j; m < i; ++m) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    double a[2], b[2];
                    const int kk = 4 * k4 + (lane & 3);
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) a[mi] = S[(16 * i + 8 * mi + (lane >> 2)) * PD_LD + 16 * m + kk];
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) {
                        const int n = 8 * ni + (lane >> 2);
                        b[ni] = (m == j) ? Dv[j * 16 * PD_DV_LD + kk * PD_DV_LD + n]      // X_jj = Dinv_j [k][n]
                                         : S[(16 * j + n) * PD_LD + 16 * m + kk];       // X_mj[k][n], stored transposed
                    }
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) dmma884(acc[k4][mi][ni][0], acc[k4][mi][ni][1], a[mi], b[ni]);
                }
            }
            // P -> per-warp scratch (as a B operand: [k][n]), then Y = -Dinv_i P
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const int rr = 8 * mi + (lane >> 2), cc = 8 * ni + 2 * (lane & 3);
                    sc[rr * PD_DV_LD + cc] = (acc[0][mi][ni][0] + acc[1][mi][ni][0]) + (acc[2][mi][ni][0] + acc[3][mi][ni][0]);
                    sc[rr * PD_DV_LD + cc + 1] = (acc[0][mi][ni][1] + acc[1][mi][ni][1]) + (acc[2][mi][ni][1] + acc[3][mi][ni][1]);
                }
            __syncwarp();
            double y[2][2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) y[mi][ni][0] = y[mi][ni][1] = 0.0;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                double a[2], b[2];
                const int kk = 4 * k4 + (lane & 3);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) a[mi] = -Dv[i * 16 * PD_DV_LD + (8 * mi + (lane >> 2)) * PD_DV_LD + kk];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) b[ni] = sc[kk * PD_DV_LD + 8 * ni + (lane >> 2)];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) dmma884(y[mi][ni][0], y[mi][ni][1], a[mi], b[ni]);
            }
            // X_ij[r][c] -> transposed into the upper triangle: S[16 j + c][16 i + r] (operand of the later block rows),
            // and straight to global memory from the accumulator layout (16-byte stores, 64 contiguous bytes per row):
            // the write-out loop below then has no bank-conflicted transposed reads left
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const int rr = 8 * mi + (lane >> 2), cc = 8 * ni + 2 * (lane & 3);
                    S[(16 * j + cc) * PD_LD + 16 * i + rr] = y[mi][ni][0];
                    S[(16 * j + cc + 1) * PD_LD + 16 * i + rr] = y[mi][ni][1];
                    *reinterpret_cast<double2*>(Linv + (16 * i + rr) * TB + 16 * j + cc) = make_double2(y[mi][ni][0], y[mi][ni][1]);
                }
            __syncwarp();
        }
    }
    grp_sync<G>();

    // ---- write L back (lower triangle incl. diagonal; the strict upper triangle of S holds the transposed inverse by now).
    //      On the likelihood path nobody reads the factored diagonal block again (log-det and z are taken here) unless the
    //      panel solve has to substitute with it: store_diag = 0 leaves it out then -----------------------------------------
    if (g.store_diag || s_refine) {
        for (int i = tid >> 6; i < TB; i += 4) {           // 16-byte stores; the element right of the diagonal is a zero
            const int c = 2 * (tid & 63);
            if (c <= i) {
                double2 v = *reinterpret_cast<const double2*>(S + i * PD_LD + c);
                if (c + 1 > i) v.y = 0.0;
                *reinterpret_cast<double2*>(Akk + (size_t)i * ld + c) = v;
            }
        }
    }

    // ---- rest of the inverse: the 16 x 16 diagonal blocks as full squares (the blocks below were stored by step 2).  The blocks
    //      ABOVE the diagonal blocks are zero and nobody ever writes anything else there: every Linv buffer is cleared once when
    //      it is allocated (agpn_api.cu), not 64 KB of zeros per factorised block (15 % of this kernel's stall samples) ----------
    for (int e = tid; e < 8 * 16 * 16; e += 256) {
        const int b = e >> 8, rr = (e >> 4) & 15, cc = e & 15;
        Linv[(16 * b + rr) * TB + 16 * b + cc] = Dv[b * 16 * PD_DV_LD + rr * PD_DV_LD + cc];
    }

    // ---- z_k = Linv r_k ; |z_k|^2 ; sum log L_jj ------------------------------------------------------
    if (rhs != nullptr && s_refine) {
        // ill-conditioned block: forward substitution on L itself (lower triangle of S), column by column
        for (int m = 0; m < TB; ++m) {
            if (tid == m) zs[m] = zs[m] / S[m * PD_LD + m];
            grp_sync<G>();
            if (tid > m && tid < TB) zs[tid] = fma(-S[tid * PD_LD + m], zs[m], zs[tid]);
            grp_sync<G>();
        }
        if (tid < TB) {
            zout[tid] = zs[tid];
            rhs[tid] = zs[tid];
        }
        grp_sync<G>();
    } else if (rhs != nullptr) {
        if (tid < TB) {
            const int R = tid;
            const int b0 = (R >> 4) * 16;
            double sacc = 0.0, asum = 0.0;                       // asum: the row's absolute sum (inf-norm of the inverse, for trsm's skip test)
            for (int C = 0; C < b0; ++C) {
                const double v = S[C * PD_LD + R];
                sacc = fma(v, zs[C], sacc);
                asum += fabs(v);
            }
            for (int C = b0; C <= R; ++C) {
                const double v = Dv[(R >> 4) * 16 * PD_DV_LD + (R & 15) * PD_DV_LD + (C - b0)];
                sacc = fma(v, zs[C], sacc);
                asum += fabs(v);
            }
            zout[R] = sacc;
            rhs[R] = sacc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) asum = fmax(asum, __shfl_xor_sync(0xffffffffu, asum, o));
            if ((tid & 31) == 0) s_ln[tid >> 5] = asum;
        }
        grp_sync<G>();
    }
    if (g.linvnorm != nullptr && tid < 2) {
        // inf-norm of the two 64-row halves of the inverse; infinity (never skip) when it was not formed above
        const bool have = rhs != nullptr && !s_refine;
        g.linvnorm[((size_t)mat * g.linv_blocks + (g.linv_blocks > 1 ? k : 0)) * 2 + tid] =
            have ? fmax(s_ln[2 * tid], s_ln[2 * tid + 1]) : __longlong_as_double(0x7ff0000000000000ll);
    }
    if (tid < 32) {
        double qv = 0.0, lv = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = tid + 32 * u;
            if (rhs != nullptr) {
                const double z = zout[j];
                qv = fma(z, z, qv);
            }
            lv += log(S[j * PD_LD + j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            qv += __shfl_xor_sync(0xffffffffu, qv, o);
            lv += __shfl_xor_sync(0xffffffffu, lv, o);
        }
        if (tid == 0) {
            if (rhs != nullptr) g.quad[mat] += qv;
            g.logdet[mat] += lv;
            if (s_fail && g.info[mat] == 0) g.info[mat] = 1;
        }
    }
}

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(CholArgs g, int k) {
    extern __shared__ __align__(16) double smem[];
    potrf_diag_body(g, k, (int)blockIdx.x, smem);
}
