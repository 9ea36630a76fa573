// Batched blocked FP64 Cholesky for sm_100a: replaces scipy cho_factor / cho_solve of
// artgpn (arttwo.py:284-287, :371-372) for nmat independent matrices at once.
//
// Layout: each matrix is row-major [n_pad][ld], n_pad = nblk * 128; only tiles on or below the
// block diagonal are referenced ("lower": K = L L^T, L[i][j] at i*ld + j; the reference's
// upper factor U of cho_factor(lower=False) is L^T, same diagonal).  Padding rows carry the
// identity, so the factor of the padded matrix is [L 0; 0 I] and log-det / solves are unchanged.
//
// Right-looking tile algorithm, one step per 128-wide block column k, three kernels per step
// over ALL matrices of the batch (grid.y = matrix):
//   potrf_diag_kernel   one CTA per matrix: 128x128 diagonal block factorised in registers
//                       (2-D cyclic thread ownership, one column broadcast per step through
//                       shared memory), its inverse by the same sweep, and -- folded in -- the
//                       forward-solve block z_k = L_kk^-1 r_k, |z_k|^2 and sum log L_jj
//                       (warp-shuffle reductions).
//   trsm_kernel         L_ik = A_ik * L_kk^-T as an NT GEMM with the inverted diagonal block
//                       on the FP64 tensor pipe (DMMA.8x8x4); also r_i -= L_ik z_k.
//   syrk_kernel         A_ij -= L_ik L_jk^T for all trailing tiles i >= j > k, DMMA.8x8x4,
//                       operands staged through a 4-stage cp.async shared-memory pipeline with an
//                       XOR swizzle that makes the 64-bit fragment loads bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TB 128                 // tile edge
#define GEMM_KC 16             // k extent of one pipeline stage
#define GEMM_STAGES 4
#define GEMM_STAGE_DOUBLES (2 * TB * GEMM_KC)
#define GEMM_SMEM_BYTES (GEMM_STAGES * GEMM_STAGE_DOUBLES * 8)   // 131072

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
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// swizzled offset (in doubles) of element (row, k) inside a [128][16] stage operand
__device__ __forceinline__ int swz(int row, int k) {
    return row * GEMM_KC + ((((k >> 1) ^ ((row & 3) << 1)) << 1) | (k & 1));
}

// ---------------------------------------------------------------------------------------------
// acc[mi][ni][2] += A[128 x K] * B[128 x K]^T  (both row-major, k contiguous), 256 threads,
// warp grid 2 (m) x 4 (n), warp tile 64 x 32 = 8 x 4 DMMA tiles.
// Accumulator element (mi, ni, e) of lane l lives at
//     row = wm*64 + mi*8 + (l >> 2),  col = wn*32 + ni*8 + 2*(l & 3) + e.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_load_stage(double* stage, const double* __restrict__ Ag, int lda,
                                                const double* __restrict__ Bg, int ldb, int kc, int tid) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        int idx = tid + 256 * u;
        int which = idx >> 10;
        int rem = idx & 1023;
        int row = rem >> 3;
        int ch = rem & 7;
        const double* src = which ? (Bg + (size_t)row * ldb + kc * GEMM_KC + ch * 2)
                                  : (Ag + (size_t)row * lda + kc * GEMM_KC + ch * 2);
        double* dst = stage + which * (TB * GEMM_KC) + row * GEMM_KC + ((ch ^ ((row & 3) << 1)) << 1);
        cp_async16(dst, src);
    }
}

__device__ __forceinline__ void gemm_compute_stage(const double* stage, double (&acc)[8][4][2], int wm, int wn, int lane) {
    const double* As = stage;
    const double* Bs = stage + TB * GEMM_KC;
    const int r = lane >> 2;
    const int kq = lane & 3;
#pragma unroll
    for (int ks = 0; ks < GEMM_KC / 4; ++ks) {
        double a[8], b[4];
        const int k = ks * 4 + kq;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) a[mi] = As[swz(wm * 64 + mi * 8 + r, k)];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[swz(wn * 32 + ni * 8 + r, k)];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// K must be a multiple of GEMM_KC.  All 256 threads must call.  On return every cp.async issued
// here has completed and been consumed (safe to overwrite global A / reuse smem after a barrier).
__device__ __forceinline__ void gemm_nt_tile(const double* __restrict__ Ag, int lda, const double* __restrict__ Bg, int ldb,
                                             int K, double (&acc)[8][4][2], double* smem) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp >> 2;
    const int wn = warp & 3;
    const int nch = K / GEMM_KC;
#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < nch) gemm_load_stage(smem + s * GEMM_STAGE_DOUBLES, Ag, lda, Bg, ldb, s, tid);
        cp_async_commit();
    }
    for (int kc = 0; kc < nch; ++kc) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        const int nxt = kc + GEMM_STAGES - 1;
        if (nxt < nch) gemm_load_stage(smem + (nxt % GEMM_STAGES) * GEMM_STAGE_DOUBLES, Ag, lda, Bg, ldb, nxt, tid);
        cp_async_commit();
        gemm_compute_stage(smem + (kc % GEMM_STAGES) * GEMM_STAGE_DOUBLES, acc, wm, wn, lane);
    }
    cp_async_wait<0>();
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// trailing update: A_ij -= L_ik L_jk^T, i >= j > k
// grid.x = nt*(nt+1)/2 with nt = nblk-k-1, grid.y = nmat
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) syrk_kernel(CholArgs g, int k) {
    extern __shared__ __align__(16) double smem[];
    const int t = blockIdx.x;
    int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    const int bj = t - bi * (bi + 1) / 2;
    const int ti = k + 1 + bi, tj = k + 1 + bj;
    double* A = g.A + (size_t)blockIdx.y * g.mat_stride;
    const int ld = g.ld;
    const double* Ai = A + (size_t)ti * TB * ld + (size_t)k * TB;
    const double* Aj = A + (size_t)tj * TB * ld + (size_t)k * TB;
    double* C = A + (size_t)ti * TB * ld + (size_t)tj * TB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    double acc[8][4][2];
    // acc = -C, so that -(acc + A B^T) = C - A B^T
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int row = wm * 64 + mi * 8 + (lane >> 2);
            const int col = wn * 32 + ni * 8 + 2 * (lane & 3);
            const double2 c = *reinterpret_cast<const double2*>(C + (size_t)row * ld + col);
            acc[mi][ni][0] = -c.x;
            acc[mi][ni][1] = -c.y;
        }
    gemm_nt_tile(Ai, ld, Aj, ld, TB, acc, smem);
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int row = wm * 64 + mi * 8 + (lane >> 2);
            const int col = wn * 32 + ni * 8 + 2 * (lane & 3);
            double2 c;
            c.x = -acc[mi][ni][0];
            c.y = -acc[mi][ni][1];
            *reinterpret_cast<double2*>(C + (size_t)row * ld + col) = c;
        }
}

// ---------------------------------------------------------------------------------------------
// panel solve: A_ik <- A_ik * Linv_kk^T (i > k), and rhs_i -= L_ik z_k
// grid.x = nblk-k-1, grid.y = nmat
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) trsm_kernel(CholArgs g, int k) {
    extern __shared__ __align__(16) double smem[];
    const int ti = k + 1 + blockIdx.x;
    const int mat = blockIdx.y;
    double* A = g.A + (size_t)mat * g.mat_stride;
    const int ld = g.ld;
    double* Aik = A + (size_t)ti * TB * ld + (size_t)k * TB;
    const double* Linv = g.Linv + ((size_t)mat * g.linv_blocks + (g.linv_blocks > 1 ? k : 0)) * (TB * TB);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    gemm_nt_tile(Aik, ld, Linv, TB, TB, acc, smem);
    // gemm_nt_tile ended with a barrier after all loads completed: in-place store is safe
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int row = wm * 64 + mi * 8 + (lane >> 2);
            const int col = wn * 32 + ni * 8 + 2 * (lane & 3);
            double2 c;
            c.x = acc[mi][ni][0];
            c.y = acc[mi][ni][1];
            *reinterpret_cast<double2*>(Aik + (size_t)row * ld + col) = c;
        }
    if (g.rhs != nullptr) {
        double* zs = smem;              // [128]   z_k
        double* red = smem + TB;        // [4][128]
        double* rhs = g.rhs + (size_t)mat * ((size_t)g.nblk * TB);
        if (tid < TB) zs[tid] = rhs[(size_t)k * TB + tid];
        __syncthreads();
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            double s = 0.0;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int col = wn * 32 + ni * 8 + 2 * (lane & 3);
                s = fma(acc[mi][ni][0], zs[col], s);
                s = fma(acc[mi][ni][1], zs[col + 1], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if ((lane & 3) == 0) red[wn * TB + wm * 64 + mi * 8 + (lane >> 2)] = s;
        }
        __syncthreads();
        if (tid < TB) {
            const double s = ((red[tid] + red[TB + tid]) + red[2 * TB + tid]) + red[3 * TB + tid];
            rhs[(size_t)ti * TB + tid] -= s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// diagonal block: factor, invert, forward-solve block, log-det
// grid.x = nmat, 256 threads as a 16 x 16 grid; thread (ti, tc) owns entries
// (ti + 16 a, tc + 16 b), a, b in 0..7, of the 128 x 128 block in registers.
// ---------------------------------------------------------------------------------------------
#define LS_LD 129
#define POTRF_SMEM_BYTES ((TB * LS_LD + 2 * TB + 3 * TB) * 8)

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(CholArgs g, int k) {
    extern __shared__ __align__(16) double smem[];
    double* Ls = smem;                       // [128][129] scaled factor
    double* colbuf = Ls + TB * LS_LD;        // [2][128]
    double* piv = colbuf + 2 * TB;           // [128] pivots d_j
    double* rs = piv + TB;                   // [128] 1/sqrt(d_j)
    double* zs = rs + TB;                    // [128] rhs block / z
    __shared__ int s_fail;

    const int mat = blockIdx.x;
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tc = tid & 15;
    const int ld = g.ld;
    double* Akk = g.A + (size_t)mat * g.mat_stride + (size_t)k * TB * ld + (size_t)k * TB;
    if (tid == 0) s_fail = 0;

    double a[8][8];
#pragma unroll
    for (int aa = 0; aa < 8; ++aa)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const int i = ti + 16 * aa, c = tc + 16 * bb;
            a[aa][bb] = (i >= c) ? Akk[(size_t)i * ld + c] : 0.0;
        }

    // ---- factor: right-looking, column j broadcast unscaled, update uses 1/d_j -----------------
    for (int j = 0; j < TB; ++j) {
        double* cb = colbuf + (j & 1) * TB;
        const int bj = j >> 4;
        if (tc == (j & 15)) {
#pragma unroll
            for (int bb = 0; bb < 8; ++bb)
                if (bb == bj) {
#pragma unroll
                    for (int aa = 0; aa < 8; ++aa) cb[ti + 16 * aa] = a[aa][bb];
                }
        }
        __syncthreads();
        double d = cb[j];
        if (!(d > 0.0) || isinf(d)) {      // LAPACK dpotrf: ajj <= 0 or NaN -> info = j+1
            if (tid == 0) s_fail = 1;
            d = 1.0;
        }
        if (tid == 0) piv[j] = d;
        const double invd = 1.0 / d;
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const int c = tc + 16 * bb;
            if (c > j) {
                const double lc = cb[c] * invd;
#pragma unroll
                for (int aa = 0; aa < 8; ++aa) {
                    const int i = ti + 16 * aa;
                    if (i >= c) a[aa][bb] = fma(-cb[i], lc, a[aa][bb]);
                }
            }
        }
    }
    __syncthreads();
    if (tid < TB) rs[tid] = 1.0 / sqrt(piv[tid]);
    if (g.rhs != nullptr && tid < TB) zs[tid] = g.rhs[(size_t)mat * ((size_t)g.nblk * TB) + (size_t)k * TB + tid];
    __syncthreads();
    // scale columns, publish L to shared memory and write it back
#pragma unroll
    for (int aa = 0; aa < 8; ++aa)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const int i = ti + 16 * aa, c = tc + 16 * bb;
            double v = 0.0;
            if (i > c) v = a[aa][bb] * rs[c];
            else if (i == c) v = sqrt(piv[c]);
            Ls[i * LS_LD + c] = v;
            if (i >= c) Akk[(size_t)i * ld + c] = v;
        }

    // ---- inverse: X = L^-1 by the same sweep (row j of X final at step j) ----------------------
    double (&x)[8][8] = a;
#pragma unroll
    for (int aa = 0; aa < 8; ++aa)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) x[aa][bb] = (ti + 16 * aa == tc + 16 * bb) ? 1.0 : 0.0;
    for (int j = 0; j < TB; ++j) {
        double* rb = colbuf + (j & 1) * TB;
        const int aj = j >> 4;
        if (ti == (j & 15)) {
#pragma unroll
            for (int aa = 0; aa < 8; ++aa)
                if (aa == aj) {
#pragma unroll
                    for (int bb = 0; bb < 8; ++bb) {
                        const double v = x[aa][bb] * rs[j];
                        x[aa][bb] = v;
                        rb[tc + 16 * bb] = v;
                    }
                }
        }
        __syncthreads();   // also orders the Ls writes above before the first read below
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            const int i = ti + 16 * aa;
            if (i > j) {
                const double lij = Ls[i * LS_LD + j];
#pragma unroll
                for (int bb = 0; bb < 8; ++bb) {
                    const int c = tc + 16 * bb;
                    if (c <= j) x[aa][bb] = fma(-lij, rb[c], x[aa][bb]);
                }
            }
        }
    }
    // write the inverse (full tile, zeros above the diagonal)
    double* Linv = g.Linv + ((size_t)mat * g.linv_blocks + (g.linv_blocks > 1 ? k : 0)) * (TB * TB);
#pragma unroll
    for (int aa = 0; aa < 8; ++aa)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const int i = ti + 16 * aa, c = tc + 16 * bb;
            Linv[i * TB + c] = (i >= c) ? x[aa][bb] : 0.0;
        }

    // ---- z_k = X r_k ; |z_k|^2 ; sum log L_jj -------------------------------------------------
    __syncthreads();
    if (g.rhs != nullptr) {
        double* zout = colbuf;  // [128] reuse
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            double s = 0.0;
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) s = fma(x[aa][bb], zs[tc + 16 * bb], s);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (tc == 0) zout[ti + 16 * aa] = s;
        }
        __syncthreads();
        if (tid < TB) g.rhs[(size_t)mat * ((size_t)g.nblk * TB) + (size_t)k * TB + tid] = zout[tid];
        if (tid < 32) {
            double q = 0.0, l = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double z = zout[tid + 32 * u];
                q = fma(z, z, q);
                l += log(sqrt(piv[tid + 32 * u]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                q += __shfl_xor_sync(0xffffffffu, q, o);
                l += __shfl_xor_sync(0xffffffffu, l, o);
            }
            if (tid == 0) {
                g.quad[mat] += q;
                g.logdet[mat] += l;
                if (s_fail && g.info[mat] == 0) g.info[mat] = 1;
            }
        }
    } else if (tid < 32) {
        double l = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) l += log(sqrt(piv[tid + 32 * u]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if (tid == 0) {
            g.logdet[mat] += l;
            if (s_fail && g.info[mat] == 0) g.info[mat] = 1;
        }
    }
}
