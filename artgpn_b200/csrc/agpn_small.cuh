// K0 -- fused small-N likelihood kernel: for N <= 192 the whole per-(walker,
// series) work of network.log_likelihood (arttwo.py:271-287) -- covariance block assembly, Cholesky,
// forward solve, log-det -- happens inside ONE CTA's shared memory; nothing but the residual is read
// from HBM and three scalars are written.  CoRoT-7 shape (N = 167, 3 series, 256 walkers) = 768 CTAs.
//
// Storage: the lower block triangle of K in 16 x 16 blocks, each [16][20] doubles (stride 20 makes the
// DMMA fragment loads bank-conflict free); block (bi, bj) at index bi (bi + 1) / 2 + bj.  N is padded
// to a multiple of 16 with the identity.  Factorisation = the panel scheme of potrf_diag_kernel
// (serial 16 x 16 block by warp 0 with look-ahead, one thread per row for the panel, rank-16 updates
// on DMMA.8x8x4); the residual rides along as one extra row, so z = L^-1 r needs no separate solve.
#pragma once
#include "agpn_assemble.cuh"
#include "agpn_chol.cuh"

#define SM_BLK 320                 // doubles per 16 x 16 block (16 rows x stride 20)
#define SM_MAXNB 12                // N <= 192

struct SmallArgs {
    const double* theta;   // [W][D]
    const double* t;       // [N]
    const double* yerr;    // [p][N]
    const double* rhs;     // residuals [(w * p + s)][rhs_stride]
    int rhs_stride;
    int N, nb;             // nb = ceil(N / 16)
    int fast;              // exp-family fast formula usable (same rule as assemble_fast_kernel)
    int exact;             // reference operation order (kern_value_exact); implies !fast
    double torigin;
    double* logdet;        // [W * p]
    double* quad;          // [W * p]
    int* info;             // [W * p]
    int* kflag;            // [W * p]
};

__host__ __device__ inline size_t small_smem_bytes(int nb, int q) {
    const size_t nblocks = (size_t)nb * (nb + 1) / 2;
    return (nblocks * SM_BLK + 3 * (size_t)nb * 16 + 2 * (size_t)q * nb * 16 + 64) * sizeof(double);
}

__device__ __forceinline__ double* sm_blk(double* Sb, int bi, int bj) { return Sb + (bi * (bi + 1) / 2 + bj) * SM_BLK; }

// 16 x 16 diagonal block D ([16][20]) factorised in place by one warp (see potrf16_warp)
__device__ __forceinline__ void potrf16_warp_p(double* D, double* invd16, double* ub, int lane, int* s_fail) {
    const int r = lane & 15;
    double d[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) d[c] = D[r * 20 + c];
    double myrs = 1.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        double* u = ub + (j & 1) * 16;
        u[r] = d[j];
        __syncwarp();
        double piv = u[j];
        if (!(piv > 0.0) || isinf(piv)) {
            bad = 1;
            piv = 1.0;
        }
        const double rs = rsqrt(piv);
        const double t = d[j] * (rs * rs);
#pragma unroll
        for (int c = j + 1; c < 16; ++c) d[c] = fma(-t, u[c], d[c]);
        d[j] = (j == r) ? piv * rs : d[j] * rs;
        if (j == r) myrs = rs;
    }
    if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c <= r) D[r * 20 + c] = d[c];
        invd16[r] = myrs;
    }
    if (bad && lane == 0) *s_fail = 1;
}

// C ([16][20]) -= Xi Xj^T, Xi / Xj = panel blocks [16][20], one warp
__device__ __forceinline__ void block_update_p(double* C, const double* Xi, const double* Xj, int lane) {
    double acc[2][2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const double2 c = *reinterpret_cast<const double2*>(C + (8 * mi + (lane >> 2)) * 20 + 8 * ni + 2 * (lane & 3));
            acc[mi][ni][0] = c.x;
            acc[mi][ni][1] = c.y;
        }
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        double a[2], b[2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) a[mi] = -Xi[(8 * mi + (lane >> 2)) * 20 + 4 * k4 + (lane & 3)];
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) b[ni] = Xj[(8 * ni + (lane >> 2)) * 20 + 4 * k4 + (lane & 3)];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
            *reinterpret_cast<double2*>(C + (8 * mi + (lane >> 2)) * 20 + 8 * ni + 2 * (lane & 3)) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// grid (p, W), 256 threads, dynamic shared memory small_smem_bytes(nb, q)
__global__ void __launch_bounds__(256) small_fused_kernel(const __grid_constant__ Structure S, const SmallArgs g) {
    extern __shared__ __align__(16) double smem[];
    __shared__ PKern s_node[AGPN_MAXQ];
    __shared__ PKern s_weight[AGPN_MAXQ];
    __shared__ double s_fa[AGPN_MAXQ][3];        // fast path: a_j, c_j + cw_j, amp
    __shared__ RKern s_rnode[AGPN_MAXQ], s_rweight[AGPN_MAXQ];
    __shared__ int s_fail, s_bad;
    __shared__ __align__(128) double s_etab[16];   // 2^(j/16) for exp_nonpos

    const int series = blockIdx.x, w = blockIdx.y;
    const int mat = w * S.p + series;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = S.q, nb = g.nb, N = g.N, np = nb * 16;
    const int nblocks = nb * (nb + 1) / 2;
    double* Sb = smem;
    double* invd = Sb + (size_t)nblocks * SM_BLK;     // [np]
    double* rr = invd + np;                           // [np] residual -> z
    double* tt = rr + np;                             // [np] times
    double* trig = tt + np;                           // [q][2][np]
    double* scr = trig + 2 * q * np;                  // [64]
    const double* th = g.theta + (size_t)w * S.D;

    if (tid == 0) { s_fail = 0; s_bad = 0; }
    if (tid >= 240) s_etab[tid - 240] = c_exp2_16[tid - 240];
    if (tid < q) {
        const PKern kn = prep_kernel(0, S.node_kind[tid], th + S.node_off[tid]);
        const int widx = tid + q * series;
        const PKern kw = prep_kernel(1, S.weight_kind[widx], th + S.weight_off[widx]);
        s_node[tid] = kn;
        s_weight[tid] = kw;
        s_rnode[tid] = prep_raw(0, S.node_kind[tid], th + S.node_off[tid]);
        s_rweight[tid] = prep_raw(1, S.weight_kind[widx], th + S.weight_off[widx]);
        s_fa[tid][0] = (kn.kind == K_SE) ? 0.0 : kn.a;
        s_fa[tid][1] = ((kn.kind == K_SE) ? kn.a : kn.c) + ((kw.kind == K_SE) ? kw.a : 0.0);
        s_fa[tid][2] = kw.amp;
    }
    for (int n = tid; n < np; n += 256) {
        tt[n] = g.t[n < N ? n : N - 1];
        rr[n] = n < N ? g.rhs[(size_t)mat * g.rhs_stride + n] : 0.0;
    }
    __syncthreads();
    if (g.fast) {
        for (int u = tid; u < q * np; u += 256) {
            const int j = u / np, n = u % np;
            double sv = 0.0, cv = 1.0;
            if (s_node[j].kind != K_SE) phase_sincos(tt[n] - g.torigin, s_node[j].per, &sv, &cv);
            trig[(2 * j) * np + n] = sv;
            trig[(2 * j + 1) * np + n] = cv;
        }
        __syncthreads();
    }

    // ---- assembly: one entry per thread per block (arttwo.py:271-281) ------------------------------
    {
        const int r = tid >> 4, c = tid & 15;
        const double jt = th[S.jit_off + series];
        int bad = 0;
        int bi = 0, bj = 0;
        for (int t = 0; t < nblocks; ++t) {
            const int gi = 16 * bi + r, gj = 16 * bj + c;
            double v;
            if (gi < N && gj < N) {
                const double t1 = tt[gi], t2 = tt[gj];
                const double rl = t1 - t2;
                v = 0.0;
                if (g.fast) {
                    const double r2 = rl * rl;
                    for (int j = 0; j < q; ++j) {
                        const double sn = fma(trig[(2 * j) * np + gi], trig[(2 * j + 1) * np + gj],
                                              -trig[(2 * j + 1) * np + gi] * trig[(2 * j) * np + gj]);
                        v = fma(s_fa[j][2], exp_nonpos(fma(s_fa[j][1], r2, s_fa[j][0] * sn * sn), s_etab), v);
                    }
                } else if (g.exact) {
                    const bool dg = gi == gj;
                    for (int j = 0; j < q; ++j)
                        v = __dadd_rn(v, __dmul_rn(kern_value_exact(s_rweight[j], rl, t1, t2, dg),
                                                   kern_value_exact(s_rnode[j], rl, t1, t2, dg)));
                } else {
                    const bool dg = gi == gj;
                    for (int j = 0; j < q; ++j) {
                        const PKern& kn = s_node[j];
                        const PKern& kw = s_weight[j];
                        const double f = kn.kind == K_CONSTANT ? kn.amp : kn.amp * kern_shape(kn, rl, t1, t2, dg);
                        const double wv = kw.kind == K_CONSTANT ? kw.amp : kw.amp * kern_shape(kw, rl, t1, t2, dg);
                        v = fma(wv, f, v);
                    }
                }
                if (gi == gj) {
                    const double e = g.yerr[(size_t)series * N + gi];
                    v = __dadd_rn(v, __dadd_rn(__dmul_rn(e, e), __dmul_rn(jt, jt)));     // k += (err^2 + jit^2), arttwo.py:280-281
                }
                if (!isfinite(v)) bad = 1;
            } else {
                v = (gi == gj) ? 1.0 : 0.0;
            }
            Sb[(size_t)t * SM_BLK + r * 20 + c] = v;
            if (++bj > bi) { bj = 0; ++bi; }
        }
        if (bad) s_bad = 1;
    }
    __syncthreads();

    // ---- factorisation with the residual as an extra row ---------------------------------------------
    if (warp == 0) potrf16_warp_p(sm_blk(Sb, 0, 0), invd, scr, lane, &s_fail);
    for (int pb = 0; pb < nb; ++pb) {
        __syncthreads();
        const double* D = sm_blk(Sb, pb, pb);
        const int nrows = (nb - pb - 1) * 16;
        if (tid <= nrows) {                                    // thread nrows carries the residual row
            double* rowp = (tid < nrows) ? sm_blk(Sb, pb + 1 + (tid >> 4), pb) + (tid & 15) * 20 : rr + pb * 16;
            double x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = rowp[j];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                x[m] *= invd[pb * 16 + m];
#pragma unroll
                for (int j = m + 1; j < 16; ++j) x[j] = fma(-x[m], D[j * 20 + m], x[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) rowp[j] = x[j];
        }
        __syncthreads();
        if (pb == nb - 1) break;
        const int nt = nb - pb - 1;
        const int ntr = nt * (nt + 1) / 2;
        if (warp == 0) {
            block_update_p(sm_blk(Sb, pb + 1, pb + 1), sm_blk(Sb, pb + 1, pb), sm_blk(Sb, pb + 1, pb), lane);
            __syncwarp();
            potrf16_warp_p(sm_blk(Sb, pb + 1, pb + 1), invd + (pb + 1) * 16, scr, lane, &s_fail);
        } else {
            for (int t = warp; t < ntr; t += 7) {
                int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
                while (bi * (bi + 1) / 2 > t) --bi;
                const int bj = t - bi * (bi + 1) / 2;
                block_update_p(sm_blk(Sb, pb + 1 + bi, pb + 1 + bj), sm_blk(Sb, pb + 1 + bi, pb), sm_blk(Sb, pb + 1 + bj, pb), lane);
            }
            // residual row: r[cols of later blocks] -= z_panel . X^T   (threads of warps 1..7)
            const int u = tid - 32;
            if (u < nt * 16) {
                const double* X = sm_blk(Sb, pb + 1 + (u >> 4), pb) + (u & 15) * 20;
                double sacc = rr[(pb + 1) * 16 + u];
#pragma unroll
                for (int m = 0; m < 16; ++m) sacc = fma(-rr[pb * 16 + m], X[m], sacc);
                rr[(pb + 1) * 16 + u] = sacc;
            }
        }
    }
    // ---- log-det, |z|^2 (fixed order) ------------------------------------------------------------------
    if (tid < 32) {
        double qv = 0.0, lv = 0.0;
        for (int n = tid; n < np; n += 32) {
            const double z = rr[n];
            qv = fma(z, z, qv);
            lv += log(sm_blk(Sb, n >> 4, n >> 4)[(n & 15) * 21]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            qv += __shfl_xor_sync(0xffffffffu, qv, o);
            lv += __shfl_xor_sync(0xffffffffu, lv, o);
        }
        if (tid == 0) {
            g.quad[mat] = qv;
            g.logdet[mat] = lv;
            g.info[mat] = s_fail;
            g.kflag[mat] = s_bad;
        }
    }
}
