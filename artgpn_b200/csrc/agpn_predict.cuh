// Predictive mean / variance (sm_100a): replaces network.prediction (arttwo.py:367-396).
//
// The reference builds K*[M][N], K**[M][M] and M triangular-solve pairs and keeps only the
// diagonal.  The same numbers are
//     V = K* L^-T  (rows = prediction points),  mu = V z + m(t*),  z = L^-1 r,
//     var = k** + s^2 - rowsum(V o V),          k**_m = sum_j W_dj(t*_m, t*_m) F_j(t*_m, t*_m).
// predict_kernel: one CTA owns 128 prediction points and sweeps the block columns k of L
// left-looking:  T = K*[rows, k] - sum_{j<k} V[rows, j] L[k, j]^T   (K* tile generated in
// registers, never stored; the sum is one NT GEMM of depth 128 k on DMMA.8x8x4)
//                V[rows, k] = T Linv_kk^T                            (second NT GEMM)
// and folds the dot products with z and the row sums of squares into the same pass.
#pragma once
#include "agpn_chol.cuh"
#include "agpn_eval.cuh"

struct PredArgs {
    const double* theta;     // [D] single row
    const double* tstar;     // [M]
    const double* ttrain;    // [N]
    int M, N, n_pad, nblk;
    int series;
    int square;              // whole prediction grid has N points (WhiteNoise quirk inside K*)
    int row_off = 0;         // index of this call's first prediction point inside the whole grid
    int exact;               // reference operation order for the K* entries
    const double* L;         // factor [n_pad][ld]
    int ld;
    const double* Linv;      // [nblk][128][128]
    const double* z;         // [n_pad]
    double* V;               // workspace [ceil(M/128)*128][n_pad]
    const double* mstar;     // [M] mean at t*
    const double* kss;       // [M] diag of K**
    double* mean_out;        // [M]
    double* std_out;         // [M]
};

__global__ void __launch_bounds__(256, 2) predict_kernel(const __grid_constant__ Structure S, const PredArgs g) {
    extern __shared__ __align__(16) double smem[];
    __shared__ PKern s_node[AGPN_MAXQ];
    __shared__ PKern s_weight[AGPN_MAXQ];
    __shared__ RKern s_rnode[AGPN_MAXQ], s_rweight[AGPN_MAXQ];
    __shared__ double s_ts[TB], s_tn[TB], s_z[TB];
    __shared__ double s_red[2][2][TB];
    __shared__ __align__(8) unsigned long long bars[2 * GEMM_STAGES];
    GemmPipe pipe = gemm_pipe_init(bars);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int q = S.q;
    const int m0 = blockIdx.x * TB;
    if (tid < q) {
        s_node[tid] = prep_kernel(0, S.node_kind[tid], g.theta + S.node_off[tid]);
        s_rnode[tid] = prep_raw(0, S.node_kind[tid], g.theta + S.node_off[tid]);
    } else if (tid >= 32 && tid < 32 + q) {
        const int widx = (tid - 32) + q * g.series;
        s_weight[tid - 32] = prep_kernel(1, S.weight_kind[widx], g.theta + S.weight_off[widx]);
        s_rweight[tid - 32] = prep_raw(1, S.weight_kind[widx], g.theta + S.weight_off[widx]);
    }
    if (tid < TB) {
        const int gm = m0 + tid;
        s_ts[tid] = g.tstar[gm < g.M ? gm : g.M - 1];
    }
    double* Vrow = g.V + (size_t)m0 * g.n_pad;
    double mu_acc = 0.0, v_acc = 0.0;     // threads < 128: one prediction point each
    double acc[4][4][2];

    for (int k = 0; k < g.nblk; ++k) {
        __syncthreads();
        if (tid < TB) {
            const int gn = k * TB + tid;
            s_tn[tid] = g.ttrain[gn < g.N ? gn : g.N - 1];
            s_z[tid] = g.z[gn];
        }
        __syncthreads();
        double* Vt = Vrow + (size_t)k * TB;
        // ---- T[:, half] = K*[rows, k] - V[rows, 0:k] L[k, 0:k]^T, two 64-column halves -------------
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int row = wm * 32 + mi * 8 + (lane >> 2);
                const int gm = m0 + row;
                const double t1 = s_ts[row];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = h * BN + wn * 32 + ni * 8 + 2 * (lane & 3) + e;
                        const int gn = k * TB + col;
                        double v = 0.0;
                        if (gm < g.M && gn < g.N) {
                            const double t2 = s_tn[col];
                            const double r = t1 - t2;
                            const bool dg = g.square && gm + g.row_off == gn;
                            for (int j = 0; j < q; ++j) {
                                if (g.exact) {
                                    v = __dadd_rn(v, __dmul_rn(kern_value_exact(s_rweight[j], r, t1, t2, dg),
                                                               kern_value_exact(s_rnode[j], r, t1, t2, dg)));
                                    continue;
                                }
                                const PKern& kn = s_node[j];
                                const PKern& kw = s_weight[j];
                                const double f = kn.kind == K_CONSTANT ? kn.amp : kn.amp * kern_shape(kn, r, t1, t2, dg);
                                const double wv = kw.kind == K_CONSTANT ? kw.amp : kw.amp * kern_shape(kw, r, t1, t2, dg);
                                v = fma(wv, f, v);
                            }
                        }
                        acc[mi][ni][e] = -v;
                    }
                }
            }
            if (k > 0) gemm_nt_tile(pipe, Vrow, g.n_pad, g.L + ((size_t)k * TB + h * BN) * g.ld, g.ld, k * TB, acc, smem);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int row = wm * 32 + mi * 8 + (lane >> 2);
                    const int col = h * BN + wn * 32 + ni * 8 + 2 * (lane & 3);
                    *reinterpret_cast<double2*>(Vt + (size_t)row * g.n_pad + col) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
                }
        }
        __syncthreads();
        // ---- V[rows, k] = T Linv_kk^T : right half first (reads all of T), left half K = 64 -------
        double pm[4] = {0.0, 0.0, 0.0, 0.0}, pv[4] = {0.0, 0.0, 0.0, 0.0};
        for (int pass = 0; pass < 2; ++pass) {
            const int h = 1 - pass;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            gemm_nt_tile(pipe, Vt, g.n_pad, g.Linv + (size_t)k * TB * TB + (size_t)h * BN * TB, TB, (h + 1) * BN, acc, smem);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int row = wm * 32 + mi * 8 + (lane >> 2);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int col = h * BN + wn * 32 + ni * 8 + 2 * (lane & 3);
                    const double x0 = acc[mi][ni][0], x1 = acc[mi][ni][1];
                    *reinterpret_cast<double2*>(Vt + (size_t)row * g.n_pad + col) = make_double2(x0, x1);
                    pm[mi] = fma(x0, s_z[col], pm[mi]);
                    pm[mi] = fma(x1, s_z[col + 1], pm[mi]);
                    pv[mi] = fma(x0, x0, pv[mi]);
                    pv[mi] = fma(x1, x1, pv[mi]);
                }
            }
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            double sm = pm[mi], sv = pv[mi];
            sm += __shfl_xor_sync(0xffffffffu, sm, 1);
            sm += __shfl_xor_sync(0xffffffffu, sm, 2);
            sv += __shfl_xor_sync(0xffffffffu, sv, 1);
            sv += __shfl_xor_sync(0xffffffffu, sv, 2);
            if ((lane & 3) == 0) {
                const int row = wm * 32 + mi * 8 + (lane >> 2);
                s_red[0][wn][row] = sm;
                s_red[1][wn][row] = sv;
            }
        }
        __syncthreads();
        if (tid < TB) {
            mu_acc += s_red[0][0][tid] + s_red[0][1][tid];
            v_acc += s_red[1][0][tid] + s_red[1][1][tid];
        }
    }
    if (tid < TB) {
        const int gm = m0 + tid;
        if (gm < g.M) {
            const double jt = g.theta[S.jit_off + g.series];
            g.mean_out[gm] = mu_acc + g.mstar[gm];
            g.std_out[gm] = sqrt(g.kss[gm] + jt * jt - v_acc);
        }
    }
}

// diag of K** (arttwo.py:385-387): every kernel at zero lag on the (square) M x M grid diagonal
__global__ void kss_kernel(const __grid_constant__ Structure S, const double* theta, int series,
                           const double* tstar, int M, double* kss) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double t = tstar[m];
    double v = 0.0;
    for (int j = 0; j < S.q; ++j) {
        const int widx = j + S.q * series;
        const PKern kn = prep_kernel(0, S.node_kind[j], theta + S.node_off[j]);
        const PKern kw = prep_kernel(1, S.weight_kind[widx], theta + S.weight_off[widx]);
        v += (kw.amp * kern_shape(kw, 0.0, t, t, true)) * (kn.amp * kern_shape(kn, 0.0, t, t, true));
    }
    kss[m] = v;
}
