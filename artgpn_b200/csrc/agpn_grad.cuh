// Analytic gradient of the marginal log-likelihood (the reference only carries the
// -- partly broken -- derivative kernel classes of weight.py:81-931 and never assembles a gradient).  Re-derived here:
//     LL = sum_i  -1/2 r_i^T K_i^-1 r_i - 1/2 log det K_i - N/2 log 2 pi,   K_i = sum_j W_ij o F_j + diag(sigma_i^2 + s_i^2)
//     dLL / d theta   = 1/2 sum_i sum_ab G_i[a][b] dK_i[a][b] / d theta,      G_i = alpha_i alpha_i^T - K_i^-1,  alpha_i = K_i^-1 r_i
//     dLL / d phi     = sum_a alpha_i[a] dm_i(t_a) / d phi                    (phi a parameter of the mean of series i)
// K_i^-1 and alpha_i come from the factorisation kernels themselves: eliminating the first N columns of [[K_i, .], [I, 0]] with
// right-hand side [r_i; 0] leaves -K_i^-1 in the trailing block and -alpha_i in the trailing right-hand side (the same partial
// blocked Cholesky agpn_predict_cov uses).  grad_tile_kernel then contracts G with the derivative tiles, generated on
// the fly from closed forms; nothing of size N^2 per parameter is ever stored.
//
// Closed forms (r = t_a - t_b, x = pi |r| / P, e = the kernel's value):
//   Constant node c:            F = c                          dF/dc = 1
//   SquaredExponential(ell):    e = exp(-r^2 / (2 ell^2))       de/dell = e r^2 / ell^3
//   Periodic(P, ell):           e = exp(-2 sin^2 x / ell^2)     de/dP = e (2 pi |r| / (P^2 ell^2)) sin 2x     de/dell = e 4 sin^2 x / ell^3
//   QuasiPeriodic(ell_e, P, ell_p): product of the two above (ell -> ell_e, ell_p)
//   Cosine(P):                  e = cos(2x)                     de/dP = sin(2x) 2 pi |r| / P^2
//   Exponential(ell):           e = exp(-|r| / ell)             de/dell = e |r| / ell^2
//   Matern32(ell), a = sqrt3 |r| / ell:  e = (1 + a) exp(-a)    de/dell = a^2 exp(-a) / ell
//   Matern52(ell), a = sqrt5 |r| / ell:  e = (1 + a + a^2/3) exp(-a)   de/dell = a^2 (1 + a) exp(-a) / (3 ell)
// Weight kernels are weight^2 times the same shapes (weight.py), Constant weight c^2: d/dc = 2c.
// Means: Constant (dm/dc = 1), Linear(slope, intercept) (dm/dslope = t - mean(t), dm/dintercept = 1).
// Jitter s_i: dK/ds = 2 s I.  Other classes (WhiteNoise, RationalQuadratic, RQP, Linear, GammaExp, Polynomial kernels;
// Parabola, Cubic, Sine, Cosine, Keplerian means) are reported as unsupported by agpn_loglike_grad.
#pragma once
#include "agpn_eval.cuh"

#define GRAD_MAXD 96            // parameters per theta row the gradient kernels handle

__host__ __device__ inline bool grad_shape_supported(int kind) {
    return kind == K_CONSTANT || kind == K_SE || kind == K_PERIODIC || kind == K_QP || kind == K_COSINE || kind == K_EXP ||
           kind == K_M32 || kind == K_M52;
}
__host__ __device__ inline bool grad_mean_supported(int kind) { return kind == M_CONSTANT || kind == M_LINEAR; }

// value and derivatives with respect to the shape parameters h[0..] (node.py order) of one kernel shape at lag r
__device__ __forceinline__ double shape_grad(int kind, const double* h, double r, double (&d)[3]) {
    const double ar = fabs(r);
    d[0] = d[1] = d[2] = 0.0;
    switch (kind) {
        case K_SE: {
            const double e = exp(-0.5 * r * r / (h[0] * h[0]));
            d[0] = e * r * r / (h[0] * h[0] * h[0]);
            return e;
        }
        case K_PERIODIC: {
            double s, c;
            sincos(AGPN_PI * ar / h[0], &s, &c);
            const double e = exp(-2.0 * s * s / (h[1] * h[1]));
            d[0] = e * (2.0 * AGPN_PI * ar / (h[0] * h[0] * h[1] * h[1])) * (2.0 * s * c);
            d[1] = e * 4.0 * s * s / (h[1] * h[1] * h[1]);
            return e;
        }
        case K_QP: {                        // (ell_e, P, ell_p)
            double s, c;
            sincos(AGPN_PI * ar / h[1], &s, &c);
            const double e = exp(-2.0 * s * s / (h[2] * h[2]) - r * r / (2.0 * h[0] * h[0]));
            d[0] = e * r * r / (h[0] * h[0] * h[0]);
            d[1] = e * (2.0 * AGPN_PI * ar / (h[1] * h[1] * h[2] * h[2])) * (2.0 * s * c);
            d[2] = e * 4.0 * s * s / (h[2] * h[2] * h[2]);
            return e;
        }
        case K_COSINE: {
            double s, c;
            sincos(2.0 * AGPN_PI * ar / h[0], &s, &c);
            d[0] = s * 2.0 * AGPN_PI * ar / (h[0] * h[0]);
            return c;
        }
        case K_EXP: {
            const double e = exp(-ar / h[0]);
            d[0] = e * ar / (h[0] * h[0]);
            return e;
        }
        case K_M32: {
            const double a = 1.7320508075688772 * ar / h[0];
            const double ex = exp(-a);
            d[0] = a * a * ex / h[0];
            return (1.0 + a) * ex;
        }
        case K_M52: {
            const double a = 2.23606797749979 * ar / h[0];
            const double ex = exp(-a);
            d[0] = a * a * (1.0 + a) * ex / (3.0 * h[0]);
            return (1.0 + a + a * a / 3.0) * ex;
        }
        default:
            return 0.0;
    }
}

struct GradArgs {
    const double* theta;     // [D] one row
    const double* t;         // [N]
    int N;
    int series;
    const double* Kinv_neg;  // trailing block of the augmented matrix: -K^-1, lower triangle valid; [.][ld]
    int ld;
    const double* alpha_neg; // [N] -alpha
    double tmean;
    double* partial;         // [blocks][GRAD_MAXD] per-block sums (reduced in a fixed order by grad_reduce_kernel)
};

// grid: lower-triangular 64 x 64 tiles of the N x N block; 256 threads, each a 4 x 4 patch
__global__ void __launch_bounds__(256) grad_tile_kernel(const __grid_constant__ Structure S, const GradArgs g) {
    __shared__ double red[8][GRAD_MAXD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    const int bj = t - bi * (bi + 1) / 2;
    const int q = S.q, D = S.D, i = g.series;
    double acc[GRAD_MAXD];
    for (int m = 0; m < D; ++m) acc[m] = 0.0;
    const int a0 = bi * 64 + (tid >> 4) * 4, b0 = bj * 64 + (tid & 15) * 4;
    for (int da = 0; da < 4; ++da) {
        const int a = a0 + da;
        if (a >= g.N) continue;
        const double ta = g.t[a], al_a = -g.alpha_neg[a];
        for (int db = 0; db < 4; ++db) {
            const int b = b0 + db;
            if (b > a || b >= g.N) continue;
            // G_ab, counted once for the diagonal and twice for a > b (G and dK are symmetric); overall factor 1/2
            const double G = (al_a * (-g.alpha_neg[b]) + g.Kinv_neg[(size_t)a * g.ld + b]) * ((a == b) ? 0.5 : 1.0);
            const double r = ta - g.t[b];
            for (int j = 0; j < q; ++j) {
                const int nk = S.node_kind[j], widx = j + q * i, wk = S.weight_kind[widx];
                const double* hn = g.theta + S.node_off[j];
                const double* hw = g.theta + S.weight_off[widx];
                double dn[3], dw[3];
                const double F = (nk == K_CONSTANT) ? hn[0] : shape_grad(nk, hn, r, dn);
                double Wv, amp = 1.0, sh = 1.0;
                if (wk == K_CONSTANT) {
                    Wv = hw[0] * hw[0];
                } else {
                    amp = hw[0];
                    sh = shape_grad(wk, hw + 1, r, dw);
                    Wv = amp * amp * sh;
                }
                const double GW = G * Wv, GF = G * F;
                if (nk == K_CONSTANT) {
                    acc[S.node_off[j]] += GW;
                } else {
                    const int nn = kernel_npars(0, nk);
                    for (int u = 0; u < nn; ++u) acc[S.node_off[j] + u] += GW * dn[u];
                }
                if (wk == K_CONSTANT) {
                    acc[S.weight_off[widx]] += GF * 2.0 * hw[0];
                } else {
                    acc[S.weight_off[widx]] += GF * 2.0 * amp * sh;
                    const int nw = kernel_npars(0, wk);
                    for (int u = 0; u < nw; ++u) acc[S.weight_off[widx] + 1 + u] += GF * amp * amp * dw[u];
                }
            }
            if (a == b) {
                const double jt = g.theta[S.jit_off + i];
                acc[S.jit_off + i] += G * 2.0 * jt;                // G carries the diagonal's factor 1/2: d K_aa / ds = 2 s
            }
        }
        // mean parameters: dLL / dphi = sum_a alpha_a dm(t_a) / dphi -- once per row a (first tile column, first patch column)
        if (bj == 0 && (tid & 15) == 0 && S.use_means) {
            for (int tt = S.mean_first[i]; tt < S.mean_first[i + 1]; ++tt) {
                if (S.mean_kind[tt] == M_CONSTANT) acc[S.mean_off[tt]] += al_a;
                else if (S.mean_kind[tt] == M_LINEAR) {
                    acc[S.mean_off[tt]] += al_a * (ta - g.tmean);
                    acc[S.mean_off[tt] + 1] += al_a;
                }
            }
        }
    }
    // block reduction in a fixed order: lanes by shuffles, warps through shared memory
    for (int m = 0; m < D; ++m) {
        double v = acc[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][m] = v;
    }
    __syncthreads();
    if (tid < D) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][tid];
        g.partial[(size_t)blockIdx.x * GRAD_MAXD + tid] = v;
    }
}

// grad[m] += sum over blocks (fixed order); one thread per parameter
__global__ void grad_reduce_kernel(const double* partial, int nblocks, int D, double* grad) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= D) return;
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partial[(size_t)b * GRAD_MAXD + m];
    grad[m] += v;
}

// augmented matrix around an assembled K (top-left block already written): bottom-left = I, bottom-right = 0 (lower
// block triangle only), and the right-hand side tail = 0
__global__ void grad_aug_fill_kernel(double* aug, int T, int npad) {
    const size_t total = (size_t)npad * T;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int rr = (int)(idx / T), cc = (int)(idx % T);
        aug[(size_t)(npad + rr) * T + cc] = (cc == rr) ? 1.0 : 0.0;
    }
}
