// Covariance-block assembly and mean-residual kernels (sm_100a).
//
// assemble_kernel replaces the double loop of network.log_likelihood (arttwo.py:270-281),
// network._covariance_matrix (arttwo.py:151-192) and the K* loop of network.prediction
// (arttwo.py:373-384):   K_i = sum_j W_ij o F_j  (+ diag(sigma_i^2 + s_i^2)).
// One CTA produces one 128 x 128 tile for ALL requested series of one walker: the node kernels
// F_j do not depend on the series, so they are evaluated once per entry and reused (the
// reference recomputes them p times, arttwo.py:277).  For the likelihood path only tiles on or
// below the block diagonal are produced (the factorisation never reads the others).
//
// mean_kernel replaces network._mean / the residual y - m (arttwo.py:106-126, :262-264).
#pragma once
#include "agpn_eval.cuh"

#define ASM_TILE 128
#define ASM_NTRIG 4     // node kernels whose sin(pi r/P) term uses the angle-addition tables

struct AsmArgs {
    const double* theta;     // [W][D]
    const double* ta;        // row grid  [na]
    const double* tb;        // column grid [nb]
    int na, nb;              // valid extents
    int rows_out, cols_out;  // extents that may be written (>= na, nb when padded)
    const double* yerr;      // [p][N] raw sigma or nullptr
    double* out;             // matrices [(w * nser + s)] * mat_stride
    long long mat_stride;
    int ld;
    int series0, nser;
    int lower_only;          // tiles enumerated over the lower block triangle (square grids)
    int square;              // WhiteNoise quirk: lag array is square
    int row_off = 0;         // rows are a slice of a larger (square) grid: the quirk's diagonal is row + row_off == column
    int add_err;             // + yerr^2 on the diagonal
    int add_jit;             // + jitter^2 on the diagonal
    int pad_identity;        // entries outside [na][nb]: 1 on the diagonal, else 0
    int use_trig;            // angle-addition tables for node kernels with a sin term
    int vec2;                // 16-byte stores allowed
    int exact;               // reference operation order, no hoisting / tables / fma (kern_value_exact)
    double torigin;          // phase origin of the tables
    int* kflag;              // [W * nser] set to 1 if a produced entry is not finite (or nullptr)
    // assemble_fast_kernel only, lower_only grids: an upper bound on |entry| over each off-diagonal tile, from the decay of
    // the squared-exponential factors at the tile's smallest lag -- [W * nser][tiles] (the panel solve's bound test reads
    // this instead of the tile).  tmax_valid is set by launch_assemble when the kernel that ran has written it.
    double* tmax = nullptr;
    int tmax_valid = 0;
};

__global__ void __launch_bounds__(256) assemble_kernel(const __grid_constant__ Structure S, const AsmArgs g) {
    __shared__ PKern s_node[AGPN_MAXQ];
    __shared__ PKern s_weight[AGPN_MAXP * AGPN_MAXQ];
    __shared__ int s_slot[AGPN_MAXQ];
    __shared__ double s_ta[ASM_TILE], s_tb[ASM_TILE];
    __shared__ double s_diag[AGPN_MAXP][ASM_TILE];
    __shared__ double s_trig[ASM_NTRIG][4][ASM_TILE];   // [slot][sin row, cos row, sin col, cos col]
    // exact mode: raw parameters alias the table storage (tables are not used then)
    RKern* s_rnode = reinterpret_cast<RKern*>(&s_trig[0][0][0]);
    RKern* s_rweight = s_rnode + AGPN_MAXQ;

    const int tid = threadIdx.x;
    const int w = blockIdx.y;
    const int q = S.q;
    int bi, bj;
    if (g.lower_only) {
        const int t = blockIdx.x;
        bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        while (bi * (bi + 1) / 2 > t) --bi;
        bj = t - bi * (bi + 1) / 2;
    } else {
        const int tcols = (g.cols_out + ASM_TILE - 1) / ASM_TILE;
        bi = blockIdx.x / tcols;
        bj = blockIdx.x % tcols;
    }
    const int i0 = bi * ASM_TILE, j0 = bj * ASM_TILE;
    const double* th = g.theta + (size_t)w * S.D;

    // ---- per-CTA preparation -------------------------------------------------------------------
    if (tid < q) {
        s_node[tid] = prep_kernel(0, S.node_kind[tid], th + S.node_off[tid]);
        if (g.exact) s_rnode[tid] = prep_raw(0, S.node_kind[tid], th + S.node_off[tid]);
    } else if (tid >= 32 && tid < 32 + g.nser * q) {
        const int u = tid - 32;
        const int s = u / q, j = u % q;
        const int widx = j + q * (g.series0 + s);
        s_weight[u] = prep_kernel(1, S.weight_kind[widx], th + S.weight_off[widx]);
        if (g.exact) s_rweight[u] = prep_raw(1, S.weight_kind[widx], th + S.weight_off[widx]);
    }
    if (tid == 64) {
        int slot = 0;
        for (int j = 0; j < AGPN_MAXQ; ++j) {
            int v = -1;
            if (j < q && g.use_trig && !g.exact && kernel_has_sin(S.node_kind[j]) && slot < ASM_NTRIG) v = slot++;
            s_slot[j] = v;
        }
    }
    if (tid < ASM_TILE) {
        const int gi = i0 + tid;
        s_ta[tid] = g.ta[gi < g.na ? gi : g.na - 1];
    } else {
        const int gj = j0 + tid - ASM_TILE;
        s_tb[tid - ASM_TILE] = g.tb[gj < g.nb ? gj : g.nb - 1];
    }
    if (bi == bj && (g.add_err || g.add_jit)) {
        for (int u = tid; u < g.nser * ASM_TILE; u += 256) {
            const int s = u / ASM_TILE, r = u % ASM_TILE;
            const int gi = i0 + r;
            double v = 0.0;
            if (gi < g.na) {
                if (g.add_err) {
                    const double e = g.yerr[(size_t)(g.series0 + s) * g.na + gi];
                    v += e * e;
                }
                if (g.add_jit) {
                    const double jt = th[S.jit_off + g.series0 + s];
                    v += jt * jt;
                }
            }
            s_diag[s][r] = v;
        }
    }
    __syncthreads();
    if (g.use_trig && !g.exact) {
        for (int j = 0; j < q; ++j) {
            const int slot = s_slot[j];
            if (slot < 0) continue;
            const double per = s_node[j].per;
            double sv, cv;
            if (tid < ASM_TILE) {
                phase_sincos(s_ta[tid] - g.torigin, per, &sv, &cv);
                s_trig[slot][0][tid] = sv;
                s_trig[slot][1][tid] = cv;
            } else {
                phase_sincos(s_tb[tid - ASM_TILE] - g.torigin, per, &sv, &cv);
                s_trig[slot][2][tid - ASM_TILE] = sv;
                s_trig[slot][3][tid - ASM_TILE] = cv;
            }
        }
        __syncthreads();
    }

    // ---- entries ---------------------------------------------------------------------------------
    const int tx = tid & 63, ty = tid >> 6;
    const int c0 = 2 * tx;
    const double tb0 = s_tb[c0], tb1 = s_tb[c0 + 1];
    const int gj0 = j0 + c0;
    unsigned bad = 0;
    for (int rr = 0; rr < ASM_TILE / 4; ++rr) {
        const int row = ty + 4 * rr;
        const int gi = i0 + row;
        if (gi >= g.rows_out) break;
        const double t1 = s_ta[row];
        const double r0 = t1 - tb0, r1 = t1 - tb1;
        const bool dg0 = g.square && (gi + g.row_off == gj0), dg1 = g.square && (gi + g.row_off == gj0 + 1);
        double F0[AGPN_MAXQ], F1[AGPN_MAXQ];
#pragma unroll
        for (int j = 0; j < AGPN_MAXQ; ++j) {
            if (j < q) {
                const PKern& kn = s_node[j];
                const int slot = s_slot[j];
                if (g.exact) {
                    F0[j] = kern_value_exact(s_rnode[j], r0, t1, tb0, dg0);
                    F1[j] = kern_value_exact(s_rnode[j], r1, t1, tb1, dg1);
                } else if (slot >= 0) {
                    const double sr = s_trig[slot][0][row], cr = s_trig[slot][1][row];
                    const double sa = sr * s_trig[slot][3][c0] - cr * s_trig[slot][2][c0];
                    const double sb = sr * s_trig[slot][3][c0 + 1] - cr * s_trig[slot][2][c0 + 1];
                    F0[j] = kn.amp * kern_shape_with_sin(kn, r0, sa);
                    F1[j] = kn.amp * kern_shape_with_sin(kn, r1, sb);
                } else if (kn.kind == K_CONSTANT) {
                    F0[j] = F1[j] = kn.amp;
                } else {
                    F0[j] = kn.amp * kern_shape(kn, r0, t1, tb0, dg0);
                    F1[j] = kn.amp * kern_shape(kn, r1, t1, tb1, dg1);
                }
            } else {
                F0[j] = F1[j] = 0.0;
            }
        }
        for (int s = 0; s < g.nser; ++s) {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int j = 0; j < AGPN_MAXQ; ++j) {
                if (j < q) {
                    const PKern& kw = s_weight[s * q + j];
                    if (g.exact) {          // k_ii = k_ii + (w * f_hat), one rounding per operation (arttwo.py:279)
                        v0 = __dadd_rn(v0, __dmul_rn(kern_value_exact(s_rweight[s * q + j], r0, t1, tb0, dg0), F0[j]));
                        v1 = __dadd_rn(v1, __dmul_rn(kern_value_exact(s_rweight[s * q + j], r1, t1, tb1, dg1), F1[j]));
                    } else if (kw.kind == K_CONSTANT) {
                        v0 = fma(kw.amp, F0[j], v0);
                        v1 = fma(kw.amp, F1[j], v1);
                    } else {
                        v0 = fma(kw.amp * kern_shape(kw, r0, t1, tb0, dg0), F0[j], v0);
                        v1 = fma(kw.amp * kern_shape(kw, r1, t1, tb1, dg1), F1[j], v1);
                    }
                }
            }
            if (bi == bj && (g.add_err || g.add_jit)) {
                if (gi == gj0) v0 += s_diag[s][row];
                if (gi == gj0 + 1) v1 += s_diag[s][row];
            }
            const bool in0 = (gi < g.na) && (gj0 < g.nb), in1 = (gi < g.na) && (gj0 + 1 < g.nb);
            if (!in0) v0 = (g.pad_identity && gi == gj0) ? 1.0 : 0.0;
            if (!in1) v1 = (g.pad_identity && gi == gj0 + 1) ? 1.0 : 0.0;
            if (!isfinite(v0) || !isfinite(v1)) bad |= 1u << s;
            double* dst = g.out + (size_t)((size_t)w * g.nser + s) * g.mat_stride + (size_t)gi * g.ld + gj0;
            if (g.vec2 && gj0 + 1 < g.cols_out) {
                *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            } else {
                if (gj0 < g.cols_out) dst[0] = v0;
                if (gj0 + 1 < g.cols_out) dst[1] = v1;
            }
        }
    }
    if (bad && g.kflag != nullptr) {
        for (int s = 0; s < g.nser; ++s)
            if (bad & (1u << s)) g.kflag[(size_t)w * g.nser + s] = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path of the likelihood assembly: every node kernel is SquaredExponential / Periodic /
// QuasiPeriodic (value exp(a s^2 + c r^2), s = sin(pi r / P)) and every weight is Constant or
// SquaredExponential (value w^2 exp(cw r^2)), so  W_ij F_j = w_ij^2 exp(a_j s_j^2 + (c_j + cw_ij) r^2):
// one exponential per node (Constant weights: shared by all series) or per (series, node).
// The sine comes from per-tile angle-addition tables, the exponential from a branch-free
// polynomial (arguments are <= 0 here).  Same output layout as assemble_kernel (lower block
// triangle, padding = identity); P = number of series (compile time), WSE = some weight is SE.
// ---------------------------------------------------------------------------------------------
#define ASMF_MAXQ 4

// e^x for x <= 0 (NaN propagates, x < -708 flushes to 0): x = (16 k + j) ln2 / 16 + r, |r| <= ln2 / 32,
// e^x = 2^k * 2^(j/16) * P6(r).  The sixteen 2^(j/16) sit in shared memory (`tab`: 128 bytes = each entry on its own pair
// of banks, equal indices broadcast -> conflict free); P6 is the degree-6 Chebyshev fit of e^r on the interval (exact
// rational arithmetic, max error 7e-18) -- 11 FP64 operations instead of the 17 of a table-free degree-11 polynomial.
__constant__ double c_exp2_16[16] = {
    0x1.0000000000000p+0, 0x1.0b5586cf9890fp+0, 0x1.172b83c7d517bp+0, 0x1.2387a6e756238p+0,
    0x1.306fe0a31b715p+0, 0x1.3dea64c123422p+0, 0x1.4bfdad5362a27p+0, 0x1.5ab07dd485429p+0,
    0x1.6a09e667f3bcdp+0, 0x1.7a11473eb0187p+0, 0x1.8ace5422aa0dbp+0, 0x1.9c49182a3f090p+0,
    0x1.ae89f995ad3adp+0, 0x1.c199bdd85529cp+0, 0x1.d5818dcfba487p+0, 0x1.ea4afa2a490dap+0};

__device__ __forceinline__ double exp_nonpos(double x, const double* __restrict__ tab) {
    const double t = fma(x, 23.083120654223414, 6755399441055744.0);          // 16 / ln2; low word of t = n = 16 k + j
    const int n = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -0.04332169877307024, x);                               // ln2 / 16, high 32 bits (kd * hi is exact)
    r = fma(kd, -1.1926343307941173e-11, r);
    double p = 0x1.6c181f479b148p-10;
    p = fma(p, r, 0x1.11126eecbabeap-7);
    p = fma(p, r, 0x1.55555554ad3e3p-5);
    p = fma(p, r, 0x1.555555540526ep-3);
    p = fma(p, r, 0x1.0000000000003p-1);
    p = fma(p, r, 0x1.000000000000ap+0);
    p = fma(p, r, 1.0);
    const double tv = tab[n & 15];
    // tv * 2^k by adding k to the exponent field: x >= -708 keeps the result normal
    const double scale = __hiloint2double(__double2hiint(tv) + ((n >> 4) << 20), __double2loint(tv));
    return (x < -708.0) ? 0.0 : p * scale;
}

template <int P, bool WSE, bool EDGE>
__device__ __forceinline__ void asm_fast_rows(const AsmArgs& g, int q, int w, int bi, int bj, int i0, int j0,
                                              const double (*s_a)[2], const double (*s_amp)[ASMF_MAXQ],
                                              const double (*s_cs)[ASMF_MAXQ], const double (*s_rowt)[2][ASM_TILE],
                                              const double (*s_colt)[2][ASM_TILE], const double* s_ta, const double* s_tb,
                                              const double (*s_diag)[ASM_TILE], const double* s_etab, double (&chk)[P]) {
    const int tid = threadIdx.x;
    const int tx = tid & 63, ty = tid >> 6;
    const int c0 = 2 * tx;
    const double tb0 = s_tb[c0], tb1 = s_tb[c0 + 1];
    const int gj0 = j0 + c0;
    for (int rr = 0; rr < ASM_TILE / 4; ++rr) {
        const int row = ty + 4 * rr;
        const int gi = i0 + row;
        const double t1 = s_ta[row];
        const double r0 = t1 - tb0, r1 = t1 - tb1;
        const double rr0 = r0 * r0, rr1 = r1 * r1;
        double v0[P], v1[P];
#pragma unroll
        for (int s = 0; s < P; ++s) v0[s] = v1[s] = 0.0;
        for (int j = 0; j < q; ++j) {
            const double sr = s_rowt[j][0][row], cr = s_rowt[j][1][row];
            const double2 sc = *reinterpret_cast<const double2*>(&s_colt[j][0][c0]);
            const double2 cc = *reinterpret_cast<const double2*>(&s_colt[j][1][c0]);
            const double s0 = fma(sr, cc.x, -cr * sc.x), s1 = fma(sr, cc.y, -cr * sc.y);
            const double a = s_a[j][0];
            const double q0 = a * s0 * s0, q1 = a * s1 * s1;
            if (!WSE) {
                const double c = s_a[j][1];
                const double e0 = exp_nonpos(fma(c, rr0, q0), s_etab), e1 = exp_nonpos(fma(c, rr1, q1), s_etab);
#pragma unroll
                for (int s = 0; s < P; ++s) {
                    v0[s] = fma(s_amp[s][j], e0, v0[s]);
                    v1[s] = fma(s_amp[s][j], e1, v1[s]);
                }
            } else {
#pragma unroll
                for (int s = 0; s < P; ++s) {
                    const double c = s_cs[s][j];
                    v0[s] = fma(s_amp[s][j], exp_nonpos(fma(c, rr0, q0), s_etab), v0[s]);
                    v1[s] = fma(s_amp[s][j], exp_nonpos(fma(c, rr1, q1), s_etab), v1[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < P; ++s) {
            double a0 = v0[s], a1 = v1[s];
            if (EDGE) {
                if (bi == bj) {
                    if (gi == gj0) a0 += s_diag[s][row];
                    if (gi == gj0 + 1) a1 += s_diag[s][row];
                }
                const bool in0 = (gi < g.na) && (gj0 < g.nb), in1 = (gi < g.na) && (gj0 + 1 < g.nb);
                if (!in0) a0 = (gi == gj0) ? 1.0 : 0.0;
                if (!in1) a1 = (gi == gj0 + 1) ? 1.0 : 0.0;
            }
            chk[s] += a0 + a1;
            double* dst = g.out + (size_t)((size_t)w * P + s) * g.mat_stride + (size_t)gi * g.ld + gj0;
            *reinterpret_cast<double2*>(dst) = make_double2(a0, a1);
        }
    }
}

template <int P, bool WSE>
__global__ void __launch_bounds__(256, 2) assemble_fast_kernel(const __grid_constant__ Structure S, const AsmArgs g) {
    __shared__ double s_a[ASMF_MAXQ][2];                       // a_j, c_j
    __shared__ double s_b[ASMF_MAXQ];                          // P_j (0: no sine term)
    __shared__ double s_amp[AGPN_MAXP][ASMF_MAXQ];             // w_ij^2
    __shared__ double s_cs[AGPN_MAXP][ASMF_MAXQ];              // c_j + cw_ij
    __shared__ double s_ta[ASM_TILE], s_tb[ASM_TILE];
    __shared__ double s_diag[AGPN_MAXP][ASM_TILE];
    __shared__ __align__(16) double s_rowt[ASMF_MAXQ][2][ASM_TILE];   // sin, cos of b (t_row - t0)
    __shared__ __align__(16) double s_colt[ASMF_MAXQ][2][ASM_TILE];   // sin, cos of b (t_col - t0)
    __shared__ __align__(128) double s_etab[16];                       // 2^(j/16) for exp_nonpos

    const int tid = threadIdx.x;
    const int w = blockIdx.y;
    const int q = S.q;
    const int t = blockIdx.x;
    if (tid >= 240) s_etab[tid - 240] = c_exp2_16[tid - 240];
    int bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    const int bj = t - bi * (bi + 1) / 2;
    const int i0 = bi * ASM_TILE, j0 = bj * ASM_TILE;
    const double* th = g.theta + (size_t)w * S.D;

    if (tid < q) {
        const PKern k = prep_kernel(0, S.node_kind[tid], th + S.node_off[tid]);
        s_a[tid][0] = (k.kind == K_SE) ? 0.0 : k.a;
        s_a[tid][1] = (k.kind == K_SE) ? k.a : k.c;            // SE keeps its -1/(2 ell^2) in .a
        s_b[tid] = (k.kind == K_SE) ? 0.0 : k.per;
    }
    __syncthreads();
    if (tid >= 32 && tid < 32 + P * q) {
        const int u = tid - 32;
        const int s = u / q, j = u % q;
        const int widx = j + q * (g.series0 + s);
        const PKern k = prep_kernel(1, S.weight_kind[widx], th + S.weight_off[widx]);
        s_amp[s][j] = k.amp;
        s_cs[s][j] = s_a[j][1] + ((k.kind == K_SE) ? k.a : 0.0);
    }
    if (tid < ASM_TILE) {
        const int gi = i0 + tid;
        s_ta[tid] = g.ta[gi < g.na ? gi : g.na - 1];
    } else {
        const int gj = j0 + tid - ASM_TILE;
        s_tb[tid - ASM_TILE] = g.tb[gj < g.nb ? gj : g.nb - 1];
    }
    if (bi == bj) {
        for (int u = tid; u < P * ASM_TILE; u += 256) {
            const int s = u / ASM_TILE, r = u % ASM_TILE;
            const int gi = i0 + r;
            double v = 0.0;
            if (gi < g.na) {
                const double e = g.yerr[(size_t)(g.series0 + s) * g.na + gi];
                const double jt = th[S.jit_off + g.series0 + s];
                v = e * e + jt * jt;
            }
            s_diag[s][r] = v;
        }
    }
    __syncthreads();
    for (int j = 0; j < q; ++j) {
        const double per = s_b[j];
        const bool nosin = S.node_kind[j] == K_SE;
        double sv = 0.0, cv = 1.0;
        if (tid < ASM_TILE) {
            if (!nosin) phase_sincos(s_ta[tid] - g.torigin, per, &sv, &cv);
            s_rowt[j][0][tid] = sv;
            s_rowt[j][1][tid] = cv;
        } else {
            if (!nosin) phase_sincos(s_tb[tid - ASM_TILE] - g.torigin, per, &sv, &cv);
            s_colt[j][0][tid - ASM_TILE] = sv;
            s_colt[j][1][tid - ASM_TILE] = cv;
        }
    }
    __syncthreads();

    if (g.tmax != nullptr && bi != bj) {                       // CTA-uniform
        // every entry is sum_j amp_sj exp(c_sj r^2 + a_j sin^2(..)) with a_j, c_sj <= 0:  <= sum_j amp_sj exp(c_sj r_min^2)
        __shared__ double s_mm[8][2];
        const double v = (tid < ASM_TILE) ? s_ta[tid] : s_tb[tid - ASM_TILE];
        double lo = v, hi = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((tid & 31) == 0) { s_mm[tid >> 5][0] = lo; s_mm[tid >> 5][1] = hi; }
        __syncthreads();
        if (tid < P) {
            const double lo_a = fmin(fmin(s_mm[0][0], s_mm[1][0]), fmin(s_mm[2][0], s_mm[3][0]));
            const double hi_a = fmax(fmax(s_mm[0][1], s_mm[1][1]), fmax(s_mm[2][1], s_mm[3][1]));
            const double lo_b = fmin(fmin(s_mm[4][0], s_mm[5][0]), fmin(s_mm[6][0], s_mm[7][0]));
            const double hi_b = fmax(fmax(s_mm[4][1], s_mm[5][1]), fmax(s_mm[6][1], s_mm[7][1]));
            const double rmin = fmax(0.0, fmax(lo_a - hi_b, lo_b - hi_a));
            double bound = 0.0;
            for (int j = 0; j < q; ++j) bound += fabs(s_amp[tid][j]) * exp(s_cs[tid][j] * rmin * rmin);
            g.tmax[((size_t)w * P + tid) * gridDim.x + t] = bound * 1.000001;       // NaN parameters give NaN: "unknown"
        }
    }
    double chk[P];
#pragma unroll
    for (int s = 0; s < P; ++s) chk[s] = 0.0;
    const bool edge = (bi == bj) || (i0 + ASM_TILE > g.na) || (j0 + ASM_TILE > g.nb);
    if (edge)
        asm_fast_rows<P, WSE, true>(g, q, w, bi, bj, i0, j0, s_a, s_amp, s_cs, s_rowt, s_colt, s_ta, s_tb, s_diag, s_etab, chk);
    else
        asm_fast_rows<P, WSE, false>(g, q, w, bi, bj, i0, j0, s_a, s_amp, s_cs, s_rowt, s_colt, s_ta, s_tb, s_diag, s_etab, chk);
    if (g.kflag != nullptr) {
#pragma unroll
        for (int s = 0; s < P; ++s)
            if (!isfinite(chk[s])) g.kflag[(size_t)w * P + s] = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// Generic assembly, blocked by kernel class (round 2).  assemble_kernel above evaluates every entry through one out-of-line
// `switch` per (entry, node) and per (entry, series, node) with library pow / exp -- 12.7 % of the HBM roof on a
// Matern52 + RationalQuadratic network.  Here the dispatch on the kernel class is hoisted out of the entry loop: a thread owns
// 2 adjacent columns x 4 rows (8 entries) at a time, and for each node / weight kernel ONE switch selects a loop over those
// 8 entries that is specialised at compile time for the class (shape_block<KIND>); exponentials of non-positive arguments
// use the table-based exp_nonpos, powers exp(alpha log B).  Same output layout, flags and quirks as assemble_kernel
// (which stays for the exact mode and for more than four series per launch).
// ---------------------------------------------------------------------------------------------
#define ASMB_R 4        // rows per thread per block of entries

__device__ __forceinline__ double exp_any(double x, const double* __restrict__ tab) { return x <= 0.0 ? exp_nonpos(x, tab) : exp(x); }

template <int KIND>
__device__ __forceinline__ void shape_block(const PKern& k, const double (&r)[ASMB_R][2], const double (&t1)[ASMB_R], const double (&t2)[2],
                                            const bool (&dg)[ASMB_R][2], const double (*sinv)[2], const double* __restrict__ etab,
                                            double (&out)[ASMB_R][2]) {
#pragma unroll
    for (int u = 0; u < ASMB_R; ++u)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double rv = r[u][e], ar = fabs(rv);
            double v;
            if (KIND == K_CONSTANT) v = 1.0;
            else if (KIND == K_WHITENOISE) v = dg[u][e] ? 1.0 : 0.0;
            else if (KIND == K_SE) v = exp_any(k.a * rv * rv, etab);
            else if (KIND == K_PERIODIC) {
                const double sn = sinv ? sinv[u][e] : sin(k.b * ar);
                v = exp_any(k.a * sn * sn, etab);
            } else if (KIND == K_QP) {
                const double sn = sinv ? sinv[u][e] : sin(k.b * ar);
                v = exp_any(fma(k.a * sn, sn, k.c * rv * rv), etab);
            } else if (KIND == K_RQ) {
                const double B = fma(k.a * rv, rv, 1.0);
                v = (B > 0.0) ? exp_any(-k.b * log(B), etab) : 1.0 / pow(B, k.b);
            } else if (KIND == K_RQP) {
                const double sn = sinv ? sinv[u][e] : sin(k.b * ar);
                v = kern_shape_with_sin(k, rv, sn);
            } else if (KIND == K_COSINE) v = cos(k.a * ar);
            else if (KIND == K_EXP) v = exp_any(k.a * ar, etab);
            else if (KIND == K_M32) {
                const double x = k.a * ar;
                v = (1.0 + x) * exp_any(-x, etab);
            } else if (KIND == K_M52) v = (1.0 + (k.a * ar + 5.0 * rv * rv) * k.b) * exp_any(k.c * ar, etab);
            else if (KIND == K_LINEAR) v = (t1[u] - k.a) * (t2[e] - k.a);
            else if (KIND == K_GAMMAEXP) v = exp(-pow(ar * k.a, k.b));
            else v = pow(k.a * t1[u] * t2[e] + k.b, k.c);          // K_POLY
            out[u][e] = v;
        }
}

__device__ __forceinline__ void shape_block_dispatch(const PKern& k, const double (&r)[ASMB_R][2], const double (&t1)[ASMB_R],
                                                     const double (&t2)[2], const bool (&dg)[ASMB_R][2], const double (*sinv)[2],
                                                     const double* __restrict__ etab, double (&out)[ASMB_R][2]) {
    switch (k.kind) {
        case K_CONSTANT: shape_block<K_CONSTANT>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_WHITENOISE: shape_block<K_WHITENOISE>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_SE: shape_block<K_SE>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_PERIODIC: shape_block<K_PERIODIC>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_QP: shape_block<K_QP>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_RQ: shape_block<K_RQ>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_RQP: shape_block<K_RQP>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_COSINE: shape_block<K_COSINE>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_EXP: shape_block<K_EXP>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_M32: shape_block<K_M32>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_M52: shape_block<K_M52>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_LINEAR: shape_block<K_LINEAR>(k, r, t1, t2, dg, sinv, etab, out); break;
        case K_GAMMAEXP: shape_block<K_GAMMAEXP>(k, r, t1, t2, dg, sinv, etab, out); break;
        default: shape_block<K_POLY>(k, r, t1, t2, dg, sinv, etab, out); break;
    }
}

#ifndef ASMB_MINB
#define ASMB_MINB 2             // two CTAs per SM (128 registers, some spills): 1.20 vs 1.36 ms on the Matern52 + RationalQuadratic probe
#endif
template <int NSER>
__global__ void __launch_bounds__(256, ASMB_MINB) assemble_block_kernel(const __grid_constant__ Structure S, const AsmArgs g) {
    __shared__ PKern s_node[AGPN_MAXQ];
    __shared__ PKern s_weight[NSER * AGPN_MAXQ];
    __shared__ int s_slot[AGPN_MAXQ];
    __shared__ double s_ta[ASM_TILE], s_tb[ASM_TILE];
    __shared__ double s_diag[NSER][ASM_TILE];
    __shared__ double s_trig[ASM_NTRIG][4][ASM_TILE];   // [slot][sin row, cos row, sin col, cos col]
    __shared__ __align__(128) double s_etab[16];

    const int tid = threadIdx.x;
    const int w = blockIdx.y;
    const int q = S.q;
    if (tid >= 240) s_etab[tid - 240] = c_exp2_16[tid - 240];
    int bi, bj;
    if (g.lower_only) {
        const int t = blockIdx.x;
        bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        while (bi * (bi + 1) / 2 > t) --bi;
        bj = t - bi * (bi + 1) / 2;
    } else {
        const int tcols = (g.cols_out + ASM_TILE - 1) / ASM_TILE;
        bi = blockIdx.x / tcols;
        bj = blockIdx.x % tcols;
    }
    const int i0 = bi * ASM_TILE, j0 = bj * ASM_TILE;
    const double* th = g.theta + (size_t)w * S.D;

    if (tid < q) {
        s_node[tid] = prep_kernel(0, S.node_kind[tid], th + S.node_off[tid]);
    } else if (tid >= 32 && tid < 32 + NSER * q) {
        const int u = tid - 32;
        const int sidx = u / q, j = u % q;
        const int widx = j + q * (g.series0 + sidx);
        s_weight[u] = prep_kernel(1, S.weight_kind[widx], th + S.weight_off[widx]);
    }
    if (tid == 64) {
        int slot = 0;
        for (int j = 0; j < AGPN_MAXQ; ++j) {
            int v = -1;
            if (j < q && g.use_trig && kernel_has_sin(S.node_kind[j]) && slot < ASM_NTRIG) v = slot++;
            s_slot[j] = v;
        }
    }
    if (tid < ASM_TILE) {
        const int gi = i0 + tid;
        s_ta[tid] = g.ta[gi < g.na ? gi : g.na - 1];
    } else {
        const int gj = j0 + tid - ASM_TILE;
        s_tb[tid - ASM_TILE] = g.tb[gj < g.nb ? gj : g.nb - 1];
    }
    if (bi == bj && (g.add_err || g.add_jit)) {
        for (int u = tid; u < NSER * ASM_TILE; u += 256) {
            const int sidx = u / ASM_TILE, r = u % ASM_TILE;
            const int gi = i0 + r;
            double v = 0.0;
            if (gi < g.na) {
                if (g.add_err) {
                    const double e = g.yerr[(size_t)(g.series0 + sidx) * g.na + gi];
                    v += e * e;
                }
                if (g.add_jit) {
                    const double jt = th[S.jit_off + g.series0 + sidx];
                    v += jt * jt;
                }
            }
            s_diag[sidx][r] = v;
        }
    }
    __syncthreads();
    if (g.use_trig) {
        for (int j = 0; j < q; ++j) {
            const int slot = s_slot[j];
            if (slot < 0) continue;
            const double per = s_node[j].per;
            double sv, cv;
            if (tid < ASM_TILE) {
                phase_sincos(s_ta[tid] - g.torigin, per, &sv, &cv);
                s_trig[slot][0][tid] = sv;
                s_trig[slot][1][tid] = cv;
            } else {
                phase_sincos(s_tb[tid - ASM_TILE] - g.torigin, per, &sv, &cv);
                s_trig[slot][2][tid - ASM_TILE] = sv;
                s_trig[slot][3][tid - ASM_TILE] = cv;
            }
        }
        __syncthreads();
    }

    const int tx = tid & 63, ty = tid >> 6;
    const int c0 = 2 * tx;
    const double t2[2] = {s_tb[c0], s_tb[c0 + 1]};
    const int gj0 = j0 + c0;
    unsigned bad = 0;
#pragma unroll 1
    for (int rb = 0; rb < ASM_TILE / 4; rb += ASMB_R) {
        // rows ty + 4 (rb + u), u < ASMB_R
        double r[ASMB_R][2], t1[ASMB_R], v[NSER][ASMB_R][2];
        bool dg[ASMB_R][2];
#pragma unroll
        for (int u = 0; u < ASMB_R; ++u) {
            const int row = ty + 4 * (rb + u);
            const int gi = i0 + row;
            t1[u] = s_ta[row];
            r[u][0] = t1[u] - t2[0];
            r[u][1] = t1[u] - t2[1];
            dg[u][0] = g.square && (gi + g.row_off == gj0);
            dg[u][1] = g.square && (gi + g.row_off == gj0 + 1);
#pragma unroll
            for (int sidx = 0; sidx < NSER; ++sidx) v[sidx][u][0] = v[sidx][u][1] = 0.0;
        }
        for (int j = 0; j < q; ++j) {
            const PKern& kn = s_node[j];
            const int slot = s_slot[j];
            double sinv[ASMB_R][2], F[ASMB_R][2];
            if (slot >= 0) {
#pragma unroll
                for (int u = 0; u < ASMB_R; ++u) {
                    const int row = ty + 4 * (rb + u);
                    const double sr = s_trig[slot][0][row], cr = s_trig[slot][1][row];
                    sinv[u][0] = sr * s_trig[slot][3][c0] - cr * s_trig[slot][2][c0];
                    sinv[u][1] = sr * s_trig[slot][3][c0 + 1] - cr * s_trig[slot][2][c0 + 1];
                }
            }
            shape_block_dispatch(kn, r, t1, t2, dg, slot >= 0 ? sinv : nullptr, s_etab, F);
#pragma unroll
            for (int sidx = 0; sidx < NSER; ++sidx) {
                const PKern& kw = s_weight[sidx * q + j];
                const double amp = kw.amp * kn.amp;
                if (kw.kind == K_CONSTANT) {
#pragma unroll
                    for (int u = 0; u < ASMB_R; ++u) {
                        v[sidx][u][0] = fma(amp, F[u][0], v[sidx][u][0]);
                        v[sidx][u][1] = fma(amp, F[u][1], v[sidx][u][1]);
                    }
                } else {
                    double Wt[ASMB_R][2];
                    shape_block_dispatch(kw, r, t1, t2, dg, nullptr, s_etab, Wt);
#pragma unroll
                    for (int u = 0; u < ASMB_R; ++u) {
                        v[sidx][u][0] = fma(amp * Wt[u][0], F[u][0], v[sidx][u][0]);
                        v[sidx][u][1] = fma(amp * Wt[u][1], F[u][1], v[sidx][u][1]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < ASMB_R; ++u) {
            const int row = ty + 4 * (rb + u);
            const int gi = i0 + row;
            if (gi >= g.rows_out) continue;
#pragma unroll
            for (int sidx = 0; sidx < NSER; ++sidx) {
                double v0 = v[sidx][u][0], v1 = v[sidx][u][1];
                if (bi == bj && (g.add_err || g.add_jit)) {
                    if (gi == gj0) v0 += s_diag[sidx][row];
                    if (gi == gj0 + 1) v1 += s_diag[sidx][row];
                }
                const bool in0 = (gi < g.na) && (gj0 < g.nb), in1 = (gi < g.na) && (gj0 + 1 < g.nb);
                if (!in0) v0 = (g.pad_identity && gi == gj0) ? 1.0 : 0.0;
                if (!in1) v1 = (g.pad_identity && gi == gj0 + 1) ? 1.0 : 0.0;
                if (!isfinite(v0) || !isfinite(v1)) bad |= 1u << sidx;
                double* dst = g.out + (size_t)((size_t)w * NSER + sidx) * g.mat_stride + (size_t)gi * g.ld + gj0;
                if (g.vec2 && gj0 + 1 < g.cols_out) {
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    if (gj0 < g.cols_out) dst[0] = v0;
                    if (gj0 + 1 < g.cols_out) dst[1] = v1;
                }
            }
        }
    }
    if (bad && g.kflag != nullptr) {
#pragma unroll
        for (int sidx = 0; sidx < NSER; ++sidx)
            if (bad & (1u << sidx)) g.kflag[(size_t)w * NSER + sidx] = 1;
    }
}

// ---------------------------------------------------------------------------------------------
struct MeanArgs {
    const double* theta;   // [W][D]
    const double* t;       // [n]
    int n, n_pad;
    double tmean, t0;
    const double* y;       // [p][n] (subtract mode) or nullptr
    double* out;           // [(w * nser + s)][n_pad]
    int series0, nser;
    int subtract;          // 1: out = y - m ; 0: out = m
    int* rflag;            // [W * nser] non-finite residual flag (or nullptr)
};

// grid (nser, W), 256 threads
__global__ void __launch_bounds__(256) mean_kernel(const __grid_constant__ Structure S, const MeanArgs g) {
    __shared__ int s_halley[AGPN_MAXTERMS];
    const int s = blockIdx.x, w = blockIdx.y;
    const int series = g.series0 + s;
    const double* th = g.theta + (size_t)w * S.D;
    const int first = S.use_means ? S.mean_first[series] : 0;
    const int last = S.use_means ? S.mean_first[series + 1] : 0;
    const int tid = threadIdx.x;
    for (int tt = first; tt < last; ++tt) {
        int any = 0;
        if (S.mean_kind[tt] == M_KEPLERIAN) {
            int local = 0;
            for (int n = tid; n < g.n; n += 256) {
                double E0;
                const double M1 = kepler_m1(th + S.mean_off[tt], g.t[n], g.t0, &E0);
                local |= (fabs(M1) > 1e-10) ? 1 : 0;
            }
            any = __syncthreads_or(local);
        }
        if (tid == 0) s_halley[tt - first] = any;
    }
    __syncthreads();
    double* out = g.out + ((size_t)w * g.nser + s) * g.n_pad;
    int bad = 0;
    for (int n = tid; n < g.n_pad; n += 256) {
        double v = 0.0;
        if (n < g.n) {
            const double t = g.t[n];
            double m = 0.0;
            for (int tt = first; tt < last; ++tt)
                m += mean_term(S.mean_kind[tt], th + S.mean_off[tt], t, g.tmean, g.t0, s_halley[tt - first] != 0);
            v = g.subtract ? (g.y[(size_t)series * g.n + n] - m) : m;
            if (!isfinite(v)) bad = 1;
        }
        out[n] = v;
    }
    if (bad && g.rflag != nullptr) g.rflag[(size_t)w * g.nser + s] = 1;
}

// element-wise value of one kernel object (node.py / weight.py __call__)
__global__ void kernel_eval_kernel(PKern k, const double* r, const double* t1, const double* t2,
                                   int rows, int cols, int square, double* out) {
    const size_t n = (size_t)rows * cols;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / cols), j = (int)(idx % cols);
        const double rv = r ? r[idx] : 0.0;
        const double a = t1 ? t1[i] : 0.0, b = t2 ? t2[j] : 0.0;
        out[idx] = k.amp * kern_shape(k, rv, a, b, square && i == j);
    }
}

// per-walker reduction of the per-(walker, series) pieces -> log-likelihood (arttwo.py:285-290)
__global__ void finalize_kernel(int W, int p, double half_n_log2pi, const double* quad, const double* logdet,
                                const int* info, const int* kflag, const int* rflag, double* out, int* status) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double ll = 0.0;
    int st = 0;
    for (int s = 0; s < p; ++s) {
        const int m = w * p + s;
        if (kflag[m]) { st = 2; break; }      // cho_factor check_finite -> ValueError
        if (info[m]) { st = 1; break; }       // LinAlgError -> -inf (arttwo.py:288-289)
        if (rflag[m]) { st = 2; break; }      // cho_solve check_finite on the residual
        ll += -0.5 * quad[m] - logdet[m] - half_n_log2pi;
    }
    out[w] = st == 0 ? ll : (st == 1 ? -INFINITY : NAN);
    status[w] = st;
}
