// Device-side evaluation of artgpn's node / weight kernels and mean functions.
//
// Formulas follow the reference (cited per case, file:line relative to /root/reference):
// node.py:36-327, weight.py:65-872 (value classes only), mean.py:68-202.
// A kernel is "prepared" once per (walker, kernel) into a PKern (all divisions hoisted) and then
// evaluated per matrix entry.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define AGPN_MAXQ 8          // node kernels per network
#define AGPN_MAXP 8          // series per network
#define AGPN_MAXTERMS 32     // mean terms over all series
#define AGPN_NKINDS 14
#define AGPN_NMEANKINDS 7

enum : int {
    K_CONSTANT = 0, K_WHITENOISE = 1, K_SE = 2, K_PERIODIC = 3, K_QP = 4, K_RQ = 5, K_RQP = 6,
    K_COSINE = 7, K_EXP = 8, K_M32 = 9, K_M52 = 10, K_LINEAR = 11, K_GAMMAEXP = 12, K_POLY = 13
};
enum : int { M_CONSTANT = 0, M_LINEAR = 1, M_PARABOLA = 2, M_CUBIC = 3, M_SINE = 4, M_COSINE = 5, M_KEPLERIAN = 6 };

// network structure shared by all walkers of a context (kernel classes, theta offsets)
struct Structure {
    int p, q, D;
    int use_means;
    int jit_off;
    int node_kind[AGPN_MAXQ];
    int node_off[AGPN_MAXQ];
    int weight_kind[AGPN_MAXP * AGPN_MAXQ];   // list index j + q*i (arttwo.py:274)
    int weight_off[AGPN_MAXP * AGPN_MAXQ];
    int mean_first[AGPN_MAXP + 1];            // terms of series i: [mean_first[i], mean_first[i+1])
    int mean_kind[AGPN_MAXTERMS];
    int mean_off[AGPN_MAXTERMS];
};

__host__ __device__ inline int kernel_npars(int family, int kind) {
    // node.py ctor arities; weight.py = amplitude + same, except Constant / WhiteNoise
    const int n[AGPN_NKINDS] = {1, 1, 1, 2, 3, 2, 4, 1, 1, 1, 1, 1, 2, 3};
    if (kind < 0 || kind >= AGPN_NKINDS) return -1;
    if (family == 0 || kind == K_CONSTANT || kind == K_WHITENOISE) return n[kind];
    return n[kind] + 1;
}

__host__ __device__ inline int mean_npars(int kind) {
    const int n[AGPN_NMEANKINDS] = {1, 2, 3, 4, 4, 4, 5};
    if (kind < 0 || kind >= AGPN_NMEANKINDS) return -1;
    return n[kind];
}

// prepared kernel: value = amp * shape(r, t1, t2)
struct PKern {
    int kind;
    int family;     // 0 node, 1 weight (RQP differs, node.py:186-189 vs weight.py:420-423)
    double amp;
    double a, b, c, d;
    double per;     // the period P itself for kernels with a sin(pi |r| / P) term (phase-reduced tables), else 0
};

#define AGPN_PI 3.141592653589793238462643383279502884

__host__ __device__ inline PKern prep_kernel(int family, int kind, const double* pars) {
    PKern k;
    k.kind = kind;
    k.family = family;
    k.amp = 1.0;
    k.a = k.b = k.c = k.d = 0.0;
    k.per = 0.0;
    const double* h = pars;
    if (family == 1 && kind != K_CONSTANT && kind != K_WHITENOISE) {
        k.amp = pars[0] * pars[0];          // weight**2 (weight.py:141 etc.)
        h = pars + 1;
    }
    switch (kind) {
        case K_CONSTANT:                    // node.py:50-51 (c) ; weight.py:78-79 (c**2)
            k.amp = family == 0 ? h[0] : h[0] * h[0];
            break;
        case K_WHITENOISE:                  // node.py:68-72 ; weight.py:106-110
            k.amp = h[0] * h[0];
            break;
        case K_SE:                          // exp(-0.5 r^2/ell^2)
            k.a = -0.5 / (h[0] * h[0]);
            break;
        case K_PERIODIC:                    // exp(-2 sin^2(pi|r|/P)/ell^2), pars (P, ell)
            k.b = AGPN_PI / h[0];
            k.per = h[0];
            k.a = -2.0 / (h[1] * h[1]);
            break;
        case K_QP:                          // pars (ell_e, P, ell_p)
            k.c = -1.0 / (2.0 * h[0] * h[0]);
            k.b = AGPN_PI / h[1];
            k.per = h[1];
            k.a = -2.0 / (h[2] * h[2]);
            break;
        case K_RQ:                          // 1/(1 + r^2/(2 alpha ell^2))^alpha, pars (alpha, ell)
            k.a = 1.0 / (2.0 * h[0] * h[1] * h[1]);
            k.b = h[0];
            break;
        case K_RQP:                         // pars (alpha, ell_e, P, ell_p)
            k.d = h[0];
            k.c = 1.0 / (2.0 * h[0] * h[1] * h[1]);
            k.b = AGPN_PI / h[2];
            k.per = h[2];
            k.a = -2.0 / (h[3] * h[3]);
            break;
        case K_COSINE:                      // cos(2 pi |r| / P)
            k.a = 2.0 * AGPN_PI / h[0];
            break;
        case K_EXP:                         // exp(-|r|/ell)
            k.a = -1.0 / h[0];
            break;
        case K_M32:                         // (1 + sqrt3 |r|/ell) exp(-sqrt3 |r|/ell)
            k.a = sqrt(3.0) / h[0];
            break;
        case K_M52:                         // (1 + (3 sqrt5 ell |r| + 5 r^2)/(3 ell^2)) exp(-sqrt5 |r|/ell)
            k.a = 3.0 * sqrt(5.0) * h[0];
            k.b = 1.0 / (3.0 * h[0] * h[0]);
            k.c = -sqrt(5.0) / h[0];
            break;
        case K_LINEAR:                      // (t1 - c)(t2 - c)
            k.a = h[0];
            break;
        case K_GAMMAEXP:                    // exp(-(|r|/ell)^gamma), pars (gamma, ell)
            k.b = h[0];
            k.a = 1.0 / h[1];
            break;
        case K_POLY:                        // (a t1 t2 + b)^c
            k.a = h[0];
            k.b = h[1];
            k.c = h[2];
            break;
        default:
            break;
    }
    return k;
}

// true if the kernel contains the sin(pi |r| / P) term (candidates for the angle-addition tables)
__host__ __device__ inline bool kernel_has_sin(int kind) {
    return kind == K_PERIODIC || kind == K_QP || kind == K_RQP;
}

#ifdef __CUDACC__
// sin and cos of pi dt / P for the angle-addition tables, without losing the phase to the size of the argument
// (dt / P reaches hundreds of periods: a plain sincos(b * dt) carries eps * |b dt| ~ 1e-13 of argument rounding).
// n = rint(dt / P) whole periods are removed exactly (fma: rem = dt - n P with one rounding of a small number),
// sincospi takes the remainder in units of pi, and sin(pi n + y) = (-1)^n sin y, cos likewise.
__device__ __forceinline__ void phase_sincos(double dt, double P, double* s, double* c) {
    const double n = rint(dt / P);
    const double rem = fma(-n, P, dt);
    double sv, cv;
    sincospi(rem / P, &sv, &cv);
    const bool odd = (((long long)n) & 1ll) != 0;
    *s = odd ? -sv : sv;
    *c = odd ? -cv : cv;
}

// shape value when sin(pi |r|/P) is already known (s); used by the tiled assembly fast path
__device__ __forceinline__ double kern_shape_with_sin(const PKern& k, double r, double s) {
    switch (k.kind) {
        case K_PERIODIC:
            return exp(k.a * s * s);
        case K_QP:
            return exp(fma(k.a * s, s, k.c * r * r));
        default: {                          // K_RQP
            double e = exp(k.a * s * s);
            double B = fma(k.c * r, r, 1.0);
            if (k.family == 0) {
                double sg = (B > 0.0) ? 1.0 : ((B < 0.0) ? -1.0 : 0.0);
                return e / (sg * pow(fabs(B), k.d));
            }
            return e / pow(B, k.d);
        }
    }
}

// shape value of a prepared kernel at lag r (or (t1, t2) for Linear / Polynomial);
// on_diag = "lag array is square and row index == column index" (WhiteNoise quirk)
__device__ __noinline__ double kern_shape(const PKern& k, double r, double t1, double t2, bool on_diag) {
    const double ar = fabs(r);
    switch (k.kind) {
        case K_CONSTANT:
            return 1.0;
        case K_WHITENOISE:
            return on_diag ? 1.0 : 0.0;
        case K_SE:
            return exp(k.a * r * r);
        case K_PERIODIC:
        case K_QP:
        case K_RQP:
            return kern_shape_with_sin(k, r, sin(k.b * ar));
        case K_RQ:
            return 1.0 / pow(fma(k.a * r, r, 1.0), k.b);
        case K_COSINE:
            return cos(k.a * ar);
        case K_EXP:
            return exp(k.a * ar);
        case K_M32: {
            double x = k.a * ar;
            return (1.0 + x) * exp(-x);
        }
        case K_M52:
            return (1.0 + (k.a * ar + 5.0 * r * r) * k.b) * exp(k.c * ar);
        case K_LINEAR:
            return (t1 - k.a) * (t2 - k.a);
        case K_GAMMAEXP:
            return exp(-pow(ar * k.a, k.b));
        case K_POLY:
            return pow(k.a * t1 * t2 + k.b, k.c);
        default:
            return 0.0;
    }
}

// ---- "exact" evaluation: the reference's own operation order, one IEEE rounding per numpy operation
// (explicit __d*_rn intrinsics: no hoisted reciprocals, no fma contraction).  Entries then agree with
// numpy up to the last-ulp differences between libdevice and glibc sin / cos / exp / pow.  Used when
// parity is wanted on ill-conditioned covariances, where a 1e-14 relative difference in K is amplified
// by cond(K) in the log-likelihood.  ~4x slower than the hoisted form.
struct RKern {
    int kind;
    int family;
    double w2;        // weight**2 (weight kernels), else 1
    double h[4];      // remaining .pars in constructor order
};

__host__ __device__ inline RKern prep_raw(int family, int kind, const double* pars) {
    RKern k;
    k.kind = kind;
    k.family = family;
    k.w2 = 1.0;
    const double* h = pars;
    if (family == 1 && kind != K_CONSTANT && kind != K_WHITENOISE) {
        k.w2 = pars[0] * pars[0];
        h = pars + 1;
    }
    const int n = kernel_npars(0, kind);
    for (int i = 0; i < 4; ++i) k.h[i] = i < n ? h[i] : 0.0;
    return k;
}

__device__ __noinline__ double kern_value_exact(const RKern& k, double r, double t1, double t2, bool on_diag) {
    const double ar = fabs(r);
    const double* h = k.h;
    const bool wt = k.family == 1;
    double shape;
    switch (k.kind) {
        case K_CONSTANT:                    // node: c ; weight: c**2
            return wt ? __dmul_rn(h[0], h[0]) : h[0];
        case K_WHITENOISE:
            return on_diag ? __dmul_rn(h[0], h[0]) : 0.0;
        case K_SE:                          // exp(-0.5 * r**2 / ell**2)
            shape = exp(__ddiv_rn(__dmul_rn(-0.5, __dmul_rn(r, r)), __dmul_rn(h[0], h[0])));
            break;
        case K_PERIODIC: {                  // exp(-2 * sin(pi*|r|/P)**2 / ell**2)
            const double sn = sin(__ddiv_rn(__dmul_rn(AGPN_PI, ar), h[0]));
            shape = exp(__ddiv_rn(__dmul_rn(-2.0, __dmul_rn(sn, sn)), __dmul_rn(h[1], h[1])));
            break;
        }
        case K_QP: {                        // exp(-2*sin(pi*|r|/P)**2/ell_p**2 - r**2/(2*ell_e**2))
            const double sn = sin(__ddiv_rn(__dmul_rn(AGPN_PI, ar), h[1]));
            const double a = -__ddiv_rn(__dmul_rn(2.0, __dmul_rn(sn, sn)), __dmul_rn(h[2], h[2]));
            const double b = __ddiv_rn(__dmul_rn(r, r), __dmul_rn(2.0, __dmul_rn(h[0], h[0])));
            shape = exp(__dsub_rn(a, b));
            break;
        }
        case K_RQ: {                        // 1 / (1 + r**2/(2*alpha*ell**2))**alpha
            const double base = __dadd_rn(1.0, __ddiv_rn(__dmul_rn(r, r), __dmul_rn(__dmul_rn(2.0, h[0]), __dmul_rn(h[1], h[1]))));
            const double pw = pow(base, h[0]);
            return wt ? __ddiv_rn(k.w2, pw) : __ddiv_rn(1.0, pw);
        }
        case K_RQP: {                       // pars (alpha, ell_e, P, ell_p)
            const double sn = sin(__ddiv_rn(__dmul_rn(AGPN_PI, ar), h[2]));
            const double e = exp(-__ddiv_rn(__dmul_rn(2.0, __dmul_rn(sn, sn)), __dmul_rn(h[3], h[3])));
            const double b = __dadd_rn(1.0, __ddiv_rn(__dmul_rn(r, r), __dmul_rn(__dmul_rn(2.0, h[0]), __dmul_rn(h[1], h[1]))));
            if (wt) return __ddiv_rn(__dmul_rn(k.w2, e), pow(b, h[0]));
            const double sg = (b > 0.0) ? 1.0 : ((b < 0.0) ? -1.0 : 0.0);
            return __ddiv_rn(e, __dmul_rn(sg, pow(fabs(b), h[0])));
        }
        case K_COSINE:                      // cos(2*pi*|r| / P)
            shape = cos(__ddiv_rn(__dmul_rn(2.0 * AGPN_PI, ar), h[0]));
            break;
        case K_EXP:                         // exp(-|r|/ell)
            shape = exp(__ddiv_rn(-ar, h[0]));
            break;
        case K_M32: {                       // (1 + sqrt3*|r|/ell) * exp(-sqrt3*|r|/ell)
            const double s3 = 1.7320508075688772;
            const double poly = __dadd_rn(1.0, __ddiv_rn(__dmul_rn(s3, ar), h[0]));
            const double e = exp(__ddiv_rn(__dmul_rn(-s3, ar), h[0]));
            return wt ? __dmul_rn(__dmul_rn(k.w2, poly), e) : __dmul_rn(poly, e);
        }
        case K_M52: {                       // (1 + (3*sqrt5*ell*|r| + 5*|r|**2)/(3*ell**2)) * exp(-sqrt5*|r|/ell)
            const double s5 = 2.23606797749979;
            const double num = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(3.0, s5), h[0]), ar), __dmul_rn(5.0, __dmul_rn(ar, ar)));
            const double poly = __dadd_rn(1.0, __ddiv_rn(num, __dmul_rn(3.0, __dmul_rn(h[0], h[0]))));
            const double e = exp(__ddiv_rn(__dmul_rn(-s5, ar), h[0]));
            return wt ? __dmul_rn(__dmul_rn(k.w2, poly), e) : __dmul_rn(poly, e);
        }
        case K_LINEAR: {                    // (t1 - c) * (t2 - c)
            const double a = __dsub_rn(t1, h[0]), b = __dsub_rn(t2, h[0]);
            return wt ? __dmul_rn(__dmul_rn(k.w2, a), b) : __dmul_rn(a, b);
        }
        case K_GAMMAEXP:                    // exp(-(|r|/ell)**gamma), pars (gamma, ell)
            shape = exp(-pow(__ddiv_rn(ar, h[1]), h[0]));
            break;
        case K_POLY:                        // (a*t1*t2 + b)**c
            shape = pow(__dadd_rn(__dmul_rn(__dmul_rn(h[0], t1), t2), h[1]), h[2]);
            break;
        default:
            return 0.0;
    }
    return wt ? __dmul_rn(k.w2, shape) : shape;
}

// ---- mean functions (mean.py) --------------------------------------------------------------
// Keplerian first-guess residual M1 (mean.py:175-178); E0 returned through *E0out
__device__ __forceinline__ double kepler_m1(const double* h, double t, double t0, double* E0out) {
    const double P = h[0], e = h[2], phi = h[4];
    const double T0 = t0 - (P * phi) / (2.0 * AGPN_PI);
    const double M0 = 2.0 * AGPN_PI * (t - T0) / P;
    const double E0 = M0 + e * sin(M0) + 0.5 * (e * e) * sin(2.0 * M0);
    *E0out = E0;
    return E0 - e * sin(E0) - M0;
}

// One mean term at time t.  tmean = mean of the grid (mean.py:92), t0 = first grid point
// (mean.py:175), halley = "some element of the grid had |M1| > 1e-10" (mean.py:180-198: exactly
// one third-order correction applied to every element).
__device__ __noinline__ double mean_term(int kind, const double* h, double t, double tmean, double t0, bool halley) {
    switch (kind) {
        case M_CONSTANT:
            return h[0];
        case M_LINEAR:
            return h[0] * (t - tmean) + h[1];
        case M_PARABOLA:                                   // np.polyval Horner from 0 (mean.py:107)
            return ((0.0 * t + h[0]) * t + h[1]) * t + h[2];
        case M_CUBIC:
            return (((0.0 * t + h[0]) * t + h[1]) * t + h[2]) * t + h[3];
        case M_SINE:
            return h[0] * sin((2.0 * AGPN_PI * t / h[1]) + h[2]) + h[3];
        case M_COSINE:
            return h[0] * h[0] * cos((2.0 * AGPN_PI * t / h[1]) + h[2]) + h[3];
        case M_KEPLERIAN: {
            const double K = h[1], e = h[2], w = h[3];
            double E0;
            const double M1 = kepler_m1(h, t, t0, &E0);
            if (halley) {
                const double M1p = 1.0 - e * cos(E0);
                const double M1pp = e * sin(E0);
                const double M1ppp = 1.0 - M1p;
                const double d1 = -M1 / M1p;
                const double d2 = -M1 / (M1p + d1 * M1pp / 2.0);
                const double d3 = -M1 / (M1p + d2 * M1pp / 2.0 + d2 * d2 * M1ppp / 6.0);
                E0 = E0 + d3;
            }
            const double nu = 2.0 * atan(sqrt((1.0 + e) / (1.0 - e)) * tan(E0 / 2.0));
            return K * (e * cos(w) + cos(w + nu));
        }
        default:
            return 0.0;
    }
}
#endif  // __CUDACC__
