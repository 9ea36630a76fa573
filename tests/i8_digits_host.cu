// Host-side check of the digit algebra of the INT8-sliced trailing update (no GPU needed; built by
// tests/test_i8_digits_host.py with nvcc): the eight digits i8_digits2 produces must satisfy
//     rint(x) = sum_s a_s 128^(7-s),  a_1 .. a_7 in [-64, 63],  |a_0| <= 127
// and the epilogue's i8_fold4 must be exact; a small end-to-end case checks that the 36 slice pairs (s + t <= 7)
// reproduce a rank-K update of doubles to 2^-53 of the row scales.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../artgpn_b200/csrc/agpn_i8.cuh"

static int s8(unsigned b) { return (int)(signed char)(b & 255u); }

int main() {
    uint64_t st = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    long long bad = 0, checked = 0;
    const double qmax = 127.0 * 562949953421312.0;
    for (int it = 0; it < 400000; ++it) {
        double x0, x1;
        if (it < 12) {
            const double e[12] = {0.0, 1.0, -1.0, 63.0, 64.0, -64.0, -65.0, 0.5, 1.5, -0.5, qmax, -qmax};
            x0 = e[it]; x1 = e[11 - it];
        } else {
            const int sh0 = (int)(rnd() % 57), sh1 = (int)(rnd() % 57);
            x0 = ((double)(long long)(rnd() >> 8) / 72057594037927936.0 * 2.0 - 1.0) * std::ldexp(127.0, sh0 - 7);
            x1 = ((double)(long long)(rnd() >> 8) / 72057594037927936.0 * 2.0 - 1.0) * std::ldexp(127.0, sh1 - 7);
        }
        unsigned piece[I8_NSL];
        i8_digits2(x0, x1, piece);
        for (int v = 0; v < 2; ++v) {
            const double x = v ? x1 : x0;
            long long q = 0;
            bool ok = true;
            for (int s = 0; s < I8_NSL; ++s) {
                const int a = s8(piece[s] >> (8 * v));
                if (s > 0 && (a < -64 || a > 63)) ok = false;
                q = q * 128 + a;
            }
            if (q != llrint(x)) ok = false;
            if ((piece[0] >> 16) || (piece[7] >> 16)) ok = false;      // pieces are 16 bit
            ++checked;
            if (!ok && bad++ < 5) printf("digits wrong for x = %.17g: q = %lld, rint = %lld\n", x, q, llrint(x));
        }
    }
    // i8_fold4: exact for every int32 diagonal a depth of 32768 columns can produce
    for (int it = 0; it < 200000; ++it) {
        int s[4];
        for (int k = 0; k < 4; ++k) s[k] = (int)((long long)(rnd() % 1073741824ull) - 536870912ll);
        const long double ref = (((long double)s[0] * 128 + s[1]) * 128 + s[2]) * 128 + s[3];
        ++checked;
        if ((long double)i8_fold4(s[0], s[1], s[2], s[3]) != ref && bad++ < 5) printf("fold4 wrong\n");
    }
    // end to end: C = sum_k L_ik L_jk from digits, diagonals d <= 7 only, vs long double
    {
        const int K = 1024;
        double worst = 0.0;
        for (int trial = 0; trial < 8; ++trial) {
            static double a[1024], b[1024];
            static int da[1024][8], db[1024][8];
            const double ra = 3.0 + trial, rb = 0.01 * (1 + trial);       // row scales
            long double exact = 0.0L;
            for (int k = 0; k < K; ++k) {
                a[k] = ((double)(rnd() >> 11) / 9007199254740992.0 * 2.0 - 1.0) * ra * std::ldexp(1.0, -(int)(rnd() % 30));
                b[k] = ((double)(rnd() >> 11) / 9007199254740992.0 * 2.0 - 1.0) * rb * std::ldexp(1.0, -(int)(rnd() % 30));
                exact += (long double)a[k] * (long double)b[k];
                unsigned pa[8], pb[8];
                i8_digits2(a[k] * (I8_Q / ra), b[k] * (I8_Q / rb), pa);
                for (int s = 0; s < 8; ++s) { da[k][s] = s8(pa[s]); db[k][s] = s8(pa[s] >> 8); }
                (void)pb;
            }
            int S[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < K; ++k)
                for (int s = 0; s < 8; ++s)
                    for (int t = 0; s + t < 8; ++t) S[s + t] += da[k][s] * db[k][t];
            const double H = i8_fold4(S[0], S[1], S[2], S[3]), Lo = i8_fold4(S[4], S[5], S[6], S[7]);
            const double val = std::fma(H, 268435456.0, Lo);
            const double ei = ra * (1.0 / (126.0 * 16777216.0)), ej = 0.5 * rb * (1.0 / (126.0 * 16777216.0));
            const double got = (val * ei) * ej;
            const double err = (double)fabsl((long double)got - exact) / (ra * rb);
            if (err > worst) worst = err;
        }
        ++checked;
        printf("end-to-end worst error / (r_i r_j) = %.3e\n", worst);
        if (!(worst < 4.0 * 1.1102230246251565e-16)) { ++bad; printf("end-to-end error too large\n"); }
    }
    printf("checked %lld bad %lld\n", checked, bad);
    return bad ? 1 : 0;
}
