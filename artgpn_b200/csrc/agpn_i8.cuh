// Trailing update of the blocked Cholesky on the INT8 tensor path (tcgen05.mma.kind::i8, TMEM accumulators):
// the FP64 rank-k update  C_ij -= sum_k L_ik L_jk  evaluated as exact integer products of 7-bit slices
// (Ozaki-style error-free splitting) -- several times the DMMA.8x8x4 rate of the FP64 tensor pipe.
//
// Numbers.  Row i of L is scaled by r_i >= max_k |L_ik| (r_i = sqrt(K_ii): sum_k L_ik^2 = K_ii) and stored as the
// integer q_ik = rint(L_ik / r_i * Q), Q = 126 * 2^49, split into 8 signed base-128 digits a_0 .. a_7
// (q = sum_s a_s 128^(7-s), |a_s| <= 64 for s > 0, |a_0| <= 127): 56 bits below the row scale.  Then
//     sum_k q_ik q_jk = sum_d 128^(14-d) S_d,     S_d = sum_{s+t=d} sum_k a_s(i,k) a_t(j,k)    (exact in int32),
// and the diagonals d = 0 .. 7 (36 slice pairs) carry everything above 2^-53 of r_i r_j (the dropped d >= 8 terms
// are ~2^-53 r_i r_j rms at K = 1024).  One TMEM accumulator (128 lanes x 64 columns of int32) per diagonal: 8 x 64 =
// all 512 TMEM columns, which fixes the CTA tile at 128 x 64.  Consecutive B slices are consecutive rows of the
// canonical K-major shared-memory layout, so ONE tcgen05.mma with N = 64 n multiplies A_s with slices t0 .. t0+n-1
// and accumulates into the n adjacent accumulators s+t0 .. s+t0+n-1: 12 MMAs per 32-byte K step instead of 36.
// The epilogue folds the 8 integer diagonals into one double (H = S_0..S_3 and Lo = S_4..S_7 are exact in a double,
// val = H 2^28 + Lo is rounded once) and applies r_i r_j / Q^2.
//
// Data movement.  The panel solve (trsm_kernel, agpn_chol.cuh; or i8_slice_kernel below with AGPN_I8_FUSE=0) writes
// the digits in the tensor core's own operand layout (UMMA canonical
// K-major, no swizzle: 8-row x 16-byte core matrices, LBO = 128 B, SBO = 256 B), tile by tile:
//     slices[mat][kc32 = (column - first column of the group) / 32][row tile of 128][slice s][4096 B],
//     in-tile (r, kb) -> (r/8) 256 + (kb/16) 128 + (r%8) 16 + kb%16
// (only the CURRENT group's block columns are kept: every update reads columns of the group being eliminated),
// so a stage is filled by plain 1-D bulk copies (cp.async.bulk, no tensor map): 32 KB of A (all 8 slices of 128
// rows, contiguous) and 8 x 2 KB of B.  The two CTAs that own the left / right 64 columns of a 128 x 128 tile form a
// cluster and need the same A: each fetches half of it and multicasts it to both (L2 -> SM bytes per CTA and K step:
// 32 instead of 48 KB).  3 stages x 48 KB; warp 0 = producer, warp 1 = MMA issuer, warps 2.. = epilogue
// (TMEM lane quarter = warp % 4; the persistent kernel runs eight: two per quarter, each half of the 64 columns); tcgen05.commit (multicast to both CTAs' `empty` barriers) hands stages back.
// The C tile travels by bulk copies too: every epilogue thread fetches ITS row (512 contiguous bytes) into a
// padded shared-memory tile before the main loop starts, folds the eight integer diagonals of that row into it
// after the last MMA and sends the row back with one bulk store -- no thread-per-row strided global access, no
// barrier between the epilogue threads, and the C latency hides behind the main loop.
// Two kernels: syrk_i8_kernel (one CTA pair per tile) and syrk_i8_persistent_kernel (default: 74 clusters walk over the
// (matrix, tile) pairs of a launch, the producer runs ahead across tile boundaries).  The reference computes the same
// update inside LAPACK dpotrf (scipy cho_factor, arttwo.py:284).
// Measured building blocks: tools/i8_probe.cu, profiles/r01_i8_probe_*.txt; the product kernels: profiles/r01_i8_ncu_summary.md.
#pragma once
#include <cstring>
#include "agpn_chol.cuh"
#include "agpn_tma.cuh"     // cluster helpers, mbar_expect_tx

#define I8_KB 32                                                 // bytes of K per stage = one kind::i8 K step
#ifndef I8_STAGES
#define I8_STAGES 3
#endif
#define I8_A_BYTES (TB * I8_KB)                                  // 4096 per slice
#define I8_B_BYTES (BN * I8_KB)                                  // 2048 per slice
#define I8_STAGE_BYTES (I8_NSL * (I8_A_BYTES + I8_B_BYTES))      // 49152
#define I8_CROW_BYTES (BN * 8 + 16)                              // padded C row: 16-byte accesses of a warp, one row per lane, conflict free
#define I8_SMEM_BYTES (I8_STAGES * I8_STAGE_BYTES + TB * I8_CROW_BYTES + 1024)   // + slack for 1024-byte alignment
#define I8_THREADS 192                                           // per-tile kernel, peak probe
#ifndef I8_EPW
#define I8_EPW 8                                                 // epilogue warps of the persistent kernel: 4 or 8 (two per TMEM lane quarter)
#endif
#define I8_PTHREADS (64 + 32 * I8_EPW)
#ifndef I8_PREFETCH
#define I8_PREFETCH 0
#endif

// ---- row scales r_i >= max_k |L_ik|: qmul = Q / r, e = r / (126 2^24) ---------------------------------------------
// rows i < n_sq: r_i = sqrt(K_ii) from the diagonal of the assembled matrix (sum_k L_ik^2 = K_ii);
// rows n_sq <= i (the K* rows appended for prediction): r_i = sqrt(k**_ii) from `ext` (|K* L^-T|_i^2 <= k**_ii)
// grid (n_rows / 256, nmat)
__global__ void __launch_bounds__(256) i8_rowscale_kernel(CholArgs g, int n_sq, const double* __restrict__ ext, int n_ext) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = g.nblk * TB;
    if (i >= n) return;
    double d = 1.0;
    if (i < n_sq) d = g.A[(size_t)blockIdx.y * g.mat_stride + (size_t)i * g.ld + i];
    else if (ext != nullptr && i - n_sq < n_ext) d = ext[i - n_sq];
    // a block that is not positive definite / not finite is flagged by the factorisation; its slices only have to be harmless
    const double r = (d > 0.0 && d < 1.0e300) ? sqrt(d) * (1.0 + 1.0e-9) : 1.0;
    double* rs = g.rscale + (size_t)blockIdx.y * 2 * n;
    rs[i] = I8_Q / r;
    rs[n + i] = r * (1.0 / (126.0 * 16777216.0));
}

// sum of squares of the rows of the K* L^-T block (rows row0 .., n_cols columns) and the predictive mean / std
// (arttwo.py:388-395): mean = m(t*) + K* K^-1 r = m* - rhs_aug ; var = k** + s^2 - |K* L^-T|_row^2
__global__ void __launch_bounds__(256) predict_finish_kernel(const double* __restrict__ V, int ld, int n_cols, int M,
                                                             const double* __restrict__ rhs_aug, const double* __restrict__ mstar,
                                                             const double* __restrict__ kss, const double* __restrict__ theta,
                                                             int jit_index, double* __restrict__ mean_out, double* __restrict__ std_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= M) return;
    const double* v = V + (size_t)row * ld;
    double s = 0.0;
    for (int c = 2 * lane; c < n_cols; c += 64) {
        const double2 x = *reinterpret_cast<const double2*>(v + c);
        s = fma(x.x, x.x, s);
        s = fma(x.y, x.y, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const double jt = theta[jit_index];
        mean_out[row] = mstar[row] - rhs_aug[row];
        std_out[row] = sqrt(kss[row] + jt * jt - s);
    }
}

// ---- slicing of block column k (rows below the diagonal block): L -> 8 int8 digit planes --------------------------
// (k0 = first block column of k's group)  grid (nblk - k - 1, nmat), 256 threads; a thread task = 16 consecutive columns of one row (one 16-byte chunk per slice)
__global__ void __launch_bounds__(256) i8_slice_kernel(CholArgs g, int k, int k0) {
    const int rt = k + 1 + blockIdx.x;
    const int n = g.nblk * TB;
    const double* A = g.A + (size_t)blockIdx.y * g.mat_stride;
    const double* qmul = g.rscale + (size_t)blockIdx.y * 2 * n;
    signed char* S = g.S8 + (size_t)blockIdx.y * i8_slices_per_matrix(g.nblk, g.sgroup);
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int id = it * 256 + threadIdx.x;
        const int r8 = id & 7, ch = (id >> 3) & 7, rg = id >> 6;       // ch: 16-column chunk 0..7 of the block column
        const int row = rg * 8 + r8;
        const double f = qmul[rt * TB + row];
        const double* src = A + (size_t)(rt * TB + row) * g.ld + (size_t)k * TB + ch * 16;
        unsigned w[I8_NSL][4];
#pragma unroll
        for (int c2 = 0; c2 < 8; ++c2) {
            const double2 v = *reinterpret_cast<const double2*>(src + 2 * c2);
            unsigned piece[I8_NSL];
            i8_digits2(v.x * f, v.y * f, piece);
#pragma unroll
            for (int s = 0; s < I8_NSL; ++s) {
                if (c2 & 1) w[s][c2 >> 1] |= piece[s] << 16;
                else w[s][c2 >> 1] = piece[s];
            }
        }
        const int kc32 = 4 * (k - k0) + (ch >> 1);
        signed char* dst = S + i8_tile_off(kc32, rt, g.nblk) + (size_t)rg * 256 + (ch & 1) * 128 + r8 * 16;
#pragma unroll
        for (int s = 0; s < I8_NSL; ++s)
            *reinterpret_cast<uint4*>(dst + s * I8_A_BYTES) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

// ---- tcgen05 / bulk-copy wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(unsigned dst, const void* src, unsigned bytes, unsigned mbar, unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "h"(mask) : "memory");
}
// K-major, no swizzle, LBO = 128 B, SBO = 256 B, descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long i8_desc(unsigned saddr) {
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(unsigned mbar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
// wait for a barrier that completes only after milliseconds of tensor work (the trap of mbar_wait is for protocol bugs)
__device__ __forceinline__ void mbar_wait_long(unsigned addr, unsigned parity) {
    unsigned done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++spins > (1ull << 32)) __trap();
    }
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// exact int32 -> double on the FP64 add pipe: the bits of 2^52 + 2^31 + v, minus the bias (no I2F)
__device__ __forceinline__ double i8_i2d(int v) {
    return __hiloint2double(0x43300000, (int)((unsigned)v ^ 0x80000000u)) - 4503601774854144.0;
}

// four int32 diagonals -> the double ((s0 128 + s1) 128 + s2) 128 + s3, exactly: folded in 64-bit integers (|.| < 2^47 for
// K <= 32768) on the integer pipe, converted by adding the bits of 1.5 * 2^52 and subtracting it again (one DADD)
__host__ __device__ __forceinline__ double i8_fold4(int s0, int s1, int s2, int s3) {
    const long long h = (((long long)s0 * 128 + s1) * 128 + s2) * 128 + s3;
#ifdef __CUDA_ARCH__
    return __longlong_as_double(h + 0x4338000000000000ll) - 6755399441055744.0;
#else
    const long long bits = h + 0x4338000000000000ll;
    double d;
    memcpy(&d, &bits, sizeof(d));
    return d - 6755399441055744.0;
#endif
}

// A slice sa times B slices tb .. tb+n-1 of one K step: the schedule that covers the 36 pairs with 12 MMAs
#define I8_MMA(sa, tb, n, first)                                                                                         \
    mma_i8(tmem + ((sa) + (tb)) * BN, i8_desc(a0 + (sa) * I8_A_BYTES), i8_desc(b0 + (tb) * I8_B_BYTES),                     \
           idesc0 | ((unsigned)((BN * (n)) >> 3) << 17), (first) ? accum0 : 1u)

// same launch geometry as syrk_kernel / syrk_tma_kernel:
//   grid.x = 2 * tiles (cluster = left / right half of one 128 x 128 tile), grid.y = matrices
//   kfirst = first block column of the update = first block column of the slice storage (the group start)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
syrk_i8_kernel(CholArgs g, int kfirst, int depth, int j0, int narrow) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * I8_STAGES + 2];
    __shared__ unsigned tmem_base_s;
    __shared__ double ecol[BN];
    (void)kfirst;

    const int t = blockIdx.x >> 1;
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
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_ctarank();
    const int h = (int)rank;                                        // which 64-column half of the tile
    const int n = g.nblk * TB;
    const double* esc = g.rscale + (size_t)blockIdx.y * 2 * n + n;

    const unsigned sraw = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned sbase = (sraw + 1023u) & ~1023u;
    const unsigned cs = sbase + I8_STAGES * I8_STAGE_BYTES;         // C tile, rows of I8_CROW_BYTES
    const unsigned full = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned empty = full + 8 * I8_STAGES, done = full + 16 * I8_STAGES, cfull = done + 8;
    // right half of a diagonal tile: rows 0..63 lie strictly above the diagonal, never read or written
    const bool half_tile = (bi == bj && h == 1);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 2);                            // one tcgen05.commit from each CTA of the pair
        }
        mbar_init(done, 1);
        mbar_init(cfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < BN) ecol[tid] = 0.5 * esc[tj * TB + h * BN + tid];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                             // both CTAs' barriers exist before any multicast / remote commit
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem = tmem_base_s;
    const int nks = depth * (TB / I8_KB);

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const signed char* S = g.S8 + (size_t)blockIdx.y * i8_slices_per_matrix(g.nblk, g.sgroup);
            for (int ks = 0; ks < nks; ++ks) {
                const int s = ks % I8_STAGES, u = ks / I8_STAGES;
                if (u > 0) mbar_wait(empty + 8 * s, (u - 1) & 1);
                mbar_expect_tx(full + 8 * s, I8_STAGE_BYTES);
                const unsigned dst = sbase + s * I8_STAGE_BYTES;
#if I8_PREFETCH > 0
                if (ks + I8_PREFETCH < nks) {                      // pull a later K step's digit planes into L2
                    bulk_prefetch_l2(S + i8_tile_off(ks + I8_PREFETCH, ti, g.nblk) + rank * (4 * I8_A_BYTES), 4 * I8_A_BYTES);
                    bulk_prefetch_l2(S + i8_tile_off(ks + I8_PREFETCH, tj, g.nblk) + rank * (4 * I8_A_BYTES), 4 * I8_A_BYTES);
                }
#endif
                // my half of the A stage (slices 4 rank .. 4 rank + 3), delivered to both CTAs
                bulk_g2s_mc(dst + rank * (4 * I8_A_BYTES), S + i8_tile_off(ks, ti, g.nblk) + rank * (4 * I8_A_BYTES), 4 * I8_A_BYTES,
                            full + 8 * s, (unsigned short)3);
                const signed char* bsrc = S + i8_tile_off(ks, tj, g.nblk) + h * I8_B_BYTES;
#pragma unroll
                for (int sl = 0; sl < I8_NSL; ++sl)
                    bulk_g2s(dst + I8_NSL * I8_A_BYTES + sl * I8_B_BYTES, bsrc + sl * I8_A_BYTES, I8_B_BYTES, full + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            // kind::i8: D = s32 (2 << 4), A, B = signed 8 bit (1 << 7, 1 << 10), K-major both, M = 128
            const unsigned idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TB >> 4) << 24);
            for (int ks = 0; ks < nks; ++ks) {
                const int s = ks % I8_STAGES;
                mbar_wait(full + 8 * s, (ks / I8_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned a0 = sbase + s * I8_STAGE_BYTES, b0 = a0 + I8_NSL * I8_A_BYTES;
                const unsigned accum0 = ks > 0 ? 1u : 0u;
                I8_MMA(0, 0, 4, true);
                I8_MMA(0, 4, 4, true);
                I8_MMA(1, 0, 4, false);
                I8_MMA(1, 4, 3, false);
                I8_MMA(2, 0, 4, false);
                I8_MMA(2, 4, 2, false);
                I8_MMA(3, 0, 4, false);
                I8_MMA(3, 4, 1, false);
                I8_MMA(4, 0, 4, false);
                I8_MMA(5, 0, 3, false);
                I8_MMA(6, 0, 2, false);
                I8_MMA(7, 0, 1, false);
                umma_commit_mc(empty + 8 * s, (unsigned short)3);
            }
            umma_commit(done);
        }
    } else {
        // ---------------- epilogue: thread = one row of the tile (TMEM lane) ----------------
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool skip = half_tile && q < 2;                       // warp-uniform
        double* Crow = g.A + (size_t)blockIdx.y * g.mat_stride + (size_t)(ti * TB + row) * g.ld + (size_t)tj * TB + h * BN;
        const unsigned crow = cs + row * I8_CROW_BYTES;
        if (warp == 2 && lane == 0) mbar_expect_tx(cfull, (half_tile ? TB / 2 : TB) * BN * 8);
        if (!skip) {
            bulk_g2s(crow, Crow, BN * 8, cfull);
            const double ei = esc[ti * TB + row];
            double* cgen = reinterpret_cast<double*>(smem_raw + (crow - sraw));
            mbar_wait(done, 0);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            mbar_wait(cfull, 0);
            __syncwarp();
            const unsigned tlane = tmem + ((unsigned)(q * 32) << 16);
#pragma unroll 1
            for (int c8 = 0; c8 < BN / 8; ++c8) {
                int v[I8_NSL][8];
#pragma unroll
                for (int d = 0; d < I8_NSL; ++d) tmem_ld8(tlane + d * BN + c8 * 8, v[d]);
                tmem_ld_wait();
#pragma unroll
                for (int j2 = 0; j2 < 8; j2 += 2) {
                    double2 cc = *reinterpret_cast<double2*>(cgen + c8 * 8 + j2);
                    double r2[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int j = j2 + u;
                        const double H = i8_fold4(v[0][j], v[1][j], v[2][j], v[3][j]);
                        const double Lo = i8_fold4(v[4][j], v[5][j], v[6][j], v[7][j]);
                        const double val = fma(H, 268435456.0, Lo);      // H 2^28 + Lo = sum_d 128^(7-d) S_d, one rounding
                        r2[u] = (val * ei) * ecol[c8 * 8 + j];
                    }
                    cc.x -= r2[0];
                    cc.y -= r2[1];
                    *reinterpret_cast<double2*>(cgen + c8 * 8 + j2) = cc;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            bulk_s2g(Crow, crow, BN * 8);
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();        // nobody leaves while the peer may still multicast into / commit onto this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---- INT8 tensor-path issue peak (roofline denominator measured in the same run, like dmma_peak_kernel) ------------
// one CTA per SM; operands resident in shared memory (zeros), M = 128, N = 256, K = 32 kind::i8 MMAs back to back
// into the two halves of the 512 TMEM columns
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, int* sink) {
    __shared__ __align__(1024) unsigned char ops[TB * I8_KB + 256 * I8_KB];      // A 4 KB | B 8 KB
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (int)sizeof(ops) / 4; i += blockDim.x) reinterpret_cast<unsigned*>(ops)[i] = 0x01010101u;
    const unsigned done = (unsigned)__cvta_generic_to_shared(&bar);
    if (tid == 0) {
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem = tmem_base_s;
    if (tid == 0) {
        const unsigned a0 = (unsigned)__cvta_generic_to_shared(ops), b0 = a0 + TB * I8_KB;
        const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(256 >> 3) << 17) | ((unsigned)(TB >> 4) << 24);
        const unsigned long long ad = i8_desc(a0), bd = i8_desc(b0);
        for (int it = 0; it < iters; ++it) mma_i8(tmem + (it & 1) * 256, ad, bd, idesc, it > 1 ? 1u : 0u);
        umma_commit(done);
    }
    mbar_wait_long(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    int v[8];
    tmem_ld8(tmem + ((unsigned)(warp * 32) << 16), v);
    tmem_ld_wait();
    if (sink && v[0] == -12345) sink[blockIdx.x * blockDim.x + tid] = v[1];     // keeps the accumulators observable
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---- persistent variant: a cluster walks over many tiles ------------------------------------------------------------
// Same tile kernel, but the per-tile fixed cost (CTA launch, TMEM allocation, barrier set-up, two cluster syncs and
// above all the pipeline fill from L2 / DRAM) is paid once per CTA instead of once per tile: the producer warp runs
// ahead into the next tile's K steps while the epilogue warps drain the accumulators, the stage / phase counters run
// on across tiles, and one extra mbarrier (`tfree`) tells the MMA thread when TMEM may be overwritten.
//   grid.x = 2 * clusters (<= one CTA per SM), tiles of the launch: (matrix, tile) pairs, tile index fastest
//   ntiles = tiles per matrix: narrow: tiles (j0 + bi0 + t, j0), t < ntiles; wide: nt (nt + 1) / 2
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_PTHREADS, 1)
syrk_i8_persistent_kernel(CholArgs g, int depth, int j0, int narrow, int ntiles, int total, int bi0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * I8_STAGES + 3];
    __shared__ unsigned tmem_base_s;
    __shared__ double ecol[2][BN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_ctarank();
    const int h = (int)rank;                                        // which 64-column half of the tile
    const int n = g.nblk * TB;
    const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    const unsigned sraw = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned sbase = (sraw + 1023u) & ~1023u;
    const unsigned cs = sbase + I8_STAGES * I8_STAGE_BYTES;         // C tile, rows of I8_CROW_BYTES
    const unsigned full = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned empty = full + 8 * I8_STAGES, done = full + 16 * I8_STAGES, cfull = done + 8, tfree = done + 16;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 2);                            // one tcgen05.commit from each CTA of the pair
        }
        mbar_init(done, 1);
        mbar_init(cfull, 1);
        mbar_init(tfree, I8_EPW);                                   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                             // both CTAs' barriers exist before any multicast / remote commit
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem = tmem_base_s;
    const int nks = depth * (TB / I8_KB);

    // Round r of the walk hands tile r * ncl + c' to cluster (c' + r) mod ncl (g.tile_order == 2): with the plain c' = cluster
    // assignment a cluster only ever sees the tile rows t = (cluster + 74 r) mod ntiles -- half of them when ntiles is even -- and the
    // clusters that keep drawing the tiles next to the diagonal finish last; the skew makes every cluster cycle through all rows
    const bool skew = g.tile_order == 2;
    // tile w -> (matrix, block row ti, block column tj)
    auto decode = [&](int w, int& mat, int& ti, int& tj, bool& diag) {
        int t;
        if (g.tile_order == 1 && narrow) {
            // matrix index fastest: the tiles next to the diagonal (the ones with active K steps in a banded factor) of all
            // matrices come first and spread evenly over the clusters, the far tiles that are only looked at come last
            const int nm = total / ntiles;
            t = w / nm;
            mat = w - t * nm;
        } else {
            mat = w / ntiles;
            t = w - mat * ntiles;
        }
        int bi, bj;
        if (narrow) {
            bi = t + bi0;                                           // bi0 = 1: the tiles below the diagonal tile only
            bj = 0;
        } else {
            bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
            while (bi * (bi + 1) / 2 > t) --bi;
            bj = t - bi * (bi + 1) / 2;
        }
        ti = j0 + bi;
        tj = j0 + bj;
        diag = bi == bj;
    };

    if (warp == 0) {
        // ---------------- producer: runs ahead across tile boundaries ----------------
        if (lane == 0) {
            unsigned it = 0;                                        // K steps issued so far (all tiles)
            for (int rnd = 0; rnd * ncl < total; ++rnd) {
                const int w = rnd * ncl + (skew ? (cl + ncl - rnd % ncl) % ncl : cl);
                if (w >= total) continue;
                int mat, ti, tj;
                bool diag;
                decode(w, mat, ti, tj, diag);
                const signed char* S = g.S8 + (size_t)mat * i8_slices_per_matrix(g.nblk, g.sgroup);
                for (int wd = 0; wd * 64 < nks; ++wd) {
                    unsigned long long act = i8_active_word(g, mat, ti, tj, wd, nks);     // K steps with non-zero digits on both sides
                    while (act) {
                        const int ks = wd * 64 + __ffsll((long long)act) - 1;
                        act &= act - 1ull;
                        const unsigned s = it % I8_STAGES, u = it / I8_STAGES;
                        if (u > 0) mbar_wait(empty + 8 * s, (u - 1) & 1);
                        mbar_expect_tx(full + 8 * s, I8_STAGE_BYTES);
                        const unsigned dst = sbase + s * I8_STAGE_BYTES;
                        bulk_g2s_mc(dst + rank * (4 * I8_A_BYTES), S + i8_tile_off(ks, ti, g.nblk) + rank * (4 * I8_A_BYTES), 4 * I8_A_BYTES,
                                    full + 8 * s, (unsigned short)3);
                        const signed char* bsrc = S + i8_tile_off(ks, tj, g.nblk) + h * I8_B_BYTES;
#pragma unroll
                        for (int sl = 0; sl < I8_NSL; ++sl)
                            bulk_g2s(dst + I8_NSL * I8_A_BYTES + sl * I8_B_BYTES, bsrc + sl * I8_A_BYTES, I8_B_BYTES, full + 8 * s);
                        ++it;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const unsigned idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TB >> 4) << 24);
            unsigned it = 0, tile = 0;
            for (int rnd = 0; rnd * ncl < total; ++rnd) {
                const int w = rnd * ncl + (skew ? (cl + ncl - rnd % ncl) % ncl : cl);
                if (w >= total) continue;
                int mat, ti, tj;
                bool diag;
                decode(w, mat, ti, tj, diag);
                if (!i8_any_active(g, mat, ti, tj, nks)) continue;   // nothing but exact zeros to add: the tile is left alone
                if (tile > 0) {                                     // the epilogue has drained the previous tile's accumulators
                    mbar_wait(tfree, (tile - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                bool first_step = true;
                for (int wd = 0; wd * 64 < nks; ++wd) {
                  unsigned long long act = i8_active_word(g, mat, ti, tj, wd, nks);
                  while (act) {
                    act &= act - 1ull;
                    const unsigned s = it % I8_STAGES;
                    mbar_wait(full + 8 * s, (it / I8_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const unsigned a0 = sbase + s * I8_STAGE_BYTES, b0 = a0 + I8_NSL * I8_A_BYTES;
                    const unsigned accum0 = first_step ? 0u : 1u;
                    first_step = false;
                    I8_MMA(0, 0, 4, true);
                    I8_MMA(0, 4, 4, true);
                    I8_MMA(1, 0, 4, false);
                    I8_MMA(1, 4, 3, false);
                    I8_MMA(2, 0, 4, false);
                    I8_MMA(2, 4, 2, false);
                    I8_MMA(3, 0, 4, false);
                    I8_MMA(3, 4, 1, false);
                    I8_MMA(4, 0, 4, false);
                    I8_MMA(5, 0, 3, false);
                    I8_MMA(6, 0, 2, false);
                    I8_MMA(7, 0, 1, false);
                    umma_commit_mc(empty + 8 * s, (unsigned short)3);
                    ++it;
                  }
                }
                umma_commit(done);
                ++tile;
            }
        }
    } else {
        // ---------------- epilogue: thread = one row of the tile (TMEM lane) ----------------
        const int q = warp & 3;                                     // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const int et = tid - 64;                                    // 0 .. 32 I8_EPW - 1 among the epilogue threads
        constexpr int NCH = I8_EPW / 4;                             // warps per lane quarter: each takes BN / NCH columns of the row
        constexpr int CW = BN / NCH;
        const int ch = (warp - 2) >> 2;
        const unsigned crow = cs + row * I8_CROW_BYTES + ch * (CW * 8);
        double* cgen = reinterpret_cast<double*>(smem_raw + (crow - sraw));
        const unsigned tlane = tmem + ((unsigned)(q * 32) << 16);
        unsigned tile = 0;
        for (int rnd = 0; rnd * ncl < total; ++rnd) {
            const int w = rnd * ncl + (skew ? (cl + ncl - rnd % ncl) % ncl : cl);
            if (w >= total) continue;
            int mat, ti, tj;
            bool diag;
            decode(w, mat, ti, tj, diag);
            if (!i8_any_active(g, mat, ti, tj, nks)) continue;       // same decision as the producer and the MMA thread
            const bool half_tile = diag && h == 1;                  // rows 0..63 lie strictly above the diagonal
            const bool skip = half_tile && q < 2;                   // warp-uniform
            const double* esc = g.rscale + (size_t)mat * 2 * n + n;
            double* Crow = g.A + (size_t)mat * g.mat_stride + (size_t)(ti * TB + row) * g.ld + (size_t)tj * TB + h * BN + ch * CW;
            const int eb = tile & 1;
            if (et < BN) ecol[eb][et] = 0.5 * esc[tj * TB + h * BN + et];
            if (et == 0) mbar_expect_tx(cfull, (half_tile ? TB / 2 : TB) * BN * 8);
            if (!skip) bulk_g2s(crow, Crow, CW * 8, cfull);
            const double ei = esc[ti * TB + row];
            const bool record = narrow && g.tmaxU != nullptr;       // the panel solve's bound test reads the tile's maximum
            const double qrow = record ? g.rscale[(size_t)mat * 2 * n + ti * TB + row] : 0.0;     // Q / r_i
            double rmax = 0.0;
            named_bar_sync(1, 32 * I8_EPW);                         // ecol[eb] visible to all epilogue warps
            mbar_wait(done, tile & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            mbar_wait(cfull, tile & 1);                             // every epilogue thread: nobody may post the next tile's bytes into this phase
            if (!skip) {
#pragma unroll 1
                for (int c8 = 0; c8 < CW / 8; ++c8) {
                    const int cc0 = ch * CW + c8 * 8;               // first of these eight columns inside the CTA's 64
                    int v[I8_NSL][8];
#pragma unroll
                    for (int d = 0; d < I8_NSL; ++d) tmem_ld8(tlane + d * BN + cc0, v[d]);
                    tmem_ld_wait();
                    if (c8 == CW / 8 - 1) {                         // last read of the accumulators: hand TMEM back
                        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tfree);
                    }
#pragma unroll
                    for (int j2 = 0; j2 < 8; j2 += 2) {
                        double2 cc = *reinterpret_cast<double2*>(cgen + c8 * 8 + j2);
                        double r2[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int j = j2 + u;
                            const double H = i8_fold4(v[0][j], v[1][j], v[2][j], v[3][j]);
                            const double Lo = i8_fold4(v[4][j], v[5][j], v[6][j], v[7][j]);
                            const double val = fma(H, 268435456.0, Lo);
                            r2[u] = (val * ei) * ecol[eb][cc0 + j];
                        }
                        cc.x -= r2[0];
                        cc.y -= r2[1];
                        *reinterpret_cast<double2*>(cgen + c8 * 8 + j2) = cc;
                        const double m2 = fabs(cc.x) + fabs(cc.y);  // bounds the larger one and keeps a NaN
                        rmax = (m2 > rmax || m2 != m2) ? m2 : rmax;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                bulk_s2g(Crow, crow, CW * 8);
                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                if (record) {
                    // an upper bound is all the reader needs: round up to a float, whose bit patterns order like the values for
                    // non-negative numbers (NaN above them all) -- one redux.sync per warp, widened back to a double's bits
                    const unsigned fb = __reduce_max_sync(0xffffffffu, __float_as_uint(__double2float_ru(rmax * qrow)));
                    if (lane == 0) {
                        const double v = (double)__uint_as_float(fb);
                        atomicMax(g.tmaxU + ((size_t)mat * g.nblk + ti) * g.nblk + tj, (unsigned long long)__double_as_longlong(v));
                    }
                }
                asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");       // the row buffer is free for the next tile's load
            } else {
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tfree);
            }
            ++tile;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();        // nobody leaves while the peer may still multicast into / commit onto this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}
