// TMA + thread-block-cluster variant of the trailing update (sm_100a).
//
// Same mathematics and tile shape as syrk_kernel (CTA tile 128 x 64, DMMA.8x8x4, two CTAs per SM),
// but the operand pipeline is Blackwell-native:
//   * the two CTAs that own the left / right 64-column halves of one 128 x 128 tile form a CLUSTER of
//     2; they need the same A operand (L_i,*), so the leader CTA fetches each A chunk ONCE with a
//     multicast TMA load (cp.async.bulk.tensor ... .multicast::cluster) that lands in both CTAs'
//     shared memory -- one third less L2 -> SM traffic;
//   * a dedicated producer warp issues the TMA loads (one instruction per 128 x 16 box, hardware
//     128-byte swizzle), 8 consumer warps only do LDS + DMMA; stages are handed over with mbarriers
//     (full: TMA complete_tx bytes; empty: one arrival per consumer warp, the non-leader's warps also
//     arrive remotely on the leader's barrier because the leader overwrites both CTAs' A buffers).
// With the 128-byte swizzle (16-byte chunk index ^= row & 7) the DMMA fragment rows are assigned
// in the order 0,2,4,6,1,3,5,7 inside each group of 8 rows, which makes the 64-bit fragment loads of
// both operands bank-conflict free; the same permutation is applied when the C tile is read / written.
#pragma once
#include <cuda.h>      // CUtensorMap (types only; the encode function is fetched through the runtime)
#include "agpn_chol.cuh"

#define TMA_STAGE_BYTES ((TB + BN) * GEMM_KC * 8)                 // 24576
#define TMA_SMEM_BYTES (GEMM_STAGES * TMA_STAGE_BYTES + 1024)     // + slack for 1024-byte alignment
#define TMA_THREADS 288                                           // 8 consumer warps + 1 producer warp

__device__ __forceinline__ int perm8(int i) { return ((i & 3) << 1) | (i >> 2); }

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx(unsigned addr, unsigned bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned local_addr, unsigned target_cta) {
    asm volatile("{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n"
                 " mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}\n" ::"r"(local_addr), "r"(target_cta) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar,
                                               unsigned short mask) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
                 " [%0], [%1, {%2, %3, %4}], [%5], %6;\n"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar), "h"(mask) : "memory");
}

// grid.x = 2 * tiles (cluster = the two halves of a tile), grid.y = matrices of this launch
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TMA_THREADS, 2)
syrk_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, CholArgs g,
                int kfirst, int depth, int j0, int narrow, int mc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * GEMM_STAGES];

    const int t = blockIdx.x >> 1;
    const int h = blockIdx.x & 1;                     // == rank in the cluster
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
    const bool leader = rank == 0;

    const unsigned sbase = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    const unsigned full = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned empty = full + 8 * GEMM_STAGES;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, (leader && mc) ? 16 : 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    cluster_sync_all();

    const int nch = depth * TB / GEMM_KC;
    if (warp == 8) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const int matz = g.mat0 + blockIdx.y;
            for (int kc = 0; kc < nch; ++kc) {
                const int s = kc % GEMM_STAGES;
                const int u = kc / GEMM_STAGES;
                if (u > 0) mbar_wait(empty + 8 * s, (u - 1) & 1);     // leader: both CTAs released the stage
                mbar_expect_tx(full + 8 * s, TMA_STAGE_BYTES);
                const int kcol = kfirst * TB + kc * GEMM_KC;
                const unsigned dstA = sbase + s * TMA_STAGE_BYTES;
                if (mc) {
                    if (leader) tma_load_3d_mc(dstA, &tmA, kcol, ti * TB, matz, full + 8 * s, (unsigned short)3);
                } else {
                    tma_load_3d(dstA, &tmA, kcol, ti * TB, matz, full + 8 * s);
                }
                tma_load_3d(dstA + TB * GEMM_KC * 8, &tmB, kcol, tj * TB + h * BN, matz, full + 8 * s);
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int wm = warp >> 1, wn = warp & 1;
        const bool active = !(bi == bj && h == 1 && wm < 2);
        double* A = g.A0 + (size_t)(g.mat0 + blockIdx.y) * g.mat_stride;
        const int ld = g.ld;
        double* C = A + (size_t)ti * TB * ld + (size_t)tj * TB + h * BN;
        const int pr = perm8(lane >> 2);
        const int pc0 = perm8(2 * (lane & 3)), pc1 = perm8(2 * (lane & 3) + 1);
        double acc[4][4][2];
        if (active) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const double* cp = C + (size_t)(wm * 32 + mi * 8 + pr) * ld + wn * 32 + ni * 8;
                    acc[mi][ni][0] = -cp[pc0];
                    acc[mi][ni][1] = -cp[pc1];
                }
        }
        // per-thread swizzled byte offsets of its fragment element for the 4 k-steps of a chunk
        unsigned koff[GEMM_KC / 4];
#pragma unroll
        for (int ks = 0; ks < GEMM_KC / 4; ++ks) {
            const int k = ks * 4 + (lane & 3);
            koff[ks] = (unsigned)(pr * 128 + (((k >> 1) ^ pr) << 4) + (k & 1) * 8);
        }
        // plain (non-volatile) shared loads so that the compiler can hoist the next k-step's fragment
        // loads above the current step's DMMAs (an asm volatile ld.shared exposed 10 % LDS stalls)
        const unsigned char* sgen = smem_raw + (sbase - (unsigned)__cvta_generic_to_shared(smem_raw));
        const unsigned a_row0 = (unsigned)(wm * 32) * 128u;
        const unsigned b_row0 = (unsigned)(TB * GEMM_KC * 8) + (unsigned)(wn * 32) * 128u;
        for (int kc = 0; kc < nch; ++kc) {
            const int s = kc % GEMM_STAGES;
            mbar_wait(full + 8 * s, (kc / GEMM_STAGES) & 1);
            if (active) {
                const unsigned char* st = sgen + s * TMA_STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < GEMM_KC / 4; ++ks) {
                    double a[4], b[4];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = *reinterpret_cast<const double*>(st + a_row0 + mi * 1024u + koff[ks]);
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) b[ni] = *reinterpret_cast<const double*>(st + b_row0 + ni * 1024u + koff[ks]);
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty + 8 * s);
                if (mc && !leader) mbar_arrive_remote(empty + 8 * s, 0);
            }
        }
        if (active) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    double* cp = C + (size_t)(wm * 32 + mi * 8 + pr) * ld + wn * 32 + ni * 8;
                    cp[pc0] = -acc[mi][ni][0];
                    cp[pc1] = -acc[mi][ni][1];
                }
        }
    }
    // nobody leaves while the peer may still multicast into / arrive on this CTA's shared memory
    if (mc) cluster_sync_all();
}
