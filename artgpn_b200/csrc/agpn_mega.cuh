// One persistent kernel for the whole left-looking factorisation of a batch (replaces the per-column launch chain
// potrf_diag_kernel / trsm_kernel / syrk_i8_persistent_kernel of run_cholesky for the INT8-sliced path).
//
// Why.  With few matrices per GPU (the 8-GPU strong-scaling shard: 16 walkers = 48 matrices; the reference's own caller
// runs half-ensembles of 15, examples/Sun/01_emcee.py:111-122) the per-column chain of three launches is latency bound:
// every launch waits for the slowest CTA of the previous one, the one-CTA-per-matrix diagonal blocks leave most SMs idle,
// and the persistent update CTAs of one sub-batch stream keep the SMs the next stream's diagonal blocks are waiting for
// (measured: 3.2 ms per 16-walker step against 21.1 / 8 = 2.6 ms).  Here the unit of scheduling is a TILE TASK with
// explicit dependencies instead of a launch:
//     D(k)        diagonal block k of a matrix: factor, inverse, forward-solve block   (potrf_diag_body)
//     P(i, k)     panel solve of tile (i, k), i > k, digit planes, rhs update           (trsm_body)
//     U(i, k)     left-looking update of tile (i, k) with block columns 0 .. k-1        (tcgen05 kind::i8, as agpn_i8.cuh)
// The tasks of a batch form one fixed sequence in which every dependency points backwards:
//     D(0);  for k = 0, 1, ..:  P(., k);  U(k+1, k+1);  D(k+1);  U(i > k+1, k+1)
// (all matrices inside every segment; U(k+1, k+1) and D(k+1) come BEFORE the rest of column k+1, so the latency-bound
// diagonal block of the next column overlaps the bulk of the update -- look-ahead falls out of the order).  Cluster c of
// the grid takes tasks c, c + nclusters, ..: a cluster waits only for tasks earlier in the sequence, which are owned by
// clusters that never wait for anything later, so the schedule cannot deadlock as long as every cluster eventually runs
// (one CTA per SM, grid <= SM count).  Readiness travels through three arrays of monotone counters in global memory
// (release / acquire at GPU scope): fD[mat] = diagonal blocks done, fP[mat][row] = panels done in that block row,
// fU[mat][row] = half tiles updated in that block row.
// Inside a cluster consecutive U tasks run exactly like syrk_i8_persistent_kernel (producer warp / MMA warp / four
// epilogue warps, the producer running ahead across tile boundaries); a change of task type drains the pipeline and
// re-uses the same shared memory for the FP64 kernels' bodies.  D and P tasks are CTA / thread-group tasks: the two CTAs of
// a cluster take two matrices (D), their four groups of 256 threads four tiles (P: two tiles in flight per SM, so one
// tile's loads, digit extraction and stores overlap the other's tensor work like two co-resident CTAs of trsm_kernel).
#pragma once
#include "agpn_i8.cuh"

#define MEGA_THREADS 512                        // two groups of 256: the FP64 bodies work on two tiles at once
#define MEGA_SMEM_BYTES I8_SMEM_BYTES            // >= POTRF_SMEM_BYTES >= GEMM_SMEM_BYTES

struct MegaArgs {
    int* fD;            // [nmat]
    int* fP;            // [nmat][nrows]
    int* fU;            // [nmat][nrows]
    int ncols;          // block columns to eliminate (<= nrows = g.nblk; rows beyond are appended K* rows)
    int tflags;         // trsm flags: 1 (digit planes) | 2 (FP64 panel not stored)
    unsigned long long* prof;   // nullptr, or [CTAs][16] nanosecond / count slots (AGPN_MEGA_PROF=1 diagnostics):
                                // 0 U runs, 1 U tasks, 2 D body, 3 D wait, 4 D tasks, 5 P body, 6 P wait, 7 P tasks (group 0),
                                // 8 whole kernel, 9 cluster sync before U runs, 10 producer waits for panels
};

__device__ __forceinline__ unsigned long long mega_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

enum : int { MT_D = 0, MT_P = 1, MT_UD = 2, MT_UR = 3, MT_END = 4 };

// position in the task sequence; advance() is amortised O(1) because task indices only grow
struct MegaCursor {
    int nmat, nrows, ncols;
    int k;              // column of the current segment (D(k), P(., k), U(., k))
    int type;
    long long base;     // index of the segment's first task
    int count;
    __device__ __forceinline__ int seg_count() const {
        switch (type) {
            case MT_D: return (nmat + 1) / 2;
            case MT_P: return (nmat * (nrows - k - 1) + 3) / 4;
            case MT_UD: return nmat;
            case MT_UR: return nmat * (nrows - k - 1);
            default: return 0;
        }
    }
    __device__ __forceinline__ void init(int nmat_, int nrows_, int ncols_) {
        nmat = nmat_; nrows = nrows_; ncols = ncols_;
        k = 0; type = MT_D; base = 0; count = seg_count();
    }
    // sequence: D(0) | P(0) Ud(1) D(1) Ur(1) | P(1) Ud(2) D(2) Ur(2) | ... | P(ncols-1)
    __device__ __forceinline__ void next_segment() {
        base += count;
        if (type == MT_D && k == 0) { type = MT_P; }
        else if (type == MT_P) { if (k + 1 < ncols) { type = MT_UD; k = k + 1; } else type = MT_END; }
        else if (type == MT_UD) type = MT_D;
        else if (type == MT_D) type = MT_UR;
        else if (type == MT_UR) type = MT_P;
        count = seg_count();
    }
    __device__ __forceinline__ void advance(long long t) {
        while (type != MT_END && t >= base + count) next_segment();
    }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// wait until *p >= need (a dependency that never arrives is a scheduling bug: fault instead of hanging the GPU)
__device__ __forceinline__ void mega_wait(const int* p, int need) {
    if (need <= 0) return;
    unsigned long long spins = 0;
    while (ld_acquire_gpu(p) < need) {
        __nanosleep(64);
        if (++spins > (1ull << 26)) __trap();
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MEGA_THREADS, 1)
chol_mega_kernel(CholArgs g, MegaArgs m) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * I8_STAGES + 3];
    __shared__ __align__(8) unsigned long long gbars[2][2 * GEMM_STAGES];
    __shared__ unsigned tmem_base_s;
    __shared__ double ecol[2][BN];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_ctarank();
    const int h = (int)rank;
    const int nrows = g.nblk;
    const int n = nrows * TB;
    const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    const unsigned sraw = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned sbase = (sraw + 1023u) & ~1023u;
    const unsigned cs = sbase + I8_STAGES * I8_STAGE_BYTES;         // C tile, rows of I8_CROW_BYTES
    double* fsm = reinterpret_cast<double*>(smem_raw + (sbase - sraw));     // the FP64 bodies' view of the same memory
    const unsigned full = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned empty = full + 8 * I8_STAGES, done = full + 16 * I8_STAGES, cfull = done + 8, tfree = done + 16;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, 2);
        }
        mbar_init(done, 1);
        mbar_init(cfull, 1);
        mbar_init(tfree, 4);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    const int grp = tid >> 8;                                       // thread group of the FP64 bodies
    GemmPipe pipe = gemm_pipe_init<true>(gbars[grp]);               // contains a group barrier
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem = tmem_base_s;

    // counters of the update pipeline: they run on across tiles and across runs of U tasks
    unsigned it_p = 0, it_m = 0, tile_m = 0, tile_e = 0;
    unsigned long long* prof = m.prof ? m.prof + (size_t)blockIdx.x * 16 : nullptr;
    const unsigned long long t_kernel0 = prof ? mega_now() : 0ull;

    MegaCursor cur;
    cur.init(g.nmat, nrows, m.ncols);
    long long t = cl;
    cur.advance(t);
    while (cur.type != MT_END) {
        if (cur.type == MT_UD || cur.type == MT_UR) {
            // ---------------- a run of consecutive update tasks of this cluster ----------------
            const unsigned long long tp0 = (prof && tid == 192) ? mega_now() : 0ull;
            cluster_sync_all();                                     // both CTAs' shared memory is free for the operand stages
            const unsigned long long tp1 = (prof && tid == 192) ? mega_now() : 0ull;
            const long long t0 = t;
            if (warp == 0) {
                // producer: waits for the digit planes of the two block rows, then streams them; runs ahead across tiles
                if (lane == 0) {
                    MegaCursor c2 = cur;
                    long long tt = t0;
                    while (c2.type == MT_UD || c2.type == MT_UR) {
                        const int j = (int)(tt - c2.base);
                        const int k = c2.k;
                        const int mat = (c2.type == MT_UD) ? j : j % c2.nmat;
                        const int ti = (c2.type == MT_UD) ? k : k + 1 + j / c2.nmat;
                        const unsigned long long tw0 = prof ? mega_now() : 0ull;
                        mega_wait(m.fP + (size_t)mat * nrows + ti, k);
                        mega_wait(m.fP + (size_t)mat * nrows + k, k);
                        if (prof) prof[10] += mega_now() - tw0;
                        asm volatile("fence.proxy.async;\n" ::: "memory");
                        const signed char* S = g.S8 + (size_t)mat * i8_slices_per_matrix(g.nblk, g.sgroup);
                        const int nks = k * (TB / I8_KB);
                        for (int ks = 0; ks < nks; ++ks, ++it_p) {
                            const unsigned s = it_p % I8_STAGES, u = it_p / I8_STAGES;
                            if (u > 0) mbar_wait(empty + 8 * s, (u - 1) & 1);
                            mbar_expect_tx(full + 8 * s, I8_STAGE_BYTES);
                            const unsigned dst = sbase + s * I8_STAGE_BYTES;
                            bulk_g2s_mc(dst + rank * (4 * I8_A_BYTES), S + i8_tile_off(ks, ti, g.nblk) + rank * (4 * I8_A_BYTES),
                                        4 * I8_A_BYTES, full + 8 * s, (unsigned short)3);
                            const signed char* bsrc = S + i8_tile_off(ks, k, g.nblk) + h * I8_B_BYTES;
#pragma unroll
                            for (int sl = 0; sl < I8_NSL; ++sl)
                                bulk_g2s(dst + I8_NSL * I8_A_BYTES + sl * I8_B_BYTES, bsrc + sl * I8_A_BYTES, I8_B_BYTES, full + 8 * s);
                        }
                        tt += ncl;
                        c2.advance(tt);
                    }
                }
            } else if (warp == 1) {
                if (lane == 0) {
                    const unsigned idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TB >> 4) << 24);
                    MegaCursor c2 = cur;
                    long long tt = t0;
                    while (c2.type == MT_UD || c2.type == MT_UR) {
                        const int nks = c2.k * (TB / I8_KB);
                        if (tile_m > 0) {                           // the epilogue has drained the previous tile's accumulators
                            mbar_wait(tfree, (tile_m - 1) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        }
                        for (int ks = 0; ks < nks; ++ks, ++it_m) {
                            const unsigned s = it_m % I8_STAGES;
                            mbar_wait(full + 8 * s, (it_m / I8_STAGES) & 1);
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
                        ++tile_m;
                        tt += ncl;
                        c2.advance(tt);
                    }
                }
            } else if (warp < 6) {
                // epilogue: thread = one row of the tile (TMEM lane)
                const int q = warp & 3;
                const int row = q * 32 + lane;
                const int et = tid - 64;
                const unsigned crow = cs + row * I8_CROW_BYTES;
                double* cgen = reinterpret_cast<double*>(smem_raw + (crow - sraw));
                const unsigned tlane = tmem + ((unsigned)(q * 32) << 16);
                MegaCursor c2 = cur;
                long long tt = t0;
                while (c2.type == MT_UD || c2.type == MT_UR) {
                    const int j = (int)(tt - c2.base);
                    const int k = c2.k;
                    const bool diag = c2.type == MT_UD;
                    const int mat = diag ? j : j % c2.nmat;
                    const int ti = diag ? k : k + 1 + j / c2.nmat;
                    const bool half_tile = diag && h == 1;          // rows 0..63 lie strictly above the diagonal
                    const bool skip = half_tile && q < 2;           // warp-uniform
                    const double* esc = g.rscale + (size_t)mat * 2 * n + n;
                    double* Crow = g.A + (size_t)mat * g.mat_stride + (size_t)(ti * TB + row) * g.ld + (size_t)k * TB + h * BN;
                    const int eb = tile_e & 1;
                    if (et < BN) ecol[eb][et] = 0.5 * esc[k * TB + h * BN + et];
                    if (et == 0) mbar_expect_tx(cfull, (half_tile ? TB / 2 : TB) * BN * 8);
                    if (!skip) bulk_g2s(crow, Crow, BN * 8, cfull);
                    const double ei = esc[ti * TB + row];
                    named_bar_sync(1, 128);
                    mbar_wait_long(done, tile_e & 1);               // may include the wait for other clusters' panels
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    mbar_wait(cfull, tile_e & 1);
                    if (!skip) {
#pragma unroll 1
                        for (int c8 = 0; c8 < BN / 8; ++c8) {
                            int v[I8_NSL][8];
#pragma unroll
                            for (int d = 0; d < I8_NSL; ++d) tmem_ld8(tlane + d * BN + c8 * 8, v[d]);
                            tmem_ld_wait();
                            if (c8 == BN / 8 - 1) {
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
                                    const int jj = j2 + u;
                                    const double H = i8_fold4(v[0][jj], v[1][jj], v[2][jj], v[3][jj]);
                                    const double Lo = i8_fold4(v[4][jj], v[5][jj], v[6][jj], v[7][jj]);
                                    const double val = fma(H, 268435456.0, Lo);
                                    r2[u] = (val * ei) * ecol[eb][c8 * 8 + jj];
                                }
                                cc.x -= r2[0];
                                cc.y -= r2[1];
                                *reinterpret_cast<double2*>(cgen + c8 * 8 + j2) = cc;
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        bulk_s2g(Crow, crow, BN * 8);
                        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");      // written, not only read: the tile is published below
                        asm volatile("fence.proxy.async;\n" ::: "memory");
                    } else {
                        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tfree);
                    }
                    // publish this half tile: every epilogue thread's row is in global memory
                    __threadfence();
                    named_bar_sync(2, 128);
                    if (et == 0) red_release_gpu_add(m.fU + (size_t)mat * nrows + ti, 1);
                    ++tile_e;
                    tt += ncl;
                    c2.advance(tt);
                }
            }
            // every role walks the same run: find its end
            int nrun = 0;
            while (cur.type == MT_UD || cur.type == MT_UR) {
                t += ncl;
                cur.advance(t);
                ++nrun;
            }
            // roles that did the walking above hold the advanced counters only in their own registers: tile / it counters
            // are per-role, so nothing has to be exchanged -- but all roles must have left the run before shared memory
            // changes hands
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (prof && tid == 192) {
                prof[0] += mega_now() - tp1;
                prof[1] += nrun;
                prof[9] += tp1 - tp0;
            }
        } else if (cur.type == MT_D) {
            // diagonal blocks: one matrix per CTA (thread group 0; the block needs most of the shared memory)
            const int k = cur.k;
            const int mat = 2 * (int)(t - cur.base) + h;
            if (mat < g.nmat && grp == 0) {
                const unsigned long long td0 = (prof && tid == 0) ? mega_now() : 0ull;
                if ((tid & 255) == 0) mega_wait(m.fU + (size_t)mat * nrows + k, 2 * k);
                const unsigned long long td1 = (prof && tid == 0) ? mega_now() : 0ull;
                grp_sync<true>();
                potrf_diag_body<true>(g, k, mat, fsm);
                __threadfence();
                grp_sync<true>();
                if ((tid & 255) == 0) st_release_gpu(m.fD + mat, k + 1);
                if (prof && tid == 0) {
                    prof[2] += mega_now() - td1;
                    prof[3] += td1 - td0;
                    prof[4] += 1;
                }
            }
            __syncthreads();
            t += ncl;
            cur.advance(t);
        } else {    // MT_P
            // panel solves: four tiles per cluster task, one per thread group of each CTA
            const int k = cur.k;
            const int idx = 4 * (int)(t - cur.base) + 2 * h + grp;
            if (idx < g.nmat * (nrows - k - 1)) {
                const int ti = k + 1 + idx / g.nmat, mat = idx % g.nmat;
                const unsigned long long tq0 = (prof && tid == 0) ? mega_now() : 0ull;
                if ((tid & 255) == 0) {
                    mega_wait(m.fD + mat, k + 1);
                    mega_wait(m.fU + (size_t)mat * nrows + ti, 2 * k);
                }
                const unsigned long long tq1 = (prof && tid == 0) ? mega_now() : 0ull;
                grp_sync<true>();
                trsm_body<true>(g, k, 0, m.tflags, ti, mat, fsm + (size_t)grp * (GEMM_SMEM_BYTES / 8), pipe);
                __threadfence();
                grp_sync<true>();
                if ((tid & 255) == 0) st_release_gpu(m.fP + (size_t)mat * nrows + ti, k + 1);
                if (prof && tid == 0) {
                    prof[5] += mega_now() - tq1;
                    prof[6] += tq1 - tq0;
                    prof[7] += 1;
                }
            }
            __syncthreads();
            t += ncl;
            cur.advance(t);
        }
    }
    if (prof && tid == 0) prof[8] += mega_now() - t_kernel0;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();        // nobody leaves while the peer may still multicast into / commit onto this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}
