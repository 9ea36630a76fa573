// Probe for an integer-sliced (Ozaki-style) FP64 trailing update on the tcgen05 INT8 tensor path (sm_100a).
//
// Question it answers before any product code is written: with 8 signed 7-bit slices per operand a
// double-precision rank-k update needs the 36 slice pairs (s, t), s + t <= 7, accumulated exactly in
// int32 -- one TMEM accumulator per diagonal d = s + t.  8 accumulators x N columns must fit the 512
// TMEM columns, so N = 64 (M = 128).  What does `tcgen05.mma.kind::i8` deliver in that shape, operands
// in shared memory (SS), fed by 1-D bulk copies from L2?  And are the descriptors right (checked
// against a CPU integer product)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -o tools/i8_probe tools/i8_probe.cu
//   ./tools/i8_probe            -> one JSON line per experiment
//
// Slice storage (what the product's slicing kernel would write), per matrix, int8:
//   plane (kc, s) = 64-byte K chunk kc of slice s, all rows:   base + ((kc * 8 + s) * rows) * 64
//   inside a plane: row r, byte kb -> (r / 8) * 512 + (kb / 16) * 128 + (r % 8) * 16 + (kb % 16)
// i.e. the UMMA canonical K-major no-swizzle layout (8 x 16-byte core matrices, LBO = 128 B between the
// 16-byte K chunks, SBO = 512 B between 8-row groups), so any 8-aligned row range of a plane is one
// contiguous block that a single cp.async.bulk drops into shared memory ready for the tensor core.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <array>
#include <initializer_list>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static constexpr int NSL = 8;          // slices
static constexpr int KC = 64;          // bytes of K per pipeline stage (2 MMAs of K = 32)
static constexpr int TM = 128;         // tile rows (TMEM lanes)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned a, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(a), "r"(bytes) : "memory");
}
// bounded wait: a wrong descriptor must not hang the box
__device__ __forceinline__ bool mbar_wait(unsigned a, unsigned parity, int* err) {
    for (long long spin = 0; spin < (1ll << 26); ++spin) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    if (err) atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
    // K-major, SWIZZLE_NONE, LBO = 128 B, SBO = 512 B, descriptor version 1 (sm_100)
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (32ull << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// One CTA = one 128 x N output tile, NACC accumulators (diagonal d -> accumulator d % NACC).
// warp 0: bulk-copy producer (lane 0), warp 1: MMA issuer (lane 0), warps 2..5: epilogue (TMEM lanes 32*(warp&3)).
// feed = 1: stages streamed from global (the real main loop);  feed = 0: operands loaded once, MMAs re-issued on
// the same stage `nchunks` times (pure tensor-pipe rate).
template <int N, int NACC, int NST>
__global__ void __launch_bounds__(192, 1)
i8_tile_kernel(const int8_t* __restrict__ slices, int rows, int nch_data, int rowA, int rowB, int nchunks, int feed,
               int* __restrict__ out, long long* __restrict__ cycles, int* __restrict__ err) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr unsigned A_BYTES = TM * KC;              // per slice
    constexpr unsigned B_BYTES = N * KC;
    constexpr unsigned STAGE = NSL * (A_BYTES + B_BYTES);
    __shared__ __align__(8) unsigned long long bars[2 * NST + 1];
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned sbase = smem_u32(smem);
    const unsigned full = smem_u32(bars), empty = full + 8 * NST, done = full + 16 * NST;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem = tmem_base_s;

    // every CTA takes its own row tiles of the 2048-row panel (K = 1024, 16 MB of slices: L2 resident, like one
    // matrix's group panel); the K chunks wrap around so the stream can be made as long as the timing needs
    const int8_t* base = slices;
    rowA += (blockIdx.x % 16) * TM;
    rowB = (rowB + (blockIdx.x * 7 % 32) * 64) % rows;
    const int nstream = feed ? nchunks : 1;
    long long t0 = 0;
    if (warp == 0) {
        if (lane == 0) {
            for (int kc = 0; kc < nstream; ++kc) {
                const int s = kc % NST, u = kc / NST;
                if (u > 0 && !mbar_wait(empty + 8 * s, (u - 1) & 1, err)) break;
                mbar_expect_tx(full + 8 * s, STAGE);
                const unsigned dst = sbase + s * STAGE;
                for (int sl = 0; sl < NSL; ++sl) {
                    const int8_t* plane = base + ((size_t)((kc % nch_data) * NSL + sl) * rows) * KC;
                    bulk_g2s(dst + sl * A_BYTES, plane + (size_t)rowA * KC, A_BYTES, full + 8 * s);
                    bulk_g2s(dst + NSL * A_BYTES + sl * B_BYTES, plane + (size_t)rowB * KC, B_BYTES, full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(TM >> 4) << 24);
            unsigned used = 0;          // bit d set once accumulator d holds data
            t0 = clock64();
            for (int kc = 0; kc < nchunks; ++kc) {
                const int s = feed ? kc % NST : 0;
                if (feed || kc == 0) {
                    if (!mbar_wait(full + 8 * s, feed ? (kc / NST) & 1 : 0, err)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                const unsigned a0 = sbase + s * STAGE, b0 = a0 + NSL * A_BYTES;
#pragma unroll
                for (int ks = 0; ks < KC / 32; ++ks) {
#pragma unroll
                    for (int sa = 0; sa < NSL; ++sa) {
                        const unsigned long long ad = umma_desc(a0 + sa * A_BYTES + ks * 256);
#pragma unroll
                        for (int sb = 0; sb < NSL; ++sb) {
                            if (sa + sb >= NSL) continue;
                            const int d = (sa + sb) % NACC;
                            mma_i8(tmem + d * N, ad, umma_desc(b0 + sb * B_BYTES + ks * 256), idesc, (used >> d) & 1u);
                            used |= 1u << d;
                        }
                    }
                }
                if (feed) umma_commit(empty + 8 * s);
            }
            umma_commit(done);
        }
    }
    // everybody: wait for the accumulators
    const bool okd = mbar_wait(done, 0, err);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp == 1 && lane == 0 && cycles) cycles[blockIdx.x] = clock64() - t0;
    if (warp >= 2 && okd && out && blockIdx.x == 0) {
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        __syncwarp();
        for (int d = 0; d < NACC; ++d)
            for (int c = 0; c < N; c += 16) {
                int v[16];
                tmem_ld16(tmem + ((unsigned)(q * 32) << 16) + d * N + c, v);
                for (int j = 0; j < 16; ++j) out[((size_t)d * TM + row) * N + c + j] = v[j];
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}


// ------------------------------------------------------------------------------------------------------------
// v2: the product main loop.  Same 8 x 64-column accumulators, but (a) consecutive B slices are consecutive rows
// of the canonical layout, so ONE MMA with N = 64 n multiplies A_s with slices t0 .. t0+n-1 and lands in the n
// adjacent accumulators of diagonals s+t0 .. s+t0+n-1: 12 wide MMAs per 32-byte K step instead of 36 narrow ones
// (shared-memory operand reads drop from 192 to 104 B/clk); (b) the CN CTAs of a cluster own CN adjacent 64-column
// tiles of the same 128 rows and fetch the A stage once, each CTA multicasting 8 / CN slices to all of them
// (L2 -> SM bytes per CTA and stage: 64 / CN + 32 KB instead of 96 KB).
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(unsigned dst, const void* src, unsigned bytes, unsigned mbar, unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(unsigned mbar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(mbar), "h"(mask) : "memory");
}

// MMA schedule of one 32-byte K step: entries (sa, t0, n) = A slice sa times B slices t0 .. t0+n-1 (N = 64 n) into the
// accumulators of diagonals sa+t0 ..; `first` marks the entries that touch their columns first (they overwrite on the very first K step)
struct Sched { int n; unsigned char sa[40], t0[40], nn[40], first[40]; };
// the same schedules as compile-time tables (one thread issues every MMA: the issue loop has to be straight-line code)
template <int ID> struct SchedC;
template <> struct SchedC<0> { static constexpr int n = 12; };   // wide12
template <> struct SchedC<1> { static constexpr int n = 20; };   // disjoint20
template <> struct SchedC<2> { static constexpr int n = 16; };   // disjoint16
template <> struct SchedC<3> { static constexpr int n = 36; };   // narrow36 (diagonal round-robin)
template <int ID> __device__ __forceinline__ constexpr int sched_entry(int e, int f) {
    // f: 0 = sa, 1 = t0, 2 = n, 3 = first
    if (ID == 0) { constexpr int T[12][4] = {{0,0,4,1},{0,4,4,1},{1,0,4,0},{1,4,3,0},{2,0,4,0},{2,4,2,0},{3,0,4,0},{3,4,1,0},{4,0,4,0},{5,0,3,0},{6,0,2,0},{7,0,1,0}}; return T[e][f]; }
    if (ID == 1) { constexpr int T[20][4] = {{0,0,2,1},{0,6,2,1},{0,4,2,1},{2,4,2,0},{2,2,2,0},{4,2,2,0},{4,0,2,0},{6,0,2,0},{0,2,2,1},{1,6,1,0},{2,0,2,0},{1,4,2,0},{3,4,1,0},{1,2,2,0},
                                             {3,2,2,0},{5,2,1,0},{1,0,2,0},{3,0,2,0},{5,0,2,0},{7,0,1,0}}; return T[e][f]; }
    if (ID == 2) { constexpr int T[16][4] = {{0,0,4,1},{0,4,4,1},{1,0,3,0},{1,3,4,0},{2,0,2,0},{2,4,2,0},{2,2,2,0},{3,3,2,0},{3,1,2,0},{5,1,2,0},{5,0,1,0},{6,0,2,0},{3,0,1,0},{4,3,1,0},{4,0,3,0},{7,0,1,0}}; return T[e][f]; }
    // ID == 3: e < 8: (0, e, 1, first); then s = 1..7, d = s..7
    if (e < 8) return f == 0 ? 0 : f == 1 ? e : f == 2 ? 1 : 1;
    int idx = 8;
    for (int sa = 1; sa < 8; ++sa)
        for (int d = sa; d < 8; ++d) { if (idx == e) return f == 0 ? sa : f == 1 ? d - sa : f == 2 ? 1 : 0; ++idx; }
    return 0;
}

template <int CN, int SID>
__global__ void __launch_bounds__(192, 1)
i8_wide_kernel(const int8_t* __restrict__ slices, int rows, int nch_data, int nchunks, int feed,
               int* __restrict__ out, long long* __restrict__ cycles, int* __restrict__ err) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int N = 64;
    constexpr unsigned A_BYTES = TM * KC, B_BYTES = N * KC;
    constexpr unsigned STAGE = NSL * (A_BYTES + B_BYTES);
    constexpr int NST = 2;
    __shared__ __align__(8) unsigned long long bars[2 * NST + 1];
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned rank = CN > 1 ? cluster_rank() : 0u;
    const int cl = blockIdx.x / CN;
    const unsigned short mask = (unsigned short)((1u << CN) - 1u);
    const unsigned sbase = smem_u32(smem);
    const unsigned full = smem_u32(bars), empty = full + 8 * NST, done = full + 16 * NST;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, CN); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (CN > 1) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem = tmem_base_s;

    const int rowA = (cl % 16) * TM;
    const int rowB = ((cl * 5 % 8) * 256 + (int)rank * 64) % rows;
    long long t0 = 0;
    if (warp == 0) {
        if (lane == 0) {
            for (int kc = 0; kc < (feed ? nchunks : 1); ++kc) {
                const int s = kc % NST, u = kc / NST;
                if (u > 0 && !mbar_wait(empty + 8 * s, (u - 1) & 1, err)) break;
                mbar_expect_tx(full + 8 * s, STAGE);
                const unsigned dst = sbase + s * STAGE;
                const int8_t* chunk = slices + ((size_t)((kc % nch_data) * NSL) * rows) * KC;
                for (int sl = (int)rank; sl < NSL; sl += CN) {
                    const int8_t* src = chunk + ((size_t)sl * rows + rowA) * KC;
                    if (CN > 1) bulk_g2s_mc(dst + sl * A_BYTES, src, A_BYTES, full + 8 * s, mask);
                    else bulk_g2s(dst + sl * A_BYTES, src, A_BYTES, full + 8 * s);
                }
                for (int sl = 0; sl < NSL; ++sl)
                    bulk_g2s(dst + NSL * A_BYTES + sl * B_BYTES, chunk + ((size_t)sl * rows + rowB) * KC, B_BYTES, full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TM >> 4) << 24);
            t0 = clock64();
            for (int kc = 0; kc < nchunks; ++kc) {
                const int s = feed ? kc % NST : 0;
                if (feed || kc == 0) {
                    if (!mbar_wait(full + 8 * s, feed ? (kc / NST) & 1 : 0, err)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                const unsigned a0 = sbase + s * STAGE, b0 = a0 + NSL * A_BYTES;
#pragma unroll
                for (int ks = 0; ks < KC / 32; ++ks) {
#pragma unroll
                    for (int e = 0; e < SchedC<SID>::n; ++e) {
                        constexpr int dummy = 0; (void)dummy;
                        const int sa = sched_entry<SID>(e, 0), tb = sched_entry<SID>(e, 1), n = sched_entry<SID>(e, 2);
                        const unsigned idesc = idesc0 | ((unsigned)((64 * n) >> 3) << 17);
                        const unsigned accum = (kc > 0 || ks > 0 || !sched_entry<SID>(e, 3)) ? 1u : 0u;
                        mma_i8(tmem + (sa + tb) * N, umma_desc(a0 + sa * A_BYTES + ks * 256), umma_desc(b0 + tb * B_BYTES + ks * 256), idesc, accum);
                    }
                }
                if (feed) { if (CN > 1) umma_commit_mc(empty + 8 * s, mask); else umma_commit(empty + 8 * s); }
            }
            umma_commit(done);
        }
    }
    const bool okd = mbar_wait(done, 0, err);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp == 1 && lane == 0 && cycles) cycles[blockIdx.x] = clock64() - t0;
    if (warp >= 2 && okd && out && cl == 0) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        __syncwarp();
        for (int d = 0; d < NSL; ++d)
            for (int c = 0; c < N; c += 16) {
                int v[16];
                tmem_ld16(tmem + ((unsigned)(q * 32) << 16) + d * N + c, v);
                for (int j = 0; j < 16; ++j) out[(((size_t)rank * NSL + d) * TM + row) * N + c + j] = v[j];
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (CN > 1) cluster_sync_all();      // nobody leaves while a peer may still multicast into / arrive on this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

static inline size_t plane_off(int kc, int s, int rows) { return ((size_t)(kc * NSL + s) * rows) * KC; }
static inline size_t in_plane(int r, int kb) { return (size_t)(r / 8) * 512 + (kb / 16) * 128 + (r % 8) * 16 + (kb % 16); }

template <int N, int NACC, int NST>
static void run(const char* name, int nchunks_check, int nchunks_perf, int sms) {
    const int rows = 2048, nch_data = 16;
    const size_t bytes = (size_t)nch_data * NSL * rows * KC;
    std::vector<int8_t> h(bytes);
    uint32_t st = 12345u;
    for (size_t i = 0; i < bytes; ++i) { st = st * 1664525u + 1013904223u; h[i] = (int8_t)((int)((st >> 24) & 127) - 64); }
    int8_t* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
    int* dout; CK(cudaMalloc(&dout, sizeof(int) * NACC * TM * N));
    long long* dcyc; CK(cudaMalloc(&dcyc, sizeof(long long) * sms));
    int* derr; CK(cudaMalloc(&derr, sizeof(int))); CK(cudaMemset(derr, 0, sizeof(int)));
    const size_t smem = NST * (size_t)NSL * (TM + N) * KC;
    CK(cudaFuncSetAttribute(i8_tile_kernel<N, NACC, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rowA = 0, rowB = 192;

    // ---- correctness (streamed stages) ----
    i8_tile_kernel<N, NACC, NST><<<1, 192, smem>>>(d, rows, nch_data, rowA, rowB, nchunks_check, 1, dout, nullptr, derr);
    CK(cudaDeviceSynchronize());
    int herr = 0; CK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> got((size_t)NACC * TM * N); CK(cudaMemcpy(got.data(), dout, got.size() * sizeof(int), cudaMemcpyDeviceToHost));
    long long bad = 0;
    if (!herr) {
        std::vector<long long> ref((size_t)NACC * TM * N, 0);
        for (int kc = 0; kc < nchunks_check; ++kc)
            for (int sa = 0; sa < NSL; ++sa)
                for (int sb = 0; sa + sb < NSL; ++sb) {
                    const int dd = (sa + sb) % NACC;
                    const int8_t* pa = h.data() + plane_off(kc, sa, rows);
                    const int8_t* pb = h.data() + plane_off(kc, sb, rows);
                    for (int i = 0; i < TM; ++i)
                        for (int j = 0; j < N; ++j) {
                            long long acc = 0;
                            for (int kb = 0; kb < KC; ++kb) acc += (long long)pa[in_plane(rowA + i, kb)] * pb[in_plane(rowB + j, kb)];
                            ref[((size_t)dd * TM + i) * N + j] += acc;
                        }
                }
        for (size_t i = 0; i < ref.size(); ++i) bad += (long long)got[i] != ref[i];
    }
    // ---- tensor-pipe rate: all SMs, operands resident (feed = 0) and streamed (feed = 1) ----
    float ms_res = -1, ms_feed = -1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int feed = 0; feed < 2 && !herr; ++feed) {
        for (int rep = 0; rep < 3; ++rep) {      // last repetition is the one reported
            CK(cudaEventRecord(e0));
            i8_tile_kernel<N, NACC, NST><<<sms, 192, smem>>>(d, rows, nch_data, rowA, rowB, nchunks_perf, feed, nullptr, dcyc, derr);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(feed ? &ms_feed : &ms_res, e0, e1));
        }
        CK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    }
    std::vector<long long> cyc(sms); CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    const double macs_per_cta = (double)nchunks_perf * 36.0 * (KC / 32) * TM * N * 32.0;
    const double pops_res = ms_res > 0 ? 2.0 * macs_per_cta * sms / (ms_res * 1e-3) / 1e15 : 0;
    const double pops_feed = ms_feed > 0 ? 2.0 * macs_per_cta * sms / (ms_feed * 1e-3) / 1e15 : 0;
    // FP64-equivalent rate: 36 int8 MACs stand for one FP64 FMA
    printf("{\"probe\": \"%s\", \"N\": %d, \"accumulators\": %d, \"wait_timeout\": %d, \"mismatches\": %lld, \"checked\": %d, "
           "\"ms_resident\": %.4f, \"int8_pops_resident\": %.3f, \"ms_streamed\": %.4f, \"int8_pops_streamed\": %.3f, "
           "\"fp64_equiv_tflops_streamed\": %.2f, \"mac_per_clk_per_sm_streamed(last CTA clock)\": %.1f, \"sms\": %d, \"nchunks\": %d}\n",
           name, N, NACC, herr, bad, NACC * TM * N, ms_res, pops_res, ms_feed, pops_feed, pops_feed * 1e3 / 36.0,
           cyc[0] > 0 ? macs_per_cta / (double)cyc[0] : 0.0, sms, nchunks_perf);
    fflush(stdout);
    cudaFree(d); cudaFree(dout); cudaFree(dcyc); cudaFree(derr);
}

template <int CN, int SID>
static void run_wide(const char* sname, const Sched& sch, int nchunks_check, int nchunks_perf, int sms) {
    constexpr int N = 64;
    const int rows = 2048, nch_data = 16;
    const size_t bytes = (size_t)nch_data * NSL * rows * KC;
    std::vector<int8_t> h(bytes);
    uint32_t st = 777u + CN;
    for (size_t i = 0; i < bytes; ++i) { st = st * 1664525u + 1013904223u; h[i] = (int8_t)((int)((st >> 24) & 127) - 64); }
    int8_t* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
    int* dout; CK(cudaMalloc(&dout, sizeof(int) * CN * NSL * TM * N));
    const int grid = sms / CN * CN;
    long long* dcyc; CK(cudaMalloc(&dcyc, sizeof(long long) * grid));
    int* derr; CK(cudaMalloc(&derr, sizeof(int))); CK(cudaMemset(derr, 0, sizeof(int)));
    const size_t smem = 2 * (size_t)NSL * (TM + N) * KC;
    CK(cudaFuncSetAttribute(i8_wide_kernel<CN, SID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto launch = [&](int g, int nch, int feed, int* o, long long* cy) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(g); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CN; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, i8_wide_kernel<CN, SID>, (const int8_t*)d, rows, nch_data, nch, feed, o, cy, derr));
    };
    launch(CN, nchunks_check, 1, dout, nullptr);
    CK(cudaDeviceSynchronize());
    int herr = 0; CK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> got((size_t)CN * NSL * TM * N); CK(cudaMemcpy(got.data(), dout, got.size() * sizeof(int), cudaMemcpyDeviceToHost));
    long long bad = 0;
    if (!herr) {
        for (int r = 0; r < CN; ++r) {
            const int rowA = 0, rowB = r * 64;
            std::vector<long long> ref((size_t)NSL * TM * N, 0);
            for (int kc = 0; kc < nchunks_check; ++kc)
                for (int sa = 0; sa < NSL; ++sa)
                    for (int sb = 0; sa + sb < NSL; ++sb) {
                        const int8_t* pa = h.data() + plane_off(kc, sa, rows);
                        const int8_t* pb = h.data() + plane_off(kc, sb, rows);
                        for (int i = 0; i < TM; ++i)
                            for (int j = 0; j < N; ++j) {
                                long long acc = 0;
                                for (int kb = 0; kb < KC; ++kb) acc += (long long)pa[in_plane(rowA + i, kb)] * pb[in_plane(rowB + j, kb)];
                                ref[((size_t)(sa + sb) * TM + i) * N + j] += acc;
                            }
                    }
            for (size_t i = 0; i < ref.size(); ++i) bad += (long long)got[(size_t)r * NSL * TM * N + i] != ref[i];
        }
    }
    float ms = -1, ms_res = -1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int feed = 0; feed < 2; ++feed)
        for (int rep = 0; rep < 3 && !herr; ++rep) {
            CK(cudaEventRecord(e0));
            launch(grid, nchunks_perf, feed, nullptr, dcyc);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(feed ? &ms : &ms_res, e0, e1));
        }
    (void)sch;
    CK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<long long> cyc(grid); CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    const double macs_per_cta = (double)nchunks_perf * 36.0 * (KC / 32) * TM * N * 32.0;
    const double pops = ms > 0 ? 2.0 * macs_per_cta * grid / (ms * 1e-3) / 1e15 : 0;
    const double l2_tbs = ms > 0 ? (double)nchunks_perf * grid * (64.0 / CN + 32.0) * 1024.0 / (ms * 1e-3) / 1e12 : 0;
    printf("{\"probe\": \"schedule %s (%d MMAs per K step), cluster of %d multicasting A\", \"wait_timeout\": %d, \"mismatches\": %lld, \"checked\": %d, "
           "\"ms_resident\": %.4f, \"ms\": %.4f, \"int8_pops\": %.3f, \"fp64_equiv_tflops\": %.2f, \"l2_to_sm_tb_per_s\": %.2f, \"mac_per_clk_cta0\": %.1f, \"ctas\": %d, \"nchunks\": %d}\n",
           sname, sch.n, CN, herr, bad, CN * NSL * TM * N, ms_res, ms, pops, pops * 1e3 / 36.0, l2_tbs, cyc[0] > 0 ? macs_per_cta / (double)cyc[0] : 0.0, grid, nchunks_perf);
    fflush(stdout);
    cudaFree(d); cudaFree(dout); cudaFree(dcyc); cudaFree(derr);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", p.name, sms, p.clockRate / 1000);
    run<64, 8, 2>("8 diagonals x 64 columns (the product shape)", 4, 2048, sms);
    run<128, 4, 1>("4 x 128 columns (two-pass shape; one stage, streamed figure not pipelined)", 4, 2048, sms);
    run<256, 2, 1>("2 x 256 columns (widest MMA, reference rate; one stage)", 4, 1024, sms);
    auto finish = [](Sched& s) {
        bool touched[8] = {};
        for (int e = 0; e < s.n; ++e) {
            int nt = 0;
            for (int b = s.sa[e] + s.t0[e]; b < s.sa[e] + s.t0[e] + s.nn[e]; ++b) nt += touched[b];
            if (nt != 0 && nt != s.nn[e]) { printf("{\"error\": \"schedule entry %d mixes fresh and used accumulators\"}\n", e); exit(1); }
            s.first[e] = nt == 0;
            for (int b = s.sa[e] + s.t0[e]; b < s.sa[e] + s.t0[e] + s.nn[e]; ++b) touched[b] = true;
        }
    };
    auto mk = [&](std::initializer_list<std::array<int, 3>> l) {
        Sched s{}; s.n = 0;
        for (auto& e : l) { s.sa[s.n] = (unsigned char)e[0]; s.t0[s.n] = (unsigned char)e[1]; s.nn[s.n] = (unsigned char)e[2]; ++s.n; }
        finish(s);
        return s;
    };
    // widest: 12 MMAs, consecutive MMAs overlap in TMEM columns
    Sched wide12 = mk({{0,0,4},{0,4,4},{1,0,4},{1,4,3},{2,0,4},{2,4,2},{3,0,4},{3,4,1},{4,0,4},{5,0,3},{6,0,2},{7,0,1}});
    // 20 MMAs of <= 128 columns, consecutive MMAs touch disjoint accumulator columns (cyclically)
    Sched dis20 = mk({{0,0,2},{0,6,2},{0,4,2},{2,4,2},{2,2,2},{4,2,2},{4,0,2},{6,0,2},{0,2,2},{1,6,1},{2,0,2},{1,4,2},{3,4,1},{1,2,2},
                      {3,2,2},{5,2,1},{1,0,2},{3,0,2},{5,0,2},{7,0,1}});
    // 16 MMAs mixed widths, consecutive disjoint
    Sched dis16 = mk({{0,0,4},{0,4,4},{1,0,3},{1,3,4},{2,0,2},{2,4,2},{2,2,2},{3,3,2},{3,1,2},{5,1,2},{5,0,1},{6,0,2},{3,0,1},{4,3,1},{4,0,3},{7,0,1}});
    // 36 single-diagonal MMAs, diagonal-major order d = 0..7 cycling s (accumulator reuse distance >= 1 .. 8)
    Sched nar36{};
    {   // first the 8 MMAs (s = 0, t = d) that initialise every diagonal, then the rest round-robin over diagonals
        for (int d = 0; d < 8; ++d) { nar36.sa[nar36.n] = 0; nar36.t0[nar36.n] = (unsigned char)d; nar36.nn[nar36.n] = 1; ++nar36.n; }
        for (int sa = 1; sa < 8; ++sa)
            for (int d = sa; d < 8; ++d) { nar36.sa[nar36.n] = (unsigned char)sa; nar36.t0[nar36.n] = (unsigned char)(d - sa); nar36.nn[nar36.n] = 1; ++nar36.n; }
        finish(nar36);
    }
    run_wide<1, 0>("wide12", wide12, 4, 2048, sms);
    run_wide<1, 1>("disjoint20", dis20, 4, 2048, sms);
    run_wide<1, 2>("disjoint16", dis16, 4, 2048, sms);
    run_wide<1, 3>("narrow36", nar36, 4, 2048, sms);
    run_wide<2, 1>("disjoint20", dis20, 4, 2048, sms);
    run_wide<2, 2>("disjoint16", dis16, 4, 2048, sms);
    return 0;
}
