// Probe for an integer-sliced (Ozaki-style) FP64 trailing update on the tcgen05 INT8 tensor path (sm_100a).
//
// Question it answers before any product code is written: with 8 signed 7-bit slices per operand a
// double-precision rank-k update needs the 36 slice pairs (s, t), s + t <= 7, accumulated exactly in
// int32 -- one TMEM accumulator per diagonal d = s + t.  8 accumulators x N columns must fit the 512
// TMEM columns, so N = 64 (M = 128).  What does `tcgen05.mma.kind::i8` deliver in that shape, operands
// in shared memory (SS), fed by 1-D bulk copies from L2?  And are the descriptors right (checked
// against a CPU integer product)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/i8_probe tools/i8_probe.cu
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

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", p.name, sms, p.clockRate / 1000);
    run<64, 8, 2>("8 diagonals x 64 columns (the product shape)", 4, 2048, sms);
    run<128, 4, 1>("4 x 128 columns (two-pass shape; one stage, streamed figure not pipelined)", 4, 2048, sms);
    run<256, 2, 1>("2 x 256 columns (widest MMA, reference rate; one stage)", 4, 1024, sms);
    return 0;
}
