// ubench_umma.cu — what bounds tcgen05.mma kind::i8 in SS mode?  ("an SS-mode M = N = 128, K = 32
// instruction reads 4 KB of A + 4 KB of B per 64 clocks = the SM's whole shared-memory read bandwidth".)
// One CTA per SM, one warp issues back-to-back MMAs over a ring of shared-memory operand slabs (contents irrelevant),
// accumulators in TMEM, no epilogue, optional concurrent shared-memory writes from the other warps at the TMA ring's
// rate.  Prints clocks per MMA instruction and the int8 rate for each variant as one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr uint32_t kDescHi = 0x40004040u;
__device__ __forceinline__ uint32_t desc_lo(uint32_t a) { return ((a >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
    return pred;
}
template <bool ACC>
__device__ __forceinline__ void umma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t leader)
{
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 e, %4, 0;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 da, {%1, %6};\n\tmov.b64 db, {%2, %6};\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(leader), "n"(ACC ? 1 : 0), "r"(kDescHi) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t leader)
{
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\tsetp.ne.b32 e, %4, 0;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 db, {%2, %6};\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(leader), "n"(ACC ? 1 : 0), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void commit_if(uint64_t *bar, uint32_t leader)
{
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}

// MODE 0: SS, N = 128   1: SS, N = 256   2: A from TMEM, N = 128   3: A from TMEM, N = 256
template <int MODE, int WRITERS>
__global__ void __launch_bounds__(WRITERS >= 32 ? 96 : 64 + WRITERS * 32 + (WRITERS == 0 ? 32 : 0), 1) k(int tiles, long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int N = (MODE & 1) ? 256 : 128;
    constexpr int STAGE = 2 * N * 128;                 // a train tile: 2 K-slabs x N rows x 128 B
    constexpr int NST = (MODE & 1) ? 2 : 4;
    uint8_t *sA = smem;                                // 2 row blocks x 2 slabs x 128 rows x 128 B = 64 KB
    uint8_t *sB = smem + 65536;                        // NST stages
    uint8_t *sW = sB + NST * STAGE;                    // 16 KB scratch the writer warps stream into
    __shared__ uint64_t bar, bar2[3], bar3;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2[0], 1); mbar_init(&bar2[1], 1); mbar_init(&bar2[2], 1); mbar_init(&bar3, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        const uint32_t leader = elect_one();
        const uint32_t a0 = desc_lo(smem_u32(sA)), b0 = desc_lo(smem_u32(sB));
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const uint32_t b_lo = b0 + (t % NST) * (STAGE >> 4);
            if (WRITERS == 40) {   // the two row blocks' instruction streams interleaved (what two issuing warps produce)
                const uint32_t d0 = tmem + (uint32_t)(((t & 1) * 2 + 0) * N) % 512u, d1 = tmem + (uint32_t)(((t & 1) * 2 + 1) * N) % 512u;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t bk = b_lo + (kk >> 2) * (N * 128 >> 4) + (kk & 3) * 2;
                    const uint32_t ak0 = a0 + (kk >> 2) * (128 * 128 >> 4) + (kk & 3) * 2, ak1 = ak0 + (2 * 128 * 128 >> 4);
                    if (kk == 0) { umma_ss<false>(d0, ak0, bk, idesc, leader); umma_ss<false>(d1, ak1, bk, idesc, leader); }
                    else { umma_ss<true>(d0, ak0, bk, idesc, leader); umma_ss<true>(d1, ak1, bk, idesc, leader); }
                }
                continue;
            }
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
                // accumulators: N = 128: 4 x 128 columns (as the matcher); N = 256: 2 x 256; TMEM-A variants keep A in the last 128 columns
                const uint32_t d = tmem + ((MODE >= 2) ? (uint32_t)(((t & 1) * 2 + rb) % ((MODE & 1) ? 1 : 3)) * N : (uint32_t)(((t & 1) * 2 + rb) * N) % 512u);
                const uint32_t a_lo = a0 + rb * (2 * 128 * 128 >> 4);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t bk = b_lo + (kk >> 2) * (N * 128 >> 4) + (kk & 3) * 2;
                    if (MODE < 2) {
                        const uint32_t ak = a_lo + (kk >> 2) * (128 * 128 >> 4) + (kk & 3) * 2;
                        if (kk == 0) umma_ss<false>(d, ak, bk, idesc, leader); else umma_ss<true>(d, ak, bk, idesc, leader);
                    } else {
                        const uint32_t at = tmem + 384u + (uint32_t)(rb * 64 + kk * 8);   // 8 columns (32 B) per K step
                        if (kk == 0) umma_ts<false>(d, at, bk, idesc, leader); else umma_ts<true>(d, at, bk, idesc, leader);
                    }
                }
                if (WRITERS >= 32) commit_if(&bar2[rb], leader);                 // one commit per 8 MMAs (the matcher's tfull)
                if (WRITERS == 33) {
                    uint32_t done;
                    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(smem_u32(&bar3)), "r"(1u) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (done == 12345u) sink[1] = 1;
                }
            }
            if (WRITERS >= 32) commit_if(&bar2[2], leader);                       // and one per tile (the matcher's empty)
        }
        commit_if(&bar, leader);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) cycles[0] = t1 - t0;
    } else if (warp >= 2 && WRITERS == 16) {
        // 16 "epilogue" warps: one packed 64-column TMEM load per warp per pace clocks (the matcher's epilogue reads every
        // accumulator once: 8 warps x 8 KB per 512-clock MMA group), plus `work` dependent VIMNMX to stand in for the scan
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 2) >> 2) * 64) % 512u;
        uint32_t acc = threadIdx.x;
        const int reps = tiles;   // one load per tile per warp, as in the matcher
        for (int r = 0; r < reps; ++r) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int m = 0; m < 32; ++m) acc = __vmaxs2(acc, v[m]);
            const long long t0 = clock64();
            while (clock64() - t0 < 700) { }   // ~ the rest of the 1024-clock tile period
        }
        if (acc == 0xdeadbeef) sink[0] = acc;
    } else if (warp >= 2 && WRITERS > 0 && WRITERS < 32) {
        // stream 128-bit stores into shared memory while the MMAs run (stands in for the TMA ring refills)
        uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
        uint4 *dst = (uint4 *)sW + (threadIdx.x - 64);
        const int reps = tiles * 8;   // tuned by the host: bytes written = reps * WRITERS * 512
        for (int r = 0; r < reps; ++r) { v.x += r; dst[(r & 3) * WRITERS * 32] = v; }
        if (v.x == 0xdeadbeef) sink[0] = dst[0].y;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int MODE, int WRITERS>
void run(const char *name, int sms, long long *d_cyc, uint32_t *sink, bool last = false)
{
    constexpr int N = (MODE & 1) ? 256 : 128;
    constexpr int NST = (MODE & 1) ? 2 : 4;
    const size_t smem = 65536 + (size_t)NST * 2 * N * 128 + 16384 + 1024;
    auto kern = k<MODE, WRITERS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int threads = WRITERS >= 32 ? 96 : 64 + WRITERS * 32 + (WRITERS == 0 ? 32 : 0);
    const int tiles = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<sms, threads, smem>>>(tiles, d_cyc, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a); kern<<<sms, threads, smem>>>(tiles, d_cyc, sink); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double mmas = (double)tiles * 16;
    const double ops = mmas * 128.0 * N * 32 * 2 * sms;
    printf("\"%s\": {\"clk_per_mma\": %.2f, \"int8_pops\": %.3f, \"ms\": %.3f, \"err\": \"%s\"}%s", name, cyc / mmas, ops / (best * 1e-3) / 1e15, best,
           cudaGetErrorString(cudaGetLastError()), last ? "" : ", ");
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    long long *d_cyc; uint32_t *sink; cudaMalloc(&d_cyc, 8); cudaMalloc(&sink, 4);
    printf("{\"gpu\": \"%s\", \"sms\": %d, ", p.name, p.multiProcessorCount);
    run<0, 0>("ss_n128", p.multiProcessorCount, d_cyc, sink);
    run<0, 2>("ss_n128_plus_smem_writes_2w", p.multiProcessorCount, d_cyc, sink);
    run<0, 8>("ss_n128_plus_smem_writes_8w", p.multiProcessorCount, d_cyc, sink);
    run<1, 0>("ss_n256", p.multiProcessorCount, d_cyc, sink);
    run<1, 8>("ss_n256_plus_smem_writes_8w", p.multiProcessorCount, d_cyc, sink);
    run<2, 0>("ts_n128", p.multiProcessorCount, d_cyc, sink);
    run<2, 8>("ts_n128_plus_smem_writes_8w", p.multiProcessorCount, d_cyc, sink);
    run<3, 0>("ts_n256", p.multiProcessorCount, d_cyc, sink);
    run<0, 16>("ss_n128_plus_16_ldtm_warps", p.multiProcessorCount, d_cyc, sink);
    run<0, 32>("ss_n128_commit_per_8_mma", p.multiProcessorCount, d_cyc, sink);
    run<0, 33>("ss_n128_commit_wait_fence_per_8_mma", p.multiProcessorCount, d_cyc, sink);
    run<1, 32>("ss_n256_commit_per_8_mma", p.multiProcessorCount, d_cyc, sink);
    run<0, 40>("ss_n128_two_accumulators_interleaved", p.multiProcessorCount, d_cyc, sink, true);
    printf("}\n");
    return 0;
}
