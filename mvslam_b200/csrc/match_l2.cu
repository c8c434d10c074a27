// match_l2.cu — float-descriptor (NORM_L2) brute-force kNN(2) for BASELINE config 4 (SURF-style
// 64-float descriptors, 32k x 32k): cv::BFMatcher(NORM_L2).knnMatch(k=2) semantics
// (reference source/vision/visual-feature.cpp:59-62 with a float VisualFeatureConfig::MatcherNormType).
//
//   d^2(q,t) = |q|^2 + |t|^2 - 2 q.t : the contraction S = Q T^T runs on the 5th-gen tensor cores
//   (tcgen05.mma kind::tf32, FP32 accumulators in TMEM, operands staged by TMA with 128B swizzle).  One more
//   K step of 8 columns adds the bias -|t|^2 / 2 inside the tensor core (the train side carries it split
//   into three TF32-exact pieces, the query side a constant 1 1 1 0 pattern), so the accumulator is
//   S' = q.t - |t|^2 / 2 = -a / 2 with a = |t|^2 - 2 q.t the approximate score: the epilogue reads the
//   accumulators back with tcgen05.ld and streams them through a running (best, second-best) pair per row
//   with three-input FMNMX only -- no multiply-add, no norm loads; every column whose score is within
//   2E of the running second-best is appended to the row's candidate list, E being a rigorous bound of the
//   TF32 rounding error of a (|S_tf32 - S| <= |q||t| 2^-9).  The 32k x 32k matrix is never materialised.
//   Any column that is NOT listed is therefore provably farther than both of the two approximately-best
//   columns, so the exact top-2 is inside the list: a second kernel re-ranks the listed columns with exact
//   FP32 distances (sequential fmaf over the dimension), (distance, index) tie-break.  Only a list
//   overflow (heavily duplicated data) sends a query to the exact brute-force kernel.
//
// Warp roles of the GEMM kernel (320 threads, 1 CTA per SM-resident tile of 128 queries):
//   warp 0   TMA producer (A once, B tiles through a STAGES-deep mbarrier ring)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer, double-buffered accumulators
//   warps 2-9 epilogue: thread <-> query row (TMEM lane) x half of the tile's columns; a 32-column chunk
//             costs 32 FFMA + a min tree + one warp vote unless some lane has a column to list
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "l2.h"

namespace mvs {

namespace {

constexpr int BM = 128;         // query rows per CTA  (UMMA M)
constexpr int BN = 128;         // train rows per tile (UMMA N)
constexpr int KSLAB = 32;       // floats per 128-byte swizzle slab
constexpr int KC = 64;          // candidate slots per (row, split, column-half) list
constexpr int EPI_WARPS = 8;         // two per TMEM lane quarter: each takes half of a tile's columns
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr uint32_t TMEM_COLS = 2 * BN;   // two accumulator buffers

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand, 128-byte swizzle, dense slab of [rows][128 B]: LBO = 16 B (unused for swizzled K-major),
// SBO = 1024 B (8 rows x 128 B), descriptor version 1 (Blackwell), layout_type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (>>4)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset (>>4)
    d |= (uint64_t)1 << 46;                 // version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// the bias slabs: [rows][32 B] (one K step of 8 TF32 columns), 32-byte swizzle, SBO = 256 B (8 rows x 32 B), layout_type 6
__device__ __forceinline__ uint64_t make_smem_desc32(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                 // SWIZZLE_32B
    return d;
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = BN
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate) : "memory");
}

// 32 accumulator columns of the thread's TMEM lane; asynchronous until tmem_wait64
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
// completes both loads; the registers are in/out operands so that no use of them is scheduled above the wait
__device__ __forceinline__ void tmem_wait64(uint32_t (&v)[32], uint32_t (&w)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]),
                   "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(w[18]), "+r"(w[19]), "+r"(w[20]), "+r"(w[21]), "+r"(w[22]), "+r"(w[23]), "+r"(w[24]), "+r"(w[25]), "+r"(w[26]), "+r"(w[27]), "+r"(w[28]), "+r"(w[29]), "+r"(w[30]), "+r"(w[31])
                 :: "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));   // FMNMX3
    return d;
}

// ------------------------------------------------------------------------------------------ small kernels
// squared norms (+inf padding) and their maximum (bit pattern order == value order for floats >= 0)
// and the bias row of every descriptor for the contraction: -|x|^2 / 2 as three TF32-exact pieces (13 low mantissa bits
// clear, so the tensor core's FP32 -> TF32 conversion keeps them whatever its rounding; what is lost is below 2^-30 |x|^2);
// padding rows get a large negative bias: such a column never reaches a candidate list
constexpr float kPadBias = -1e30f;
__device__ __forceinline__ float tf32_head(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void l2_norms_kernel(const float *x, int n, int n_padded, int ld, int dim, float *out, float *bias /* [n_padded][8] */,
                                unsigned int *max_bits)
{
    // 16 lanes per row: consecutive lanes read consecutive floats (the order of this sum is free: the norm only enters the
    // error band and the approximate score, never an exact distance)
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, sub = threadIdx.x & 15;
    float s = 0.f;
    if (i < n) {      // rows are 16-byte aligned with a pitch of a multiple of 32 floats, zero beyond dim
        const float4 *row = reinterpret_cast<const float4 *>(x + (size_t)i * ld);
#pragma unroll 2
        for (int k = sub; k < (ld >> 2); k += 16) {
            const float4 v = __ldg(row + k);
            s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        }
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    // maximum of the squared norms: one atomic per block, and only when it can matter (an atomic per warp on one word
    // serialised: ~10 us of a 14 us kernel)
    __shared__ unsigned int s_max[8];
    unsigned m = (i < n) ? __float_as_uint(s) : 0u;
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned bm = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) bm = max(bm, s_max[w]);
        if (bm > __ldcg(max_bits)) atomicMax(max_bits, bm);
    }
    if (sub != 0 || i >= n_padded) return;
    float4 *brow = reinterpret_cast<float4 *>(bias + (size_t)i * 8);
    brow[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= n) { out[i] = CUDART_INF_F; brow[0] = make_float4(kPadBias, 0.f, 0.f, 0.f); return; }
    out[i] = s;
    const float h = -0.5f * s;
    const float h0 = tf32_head(h), r0 = h - h0;          // exact: h0 is the head of h
    const float h1 = tf32_head(r0), r1 = r0 - h1;
    brow[0] = make_float4(h0, h1, tf32_head(r1), 0.f);
}

__global__ void l2_pad_kernel(const float *src, int n, int dim, float *dst, int ld)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * ld) return;
    const int r = (int)(i / ld), c = (int)(i % ld);
    dst[i] = c < dim ? src[(size_t)r * dim + c] : 0.f;
}

// Half-width of the candidate band: a bound of |a_tf32 - a| / 2 for a = |t|^2 - 2 q.t computed as -2 (q.t - |t|^2 / 2) by the
// TF32 contraction.  q.t: both operands keep 10 mantissa bits (2 x 2^-10 |q||t| when truncated; 0.0041 / 2 leaves a factor
// of two in hand); the bias pieces are TF32-exact (what they drop is below 2^-30 |t|^2); the FP32 accumulator holds values up
// to |q||t| + |t|^2 / 2 through nine additions (K / 8 + 1 instructions), each within 2^-23 of that: the 4e-6 term.
__device__ __forceinline__ float l2_half_band(float qn, float tmax)
{
    return qn * tmax * 0.0041f + 4e-6f * (tmax * tmax + 2.f * qn * tmax);
}

// ------------------------------------------------------------------------------------------ GEMM + top-KC
template <int KSLABS, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
l2_gemm_topk_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapT,
                    const __grid_constant__ CUtensorMap mapTb /* bias rows of the train side, padded to a multiple of BN */,
                    const float *__restrict__ q2, const unsigned int *__restrict__ tmax2_bits, int nq, int nt,
                    int tiles_per_split, int splits, float *__restrict__ cand_val, int32_t *__restrict__ cand_idx, int32_t *__restrict__ cand_cnt)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: A (KSLABS x 16 KB), B (STAGES x KSLABS x 16 KB), bias slabs (A: 4 KB constant, B: STAGES x 4 KB), barriers
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sB = sA + KSLABS * BM * 128;
    uint8_t *sAb = sB + STAGES * KSLABS * BN * 128;
    uint8_t *sBb = sAb + BM * 32;
    uint64_t *bars = (uint64_t *)(sBb + STAGES * BN * 32);
    uint64_t *barA = bars, *full = bars + 1, *empty = full + STAGES, *tfull = empty + STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BM;
    const int split = blockIdx.y;
    const int tiles_total = (nt + BN - 1) / BN;
    const int tile_begin = split * tiles_per_split;
    const int tile_end = min(tiles_total, tile_begin + tiles_per_split);
    const int n_tiles = max(0, tile_end - tile_begin);

    if (threadIdx.x == 0) {
        mbar_init(barA, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // query-side bias slab: every 16-byte chunk is (1, 1, 1, 0) -- the 32-byte swizzle only permutes 16-byte chunks, so the
    // pattern is its own swizzled image; written through the generic proxy, read by the tensor core through the async proxy
    for (int i = threadIdx.x; i < BM * 2; i += GEMM_THREADS) reinterpret_cast<float4 *>(sAb)[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && n_tiles > 0) {
            mbar_expect_tx(barA, KSLABS * BM * 128);
            for (int ks = 0; ks < KSLABS; ++ks) tma_load_2d(&mapQ, barA, sA + ks * BM * 128, ks * KSLAB, q0);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % STAGES;
                if (i >= STAGES) mbar_wait(empty + s, ((i / STAGES) - 1) & 1);
                mbar_expect_tx(full + s, KSLABS * BN * 128 + BN * 32);
                for (int ks = 0; ks < KSLABS; ++ks)
                    tma_load_2d(&mapT, full + s, sB + (s * KSLABS + ks) * BN * 128, ks * KSLAB, (tile_begin + i) * BN);
                tma_load_2d(&mapTb, full + s, sBb + s * BN * 32, 0, (tile_begin + i) * BN);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && n_tiles > 0) {
            mbar_wait(barA, 0);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % STAGES, acc = i & 1;
                if (i >= 2) mbar_wait(tempty + acc, ((i >> 1) - 1) & 1);   // epilogue drained this accumulator
                mbar_wait(full + s, (i / STAGES) & 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
                for (int ks = 0; ks < KSLABS; ++ks) {
                    const uint32_t a_addr = smem_u32(sA + ks * BM * 128);
                    const uint32_t b_addr = smem_u32(sB + (s * KSLABS + ks) * BN * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle atom
                        umma_tf32(d_tmem, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), (ks | k) ? 1u : 0u);
                }
                umma_tf32(d_tmem, make_smem_desc32(smem_u32(sAb)), make_smem_desc32(smem_u32(sBb + s * BN * 32)), 1u);   // - |t|^2 / 2
                tcgen05_commit(empty + s);     // B stage reusable once these MMAs retire
                tcgen05_commit(tfull + acc);   // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: thread <-> (row, column half) =====
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // which 64 columns of every 128-column tile
        const int row = quarter * 32 + lane;             // row within the tile == TMEM lane
        const int q = q0 + row;
        const size_t list = ((size_t)min(q, nq - 1) * splits + split) * 2 + half;
        float *lv = cand_val + list * KC;
        int32_t *li = cand_idx + list * KC;
        // The accumulator is S' = q.t - |t|^2 / 2 = -a / 2 (a = |t|^2 - 2 q.t, the approximate score): the scan keeps the two
        // LARGEST S' and lists every column with S' >= second - E, i.e. a <= second-best a + 2E.  E bounds the error of a / 2:
        // each operand keeps 10 mantissa bits, the bias pieces are exact, nine FP32 accumulation steps.
        const float qn = sqrtf(q2[min(q, nq - 1)]);
        const float tmax = sqrtf(__uint_as_float(*tmax2_bits));
        const float E = l2_half_band(qn, tmax);
        // rows beyond nq (zero-filled by TMA) never list anything: their threshold is +inf
        float c1 = -CUDART_INF_F, c2 = -CUDART_INF_F, thr = (q < nq) ? -CUDART_INF_F : CUDART_INF_F;
        int cnt = 0;
        for (int i = 0; i < n_tiles; ++i) {
            const int acc = i & 1;
            const int col0 = (tile_begin + i) * BN + half * (BN / 2);
            mbar_wait(tfull + acc, (i >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
            // both 32-column chunks are fetched at once and the accumulator goes back to the MMA warp as soon as they have
            // arrived (one arrival per warp): the tensor pipe never waits for the scan below
            uint32_t v0[32], v1[32];
            tmem_ld32_issue(taddr, v0);
            tmem_ld32_issue(taddr + 32, v1);
            tmem_wait64(v0, v1);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t (&v)[32] = c ? v1 : v0;
                const int c0 = 32 * c;
                float m4[8];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    m4[j4] = fmax3(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                   fmaxf(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])));
                const float m = fmax3(fmax3(m4[0], m4[1], m4[2]), fmax3(m4[3], m4[4], m4[5]), fmaxf(m4[6], m4[7]));
                if (__any_sync(0xFFFFFFFFu, m >= thr)) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        if (m4[j4] >= thr) {
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const float sv = __uint_as_float(v[4 * j4 + jj]);
                                const int col = col0 + c0 + 4 * j4 + jj;
                                if (sv >= thr && col < nt) {     // padding columns (large negative bias) pass only while thr = -inf
                                    if (cnt < KC) { lv[cnt] = -2.f * sv; li[cnt] = col; }
                                    ++cnt;
                                    const float lo = fminf(c1, sv);
                                    c1 = fmaxf(c1, sv);
                                    c2 = fmaxf(c2, lo);
                                    thr = c2 - E;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (q < nq) cand_cnt[list] = cnt;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ exact re-rank
// exact squared distance: sequential fmaf over the dimension (the order is part of the result's definition).  Rows are 16-byte
// aligned with a pitch of a multiple of 32 floats (zero padding beyond dim adds exact zeros), so they are fetched as float4 with
// several loads in flight: the sum is a latency chain, the loads need not be.
__device__ __forceinline__ float exact_d2(const float *__restrict__ q, const float *__restrict__ t, int dim)
{
    const float4 *q4 = reinterpret_cast<const float4 *>(q), *t4 = reinterpret_cast<const float4 *>(t);
    float s = 0.f;
    const int n4 = (dim + 3) >> 2;
#pragma unroll 4
    for (int k = 0; k < n4; ++k) {
        const float4 a = __ldg(q4 + k), b = __ldg(t4 + k);
        const float e0 = a.x - b.x, e1 = a.y - b.y, e2 = a.z - b.z, e3 = a.w - b.w;
        s = fmaf(e0, e0, s); s = fmaf(e1, e1, s); s = fmaf(e2, e2, s); s = fmaf(e3, e3, s);
    }
    return s;
}

__device__ __forceinline__ unsigned long long key_of(float d2, int idx)
{
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx;   // d2 >= 0: bit order == value order
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, k, off);
        k = o < k ? o : k;
    }
    return k;
}

// one warp per query: lists -> global second-best approximate score a2 -> exact distances of the entries
// with a <= a2 + 2E -> top-2 by (distance, index).  A list that overflowed flags the query for the exact kernel.
__global__ void l2_rerank_kernel(const float *__restrict__ Q, const float *__restrict__ T, int ldq, int ldt, int dim,
                                 const float *__restrict__ q2, const unsigned int *__restrict__ tmax2_bits, int nq, int nt,
                                 int lists,
                                 const float *__restrict__ cand_val, const int32_t *__restrict__ cand_idx,
                                 const int32_t *__restrict__ cand_cnt,
                                 int32_t *__restrict__ out_idx, float *__restrict__ out_d2, float *__restrict__ out_dist /* or null */,
                                 int32_t *__restrict__ fb_list, unsigned int *__restrict__ n_fallback)
{
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float qn = sqrtf(q2[q]);
    const float tmax = sqrtf(__uint_as_float(*tmax2_bits));
    const float twoE = 2.f * l2_half_band(qn, tmax);
    // The query's lists are one flat run of lists x KC slots; slot s belongs to list s / KC and is live below that list's count.
    // Lane l holds the count of list l (lists <= 32), the slots are walked four per lane at a time with their loads issued
    // together: the kernel is a chain of dependent memory round trips (counts -> scores -> indexes -> rows), so what matters
    // is that each link is one round trip, not one per slot.
    static_assert(KC == 64, "slot -> list mapping below assumes 64 slots per list");
    const int nslots = lists * KC;
    const float *cv = cand_val + (size_t)q * nslots;
    const int32_t *ci = cand_idx + (size_t)q * nslots;
    const int my_cnt = lane < lists ? cand_cnt[(size_t)q * lists + lane] : 0;
    const bool overflow = __any_sync(0xFFFFFFFFu, my_cnt > KC);
    float a1 = CUDART_INF_F, a2 = CUDART_INF_F;
    for (int base = 0; base < nslots; base += 128) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sl = base + 32 * j + lane;
            const int c = min(__shfl_sync(0xFFFFFFFFu, my_cnt, (base >> 6) + (j >> 1)), KC);
            v[j] = (sl < nslots && (sl & (KC - 1)) < c) ? __ldg(cv + sl) : CUDART_INF_F;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (v[j] < a1) { a2 = a1; a1 = v[j]; } else if (v[j] < a2) a2 = v[j];
        }
    }
    // warp-wide second smallest approximate score
    float m1 = a1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m1 = fminf(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, off));
    const int owner = __ffs(__ballot_sync(0xFFFFFFFFu, a1 == m1)) - 1;
    float m2 = (lane == owner) ? a2 : a1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m2 = fminf(m2, __shfl_xor_sync(0xFFFFFFFFu, m2, off));
    const float keep = m2 + twoE;
    unsigned long long b1 = ~0ull, b2 = ~0ull;
    const float *qrow = Q + (size_t)q * ldq;
    for (int base = 0; base < nslots; base += 128) {
        float v[4];
        int id[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sl = base + 32 * j + lane;
            const int c = min(__shfl_sync(0xFFFFFFFFu, my_cnt, (base >> 6) + (j >> 1)), KC);
            const bool live = sl < nslots && (sl & (KC - 1)) < c;
            v[j] = live ? __ldg(cv + sl) : CUDART_INF_F;       // cached by the first walk
            id[j] = live ? __ldg(ci + sl) : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!(v[j] <= keep)) continue;
            const unsigned long long k = key_of(exact_d2(qrow, T + (size_t)id[j] * ldt, dim), id[j]);
            if (k < b1) { b2 = b1; b1 = k; } else if (k < b2) b2 = k;
        }
    }
    const unsigned long long g1 = warp_min_u64(b1);
    const unsigned long long g2 = warp_min_u64((b1 == g1) ? b2 : b1);
    if (lane == 0) {
        out_idx[2 * q] = (g1 == ~0ull) ? -1 : (int)(g1 & 0xFFFFFFFFu);
        out_idx[2 * q + 1] = (g2 == ~0ull) ? -1 : (int)(g2 & 0xFFFFFFFFu);
        const float d1 = __uint_as_float((unsigned)(g1 >> 32)), d2 = __uint_as_float((unsigned)(g2 >> 32));
        out_d2[2 * q] = d1;
        out_d2[2 * q + 1] = d2;
        if (out_dist) { out_dist[2 * q] = sqrtf(d1); out_dist[2 * q + 1] = sqrtf(d2); }
        if (overflow) fb_list[atomicAdd(n_fallback, 1u)] = q;      // rare: the exact kernel redoes this query
    }
}

// exact brute force for the queries whose candidate list overflowed: the CTAs walk the list the re-rank kernel wrote
__global__ void l2_fallback_kernel(const float *__restrict__ Q, const float *__restrict__ T, int ldq, int ldt, int dim,
                                   int nt, const int32_t *__restrict__ fb_list, const unsigned int *__restrict__ n_fallback,
                                   int32_t *__restrict__ out_idx, float *__restrict__ out_d2, float *__restrict__ out_dist /* or null */)
{
    __shared__ unsigned long long s1[8], s2[8];
    const unsigned int n = *n_fallback;
    for (unsigned int item = blockIdx.x; item < n; item += gridDim.x) {
        const int q = fb_list[item];
        unsigned long long b1 = ~0ull, b2 = ~0ull;
        for (int t = threadIdx.x; t < nt; t += blockDim.x) {
            const unsigned long long k = key_of(exact_d2(Q + (size_t)q * ldq, T + (size_t)t * ldt, dim), t);
            if (k < b1) { b2 = b1; b1 = k; } else if (k < b2) b2 = k;
        }
        const unsigned long long g1 = warp_min_u64(b1);
        const unsigned long long g2 = warp_min_u64((b1 == g1) ? b2 : b1);
        if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = g1; s2[threadIdx.x >> 5] = g2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long f1 = ~0ull, f2 = ~0ull;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
                const unsigned long long c[2] = {s1[w], s2[w]};
                for (int j = 0; j < 2; ++j) { if (c[j] < f1) { f2 = f1; f1 = c[j]; } else if (c[j] < f2) f2 = c[j]; }
            }
            const float d1 = __uint_as_float((unsigned)(f1 >> 32)), d2 = __uint_as_float((unsigned)(f2 >> 32));
            out_idx[2 * q] = (f1 == ~0ull) ? -1 : (int)(f1 & 0xFFFFFFFFu);
            out_idx[2 * q + 1] = (f2 == ~0ull) ? -1 : (int)(f2 & 0xFFFFFFFFu);
            out_d2[2 * q] = d1;
            out_d2[2 * q + 1] = d2;
            if (out_dist) { out_dist[2 * q] = sqrtf(d1); out_dist[2 * q + 1] = sqrtf(d2); }
        }
        __syncthreads();
    }
}

// Lowe ratio + max_dist (+ cross-check) on float distances, then sort by (distance, queryIdx); one CTA.
// keys live in global memory (up to 2^22 queries); 64-bit key = distance bits << 32 | query.
__global__ void __launch_bounds__(1024)
l2_filter_sort_kernel(const int32_t *idx, const float *dist, int nq, double ratio, double max_dist,
                      const int32_t *rev_idx, unsigned long long *keys, mvs_match *out, int32_t *n_out)
{
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        const int i1 = idx[2 * q], i2 = idx[2 * q + 1];
        if (i1 < 0 || i2 < 0) continue;
        const double d1 = (double)dist[2 * q], d2 = (double)dist[2 * q + 1];
        bool keep = (d1 < ratio * d2) && ((max_dist < 0) || (d1 <= max_dist));   // visual-feature.cpp:66-68
        if (keep && rev_idx) keep = (rev_idx[2 * i1] == q);
        if (keep) keys[atomicAdd(&s_count, 1)] = ((unsigned long long)__float_as_uint(dist[2 * q]) << 32) | (unsigned)q;
    }
    __syncthreads();
    const int m = s_count;
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = m + threadIdx.x; i < n2; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long x = keys[i], y = keys[l];
                    if ((x > y) == ((i & k) == 0)) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int q = (int)(keys[i] & 0xFFFFFFFFu);
        mvs_match mm;
        mm.query = q; mm.train = idx[2 * q]; mm.distance = dist[2 * q];
        out[i] = mm;
    }
    if (threadIdx.x == 0) *n_out = m;
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows][ld] float matrix, box = 32 floats (one 128-byte swizzle slab) x box_rows; out-of-range rows read as zero
bool make_map(CUtensorMap *m, const float *ptr, int rows, int ld, int kpad, int box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)KSLAB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows][8] float bias rows, box = one 32-byte row x BN rows, 32-byte swizzle (the layout make_smem_desc32 describes)
bool make_bias_map(CUtensorMap *m, const float *ptr, int rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {8, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {8 * sizeof(float)};
    cuuint32_t box[2] = {8, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

enum { B_Q = 0, B_T, B_NORM, B_CAND_V, B_CAND_I, B_OUT, B_MISC, B_KEYS };

cudaError_t ensure(L2Workspace &ws, int i, size_t bytes)
{
    if (bytes <= ws.cap[i]) return cudaSuccess;
    if (ws.buf[i]) cudaFree(ws.buf[i]);
    ws.buf[i] = nullptr; ws.cap[i] = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&ws.buf[i], want);
    if (e == cudaSuccess) ws.cap[i] = want;
    return e;
}

template <int KSLABS>
cudaError_t launch_gemm(const CUtensorMap &mq, const CUtensorMap &mt, const CUtensorMap &mtb, const float *q2, const unsigned int *tmax, int nq,
                        int nt, int tiles_per_split, int splits, float *cv, int32_t *ci, int32_t *cc, cudaStream_t s)
{
    constexpr int STAGES = 2;   // 2 CTAs per SM (<= 113 KB each): one CTA's MMA/TMA overlaps the other's epilogue
    const size_t smem = 1024 + (size_t)KSLABS * BM * 128 + (size_t)STAGES * KSLABS * BN * 128 + (size_t)BM * 32 + (size_t)STAGES * BN * 32 +
                        (1 + 2 * STAGES + 4) * sizeof(uint64_t) + 16;
    auto kern = l2_gemm_topk_kernel<KSLABS, STAGES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((nq + BM - 1) / BM, splits);
    kern<<<grid, GEMM_THREADS, smem, s>>>(mq, mt, mtb, q2, tmax, nq, nt, tiles_per_split, splits, cv, ci, cc);
    return cudaGetLastError();
}

#define L2CK(call)                                                                     \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e__); return MVS_E_CUDA; } \
    } while (0)

// train-dimension splits.  Every split restarts the per-row candidate stream, and short streams make the
// epilogue's slow path fire more often (measured on 32k x 32k: 1 split 0.35 ms, 2: 0.39, 4: 0.46, 8: 0.52), so:
// as few splits as still give every SM a CTA or two.
void plan_splits(int na, int nb, int &splits, int &tiles_per_split)
{
    const int tiles_total = (nb + BN - 1) / BN;
    const int qtiles = (na + BM - 1) / BM;
    const int smax = std::max(1, std::min(16, tiles_total / 8));
    splits = std::max(1, std::min(smax, (222 + qtiles - 1) / qtiles));
    if (const char *e = getenv("MVS_L2_SPLITS")) splits = std::max(1, std::min(smax, atoi(e)));   // tuning override
    tiles_per_split = (tiles_total + splits - 1) / splits;
    splits = (tiles_total + tiles_per_split - 1) / tiles_per_split;
}

// tensor-core pass: top-KC candidates per (row of A, split of B) into ws.buf[B_CAND_V/B_CAND_I]
int gemm_candidates(L2Workspace &ws, cudaStream_t stream, const float *dA, int na, const float *dB, int nb, int ld,
                    int kpad, const float *d_biasB /* [nb rounded up to BN][8] */, const float *d_normA, const unsigned int *bmax, int splits,
                    int tiles_per_split, std::string &err)
{
    L2CK(ensure(ws, B_CAND_V, (size_t)na * splits * 2 * KC * sizeof(float)));
    L2CK(ensure(ws, B_CAND_I, (size_t)na * splits * 2 * (KC + 1) * sizeof(int32_t)));
    CUtensorMap mq, mt, mtb;
    if (!make_map(&mq, dA, na, ld, kpad, BM) || !make_map(&mt, dB, nb, ld, kpad, BN) || !make_bias_map(&mtb, d_biasB, (nb + BN - 1) / BN * BN)) {
        err = "cuTensorMapEncodeTiled failed";
        return MVS_E_CUDA;
    }
    float *cv = (float *)ws.buf[B_CAND_V];
    int32_t *ci = (int32_t *)ws.buf[B_CAND_I];
    int32_t *cc = ci + (size_t)na * splits * 2 * KC;
    cudaError_t e;
    switch (kpad / KSLAB) {
    case 1: e = launch_gemm<1>(mq, mt, mtb, d_normA, bmax, na, nb, tiles_per_split, splits, cv, ci, cc, stream); break;
    case 2: e = launch_gemm<2>(mq, mt, mtb, d_normA, bmax, na, nb, tiles_per_split, splits, cv, ci, cc, stream); break;
    case 3: e = launch_gemm<3>(mq, mt, mtb, d_normA, bmax, na, nb, tiles_per_split, splits, cv, ci, cc, stream); break;
    case 4: e = launch_gemm<4>(mq, mt, mtb, d_normA, bmax, na, nb, tiles_per_split, splits, cv, ci, cc, stream); break;
    default: err = "descriptor dimension above 128 floats"; return MVS_E_UNSUPPORTED;
    }
    L2CK(e);
    return MVS_OK;
}

}  // namespace

void L2Workspace::release()
{
    for (int i = 0; i < 8; ++i) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
    for (int i = 0; i < 6; ++i) { if (ev[i]) cudaEventDestroy(ev[i]); ev[i] = nullptr; }
}

int l2_knn2(L2Workspace &ws, cudaStream_t stream, const float *query, int nq, const float *train, int nt, int dim,
            int32_t *idx, float *dist, const mvs_match_params *mp, mvs_match *out, int capacity, int *n_out,
            int *n_launches, std::string &err)
{
    int nl = 0;
    if (n_launches) *n_launches = 0;
    for (int i = 0; i < 6; ++i) if (!ws.ev[i]) cudaEventCreate(&ws.ev[i]);
    cudaEventRecord(ws.ev[0], stream);
    float gemm_ms_total = 0.f;
    if (n_out) *n_out = 0;
    if (dim < 1 || dim > 128) { err = "float descriptors: 1 <= dim <= 128 supported"; return MVS_E_UNSUPPORTED; }
    if (nq > (1 << 22) || nt > (1 << 22)) { err = "more than 2^22 descriptors per side"; return MVS_E_UNSUPPORTED; }
    const int kpad = ((dim + KSLAB - 1) / KSLAB) * KSLAB;
    const bool cross = mp && mp->cross_check;
    // Operands: descriptor sets that are already in device memory with a TMA-compatible layout (rows of a multiple of 32 floats,
    // 16-byte aligned) are read where they are; host buffers are copied, and rows of another length are padded to kpad floats.
    auto resident = [&](const float *p) {
        if (kpad != dim || ((uintptr_t)p & 15u)) return false;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeDevice;
    };
    const float *dQ = query, *dT = train;
    const float *const src[2] = {query, train};
    const float **dst[2] = {&dQ, &dT};
    const int rows[2] = {nq, nt};
    for (int k = 0; k < 2; ++k) {
        if (resident(src[k])) continue;
        const int bi = k == 0 ? B_Q : B_T;
        L2CK(ensure(ws, bi, (size_t)rows[k] * kpad * sizeof(float) + (kpad == dim ? 0 : (size_t)rows[k] * dim * sizeof(float))));
        float *d = (float *)ws.buf[bi];
        if (kpad == dim) {
            L2CK(cudaMemcpyAsync(d, src[k], (size_t)rows[k] * dim * sizeof(float), cudaMemcpyDefault, stream));
        } else {
            float *raw = d + (size_t)rows[k] * kpad;
            L2CK(cudaMemcpyAsync(raw, src[k], (size_t)rows[k] * dim * sizeof(float), cudaMemcpyDefault, stream));
            l2_pad_kernel<<<(unsigned)(((size_t)rows[k] * kpad + 255) / 256), 256, 0, stream>>>(raw, rows[k], dim, d, kpad);
            nl += 1;
        }
        *dst[k] = d;
    }
    // norms, |t|max (host reduction of nt floats: part of the error bound, not of the distance computation)
    // squared norms, each array padded with +inf to a multiple of the column tile (padded columns never win);
    // their maxima stay on the device (error bound of the TF32 contraction)
    const size_t nqp = ((size_t)nq + BN - 1) / BN * BN, ntp = ((size_t)nt + BN - 1) / BN * BN;
    L2CK(ensure(ws, B_NORM, (nqp + ntp) * 9 * sizeof(float)));   // squared norms, then the [row][8] bias rows of the contraction
    L2CK(ensure(ws, B_MISC, 128));
    unsigned int *d_nfb = (unsigned int *)ws.buf[B_MISC];      // [0..1] fallback counters, [4] n_out, [8..9] max |q|^2, |t|^2
    int32_t *d_nout = (int32_t *)(d_nfb + 4);
    unsigned int *d_max = d_nfb + 8;
    L2CK(cudaMemsetAsync(d_nfb, 0, 128, stream));
    float *nQ = (float *)ws.buf[B_NORM], *nT = nQ + nqp, *bQ = nT + ntp, *bT = bQ + 8 * nqp;
    l2_norms_kernel<<<(unsigned)((nqp * 16 + 255) / 256), 256, 0, stream>>>(dQ, nq, (int)nqp, kpad, dim, nQ, bQ, d_max);
    l2_norms_kernel<<<(unsigned)((ntp * 16 + 255) / 256), 256, 0, stream>>>(dT, nt, (int)ntp, kpad, dim, nT, bT, d_max + 1);
    nl += 2;

    const int passes = cross ? 2 : 1;
    // outputs: forward idx/d2/dist/fallback list, reverse idx/d2/fallback list
    const size_t per_f = (size_t)nq * (2 * 4 + 2 * 4 + 2 * 4 + 4) + 64, per_r = (size_t)nt * (2 * 4 + 2 * 4 + 4) + 64;
    L2CK(ensure(ws, B_OUT, per_f + per_r + 64));
    uint8_t *ob = (uint8_t *)ws.buf[B_OUT];
    int32_t *f_idx = (int32_t *)ob; float *f_d2 = (float *)(f_idx + 2 * (size_t)nq); float *f_dist = f_d2 + 2 * (size_t)nq;
    int32_t *f_list = (int32_t *)(f_dist + 2 * (size_t)nq);
    uint8_t *rb = ob + ((per_f + 15) & ~(size_t)15);
    int32_t *r_idx = (int32_t *)rb; float *r_d2 = (float *)(r_idx + 2 * (size_t)nt); int32_t *r_list = (int32_t *)(r_d2 + 2 * (size_t)nt);

    for (int pass = 0; pass < passes; ++pass) {
        const float *A = pass == 0 ? dQ : dT, *Bm = pass == 0 ? dT : dQ;
        const int na = pass == 0 ? nq : nt, nb = pass == 0 ? nt : nq;
        float *nA = pass == 0 ? nQ : nT, *biasB = pass == 0 ? bT : bQ;
        int32_t *o_idx = pass == 0 ? f_idx : r_idx; float *o_d2 = pass == 0 ? f_d2 : r_d2; int32_t *o_list = pass == 0 ? f_list : r_list;
        float *o_dist = pass == 0 ? f_dist : nullptr;
        const unsigned int *bmax = pass == 0 ? d_max + 1 : d_max;
        int splits, tps;
        plan_splits(na, nb, splits, tps);
        cudaEventRecord(ws.ev[2 + 2 * pass], stream);
        int st = gemm_candidates(ws, stream, A, na, Bm, nb, kpad, kpad, biasB, nA, bmax, splits, tps, err);
        if (st != MVS_OK) return st;
        cudaEventRecord(ws.ev[3 + 2 * pass], stream);
        nl += 1;
        l2_rerank_kernel<<<(na + 7) / 8, 256, 0, stream>>>(A, Bm, kpad, kpad, dim, nA, bmax, na, nb, splits * 2,
                                                           (const float *)ws.buf[B_CAND_V], (const int32_t *)ws.buf[B_CAND_I],
                                                           (const int32_t *)ws.buf[B_CAND_I] + (size_t)na * splits * 2 * KC,
                                                           o_idx, o_d2, o_dist, o_list, d_nfb + pass);
        // queries whose candidate list overflowed (heavily duplicated data): exact brute force over the list the re-rank wrote
        l2_fallback_kernel<<<std::min(na, 592), 256, 0, stream>>>(A, Bm, kpad, kpad, dim, nb, o_list, d_nfb + pass, o_idx, o_d2, o_dist);
        nl += 2;
    }
    L2CK(cudaGetLastError());
    if (idx) L2CK(cudaMemcpyAsync(idx, f_idx, (size_t)nq * 2 * sizeof(int32_t), cudaMemcpyDefault, stream));   // host or device (UVA)
    if (dist) L2CK(cudaMemcpyAsync(dist, f_dist, (size_t)nq * 2 * sizeof(float), cudaMemcpyDefault, stream));
    if (mp && n_out) {
        int n2 = 1;
        while (n2 < nq) n2 <<= 1;
        L2CK(ensure(ws, B_KEYS, (size_t)n2 * sizeof(unsigned long long) + (size_t)nq * sizeof(mvs_match)));
        unsigned long long *keys = (unsigned long long *)ws.buf[B_KEYS];
        mvs_match *dm = (mvs_match *)(keys + n2);
        l2_filter_sort_kernel<<<1, 1024, 0, stream>>>(f_idx, f_dist, nq, mp->ratio, mp->max_dist, cross ? r_idx : nullptr, keys, dm, d_nout);
        nl += 1;
        L2CK(cudaGetLastError());
        int32_t m = 0;
        L2CK(cudaMemcpyAsync(&m, d_nout, sizeof(m), cudaMemcpyDeviceToHost, stream));
        L2CK(cudaStreamSynchronize(stream));
        *n_out = m;
        if (m > capacity) { err = "match capacity too small"; return MVS_E_CAPACITY; }
        if (m > 0) {
            if (!out) { err = "null output"; return MVS_E_BAD_ARG; }
            L2CK(cudaMemcpyAsync(out, dm, (size_t)m * sizeof(mvs_match), cudaMemcpyDeviceToHost, stream));
        }
    }
    cudaEventRecord(ws.ev[1], stream);
    unsigned int hfb[2] = {0, 0};
    L2CK(cudaMemcpyAsync(hfb, d_nfb, sizeof(hfb), cudaMemcpyDeviceToHost, stream));
    L2CK(cudaStreamSynchronize(stream));
    float tot_ms = 0.f;
    cudaEventElapsedTime(&tot_ms, ws.ev[0], ws.ev[1]);
    for (int pass = 0; pass < passes; ++pass) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ws.ev[2 + 2 * pass], ws.ev[3 + 2 * pass]) == cudaSuccess) gemm_ms_total += ms;
    }
    ws.stats[0] = hfb[0]; ws.stats[1] = hfb[1];
    ws.stats[2] = (uint64_t)(gemm_ms_total * 1000.f); ws.stats[3] = (uint64_t)(tot_ms * 1000.f);
    if (n_launches) *n_launches = nl;
    return MVS_OK;
}

}  // namespace mvs
