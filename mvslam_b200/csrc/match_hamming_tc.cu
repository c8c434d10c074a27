// match_hamming_tc.cu — 256-bit Hamming kNN(2) as a dense contraction on the 5th-gen tensor cores.
//
// Same contract as knn2_hamming_kernel (match_hamming.cu; cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) of
// reference source/vision/visual-feature.cpp:59-62): per query the two smallest (distance, trainIdx) keys, bit-exact.
//
//   Every descriptor bit b is stored as the signed byte s(b) = +8 / -8 (expand_desc_kernel, 256 B per descriptor).
//   For two descriptors  sum_k s(q_k) s(t_k) = 64 (256 - 2 hamming(q, t)):  the products are +-64 and the partial sums
//   small integers, exact in the S32 accumulators of kind::i8, so  hamming  is the popcount distance, not an
//   approximation.  (kind::f8f6f4 with E4M3 +-1 was measured too: exact as well, 13 % slower.)
//
//   The contraction runs as tcgen05.mma (M = N = 128, K = 8 x 32, plus a ninth K step that adds the column index, see
//   the kernel) with both operands staged by TMA (128-byte swizzle) and the accumulators double-buffered in TMEM.  The
//   epilogue never materialises the score matrix: thread <-> (query row, column half) keeps a running (best, second)
//   pair on packed 16-bit keys, 1.75 ALU instructions per accumulator.
//
// Warp roles (320 threads, 2 CTAs per SM so that one CTA's TMA/MMA overlaps the other's epilogue):
//   warp 0   TMA producer (query tile once, train tiles through a STAGES-deep mbarrier ring)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-9 epilogue
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace mvs {

namespace {

constexpr int BM = 128;              // query rows per CTA  (UMMA M)
constexpr int BN = 128;              // train rows per tile (UMMA N)
constexpr int SLAB = 128;            // bytes of K per 128-byte swizzle slab
constexpr int KSLABS = 2;            // 256 one-byte elements per descriptor
constexpr int EPI_WARPS = 8;         // two per TMEM lane quarter: each takes half of a tile's columns
constexpr int TC_THREADS = 64 + EPI_WARPS * 32;
constexpr uint32_t TMEM_COLS = 2 * BN;   // two accumulator buffers
constexpr int TC_MAX_TRAIN = 32768;  // trainIdx field of the thread's running key (hamming * 32768 + trainIdx)

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

// K-major operand, 128-byte swizzle, dense slab of [rows][128 B]: LBO unused, SBO = 1024 B (8 rows x 128 B),
// descriptor version 1 (Blackwell), layout_type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: kind::i8, S8 x S8 -> S32, A and B K-major, M = 128, N = BN
constexpr uint32_t kInstrDescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDescI8), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void top2(uint32_t &b1, uint32_t &b2, uint32_t key)
{
    const uint32_t hi = max(b1, key);
    b1 = min(b1, key);
    b2 = min(b2, hi);
}

// ------------------------------------------------------------------------------------------ bits -> +-8 bytes
__global__ void __launch_bounds__(256)
expand_desc_kernel(const uint32_t *__restrict__ desc, size_t word_begin, size_t n_words, uint4 *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t w = __ldg(desc + word_begin + i);
    uint32_t o[8];
#pragma unroll
    for (int nib = 0; nib < 8; ++nib) {
        const uint32_t spread = (((w >> (4 * nib)) & 15u) * 0x00204081u) & 0x01010101u;   // bit k -> byte k
        o[nib] = 0xF8F8F8F8u ^ (spread * 0xF0u);                                           // 1 -> +8, 0 -> -8
    }
    uint4 *dst = out + (word_begin + i) * 2;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ------------------------------------------------------------------------------------------ GEMM + running top-2
// Epilogue arithmetic.  With every bit stored as +-8 the contraction gives 64 S = 128 (S/2); a ninth K step over one
// constant slab (query side: 1 in byte 0 of every row; train side: 127 - column-in-tile in byte 32 of every row) adds
// 127 - c, so the accumulator itself is the signed 16-bit key
//      k16 = 128 (128 - hamming) + (127 - c)          in [-16384, 16511]
// that orders the columns of a tile by (smaller distance, then smaller index) under MAX.  Two accumulators are packed
// into one register with a single PRMT and the running (best, second) pair of both 16-bit lanes costs 2.5 VIMNMX.S16x2
// per register: 1.75 ALU instructions per accumulator and none on the FMA pipe.  At the end of a tile the four lane
// results are widened to  hamming * 32768 + trainIdx  and merged into the thread's 32-bit pair (minimum = best).
__device__ __forceinline__ uint32_t widen_key(uint32_t k16, uint32_t tile_base)
{
    const int k = (int)(short)k16;                         // empty lane: -32768 -> distance 384, dropped at the export
    return (uint32_t)(128 - (k >> 7)) * 32768u + tile_base + (127u - ((uint32_t)k & 127u));
}

template <int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 2)
knn2_hamming_tc_kernel(const __grid_constant__ CUtensorMap map, TcKnnArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int pair = blockIdx.z;
    int fq, ft;
    if (a.pairs) {  // query = pair frame (second), train = base frame (first): visual-feature.cpp:59-60
        const int2 pr = a.pairs[pair];
        fq = a.reverse ? pr.x : pr.y;
        ft = a.reverse ? pr.y : pr.x;
    } else { fq = a.reverse ? 0 : 1; ft = a.reverse ? 1 : 0; }
    const int nq = a.frame_cnt[fq], nt = a.frame_cnt[ft];
    const int q0 = blockIdx.x * BM;
    if (q0 >= nq) return;                                   // whole CTA, before any barrier or TMEM allocation
    const int row_q = a.frame_off[fq] + q0, row_t = a.frame_off[ft];
    const int n_tiles = (nt + BN - 1) / BN;

    // the launch requests no static shared memory, so the dynamic window starts 1024-byte aligned (checked)
    uint8_t *sA = smem_raw;
    uint8_t *sB = sA + KSLABS * BM * SLAB;
    uint8_t *sX = sB + STAGES * KSLABS * BN * SLAB;        // constant slab of the ninth K step
    uint64_t *bars = (uint64_t *)(sX + BN * SLAB);
    uint64_t *barA = bars, *full = bars + 1, *empty = full + STAGES, *tfull = empty + STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if (smem_u32(smem_raw) & 1023u) __trap();
        mbar_init(barA, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // constant slab in the 128-byte-swizzle layout (16-byte chunk index XOR row mod 8): row r holds 1 at byte 0 (read as
    // the query operand, K bytes 0..31) and 127 - r at byte 32 (read as the train operand, K bytes 32..63)
    for (int i = threadIdx.x; i < BN * SLAB / 16; i += TC_THREADS) {
        const int r = i >> 3, chunk = (i & 7) ^ (r & 7);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (chunk == 0) v.x = 1u;
        if (chunk == 2) v.x = (uint32_t)(127 - r);
        reinterpret_cast<uint4 *>(sX)[i] = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the tensor core
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(barA, KSLABS * BM * SLAB);
            for (int ks = 0; ks < KSLABS; ++ks) tma_load_2d(&map, barA, sA + ks * BM * SLAB, ks * SLAB, row_q);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % STAGES;
                if (i >= STAGES) mbar_wait(empty + s, ((i / STAGES) - 1) & 1);
                mbar_expect_tx(full + s, KSLABS * BN * SLAB);
                for (int ks = 0; ks < KSLABS; ++ks)
                    tma_load_2d(&map, full + s, sB + (s * KSLABS + ks) * BN * SLAB, ks * SLAB, row_t + i * BN);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            mbar_wait(barA, 0);
            const uint64_t x_a = make_smem_desc(smem_u32(sX)), x_b = make_smem_desc(smem_u32(sX) + 32);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % STAGES, acc = i & 1;
                if (i >= 2) mbar_wait(tempty + acc, ((i >> 1) - 1) & 1);   // epilogue drained this accumulator
                mbar_wait(full + s, (i / STAGES) & 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
#pragma unroll
                for (int ks = 0; ks < KSLABS; ++ks) {
                    const uint32_t a_addr = smem_u32(sA + ks * BM * SLAB);
                    const uint32_t b_addr = smem_u32(sB + (s * KSLABS + ks) * BN * SLAB);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // UMMA_K = 32 one-byte elements = 32 bytes inside the swizzle atom
                        umma_i8(d_tmem, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), (ks | k) ? 1u : 0u);
                }
                tcgen05_commit(empty + s);     // train stage reusable once these MMAs retire
                umma_i8(d_tmem, x_a, x_b, 1u);  // + (127 - column): the index half of the key
                tcgen05_commit(tfull + acc);   // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: thread <-> (query row, column half) =====
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // which 64 columns of every 128-column tile
        const int q = q0 + quarter * 32 + lane;          // row within the tile == TMEM lane
        uint32_t g1 = kKeyNone, g2 = kKeyNone;
        for (int i = 0; i < n_tiles; ++i) {
            const int acc = i & 1;
            const int col0 = i * BN + half * (BN / 2);
            mbar_wait(tfull + acc, (i >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
            const bool ragged = col0 + BN / 2 > nt;      // warp-uniform: only a frame's last tile
            uint32_t a1 = 0x80008000u, a2 = 0x80008000u, b1 = 0x80008000u, b2 = 0x80008000u;   // two chains for ILP
#pragma unroll 1
            for (int c0 = 0; c0 < BN / 2; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                if (!ragged) {
#pragma unroll
                    for (int m = 0; m < 16; m += 2) {
                        const uint32_t pa = __byte_perm(v[2 * m], v[2 * m + 1], 0x5410);
                        const uint32_t pb = __byte_perm(v[2 * m + 2], v[2 * m + 3], 0x5410);
                        const uint32_t la = __vmins2(a1, pa), lb = __vmins2(b1, pb);
                        a1 = __vmaxs2(a1, pa); b1 = __vmaxs2(b1, pb);
                        a2 = __vmaxs2(a2, la); b2 = __vmaxs2(b2, lb);
                    }
                } else {                                 // rows of the next frame / zero fill lose to every real key
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int c = col0 + c0 + 2 * m;
                        const uint32_t pk = __byte_perm(c < nt ? v[2 * m] : 0x8000u, c + 1 < nt ? v[2 * m + 1] : 0x8000u, 0x5410);
                        const uint32_t lo = __vmins2(a1, pk);
                        a1 = __vmaxs2(a1, pk); a2 = __vmaxs2(a2, lo);
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tempty + acc);
            // the tile's two best of each lane pair -> the thread's running pair
            const uint32_t n1 = __vmaxs2(a1, b1), n2 = __vimax3_s16x2(__vmins2(a1, b1), a2, b2);
            const uint32_t tile_base = (uint32_t)(i * BN);
            top2(g1, g2, widen_key(n1 & 0xFFFFu, tile_base));
            top2(g1, g2, widen_key(n1 >> 16, tile_base));
            top2(g1, g2, widen_key(n2 & 0xFFFFu, tile_base));
            top2(g1, g2, widen_key(n2 >> 16, tile_base));
        }
        if (q < nq) {
            const uint32_t x1 = (g1 >> 15) > 256u ? kKeyNone : (((g1 >> 15) << kIdxBits) | (g1 & 32767u));
            const uint32_t x2 = (g2 >> 15) > 256u ? kKeyNone : (((g2 >> 15) << kIdxBits) | (g2 & 32767u));
            a.partial[((size_t)pair * 2 + half) * a.q_stride + q] = make_uint2(x1, x2);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
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

cudaError_t launch_tc(const CUtensorMap &map, const TcKnnArgs &a, int max_nq, int n_pairs, cudaStream_t s)
{
    constexpr int STAGES = 2;   // 2 CTAs per SM (<= 113 KB each)
    const size_t smem = (size_t)KSLABS * BM * SLAB + (size_t)STAGES * KSLABS * BN * SLAB + (size_t)BN * SLAB +
                        (1 + 2 * STAGES + 4) * sizeof(uint64_t) + 16;
    auto kern = knn2_hamming_tc_kernel<STAGES>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((max_nq + BM - 1) / BM, 1, n_pairs);
    kern<<<grid, TC_THREADS, smem, s>>>(map, a);
    return cudaGetLastError();
}

}  // namespace

int tc_max_train() { return TC_MAX_TRAIN; }

void launch_expand_desc(const uint4 *desc, size_t row_begin, size_t n_rows, void *desc8, cudaStream_t s)
{
    if (!n_rows) return;
    const size_t n_words = n_rows * 8;
    expand_desc_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>((const uint32_t *)desc, row_begin * 8, n_words, (uint4 *)desc8);
}

cudaError_t launch_knn2_hamming_tc(const void *desc8, size_t total_rows, const TcKnnArgs &a, int max_nq, int n_pairs, cudaStream_t s)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return cudaErrorNotSupported;
    CUtensorMap map;
    cuuint64_t gdim[2] = {256, (cuuint64_t)total_rows};
    cuuint64_t gstr[1] = {256};
    cuuint32_t box[2] = {(cuuint32_t)SLAB, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(desc8), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    return launch_tc(map, a, max_nq, n_pairs, s);
}

}  // namespace mvs
