// match_hamming_tc.cu — 256-bit Hamming kNN(2) as a dense contraction on the 5th-gen tensor cores.
//
// Same contract as knn2_hamming_kernel (match_hamming.cu; cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) of
// reference source/vision/visual-feature.cpp:59-62): per query the two smallest (distance, trainIdx) keys, bit-exact.
//
//   Every descriptor bit b is stored as the 8-bit value s(b) = +1 / -1 (expand_desc_kernel, 256 B per descriptor).
//   For two descriptors  S = sum_k s(q_k) s(t_k) = 256 - 2 hamming(q, t),  an even integer in [-256, 256]: the
//   products are +-1 and the partial sums are small integers, so the accumulation is exact in S32 (kind::i8) and in
//   FP32 (kind::f8f6f4, E4M3 +-1.0) alike and  hamming = 128 - S/2  is the popcount distance, not an approximation.
//
//   S = Q T^T runs as tcgen05.mma (M = N = 128, K = 8 x 32) with both operands staged by TMA (128-byte swizzle) and
//   the accumulators double-buffered in TMEM.  The epilogue never materialises S: thread <-> (query row, column half)
//   turns each accumulator into the sortable key  hamming * 32768 + trainIdx  with one FMA-pipe instruction and keeps a
//   running (best, second) pair with 3 integer min/max per element.  The epilogue's min/max instructions, not the
//   tensor pipe, bound the kernel.
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
constexpr int TC_MAX_TRAIN = 32768;  // trainIdx field of the epilogue key (15 bits: the key stays below 2^24, exact in FP32)
constexpr float kKeyScale = 16384.f; // key = (128 - S/2) * 32768 + idx = 4194304 - 16384 S + idx

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
// instruction descriptor: A and B K-major, M = 128, N = BN;  kind::i8: S8 x S8 -> S32;  kind::f8f6f4: E4M3 x E4M3 -> F32
constexpr uint32_t kInstrDescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kInstrDescF8 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

template <bool I8>
__device__ __forceinline__ void umma_8bit(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    if (I8)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDescI8), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDescF8), "r"(accumulate) : "memory");
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

// accumulator -> key.  FP32 accumulators: the key is an integer-valued float below 2^24, and non-negative floats order
// like their bit patterns, so the running pair is kept with integer min/max on the bits (no conversion in the loop).
template <bool I8>
__device__ __forceinline__ uint32_t make_key(uint32_t acc, int base_i, float base_f, int j)
{
    if (I8) return (uint32_t)((base_i + j) - (int)acc * (int)kKeyScale);
    return __float_as_uint(fmaf(__uint_as_float(acc), -kKeyScale, base_f + (float)j));
}

template <bool I8>
__device__ __forceinline__ uint32_t export_key(uint32_t k)
{   // epilogue key -> the (distance << kIdxBits | trainIdx) key of match_finalize_kernel
    if (k == kKeyNone) return kKeyNone;
    const uint32_t ki = I8 ? k : (uint32_t)__uint_as_float(k);
    return ((ki >> 15) << kIdxBits) | (ki & 32767u);
}

// ------------------------------------------------------------------------------------------ bits -> +-1 bytes
template <bool I8>
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
        o[nib] = I8 ? (0xFFFFFFFFu ^ (spread * 0xFEu)) : (0xB8B8B8B8u ^ (spread << 7));    // 1 -> +1, 0 -> -1
    }
    uint4 *dst = out + (word_begin + i) * 2;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ------------------------------------------------------------------------------------------ GEMM + running top-2
template <bool I8, int STAGES>
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

    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sB = sA + KSLABS * BM * SLAB;
    uint64_t *bars = (uint64_t *)(sB + STAGES * KSLABS * BN * SLAB);
    uint64_t *barA = bars, *full = bars + 1, *empty = full + STAGES, *tfull = empty + STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(barA, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
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
                        umma_8bit<I8>(d_tmem, make_smem_desc(a_addr + k * 32), make_smem_desc(b_addr + k * 32), (ks | k) ? 1u : 0u);
                }
                tcgen05_commit(empty + s);     // train stage reusable once these MMAs retire
                tcgen05_commit(tfull + acc);   // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: thread <-> (query row, column half) =====
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // which 64 columns of every 128-column tile
        const int q = q0 + quarter * 32 + lane;          // row within the tile == TMEM lane
        // two independent running pairs (even / odd columns) halve the dependent min/max chain; merged at the end
        uint32_t e1 = kKeyNone, e2 = kKeyNone, o1 = kKeyNone, o2 = kKeyNone;
        for (int i = 0; i < n_tiles; ++i) {
            const int acc = i & 1;
            const int col0 = i * BN + half * (BN / 2);
            mbar_wait(tfull + acc, (i >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
            const bool ragged = col0 + BN / 2 > nt;      // warp-uniform: only a frame's last tile
#pragma unroll 1
            for (int c0 = 0; c0 < BN / 2; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                const int base_i = 4194304 + col0 + c0;
                const float base_f = (float)base_i;
                if (!ragged) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        top2(e1, e2, make_key<I8>(v[j], base_i, base_f, j));
                        top2(o1, o2, make_key<I8>(v[j + 1], base_i, base_f, j + 1));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t key = make_key<I8>(v[j], base_i, base_f, j);
                        top2(e1, e2, (col0 + c0 + j < nt) ? key : kKeyNone);   // rows of the next frame / zero fill
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tempty + acc);
        }
        top2(e1, e2, o1);
        top2(e1, e2, o2);
        if (q < nq) a.partial[((size_t)pair * 2 + half) * a.q_stride + q] = make_uint2(export_key<I8>(e1), export_key<I8>(e2));
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

template <bool I8>
cudaError_t launch_tc(const CUtensorMap &map, const TcKnnArgs &a, int max_nq, int n_pairs, cudaStream_t s)
{
    constexpr int STAGES = 2;   // 2 CTAs per SM (<= 113 KB each)
    const size_t smem = 1024 + (size_t)KSLABS * BM * SLAB + (size_t)STAGES * KSLABS * BN * SLAB +
                        (1 + 2 * STAGES + 4) * sizeof(uint64_t) + 16;
    auto kern = knn2_hamming_tc_kernel<I8, STAGES>;
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

bool tc_kind_i8()
{
    static int kind = -1;
    if (kind < 0) {
        const char *e = getenv("MVS_TC_KIND");      // "f8": E4M3 operands with FP32 accumulators instead of S8 / S32
        kind = (e && e[0] == 'f') ? 0 : 1;
    }
    return kind == 1;
}

void launch_expand_desc(const uint4 *desc, size_t row_begin, size_t n_rows, void *desc8, cudaStream_t s)
{
    if (!n_rows) return;
    const size_t n_words = n_rows * 8;
    const unsigned blocks = (unsigned)((n_words + 255) / 256);
    if (tc_kind_i8())
        expand_desc_kernel<true><<<blocks, 256, 0, s>>>((const uint32_t *)desc, row_begin * 8, n_words, (uint4 *)desc8);
    else
        expand_desc_kernel<false><<<blocks, 256, 0, s>>>((const uint32_t *)desc, row_begin * 8, n_words, (uint4 *)desc8);
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
    return tc_kind_i8() ? launch_tc<true>(map, a, max_nq, n_pairs, s) : launch_tc<false>(map, a, max_nq, n_pairs, s);
}

}  // namespace mvs
