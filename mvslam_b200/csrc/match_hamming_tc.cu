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
//   The contraction runs as tcgen05.mma (M = N = 128, K = 8 x 32) with both operands staged by TMA (128-byte swizzle)
//   and the accumulators double-buffered in TMEM.  The epilogue never materialises the score matrix: thread <-> (query
//   row, column half) keeps the running maximum of four column STREAMS (columns c' = 0, 1, 2, 3 mod 4 of its 64-column
//   half) on packed 16-bit keys -- 0.25 three-input VIMNMX per accumulator -- and hands K2 the exact best plus the best of
//   all OTHER streams.  That second value is an upper bound of the true second distance: the true second neighbour is
//   either it or one of the 15 stream-mates of the best (same 64-column block, same index mod 4), which K2 evaluates
//   with popcounts for the few queries whose fate depends on it (match_hamming.cu, refine_second).  A full (best, second)
//   chain costs 1.25 ALU instructions per accumulator and kept the ALU pipe, not the tensor pipe, on the critical path
//   (round 2 measurement).
//
// Warp roles (640 threads, one persistent CTA per SM):
//   warp 0      train-tile producer (128-row tiles through a STAGES-deep mbarrier ring, running ahead across items)
//   warp 1, 18  tcgen05.mma issuers, one per 128-row block of the query tile (converged warp, elected lane); warp 1 also
//               allocates the TMEM
//   warps 2-17  epilogue: two groups of 8 (one per 128-row block of the query tile)
//   warp 19     query row-block producer (3-slot ring of 128-row blocks)
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
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
constexpr int RB = 2;                // 128-row blocks of a query tile (each train tile is multiplied with both)
constexpr int EPI_GROUP = 8;         // warps per epilogue group: two per TMEM lane quarter, each takes half of a tile's columns
constexpr int EPI_WARPS = RB * EPI_GROUP;  // one group per row block
constexpr int MMA_WARP_B = 2 + EPI_WARPS;  // second issuer (row block 1)
constexpr int A_WARP = MMA_WARP_B + 1;     // query row-block producer (ring of A_SLOTS 128-row blocks, a kernel template parameter)
constexpr int TC_THREADS = (A_WARP + 1) * 32;
constexpr int NACC = 2 * RB;         // TMEM accumulator buffers: double-buffered per row block
constexpr uint32_t TMEM_COLS = NACC * BN;
constexpr int TC_SPLITS = 2;         // partial results per query and train split: one per column half
constexpr int TC_MAX_TRAIN = (int)kIdxMask;  // the thread's running key is already the export format (hamming << 22 | trainIdx)

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
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand in a dense [rows][128 B] slab with 128-byte swizzle: start address
// >> 4 in bits 0-13, LBO (unused for swizzled K-major) = 1 in bits 16-29, SBO = 1024 B (8 rows x 128 B) >> 4 in bits 32-45,
// descriptor version 1 (Blackwell) in bit 46, layout type 2 = SWIZZLE_128B in bits 61-63.  Built as (desc_lo, kDescHi).
// instruction descriptor: kind::i8, S8 x S8 -> S32, A and B K-major, M = 128, N = BN
constexpr uint32_t kInstrDescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// The MMA warp runs converged and only predicates the tensor-core instructions on its elected lane: the descriptors
// stay warp-uniform values (uniform datapath, no per-instruction register -> uniform-register moves), which keeps the
// issue cost of one tcgen05.mma well below the 64 clocks it occupies the tensor pipe.
constexpr uint32_t kDescHi = 0x40004040u;   // high word: SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }

template <bool ACCUMULATE>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t leader)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 e, %4, 0;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(kInstrDescI8), "r"(leader), "n"(ACCUMULATE ? 1 : 0), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_if(uint64_t *bar, uint32_t leader)
{
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
    return pred;
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 64 accumulator columns as 32 registers: .pack::16b keeps the low 16 bits of every 32-bit column and puts two adjacent
// columns into one register (even column low, odd column high) -- the accumulators are signed 16-bit keys by construction
__device__ __forceinline__ void tmem_ld64_packed(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
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
// Epilogue arithmetic.  With every bit stored as +-8 the contraction gives 64 S = 128 (S/2), a multiple of 128 that fits
// a signed 16-bit lane.  tcgen05.ld.pack::16b delivers two adjacent accumulators per register (their low 16 bits); one
// integer multiply-add on the otherwise idle FMA pipe (x * 1 + constant, the 1 being a kernel argument so that it stays
// an IMAD) puts  63 - c'  (c' = column inside the thread's 64-column half) into the free low 7 bits of both lanes:
//      k16 = 128 (128 - hamming) + (63 - c')          in [-16384, 16447]
// orders the columns of the half by (smaller distance, then smaller index) under MAX; one three-input VIMNMX.S16x2 folds
// two registers (four accumulators) into the running maxima of their streams.  At the end of a tile the two best of the four
// stream maxima are widened to  hamming << 22 | trainIdx  and merged into the thread's 32-bit pair (minimum = best).
// (Measured alternative: a ninth K step over a constant slab that adds the index inside the MMA -- no epilogue
// instruction at all, but 12.5 % more tensor work on a kernel whose tensor pipe is 88 % busy.)
__device__ __forceinline__ uint32_t widen_key(uint32_t k16, uint32_t tile_base /* first column of the thread's half */)
{
    const int k = (int)(short)k16;                         // empty lane: -32768 -> distance 384, dropped at the export
    return ((uint32_t)(128 - (k >> 7)) << kIdxBits) + tile_base + (63u - ((uint32_t)k & 127u));
}

// Persistent kernel: one CTA per SM walks the (pair, 256-row query tile) items of the batch with a stride of gridDim.x.
// A 128-row train tile is loaded once and multiplied with BOTH 128-row blocks of the query tile: the L2 -> shared
// traffic per descriptor pair is 1 byte instead of 2 (the L2 slices cap at ~6300 B/clk chip-wide, which a 128-row
// query tile saturates at half the tensor rate).  The three roles keep running counters, so the TMA ring (STAGES train
// tiles in flight) and the four TMEM accumulators (2 row blocks x 2 steps) stay full across item boundaries.
struct ItemInfo { int pair, q0, nq, nt, row_q, row_t, n_tiles, n_rb, tile0, ts; };

#ifdef MVS_TC_PROBE   // experiment build only (tools/knn_probe.py): per-CTA clock and wall-time counters of the last launch
__device__ unsigned long long g_tc_probe[160][10];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PROBE_T0() const long long pt0__ = clock64()
#define PROBE_ADD(var) var += clock64() - pt0__
#else
#define PROBE_T0()
#define PROBE_ADD(var)
#endif

template <bool TSPLIT>
__device__ __forceinline__ bool load_item(const TcKnnArgs &a, int item, ItemInfo &it, int n_items)
{
    if (item >= n_items) return false;
    const int t_splits = TSPLIT ? a.t_splits : 1;       // compile-time 1 for the large batches: no division beyond item / q_tiles
    // item = (pair, query tile, train split): launches with fewer items than SMs cut the train dimension as well (t_splits,
    // chosen by the launcher), so that one VO pair or one large pair of descriptor sets still spreads over the chip
    const int per_pair = a.q_tiles * t_splits;
    it.pair = item / per_pair;
    const int rem = item - it.pair * per_pair;
    const int qt = rem / t_splits;
    it.ts = rem - qt * t_splits;
    it.q0 = qt * (RB * BM);
    int fq, ft;
    if (a.pairs) {  // query = pair frame (second), train = base frame (first): visual-feature.cpp:59-60
        const int2 pr = a.pairs[it.pair];
        fq = a.reverse ? pr.x : pr.y;
        ft = a.reverse ? pr.y : pr.x;
    } else { fq = a.reverse ? 0 : 1; ft = a.reverse ? 1 : 0; }
    it.nq = a.frame_cnt[fq]; it.nt = a.frame_cnt[ft];
    it.row_q = a.frame_off[fq] + it.q0; it.row_t = a.frame_off[ft];
    const int tiles_total = (it.nt + BN - 1) / BN, tiles_per = (tiles_total + t_splits - 1) / t_splits;
    it.tile0 = it.ts * tiles_per;                           // a trailing split of a short frame may own no tile: it exports "none"
    it.n_tiles = max(0, min(tiles_per, tiles_total - it.tile0));
    it.n_rb = (it.q0 + BM < it.nq) ? 2 : 1;                 // the second row block may be empty (every role skips it)
    return it.q0 < it.nq;                                   // every role skips the same items
}

// Pipeline (measured on the B200 box with clock64 counters in every role, round 2; clocks per 1024-clock train tile):
//   * the tensor pipe ran 1245 clocks per tile with a single producer lane that loaded "query tile, then that item's train
//     tiles" and waited for the previous item's MMAs before touching the query buffer: both rings ran dry at every item
//     boundary, and a.pairs / frame_cnt / frame_off were fetched on the critical path;
//   * with the boundaries gone the period was still ~1250: four epilogue warps per SM sub-partition shared an ALU pipe that
//     their (best, second) chains loaded to 82 %, so accumulators came back late.  Without any epilogue arithmetic the same
//     kernel ran 1060 clocks per tile.
// What is left (this revision, same box, 1024 Tsukuba pairs, CUDA-event time of the launch): 0.445 ms; without the TMA
// loads 0.441, additionally without the epilogue arithmetic 0.431, additionally without the TMEM loads 0.405 (the tcgen05.ld
// traffic shares the TMEM read port with the accumulating MMAs: 6 %); issuing the same MMAs back to back in a micro-benchmark
// (tools/ubench_umma.cu: 64.00 clocks per instruction in SS mode, i.e. the operands are NOT shared-memory-bandwidth bound)
// would take 0.37 ms, and ~0.01 ms is launch + TMEM allocation.
// Hence: the train ring (warp 0) never waits for anything but its own stages and runs across item boundaries; the query row
// blocks go through their own 3-slot ring (warp A_WARP), so the next item's first row block is resident before the current
// item ends and the second follows while the first is being multiplied; each row block has its own issuing warp; item
// metadata is fetched one item ahead; the epilogue keeps stream maxima only (see the file header).
template <int STAGES, int A_SLOTS, bool TSPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn2_hamming_tc_kernel(const __grid_constant__ CUtensorMap map, TcKnnArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the launch requests no static shared memory, so the dynamic window starts 1024-byte aligned (checked)
    uint8_t *sA = smem_raw;                                 // [A_SLOTS][KSLABS] query slabs, one 128-row block per slot
    uint8_t *sB = sA + A_SLOTS * KSLABS * BM * SLAB;        // [STAGES][KSLABS] train slabs
    uint64_t *bars = (uint64_t *)(sB + STAGES * KSLABS * BN * SLAB);
    uint64_t *afull = bars, *aempty = afull + A_SLOTS, *full = aempty + A_SLOTS, *empty = full + STAGES, *tfull = empty + STAGES,
             *tempty = tfull + NACC;
    uint32_t *tmem_slot = (uint32_t *)(tempty + NACC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.q_tiles * (TSPLIT ? a.t_splits : 1) * a.n_pairs;
    pdl_launch_dependents();   // K2's blocks may take the SMs this grid leaves free (small batches); they wait for this grid's end
#ifdef MVS_TC_PROBE
    long long w_a = 0, w_full = 0, w_acc = 0;
    const long long pk0 = clock64();
    const unsigned long long pg0 = gtimer();
#endif

    if (threadIdx.x == 0) {
        if (smem_u32(smem_raw) & 1023u) __trap();
        for (int s = 0; s < A_SLOTS; ++s) { mbar_init(afull + s, 1); mbar_init(aempty + s, 1); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, RB); }
        for (int b = 0; b < NACC; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, EPI_GROUP); }
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

    // mbarrier parity convention: a consumer waits for fill number n with parity n & 1; a producer waits for the n-th
    // release with parity (n & 1) ^ 1, which passes at once for n = 0 (nothing to wait for on first use).
    // Query row blocks are numbered per CTA in the order (item, row block); number u lives in slot u % A_SLOTS.
    // Every role walks the same item sequence and fetches the NEXT item's metadata while it works on the current one.
    ItemInfo it, nx;
    bool have = load_item<TSPLIT>(a, blockIdx.x, nx, n_items);
    if (warp == 0) {
        // ===== train-tile producer: runs ahead across item boundaries, bounded only by its own ring =====
        if (lane == 0) {
            uint32_t tile_no = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                it = nx;
                const bool valid = have;
                have = load_item<TSPLIT>(a, item + (int)gridDim.x, nx, n_items);
                if (!valid) continue;
                for (int i = 0; i < it.n_tiles; ++i, ++tile_no) {
                    const uint32_t s = tile_no % STAGES;
                    mbar_wait(empty + s, ((tile_no / STAGES) & 1) ^ 1);
                    mbar_expect_tx(full + s, KSLABS * BN * SLAB);
                    for (int ks = 0; ks < KSLABS; ++ks)
                        tma_load_2d(&map, full + s, sB + (s * KSLABS + ks) * BN * SLAB, ks * SLAB, it.row_t + (it.tile0 + i) * BN);
                }
            }
        }
    } else if (warp == A_WARP) {
        // ===== query row-block producer =====
        if (lane == 0) {
            uint32_t u = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                it = nx;
                const bool valid = have;
                have = load_item<TSPLIT>(a, item + (int)gridDim.x, nx, n_items);
                if (!valid) continue;
                for (int rb = 0; rb < it.n_rb; ++rb, ++u) {
                    const uint32_t sl = u % A_SLOTS;
                    mbar_wait(aempty + sl, ((u / A_SLOTS) & 1) ^ 1);
                    mbar_expect_tx(afull + sl, KSLABS * BM * SLAB);
                    for (int ks = 0; ks < KSLABS; ++ks)
                        tma_load_2d(&map, afull + sl, sA + (sl * KSLABS + ks) * BM * SLAB, ks * SLAB, it.row_q + rb * BM);
                }
            }
        }
    } else if (warp == 1 || warp == MMA_WARP_B) {
        // ===== MMA issuers (converged warps, elected lane issues): warp 1 owns row block 0, warp MMA_WARP_B row block 1 =====
        const int rb = (warp == 1) ? 0 : 1;
        const uint32_t leader = elect_one();
        const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
        uint32_t tile_no = 0, use_no = 0, u = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            it = nx;
            const bool valid = have;
            have = load_item<TSPLIT>(a, item + (int)gridDim.x, nx, n_items);
            if (!valid) continue;
            const bool active = rb < it.n_rb;             // an item may have one row block only: stay in step with the train ring
            const uint32_t sl = (u + rb) % A_SLOTS;
            { PROBE_T0(); if (active) mbar_wait(afull + sl, ((u + rb) / A_SLOTS) & 1); PROBE_ADD(w_a); }
            const uint32_t a_lo = a_lo0 + sl * (KSLABS * BM * SLAB >> 4);
            for (int i = 0; i < it.n_tiles; ++i, ++tile_no) {
                const uint32_t s = tile_no % STAGES;
                const uint32_t b_lo = b_lo0 + s * (KSLABS * BN * SLAB >> 4);
                { PROBE_T0(); mbar_wait(full + s, (tile_no / STAGES) & 1); PROBE_ADD(w_full); }
                if (active) {
                    const uint32_t acc = (use_no & 1) * RB + rb;
                    { PROBE_T0(); mbar_wait(tempty + acc, ((use_no >> 1) & 1) ^ 1); PROBE_ADD(w_acc); }   // epilogue drained this accumulator
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    // UMMA_K = 32 one-byte elements = 32 bytes inside the swizzle atom; 4 steps per 128-byte slab
                    umma_i8<false>(d_tmem, a_lo, b_lo, leader);
#pragma unroll
                    for (int k = 1; k < 4 * KSLABS; ++k)
                        umma_i8<true>(d_tmem, a_lo + (k >> 2) * (BM * SLAB >> 4) + (k & 3) * 2,
                                      b_lo + (k >> 2) * (BN * SLAB >> 4) + (k & 3) * 2, leader);
                    tcgen05_commit_if(tfull + acc, leader);      // accumulator ready for the epilogue
                    ++use_no;
                }
                tcgen05_commit_if(empty + s, leader);            // train stage reusable once both issuers' MMAs retire
            }
            if (active) tcgen05_commit_if(aempty + sl, leader);  // this row block's slot is free once its MMAs retire
            u += (uint32_t)it.n_rb;
        }
    } else {
        // ===== epilogue: thread <-> (query row, column half), warp group <-> row block =====
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access (warps 2..17)
        const int ew = warp - 2;
        const int half = (ew >> 2) & 1;                  // which 64 columns of every 128-column tile
        const int rb = ew >> 3;                          // row block of the query tile
        uint32_t use_no = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            it = nx;
            const bool valid = have;
            have = load_item<TSPLIT>(a, item + (int)gridDim.x, nx, n_items);
            if (!valid) continue;
            if (rb >= it.n_rb) continue;
            const int q = it.q0 + rb * BM + quarter * 32 + lane;   // row within the block == TMEM lane
            uint32_t g1 = kKeyNone, g2 = kKeyNone;
            for (int i = 0; i < it.n_tiles; ++i, ++use_no) {
                const uint32_t acc = (use_no & 1) * RB + rb;
                const int col0 = (it.tile0 + i) * BN + half * (BN / 2);
                { PROBE_T0(); mbar_wait(tfull + acc, (use_no >> 1) & 1); PROBE_ADD(w_acc); }
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
                const bool ragged = col0 + BN / 2 > it.nt;   // warp-uniform: only a frame's last tile
                // one packed load brings the 64 columns as 32 registers; the accumulator goes back to the MMA warp as soon as
                // they have arrived (one arrival per warp)
                uint32_t v[32];
                tmem_ld64_packed(taddr, v);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + acc);
                const uint32_t one = (uint32_t)a.one;    // opaque 1: keeps the index insertion an IMAD (FMA pipe)
#define MVS_KEY(m) (v[m] * one + (uint32_t)(((63 - (2 * (m) + 1)) << 16) | (63 - 2 * (m))))
                // register m holds columns 2m (low lane) and 2m + 1 (high lane): the even registers feed the streams
                // c' = 0, 1 (mod 4), the odd ones c' = 2, 3 (mod 4)
                uint32_t sa = 0x80008000u, sb = 0x80008000u;
                if (!ragged) {
#pragma unroll
                    for (int m = 0; m < 32; m += 4) {
                        sa = __vimax3_s16x2(sa, MVS_KEY(m), MVS_KEY(m + 2));
                        sb = __vimax3_s16x2(sb, MVS_KEY(m + 1), MVS_KEY(m + 3));
                    }
                } else {                                 // rows of the next frame / zero fill lose to every real key
#pragma unroll
                    for (int m = 0; m < 32; ++m) {
                        const int c = col0 + 2 * m;
                        const uint32_t key = MVS_KEY(m);
                        const uint32_t pk = (c < it.nt ? (key & 0xFFFFu) : 0x8000u) | (c + 1 < it.nt ? (key & 0xFFFF0000u) : 0x80000000u);
                        if (m & 1) sb = __vmaxs2(sb, pk); else sa = __vmaxs2(sa, pk);
                    }
                }
#undef MVS_KEY
                // the two best of the four stream maxima (in both 16-bit lanes of m1 / m2) -> the thread's running pair
                const uint32_t n1 = __vmaxs2(sa, sb), n2 = __vmins2(sa, sb);
                const uint32_t r1 = __byte_perm(n1, 0, 0x1032), r2 = __byte_perm(n2, 0, 0x1032);
                const uint32_t m1 = __vmaxs2(n1, r1);
                const uint32_t m2 = __vmaxs2(__vmins2(n1, r1), __vmaxs2(n2, r2));
                const uint32_t tile_base = (uint32_t)col0;
                top2(g1, g2, widen_key(m1 & 0xFFFFu, tile_base));
                top2(g1, g2, widen_key(m2 & 0xFFFFu, tile_base));
            }
            if (q < it.nq) {
                const uint32_t x1 = (g1 >> kIdxBits) > 256u ? kKeyNone : g1;
                const uint32_t x2 = (g2 >> kIdxBits) > 256u ? kKeyNone : g2;
                a.partial[((size_t)it.pair * (TC_SPLITS * (TSPLIT ? a.t_splits : 1)) + it.ts * TC_SPLITS + half) * a.q_stride + q] = make_uint2(x1, x2);
            }
        }
    }
#ifdef MVS_TC_PROBE
    // slots: 0 issuer A total clocks, 1-3 issuer A waits (query block, train tile, accumulator), 4-6 issuer B waits,
    //        7 CTA wall time (ns), 8 one epilogue warp's wait for accumulators
    if (lane == 0 && warp == 1) {
        unsigned long long *o = g_tc_probe[blockIdx.x];
        o[0] = (unsigned long long)(clock64() - pk0); o[1] = (unsigned long long)w_a; o[2] = (unsigned long long)w_full; o[3] = (unsigned long long)w_acc;
    }
    if (lane == 0 && warp == MMA_WARP_B) {
        unsigned long long *o = g_tc_probe[blockIdx.x];
        o[4] = (unsigned long long)w_a; o[5] = (unsigned long long)w_full; o[6] = (unsigned long long)w_acc;
    }
    if (lane == 0 && warp == 2) g_tc_probe[blockIdx.x][8] = (unsigned long long)w_acc;
#endif
    tcgen05_fence_before();
    __syncthreads();
#ifdef MVS_TC_PROBE
    if (threadIdx.x == 0) g_tc_probe[blockIdx.x][7] = gtimer() - pg0;
#endif
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

cudaError_t launch_tc(const CUtensorMap &map, TcKnnArgs a, int max_nq, int n_pairs, cudaStream_t s)
{
    // one persistent CTA per SM: 3 x 32 KB query row blocks + 4 x 32 KB train ring = 224 KB.  (4 query slots + a 3-deep train
    // ring, which lets BOTH row blocks of the next item load early, measured the same: 0.439 against 0.440 ms per 1024 pairs.)
    constexpr int STAGES = 4, A_SLOTS = 3;
    const size_t smem = (size_t)A_SLOTS * KSLABS * BM * SLAB + (size_t)STAGES * KSLABS * BN * SLAB +
                        (2 * A_SLOTS + 2 * STAGES + 2 * NACC) * sizeof(uint64_t) + 16;
    auto kern = knn2_hamming_tc_kernel<STAGES, A_SLOTS, false>;      // large batches: the item decode has one division
    auto kern_split = knn2_hamming_tc_kernel<STAGES, A_SLOTS, true>; // small launches cut the train dimension too
    static int sm_count[64] = {0};          // per device: the shared-memory opt-in is a per-device function attribute
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64) return e != cudaSuccess ? e : cudaErrorInvalidDevice;
    if (!sm_count[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        sm_count[dev] = n;
    }
    const int n_sm = sm_count[dev];
    a.q_tiles = (max_nq + RB * BM - 1) / (RB * BM);
    a.n_pairs = n_pairs;
    a.one = 1;
    if (a.t_splits < 1) a.t_splits = 1;
    const long items = (long)a.q_tiles * a.t_splits * n_pairs;
    if (a.t_splits > 1) kern_split<<<(unsigned)std::min<long>(items, n_sm), TC_THREADS, smem, s>>>(map, a);
    else kern<<<(unsigned)std::min<long>(items, n_sm), TC_THREADS, smem, s>>>(map, a);
    return cudaGetLastError();
}

}  // namespace

#ifdef MVS_TC_PROBE
extern "C" void mvs_debug_tc_probe_dump(unsigned long long *out /* [160][10] */)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_tc_probe, sizeof(unsigned long long) * 160 * 10);
}
#endif

int tc_max_train() { return TC_MAX_TRAIN; }
// Partial top-2 pairs per query the kernel writes for a launch of this shape: two column halves x the train splits.  The train
// dimension is cut only when the (pair, query tile) items would leave most SMs idle -- a single VO pair is 7 items -- and never
// finer than four train tiles per split (each split pays the pipeline fill and one more partial for K2 to merge).
int tc_train_splits(int max_nq, int max_nt, int n_pairs)
{
    int dev = 0, n_sm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) { int n = 0; if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) n_sm = n; }
    const long items = (long)((max_nq + RB * BM - 1) / (RB * BM)) * n_pairs;
    const int tiles = (max_nt + BN - 1) / BN;
    if (items * 2 > n_sm || tiles < 8) return 1;
    return (int)std::max<long>(1, std::min<long>(std::min<long>(n_sm / items, tiles / 4), 16));
}
int tc_splits(int max_nq, int max_nt, int n_pairs) { return TC_SPLITS * tc_train_splits(max_nq, max_nt, n_pairs); }

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
