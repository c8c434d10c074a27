// ubench.cu — measures the instruction-pipe ceilings the kernels are graded against (SURVEY.md §8d:
// "take popc rate from a micro-benchmark on the box"): POPC, LOP3, IADD3, VIMNMX, DADD, DMUL, DFMA
// per second on the whole chip, plus the mixed Hamming inner-loop ceiling.  Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8;
constexpr int ITERS = 4096;

enum Op { POPC, XOR, IADD3, MNMX, DADD, DMUL, DFMA, HAMMING, HAM_CSA3, HAM_CSA4, HAM_CSA3M, MNMX16, MNMX3, MNMX3_16, PRMT, IMAD, HMNMX2, MIX_I16_H2, FMNMX, MIX_I16_F32 };

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, double dseed)
{
    uint32_t r[ILP], acc = 0;
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { r[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; d[i] = dseed + i + threadIdx.x; }
    uint32_t q[8], b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = seed ^ (i * 0x85EBCA6Bu + threadIdx.x);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == POPC) r[i] = __popc(r[i]) + 0x55555555u;         // 1 POPC + 1 IADD per op: see note below
            if (OP == XOR) r[i] = (r[i] ^ seed) ^ (r[i] >> 1);
            if (OP == IADD3) r[i] = r[i] + seed + it;
            if (OP == MNMX) r[i] = max(min(r[i], seed + it), (uint32_t)it);
            if (OP == MNMX16) r[i] = __vmaxs2(__vmins2(r[i], seed + it), (uint32_t)it);
            if (OP == MNMX3) r[i] = max(max(r[i], seed + it), (uint32_t)it ^ seed);
            if (OP == MNMX3_16) r[i] = __vimax3_s16x2(r[i], seed + it, (uint32_t)it ^ seed);
            if (OP == PRMT) r[i] = __byte_perm(r[i], seed + it, 0x5410) ^ 0u, r[i] = __byte_perm(r[i], it, 0x1032);
            if (OP == IMAD) r[i] = r[i] * seed + it;
            if (OP == HMNMX2 || ((OP == MIX_I16_H2) && (i & 1))) {   // max.f16x2 / min.f16x2 (does it leave the ALU pipe?)
                uint32_t t;
                asm volatile("min.f16x2 %0, %1, %2;" : "=r"(t) : "r"(r[i]), "r"(seed + it));
                asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r[i]) : "r"(t), "r"((uint32_t)it));
            }
            if ((OP == MIX_I16_H2 || OP == MIX_I16_F32) && !(i & 1)) r[i] = __vmaxs2(__vmins2(r[i], seed + it), (uint32_t)it);
            if (OP == FMNMX || ((OP == MIX_I16_F32) && (i & 1))) {
                float f = __uint_as_float(r[i]);
                f = fmaxf(fminf(f, __uint_as_float(seed + it)), __uint_as_float((uint32_t)it));
                r[i] = __float_as_uint(f);
            }
            if (OP == DADD) d[i] = d[i] + dseed;
            if (OP == DMUL) d[i] = d[i] * dseed;
            if (OP == DFMA) d[i] = fma(d[i], dseed, dseed);
        }
        if (OP == HAMMING) {   // the matcher's inner loop on synthetic "train" words: 8 XOR + 8 POPC + adds + top-2
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                uint32_t dist = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) dist += __popc(q[w] ^ (r[j] + w * 0x01000193u + it));
                const uint32_t key = (dist << 22) + it * ILP + j;
                const uint32_t hi = max(b1, key);
                b1 = min(b1, key);
                b2 = min(b2, hi);
            }
        }
        if (OP == HAM_CSA3 || OP == HAM_CSA4 || OP == HAM_CSA3M) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                uint32_t x[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) x[w] = q[w] ^ (r[j] + w * 0x01000193u + it);
                // carry-save adders: (a,b,c) -> sum (weight 1), carry (weight 2); 2 LOP3 each
                const uint32_t s0 = x[0] ^ x[1] ^ x[2], c0 = (x[0] & x[1]) | (x[2] & (x[0] | x[1]));
                const uint32_t s1 = x[3] ^ x[4] ^ x[5], c1 = (x[3] & x[4]) | (x[5] & (x[3] | x[4]));
                const uint32_t s2 = s0 ^ s1 ^ x[6], c2 = (s0 & s1) | (x[6] & (s0 | s1));
                uint32_t dist;
                if (OP == HAM_CSA4) {
                    const uint32_t s3 = c0 ^ c1 ^ c2, c3 = (c0 & c1) | (c2 & (c0 | c1));
                    dist = __popc(s2) + __popc(x[7]) + 2 * __popc(s3) + 4 * __popc(c3);
                } else if (OP == HAM_CSA3) {
                    dist = __popc(s2) + __popc(x[7]) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
                } else {
                    uint32_t w2 = __popc(c0) + __popc(c1) + __popc(c2), w1 = __popc(s2) + __popc(x[7]);
                    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(dist) : "r"(w2), "r"(w1));
                }
                const uint32_t key = (dist << 22) + it * ILP + j;
                const uint32_t hi = max(b1, key);
                b1 = min(b1, key);
                b2 = min(b2, hi);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += r[i] + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + b1 + b2;
}

template <int OP>
double run(uint32_t *out, int blocks, double ops_per_iter)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<blocks, 256>>>(out, 12345u, 1.0000001);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k<OP><<<blocks, 256>>>(out, 12345u + rep, 1.0000001);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return (double)blocks * 256 * ITERS * ops_per_iter / (best * 1e-3);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8;
    uint32_t *out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d", p.name, p.multiProcessorCount, clk);
    printf(", \"popc_per_s\": %.4e", run<POPC>(out, blocks, ILP));
    printf(", \"lop3_per_s\": %.4e", run<XOR>(out, blocks, ILP * 2));
    printf(", \"iadd3_per_s\": %.4e", run<IADD3>(out, blocks, ILP));
    printf(", \"vimnmx_per_s\": %.4e", run<MNMX>(out, blocks, ILP * 2));
    printf(", \"dadd_per_s\": %.4e", run<DADD>(out, blocks, ILP));
    printf(", \"dmul_per_s\": %.4e", run<DMUL>(out, blocks, ILP));
    printf(", \"dfma_per_s\": %.4e", run<DFMA>(out, blocks, ILP));
    printf(", \"hamming256_pairs_per_s\": %.4e", run<HAMMING>(out, blocks, ILP));
    printf(", \"hamming256_csa3_pairs_per_s\": %.4e", run<HAM_CSA3>(out, blocks, ILP));
    printf(", \"hamming256_csa4_pairs_per_s\": %.4e", run<HAM_CSA4>(out, blocks, ILP));
    printf(", \"hamming256_csa3m_pairs_per_s\": %.4e", run<HAM_CSA3M>(out, blocks, ILP));
    printf(", \"vimnmx_16x2_per_s\": %.4e", run<MNMX16>(out, blocks, ILP * 2));
    printf(", \"vimnmx3_per_s\": %.4e", run<MNMX3>(out, blocks, ILP));
    printf(", \"vimnmx3_16x2_per_s\": %.4e", run<MNMX3_16>(out, blocks, ILP));
    printf(", \"prmt_per_s\": %.4e", run<PRMT>(out, blocks, ILP * 2));
    printf(", \"imad_per_s\": %.4e", run<IMAD>(out, blocks, ILP));
    printf(", \"hmnmx2_per_s\": %.4e", run<HMNMX2>(out, blocks, ILP * 2));
    printf(", \"mix_vimnmx16x2_hmnmx2_per_s\": %.4e", run<MIX_I16_H2>(out, blocks, ILP * 2));
    printf(", \"fmnmx_per_s\": %.4e", run<FMNMX>(out, blocks, ILP * 2));
    printf(", \"mix_vimnmx16x2_fmnmx_per_s\": %.4e", run<MIX_I16_F32>(out, blocks, ILP * 2));
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
