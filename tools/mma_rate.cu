// Microbenchmark: cycles per tcgen05.mma (cta_group::1, both operands in shared memory, K-major SW128) as a
// function of M, N and operand kind, one CTA per SM, no TMA traffic (operands are whatever is in smem).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I michigan_b200/csrc tools/mma_rate.cu -o tools/bin/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "mg_ptx.cuh"
using namespace mg;

__device__ int g_fill_random = 0;
__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int kind, int iters, int a_stride_slots, long long* out,
                                                      int a_row_off, int a_sbo, int b_row_off, int b_sbo) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // operand contents: zeros, or (g_fill_random) pseudo-random fp16 values in (-1, 1) - tensor-core timing turns out to depend on it
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        const uint32_t lo = 0x3000u | (h & 0x8fffu & 0x8bffu), hi = 0x3000u | ((h >> 16) & 0x8bffu);
        reinterpret_cast<uint32_t*>(smem)[i] = g_fill_random ? (lo | (hi << 16)) : 0u;
    }
    if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0 && lane == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        const uint32_t idesc = kind == 0 ? umma_idesc_tf32(M, N) : umma_idesc_16(M, N, 1);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // rotate through 4 operand slots of 16 KB (A) / 32 KB (B) like a pipeline would
            const uint32_t sa = a0 + (uint32_t)((it & 3) * a_stride_slots * 16384), sb = b0 + (uint32_t)((it & 1) * 32768);
            // halo-style operands: first row shifted by *_row_off rows (128 B each), 8-row groups *_sbo bytes apart
            const uint64_t da = umma_desc_sw128_general(sa + a_row_off * 128, a_sbo, 0), db = umma_desc_sw128_general(sb + b_row_off * 128, b_sbo, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (kind == 0) umma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                else umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Two issuing threads (lane 0 of warps 0 and 2), each with its own TMEM accumulator and operand slots: does the per-MMA
// cost of small-N shapes come from the tensor pipe or from the single issuing thread?
__global__ void __launch_bounds__(128, 1) rate2_kernel(int M, int N, int kind, int iters, int issuers, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    __shared__ long long t_end[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x30003000u;
    if (warp == 0 && lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const int who = warp >> 1;   // warp 0 -> issuer 0, warp 2 -> issuer 1
    const long long t0 = clock64();
    if ((warp == 0 || warp == 2) && lane == 0 && who < issuers) {
        const uint32_t a0 = smem_u32(smem) + who * 32768, b0 = smem_u32(smem) + 65536 + who * 49152;
        const uint32_t idesc = kind == 0 ? umma_idesc_tf32(M, N) : umma_idesc_16(M, N, 1);
        for (int it = 0; it < iters / issuers; ++it) {
            const uint64_t da = umma_desc_kmajor_sw128(a0 + (uint32_t)((it & 1) * 16384)), db = umma_desc_kmajor_sw128(b0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (kind == 0) umma_tf32(tmem + who * 256, da + 2 * k, db + 2 * k, idesc, 1u);
                else umma_f16(tmem + who * 256, da + 2 * k, db + 2 * k, idesc, 1u);
            }
        }
        umma_commit(&bar[who]);
        mbar_wait(&bar[who], 0);
        t_end[who] = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (issuers == 2 ? (t_end[0] > t_end[1] ? t_end[0] : t_end[1]) : t_end[0]) - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int fill = 0; fill < 2; ++fill) {
    cudaMemcpyToSymbol(g_fill_random, &fill, sizeof(int));
    printf("\n== operand contents: %s\n", fill ? "pseudo-random fp16" : "zeros");
    printf("kind   M    N   cycles/MMA   (ideal N/2 for M=128 f16 K16; tf32 K8)\n");
    for (int kind = 0; kind < 2; ++kind)
        for (int M : {128, 64})
            for (int N : {256, 128, 64, 32}) {
                if (M == 64 && N > 256) continue;
                rate_kernel<<<148, 128, 180 * 1024>>>(M, N, kind, iters, 1, d, 0, 1024, 0, 1024);
                cudaError_t e = cudaDeviceSynchronize();
                long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("%-5s %4d %4d   %8.1f   %s\n", kind == 0 ? "tf32" : "f16", M, N, (double)c / (iters * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    }
    cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("\none vs two issuing threads (separate accumulators), total MMAs fixed: kind N issuers -> cycles per MMA (aggregate)\n");
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {256, 128, 64})
            for (int iss = 1; iss <= 2; ++iss) {
                rate2_kernel<<<148, 128, 180 * 1024>>>(128, N, kind, iters, iss, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("%-5s N=%3d issuers=%d   %8.1f   %s\n", kind == 0 ? "tf32" : "f16", N, iss, (double)c / (iters * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    printf("\nshifted / strided operands (f16, M=128): a_row_off a_sbo b_row_off b_sbo N -> cycles/MMA\n");
    const int cfg[][4] = {{0, 1024, 0, 1024}, {1, 1024, 0, 1024}, {0, 1280, 0, 1024}, {1, 1280, 0, 1024}, {0, 2048, 0, 1024}, {1, 2048, 0, 1024},
                          {0, 1024, 1, 1024}, {0, 1024, 0, 1280}, {0, 1024, 1, 1280}, {0, 1024, 1, 2048}, {0, 1024, 11, 1280}};
    for (auto& c4 : cfg)
        for (int N : {256, 128, 64}) {
            if (c4[3] != 1024 && N * c4[3] / 8 + 4096 > 64 * 1024) continue;   // stay inside the smem window
            rate_kernel<<<148, 128, 180 * 1024>>>(128, N, 1, iters, 1, d, c4[0], c4[1], c4[2], c4[3]);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("%2d %5d %2d %5d  N=%3d   %8.1f   %s\n", c4[0], c4[1], c4[2], c4[3], N, (double)c / (iters * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
