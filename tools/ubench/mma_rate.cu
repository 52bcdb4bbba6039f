// Microbenchmark: cycles per tcgen05.mma (kind::f16, K = 16) for the small tiles of the recurrent kernels, operands
// already in place: A from shared memory (SS) or tensor memory (TS), M = 64 / 128, N = 16 .. 128.  52 MMAs back to
// back (the K = 832 of one chain step), one commit, wait.  Prints issue time and time to completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../asr_b200/csrc -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "ptx.cuh"
using namespace asrb;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode bit 0: TS (A in tensor memory); M, N as given; nmma MMAs; B (and SS A) operands walk through K blocks of 64
template <int ts, int M, int N, int ND, int DISTINCT>
__global__ void __launch_bounds__(640, 1) rate_kernel(int nmma, int reps, long long* out, int gap = 0, int warm = 0) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    const bool rnd = (nmma & 256) != 0;
    nmma &= 255;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // two bf16 in about [-1, 1): sign + exponent 0x3f0..0x3f7 + random mantissa
        reinterpret_cast<uint32_t*>(smem)[i] = rnd ? ((h & 0x807f807fu) | 0x3f003f00u) : 0x3c003c00u;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base;
    if (rnd && warp < 4) {      // random weights in every lane / column of tensor memory
        for (int c0 = 0; c0 < 512; c0 += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) { uint32_t h = (uint32_t)(threadIdx.x * 512 + c0 + j) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; r[j] = (h & 0x807f807fu) | 0x3f003f00u; }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(tm + ((uint32_t)(warp * 32) << 16) + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before_sync();
    }
    __syncthreads();
    tc_fence_after_sync();
    constexpr uint32_t idesc = umma_idesc(kFmtBF16, M, N);
    // A (SS): K blocks of [M rows x 128 B] from smem + 0; B: K blocks of [N rows x 128 B] from smem + 96 KB
    constexpr uint32_t a_blk = (uint32_t)M * 128, b_blk = (uint32_t)N * 128;
    uint8_t* sa = smem;
    uint8_t* sb = smem + (DISTINCT ? 0 : 112 * 1024);
    long long t_issue = 0, t_done = 0;
    uint32_t phase = 0;
    if (warp == 1) {
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            if (elect_one()) {
            const uint64_t b0 = umma_desc_sw128(smem_u32(sb)), a0 = umma_desc_sw128(smem_u32(sa));
            constexpr uint64_t bstep = b_blk >> 4, astep = a_blk >> 4;
#pragma unroll
            for (int i = 0; i < 52; ++i) {   // (constant indices: the issue loop is two uniform-datapath instructions per MMA)
                const int kb = i >> 2, k = i & 3;
                const uint64_t bdesc = b0 + (uint64_t)(DISTINCT ? kb : kb % 4) * bstep + 2 * k;
                if (ts) mma_ts(DISTINCT ? tm : tm + 256 + (i % ND) * N, DISTINCT ? tm + 64 + i * 8 : tm + (uint32_t)(i * 8 % 256), bdesc, idesc, i >= ND);
                else umma_f16(tm + 256 + (i % ND) * N, a0 + (uint64_t)(kb % 6) * astep + 2 * k, bdesc, idesc, i >= ND);
            }
            umma_commit(&bar);
            }
            __syncwarp();
            const long long t1 = clock64();
            mbar_wait(&bar, phase);
            phase ^= 1;
            const long long t2 = clock64();
            if (r >= 2) { t_issue += t1 - t0; t_done += t2 - t0; }
            if (gap) {       // idle gap between bursts; with `warm`, one dummy MMA (own accumulator) every `warm` cycles meanwhile
                const long long g0 = clock64();
                long long last = g0;
                while (clock64() - g0 < gap) {
                    if (warm && clock64() - last >= warm) {
                        last = clock64();
                        if (elect_one()) mma_ts(tm + 480, tm + 64, umma_desc_sw128(smem_u32(sb)), idesc, 0);
                        __syncwarp();
                    }
                }
                if (warm) {      // drain the dummies before the timed burst
                    if (elect_one()) umma_commit(&bar);
                    __syncwarp();
                    mbar_wait(&bar, phase);
                    phase ^= 1;
                }
            }
            if (threadIdx.x == 32) mbar_arrive(&bar2);     // releases the waiting warps (nmma bit 8: 16 more warps parked in try_wait)
        }
        if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t_issue / (reps - 2); out[1] = t_done / (reps - 2); }
    }
    if (warp >= 4) { uint32_t ph2 = 0; for (int r = 0; r < reps; ++r) { mbar_wait(&bar2, ph2); ph2 ^= 1; } }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int ts, int M, int N, int ND, int DISTINCT = 0>
static void run(long long* out, int threads = 128, int grid = 1, int rnd = 0, int gap = 0, int warm = 0) {
    long long h[2];
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel<ts, M, N, ND, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel<ts, M, N, ND, DISTINCT><<<grid, threads, smem>>>(52 | (rnd ? 256 : 0), 40, out, gap, warm);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%s%s M=%3d N=%3d accumulators=%d threads=%d grid=%d %s gap=%d warm=%d: 52 MMAs issue %lld cycles (%.1f each), done %lld cycles (%.1f each)\n", ts ? "TS" : "SS", DISTINCT ? " distinct operands" : "", M, N, ND, threads, grid, rnd ? "random data" : "constant data", gap, warm,
           h[0], (double)h[0] / 52, h[1], (double)h[1] / 52);
}

int main() {
    long long* out; cudaMalloc(&out, 16);
    for (int gap : {0, 500, 1000, 2000, 5000, 20000}) run<1, 64, 32, 1, 1>(out, 640, 1, 1, gap, 0);
    for (int warm : {2000, 1000, 500, 200, 100}) run<1, 64, 32, 1, 1>(out, 640, 1, 1, 5000, warm);
    run<1, 64, 32, 1, 1>(out, 640, 100, 1, 5000, 0);
    return 0;
}
