// Microbenchmark: latency of the "data is the flag" hand-over between CTAs, in isolation.
// P CTAs; every step each CTA stores one 128-byte line (its piece of the step's slot), then polls one word of every
// other CTA's line until none of them reads as "unwritten"; cycles per step = store -> all visible.
//   mode 0: fresh slot per step, filled with 0xFF before the launch (never touched before the producers write it)
//   mode 1: as 0, but every CTA prefetches its line of slot s+2 into L2 while it works on step s
//   mode 2: two slots reused alternately, validity = "word equals the step number" (lines stay hot in L2)
//   mode 3: as 0, the line is written by sixteen 2-byte stores per 32-byte sector from 4 warps (partial sectors)
//   mode 4: as 2 with the 2-byte stores of mode 3 (epoch in every 16-bit word)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_latency xchg_latency.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed16(uint16_t* p, uint16_t v) {
    asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}

// buf: [slots][P][32 words]
__global__ void __launch_bounds__(128, 1) xchg(uint32_t* buf, int steps, int mode, int spin_work, long long* out) {
    const int P = gridDim.x, me = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool epoch = (mode == 2 || mode == 4);
    const bool halfw = (mode == 3 || mode == 4);
    long long t0 = 0, acc = 0, acc_store = 0;
    for (int s = 1; s <= steps; ++s) {
        const size_t slot = epoch ? (size_t)(s & 1) : (size_t)s;
        uint32_t* line = buf + (slot * P + me) * 32;
        const uint32_t val = epoch ? (uint32_t)s : 0x3f803f80u;   // bf16 1.0, 1.0
        __syncthreads();
        if (threadIdx.x == 0) t0 = clock64();
        if (halfw) {
            // 64 halves: warp w writes halves {16*i + 4*w + (lane&3)} for i = lane>>2 ... spread like the kernel does
            const int i = lane >> 2;                       // 0..7: 16-byte piece of the line
            const int h = i * 8 + (warp & 1) * 4 + (lane & 3);
            if (warp < 2) st_relaxed16(reinterpret_cast<uint16_t*>(line) + h, epoch ? (uint16_t)s : (uint16_t)0x3f80);
        } else if (warp == 0) {
            st_relaxed(line + lane, val);
        }
        if (mode == 1 && warp == 1 && lane == 0 && s + 2 <= steps)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(buf + ((size_t)(s + 2) * P + me) * 32));
        if (threadIdx.x == 0) acc_store += clock64() - t0;
        if (warp == 0) {
            for (;;) {
                bool ok = true;
                for (int c = lane; c < P; c += 32) {
                    const uint32_t w = ld_relaxed(buf + (slot * P + c) * 32 + 31);   // last word of the peer's line
                    if (epoch) ok = ok && (halfw ? ((w >> 16) == (uint32_t)(s & 0xffff)) : (w == (uint32_t)s));
                    else       ok = ok && ((w >> 16) != 0xffffu);
                }
                if (__all_sync(0xffffffffu, ok)) break;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) acc += clock64() - t0;
        // some "work" between the steps so that the CTAs do not run in perfect lockstep with an idle memory system
        for (int k = 0; k < spin_work; ++k) asm volatile("nanosleep.u32 20;");
    }
    if (threadIdx.x == 0) { out[me * 2] = acc; out[me * 2 + 1] = acc_store; }
}

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 100, steps = argc > 2 ? atoi(argv[2]) : 500;
    uint32_t* buf;
    long long* out;
    const size_t words = (size_t)(steps + 3) * P * 32;
    CK(cudaMalloc(&buf, words * 4));
    CK(cudaMalloc(&out, P * 2 * sizeof(long long)));
    // something large in between so that the freshly filled buffer is not simply sitting in L2
    uint32_t* big;
    const size_t big_bytes = 512ull << 20;
    CK(cudaMalloc(&big, big_bytes));
    static long long h[2 * 148];
    for (int work = 0; work <= 1; ++work) {
        for (int evict = 0; evict <= 1; ++evict) {
            for (int mode = 0; mode <= 4; ++mode) {
                const bool epoch = (mode == 2 || mode == 4);
                CK(cudaMemset(buf, epoch ? 0 : 0xFF, words * 4));
                if (evict) CK(cudaMemset(big, 1, big_bytes));
                xchg<<<P, 128>>>(buf, steps, mode, work * 20, out);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, out, P * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
                double a = 0, b = 0;
                for (int i = 0; i < P; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
                printf("P=%d work=%d evict_L2=%d mode=%d: %.0f cycles store->all visible (store issue %.0f)\n", P, work, evict, mode,
                       a / P / steps, b / P / steps);
            }
        }
    }
    return 0;
}
