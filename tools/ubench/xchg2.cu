// Microbenchmark of the recurrent kernel's step hand-over in isolation ("data is the flag" protocol of rnn2.cu):
// 2 x P CTAs (two independent groups = the two directions).  Every step each CTA (1) stores its [B x 16] bf16 piece of
// the step's slab, (2) its warp 0 polls canary words until the slab looks complete, (3) copies the WHOLE slab into shared
// memory, (4) all threads scan the copy for the fill pattern (stands in for the NaN test; counted, and waited out by
// re-polling so that the loop stays correct).  Reports cycles per step and per phase.
//   store kinds : 0 = 2-byte st.relaxed.gpu, thread = (row, unit)   [16x128b epilogue mapping of rnn2.cu]
//                 1 = 2-byte plain st (weak)
//                 2 = 8-byte plain st, thread = (row, 4 units)       [32x32b mapping]
//                 3 = 16-byte plain st, thread = (row, 8 units)
//                 4 = 16-byte plain st, lane pairs cover a row's 32 bytes (chunk-major slab: still two half sectors)
//   copy kinds  : 0 = cp.async.bulk per 64-column K block (chunk-major slab), 1 = one cp.async.bulk for the whole slab,
//                 2 = per-thread 16-byte ld.volatile + st.shared
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg2 xchg2.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool half_unwritten(uint32_t w) { const uint32_t x = ~w; return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u; }
__device__ __forceinline__ uint32_t ld_vol(const void* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

constexpr int B = 64, NJ = 16;

// slab layout: [K/8 chunks][B rows][8] bf16 ; CTA p of a group owns columns 16p..16p+15 = chunks 2p, 2p+1
__global__ void __launch_bounds__(640, 1)
xchg(uint16_t* buf, int K, int steps, int store_kind, int copy_kind, int canary, long long* out, int* bad_out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ int s_bad;
    const int P = gridDim.x / 2, grp = blockIdx.x / P, me = blockIdx.x % P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t slab_elems = (size_t)K * B;
    const int nkb = (K + 63) / 64;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_bad = 0;
    }
    __syncthreads();
    long long t_step = 0, t_can = 0, t_copy = 0, t_store = 0;
    long long t0 = 0, t1 = 0;
    uint32_t phase = 0;
    int bad_total = 0;
    for (int s = 0; s < steps; ++s) {
        uint16_t* slab = buf + ((size_t)grp * steps + s) * slab_elems;
        __syncthreads();
        if (threadIdx.x == 0) t0 = clock64();
        // ---- (1) stores by the 16 "epilogue" warps ----
        if (warp >= 4) {
            const int wq = warp - 4;
            const uint16_t val = (uint16_t)(0x3c00 + (s & 0xff));
            if (store_kind <= 1) {
                const int quad = wq & 3, ug = wq >> 2, unit = 16 * me + 4 * ug + (lane & 3);
                for (int c = 0; c < 2; ++c) {
                    const int row = quad * 16 + (lane >> 2) + 8 * c;
                    uint16_t* dst = slab + ((size_t)(unit >> 3) * B + row) * 8 + (unit & 7);
                    if (unit < K) {
                        if (store_kind == 0) asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(dst), "h"(val) : "memory");
                        else *reinterpret_cast<volatile uint16_t*>(dst) = val;
                    }
                }
            } else if (store_kind == 2) {
                // thread = (row, 4 units): 16 warps x 32 lanes = 64 rows x 4 groups x 2 (half of the threads idle)
                const int e = wq * 32 + lane;
                if (e < 256) {
                    const int row = e & 63, ug = e >> 6, unit = 16 * me + 4 * ug;
                    uint16_t* dst = slab + ((size_t)(unit >> 3) * B + row) * 8 + (unit & 7);
                    const uint32_t v2 = val | ((uint32_t)val << 16);
                    if (unit < K) asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(v2), "r"(v2) : "memory");
                }
            } else if (store_kind == 4) {
                // full 32-byte sectors: two adjacent lanes write the two 16-byte halves of a row's 16 units
                const int e = wq * 32 + lane;
                if (e < 128) {
                    const int row = e >> 1, ch = e & 1, unit = 16 * me + 8 * ch;
                    uint16_t* dst = slab + ((size_t)(unit >> 3) * B + row) * 8;
                    const uint32_t v2 = val | ((uint32_t)val << 16);
                    if (unit < K) asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v2), "r"(v2), "r"(v2), "r"(v2) : "memory");
                }
            } else {
                const int e = wq * 32 + lane;
                if (e < 128) {
                    const int row = e & 63, ch = e >> 6, unit = 16 * me + 8 * ch;
                    uint16_t* dst = slab + ((size_t)(unit >> 3) * B + row) * 8;
                    const uint32_t v2 = val | ((uint32_t)val << 16);
                    if (unit < K) asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v2), "r"(v2), "r"(v2), "r"(v2) : "memory");
                }
            }
        }
        if (threadIdx.x == 0) t_store += clock64() - t0;
        // ---- (2) canary gate + (3) copy, by warp 0 ----
        if (warp == 0) {
            const int ncan = K / 16;
            if (canary) {
                const uint16_t* can = slab + ((size_t)B + (B - 1)) * 8 + 4;
                for (;;) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    int n = 0;
                    for (int i = lane; i < ncan; i += 32) w[n++] = ld_vol(can + (size_t)i * 16 * B);
                    bool ok = true;
                    for (int i = 0; i < n; ++i) ok = ok && !half_unwritten(w[i]);
                    if (__all_sync(0xffffffffu, ok)) break;
                }
            }
            if (lane == 0) { t1 = clock64(); t_can += t1 - t0; }
            if (copy_kind <= 1 && lane == 0) {
                const uint32_t total = (uint32_t)(slab_elems * 2);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(total) : "memory");
                if (copy_kind == 0) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        const int cols = min(64, K - kb * 64);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(smem_u32(smem + (size_t)kb * 64 * B * 2)), "l"(slab + (size_t)kb * 64 * B), "r"(cols * B * 2), "r"(smem_u32(&bar)) : "memory");
                    }
                } else {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(smem)), "l"(slab), "r"(total), "r"(smem_u32(&bar)) : "memory");
                }
            }
        }
        if (copy_kind == 2) {
            __syncthreads();   // (the gate) -- all threads copy
            const int nchunk = (int)(slab_elems / 8);
            for (int i = threadIdx.x; i < nchunk; i += blockDim.x) {
                uint4 v;
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slab + (size_t)i * 8) : "memory");
                reinterpret_cast<uint4*>(smem)[i] = v;
            }
            __syncthreads();
        } else {
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
        }
        if (threadIdx.x == 0) t_copy += clock64() - t1;
        // ---- (4) scan for the fill pattern; wait stragglers out ----
        {
            const int nchunk = (int)(slab_elems / 8);
            int bad = 0;
            for (int i = threadIdx.x; i < nchunk; i += blockDim.x) {
                uint4 v = reinterpret_cast<const uint4*>(smem)[i];
                if (half_unwritten(v.x) | half_unwritten(v.y) | half_unwritten(v.z) | half_unwritten(v.w)) {
                    ++bad;
                    do {
                        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slab + (size_t)i * 8) : "memory");
                    } while (half_unwritten(v.x) | half_unwritten(v.y) | half_unwritten(v.z) | half_unwritten(v.w));
                }
            }
            if (bad) atomicAdd(&s_bad, bad);
        }
        __syncthreads();
        if (threadIdx.x == 0) { t_step += clock64() - t0; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bad_total = s_bad;
        out[blockIdx.x * 4 + 0] = t_step / steps; out[blockIdx.x * 4 + 1] = t_store / steps;
        out[blockIdx.x * 4 + 2] = t_can / steps; out[blockIdx.x * 4 + 3] = t_copy / steps;
        bad_out[blockIdx.x] = bad_total;
    }
}

int main(int argc, char** argv) {
    const int steps = 300;
    long long* out; int* bad;
    CK(cudaMalloc(&out, 4 * 148 * sizeof(long long)));
    CK(cudaMalloc(&bad, 148 * sizeof(int)));
    static long long h[4 * 148]; static int hb[148];
    for (int K : {800}) {
        const int P = 50;
        uint16_t* buf;
        const size_t bytes = (size_t)2 * steps * K * B * 2;
        CK(cudaMalloc(&buf, bytes));
        const size_t smem = (size_t)K * B * 2 + 2048;
        CK(cudaFuncSetAttribute(xchg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int canary = 1; canary >= 0; --canary)
            for (int copy_kind = 0; copy_kind < 3; ++copy_kind)
                for (int store_kind = 0; store_kind < 5; ++store_kind) {
                    if (!canary && (store_kind != 0 || copy_kind == 2)) continue;
                    for (int rep = 0; rep < 2; ++rep) {
                        CK(cudaMemset(buf, 0xFF, bytes));
                        xchg<<<2 * P, 640, smem>>>(buf, K, steps, store_kind, copy_kind, canary, out, bad);
                        CK(cudaDeviceSynchronize());
                    }
                    CK(cudaMemcpy(h, out, 4 * 2 * P * sizeof(long long), cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(hb, bad, 2 * P * sizeof(int), cudaMemcpyDeviceToHost));
                    double a[4] = {0, 0, 0, 0}; long nb = 0;
                    for (int i = 0; i < 2 * P; ++i) { for (int k = 0; k < 4; ++k) a[k] += h[4 * i + k]; nb += hb[i]; }
                    printf("K=%d slab %d KB canary=%d copy=%d store=%d: step %.0f cycles (stores issued %.0f, canaries ok %.0f, copy after gate %.0f), unwritten chunks seen %.1f per CTA-step\n",
                           K, K * B * 2 / 1024, canary, copy_kind, store_kind, a[0] / (2 * P), a[1] / (2 * P), a[2] / (2 * P), a[3] / (2 * P),
                           (double)nb / (2 * P) / steps);
                    fflush(stdout);
                }
        CK(cudaFree(buf));
    }
    return 0;
}
