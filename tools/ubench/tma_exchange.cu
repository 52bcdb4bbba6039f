// Microbenchmark: the recurrent kernel's exchange pattern in isolation.  P CTAs each WRITE a 16-column slice of a
// [64 x K] bf16 matrix (8 bytes per thread), pass a grid-wide release/acquire step barrier, then each CTA TMA-fetches
// the WHOLE matrix (13 boxes of 64 rows x 128 B).  Reports cycles from "barrier passed" to "all data landed", for
// freshly written slabs versus slabs written long ago.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_exchange tma_exchange.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: fetch the slab written this iteration; 1: fetch a slab nobody wrote in this launch; 2: write, but fetch an old slab
__global__ void __launch_bounds__(320, 1)
exchange_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm4, int chunk, __nv_bfloat16* buf, uint32_t* counter, int K, int Kp, int iters, int mode,
                int store_kind, int rows, long long* out) {
    const int variant = store_kind >> 4; store_kind &= 15;
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    const int P = gridDim.x, nkb = Kp / 64;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    long long acc_fetch = 0, acc_issue = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
        if (warp >= 2) {   // 256 "epilogue" threads: row b, 4-unit group
            const int e = threadIdx.x - 64, b = e & 63, ug = e >> 6;
            if (mode != 1) {
                const int col = blockIdx.x * 16 + ug * 4;
                if (col < K) {
                    __nv_bfloat16* dst = buf + ((size_t)it * 64 + b) * Kp + col;
                    uint2 v = make_uint2(0x3f803f80u + it, 0x3f803f80u);
                    if (store_kind == 0) *reinterpret_cast<uint2*>(dst) = v;
                    else asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(v.x), "r"(v.y) : "memory");
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (e == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        } else if (warp == 0) {
            const uint32_t need = (uint32_t)P * (uint32_t)(it + 1);
            uint32_t v;
            if ((variant & 5) == 0) do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < need);
            else if (variant & 1) do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < need);
            else do { asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < need);
            if (!(variant & 2)) asm volatile("fence.proxy.async.global;" ::: "memory");
            if (variant & 8) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm4) : "memory");
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                const int slab = mode == 0 ? it : (mode == 1 ? it : (it + iters) );
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((chunk <= 1 ? nkb : (nkb + chunk - 1) / chunk * chunk) * rows * 128) : "memory");
                const int rot = (variant & 16) ? blockIdx.x : 0;
                if (chunk <= 1) for (int i = 0; i < nkb; ++i) {
                    const int kb = (i + rot) % nkb;
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(smem_u32(smem + (size_t)kb * rows * 128)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(kb * 64), "r"(0), "r"(slab) : "memory");
                }
                else { const int nch = (nkb + chunk - 1) / chunk; for (int i = 0; i < nch; ++i) {   // one 4-D box per `chunk` K blocks (blocks past the last are zero-filled)
                    const int kb = ((i + rot) % nch) * chunk;
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(smem_u32(smem + (size_t)kb * rows * 128)), "l"((uint64_t)&tm4), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(kb), "r"(slab) : "memory");
                } }
                const long long t1 = clock64();
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                const long long t2 = clock64();
                if (it >= 10) { acc_issue += t1 - t0; acc_fetch += t2 - t0; }
            }
            phase ^= 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = acc_issue / (iters - 10); out[2 * blockIdx.x + 1] = acc_fetch / (iters - 10); }
}

int main() {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    PFN enc = (PFN)fp;
    const int iters = 400, slabs = 2 * iters;
    long long* out; CK(cudaMalloc(&out, 2 * 148 * sizeof(long long)));
    uint32_t* counter; CK(cudaMalloc(&counter, 256));
    long long h[2 * 148];
    for (int variant : {0, 16})
    for (int chunk : {1, 2, 4})
    for (int rows : {32})
    for (int K : {800}) {
        const int Kp = (K + 63) / 64 * 64;
        if (Kp * 128 > 200 * 1024) { /* 2400: 38 boxes = 304 KB does not fit; fetch the first 24 */ }
        const int Kfetch = Kp * 128 <= 200 * 1024 ? Kp : 1536;
        __nv_bfloat16* buf; size_t bytes = (size_t)slabs * 64 * Kp * 2;
        CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        CUtensorMap tm;
        cuuint64_t d[3] = {(cuuint64_t)Kfetch, 64, (cuuint64_t)slabs}, s[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)64 * Kp * 2};
        cuuint32_t bx[3] = {64, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d, s, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        CUtensorMap tm4;
        cuuint64_t d4[4] = {64, (cuuint64_t)rows, (cuuint64_t)(Kfetch / 64), (cuuint64_t)slabs}, s4[3] = {(cuuint64_t)Kp * 2, 128, (cuuint64_t)64 * Kp * 2};
        cuuint32_t bx4[4] = {64, (cuuint32_t)rows, (cuuint32_t)chunk, 1}, es4[4] = {1, 1, 1, 1};
        if (enc(&tm4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, d4, s4, bx4, es4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode4 failed\n"); return 1; }
        const size_t smem = (size_t)(Kfetch / 64 + 13) * 8192 + 1024;
        CK(cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int P = (K + 15) / 16;   // 50 or 150 -> clamp to 100 for the wide case (each CTA then writes only its slice)
        const int grid = P > 100 ? 100 : P;
        for (int store_kind = 0; store_kind < 1; ++store_kind)
            for (int mode = 0; mode < 1; ++mode) {
                int sk = store_kind | (variant << 4);
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaMemset(counter, 0, 256));
                    void* args[] = {(void*)&tm, (void*)&tm4, (void*)&chunk, (void*)&buf, (void*)&counter, (void*)&K, (void*)&Kp, (void*)&iters, (void*)&mode, (void*)&sk, (void*)&rows, (void*)&out};
                    int Kk = Kfetch; args[6] = (void*)&Kk;
                    CK(cudaLaunchCooperativeKernel((void*)exchange_kernel, dim3(grid), dim3(320), args, smem, 0));
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h, out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost));
                double mi = 0, mf = 0; for (int i = 0; i < grid; ++i) { mi += h[2 * i]; mf += h[2 * i + 1]; }
                printf("variant=%d chunk=%d rows=%d K=%d fetch %d KB grid=%d store=%s mode=%s : issue %.0f cycles, all landed %.0f cycles\n", variant, chunk, rows, K, Kfetch / 64 * rows / 8, grid,
                       store_kind ? "st.cg" : "st   ", mode == 0 ? "fresh slab    " : (mode == 1 ? "no writes     " : "write,read old"), mi / grid, mf / grid);
            }
        CK(cudaFree(buf));
    }
    return 0;
}
