// Microbenchmark: how fast can one SM ingest a [rows x K] bf16 operand through TMA (128B-swizzled K blocks) when
// 1..148 CTAs fetch the SAME or DIFFERENT addresses?  (the recurrent kernel's per-step operand fetch, in isolation)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest tma_ingest.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(64, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tm, int nkb, int rows, int iters, int distinct, int slabs, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            const int slab = distinct ? (int)((blockIdx.x * 7 + it) % slabs) : (it % slabs);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nkb * rows * 128) : "memory");
            for (int kb = 0; kb < nkb; ++kb)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(smem + (size_t)kb * rows * 128)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(kb * 64), "r"(0), "r"(slab) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    PFN enc = (PFN)fp;
    const int B = 64, slabs = 512;
    long long* out; CK(cudaMalloc(&out, 148 * sizeof(long long)));
    long long h[148];
    for (int K : {832, 2432}) {
        __nv_bfloat16* buf; size_t bytes = (size_t)slabs * B * K * 2;
        CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        for (int rows : {64, 32}) {
            CUtensorMap tm;
            cuuint64_t d[3] = {(cuuint64_t)K, (cuuint64_t)B, (cuuint64_t)slabs}, s[2] = {(cuuint64_t)K * 2, (cuuint64_t)B * K * 2};
            cuuint32_t bx[3] = {64, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
            if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d, s, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
            const int nkb = K / 64;
            // the ring of the real kernel holds at most ~18 K blocks of 64 rows; here the whole operand is in flight when it fits
            const int nkb_use = nkb * rows * 128 <= 200 * 1024 ? nkb : 200 * 1024 / (rows * 128);
            const size_t smem = (size_t)nkb_use * rows * 128 + 1024;
            CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (int distinct = 0; distinct < 2; ++distinct)
                for (int grid : {1, 25, 50, 100, 148}) {
                    const int iters = 200;
                    ingest_kernel<<<grid, 64, smem>>>(tm, nkb_use, rows, iters, distinct, slabs, out);
                    CK(cudaDeviceSynchronize());
                    ingest_kernel<<<grid, 64, smem>>>(tm, nkb_use, rows, iters, distinct, slabs, out);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    const double cyc = mx / iters, kb = nkb_use * rows * 128 / 1024.0;
                    printf("K=%d rows=%d blocks=%d (%.0f KB) %s grid=%3d : %.0f cycles/fetch = %.1f B/clk/SM\n", K, rows, nkb_use, kb,
                           distinct ? "distinct" : "same    ", grid, cyc, kb * 1024 / cyc);
                }
        }
        CK(cudaFree(buf));
    }
    return 0;
}
