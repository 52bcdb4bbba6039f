// Microbenchmark: does a release (st.global + fence.acq_rel.gpu / red.release.gpu) issued by one warp wait for TMA loads
// that ANOTHER warp of the same CTA has in flight?  grid CTAs; warp 0 streams `kb` boxes of [64 rows x 128 B] from a
// buffer much larger than L2 (so the loads take a while); warp 1 stores one word, then times the fence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membar_tma membar_tma.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode bit 0: TMA loads in flight; bits 1..2: 0 = st + fence.acq_rel.gpu, 1 = st + red.release.gpu, 2 = st + fence.sc.gpu,
//                                            3 = st + membar.cta (control)
__global__ void __launch_bounds__(64, 1)
membar_kernel(const __grid_constant__ CUtensorMap tm, uint32_t* scratch, int nkb, int iters, int mode, int delay, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long acc_fetch = 0, acc_fence = 0;
    uint32_t phase = 0;
    const int kind = mode >> 1;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (warp == 0 && lane == 0 && (mode & 1)) {
            const long long t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nkb * 64 * 128) : "memory");
            const int slab = (it * gridDim.x + blockIdx.x);
            for (int kb = 0; kb < nkb; ++kb)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(smem + (size_t)kb * 8192)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(kb * 64), "r"(0), "r"(slab) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            if (it >= 5) acc_fetch += clock64() - t0;
        }
        if (warp == 1 && lane == 0) {
            const long long ts = clock64();
            while (clock64() - ts < delay) {}
            scratch[blockIdx.x * 64 + (it & 31)] = (uint32_t)it;
            const long long t0 = clock64();
            if (kind == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            else if (kind == 1) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(scratch + 8192 + blockIdx.x * 32), "r"(1u) : "memory");
            else if (kind == 2) asm volatile("fence.sc.gpu;" ::: "memory");
            else asm volatile("membar.cta;" ::: "memory");
            const long long t1 = clock64();
            if (it >= 5) acc_fence += t1 - t0;
        }
        phase ^= (mode & 1);
    }
    if (threadIdx.x == 0) out[2 * blockIdx.x] = acc_fetch / (iters - 5);
    if (threadIdx.x == 32) out[2 * blockIdx.x + 1] = acc_fence / (iters - 5);
}

int main() {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    PFN enc = (PFN)fp;
    const int iters = 40, grid = 100, nkb = 13, Kp = 832;
    const int slabs = iters * grid;
    long long* out; CK(cudaMalloc(&out, 2 * 148 * sizeof(long long)));
    uint32_t* scratch; CK(cudaMalloc(&scratch, 1 << 20)); CK(cudaMemset(scratch, 0, 1 << 20));
    __nv_bfloat16* buf; size_t bytes = (size_t)slabs * 64 * Kp * 2;      // 4000 slabs x 106 KB = 426 MB: every fetch comes from HBM
    CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    CUtensorMap tm;
    cuuint64_t d[3] = {(cuuint64_t)Kp, 64, (cuuint64_t)slabs}, s[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)64 * Kp * 2};
    cuuint32_t bx[3] = {64, 64, 1}, es[3] = {1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, d, s, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const size_t smem = (size_t)nkb * 8192 + 1024;
    CK(cudaFuncSetAttribute(membar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long h[2 * 148];
    const char* names[4] = {"st + fence.acq_rel.gpu", "st + red.release.gpu", "st + fence.sc.gpu", "st + membar.cta"};
    for (int kind = 0; kind < 4; ++kind)
        for (int tma = 0; tma < 2; ++tma)
            for (int delay : {300}) {
                int mode = tma | (kind << 1);
                int nk = nkb, it = iters;
                void* args[] = {(void*)&tm, (void*)&scratch, (void*)&nk, (void*)&it, (void*)&mode, (void*)&delay, (void*)&out};
                CK(cudaLaunchKernel((void*)membar_kernel, dim3(grid), dim3(64), args, smem, 0));
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, out, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost));
                double mf = 0, mm = 0; for (int i = 0; i < grid; ++i) { mf += h[2 * i]; mm += h[2 * i + 1]; }
                printf("%-24s TMA loads in flight: %s : fetch %6.0f cycles, fence %6.0f cycles\n", names[kind], tma ? "yes" : "no ", mf / grid, mm / grid);
            }
    return 0;
}
