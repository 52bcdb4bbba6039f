// Probe: which (TMEM lane, column) does each thread register of the tcgen05.ld shapes 16x64b / 16x128b / 16x256b /
// 32x32b receive?  TMEM is filled with tcgen05.st.32x32b (thread = lane, value = lane * 1000 + column), then read back
// with every shape and printed as a (thread, register) -> (lane, column) table.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_layout tmem_ld_layout.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(int* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    // fill: lane = 32*warp + lane, 16 columns
    const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 16; ++c) {
        const uint32_t v = (uint32_t)((warp * 32 + lane) * 1000 + c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) {   // lane quarter 1: lanes 32..63
        uint32_t r[4];
        // 16x64b.x1: 1 register
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[0 * 128 + lane * 4 + 0] = r[0];
        // 16x128b.x1: 2 registers
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[1 * 128 + lane * 4 + 0] = r[0];
        out[1 * 128 + lane * 4 + 1] = r[1];
        // 16x256b.x1: 4 registers
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out[2 * 128 + lane * 4 + i] = r[i];
        // 16x128b.x2: 4 registers
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out[3 * 128 + lane * 4 + i] = r[i];
        // 16x256b.x1 with the lane field = 16 (second half of the quarter)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr + (16u << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out[4 * 128 + lane * 4 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}

int main() {
    int* d;
    cudaMalloc(&d, 5 * 128 * sizeof(int));
    cudaMemset(d, 0xFF, 5 * 128 * sizeof(int));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static int h[5 * 128];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[5] = {"16x64b.x1", "16x128b.x1", "16x256b.x1", "16x128b.x2", "16x256b.x1 @lane+16"};
    const int nreg[5] = {1, 2, 4, 4, 4};
    for (int s = 0; s < 5; ++s) {
        printf("== %s: thread: (lane,col) per register\n", names[s]);
        for (int t = 0; t < 32; ++t) {
            printf(" t%02d:", t);
            for (int i = 0; i < nreg[s]; ++i) {
                const int v = h[s * 128 + t * 4 + i];
                printf(" (%d,%d)", v / 1000, v % 1000);
            }
            if (t % 4 == 3) printf("\n");
        }
    }
    return 0;
}
