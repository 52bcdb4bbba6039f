// Microbenchmark: (1) latency of one L2-hit load for several load flavours (dependent chain, one thread);
// (2) one-way store -> visible latency between two CTAs on different SMs (ping-pong of a flag word, one thread each)
// for several store / load flavours; (3) the same with one poller per CTA and 1..64 CTAs polling the same word.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int F>
__device__ __forceinline__ uint32_t ld_f(const uint32_t* p) {
    uint32_t v;
    if (F == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 5) asm volatile("ld.relaxed.cta.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int F>
__device__ __forceinline__ void st_f(uint32_t* p, uint32_t v) {
    if (F == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (F == 1) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (F == 2) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (F == 3) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (F == 4) asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
    if (F == 5) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int F>
__global__ void chase(const uint32_t* buf, int n, long long* out) {
    uint32_t idx = 0;
    for (int i = 0; i < 64; ++i) idx = ld_f<F>(buf + idx);   // warm L2
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) idx = ld_f<F>(buf + idx);
    const long long t1 = clock64();
    out[0] = t1 - t0;
    out[1] = idx;
}

// CTA 0 and CTA `peer` bounce a counter: even values are written by CTA 0 into flag[0], odd by the peer into flag[32]
template <int LF, int SF>
__global__ void pingpong(uint32_t* flag, int iters, int peer, long long* out) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != peer) return;
    const bool a = blockIdx.x == 0;
    uint32_t* mine = flag + (a ? 0 : 32);
    const uint32_t* theirs = flag + (a ? 32 : 0);
    const long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (a) {
            st_f<SF>(mine, (uint32_t)i);
            while (ld_f<LF>(theirs) < (uint32_t)i) {}
        } else {
            while (ld_f<LF>(theirs) < (uint32_t)i) {}
            st_f<SF>(mine, (uint32_t)i);
        }
    }
    if (a) out[0] = clock64() - t0;
}

// one writer (CTA 0) bumps a word every time all pollers have acknowledged; npoll CTAs poll it (one thread each) and
// acknowledge with a relaxed red on a second word.  cycles per round = store -> seen by all -> acks seen by the writer
__global__ void fanout(uint32_t* flag, int iters, int npoll, long long* out) {
    if (threadIdx.x != 0) return;
    if ((int)blockIdx.x > npoll) return;
    if (blockIdx.x == 0) {
        const long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            st_f<0>(flag, (uint32_t)i);
            while (ld_f<0>(flag + 32) < (uint32_t)(i * npoll)) {}
        }
        out[0] = clock64() - t0;
    } else {
        for (int i = 1; i <= iters; ++i) {
            while (ld_f<0>(flag) < (uint32_t)i) {}
            st_f<4>(flag + 32, 1u);
        }
    }
}

__global__ void get_smid(int* out) {
    if (threadIdx.x == 0) {
        uint32_t s;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        out[blockIdx.x] = (int)s;
    }
}

int main() {
    uint32_t* buf;
    long long* out;
    const int n = 1 << 16;
    CK(cudaMalloc(&buf, n * 4 * 33));
    CK(cudaMalloc(&out, 64));
    // pointer chase with a 128-byte stride permutation inside 8 MB (L2 resident)
    {
        uint32_t* h = (uint32_t*)malloc((size_t)n * 32 * 4);
        for (size_t i = 0; i < (size_t)n * 32; ++i) h[i] = 0;
        uint32_t cur = 0;
        for (int i = 0; i < n; ++i) { uint32_t nxt = (uint32_t)(((uint64_t)(i + 1) * 40503u) % n) * 32; h[cur] = nxt; cur = nxt; }
        CK(cudaMemcpy(buf, h, (size_t)n * 32 * 4, cudaMemcpyHostToDevice));
        free(h);
    }
    long long h[8];
    const char* lname[6] = {"ld.relaxed.gpu", "ld.volatile", "ld.cg", "ld.acquire.gpu", "ld.cv", "ld.relaxed.cta"};
    const char* sname[6] = {"st.relaxed.gpu", "st.volatile", "st.cg", "st.release.gpu", "red.relaxed.gpu", "st.wt"};
#define CHASE(F) chase<F><<<1, 1>>>(buf, 2000, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost)); \
    printf("L2-hit dependent load, %-16s: %.0f cycles\n", lname[F], (double)h[0] / 2000);
    CHASE(0) CHASE(1) CHASE(2) CHASE(3) CHASE(4) CHASE(5)
    uint32_t* flag = buf;
    for (int peer = 1; peer <= 147; peer += 73) {
#define PP(LF, SF) CK(cudaMemset(flag, 0, 512)); pingpong<LF, SF><<<148, 32>>>(flag, 2000, peer, out); CK(cudaDeviceSynchronize()); \
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost)); \
    printf("ping-pong CTA0<->CTA%d  %-16s + %-16s: one way %.0f cycles\n", peer, sname[SF], lname[LF], (double)h[0] / 4000);
        PP(0, 0) PP(1, 1) PP(2, 2) PP(3, 3) PP(0, 4) PP(4, 5) PP(2, 0) PP(0, 2)
    }
    // latency map: CTA 0 against every other CTA (relaxed store + volatile load), and the SM each CTA runs on
    {
        int* smid;
        CK(cudaMalloc(&smid, 148 * 4));
        get_smid<<<148, 32>>>(smid);
        CK(cudaDeviceSynchronize());
        int hs[148];
        CK(cudaMemcpy(hs, smid, sizeof(hs), cudaMemcpyDeviceToHost));
        for (int a = 0; a < 4; ++a) {            // four flag addresses, 64 KB apart (different 2 KB home grains)
            uint32_t* f = buf + (size_t)a * 16384 + 64;
            printf("flag address +%d KB: one-way cycles CTA0 (sm %d) <-> CTA p:", a * 64, hs[0]);
            for (int peer = 1; peer < 148; ++peer) {
                CK(cudaMemset(f, 0, 512));
                pingpong<1, 0><<<148, 32>>>(f, 300, peer, out);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
                if (a == 0) printf(" %d(sm%d):%.0f", peer, hs[peer], (double)h[0] / 600);
                else printf(" %.0f", (double)h[0] / 600);
            }
            printf("\n");
        }
    }
    for (int npoll = 1; npoll <= 128; npoll *= 2) {
        CK(cudaMemset(flag, 0, 512));
        fanout<<<148, 32>>>(flag, 1000, npoll, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
        printf("fan-out to %3d polling CTAs: %.0f cycles per round (store -> all saw it -> %d acks seen)\n", npoll, (double)h[0] / 1000, npoll);
    }
    return 0;
}
