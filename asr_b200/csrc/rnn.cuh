// asr_b200 -- declarations shared by the two recurrent kernels (rnn.cu: TMA-fed ring, tf32 / debug paths;
// rnn2.cu: the data-is-the-flag exchange of the bf16 product path).
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "ptx.cuh"

namespace asrb {

// warp 0: TMA producer; warp 1: MMA issuer; warps 2,3: idle (keep warp % 4 == TMEM lane quarter for the rest);
// warps 4..19: epilogue, lane quarter = warp % 4, four 4-unit groups per quarter
constexpr int kRnnCtrlWarps = 4;
constexpr int kRnnEpiWarps = 16;
constexpr int kRnnThreads = (kRnnCtrlWarps + kRnnEpiWarps) * 32;
constexpr int kRnnEpiThreads = kRnnEpiWarps * 32;
constexpr int kRnnCounterStride = 32;                  // step counters [dir], one 128-byte line each
constexpr int kRnnMaxRows = 128;   // batch rows per CTA (one MMA M tile; TMA zero-fills rows >= B)
constexpr int kRnnMaxSmem = 227 * 1024;
constexpr int kRnnMaxStages = 40;
constexpr int kRnnBarBytes = 1024;  // 2 x kRnnMaxStages ring barriers + 3 + TMEM slot, then the bias slice
constexpr int kRnnBiasOffset = 704;

struct RnnParams {
    int T, B, H, G, P, kpad, use_simt, stages, chunk;   // chunk = K blocks per pipeline stage / barrier
    int P_saved;          // slices of the saved-gates layout (= the forward's P; the split backward may pad its own P)
    const int* lengths;
    uint32_t* counters;   // [2 dirs] step counters, kRnnCounterStride words apart
    float* out_sum;       // forward, optional: [T,B,H] = sum of the two directions (blocks.py:92), written by the kernel itself
    int stage_out;        // rnn3.cu: gate-gradient outputs through shared-memory tiles + TMA stores
    int dbg;              // DEBUG timing experiments: 1 = drop the non-critical stores, 2 = drop the operand prefetch
    const float* wpack;  // packed fp32 weight slices (global copy, SIMT debug path)
    // bf16 MMA operands; rows padded to a multiple of 64 elements (128 bytes) so that every 128-byte box row TMA
    // fetches is ONE aligned L2 line (H=800 rows of 1600 bytes would put every odd row across two lines)
    __nv_bfloat16* hbf;   // [2,T+2,B,Hp] bf16 copy of hseq (forward, bf16 mode): the next step's MMA operand
    __nv_bfloat16* dghbf; // [2,T,B,Gp]  bf16 copy of dgh  (backward, bf16 mode)
    int Hp, Gp;
    long long* trace;     // DEBUG: [gridDim][T][16] SM-clock stamps per step (asrb_debug_rnn_trace), else NULL
    // forward
    const float* gi;     // [T,B,2,G]
    const float* b_hh;   // [2,G]
    float* hseq;         // [2,T+2,B,H]
    float* cseq;         // [2,T+2,B,H] (LSTM)
    float* saved;        // [2,T,B,4,H]
    // backward
    const float* dout;   // [T,B,H]
    // gate gradients for the dgrad / wgrad GEMMs: fp32 in tf32 mode, bf16 in bf16 mode (the backward GEMMs then run with
    // bf16 operands: the loss only depends on the forward, and half the bytes leave the kernel)
    void* dgi;           // [T,B,2,G]
    float* dgh;          // [2,T,B,G]  (tf32 / debug modes: MMA operand; NULL in bf16 mode)
    void* dgiT;          // [2G, ldT]  transposed gate gradients (row = dir*G + gate*H + unit, column = t*B + b)
    void* dghT;          // [2G, ldT]  GRU: transposed hidden-side gate gradients (n rows differ from dgiT); NULL for LSTM
    long long ldT;
};

template <int CELL, int NJ>
struct RnnShape {
    static constexpr int kGates = (CELL == ASRB_RNN_GRU) ? 3 : 4;
    static constexpr int kNpadF = ((kGates * NJ + 15) / 16) * 16;
    static constexpr int kNpadB = 16;
    static_assert(NJ <= 16 && kNpadF <= 64, "slice too wide");
};

// ------------------------------------------------------------------------------------------------
// the recurrence
// ------------------------------------------------------------------------------------------------
extern long long* g_rnn_trace;
extern int g_rnn_dbg;
extern int g_rnn_ksplit;   // largest backward K split allowed: 0 none, 2 CTA pairs (default), 4 clusters of four -- the
                        // 1.5 k cycles the MMA phase gains with four are lost again in the longer exchange (measured)
extern int g_rnn_chunk;   // DEBUG: K blocks per pipeline barrier (0 = automatic)

// fast gate non-linearities on the flush-to-zero MUFU forms (ex2.approx.ftz + rcp.approx.ftz: ~1e-6 absolute error);
// __expf / __fdividef spend three more instructions per call on denormal scaling and range checks
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fsigmoid(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float ftanh(float x) { return fmaf(-2.f, rcp_ftz(1.f + ex2_ftz(2.8853900817779268f * x)), 1.f); }

__device__ __forceinline__ void ld4(float* dst, const float* src) {
    const float4 t = *reinterpret_cast<const float4*>(src);
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
}
__device__ __forceinline__ void ldg4(float* dst, const float* src) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src));
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
}
__device__ __forceinline__ void st4(float* dst, const float* v) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* dst, const float* v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst) = u;
}

#define ASRB_TRACE(slot, step)                                                                   \
    do {                                                                                         \
        if (p.trace) p.trace[((size_t)blockIdx.x * p.T + (step)) * 16 + (slot)] = clock64();     \
    } while (0)

// Step-barrier poll with a watchdog: the CTAs of a launch spin on counters the other CTAs bump, which needs the whole grid
// resident at once (checked against the occupancy API before every launch).  Should that ever not hold -- another spinning
// kernel on the device, a MIG slice smaller than reported -- the kernel traps after kRnnWatchdogNs instead of hanging the
// stream forever: the launch then fails loudly with a CUDA error.
constexpr unsigned long long kRnnWatchdogNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void poll_counter(const uint32_t* counter, uint32_t need) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_u32(counter) < need) {
        if ((++spins & 0x3FFFu) == 0) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kRnnWatchdogNs) __trap();
        }
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}


// ---- helpers of the kernels that keep the weight slice in tensor memory / read the accumulator with all 32 lanes
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
// 16 TMEM lanes x 4 columns: thread t receives column t%4 of lanes t/4 (a) and t/4 + 8 (b)  (tools/ubench/tmem_ld_layout.cu)
__device__ __forceinline__ void tmem_ld_16x128b(uint32_t taddr, float& a, float& b) {
    uint32_t ra, rb;
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(ra), "=r"(rb) : "r"(taddr) : "memory");
    a = __uint_as_float(ra);
    b = __uint_as_float(rb);
}
// 16 TMEM lanes x 16 columns: thread t receives, for lanes t/4 (v[0..1], v[4..5]) and t/4 + 8 (v[2..3], v[6..7]),
// columns 2*(t%4) + {0,1} (v[0..3]) and 8 + 2*(t%4) + {0,1} (v[4..7])   (tools/ubench/tmem_ld_layout.cu)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// registers -> TMEM: thread i of the warp writes 8 consecutive 32-bit columns of lane (base lane + i)
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (M x 16 bf16 = 8 columns) is read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// the same, issued only when `enable` != 0 (a predicated instruction: no branch around it, so a run of MMAs with constant
// operand offsets stays straight-line code in the uniform datapath)
__device__ __forceinline__ void umma_f16_ts_if(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t enable) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(enable)
        : "memory");
}

struct RnnPlan {
    int nj, P, npad_f, npad_b, kpad_f, kpad_b, stages_f, stages_b, chunk_f, chunk_b, mrows, bf16, ksplit, P_b;   // ksplit: 0, 2, 4
    int ts_bwd;   // the backward recurrence runs rnn3.cu (weights in tensor memory, clusters of 4): packed layout of ks = 4
    size_t smem_f, smem_b;
};


// rnn2.cu: launches the bf16 exchange-by-data kernel for (cell, plan)
int rnn2_dispatch(bool bwd, int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream);
// rnn3.cu: forward recurrence, weights in tensor memory, two interleaved half-batch chains (B <= 64, H <= 896, nj == 16)
int rnn3_forward(int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream);
// rnn3.cu: backward recurrence, weights in tensor memory, K split over clusters of 4, two chains (plan.ts_bwd)
int rnn3_backward(int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream);
long long rnn3_redo_count();      // launches whose second pass had to run (verified hand-over, rnn3.cu); < 0: CUDA error

}  // namespace asrb
