// asr_b200 -- persistent bidirectional GRU / LSTM recurrence, bf16 product path: "the data is the flag".
//
// Same decomposition as rnn.cu (CTA (dir, p) owns NJ hidden units, its W_hh slice is resident in shared memory, one
// launch covers all T steps of both directions; modules/blocks.py:87-89 of the reference), but the step-to-step
// hand-over between the CTAs has no counter, no release/acquire pair and no fence on the critical path:
//
//   * the bf16 exchange buffer (hbf [2,T+2,B,Hp] / dghbf [2,T,B,Gp]: one slot per time step, never reused within a
//     launch) is filled with the bit pattern 0xFFFF before the launch -- a value the producers can never write: h lies
//     in (-1, 1), and cvt.rn.bf16 turns a NaN gradient into the canonical 0x7FFF;
//   * a producer thread stores its 4 hidden units of one batch row (8 bytes) as soon as it has them
//     (st.relaxed.gpu: straight to L2);
//   * every CTA's 16 epilogue warps fetch the whole previous state with 16-byte ld.relaxed.gpu loads (L1 bypassed),
//     re-issue a load for as long as one of its eight values still reads 0xFFFF, and write the chunk into shared
//     memory in the tcgen05 K-major / 128-byte-swizzle layout; one mbarrier per 128-byte K block hands it to the MMA
//     warp, so the tensor core starts on the first K blocks while the rest are in flight.
//
// Every 16-bit word is individually either "not yet written" or final, so correctness needs no ordering between
// different words -- which is exactly what made the counter protocol slow (rnn.cu, DESIGN.md section 6: MEMBAR.ALL.GPU
// before the release, ~2 k cycles release -> acquire visibility, proxy fence, first-TMA latency: ~5 k of the 10 k cycles
// of a step in which the SM did nothing).
//
// The recurrent product itself (tcgen05.mma kind::f16, M = batch rows, N = the slice's gate columns, fp32 accumulate in
// TMEM), the gate math, the saved-activation layout and the backward K split over CTA pairs (partial sums exchanged
// through distributed shared memory) are those of rnn.cu.
#include "rnn.cuh"

namespace asrb {


__device__ __forceinline__ uint4 ld_relaxed_u4(const void* ptr) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const void* ptr) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_bf16(__nv_bfloat16* dst, float v) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(dst), "h"(*reinterpret_cast<const uint16_t*>(&b)) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// 16 TMEM lanes x 4 columns: thread t receives column t%4 of lanes t/4 (a) and t/4 + 8 (b)  (tools/ubench/tmem_ld_layout.cu)
__device__ __forceinline__ void tmem_ld_16x128b(uint32_t taddr, float& a, float& b) {
    uint32_t ra, rb;
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(ra), "=r"(rb) : "r"(taddr) : "memory");
    a = __uint_as_float(ra);
    b = __uint_as_float(rb);
}
// does one of the two 16-bit halves of w still hold the fill pattern 0xFFFF?  ("has a zero half" of ~w)
__device__ __forceinline__ bool half_unwritten(uint32_t w) {
    const uint32_t x = ~w;
    return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u;
}
__device__ __forceinline__ bool chunk_unwritten(const uint4& v) {
    return half_unwritten(v.x) | half_unwritten(v.y) | half_unwritten(v.z) | half_unwritten(v.w);
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// barrier / small-state block of this kernel (after the A ring)
constexpr int kX2BarBytes = 1536;
constexpr int kX2TmaBarOff = 2 * kRnnMaxStages;        // in uint64 units: tma_bar[kRnnMaxStages]
constexpr int kX2MiscOff = 3 * kRnnMaxStages;          // w_bar, tfull_bar, tmem slot, x_bar[2], patch counter
constexpr int kX2BiasOffset = 1024;                    // bytes: [kGates][NJ] floats

// trace slots of this kernel (asrb_debug_rnn_trace): 4 step top, 0 canary words valid (TMA warp), 1 TMA loads issued,
// 2 MMA warp: first K block handed over, 3 MMA warp: last commit issued, 10 last K block validated, 5 accumulator
// complete, 6 accumulator in registers, 11 exchange done (K split), 7 operand stored, 8 other stores issued,
// 9 cumulative count of chunks the CTA had to re-fetch because the TMA copy still held the fill pattern
template <int CELL, int NJ, bool BWD, int MROWS, int KS>
__global__ void __launch_bounds__(kRnnThreads, 1)
rnn_rec2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const RnnParams p) {
    static_assert(KS == 1 || BWD, "the K split exists for the backward recurrence only");
    constexpr bool KSPLIT = KS > 1;
    using S = RnnShape<CELL, NJ>;
    constexpr int kGates = S::kGates;
    constexpr int NPAD = BWD ? (KSPLIT ? KS * NJ : S::kNpadB) : S::kNpadF;
    constexpr int kTmemCols = 64;
    constexpr int KBE = 64;                      // bf16 elements per 128-byte K block
    constexpr int kSlotBytes = MROWS * 128;      // one K block of the A tile: MROWS rows x 128 B
    constexpr int kRowsPerWarp = MROWS / 4;      // TMEM lane quarter -> batch rows (M=64: 16 lanes of each quarter)
    constexpr int kRowGroups = MROWS / 64;       // 4-row groups of a K block a validating warp owns
    constexpr int kCells = MROWS / 32;           // (batch row, hidden unit) cells of an epilogue thread

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = p.kpad / KBE;
    const int NS = p.stages;                                  // K-block slots of the A ring (NS == nkb: no reuse inside a step)
    const bool ring = NS < nkb;
    uint8_t* smem_w = smem;                                   // nkb x [NPAD rows x 128 B]
    uint8_t* smem_a = smem_w + (size_t)nkb * NPAD * 128;      // NS x [MROWS rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)NS * kSlotBytes);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kX2BiasOffset);   // [kGates][NJ] (forward)
    uint64_t* full_bar = bars;                    // K block validated by all 16 warps -> MMA warp
    uint64_t* empty_bar = bars + kRnnMaxStages;   // MMAs that read the slot have completed -> TMA warp (ring only)
    uint64_t* tma_bar = bars + kX2TmaBarOff;      // the TMA copy of the K block has landed -> validating warps
    uint64_t* w_bar = bars + kX2MiscOff;
    uint64_t* tfull_bar = w_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 2);
    uint64_t* x_bar = w_bar + 3;                  // [2] KSPLIT: the peer's partial sums of parity 0 / 1 have arrived
    uint32_t* patch_count = reinterpret_cast<uint32_t*>(w_bar + 5);
    float* xbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kX2BarBytes);   // [2][KS sources][MROWS][NJ]
    float* xstage = xbuf + 2 * KS * MROWS * NJ;   // [2][KS-1 peers][MROWS][NJ]: our partial sums of each peer's units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, B = p.B, H = p.H, G = p.G, P = p.P;
    const int dir = blockIdx.x / P, pidx = blockIdx.x % P;
    const int j0 = pidx * NJ;
    const uint32_t crank = KSPLIT ? cluster_ctarank() : 0u;      // = pidx % KS: which part of K this CTA multiplies
    const int K = BWD ? G : H;                                   // columns of the exchanged operand
    const int pitch = BWD ? p.Gp : p.Hp;
    const int kcol0 = KSPLIT ? (int)crank * p.kpad : 0;          // first operand column of this CTA's K part
    const __nv_bfloat16* xchg = BWD ? p.dghbf : p.hbf;
    auto t_of = [&](int s) { return (BWD ? (dir == 0) : (dir == 1)) ? (T - 1 - s) : s; };
    // slab (time slot of the exchange buffer) holding the operand of sequential step s: what step s-1 produced
    auto slab_of = [&](int s) {
        const int tp = t_of(s - 1);
        return BWD ? (dir * T + tp) : (dir * (T + 2) + tp + 1);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmA);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full_bar[i], kRnnEpiWarps);
            mbar_init(&empty_bar[i], 1);
            mbar_init(&tma_bar[i], 1);
        }
        mbar_init(w_bar, 1);
        mbar_init(tfull_bar, 1);
        if (KSPLIT) {
            mbar_init(&x_bar[0], 1);   // armed by one local thread with the byte count the peer will send
            mbar_init(&x_bar[1], 1);
        }
        *patch_count = 0u;
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if constexpr (KSPLIT) cluster_sync_all();   // the peer's exchange barriers exist before anybody arrives on them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA warp: weights once, then every step's operand as soon as it looks complete ==========
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, (uint32_t)(nkb * NPAD * 128));
            for (int kb = 0; kb < nkb; ++kb)
                tma_load_2d(smem_w + (size_t)kb * NPAD * 128, &tmW, w_bar, kb * KBE, (dir * P + pidx) * NPAD);
        }
        __syncwarp();
        // canary words: the last two units of every 16-column group of our K part, in the last batch row -- a heuristic
        // for "the producers have stored this step" (the validation below is what guarantees it)
        const int kcols = min(p.kpad, K - kcol0);
        const int ncan = kcols / 16;
        int slot = 0;
        uint32_t use = 0;
        for (int s = 1; s < T; ++s) {
            const int slab = slab_of(s);
            const __nv_bfloat16* can = xchg + ((size_t)slab * B + (B - 1)) * pitch + kcol0 + 12;
            for (;;) {
                bool ok = true;
                for (int i = lane; i < ncan; i += 32) ok = ok && !half_unwritten(ld_relaxed_u32(can + 16 * i));
                if (__all_sync(0xffffffffu, ok)) break;
            }
            if (lane == 0) ASRB_TRACE(0, s);
            // the slots are free once our own MMAs of the previous step have read them (the canaries of the other CTAs
            // say nothing about that); with a ring the per-slot barrier below covers it
            if (!ring && s > 1) mbar_wait(tfull_bar, (uint32_t)(s & 1));
            for (int kb = 0; kb < nkb; ++kb) {
                if (ring && use > 0) mbar_wait(&empty_bar[slot], (use - 1) & 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&tma_bar[slot], (uint32_t)kSlotBytes);
                    tma_load_3d(smem_a + (size_t)slot * kSlotBytes, &tmA, &tma_bar[slot], kcol0 + kb * KBE, 0, slab);
                }
                __syncwarp();
                if (++slot == NS) { slot = 0; ++use; }
            }
            if (lane == 0) ASRB_TRACE(1, s);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        constexpr uint32_t idesc = umma_idesc(kFmtBF16, MROWS, NPAD);
        mbar_wait(w_bar, 0);
        int slot = 0;
        uint32_t phase = 0;
        for (int s = 1; s < T; ++s) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[slot], phase);
                if (kb == 0 && lane == 0) ASRB_TRACE(2, s);
                tc_fence_after_sync();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + (size_t)slot * kSlotBytes));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_w + (size_t)kb * NPAD * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // 4 x 32-byte K slices (K = 16 bf16) per 128-byte block
                        umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    if (ring) umma_commit(&empty_bar[slot]);
                    if (kb == nkb - 1) umma_commit(tfull_bar);
                }
                __syncwarp();
                if (++slot == NS) { slot = 0; phase ^= 1; }
            }
            if (lane == 0) ASRB_TRACE(3, s);
        }
    } else if (warp >= kRnnCtrlWarps) {
        // ===================== validation + epilogue: 16 warps =====================
        // epilogue role: lane quarter q = warp % 4 (a warp may only read TMEM lanes 32*(warp%4)..+31), unit group ug;
        // the accumulator is read with tcgen05.ld.16x128b so that ALL 32 lanes hold cells also in the M=64 layout (16
        // TMEM lanes per quarter): thread = hidden unit 4*ug + lane%4 of batch rows lane/4 + 8c of the quarter.
        // validation role: warp wq owns the 4-row groups wq (and wq+16) of every K block; lane = (row, 16-byte chunk).
        constexpr int NV = NJ / 4;
        const int quad = warp & 3;
        const int wq = warp - kRnnCtrlWarps;
        const int ug = wq >> 2, ul = lane & 3;
        const int hl = wq * 32 + lane;                        // 0..511
        const int ju = 4 * ug + ul;                           // unit within the slice
        const int unit = j0 + ju;
        const bool uvalid = (ug < NV) && (unit < H);
        const bool warp_ld = (quad * kRowsPerWarp < B) && (ug < NV);   // warp-uniform: this warp reads the accumulator
        int row[kCells], len[kCells];
        bool cellok[kCells];
#pragma unroll
        for (int c = 0; c < kCells; ++c) {
            row[c] = quad * kRowsPerWarp + (lane >> 2) + 8 * c;
            cellok[c] = uvalid && row[c] < B;
            len[c] = cellok[c] ? p.lengths[row[c]] : 0;
        }
        const size_t slotHB = (size_t)B * H;

        // validation geometry
        const int frow = lane >> 3, fchunk = lane & 7;
        int f_row[kRowGroups];
        bool f_rowok[kRowGroups];
        uint32_t f_soff[kRowGroups];
#pragma unroll
        for (int r = 0; r < kRowGroups; ++r) {
            f_row[r] = 4 * (wq + kRnnEpiWarps * r) + frow;
            f_rowok[r] = f_row[r] < B;
            f_soff[r] = (uint32_t)(f_row[r] * 128 + ((fchunk ^ (f_row[r] & 7)) << 4));
        }
        const uint32_t smem_a_u32 = smem_u32(smem_a);
        int f_slot = 0;
        uint32_t f_phase = 0;

        float xsend[KS > 1 ? KS - 1 : 1][kCells] = {};
        uint32_t xphase[2] = {0u, 0u};
        float state_h[kCells], state_c[kCells];   // fwd: h / c of the previous step ; bwd: direct dh / dc carries
#pragma unroll
        for (int c = 0; c < kCells; ++c) state_h[c] = state_c[c] = 0.f;

        if constexpr (!BWD) {
            // zero boundary slots 0 and T+1 of our cells (read by the dW_hh product of the backward pass)
#pragma unroll
            for (int c = 0; c < kCells; ++c) {
                if (cellok[c]) {
                    const size_t o = ((size_t)dir * (T + 2)) * slotHB + (size_t)row[c] * H + unit;
                    p.hseq[o] = 0.f;
                    p.hseq[o + (size_t)(T + 1) * slotHB] = 0.f;
                    if constexpr (CELL == ASRB_RNN_LSTM) {
                        p.cseq[o] = 0.f;
                        p.cseq[o + (size_t)(T + 1) * slotHB] = 0.f;
                    }
                }
            }
            for (int i = hl; i < kGates * NJ; i += kRnnEpiThreads) {
                const int g = i / NJ, jj = i % NJ;
                s_bias[i] = (j0 + jj < H) ? p.b_hh[(size_t)dir * G + g * H + j0 + jj] : 0.f;
            }
            named_bar_sync(3, kRnnEpiThreads);
        }
        float bias[BWD ? 1 : kGates];
        if constexpr (!BWD) {
#pragma unroll
            for (int g = 0; g < kGates; ++g) bias[g] = s_bias[g * NJ + (ju < NJ ? ju : 0)];
        }

        for (int s = 0; s < T; ++s) {
            const int t = t_of(s);
            bool active[kCells];
#pragma unroll
            for (int c = 0; c < kCells; ++c) active[c] = cellok[c] && (t < len[c]);
            if (hl == 0) ASRB_TRACE(4, s);
            if constexpr (KSPLIT) {
                if (hl == 0 && s > 0) mbar_arrive_expect_tx(&x_bar[s & 1], (uint32_t)((KS - 1) * MROWS * NJ * 4));
            }
            constexpr int kAccG = BWD ? 1 : kGates;
            float acc[kAccG][kCells];
#pragma unroll
            for (int g = 0; g < kAccG; ++g)
#pragma unroll
                for (int c = 0; c < kCells; ++c) acc[g][c] = 0.f;

            // ---- operand prefetch (independent of the recurrent product) ----
            constexpr int kIn = BWD ? 6 : kGates;        // fwd: gi gates ; bwd: 4 saved + dout + previous state
            float in[kIn][kCells];
            float ct[kCells];
#pragma unroll
            for (int c = 0; c < kCells; ++c) {
#pragma unroll
                for (int q = 0; q < kIn; ++q) in[q][c] = 0.f;
                ct[c] = 0.f;
                if (active[c]) {
                    if constexpr (!BWD) {
                        const float* g = p.gi + (((size_t)t * B + row[c]) * 2 + dir) * G + unit;
#pragma unroll
                        for (int q = 0; q < kGates; ++q) in[q][c] = __ldg(g + (size_t)q * H);
                    } else {
                        const float* sv = p.saved + ((((size_t)dir * T + t) * p.P_saved + pidx) * 4) * (size_t)(NV * B * 4) +
                                          (size_t)ug * (B * 4) + (size_t)row[c] * 4 + ul;
                        const int tprev_slot = (dir == 0) ? t : t + 2;   // slot of the step that preceded t in forward order
                        const size_t ro = (size_t)row[c] * H + unit;
#pragma unroll
                        for (int q = 0; q < 4; ++q) in[q][c] = sv[(size_t)q * NV * (B * 4)];
                        in[4][c] = __ldg(p.dout + (size_t)t * slotHB + ro);
                        in[5][c] = (CELL == ASRB_RNN_GRU ? p.hseq : p.cseq)[((size_t)dir * (T + 2) + tprev_slot) * slotHB + ro];
                        if constexpr (CELL == ASRB_RNN_LSTM) ct[c] = p.cseq[((size_t)dir * (T + 2) + t + 1) * slotHB + ro];
                    }
                }
            }

            if (s > 0) {
                // ---- validate the TMA copy of the previous step's state, K block by K block ----
                {
                    const __nv_bfloat16* src = xchg + (size_t)slab_of(s) * B * pitch + kcol0 + fchunk * 8;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&tma_bar[f_slot], f_phase);
                        const bool colok = kcol0 + kb * KBE + fchunk * 8 < K;
#pragma unroll
                        for (int r = 0; r < kRowGroups; ++r) {
                            if (f_rowok[r] && colok) {
                                const uint32_t sa = smem_a_u32 + (uint32_t)f_slot * kSlotBytes + f_soff[r];
                                uint4 v = ld_shared_u4(sa);
                                if (chunk_unwritten(v)) {
                                    // the copy overtook the producer: fetch the chunk ourselves until it is there
                                    const __nv_bfloat16* g = src + (size_t)f_row[r] * pitch + kb * KBE;
                                    do { v = ld_relaxed_u4(g); } while (chunk_unwritten(v));
                                    st_shared_u4(sa, v);
                                    fence_proxy_async_smem();      // generic-proxy store -> the tensor core's async-proxy reads
                                    atomicAdd(patch_count, 1u);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_bar[f_slot]);
                        if (++f_slot == NS) { f_slot = 0; f_phase ^= 1; }
                    }
                    if (hl == 0) ASRB_TRACE(10, s);
                }

                // ---- recurrent product for this step ----
                mbar_wait(tfull_bar, (uint32_t)((s - 1) & 1));
                if (hl == 0) ASRB_TRACE(5, s);
                tc_fence_after_sync();
                if (warp_ld) {
#pragma unroll
                    for (int h2 = 0; h2 < kCells / 2; ++h2) {
                        const uint32_t taddr = tmem_base + (uint32_t(quad * 32 + 16 * h2) << 16) + 4 * ug;
                        if constexpr (KSPLIT) {
                            tmem_ld_16x128b(taddr + crank * NJ, acc[0][2 * h2], acc[0][2 * h2 + 1]);   // our own units
#pragma unroll
                            for (int q = 1; q < KS; ++q)                                               // every peer's units
                                tmem_ld_16x128b(taddr + ((crank + q) % KS) * NJ, xsend[q - 1][2 * h2], xsend[q - 1][2 * h2 + 1]);
                        } else {
#pragma unroll
                            for (int g = 0; g < kAccG; ++g)
                                tmem_ld_16x128b(taddr + (BWD ? 0 : g * NJ), acc[g][2 * h2], acc[g][2 * h2 + 1]);
                        }
                    }
                    tmem_ld_wait();
                }
                tc_fence_before_sync();   // our next full-barrier arrival orders these reads before the next step's MMAs
                if (hl == 0) ASRB_TRACE(6, s);
                if constexpr (KSPLIT) {
                    // exchange of the partial sums through distributed shared memory (see rnn.cu for the alternatives
                    // measured): staged locally, ONE bulk copy per peer, double-buffered by step parity
                    const int par = s & 1;
                    if (warp_ld) {
#pragma unroll
                        for (int c = 0; c < kCells; ++c)
#pragma unroll
                            for (int q = 1; q < KS; ++q)
                                xstage[(((size_t)par * (KS - 1) + (q - 1)) * MROWS + row[c]) * NJ + ju] = xsend[q - 1][c];
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(5, kRnnEpiThreads);
                    if (hl == 0) {
#pragma unroll
                        for (int q = 1; q < KS; ++q) {
                            const uint32_t peer = (crank + q) % KS;
                            dsmem_bulk_copy(map_to_cta(xbuf + ((size_t)par * KS + crank) * MROWS * NJ, peer),
                                            xstage + ((size_t)par * (KS - 1) + (q - 1)) * MROWS * NJ,
                                            (uint32_t)(MROWS * NJ * 4), map_to_cta(&x_bar[par], peer));
                        }
                    }
                    mbar_wait_cluster(&x_bar[par], xphase[par]);
                    xphase[par] ^= 1u;
                    if (warp_ld) {
#pragma unroll
                        for (int c = 0; c < kCells; ++c)
#pragma unroll
                            for (int q = 1; q < KS; ++q)
                                acc[0][c] += xbuf[(((size_t)par * KS + (crank + q) % KS) * MROWS + row[c]) * NJ + ju];
                    }
                    if (hl == 0) ASRB_TRACE(11, s);
                }
            }

            // ---- cell math (registers only) ----
            if constexpr (!BWD) {
                float hn[kCells], cn[kCells], sv[4][kCells];
#pragma unroll
                for (int c = 0; c < kCells; ++c) {
                    float h_ = 0.f, c_ = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    if (active[c]) {
                        if constexpr (CELL == ASRB_RNN_GRU) {
                            const float gn = acc[2][c] + bias[2];
                            const float r = fsigmoid(in[0][c] + acc[0][c] + bias[0]);
                            const float z = fsigmoid(in[1][c] + acc[1][c] + bias[1]);
                            const float n = ftanh(in[2][c] + r * gn);
                            h_ = (1.f - z) * n + z * state_h[c];
                            s0 = r; s1 = z; s2 = n; s3 = gn;
                        } else {
                            const float gi_ = fsigmoid(in[0][c] + acc[0][c] + bias[0]);
                            const float gf = fsigmoid(in[1][c] + acc[1][c] + bias[1]);
                            const float gg = ftanh(in[2][c] + acc[2][c] + bias[2]);
                            const float go = fsigmoid(in[3][c] + acc[kGates - 1][c] + bias[kGates - 1]);
                            c_ = gf * state_c[c] + gi_ * gg;
                            h_ = go * ftanh(c_);
                            s0 = gi_; s1 = gf; s2 = gg; s3 = go;
                        }
                    }
                    hn[c] = h_; cn[c] = c_;
                    sv[0][c] = s0; sv[1][c] = s1; sv[2][c] = s2; sv[3][c] = s3;
                    state_h[c] = h_;
                    state_c[c] = c_;
                }
                // (1) the next step's operand: straight to L2, visible to the other CTAs as it lands
#pragma unroll
                for (int c = 0; c < kCells; ++c)
                    if (cellok[c]) st_relaxed_bf16(p.hbf + (((size_t)dir * (T + 2) + t + 1) * B + row[c]) * p.Hp + unit, hn[c]);
                if (hl == 0) ASRB_TRACE(7, s);
                // (2) the stores nobody waits for
#pragma unroll
                for (int c = 0; c < kCells; ++c) {
                    if (cellok[c]) {
                        const size_t o = ((size_t)dir * (T + 2) + t + 1) * slotHB + (size_t)row[c] * H + unit;
                        float* svp = p.saved + ((((size_t)dir * T + t) * P + pidx) * 4) * (size_t)(NV * B * 4) +
                                     (size_t)ug * (B * 4) + (size_t)row[c] * 4 + ul;
                        p.hseq[o] = hn[c];
                        if constexpr (CELL == ASRB_RNN_LSTM) p.cseq[o] = cn[c];
#pragma unroll
                        for (int q = 0; q < 4; ++q) svp[(size_t)q * NV * (B * 4)] = sv[q][c];
                    }
                }
                if (hl == 0) ASRB_TRACE(8, s);
            } else {
                float dg[4][kCells], eg2[kCells];
#pragma unroll
                for (int c = 0; c < kCells; ++c) {
                    const float carry = acc[0][c] + state_h[c];
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, e2 = 0.f;
                    if (active[c]) {
                        const float dh = carry + in[4][c];
                        if constexpr (CELL == ASRB_RNN_GRU) {
                            const float r = in[0][c], z = in[1][c], n = in[2][c], gn = in[3][c], hp = in[5][c];
                            const float dn = dh * (1.f - z) * (1.f - n * n);
                            d2 = dn;                          // d gi_n
                            e2 = dn * r;                      // d gh_n
                            d1 = dh * (hp - n) * z * (1.f - z);
                            d0 = dn * gn * r * (1.f - r);
                            state_h[c] = dh * z;
                        } else {
                            const float gi_ = in[0][c], gf = in[1][c], gg = in[2][c], go = in[3][c], cp = in[5][c];
                            const float tcv = ftanh(ct[c]);
                            const float dc = state_c[c] + dh * go * (1.f - tcv * tcv);
                            d0 = dc * gg * gi_ * (1.f - gi_);
                            d1 = dc * cp * gf * (1.f - gf);
                            d2 = dc * gi_ * (1.f - gg * gg);
                            d3 = dh * tcv * go * (1.f - go);
                            e2 = d2;
                            state_c[c] = dc * gf;
                            state_h[c] = 0.f;
                        }
                    } else {
                        state_h[c] = carry;  // gradient passes an inactive step untouched
                    }
                    dg[0][c] = d0; dg[1][c] = d1; dg[2][c] = d2; dg[3][c] = d3; eg2[c] = e2;
                }
#pragma unroll
                for (int c = 0; c < kCells; ++c) {   // (1) next step's operand
                    if (cellok[c]) {
                        __nv_bfloat16* o = p.dghbf + (((size_t)dir * T + t) * B + row[c]) * p.Gp + unit;
#pragma unroll
                        for (int q = 0; q < kGates; ++q) st_relaxed_bf16(o + (size_t)q * H, (q == 2) ? eg2[c] : dg[q][c]);
                    }
                }
                if (hl == 0) ASRB_TRACE(7, s);
#pragma unroll
                for (int c = 0; c < kCells; ++c) {   // (2) outputs only later kernels read
                    if (cellok[c]) {
                        __nv_bfloat16* dgi = reinterpret_cast<__nv_bfloat16*>(p.dgi) + (((size_t)t * B + row[c]) * 2 + dir) * G + unit;
                        // transposed copies for the weight-gradient GEMMs (row = gate row, column = t*B + b)
                        __nv_bfloat16* gT = reinterpret_cast<__nv_bfloat16*>(p.dgiT) + ((size_t)dir * G + unit) * p.ldT + (size_t)t * B + row[c];
                        __nv_bfloat16* hT = p.dghT ? reinterpret_cast<__nv_bfloat16*>(p.dghT) + ((size_t)dir * G + unit) * p.ldT + (size_t)t * B + row[c] : nullptr;
#pragma unroll
                        for (int q = 0; q < kGates; ++q) {
                            const __nv_bfloat16 v = __float2bfloat16_rn(dg[q][c]);
                            dgi[(size_t)q * H] = v;
                            gT[(size_t)q * H * p.ldT] = v;
                            if (hT) hT[(size_t)q * H * p.ldT] = __float2bfloat16_rn((q == 2) ? eg2[c] : dg[q][c]);
                        }
                    }
                }
                if (hl == 0) ASRB_TRACE(8, s);
            }
            if (hl == 0 && p.trace) p.trace[((size_t)blockIdx.x * p.T + s) * 16 + 9] = *reinterpret_cast<volatile uint32_t*>(patch_count);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
    if constexpr (KSPLIT) cluster_sync_all();   // nobody leaves while the peer may still write into its exchange buffer
}

template <int CELL, int NJ, bool BWD, int MROWS, int KS = 1>
static int rnn2_launch(const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    using S = RnnShape<CELL, NJ>;
    constexpr bool KSPLIT = KS > 1;
    constexpr int NPAD = BWD ? (KSPLIT ? KS * NJ : S::kNpadB) : S::kNpadF;
    const int Pk = (BWD && KSPLIT) ? pl.P_b : pl.P;
    prm.P_saved = pl.P;
    prm.P = Pk;
    const int kpad = BWD ? pl.kpad_b : pl.kpad_f;
    const int nkb = kpad / 64;
    // shared memory: weights + barrier block + K-split exchange buffers + as many K-block slots of the operand as fit
    const size_t fixed = 1024 + kX2BarBytes;
    const size_t xb = KSPLIT ? (size_t)2 * (2 * KS - 1) * MROWS * NJ * 4 : 0;
    const size_t wbytes = (size_t)NPAD * kpad * 2;
    if (wbytes + fixed + xb + (size_t)MROWS * 128 > (size_t)kRnnMaxSmem) return ASRB_ERR_UNSUPPORTED;
    int slots = (int)(((size_t)kRnnMaxSmem - fixed - xb - wbytes) / ((size_t)MROWS * 128));
    if (slots > nkb) slots = nkb;
    if (slots > kRnnMaxStages) slots = kRnnMaxStages;
    prm.kpad = kpad;
    prm.stages = slots;
    prm.chunk = 1;
    const size_t smem = wbytes + fixed + xb + (size_t)slots * MROWS * 128;
    CUtensorMap tmW, tmA;
    {
        uint64_t d[2] = {(uint64_t)kpad, (uint64_t)2 * Pk * NPAD}, s[1] = {(uint64_t)kpad * 2};
        uint32_t bx[2] = {64, (uint32_t)NPAD};
        int rc = make_tmap_bf16(&tmW, wpack, 2, d, s, bx);
        if (rc) return rc;
    }
    {
        const int K = BWD ? prm.G : prm.H;
        const int slabs = BWD ? 2 * prm.T : 2 * (prm.T + 2);
        const uint64_t pitch = (uint64_t)(BWD ? prm.Gp : prm.Hp);     // columns >= K / rows >= B: TMA zero fill
        uint64_t d[3] = {(uint64_t)K, (uint64_t)prm.B, (uint64_t)slabs};
        uint64_t s[2] = {pitch * 2, (uint64_t)prm.B * pitch * 2};
        uint32_t bx[3] = {64, (uint32_t)MROWS, 1};
        int rc = make_tmap_bf16(&tmA, BWD ? (const void*)prm.dghbf : (const void*)prm.hbf, 3, d, s, bx);
        if (rc) return rc;
    }
    auto kern = rnn_rec2_kernel<CELL, NJ, BWD, MROWS, KS>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // every CTA of the grid spins on data the others produce: all of them have to be resident at once
    int dev = 0, sms = 0, per_sm = 0;
    ASRB_CUDA_OK(cudaGetDevice(&dev));
    ASRB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ASRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRnnThreads, smem));
    if (2 * Pk > sms * per_sm) return ASRB_ERR_UNSUPPORTED;
    // the fill pattern of the exchange buffer: the whole buffer, one slot per time step
    {
        void* xbuf = BWD ? (void*)prm.dghbf : (void*)prm.hbf;
        const size_t bytes = (size_t)2 * (BWD ? prm.T : prm.T + 2) * prm.B * (BWD ? prm.Gp : prm.Hp) * 2;
        ASRB_CUDA_OK(cudaMemsetAsync(xbuf, 0xFF, bytes, stream));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * Pk);
    cfg.blockDim = dim3(kRnnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = KS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = KSPLIT ? 1 : 0;
    prm.dbg = g_rnn_dbg;
    ASRB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmW, tmA, prm));
    return 0;
}

template <bool BWD>
static int rnn2_dispatch_t(int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
#define ASRB_RNN2_CASE(C, N)                                                                             \
    if (cell == C && pl.nj == N) {                                                                       \
        if constexpr (BWD && N == 16) {                                                                  \
            if (pl.ksplit == 2) {                                                                        \
                if (pl.mrows == 64) return rnn2_launch<C, N, true, 64, 2>(pl, prm, wpack, stream);       \
                return rnn2_launch<C, N, true, 128, 2>(pl, prm, wpack, stream);                          \
            }                                                                                            \
            if (pl.ksplit == 4) {                                                                        \
                if (pl.mrows == 64) return rnn2_launch<C, N, true, 64, 4>(pl, prm, wpack, stream);       \
                return rnn2_launch<C, N, true, 128, 4>(pl, prm, wpack, stream);                          \
            }                                                                                            \
        }                                                                                                \
        if (pl.mrows == 64) return rnn2_launch<C, N, BWD, 64>(pl, prm, wpack, stream);                   \
        return rnn2_launch<C, N, BWD, 128>(pl, prm, wpack, stream);                                      \
    }
    ASRB_RNN2_CASE(ASRB_RNN_GRU, 8) ASRB_RNN2_CASE(ASRB_RNN_GRU, 12) ASRB_RNN2_CASE(ASRB_RNN_GRU, 16)
    ASRB_RNN2_CASE(ASRB_RNN_LSTM, 8) ASRB_RNN2_CASE(ASRB_RNN_LSTM, 12) ASRB_RNN2_CASE(ASRB_RNN_LSTM, 16)
#undef ASRB_RNN2_CASE
    return ASRB_ERR_UNSUPPORTED;
}

int rnn2_dispatch(bool bwd, int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    return bwd ? rnn2_dispatch_t<true>(cell, pl, prm, wpack, stream) : rnn2_dispatch_t<false>(cell, pl, prm, wpack, stream);
}

}  // namespace asrb
