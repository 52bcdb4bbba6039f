// asr_b200 -- persistent bidirectional GRU / LSTM recurrence, bf16 product path: "the data is the flag".
//
// Same decomposition as rnn.cu (CTA (dir, p) owns NJ hidden units, its W_hh slice is resident in shared memory, one
// launch covers all T steps of both directions; modules/blocks.py:87-89 of the reference), but the step-to-step
// hand-over between the CTAs has no counter, no release/acquire pair and no fence on the critical path:
//
//   * the bf16 exchange buffer (one slot per time step, never reused within a launch) is filled with the bit pattern
//     0xFFFF before the launch -- a bf16 NaN the producers can never write: h lies in (-1, 1), and cvt.rn.bf16 turns a
//     NaN gradient into the canonical 0x7FFF;
//   * a producer thread stores each value (2 bytes, st.relaxed.gpu: straight to L2) as soon as it has it; the slot is
//     row-major [B][Kp] like in rnn.cu and reaches shared memory through the same 128-byte-swizzle TMA boxes.  (A
//     "chunk-major" slot [K/8][B][8] = the canonical K-major layout WITHOUT swizzle, filled by plain bulk copies, was
//     built and measured: the stores shrink to one line per warp instruction, but tcgen05.mma reads an unswizzled
//     M=64 operand about five times slower -- 140 instead of 28 cycles per K=16 step -- which costs far more.)
//   * the TMA warp of every CTA polls a few canary words of the slot (the last units each producer writes) and then
//     streams the slot; the MMAs run K block by K block as the boxes land;
//   * a value the copy overtook is still 0xFFFF = NaN, and NaN * w = NaN: the accumulator row of that batch row comes out
//     NaN in every column.  The epilogue threads test the accumulator values they read anyway, OR the result over the CTA
//     in the barrier they need anyway, and the step's copy + product is simply repeated (measured: a handful of times per
//     launch).  A NaN that is really in the data ends the retries after kX2MaxTries and propagates like in the reference.
//
// Every 16-bit word is individually either "not yet written" or final, so correctness needs no ordering between
// different words -- which is exactly what made the counter protocol slow (rnn.cu, DESIGN.md section 6: MEMBAR.ALL.GPU
// before the release, ~2 k cycles release -> acquire visibility, proxy fence: ~4 k of the 10 k cycles of a step).
//
// Second rule of this kernel: the load/store unit is kept free for the hand-over.  Scattered global accesses queue in
// order in the LSU (8..16 cycles per warp instruction that touches 8..16 lines) and every mbarrier / shared-memory
// operation queues behind them, so: the input-side gate pre-activations arrive by TMA (64-byte-swizzle boxes, a step
// ahead), the operand stores touch ONE line per warp instruction, and the stores nobody waits for (fp32 state, saved
// activations, gate gradients) are held in registers until the next step's copies have been issued.
//
// The recurrent product itself (tcgen05.mma kind::f16, M = batch rows, N = the slice's gate columns, fp32 accumulate in
// TMEM), the cell math, the saved-activation layout and the backward K split over CTA pairs (partial sums exchanged
// through distributed shared memory) are those of rnn.cu.
#include "rnn.cuh"

namespace asrb {

// polling load: always served by L2 (measured per L2 hit: ld.volatile 291 cycles, ld.relaxed.gpu 478; tools/ubench/pingpong.cu)
__device__ __forceinline__ uint32_t ld_volatile_u32(const void* ptr) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_bf16(__nv_bfloat16* dst, float v) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(dst), "h"(*reinterpret_cast<const uint16_t*>(&b)) : "memory");
}
// does one of the two 16-bit halves of w still hold the fill pattern 0xFFFF?  ("has a zero half" of ~w)
__device__ __forceinline__ bool half_unwritten(uint32_t w) {
    const uint32_t x = ~w;
    return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u;
}
// CTA-wide OR of a predicate inside a named barrier
__device__ __forceinline__ bool bar_red_or(int id, int nthreads, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.or.pred q, %2, %3, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out) : "r"((uint32_t)pred), "r"(id), "r"(nthreads) : "memory");
    return out != 0;
}
// barrier / small-state block of this kernel (after the A ring)
constexpr int kX2BarBytes = 1536;
constexpr int kX2MiscOff = 2 * kRnnMaxStages;          // in uint64 units, after full_bar[] and empty_bar[]
constexpr int kX2BiasOffset = 1024;                    // bytes: [kGates][NJ] floats
constexpr int kX2MaxTries = 8;
constexpr int kX2MaxCanary = 4;                        // canary words per lane of the TMA warp

// trace slots of this kernel (asrb_debug_rnn_trace): 4 step top, 10 copies announced (go), 0 canary words valid (TMA
// warp), 1 copies issued, 2 MMA warp: first K block landed, 3 MMA warp: last commit issued, 5 accumulator complete,
// 6 accumulator in registers + verdict, 11 exchange done (K split), 7 operand stored, 8 deferred stores issued,
// 9 cumulative number of repeated steps of this CTA
// TS (forward, batch <= 64, K <= 896): the weight slice lives in TENSOR memory as the A operand (M = 64 gate rows, loaded
// once), the batch is the N dimension and comes from shared memory as the B operand.  The recurrent product of rnn.cu is
// bound by shared-memory bandwidth (per step the previous state is written by TMA, read back by the tensor core, and the
// resident weight slice is read as well: 281 KB at ~96 B/clk); with the weights in TMEM the tensor core reads only the
// state.  The accumulator comes out transposed (lane = gate row, column = batch row) and goes through a small
// shared-memory tile to the (batch row, unit) threads of the cell math.
template <int CELL, int NJ, bool BWD, int MROWS, int KS, bool TS>
__global__ void __launch_bounds__(kRnnThreads, 1)
rnn_rec2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ CUtensorMap tmGi, const RnnParams p) {
    static_assert(KS == 1 || BWD, "the K split exists for the backward recurrence only");
    static_assert(!TS || (!BWD && MROWS == 64 && NJ == 16), "TS: forward, one M=64 tile of gate rows, batch <= 64");
    constexpr bool KSPLIT = KS > 1;
    using S = RnnShape<CELL, NJ>;
    constexpr int kGates = S::kGates;
    constexpr int NPAD = BWD ? (KSPLIT ? KS * NJ : S::kNpadB) : S::kNpadF;
    constexpr int kTmemCols = TS ? 512 : 64;
    constexpr int kDtStride = 72;                // floats per gate row of the transposed accumulator tile (TS)
    constexpr int KBE = 64;                      // bf16 elements per K block (8 chunks of 8)
    constexpr int kRowsPerWarp = MROWS / 4;      // TMEM lane quarter -> batch rows (M=64: 16 lanes of each quarter)
    constexpr int kCells = MROWS / 32;           // (batch row, hidden unit) cells of an epilogue thread

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int T = p.T, B = p.B, H = p.H, G = p.G, P = p.P;
    const int nkb = p.kpad / KBE;
    const int NS = p.stages;                                  // K-block slots of the A ring (NS == nkb: no reuse inside a step)
    const bool ring = NS < nkb;
    constexpr uint32_t slot_bytes = MROWS * 128;              // one K block of the A tile: MROWS rows x 128 B (swizzled)
    const int gi_region = BWD ? 0 : round_up(B * 64, 512);    // one gate's [B rows x NJ floats] box (64-byte swizzle atom: 512 B)
    uint8_t* smem_w = smem;                                   // nkb x [NPAD rows x 128 B], 128-byte swizzle (TMA); TS: the
                                                              // transposed accumulator tile [64 gate rows][kDtStride] instead
    uint8_t* smem_gi = smem_w + (TS ? (size_t)64 * kDtStride * 4 : (size_t)nkb * NPAD * 128);   // forward: [2][kGates][gi_region]
    uint8_t* smem_a = smem_gi + (size_t)2 * kGates * gi_region;   // NS x [MROWS rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)NS * slot_bytes);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kX2BiasOffset);   // [kGates][NJ] (forward)
    uint64_t* full_bar = bars;                    // the bulk copy of the K block has landed -> MMA warp
    uint64_t* empty_bar = bars + kRnnMaxStages;   // MMAs that read the slot have completed -> TMA warp (ring only)
    uint64_t* w_bar = bars + kX2MiscOff;
    uint64_t* tfull_bar = w_bar + 1;              // accumulator complete -> epilogue
    uint64_t* verdict_bar = w_bar + 2;            // epilogue: step accepted / to be repeated -> TMA and MMA warps
    uint64_t* epi_bar = w_bar + 3;                // the epilogue warps have stored the step's operand -> TMA warp
    uint64_t* go_bar = w_bar + 4;                 // the step's copies are being issued -> epilogue (deferred stores may go)
    uint64_t* gi_bar = w_bar + 5;                 // [2] forward: the gate pre-activations of step parity 0 / 1 have landed
    uint64_t* x_bar = w_bar + 7;                  // [2] KSPLIT: the peer's partial sums of parity 0 / 1 have arrived
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 9);
    volatile uint32_t* verdict = tmem_slot + 1;   // 1 = repeat the step
    volatile uint32_t* retry_count = tmem_slot + 2;
    float* xbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kX2BarBytes);   // [2][KS sources][MROWS][NJ]
    float* xstage = xbuf + 2 * KS * MROWS * NJ;   // [2][KS-1 peers][MROWS][NJ]: our partial sums of each peer's units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.x / P, pidx = blockIdx.x % P;
    const int j0 = pidx * NJ;
    const uint32_t crank = KSPLIT ? cluster_ctarank() : 0u;      // = pidx % KS: which part of K this CTA multiplies
    const int K = BWD ? G : H;                                   // columns of the exchanged operand
    const int pitch = BWD ? p.Gp : p.Hp;                         // row pitch of a slab (K rounded up to 64)
    const int kcol0 = KSPLIT ? (int)crank * p.kpad : 0;          // first operand column of this CTA's K part
    const int kcols = min(p.kpad, K - kcol0);                    // operand columns of this CTA's K part (may be <= 0)
    const __nv_bfloat16* xchg = BWD ? p.dghbf : p.hbf;           // [slabs][B][pitch]
    const size_t slab_elems = (size_t)B * pitch;
    auto t_of = [&](int s) { return (BWD ? (dir == 0) : (dir == 1)) ? (T - 1 - s) : s; };
    // slab (time slot of the exchange buffer) written at sequential step s
    auto slab_of = [&](int s) {
        const int t = t_of(s);
        return BWD ? (dir * T + t) : (dir * (T + 2) + t + 1);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmA);
        if (!BWD) tma_prefetch_desc(&tmGi);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(w_bar, 1);
        mbar_init(tfull_bar, 1);
        mbar_init(verdict_bar, 1);
        mbar_init(epi_bar, kRnnEpiWarps);
        mbar_init(go_bar, 1);
        mbar_init(&gi_bar[0], 1);
        mbar_init(&gi_bar[1], 1);
        if (KSPLIT) {
            mbar_init(&x_bar[0], 1);   // armed by one local thread with the byte count the peer will send
            mbar_init(&x_bar[1], 1);
        }
        *verdict = 0u;
        *retry_count = 0u;
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if constexpr (KSPLIT) cluster_sync_all();   // the peer's exchange barriers exist before anybody arrives on them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA warp =====================
        auto load_gi = [&](int s) {   // forward: the slice's gate pre-activations of step s -> buffer s & 1
            const int t = t_of(s);
            mbar_arrive_expect_tx(&gi_bar[s & 1], (uint32_t)(kGates * B * NJ * 4));
#pragma unroll
            for (int q = 0; q < kGates; ++q)
                tma_load_3d(smem_gi + (size_t)((s & 1) * kGates + q) * gi_region, &tmGi, &gi_bar[s & 1],
                            dir * G + q * H + j0, 0, t);
        };
        if (elect_one()) {
            if (!TS) {
                mbar_arrive_expect_tx(w_bar, (uint32_t)(nkb * NPAD * 128));
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(smem_w + (size_t)kb * NPAD * 128, &tmW, w_bar, kb * KBE, (dir * P + pidx) * NPAD);
            }
            if (!BWD) {
                load_gi(0);
                if (T > 1) load_gi(1);
            }
        }
        __syncwarp();
        // canary words: the last two units of every 16-column group of our K part, last batch row -- a heuristic for
        // "the producers have stored this step" (the NaN test of the epilogue is what guarantees it)
        const int ncan = kcols > 0 ? min(kcols / 16, 32 * kX2MaxCanary) : 0;
        int slot = 0;
        uint32_t use = 0, att = 0;
        for (int s = 1; s < T; ++s) {
            // our own epilogue has stored step s-1: its reads of the gate buffer and our MMAs of that step are done
            mbar_wait(epi_bar, (uint32_t)((s - 1) & 1));
            const int slab = slab_of(s - 1);
            // canary of 16-column group i: units 12, 13 of the group, last batch row; all loads of a round in flight at once
            const __nv_bfloat16* can = xchg + (size_t)slab * slab_elems + (size_t)(B - 1) * pitch + kcol0 + 12;
            for (;;) {
                uint32_t w[kX2MaxCanary];
#pragma unroll
                for (int i = 0; i < kX2MaxCanary; ++i) {
                    const int ci = lane + 32 * i;
                    w[i] = ci < ncan ? ld_volatile_u32(can + 16 * ci) : 0u;
                }
                bool ok = true;
#pragma unroll
                for (int i = 0; i < kX2MaxCanary; ++i) ok = ok && !half_unwritten(w[i]);
                if (__all_sync(0xffffffffu, ok)) break;
            }
            if (lane == 0) {
                ASRB_TRACE(0, s);
                mbar_arrive(go_bar);
            }
            for (;;) {
                for (int kb = 0; kb < nkb; ++kb) {
                    if (ring && use > 0) mbar_wait(&empty_bar[slot], (use - 1) & 1);
                    if (elect_one()) {      // rows >= B and columns >= K of the box: TMA zero fill
                        mbar_arrive_expect_tx(&full_bar[slot], slot_bytes);
                        tma_load_3d(smem_a + (size_t)slot * slot_bytes, &tmA, &full_bar[slot], kcol0 + kb * KBE, 0, slab);
                    }
                    __syncwarp();
                    if (++slot == NS) { slot = 0; ++use; }
                }
                if (lane == 0) ASRB_TRACE(1, s);
                mbar_wait(verdict_bar, att & 1);
                ++att;
                if (*verdict == 0u) break;
            }
            if (!BWD && s + 1 < T) {
                if (elect_one()) load_gi(s + 1);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        // TS: M = 64 gate rows (A, tensor memory), N = MROWS batch rows (B, shared memory)
        constexpr uint32_t idesc = TS ? umma_idesc(kFmtBF16, 64, MROWS) : umma_idesc(kFmtBF16, MROWS, NPAD);
        mbar_wait(w_bar, 0);
        tc_fence_after_sync();
        int slot = 0;
        uint32_t phase = 0, att = 0;
        for (int s = 1; s < T; ++s) {
            for (;;) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[slot], phase);
                    if (kb == 0 && lane == 0) ASRB_TRACE(2, s);
                    tc_fence_after_sync();
                    if (elect_one()) {
                        const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + (size_t)slot * slot_bytes));
                        if constexpr (TS) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // weights: 8 TMEM columns per K = 16 step, after the accumulator
                                umma_f16_ts(tmem_base, tmem_base + 64 + (kb * 4 + k) * 8, adesc + 2 * k, idesc, (kb | k) != 0);
                        } else {
                            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_w + (size_t)kb * NPAD * 128));
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // 4 x 32-byte K slices (K = 16 bf16) per 128-byte block
                                umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        if (ring) umma_commit(&empty_bar[slot]);
                        if (kb == nkb - 1) umma_commit(tfull_bar);
                    }
                    __syncwarp();
                    if (++slot == NS) { slot = 0; phase ^= 1; }
                }
                if (lane == 0) ASRB_TRACE(3, s);
                mbar_wait(verdict_bar, att & 1);
                ++att;
                if (*verdict == 0u) break;
            }
        }
    } else if (warp >= kRnnCtrlWarps) {
        // ===================== epilogue: 16 warps =====================
        // lane quarter q = warp % 4 (a warp may only read TMEM lanes 32*(warp%4)..+31), unit group ug; the accumulator
        // is read with tcgen05.ld.16x128b so that ALL 32 lanes hold cells also in the M=64 layout (16 TMEM lanes per
        // quarter): thread = hidden unit 4*ug + lane%4 of batch rows lane/4 + 8c of the quarter.
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
        auto xoff = [&](int r, int col) { return (size_t)r * pitch + col; };   // operand element (row, column) of a slab
        // forward: shared-memory offset of (row, unit ju) inside one gate's 64-byte-swizzled box
        uint32_t gi_off[kCells];
#pragma unroll
        for (int c = 0; c < kCells; ++c)
            gi_off[c] = (uint32_t)(row[c] * 64 + ((ug ^ ((row[c] >> 1) & 3)) << 4) + ul * 4);
        const uint32_t smem_gi_u32 = smem_u32(smem_gi);

        // DEBUG (asrb_debug_rnn_dbg bit 16): hold the non-critical stores of a step back until the next step's copies are
        // being issued, instead of issuing them right behind the operand stores
        const bool gated = (p.dbg & 16) != 0;
        float xsend[KS > 1 ? KS - 1 : 1][kCells] = {};
        uint32_t xphase[2] = {0u, 0u};
        uint32_t att = 0;
        float state_h[kCells], state_c[kCells];   // fwd: h / c of the previous step ; bwd: direct dh / dc carries
#pragma unroll
        for (int c = 0; c < kCells; ++c) state_h[c] = state_c[c] = 0.f;
        // results of the previous step whose stores nobody waits for
        constexpr int kPend = BWD ? 5 : 6;        // fwd: h, c, 4 saved ; bwd: 4 dgates + the hidden-side n gradient
        float pend[kPend][kCells];
#pragma unroll
        for (int q = 0; q < kPend; ++q)
#pragma unroll
            for (int c = 0; c < kCells; ++c) pend[q][c] = 0.f;

        auto deferred_stores = [&](int t) {
            if (p.dbg & 1) return;   // DEBUG: timing experiment without the stores nobody waits for (results incomplete)
#pragma unroll
            for (int c = 0; c < kCells; ++c) {
                if (!cellok[c]) continue;
                if constexpr (!BWD) {
                    const size_t o = ((size_t)dir * (T + 2) + t + 1) * slotHB + (size_t)row[c] * H + unit;
                    float* svp = p.saved + ((((size_t)dir * T + t) * P + pidx) * 4) * (size_t)(NV * B * 4) +
                                 (size_t)ug * (B * 4) + (size_t)row[c] * 4 + ul;
                    p.hseq[o] = pend[0][c];
                    if constexpr (CELL == ASRB_RNN_LSTM) p.cseq[o] = pend[1][c];
#pragma unroll
                    for (int q = 0; q < 4; ++q) svp[(size_t)q * NV * (B * 4)] = pend[2 + q][c];
                } else {
                    __nv_bfloat16* dgi = reinterpret_cast<__nv_bfloat16*>(p.dgi) + (((size_t)t * B + row[c]) * 2 + dir) * G + unit;
                    // transposed copies for the weight-gradient GEMMs (row = gate row, column = t*B + b)
                    __nv_bfloat16* gT = reinterpret_cast<__nv_bfloat16*>(p.dgiT) + ((size_t)dir * G + unit) * p.ldT + (size_t)t * B + row[c];
                    __nv_bfloat16* hT = p.dghT ? reinterpret_cast<__nv_bfloat16*>(p.dghT) + ((size_t)dir * G + unit) * p.ldT + (size_t)t * B + row[c] : nullptr;
#pragma unroll
                    for (int q = 0; q < kGates; ++q) {
                        const __nv_bfloat16 v = __float2bfloat16_rn(pend[q][c]);
                        dgi[(size_t)q * H] = v;
                        gT[(size_t)q * H * p.ldT] = v;
                        if (hT) hT[(size_t)q * H * p.ldT] = (q == 2) ? __float2bfloat16_rn(pend[4][c]) : v;
                    }
                }
            }
        };

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
            if constexpr (TS) {
                // the weight slice -> tensor memory, once: gate row c = 16 * quarter + i sits in TMEM lane 32 * quarter + i
                // (the M = 64 data path layout, like the accumulator rows); 16 bf16 = 8 columns per store, after the 64
                // accumulator columns.  DEBUG (dbg bit 32): rows 0..63 in lanes 0..63 instead.
                if (wq < 4) {
                    const bool alt = (p.dbg & 32) != 0;
                    const int c = alt ? (quad * 32 + lane) : (16 * quad + lane);
                    const bool have = (alt ? quad < 2 : lane < 16) && c < NPAD;
                    const uint4* wrow = reinterpret_cast<const uint4*>(
                        reinterpret_cast<const __nv_bfloat16*>(p.wpack) + ((size_t)(dir * P + pidx) * NPAD + (have ? c : 0)) * p.kpad);
                    for (int k16 = 0; k16 < p.kpad / 16; ++k16) {
                        uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
                        if (have) { lo = __ldg(wrow + 2 * k16); hi = __ldg(wrow + 2 * k16 + 1); }
                        const uint32_t r[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                        tmem_st_32x8(tmem_base + (uint32_t(quad * 32) << 16) + 64 + k16 * 8, r);
                    }
                    tmem_st_wait();
                }
                tc_fence_before_sync();
            }
            named_bar_sync(3, kRnnEpiThreads);
            if (TS && hl == 0) mbar_arrive(w_bar);
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
            // the previous step's operand looks complete and its copies are on their way: the LSU is ours until they land
            if (s > 0 && gated) mbar_wait(go_bar, (uint32_t)((s - 1) & 1));
            if (hl == 0) ASRB_TRACE(10, s);
            if constexpr (KSPLIT) {
                if (hl == 0 && s > 0) mbar_arrive_expect_tx(&x_bar[s & 1], (uint32_t)((KS - 1) * MROWS * NJ * 4));
            }
            constexpr int kAccG = BWD ? 1 : kGates;
            float acc[kAccG][kCells];
#pragma unroll
            for (int g = 0; g < kAccG; ++g)
#pragma unroll
                for (int c = 0; c < kCells; ++c) acc[g][c] = 0.f;

            // ---- backward: operand prefetch from global memory (independent of the recurrent product) ----
            constexpr int kIn = BWD ? 6 : kGates;        // fwd: gi gates ; bwd: 4 saved + dout + previous state
            float in[kIn][kCells];
            float ct[kCells];
#pragma unroll
            for (int c = 0; c < kCells; ++c) {
#pragma unroll
                for (int q = 0; q < kIn; ++q) in[q][c] = 0.f;
                ct[c] = 0.f;
                if constexpr (BWD) {
                    if (active[c]) {
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
            // ---- the stores of the previous step that nobody waits for ----
            if (s > 0 && gated) deferred_stores(t_of(s - 1));
            if (hl == 0 && gated) ASRB_TRACE(8, s);

            if (s > 0) {
                // ---- recurrent product for this step: repeated while a row of the accumulator reads NaN ----
                for (int tries = 1;; ++tries) {
                    mbar_wait(tfull_bar, att & 1);
                    ++att;
                    if (hl == 0) ASRB_TRACE(5, s);
                    tc_fence_after_sync();
                    bool bad = false;
                    if constexpr (TS) {
                        // lane quarter `quad` holds gate `quad` of the slice's 16 units (rows), columns = batch rows; this
                        // warp takes columns 16*ug .. +15 and passes them on through the transposed tile
                        if (quad < kGates) {
                            float v[8];
                            tmem_ld_16x256b_x2(tmem_base + (uint32_t(quad * 32) << 16) + 16 * ug, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) bad = bad || (v[i] != v[i]);
                            float* dt = reinterpret_cast<float*>(smem_w) + (quad * 16 + (lane >> 2)) * kDtStride + 16 * ug + 2 * (lane & 3);
                            *reinterpret_cast<float2*>(dt) = make_float2(v[0], v[1]);
                            *reinterpret_cast<float2*>(dt + 8 * kDtStride) = make_float2(v[2], v[3]);
                            *reinterpret_cast<float2*>(dt + 8) = make_float2(v[4], v[5]);
                            *reinterpret_cast<float2*>(dt + 8 * kDtStride + 8) = make_float2(v[6], v[7]);
                        }
                    } else if (warp_ld) {
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
#pragma unroll
                        for (int c = 0; c < kCells; ++c) {
                            if (row[c] < B) {
#pragma unroll
                                for (int g = 0; g < kAccG; ++g) bad = bad || (acc[g][c] != acc[g][c]);
                                if constexpr (KSPLIT) {
#pragma unroll
                                    for (int q = 1; q < KS; ++q) bad = bad || (xsend[q - 1][c] != xsend[q - 1][c]);
                                }
                            }
                        }
                    }
                    tc_fence_before_sync();   // the verdict arrival orders these reads before the next MMAs
                    const bool again = bar_red_or(2, kRnnEpiThreads, bad) && tries < kX2MaxTries;
                    if (hl == 0) {
                        *verdict = again ? 1u : 0u;
                        if (again) *retry_count = *retry_count + 1u;
                        mbar_arrive(verdict_bar);
                    }
                    if (!again) break;
                }
                if constexpr (TS) {      // the barrier of the verdict has made the tile visible
                    const float* dt = reinterpret_cast<const float*>(smem_w) + (ju < NJ ? ju : 0) * kDtStride;
#pragma unroll
                    for (int g = 0; g < kAccG; ++g)
#pragma unroll
                        for (int c = 0; c < kCells; ++c) acc[g][c] = dt[g * 16 * kDtStride + row[c]];
                }
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
                mbar_wait(&gi_bar[s & 1], (uint32_t)((s >> 1) & 1));
#pragma unroll
                for (int c = 0; c < kCells; ++c) {
                    if (active[c]) {
#pragma unroll
                        for (int q = 0; q < kGates; ++q)
                            in[q][c] = ld_shared_f32(smem_gi_u32 + (uint32_t)(((s & 1) * kGates + q) * gi_region) + gi_off[c]);
                    }
                }
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
                    state_h[c] = h_;
                    state_c[c] = c_;
                    pend[0][c] = h_; pend[1][c] = c_;
                    pend[2][c] = s0; pend[3][c] = s1; pend[4][c] = s2; pend[5][c] = s3;
                }
                // the next step's operand: straight to L2, one 128-byte line per warp instruction
                __nv_bfloat16* slab = p.hbf + (size_t)slab_of(s) * slab_elems;
#pragma unroll
                for (int c = 0; c < kCells; ++c)
                    if (cellok[c]) st_relaxed_bf16(slab + xoff(row[c], unit), state_h[c]);
            } else {
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
                    pend[0][c] = d0; pend[1][c] = d1; pend[2][c] = d2; pend[3][c] = d3; pend[4][c] = e2;
                }
                __nv_bfloat16* slab = p.dghbf + (size_t)slab_of(s) * slab_elems;
#pragma unroll
                for (int c = 0; c < kCells; ++c) {
                    if (cellok[c]) {
#pragma unroll
                        for (int q = 0; q < kGates; ++q)
                            st_relaxed_bf16(slab + xoff(row[c], q * H + unit), (q == 2) ? pend[4][c] : pend[q][c]);
                    }
                }
            }
            if (hl == 0) {
                ASRB_TRACE(7, s);
                if (p.trace) p.trace[((size_t)blockIdx.x * p.T + s) * 16 + 9] = *retry_count;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(epi_bar);
            if (!gated) {      // the stores nobody waits for: they drain while the operand travels
                deferred_stores(t);
                if (hl == 0) ASRB_TRACE(8, s);
            }
        }
        if (gated) deferred_stores(t_of(T - 1));
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
    if constexpr (KSPLIT) cluster_sync_all();   // nobody leaves while the peer may still write into its exchange buffer
}

template <int CELL, int NJ, bool BWD, int MROWS, int KS = 1, bool TS = false>
static int rnn2_launch(const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    using S = RnnShape<CELL, NJ>;
    static_assert(NJ == 16, "the gate boxes are 64-byte rows (64-byte swizzle)");
    constexpr bool KSPLIT = KS > 1;
    constexpr int kGates = S::kGates;
    constexpr int NPAD = BWD ? (KSPLIT ? KS * NJ : S::kNpadB) : S::kNpadF;
    const int Pk = (BWD && KSPLIT) ? pl.P_b : pl.P;
    prm.P_saved = pl.P;
    prm.P = Pk;
    const int kpad = BWD ? pl.kpad_b : pl.kpad_f;
    const int nkb = kpad / 64;
    const int B = prm.B, K = BWD ? prm.G : prm.H;
    // shared memory: weights + gate boxes (forward) + operand ring + overrun slack + barrier block + K-split buffers
    const size_t wbytes = TS ? (size_t)64 * 72 * 4 : (size_t)NPAD * kpad * 2;   // TS: the transposed accumulator tile
    const size_t gibytes = BWD ? 0 : (size_t)2 * kGates * round_up(B * 64, 512);
    const size_t xb = KSPLIT ? (size_t)2 * (2 * KS - 1) * MROWS * NJ * 4 : 0;
    const size_t fixed = 1024 + wbytes + gibytes + kX2BarBytes + xb;
    const size_t slot_bytes = (size_t)MROWS * 128;
    if (fixed + slot_bytes > (size_t)kRnnMaxSmem) return ASRB_ERR_UNSUPPORTED;
    int slots = (int)(((size_t)kRnnMaxSmem - fixed) / slot_bytes);
    if (slots > nkb) slots = nkb;
    if (slots > kRnnMaxStages) slots = kRnnMaxStages;
    prm.kpad = kpad;
    prm.stages = slots;
    prm.chunk = 1;
    const size_t smem = fixed + (size_t)slots * slot_bytes;
    CUtensorMap tmW, tmA, tmGi;
    {
        uint64_t d[2] = {(uint64_t)kpad, (uint64_t)2 * Pk * NPAD}, s[1] = {(uint64_t)kpad * 2};
        uint32_t bx[2] = {64, (uint32_t)NPAD};
        int rc = make_tmap_bf16(&tmW, wpack, 2, d, s, bx);
        if (rc) return rc;
    }
    {
        const int slabs = BWD ? 2 * prm.T : 2 * (prm.T + 2);
        const uint64_t pitch = (uint64_t)(BWD ? prm.Gp : prm.Hp);     // columns >= K / rows >= B: TMA zero fill
        uint64_t d[3] = {(uint64_t)K, (uint64_t)B, (uint64_t)slabs};
        uint64_t s[2] = {pitch * 2, (uint64_t)B * pitch * 2};
        uint32_t bx[3] = {64, (uint32_t)MROWS, 1};
        int rc = make_tmap_bf16(&tmA, BWD ? (const void*)prm.dghbf : (const void*)prm.hbf, 3, d, s, bx);
        if (rc) return rc;
    }
    if (!BWD) {
        // gi [T][B][2G] fp32, box = [B rows][NJ columns], 64-byte swizzle (conflict-free reads by (row, unit) threads)
        PFN_encodeTiled enc = get_encode_tiled();
        if (!enc) return ASRB_ERR_DRIVER;
        if ((reinterpret_cast<uintptr_t>(prm.gi) & 15) != 0) return ASRB_ERR_ALIGNMENT;
        cuuint64_t gdim[3] = {(cuuint64_t)2 * prm.G, (cuuint64_t)B, (cuuint64_t)prm.T};
        cuuint64_t gstr[2] = {(cuuint64_t)2 * prm.G * 4, (cuuint64_t)B * 2 * prm.G * 4};
        cuuint32_t bx[3] = {(cuuint32_t)NJ, (cuuint32_t)B, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&tmGi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(prm.gi), gdim, gstr, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return ASRB_ERR_TENSORMAP;
    } else {
        tmGi = tmW;
    }
    auto kern = rnn_rec2_kernel<CELL, NJ, BWD, MROWS, KS, TS>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // every CTA of the grid spins on data the others produce: all of them have to be resident at once
    int dev = 0, sms = 0, per_sm = 0;
    ASRB_CUDA_OK(cudaGetDevice(&dev));
    ASRB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ASRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRnnThreads, smem));
    if (2 * Pk > sms * per_sm) return ASRB_ERR_UNSUPPORTED;
    // the fill pattern of the exchange buffer: one slot per time step
    {
        void* xbuf = BWD ? (void*)prm.dghbf : (void*)prm.hbf;
        const size_t bytes = (size_t)2 * (BWD ? prm.T : prm.T + 2) * B * (BWD ? prm.Gp : prm.Hp) * 2;
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
    prm.wpack = reinterpret_cast<const float*>(wpack);   // TS: read directly (bf16) when the slice goes to tensor memory
    ASRB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmW, tmA, tmGi, prm));
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
        if constexpr (!BWD) {                                                                            \
            /* weights in tensor memory: 64 accumulator columns + kpad/2 weight columns of the 512 */    \
            if (pl.mrows == 64 && pl.kpad_f <= 896 && !(g_rnn_dbg & 64))                                 \
                return rnn2_launch<C, N, false, 64, 1, true>(pl, prm, wpack, stream);                    \
        }                                                                                                \
        if (pl.mrows == 64) return rnn2_launch<C, N, BWD, 64>(pl, prm, wpack, stream);                   \
        return rnn2_launch<C, N, BWD, 128>(pl, prm, wpack, stream);                                      \
    }
    ASRB_RNN2_CASE(ASRB_RNN_GRU, 16)
    ASRB_RNN2_CASE(ASRB_RNN_LSTM, 16)
#undef ASRB_RNN2_CASE
    return ASRB_ERR_UNSUPPORTED;
}

int rnn2_dispatch(bool bwd, int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    return bwd ? rnn2_dispatch_t<true>(cell, pl, prm, wpack, stream) : rnn2_dispatch_t<false>(cell, pl, prm, wpack, stream);
}

}  // namespace asrb
