// asr_b200 -- persistent bidirectional GRU / LSTM recurrence (forward and backward) for sm_100a.
//
// Replaces the recurrent half of torch.nn.GRU / torch.nn.LSTM over packed sequences as the reference
// calls it (asr_deepspeech/modules/blocks.py:87-89; cell equations: torch.nn docs, gate row order
// [r|z|n] / [i|f|g|o]).  The input half (x W_ih^T + b_ih for all timesteps, both directions) is one big
// tensor-core GEMM done beforehand (gemm.cu); this kernel runs the strictly sequential part.
//
// One launch covers all T steps of both directions.  CTA (dir, p) owns NJ hidden units:
//   * its slice of W_hh (forward: NJ x gates rows of K=H; backward: NJ columns of W_hh, K=gates*H) is loaded
//     ONCE by TMA into shared memory in the tcgen05 K-major/128B-swizzle layout and stays there;
//   * per step it TMA-streams the previous state of ALL units (h_{t-1}: [B,H], or dgates_{t+1}: [B,G]) through
//     an mbarrier ring as the A operand, issues tcgen05.mma (M = 64 or 128 batch rows, N = slice columns) into TMEM,
//     and the 4 epilogue warps (one thread per batch row) apply the gate non-linearities, the sequence-length mask
//     and write state / saved activations;
//   * operand precision of the recurrent product: bf16 (default: kind::f16, whole previous state in flight at once,
//     the step is latency-bound and bf16 halves both the bytes and the shared-memory footprint) or tf32 (fp32 state
//     read directly); accumulation, gate math, stored states and all gradients are fp32 either way;
//   * a per-direction release/acquire counter in global memory is the step barrier between CTAs.
// Packed-sequence semantics (pack_padded_sequence / pad_packed_sequence, blocks.py:87,89): utterance b takes
// part only while t < len[b]; outputs and states at t >= len[b] are written as zeros, which is also the correct
// initial state for the reverse direction.
//
// State layout: hseq/cseq [2][T+2][B][H] with time slot t+1 holding step t and slots 0 / T+1 zero, so that
// "previous step" is a pure pointer offset for both directions (used by the dW_hh GEMM as well).
#include "rnn.cuh"

namespace asrb {

long long* g_rnn_trace = nullptr;
int g_rnn_dbg = 0;
int g_rnn_ksplit = 2;
int g_rnn_chunk = 0;

// ------------------------------------------------------------------------------------------------
// weight packing: one contiguous, zero-padded [npad, kpad] K-major matrix per (direction, CTA)
// ------------------------------------------------------------------------------------------------
// forward : row c = gate (c / nj), unit p*nj + c % nj   ->  W_hh[dir][gate*H + unit][0..H)
__device__ __forceinline__ void pack_store(float* o, float v) { *o = v; }
__device__ __forceinline__ void pack_store(__nv_bfloat16* o, float v) { *o = __float2bfloat16_rn(v); }

// one block per output row: a row of W_hh is copied (and converted) with coalesced reads and writes
template <typename OutT>
__global__ void rnn_pack_fwd_kernel(const float* __restrict__ w_hh0, const float* __restrict__ w_hh1, OutT* __restrict__ out,
                                    int H, int gates, int nj, int P, int npad, int kpad) {
    const int row = blockIdx.x;                       // (dir, p, c)
    const int c = row % npad, p = (row / npad) % P, dir = row / (npad * P);
    const int g = c / nj, j = p * nj + c % nj;
    const bool have = g < gates && j < H;
    const float* src = (dir ? w_hh1 : w_hh0) + (size_t)(have ? g * H + j : 0) * H;
    OutT* dst = out + (size_t)row * kpad;
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) pack_store(dst + k, (have && k < H) ? __ldg(src + k) : 0.f);
}
// backward: out[(dir, p)][c][kk] = W_hh[dir][k0(p) + kk][j0(p) + c]: a transposed block of W_hh per CTA, moved in 32 x 32
// tiles through shared memory (coalesced along the units on the way in, along K on the way out).
//   unsplit: j0 = p*nj, k0 = 0, rows c < nj valid ; split over ks CTAs: j0 = (p/ks)*npad, k0 = (p%ks)*kpad
template <typename OutT>
__global__ void rnn_pack_bwd_tiled_kernel(const float* __restrict__ w_hh0, const float* __restrict__ w_hh1, OutT* __restrict__ out,
                                          int H, int G, int nj, int P, int npad, int kpad, int ks) {
    __shared__ float tile[32][33];
    const int dp = blockIdx.z, p = dp % P, dir = dp / P;
    const int j0 = ks ? (p / ks) * npad : p * nj, k0 = ks ? (p % ks) * kpad : 0;
    const int cvalid = ks ? npad : nj;
    const float* w = dir ? w_hh1 : w_hh0;
    const int kk0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {            // read W[k0 + kk0 + r][j0 + c0 + x]
        const int k = k0 + kk0 + r, c = c0 + threadIdx.x, j = j0 + c;
        tile[r][threadIdx.x] = (kk0 + r < kpad && k < G && c < cvalid && j < H) ? __ldg(w + (size_t)k * H + j) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {            // write out[c0 + r][kk0 + x]
        const int c = c0 + r, kk = kk0 + threadIdx.x;
        if (c < npad && kk < kpad) pack_store(out + ((size_t)dp * npad + c) * kpad + kk, tile[threadIdx.x][r]);
    }
}

// KS > 1 (backward only): a cluster of KS = 2 or 4 CTAs shares KS*NJ hidden units.  Each CTA holds the weights of all
// KS*NJ units for 1/KS of the gate index K and streams only that part of the operand through its shared memory (the
// bandwidth that bounds the backward step: 686 KB per CTA and step at H=800 unsplit, 382 KB with KS=2, 230 KB with
// KS=4); the partial products are exchanged through distributed shared memory: every CTA sends each peer the
// 64 x NJ partial sums of that peer's units and finishes its own NJ units.
template <int CELL, int NJ, bool BWD, bool BF16, int MROWS, int KS>
__global__ void __launch_bounds__(kRnnThreads, 1)
rnn_rec_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const RnnParams p) {
    static_assert(KS == 1 || BWD, "the K split exists for the backward recurrence only");
    constexpr bool KSPLIT = KS > 1;
    using S = RnnShape<CELL, NJ>;
    constexpr int kGates = S::kGates;
    constexpr int NPAD = BWD ? (KSPLIT ? KS * NJ : S::kNpadB) : S::kNpadF;
    constexpr int kTmemCols = 64;
    constexpr int KBE = BF16 ? 64 : 32;          // elements per 128-byte K block
    constexpr int kStageBytes = MROWS * 128;     // one K block of the A tile: MROWS rows x 128 B
    // TMEM lane of tile row m: M=128 -> m ; M=64 -> (m % 16) + 32 * (m / 16)  (16 lanes of every lane quarter)
    constexpr int kRowsPerWarp = MROWS / 4;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkb = p.kpad / KBE;
    uint8_t* smem_w = smem;                                   // nkb x [NPAD rows x 128 B]
    uint8_t* smem_a = smem_w + (size_t)nkb * NPAD * 128;      // stages x chunk x [MROWS rows x 128 B]
    const int stage_bytes = p.chunk * kStageBytes;
    const int nchunks = ceil_div(nkb, p.chunk);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)p.stages * stage_bytes);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kRnnBiasOffset);   // [kGates][NJ] (forward)
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kRnnMaxStages;
    uint64_t* w_bar = bars + 2 * kRnnMaxStages;
    uint64_t* tfull_bar = w_bar + 1;
    uint64_t* tempty_bar = w_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 3);
    uint64_t* x_bar = w_bar + 4;                  // [2] KSPLIT: the peer's partial sums of parity 0 / 1 have arrived
    float* xbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kRnnBarBytes);   // [2][KS sources][MROWS][NJ] (KSPLIT)
    float* xstage = xbuf + 2 * KS * MROWS * NJ;   // [2][KS-1 peers][MROWS][NJ]: our partial sums of each peer's units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, B = p.B, H = p.H, G = p.G, P = p.P;
    const int dir = blockIdx.x / P, pidx = blockIdx.x % P;
    const int j0 = pidx * NJ;
    const bool tc = !p.use_simt;
    uint32_t* counter = p.counters + dir * kRnnCounterStride;
    const uint32_t crank = KSPLIT ? cluster_ctarank() : 0u;      // = pidx % KS: which part of K this CTA multiplies
    const int kb_off = KSPLIT ? (int)crank * nkb : 0;            // first K block (of the global operand) of this CTA
    // time index processed at sequential step s
    auto t_of = [&](int s) { return (BWD ? (dir == 0) : (dir == 1)) ? (T - 1 - s) : s; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmA);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(w_bar, 1);
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, kRnnEpiWarps);
        if (KSPLIT) {
            mbar_init(&x_bar[0], 1);   // armed by one local thread with the byte count the peer will send
            mbar_init(&x_bar[1], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if constexpr (KSPLIT) cluster_sync_all();   // the peer's exchange barriers exist before anybody arrives on them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: ONE elected thread runs the whole loop =====================
        // (per-chunk elect / __syncwarp rounds of the whole warp cost ~250 cycles each on the step's critical path: rnn3.cu)
        if (tc) {
            if (elect_one()) {
                mbar_arrive_expect_tx(w_bar, (uint32_t)(nkb * NPAD * 128));
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(smem_w + (size_t)kb * NPAD * 128, &tmW, w_bar, kb * KBE, (dir * P + pidx) * NPAD);
                int stage = 0;
                uint32_t phase = 0;
                for (int s = 1; s < T; ++s) {
                    // step barrier: every CTA of this direction has published step s-1
                    const uint32_t need = (uint32_t)P * (uint32_t)s;
                    poll_counter(counter, need);
                    ASRB_TRACE(0, s);
                    // other CTAs' generic-proxy stores (acquired above) -> visible to our async-proxy (TMA) reads.
                    // The .global form is a bare FENCE.VIEW.ASYNC.G; the unqualified one adds a MEMBAR.ALL.GPU.
                    fence_proxy_async_global();
                    ASRB_TRACE(9, s);
                    const int tp = t_of(s - 1);
                    const int slab = BWD ? (dir * T + tp) : (dir * (T + 2) + tp + 1);
                    for (int c = 0; c < nchunks; ++c) {
                        const int kb0 = c * p.chunk;
                        const int nblk = min(p.chunk, nkb - kb0);
                        uint8_t* st = smem_a + (size_t)stage * stage_bytes;
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nblk * kStageBytes));
                        for (int i = 0; i < nblk; ++i)
                            tma_load_3d(st + (size_t)i * kStageBytes, &tmA, &full_bar[stage], (kb_off + kb0 + i) * KBE, 0, slab);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    ASRB_TRACE(1, s);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        // Measured (tools/trace_rnn.py, tools/ubench): this phase is bound by shared-memory bandwidth -- every step the
        // whole previous state is written to SMEM by TMA and read back by the tensor core together with the weight slice
        // (fwd 286 KB, bwd 684 KB per CTA per step at ~90 B/clk) -- not by L2 (TMA multicast over clusters of 2/5 CTAs and
        // replicated operands changed nothing), not by the accumulator dependency (4 independent accumulators: same
        // time) and not by the number of MMAs (two half-batch passes cost the same as one).
        if (tc) {
            constexpr uint32_t idesc = umma_idesc(BF16 ? kFmtBF16 : kFmtTF32, MROWS, NPAD);
            if (elect_one()) {               // one elected thread runs the whole loop
                mbar_wait(w_bar, 0);
                int stage = 0;
                uint32_t phase = 0;
                for (int s = 1; s < T; ++s) {
                    const uint32_t it = (uint32_t)(s - 1);
                    mbar_wait(tempty_bar, (it & 1) ^ 1);
                    tc_fence_after_sync();
                    for (int c = 0; c < nchunks; ++c) {
                        const int kb0 = c * p.chunk;
                        const int nblk = min(p.chunk, nkb - kb0);
                        mbar_wait(&full_bar[stage], phase);
                        if (c == 0) ASRB_TRACE(2, s);
                        tc_fence_after_sync();
                        const uint32_t a0 = smem_u32(smem_a + (size_t)stage * stage_bytes);
                        const uint32_t b0 = smem_u32(smem_w + (size_t)kb0 * NPAD * 128);
                        for (int i = 0; i < nblk; ++i) {
                            const uint64_t adesc = umma_desc_sw128(a0 + i * kStageBytes);
                            const uint64_t bdesc = umma_desc_sw128(b0 + i * NPAD * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {   // 4 x 32-byte K slices per 128-byte block (K=8 tf32 / K=16 bf16)
                                if constexpr (BF16) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (c | i | k) != 0);
                                else                umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (c | i | k) != 0);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                        if (c == nchunks - 1) umma_commit(tfull_bar);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    ASRB_TRACE(3, s);
                }
            }
            __syncwarp();
        }
    } else if (warp >= kRnnCtrlWarps) {
        // ===================== epilogue: thread = (batch row, one 4-wide group of the slice's hidden units) ==========
        // 16 warps: lane quarter q = warp % 4 (a warp may only read TMEM lanes 32*(warp%4)..+31), unit group ug.
        constexpr int NV = NJ / 4;               // 4-wide unit groups of the slice: all global traffic is 16-byte vectors
        const int quad = warp & 3;
        const int ug = (warp - kRnnCtrlWarps) >> 2;
        const int b = quad * kRowsPerWarp + lane;
        const int hl = (warp - kRnnCtrlWarps) * 32 + lane;    // epilogue thread index: 0..511
        const bool gvalid = (ug < NV) && (j0 + 4 * ug < H);
        const bool rowok = lane < kRowsPerWarp && b < B && gvalid;
        const bool warp_ld = (quad * kRowsPerWarp < B) && (ug < NV);   // warp-uniform: this warp reads the accumulator
        const int len = rowok ? p.lengths[b] : 0;
        const size_t slotHB = (size_t)B * H;
        const int u0 = 4 * ug;                   // first unit (within the slice) owned by this thread

        float xsend[KS > 1 ? KS - 1 : 1][4] = {};
        uint32_t xphase[2] = {0u, 0u};
        float state_h[4];   // fwd: h_prev of our units ; bwd: direct dh carry
        float state_c[4];   // fwd LSTM: c_prev ; bwd LSTM: dc carry
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) state_h[jj] = state_c[jj] = 0.f;

        if (!BWD && rowok) {  // zero boundary slots 0 and T+1 of our part of the slice
            float z4[4] = {0.f, 0.f, 0.f, 0.f};
            const size_t o = ((size_t)dir * (T + 2)) * slotHB + (size_t)b * H + j0 + u0;
            st4(p.hseq + o, z4);
            st4(p.hseq + o + (size_t)(T + 1) * slotHB, z4);
            if constexpr (BF16) {
                const size_t ob = (((size_t)dir * (T + 2)) * B + b) * p.Hp + j0 + u0;
                st4_bf16(p.hbf + ob, z4);
                st4_bf16(p.hbf + ob + (size_t)(T + 1) * B * p.Hp, z4);
            }
            if constexpr (CELL == ASRB_RNN_LSTM) {
                st4(p.cseq + o, z4);
                st4(p.cseq + o + (size_t)(T + 1) * slotHB, z4);
            }
        }
        if constexpr (!BWD) {                    // recurrent biases of the slice -> shared memory (broadcast reads)
            for (int i = hl; i < kGates * NJ; i += kRnnEpiThreads) {
                const int g = i / NJ, jj = i % NJ;
                s_bias[i] = (j0 + jj < H) ? p.b_hh[(size_t)dir * G + g * H + j0 + jj] : 0.f;
            }
            named_bar_sync(3, kRnnEpiThreads);
        }

        for (int s = 0; s < T; ++s) {
            const int t = t_of(s);
            const bool active = rowok && (t < len);
            if (hl == 0) ASRB_TRACE(4, s);
            if constexpr (KSPLIT) {
                if (hl == 0 && s > 0) {   // the peer sends one 16-byte vector per (row of the tile quarters in use, unit group)
                    const int quads = min(4, ceil_div(B, kRowsPerWarp));
                    mbar_arrive_expect_tx(&x_bar[s & 1], (uint32_t)((KS - 1) * MROWS * NJ * 4));
                    (void)quads;
                }
            }
            constexpr int kAccG = BWD ? 1 : kGates;      // accumulator column groups we read: gates (fwd) / units (bwd)
            float acc[kAccG][4];
#pragma unroll
            for (int g = 0; g < kAccG; ++g)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[g][jj] = 0.f;

            // ---- operand prefetch (independent of the MMA): issued before we wait for the accumulator ----
            constexpr int kIn = BWD ? 6 : kGates;        // fwd: gi gates ; bwd: 4 saved + dout + previous state
            float in[kIn][4];
#pragma unroll
            for (int q = 0; q < kIn; ++q)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) in[q][jj] = 0.f;
            float ct[4] = {0.f, 0.f, 0.f, 0.f};          // bwd LSTM: c_t
            if (active && !(p.dbg & 2)) {
                if constexpr (!BWD) {
                    const float* g = p.gi + (((size_t)t * B + b) * 2 + dir) * G + j0 + u0;
#pragma unroll
                    for (int q = 0; q < kGates; ++q) ldg4(in[q], g + (size_t)q * H);
                } else {
                    const float* sv = p.saved + ((((size_t)dir * T + t) * p.P_saved + pidx) * 4) * (size_t)(NV * B * 4) + (size_t)b * 4;
                    const float* dop = p.dout + ((size_t)t * B + b) * H + j0 + u0;
                    const int tprev_slot = (dir == 0) ? t : t + 2;  // slot of the step that preceded t in forward order
                    const float* prevp = (CELL == ASRB_RNN_GRU ? p.hseq : p.cseq) +
                                         ((size_t)dir * (T + 2) + tprev_slot) * slotHB + (size_t)b * H + j0 + u0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) ld4(in[q], sv + ((size_t)q * NV + ug) * (size_t)(B * 4));
                    ldg4(in[4], dop);
                    ld4(in[5], prevp);
                    if constexpr (CELL == ASRB_RNN_LSTM)
                        ld4(ct, p.cseq + ((size_t)dir * (T + 2) + t + 1) * slotHB + (size_t)b * H + j0 + u0);
                }
            }

            // ---- recurrent product for this step ----
            if (s > 0) {
                if (tc) {
                    mbar_wait(tfull_bar, (uint32_t)((s - 1) & 1));
                    if (hl == 0) ASRB_TRACE(5, s);
                    tc_fence_after_sync();
                    if (warp_ld) {
                        const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + u0;
                        if constexpr (KSPLIT) {
                            tmem_ld_32x4(taddr + crank * NJ, acc[0]);          // partial sums of our own units
#pragma unroll
                            for (int q = 1; q < KS; ++q)                       // ... and of every peer's units
                                tmem_ld_32x4(taddr + ((crank + q) % KS) * NJ, xsend[q - 1]);
                        } else {
#pragma unroll
                            for (int g = 0; g < kAccG; ++g) tmem_ld_32x4(taddr + (BWD ? 0 : g * NJ), acc[g]);
                        }
                        tmem_ld_wait();
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar);
                    if (hl == 0) ASRB_TRACE(6, s);
                    if constexpr (KSPLIT) {
                        // Exchange through distributed shared memory, double-buffered by step parity (a peer can only write
                        // parity q again after it has received our data of the step in between, which also means it has
                        // finished reading our staging buffer of parity q).  The partial sums are staged locally and go out
                        // as ONE bulk copy per peer.  Measured alternatives (cycles from the accumulator read to the summed
                        // result, pairs / clusters of four): per-thread st.async with complete_tx 1.2 k / 2.5 k; bulk copy
                        // 1.2 k / 2.1 k; plain DSMEM stores + mbarrier.arrive.release.cluster 2.5 k; plain stores + the
                        // hardware cluster barrier 3.1 k -- ordering at cluster scope is what costs, not the bytes.
                        const int par = s & 1;
                        if (warp_ld && lane < kRowsPerWarp) {
#pragma unroll
                            for (int q = 1; q < KS; ++q)
                                st4(xstage + (((size_t)par * (KS - 1) + (q - 1)) * MROWS + b) * NJ + u0, xsend[q - 1]);
                        }
                        fence_proxy_async_smem();                      // generic stores -> the bulk copy's async-proxy reads
                        named_bar_sync(5, kRnnEpiThreads);
                        if (hl == 0) {
                            ASRB_TRACE(12, s);
#pragma unroll
                            for (int q = 1; q < KS; ++q) {
                                const uint32_t peer = (crank + q) % KS;
                                // lands in slot [par][source = our rank] of the peer's buffer
                                dsmem_bulk_copy(map_to_cta(xbuf + ((size_t)par * KS + crank) * MROWS * NJ, peer),
                                                xstage + ((size_t)par * (KS - 1) + (q - 1)) * MROWS * NJ,
                                                (uint32_t)(MROWS * NJ * 4), map_to_cta(&x_bar[par], peer));
                            }
                            ASRB_TRACE(13, s);
                        }
                        mbar_wait_cluster(&x_bar[par], xphase[par]);
                        xphase[par] ^= 1u;
                        if (warp_ld && lane < kRowsPerWarp) {
#pragma unroll
                            for (int q = 1; q < KS; ++q) {
                                float xr[4];
                                ld4(xr, xbuf + (((size_t)par * KS + (crank + q) % KS) * MROWS + b) * NJ + u0);
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) acc[0][jj] += xr[jj];
                            }
                        }
                        if (hl == 0) ASRB_TRACE(11, s);
                    }
                } else {
                    // DEBUG path (asrb_set_debug_flags bit 1): same algorithm, plain fp32 dot products
                    if (hl == 0) {
                        const uint32_t need = (uint32_t)P * (uint32_t)s;
                        poll_counter(counter, need);
                    }
                    named_bar_sync(1, kRnnEpiThreads);
                    if (rowok) {
                        const int tp = t_of(s - 1);
                        const int K = BWD ? G : H;
                        const float* arow = BWD ? p.dgh + (((size_t)dir * T + tp) * B + b) * G
                                                : p.hseq + ((size_t)dir * (T + 2) + tp + 1) * slotHB + (size_t)b * H;
                        const float* wrow = p.wpack + ((size_t)(dir * P + pidx) * NPAD) * p.kpad;
                        for (int k = 0; k < K; ++k) {
                            const float a = __ldcg(arow + k);
#pragma unroll
                            for (int g = 0; g < kAccG; ++g)
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int col = (BWD ? 0 : g * NJ) + u0 + jj;
                                    acc[g][jj] = fmaf(a, __ldg(wrow + (size_t)col * p.kpad + k), acc[g][jj]);
                                }
                        }
                    }
                }
            }

            // ---- cell math (registers only), then 16-byte vector stores ----
            if constexpr (!BWD) {
                float hn[4], cn[4], sv[4][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float h_ = 0.f, c_ = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    if (active) {
                        const float* bs = s_bias + u0 + jj;
                        if constexpr (CELL == ASRB_RNN_GRU) {
                            const float gr = acc[0][jj] + bs[0];
                            const float gz = acc[1][jj] + bs[NJ];
                            const float gn = acc[2][jj] + bs[2 * NJ];
                            const float r = fsigmoid(in[0][jj] + gr);
                            const float z = fsigmoid(in[1][jj] + gz);
                            const float n = ftanh(in[2][jj] + r * gn);
                            h_ = (1.f - z) * n + z * state_h[jj];
                            s0 = r; s1 = z; s2 = n; s3 = gn;
                        } else {
                            const float gi_ = fsigmoid(in[0][jj] + acc[0][jj] + bs[0]);
                            const float gf = fsigmoid(in[1][jj] + acc[1][jj] + bs[NJ]);
                            const float gg = ftanh(in[2][jj] + acc[2][jj] + bs[2 * NJ]);
                            const float go = fsigmoid(in[3][jj] + acc[kGates - 1][jj] + bs[(kGates - 1) * NJ]);
                            c_ = gf * state_c[jj] + gi_ * gg;
                            h_ = go * ftanh(c_);
                            s0 = gi_; s1 = gf; s2 = gg; s3 = go;
                        }
                    }
                    hn[jj] = h_; cn[jj] = c_;
                    sv[0][jj] = s0; sv[1][jj] = s1; sv[2][jj] = s2; sv[3][jj] = s3;
                    state_h[jj] = h_;
                    state_c[jj] = c_;
                }
                // (1) the next step's MMA operand goes out first and is published; (2) the stores nobody waits for
                //     (fp32 state, saved gates) follow AFTER the release and overlap the wait for the next step
                const size_t o = ((size_t)dir * (T + 2) + t + 1) * slotHB + (size_t)b * H + j0 + u0;
                if (rowok) {
                    if constexpr (BF16) st4_bf16(p.hbf + (((size_t)dir * (T + 2) + t + 1) * B + b) * p.Hp + j0 + u0, hn);
                    else                st4(p.hseq + o, hn);
                }
                if (hl == 0) ASRB_TRACE(7, s);
                // Step barrier, cooperative-groups style: CTA barrier, then ONE thread fences (the MEMBAR.ALL.GPU inside
                // red.release covers the stores it observed through the barrier) and bumps the counter.  The consumers'
                // TMA reads are ordered by their own acquire + fence.proxy.async.global; no per-thread fence here (an
                // unqualified fence.proxy.async is a MEMBAR.ALL.GPU in every warp: measured 1400 cycles on the chain).
                named_bar_sync(2, kRnnEpiThreads);
                if (hl == 0) {
                    ASRB_TRACE(8, s);
                    red_release_add_u32(counter, 1u);
                    ASRB_TRACE(10, s);
                }
                // hold the bulk stores back until the release has been issued: a MEMBAR waits for every store in flight
                named_bar_sync(4, kRnnEpiThreads);
                if (rowok && !(p.dbg & 1)) {
                    // saved gates, slice-major [dir][t][slice][gate][group][b][4]: consecutive batch rows (= lanes) are
                    // 16 bytes apart, so a warp store covers 2 lines instead of 16
                    float* svp = p.saved + ((((size_t)dir * T + t) * P + pidx) * 4) * (size_t)(NV * B * 4) + (size_t)b * 4;
                    if constexpr (BF16) st4(p.hseq + o, hn);
                    if constexpr (CELL == ASRB_RNN_LSTM) st4(p.cseq + o, cn);
#pragma unroll
                    for (int q = 0; q < 4; ++q) st4(svp + ((size_t)q * NV + ug) * (size_t)(B * 4), sv[q]);
                }
            } else {
                float dg[4][4], eg2[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float carry = acc[0][jj] + state_h[jj];
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, e2 = 0.f;
                    if (active) {
                        const float dh = carry + in[4][jj];
                        if constexpr (CELL == ASRB_RNN_GRU) {
                            const float r = in[0][jj], z = in[1][jj], n = in[2][jj], gn = in[3][jj], hp = in[5][jj];
                            const float dn = dh * (1.f - z) * (1.f - n * n);
                            d2 = dn;                          // d gi_n
                            e2 = dn * r;                      // d gh_n
                            d1 = dh * (hp - n) * z * (1.f - z);
                            d0 = dn * gn * r * (1.f - r);
                            state_h[jj] = dh * z;
                        } else {
                            const float gi_ = in[0][jj], gf = in[1][jj], gg = in[2][jj], go = in[3][jj], cp = in[5][jj];
                            const float tcv = ftanh(ct[jj]);
                            const float dc = state_c[jj] + dh * go * (1.f - tcv * tcv);
                            d0 = dc * gg * gi_ * (1.f - gi_);
                            d1 = dc * cp * gf * (1.f - gf);
                            d2 = dc * gi_ * (1.f - gg * gg);
                            d3 = dh * tcv * go * (1.f - go);
                            e2 = d2;
                            state_c[jj] = dc * gf;
                            state_h[jj] = 0.f;
                        }
                    } else {
                        state_h[jj] = carry;  // gradient passes an inactive step untouched
                    }
                    dg[0][jj] = d0; dg[1][jj] = d1; dg[2][jj] = d2; dg[3][jj] = d3; eg2[jj] = e2;
                }
                const size_t oh = (((size_t)dir * T + t) * B + b) * G + j0 + u0;
                if (rowok) {   // (1) next step's MMA operand first
#pragma unroll
                    for (int q = 0; q < kGates; ++q) {
                        const float* hv = (q == 2) ? eg2 : dg[q];
                        if constexpr (BF16) st4_bf16(p.dghbf + (((size_t)dir * T + t) * B + b) * p.Gp + j0 + u0 + (size_t)q * H, hv);
                        if (p.dgh) st4(p.dgh + oh + (size_t)q * H, hv);
                    }
                }
                if (hl == 0) ASRB_TRACE(7, s);
                named_bar_sync(2, kRnnEpiThreads);      // see the forward branch
                if (hl == 0) {
                    ASRB_TRACE(8, s);
                    red_release_add_u32(counter, 1u);
                    ASRB_TRACE(10, s);
                }
                named_bar_sync(4, kRnnEpiThreads);
                if (rowok && !(p.dbg & 1)) {   // (2) outputs only later kernels read
                    using GT = typename std::conditional<BF16, __nv_bfloat16, float>::type;
                    GT* dgi = reinterpret_cast<GT*>(p.dgi) + (((size_t)t * B + b) * 2 + dir) * G + j0 + u0;
                    // transposed copies for the weight-gradient GEMMs: lanes (batch rows) are contiguous -> coalesced
                    GT* gT = reinterpret_cast<GT*>(p.dgiT) + ((size_t)dir * G + j0 + u0) * p.ldT + (size_t)t * B + b;
                    GT* hT = p.dghT ? reinterpret_cast<GT*>(p.dghT) + ((size_t)dir * G + j0 + u0) * p.ldT + (size_t)t * B + b : nullptr;
#pragma unroll
                    for (int q = 0; q < kGates; ++q) {
                        const float* hv = (q == 2) ? eg2 : dg[q];
                        if constexpr (BF16) st4_bf16(dgi + (size_t)q * H, dg[q]);
                        else                st4(dgi + (size_t)q * H, dg[q]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) pack_store(gT + ((size_t)q * H + e) * p.ldT, dg[q][e]);
                        if (hT) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) pack_store(hT + ((size_t)q * H + e) * p.ldT, hv[e]);
                        }
                    }
                }
            }
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

// out[t,b,:] = hseq[0][t+1][b][:] + hseq[1][t+1][b][:]     (blocks.py:92 "sum(2)")
__global__ void rnn_sum_dirs_kernel(const float* __restrict__ hseq, float* __restrict__ out, int T, long long BH) {
    const long long n = (long long)T * BH;
    const float* h0 = hseq + BH;
    const float* h1 = hseq + (long long)(T + 2) * BH + BH;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = h0[i] + h1[i];
}

// bf16 != 0: bf16 operands for the recurrent product (needs H % 8 == 0); else tf32 (H % 4 == 0)
static int rnn_make_plan(int cell, int H, int B, int bf16, RnnPlan* pl) {
    const int gates = cell == ASRB_RNN_GRU ? 3 : 4;
    const int G = gates * H;
    const int esize = bf16 ? 2 : 4, kbe = 128 / esize;
    if (B > kRnnMaxRows || H % (16 / esize) != 0) return ASRB_ERR_UNSUPPORTED;
    const int mrows = B <= 64 ? 64 : 128;
    const int stage = mrows * 128;
    // widest slice first: fewer CTAs mean less replicated state traffic and a cheaper step barrier
    const int cands[3] = {16, 12, 8};
    for (int ci = 0; ci < 3; ++ci) {
        const int nj = cands[ci];
        const int P = ceil_div(H, nj);
        if (2 * P > kNumSMs) continue;
        RnnPlan r;
        r.nj = nj; r.P = P; r.mrows = mrows; r.bf16 = bf16;
        // backward: K split over clusters of 4 (the slice count padded to a multiple of 4 with empty CTAs) or 2 CTAs
        // when the geometry allows and at least 4 ring blocks still fit beside the weights and the exchange buffers
        r.npad_f = round_up(gates * nj, 16);
        r.kpad_f = round_up(H, kbe);
        const size_t wf = (size_t)r.npad_f * r.kpad_f * esize;
        const size_t fixed = 1024 + kRnnBarBytes;
        size_t wb = 0, xb = 0;
        // rnn3.cu backward (weights in tensor memory): clusters of 4, the K quarter has to fit 448 TMEM columns
        r.ts_bwd = 0;
        {
            const int kpad4 = ceil_div(ceil_div(G, 64), 4) * 64;
            if (bf16 && nj == 16 && B <= 64 && !(g_rnn_dbg & (8 | 256 | 1024)) && kpad4 <= 896 && 2 * round_up(P, 4) <= kNumSMs)
                r.ts_bwd = 1;
        }
        const int ks_try[3] = {4, 2, 0};
        for (int ki = 0; ki < 3; ++ki) {
            const int ks = ks_try[ki];
            if (r.ts_bwd && ks != 4) continue;
            if (ks && !(bf16 && nj == 16 && (g_rnn_ksplit >= ks || r.ts_bwd))) continue;
            if (ks && 2 * round_up(P, ks) > kNumSMs) continue;
            r.ksplit = ks;
            r.P_b = ks ? round_up(P, ks) : P;
            r.npad_b = ks ? ks * nj : 16;
            r.kpad_b = ks ? ceil_div(ceil_div(G, kbe), ks) * kbe : round_up(G, kbe);
            wb = (size_t)r.npad_b * r.kpad_b * esize;
            xb = ks ? (size_t)2 * (2 * ks - 1) * mrows * nj * 4 : 0;   // receive + staging buffers of the K split
            if (!ks || r.ts_bwd || wb + fixed + xb + 4 * (size_t)stage <= (size_t)kRnnMaxSmem) break;
        }
        if (r.ts_bwd) wb = xb = 0;     // its shared-memory budget is its own (rnn3_bwd_launch)
        if (wf + fixed + stage > (size_t)kRnnMaxSmem || wb + fixed + xb + stage > (size_t)kRnnMaxSmem) continue;
        // as many K blocks in flight as fit (the whole previous state when possible): the step is latency-bound
        // K blocks that fit next to the resident weights; up to 4 blocks share one barrier / pipeline stage
        const int bf = (int)((kRnnMaxSmem - fixed - wf) / stage), bb = (int)((kRnnMaxSmem - fixed - xb - wb) / stage);
        const int nkb_f = r.kpad_f / kbe, nkb_b = r.kpad_b / kbe;
        r.chunk_f = g_rnn_chunk > 0 ? g_rnn_chunk : (bf >= 8 ? 4 : (bf >= 4 ? 2 : 1));
        r.chunk_b = g_rnn_chunk > 0 ? g_rnn_chunk : (bb >= 8 ? 4 : (bb >= 4 ? 2 : 1));
        r.stages_f = bf / r.chunk_f; r.stages_b = bb / r.chunk_b;
        const int cf = ceil_div(nkb_f, r.chunk_f), cb = ceil_div(nkb_b, r.chunk_b);
        if (r.stages_f > cf) r.stages_f = cf;
        if (r.stages_b > cb) r.stages_b = cb;
        if (r.stages_f > kRnnMaxStages) r.stages_f = kRnnMaxStages;
        if (r.stages_b > kRnnMaxStages) r.stages_b = kRnnMaxStages;
        r.smem_f = wf + fixed + (size_t)r.stages_f * r.chunk_f * stage;
        r.smem_b = wb + fixed + xb + (size_t)r.stages_b * r.chunk_b * stage;
        *pl = r;
        return 0;
    }
    return ASRB_ERR_UNSUPPORTED;
}

template <int CELL, int NJ, bool BWD, bool BF16, int MROWS, int KS = 1>
static int rnn_launch(const RnnPlan& pl, RnnParams& prm, const void* wpack, const void* a_base, asrb_stream_t stream) {
    using S = RnnShape<CELL, NJ>;
    constexpr bool KSPLIT = KS > 1;
    constexpr int NPAD = BWD ? (KSPLIT ? KS * NJ : S::kNpadB) : S::kNpadF;
    const int Pk = (BWD && KSPLIT) ? pl.P_b : pl.P;      // CTAs per direction of THIS launch
    prm.P_saved = pl.P;
    prm.P = Pk;
    constexpr int KBE = BF16 ? 64 : 32;
    constexpr int ES = BF16 ? 2 : 4;
    const int kpad = BWD ? pl.kpad_b : pl.kpad_f;
    const size_t smem = BWD ? pl.smem_b : pl.smem_f;
    prm.kpad = kpad;
    prm.stages = BWD ? pl.stages_b : pl.stages_f;
    prm.chunk = BWD ? pl.chunk_b : pl.chunk_f;
    prm.wpack = BF16 ? nullptr : reinterpret_cast<const float*>(wpack);
    CUtensorMap tmW, tmA;
    {
        uint64_t d[2] = {(uint64_t)kpad, (uint64_t)2 * Pk * NPAD}, s[1] = {(uint64_t)kpad * ES};
        uint32_t bx[2] = {KBE, (uint32_t)NPAD};
        int rc = BF16 ? make_tmap_bf16(&tmW, wpack, 2, d, s, bx) : make_tmap_f32(&tmW, wpack, 2, d, s, bx);
        if (rc) return rc;
    }
    {
        const int K = BWD ? prm.G : prm.H;
        const int slabs = BWD ? 2 * prm.T : 2 * (prm.T + 2);
        uint64_t d[3] = {(uint64_t)K, (uint64_t)prm.B, (uint64_t)slabs};
        const uint64_t pitch = BF16 ? (uint64_t)round_up(K, 64) : (uint64_t)K;   // columns >= K: TMA zero fill
        uint64_t s[2] = {pitch * ES, (uint64_t)prm.B * pitch * ES};
        uint32_t bx[3] = {KBE, (uint32_t)MROWS, 1};
        int rc = BF16 ? make_tmap_bf16(&tmA, a_base, 3, d, s, bx) : make_tmap_f32(&tmA, a_base, 3, d, s, bx);
        if (rc) return rc;
    }
    auto kern = rnn_rec_kernel<CELL, NJ, BWD, BF16, MROWS, KS>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ASRB_CUDA_OK(cudaMemsetAsync(prm.counters, 0, 2 * kRnnCounterStride * sizeof(uint32_t), stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * Pk);
    cfg.blockDim = dim3(kRnnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    // Plain (not cooperative) launch: the step barrier needs all 2P <= 148 CTAs co-resident, which one CTA per SM on an
    // otherwise idle device gives (kernels of the same stream have drained; nothing else runs beside the recurrence).
    // A cooperative launch makes the same promise formally but BLOCKS the host until the kernel can start, so the CPU
    // could not queue the next layer's kernels behind the recurrence (measured: ~0.3 ms of idle GPU after every launch).
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = KS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = KSPLIT ? 1 : 0;
    prm.dbg = g_rnn_dbg;
    {   // the step barrier spins: refuse the launch when the device cannot hold the whole grid at once
        int dev = 0, sms = 0, per_sm = 0;
        ASRB_CUDA_OK(cudaGetDevice(&dev));
        ASRB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        ASRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRnnThreads, smem));
        if (2 * Pk > sms * per_sm) return ASRB_ERR_UNSUPPORTED;
    }
    ASRB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmW, tmA, prm));
    return 0;
}

template <bool BWD>
static int rnn_dispatch(int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, const void* a_base,
                        asrb_stream_t stream) {
#define ASRB_RNN_CASE(C, N)                                                                              \
    if (cell == C && pl.nj == N) {                                                                       \
        if constexpr (BWD && N == 16) {                                                                  \
            if (pl.ksplit == 2) {                                                                        \
                if (pl.mrows == 64) return rnn_launch<C, N, true, true, 64, 2>(pl, prm, wpack, a_base, stream);  \
                return rnn_launch<C, N, true, true, 128, 2>(pl, prm, wpack, a_base, stream);             \
            }                                                                                            \
            if (pl.ksplit == 4) {                                                                        \
                if (pl.mrows == 64) return rnn_launch<C, N, true, true, 64, 4>(pl, prm, wpack, a_base, stream);  \
                return rnn_launch<C, N, true, true, 128, 4>(pl, prm, wpack, a_base, stream);             \
            }                                                                                            \
        }                                                                                                \
        if (pl.bf16) {                                                                                   \
            if (pl.mrows == 64) return rnn_launch<C, N, BWD, true, 64>(pl, prm, wpack, a_base, stream);  \
            return rnn_launch<C, N, BWD, true, 128>(pl, prm, wpack, a_base, stream);                     \
        }                                                                                                \
        if (pl.mrows == 64) return rnn_launch<C, N, BWD, false, 64>(pl, prm, wpack, a_base, stream);     \
        return rnn_launch<C, N, BWD, false, 128>(pl, prm, wpack, a_base, stream);                        \
    }
    ASRB_RNN_CASE(ASRB_RNN_GRU, 8) ASRB_RNN_CASE(ASRB_RNN_GRU, 12) ASRB_RNN_CASE(ASRB_RNN_GRU, 16)
    ASRB_RNN_CASE(ASRB_RNN_LSTM, 8) ASRB_RNN_CASE(ASRB_RNN_LSTM, 12) ASRB_RNN_CASE(ASRB_RNN_LSTM, 16)
#undef ASRB_RNN_CASE
    return ASRB_ERR_UNSUPPORTED;
}

// Which kernel runs the bf16 recurrent product (asrb_debug_rnn_dbg bits, for A/B timing):
//   forward, batch <= 64, H <= 896: rnn3.cu (weights in tensor memory, two interleaved half-batch chains) -- default;
//   bit 8: the counter + TMA kernel of this file (the default for everything rnn3.cu does not cover);
//   backward, batch <= 64, G/4 <= 896: rnn3.cu (weights in tensor memory, clusters of 4, two chains) -- default;
//     bit 1024: the backward of this file instead;
//   bit 256: the experimental exchange-by-data kernel (rnn2.cu), forward and backward.
// Inside rnn3.cu: the step hand-over is VERIFIED by default (a fast first pass that detects a stale operand + a release
// second pass that runs only then; bit 8192 forces the second pass: see rnn3.cu); bit 4096: generic stores + red.release in
// one pass.  bit 16: ONE TMA store of the operand tile +
// cp.async.bulk.wait_group + a relaxed increment -- 0.3 / 0.5 ms per layer faster (the release's MEMBAR.GPU waits for the
// other chain's TMA copies), but the completion does NOT mean the data is in L2: about one hand-over in 10^7 was consumed too
// early (tools/stress_fullsize.py), so it is an experiment only; bit 4: the same + an L2 read-back of the tile before the
// increment.  bit 1: forward without the stores nobody waits for (timing only, results incomplete); bit 2048: backward
// outputs through direct stores instead of shared-memory tiles + TMA stores; bit 32768: forward with / backward without
// the tensor-pipe lock between the two chains' MMA sequences.
static inline bool rnn2_eligible(const RnnPlan& pl, const RnnParams& prm) {
    return pl.bf16 && !prm.use_simt && (g_rnn_dbg & 256) && pl.nj == 16 && prm.H % 16 == 0;
}
static inline bool rnn3_eligible(const RnnPlan& pl, const RnnParams& prm) {
    return pl.bf16 && !prm.use_simt && !(g_rnn_dbg & (8 | 256)) && pl.nj == 16 && prm.B <= 64 && pl.kpad_f <= 896;
}

// the CUDA-core debug product reads fp32 operands: it always runs the tf32-layout variant
static inline int rnn_effective_bf16(int bf16) { return (g_debug_flags & ASRB_DEBUG_SIMT_RNN) ? 0 : (bf16 ? 1 : 0); }

}  // namespace asrb

using namespace asrb;

extern "C" {

/* floats in the saved-gates buffer (slice-major layout private to asrb_rnn_fwd / asrb_rnn_bwd) */
size_t asrb_rnn_saved_floats(int cell, int H, int B, int bf16, int T) {
    RnnPlan pl;
    if (rnn_make_plan(cell, H, B, rnn_effective_bf16(bf16), &pl)) return 0;
    return (size_t)2 * T * pl.P * 4 * pl.nj * B;
}

int asrb_rnn_plan(int cell, int H, int B, int bf16, int* nj, int* P, size_t* wpack_fwd_bytes, size_t* wpack_bwd_bytes) {
    ASRB_REQUIRE((cell == ASRB_RNN_GRU || cell == ASRB_RNN_LSTM) && H > 0 && B > 0, ASRB_ERR_BAD_ARG);
    RnnPlan pl;
    int rc = rnn_make_plan(cell, H, B, rnn_effective_bf16(bf16), &pl);
    if (rc) return rc;
    const size_t es = pl.bf16 ? 2 : 4;
    if (nj) *nj = pl.nj;
    if (P) *P = pl.P;
    if (wpack_fwd_bytes) *wpack_fwd_bytes = (size_t)2 * pl.P * pl.npad_f * pl.kpad_f * es;
    if (wpack_bwd_bytes) *wpack_bwd_bytes = (size_t)2 * pl.P_b * pl.npad_b * pl.kpad_b * es;
    return 0;
}

int asrb_rnn_pack_weights(int cell, int H, int B, int bf16, const float* w_hh_fwd, const float* w_hh_rev,
                          void* wpack_fwd, void* wpack_bwd, asrb_stream_t stream) {
    ASRB_REQUIRE(w_hh_fwd && w_hh_rev, ASRB_ERR_BAD_ARG);
    RnnPlan pl;
    int rc = rnn_make_plan(cell, H, B, rnn_effective_bf16(bf16), &pl);
    if (rc) return rc;
    const int gates = cell == ASRB_RNN_GRU ? 3 : 4;
    if (wpack_fwd) {
        const int rows = 2 * pl.P * pl.npad_f;
        if (pl.bf16) rnn_pack_fwd_kernel<<<rows, 256, 0, stream>>>(w_hh_fwd, w_hh_rev, (__nv_bfloat16*)wpack_fwd, H, gates, pl.nj, pl.P, pl.npad_f, pl.kpad_f);
        else         rnn_pack_fwd_kernel<<<rows, 256, 0, stream>>>(w_hh_fwd, w_hh_rev, (float*)wpack_fwd, H, gates, pl.nj, pl.P, pl.npad_f, pl.kpad_f);
        ASRB_LAUNCH_OK();
    }
    if (wpack_bwd) {
        const int Pb = pl.ksplit ? pl.P_b : pl.P;
        const dim3 grid(ceil_div(pl.kpad_b, 32), ceil_div(pl.npad_b, 32), 2 * Pb), block(32, 8);
        if (pl.ksplit || pl.bf16)
            rnn_pack_bwd_tiled_kernel<<<grid, block, 0, stream>>>(w_hh_fwd, w_hh_rev, (__nv_bfloat16*)wpack_bwd, H, gates * H, pl.nj, Pb, pl.npad_b, pl.kpad_b, pl.ksplit);
        else
            rnn_pack_bwd_tiled_kernel<<<grid, block, 0, stream>>>(w_hh_fwd, w_hh_rev, (float*)wpack_bwd, H, gates * H, pl.nj, Pb, pl.npad_b, pl.kpad_b, 0);
        ASRB_LAUNCH_OK();
    }
    return 0;
}

int asrb_rnn_fwd_sum(int cell, int bf16, const float* gi, const float* b_hh, const void* wpack_fwd, const int32_t* lengths,
                     float* hseq, void* hseq_bf16, float* cseq, float* saved, float* out_sum, uint32_t* counters, int T, int B,
                     int H, asrb_stream_t stream) {
    ASRB_REQUIRE(gi && b_hh && wpack_fwd && lengths && hseq && saved && counters && T > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(cell == ASRB_RNN_GRU || cseq, ASRB_ERR_BAD_ARG);
    RnnPlan pl;
    int rc = rnn_make_plan(cell, H, B, rnn_effective_bf16(bf16), &pl);
    if (rc) return rc;
    ASRB_REQUIRE(!pl.bf16 || hseq_bf16, ASRB_ERR_BAD_ARG);
    RnnParams prm = {};
    prm.T = T; prm.B = B; prm.H = H; prm.G = (cell == ASRB_RNN_GRU ? 3 : 4) * H; prm.P = pl.P;
    prm.use_simt = (g_debug_flags & ASRB_DEBUG_SIMT_RNN) ? 1 : 0;
    prm.lengths = lengths; prm.counters = counters;
    prm.gi = gi; prm.b_hh = b_hh; prm.hseq = hseq; prm.cseq = cseq; prm.saved = saved;
    prm.hbf = reinterpret_cast<__nv_bfloat16*>(hseq_bf16);
    prm.Hp = round_up(H, 64);
    prm.trace = g_rnn_trace;
    if (rnn3_eligible(pl, prm)) {
        // (A variant in which the kernel itself added both directions' h tiles into out_sum with TMA reduce-adds and stored the
        // fp32 state / saved activations through shared-memory tiles was built and measured: 12.45 ms per step for the five
        // forward launches against 11.46 + 0.27 with the separate sum kernel -- the reduce-adds cost more L2 time than the
        // streaming sum; removed.)
        prm.out_sum = nullptr;
        rc = rnn3_forward(cell, pl, prm, wpack_fwd, stream);
        if (rc == 0 && out_sum && !prm.out_sum) rc = asrb_rnn_sum_dirs(hseq, out_sum, T, B, H, stream);
        return rc;
    }
    if (rnn2_eligible(pl, prm)) rc = rnn2_dispatch(false, cell, pl, prm, wpack_fwd, stream);
    else rc = rnn_dispatch<false>(cell, pl, prm, wpack_fwd, pl.bf16 ? (const void*)hseq_bf16 : (const void*)hseq, stream);
    if (rc == 0 && out_sum) rc = asrb_rnn_sum_dirs(hseq, out_sum, T, B, H, stream);
    return rc;
}

int asrb_rnn_fwd(int cell, int bf16, const float* gi, const float* b_hh, const void* wpack_fwd, const int32_t* lengths,
                 float* hseq, void* hseq_bf16, float* cseq, float* saved, uint32_t* counters, int T, int B, int H,
                 asrb_stream_t stream) {
    return asrb_rnn_fwd_sum(cell, bf16, gi, b_hh, wpack_fwd, lengths, hseq, hseq_bf16, cseq, saved, nullptr, counters, T, B, H, stream);
}

int asrb_rnn_bwd(int cell, int bf16, const float* dout, const void* wpack_bwd, const int32_t* lengths,
                 const float* hseq, const float* cseq, const float* saved, void* dgi, float* dgh, void* dgh_bf16,
                 void* dgiT, void* dghT, long long ldT, uint32_t* counters, int T, int B, int H,
                 asrb_stream_t stream) {
    ASRB_REQUIRE(dout && wpack_bwd && lengths && hseq && saved && dgi && dgiT && counters && T > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(cell == ASRB_RNN_GRU || cseq, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(cell == ASRB_RNN_LSTM || dghT, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(ldT >= (long long)T * B, ASRB_ERR_BAD_ARG);
    RnnPlan pl;
    int rc = rnn_make_plan(cell, H, B, rnn_effective_bf16(bf16), &pl);
    if (rc) return rc;
    ASRB_REQUIRE(pl.bf16 ? (dgh_bf16 != nullptr) : (dgh != nullptr), ASRB_ERR_BAD_ARG);
    RnnParams prm = {};
    prm.T = T; prm.B = B; prm.H = H; prm.G = (cell == ASRB_RNN_GRU ? 3 : 4) * H; prm.P = pl.P;
    prm.use_simt = (g_debug_flags & ASRB_DEBUG_SIMT_RNN) ? 1 : 0;
    prm.lengths = lengths; prm.counters = counters;
    prm.hseq = const_cast<float*>(hseq); prm.cseq = const_cast<float*>(cseq); prm.saved = const_cast<float*>(saved);
    prm.dout = dout; prm.dgi = dgi; prm.dgh = dgh; prm.dgiT = dgiT; prm.dghT = (cell == ASRB_RNN_GRU) ? dghT : nullptr; prm.ldT = ldT;
    prm.dghbf = reinterpret_cast<__nv_bfloat16*>(dgh_bf16);
    prm.Gp = round_up(prm.G, 64);
    prm.trace = g_rnn_trace;
    if (pl.ts_bwd && !prm.use_simt) return rnn3_backward(cell, pl, prm, wpack_bwd, stream);
    if (rnn2_eligible(pl, prm)) return rnn2_dispatch(true, cell, pl, prm, wpack_bwd, stream);
    return rnn_dispatch<true>(cell, pl, prm, wpack_bwd, pl.bf16 ? (const void*)dgh_bf16 : (const void*)dgh, stream);
}

/* DEBUG / timing experiments: backward K split over CTA pairs on (1, default) / off (0).  Changes the packed-weight
 * layout: call before asrb_rnn_plan / asrb_rnn_pack_weights. */
int asrb_debug_rnn_ksplit(int on) { g_rnn_ksplit = on == 1 ? 2 : on; return 0; }   /* 0 off, 1 or 2 pairs, 4 clusters of four */

/* DEBUG / timing experiments: K blocks per pipeline barrier (0 = automatic) */
int asrb_debug_rnn_dbg(int bits) { g_rnn_dbg = bits; return 0; }
int asrb_debug_rnn_redos(void) {
    const long long v = asrb::rnn3_redo_count();
    return v < 0 ? -1 : (int)(v & 0x7fffffff);
}
int asrb_debug_rnn_chunk(int blocks) {
    ASRB_REQUIRE(blocks >= 0 && blocks <= 8, ASRB_ERR_BAD_ARG);
    g_rnn_chunk = blocks;
    return 0;
}

/* DEBUG: per-step SM-clock stamps of the next asrb_rnn_fwd / asrb_rnn_bwd launches into trace[grid][T][12] (NULL = off) */
int asrb_debug_rnn_trace(long long* trace) {
    g_rnn_trace = trace;
    return 0;
}

int asrb_rnn_sum_dirs(const float* hseq, float* out, int T, int B, int H, asrb_stream_t stream) {
    ASRB_REQUIRE(hseq && out && T > 0 && B > 0 && H > 0, ASRB_ERR_BAD_ARG);
    const long long n = (long long)T * B * H;
    const int grid = (int)((n + 255) / 256 < kNumSMs * 8 ? (n + 255) / 256 : kNumSMs * 8);
    rnn_sum_dirs_kernel<<<grid, 256, 0, stream>>>(hseq, out, T, (long long)B * H);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
