// asr_b200 -- forward GRU / LSTM recurrence: weights in tensor memory, two interleaved half-batch chains.
//
// Same decomposition and step barrier as rnn.cu (CTA (dir, p) owns 16 hidden units, a release/acquire counter in global
// memory is the step barrier between the CTAs of a direction; modules/blocks.py:87-89 of the reference), restructured
// around what the SM-clock traces and microbenchmarks of round 2 established:
//
//   * a step of rnn.cu is ~10 k cycles of which ~4 k are exposed hand-over latency: the all-to-all visibility of the
//     freshly written state costs ~2.5 k cycles whatever the protocol (tools/ubench/xchg2.cu: 2-byte or 16-byte stores,
//     counters or data-as-flag polling), because every word crosses to its L2 home and back, half of the time over the
//     die-to-die link (tools/ubench/pingpong.cu: 450 cycles one way on the home die, 850-1000 otherwise).  It cannot be
//     removed -- but the batch rows are independent recurrences, so it can be HIDDEN: the batch is cut into two chains
//     of 32 rows with their own step counters, TMA / MMA / epilogue warps, shared-memory tiles and accumulators, and the
//     warp scheduler runs one chain's copy, product and cell math in the shadow of the other chain's hand-over.
//   * the recurrent product of rnn.cu is bound by shared-memory bandwidth (per step the state is written by TMA and read
//     back by the tensor core together with the resident weight slice).  Here the weight slice is the A operand in
//     TENSOR memory (M = 64 gate rows x K, written once with tcgen05.st; 416 of the 512 columns at H = 800), the chain's
//     32 batch rows are the N dimension and come from shared memory as the B operand: the tensor core reads only the
//     state, and N = 32 makes an MMA cost 16 cycles.
//   * the accumulator therefore comes out transposed (TMEM lane = gate row, column = batch row); it is read with
//     tcgen05.ld.16x256b (all 32 lanes hold data in the M = 64 layout) and handed through a small shared-memory tile to
//     the (batch row, unit) threads of the cell math.
//   * the load/store unit is kept free for the hand-over: the gate pre-activations arrive by TMA (64-byte-swizzle boxes,
//     a step ahead) instead of 48 scattered wavefronts per warp.
//
// Used for batch <= 64 and H <= 896 (64 + kpad/2 TMEM columns); everything else runs rnn.cu.
#include "rnn.cuh"

namespace asrb {

__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// global += shared (fp32), 3-D tile
__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int kR3Chunk = 4;                 // K blocks per TMA barrier
constexpr int kR3MaxChunks = 4;             // kpad <= 1024
constexpr int kR3CounterStride = 16;        // step counters [dir][chain], 64 bytes apart (8 of them in the 512-byte block)

// Two chains of 32 batch rows.  Warp c < 2 is chain c's control warp (one elected thread polls the chain's step counter,
// issues its TMA copies, then its MMAs -- the steps of a chain are sequential anyway); warps 4 .. 19 are the epilogue warps,
// 8 per chain.
//
// Accumulator rows: the CTA's 16 units x 4 gate slots are laid out so that TMEM lane quarter q holds ALL gates of units
// 4q .. 4q+3 (lane 32q + 4*gate + u).  An epilogue warp (quarter q = warp % 4, batch columns 16*half .. +15 of the chain)
// reads its quarter with one tcgen05.ld.16x256b.x2 -- thread t gets rows t/4 and t/4+8, i.e. gates (0, 2) of unit t/4 for
// t/4 < 4 and gates (1, 3) of unit t/4-4 otherwise, for four batch columns -- and one round of lane^16 swaps leaves every
// thread with all gates of one unit for two batch rows: no shared-memory transpose, no barrier between the accumulator read
// and the gate math.
// zero the K-padding columns [col0, col0 + ncols) of every row of a bf16 matrix (ncols even, 4-byte aligned): one 4-byte
// store per thread.  (cudaMemset2DAsync on 64-byte rows costs ~0.1 ms per call; this is a few microseconds.)
__global__ void rnn3_zero_pad_cols_kernel(__nv_bfloat16* base, size_t rows, int pitch, int col0, int ncols) {
    const int per_row = ncols / 2;
    const size_t n = rows * (size_t)per_row;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / per_row;
        const int c = (int)(i % per_row);
        *reinterpret_cast<uint32_t*>(base + r * pitch + col0 + 2 * c) = 0u;
    }
}
// Verified hand-over (see rnn3_launch): kernel-internal bits of RnnParams::dbg, the flag word and the second pass's counters
constexpr int kR3Check = 1 << 20;           // first pass: look for the sentinel in every operand tile that arrives
constexpr int kR3Redo = 1 << 21;            // second pass: leave at once unless the first pass raised the flag
constexpr int kR3Sabotage = 1 << 22;        // TEST: CTA 0 withholds the operand tile of step 5 (its counter arrival stays)
constexpr int kR3FlagWord = 8;              // counters[8]: "a stale operand sector was consumed"
constexpr int kR3RedoCounters = 64;         // the second pass counts in counters[64..128)
__device__ unsigned long long g_rnn3_redos = 0;   // launches repeated so far (asrb_debug_rnn_redos)

// rows of a bf16 matrix: columns [0, valid) <- 0xFFFF (a NaN no cell ever produces: the sentinel), [valid, pitch) <- 0 (the
// K padding).  pitch is a multiple of 8 elements; 16 bytes per thread.
__global__ void rnn3_fill_sentinel_kernel(__nv_bfloat16* base, size_t rows, int pitch, int valid) {
    const int per_row = pitch / 8;
    const size_t n = rows * (size_t)per_row;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % per_row) * 8;
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (c0 + 2 * k < valid ? 0xFFFFu : 0u) | (c0 + 2 * k + 1 < valid ? 0xFFFF0000u : 0u);
        reinterpret_cast<uint4*>(base)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
static int rnn3_fill_sentinel(__nv_bfloat16* base, size_t rows, int pitch, int valid, asrb_stream_t stream) {
    if ((pitch & 7) || (reinterpret_cast<uintptr_t>(base) & 15)) return ASRB_ERR_ALIGNMENT;
    const size_t n = rows * (size_t)(pitch / 8);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    rnn3_fill_sentinel_kernel<<<blocks, 256, 0, stream>>>(base, rows, pitch, valid);
    ASRB_CUDA_OK(cudaGetLastError());
    return 0;
}
// one word of every 32-byte sector of a landed operand tile [nkb][32 rows][128 B, 128-byte swizzle]: is any the sentinel?
// (a sector is written whole by one TMA store, so one word stands for it; word `sec` of chunk 2*sec keeps the 32 lanes of a
// warp on 32 different banks)
__device__ __forceinline__ bool rnn3_tile_has_sentinel(const uint8_t* tile, int nkb, int el, int nthreads) {
    const uint32_t base = smem_u32(tile);
    bool bad = false;
    for (int i = el; i < nkb * 128; i += nthreads) {
        const int kb = i >> 7, r = (i >> 2) & 31, sec = i & 3;
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)(kb * 4096 + r * 128 + (((2 * sec) ^ (r & 7)) << 4) + sec * 4)));
        bad = bad || (v == 0xFFFFFFFFu);
    }
    return bad;
}
static int rnn3_zero_pad_cols(__nv_bfloat16* base, size_t rows, int pitch, int col0, int ncols, asrb_stream_t stream) {
    if (ncols <= 0) return 0;
    if ((ncols | col0 | pitch) & 1) return ASRB_ERR_ALIGNMENT;
    const size_t n = rows * (size_t)(ncols / 2);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    rnn3_zero_pad_cols_kernel<<<blocks, 256, 0, stream>>>(base, rows, pitch, col0, ncols);
    ASRB_CUDA_OK(cudaGetLastError());
    return 0;
}

// Step hand-over of a launch (asrb_debug_rnn_dbg):
//   default: VERIFIED hand-over.  Pass 1 hands the operand tile over with a TMA store + its completion + a relaxed counter
//     increment.  That is fast -- a release is a MEMBAR.GPU that waits for the other chain's TMA copies -- but "completed" is
//     not "visible in L2": about one hand-over in 10^7 is consumed too early.  So every operand slab starts as a sentinel
//     (bf16 0xFFFF, a NaN no cell produces), each consumer looks for it in every tile that lands (one word per 32-byte
//     sector, in the shadow of the MMAs), and raises a flag.  Pass 2 is the same kernel with the release hand-over: it
//     leaves at once when the flag is down and recomputes the whole launch when it is up.  Both passes write the same
//     outputs from the same inputs, so the result is the release protocol's, bit for bit.
//     bit 8192 (TEST): CTA 0 withholds its tile of step 5 in pass 1, so pass 2 must run.
//   bit 4096: generic stores + red.release in one pass.  bit 16 / bit 4: the unverified TMA-store forms (experiments).
static inline bool rnn3_verified_handover() { return !(g_rnn_dbg & (4096 | 16 | 4)); }
static inline int rnn3_first_pass_dbg(bool verified) {
    if (verified) return (g_rnn_dbg & ~2) | kR3Check | ((g_rnn_dbg & 8192) ? kR3Sabotage : 0);
    return (g_rnn_dbg & (16 | 4)) ? (g_rnn_dbg & ~2) : (g_rnn_dbg | 2);      // release unless a TMA-store form is asked for
}
static inline int rnn3_second_pass_dbg() { return ((g_rnn_dbg | 2) & ~(16 | 4)) | kR3Redo; }

template <int CELL>
__global__ void __launch_bounds__(kRnnThreads, 1)
rnn_rec3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmGi,
                const __grid_constant__ CUtensorMap tmOp, const RnnParams p) {
    constexpr int NCH = 2;
    constexpr int kR3Rows = 64 / NCH;           // batch rows of a chain = N of its MMAs
    constexpr int kR3EpiWarps = 16 / NCH;       // per chain
    constexpr int kR3EpiThreads = kR3EpiWarps * 32;
    constexpr int NJ = 16;
    constexpr int kGates = (CELL == ASRB_RNN_GRU) ? 3 : 4;
    constexpr int NPAD = ((kGates * NJ + 15) / 16) * 16;      // rows of the packed forward slice
    constexpr int kTmemCols = 512;
    constexpr int KBE = 64;
    constexpr uint32_t kSlotBytes = kR3Rows * 128;            // one K block of a chain's operand: 32 rows x 128 B (swizzled)
    constexpr int kGiRegion = kR3Rows * 64;                   // one gate's [32 rows x 16 floats] box (2 KB = 4 swizzle atoms)
    constexpr int NV = NJ / 4;

    extern __shared__ uint8_t smem_raw[];
    const long long t_entry = clock64();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int T = p.T, B = p.B, H = p.H, G = p.G, P = p.P;
    const int nkb = p.kpad / KBE;
    const int nchunks = ceil_div(nkb, kR3Chunk);
    // per chain: operand tile | gate boxes (2 step parities) | the chain's output operand tile [rows][16 units] bf16 | barriers
    const size_t a_bytes = (size_t)nkb * kSlotBytes, gi_bytes = (size_t)2 * kGates * kGiRegion;
    const size_t op_bytes = (size_t)kR3Rows * NJ * 2;
    const size_t chain_bytes = (a_bytes + gi_bytes + op_bytes + 256 + 1023) & ~size_t(1023);   // operand tiles: 1 KB aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.x / P, pidx = blockIdx.x % P;
    const int j0 = pidx * NJ;
    const int chain = warp < kRnnCtrlWarps ? warp : (warp - kRnnCtrlWarps) / kR3EpiWarps;
    const int row0 = chain * kR3Rows;                         // first batch row of the chain
    const bool chain_on = chain < NCH && row0 < B;
    uint8_t* cs = smem + (size_t)(chain < NCH ? chain : 0) * chain_bytes;
    uint8_t* smem_a = cs;                                     // [nkb][32 rows x 128 B]
    uint8_t* smem_gi = cs + a_bytes;                          // [2][kGates][kGiRegion]
    __nv_bfloat16* st_op = reinterpret_cast<__nv_bfloat16*>(cs + a_bytes + gi_bytes);   // [rows][16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cs + a_bytes + gi_bytes + op_bytes);
    uint64_t* full_bar = bars;                                // [kR3MaxChunks]
    uint64_t* tfull_bar = bars + kR3MaxChunks;
    uint64_t* gi_bar = bars + kR3MaxChunks + 1;               // [2]
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NCH * chain_bytes);   // weights are in tensor memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
    float* s_bias = reinterpret_cast<float*>(w_bar + 2);      // [4][NJ]

    uint32_t* counter = p.counters + ((p.dbg & kR3Redo) ? kR3RedoCounters : 0) + (dir * NCH + (chain < NCH ? chain : 0)) * kR3CounterStride;
    auto t_of = [&](int s) { return dir == 1 ? (T - 1 - s) : s; };
    if (p.dbg & kR3Redo) {       // second pass of a verified launch: nothing to do unless the first pass consumed a stale sector
        if (ld_relaxed_gpu_u32(p.counters + kR3FlagWord) == 0u) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_rnn3_redos, 1ull);
    }

    if (warp < NCH && lane == 0) {
        if (warp == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmGi);
            tma_prefetch_desc(&tmOp);
            mbar_init(w_bar, 1);
        }
        for (int i = 0; i < kR3MaxChunks; ++i) mbar_init(&full_bar[i], 1);
        mbar_init(tfull_bar, 1);
        mbar_init(&gi_bar[0], 1);
        mbar_init(&gi_bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base + chain * kR3Rows;      // the chain's accumulator: kR3Rows columns
    const uint32_t tmem_w = tmem_base + NCH * kR3Rows;        // weights: kpad/2 columns

    if (warp < kRnnCtrlWarps) {
        // ===================== control warp of the chain: TMA copies, then the MMAs =====================
        //   D[64 gate rows, kR3Rows batch rows] = W[tensor memory] x h^T[shared memory]
        if (chain_on) {
            constexpr uint32_t idesc = umma_idesc(kFmtBF16, 64, kR3Rows);
            auto load_gi = [&](int s) {   // the slice's gate pre-activations of step s, rows of this chain -> buffer s & 1
                const int t = t_of(s);
                mbar_arrive_expect_tx(&gi_bar[s & 1], (uint32_t)(kGates * kR3Rows * NJ * 4));
                // one box for all gates: [gate][row][16 columns] (the tensor map walks the gates, H columns apart, as its third dimension)
                tma_load_4d(smem_gi + (size_t)((s & 1) * kGates) * kGiRegion, &tmGi, &gi_bar[s & 1], dir * G + j0, row0, 0, t);
            };
            // ONE elected thread runs the chain's whole control loop: poll, copies, MMAs.  (Per-chunk elect / __syncwarp
            // rounds of the whole warp cost ~250 cycles a chunk on the chain's critical path.)  The loop is kept SMALL: this
            // path runs once a step next to sixteen epilogue warps, and its instruction fetches sit on the critical path.
            if (elect_one()) {
                load_gi(0);
                if (T > 1) load_gi(1);
                mbar_wait(w_bar, 0);           // the weight slice is in tensor memory
                tc_fence_after_sync();
                const uint64_t bdesc0 = umma_desc_sw128(smem_u32(smem_a));
                const int first = nkb % kR3Chunk ? nkb % kR3Chunk : kR3Chunk;     // K blocks of the first chunk
                for (int s = 1; s < T; ++s) {
                    // step barrier of the chain: every CTA of this direction has published step s-1 of these rows
                    poll_counter(counter, (uint32_t)P * (uint32_t)s);
                    if (chain == 0) ASRB_TRACE(0, s);
                    if (chain == 1) ASRB_TRACE(12, s);
                    fence_proxy_async_global();      // (fallback path: the others' generic-proxy stores -> our async-proxy reads)
                    const int slab = dir * (T + 2) + t_of(s - 1) + 1;
                    // (our own arrival is part of `need`: our MMAs of step s-1 have read the tile, our epilogue its gate boxes)
                    for (int c = 0; c < nchunks; ++c) {
                        const int kb0 = c ? first + (c - 1) * kR3Chunk : 0, nblk = c ? kR3Chunk : first;
                        // ONE box per chunk: [4 K blocks][32 rows][128 B] (the tensor map walks the K blocks as its third
                        // dimension).  A TMA request costs ~60 cycles on top of its bytes.
                        // (tmA2: the same tensor with a box of nkb % 4 blocks -- the FIRST chunk: the small box lands soonest)
                        mbar_arrive_expect_tx(&full_bar[c], (uint32_t)nblk * kSlotBytes);
                        tma_load_4d(smem_a + (size_t)kb0 * kSlotBytes, nblk == kR3Chunk ? &tmA : &tmA2, &full_bar[c], 0, row0, kb0, slab);
                    }
                    if (s + 1 < T) load_gi(s + 1);
                    if (chain == 0) ASRB_TRACE(1, s);
                    const uint32_t ph = (uint32_t)((s - 1) & 1);
                    for (int c = 0; c < nchunks; ++c) {
                        const int kb0 = c ? first + (c - 1) * kR3Chunk : 0, nblk = c ? kR3Chunk : first;
                        mbar_wait(&full_bar[c], ph);
                        // (one tcgen05.fence per step, after the poll: the epilogue's reads of D are done.  None per chunk: the
                        // mbarrier wait orders the TMA's writes before the MMAs' reads.)
                        if (c == 0) {
                            if (chain == 0) ASRB_TRACE(2, s);
                            tc_fence_after_sync();
                        }
                        // constant offsets from the chunk's base descriptor / weight column, predicated instead of branched:
                        // straight-line code in the uniform datapath
                        const uint64_t cdesc = bdesc0 + (uint64_t)kb0 * (kSlotBytes >> 4);
                        const uint32_t cw = tmem_w + (uint32_t)kb0 * 32;
#pragma unroll
                        for (int i = 0; i < kR3Chunk; ++i) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // 8 TMEM columns of weights per K = 16 step
                                umma_f16_ts_if(tmem_d, cw + (i * 4 + k) * 8, cdesc + (uint64_t)(i * (kSlotBytes >> 4) + 2 * k), idesc, (c | i | k) != 0, i < nblk);
                        }
                    }
                    umma_commit(tfull_bar);
                    if (chain == 0) ASRB_TRACE(3, s);
                    if (chain == 1) ASRB_TRACE(13, s);
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps of the chain: 8 warps =====================
        const int wl = (warp - kRnnCtrlWarps) % kR3EpiWarps;
        const int quad = warp & 3, half = wl >> 2;             // TMEM lane quarter = unit group; batch columns 16*half .. +15
        const int r8 = lane >> 2, hi = r8 >> 2;                // accumulator row pair of the lane; hi: keeps the odd columns
        const int ug = quad, ul = r8 & 3;
        const int el = wl * 32 + lane;                         // 0..255 within the chain
        const int ju = 4 * ug + ul, unit = j0 + ju;
        const bool uvalid = unit < H;
        int rl[2], row[2], len[2];                             // our two cells: (unit, batch rows rl[0], rl[0] + 8 of the chain)
        bool cellok[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            rl[c] = 16 * half + 2 * (lane & 3) + hi + 8 * c;
            row[c] = row0 + rl[c];
            cellok[c] = uvalid && row[c] < B;
            len[c] = cellok[c] ? p.lengths[row[c]] : 0;
        }
        const size_t slotHB = (size_t)B * H;
        uint32_t gi_off[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) gi_off[c] = (uint32_t)(rl[c] * 64 + ((ug ^ ((rl[c] >> 1) & 3)) << 4) + ul * 4);
        const uint32_t smem_gi_u32 = smem_u32(smem_gi);

        // ---- once: zero boundary slots, biases, the weight slice -> tensor memory ----
#pragma unroll
        for (int c = 0; c < 2; ++c) {
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
        if (chain == 0) {
            for (int i = el; i < 4 * NJ; i += kR3EpiThreads) {
                const int g = i / NJ, jj = i % NJ;
                s_bias[i] = (g < kGates && j0 + jj < H) ? p.b_hh[(size_t)dir * G + g * H + j0 + jj] : 0.f;
            }
            if (wl < 4) {
                // TMEM lane 32 * quarter + 4 * gate + u holds packed row gate * 16 + 4 * quarter + u (the M = 64 data path
                // uses lanes 0..15 of each quarter); 16 bf16 = 8 columns per store
                const int g = lane >> 2, c = g * NJ + 4 * quad + (lane & 3);
                const bool have = lane < 16 && g < kGates && c < NPAD && pidx < p.P_saved;
                const uint4* wrow = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.wpack) + ((size_t)(dir * p.P_saved + (have ? pidx : 0)) * NPAD + (have ? c : 0)) * p.kpad);
                for (int k0 = 0; k0 < p.kpad / 16; k0 += 4) {       // kpad is a multiple of 64: four stores per round,
                    uint4 v[8];                                      // their eight loads in flight together
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = have ? __ldg(wrow + 2 * k0 + i) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t r[8] = {v[2 * i].x, v[2 * i].y, v[2 * i].z, v[2 * i].w, v[2 * i + 1].x, v[2 * i + 1].y, v[2 * i + 1].z, v[2 * i + 1].w};
                        tmem_st_32x8(tmem_w + (uint32_t(quad * 32) << 16) + (k0 + i) * 8, r);
                    }
                }
                tmem_st_wait();
            }
            tc_fence_before_sync();
        }
        named_bar_sync(3, kRnnEpiThreads);                      // all 16 epilogue warps
        if (warp == kRnnCtrlWarps && lane == 0) mbar_arrive(w_bar);
        float bias[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) bias[g] = s_bias[g * NJ + ju];

        if (chain_on) {
            float state_h[2] = {0.f, 0.f}, state_c[2] = {0.f, 0.f};
            // hand-over protocol (asrb_debug_rnn_dbg): the launcher SETS bit 2 (generic stores + red.release) unless bit 16
            // (TMA store + completion + relaxed increment) or bit 4 (the same + L2 read-back) asks for a TMA-store form
            const bool tma_op = !(p.dbg & 2);
            const bool readback = (p.dbg & 4) != 0;
            // per-step addresses advance by constants (t moves by +-1): bases for step 0, strides in elements
            const int t0 = t_of(0);
            const long long tstep = dir == 1 ? -1 : 1;
            float* hq[2];
            float* cq[2];
            float* svq[2];
            __nv_bfloat16* opq[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const size_t o = ((size_t)dir * (T + 2) + t0 + 1) * slotHB + (size_t)row[c] * H + unit;
                hq[c] = p.hseq + o;
                cq[c] = (CELL == ASRB_RNN_LSTM) ? p.cseq + o : nullptr;
                svq[c] = p.saved + ((((size_t)dir * T + t0) * p.P_saved + pidx) * 4) * (size_t)(NV * B * 4) + (size_t)ug * (B * 4) + (size_t)row[c] * 4 + ul;
                opq[c] = p.hbf + (((size_t)dir * (T + 2) + t0 + 1) * B + row[c]) * p.Hp + unit;
            }
            const long long h_stride = tstep * (long long)slotHB;
            const long long sv_stride = tstep * (long long)p.P_saved * 4 * (long long)(NV * B * 4);
            const long long op_stride = tstep * (long long)B * p.Hp;
            const size_t sv_gate = (size_t)NV * (B * 4);
            for (int s = 0; s < T; ++s) {
                const int t = t_of(s);
                bool active[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) active[c] = cellok[c] && (t < len[c]);
                if (el == 0 && chain == 0) ASRB_TRACE(4, s);
                if (el == 0 && chain == 1) ASRB_TRACE(14, s);
                if (s == 0 && el == 0 && chain == 0 && p.trace) p.trace[((size_t)blockIdx.x * p.T) * 16 + 15] = t_entry;
                if ((s == 0 || s == T - 1) && el == 0 && chain == 0 && p.trace)      // wall clock (ns) next to the SM clock: the real SM frequency
                    p.trace[((size_t)blockIdx.x * p.T + s) * 16 + 11] = (long long)globaltimer_ns();
                // gate pre-activations of this step (TMA, a step ahead): read BEFORE the wait for the accumulator -- the box
                // landed long ago, and the barrier probe + six shared-memory loads (~280 cycles) are off the critical path here
                mbar_wait(&gi_bar[s & 1], (uint32_t)((s >> 1) & 1));
                float in[kGates][2];
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int q = 0; q < kGates; ++q)
                        in[q][c] = active[c] ? ld_shared_f32(smem_gi_u32 + (uint32_t)(((s & 1) * kGates + q) * kGiRegion) + gi_off[c]) : 0.f;
                float acc[4][2];
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[g][0] = acc[g][1] = 0.f;
                if (s > 0) {
                    if (p.dbg & kR3Check) {
                        // verified hand-over: the operand tile has landed long before its MMAs finish -- look at it meanwhile
                        for (int c = 0; c < nchunks; ++c) mbar_wait(&full_bar[c], (uint32_t)((s - 1) & 1));
                        if (rnn3_tile_has_sentinel(smem_a, nkb, el, kR3EpiThreads)) atomicOr(p.counters + kR3FlagWord, 1u);
                    }
                    mbar_wait(tfull_bar, (uint32_t)((s - 1) & 1));
                    if (el == 0 && chain == 0) ASRB_TRACE(5, s);
                    tc_fence_after_sync();
                    float v[8];
                    tmem_ld_16x256b_x2(tmem_d + (uint32_t(quad * 32) << 16) + 16 * half, v);
                    tmem_ld_wait();
                    tc_fence_before_sync();    // the counter arrival below orders these reads before the next step's MMAs
                    // v[0..1]: row r8, columns 2*(lane&3) + {0, 1}; v[2..3]: row r8 + 8; v[4..7]: the same, columns + 8.
                    // lo lanes (r8 < 4) hold gates 0 / 2 and keep the even columns; hi lanes hold gates 1 / 3 and keep the odd.
                    // send what the partner keeps, keep ours: cell c uses column 2*(lane&3) + hi + 8*c
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float mine_a = hi ? v[4 * c + 1] : v[4 * c + 0];      // gate (hi ? 1 : 0) of our column
                        const float mine_b = hi ? v[4 * c + 3] : v[4 * c + 2];      // gate (hi ? 3 : 2)
                        const float send_a = hi ? v[4 * c + 0] : v[4 * c + 1];
                        const float send_b = hi ? v[4 * c + 2] : v[4 * c + 3];
                        const float got_a = __shfl_xor_sync(0xffffffffu, send_a, 16);   // partner's gate (hi ? 0 : 1) of our column
                        const float got_b = __shfl_xor_sync(0xffffffffu, send_b, 16);   // partner's gate (hi ? 2 : 3)
                        acc[0][c] = hi ? got_a : mine_a;
                        acc[1][c] = hi ? mine_a : got_a;
                        acc[2][c] = hi ? got_b : mine_b;
                        acc[3][c] = hi ? mine_b : got_b;
                    }
                    if (el == 0 && chain == 0) ASRB_TRACE(6, s);
                }
                float hn[2], cn[2], sv[4][2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
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
                            const float go = fsigmoid(in[3][c] + acc[3][c] + bias[3]);
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
                // (1) the next step's MMA operand, published through the chain's step counter.  It leaves as ONE TMA store of
                // the chain's [rows][16 units] tile, whose completion (cp.async.bulk.wait_group) the publishing thread waits for
                // before a RELAXED counter increment: a release here is a MEMBAR.GPU, and that was measured to wait for every
                // memory operation the SM has in flight -- the OTHER chain's TMA copies included (tools/ubench/membar_tma.cu:
                // 870 cycles alone, the rest of the copies' ~3 k when they are in flight), which chained the two chains together.
                // (rows >= B are clipped by the tensor map; units >= H hold 0 and land in the zero K padding.)
                if (tma_op) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) st_op[rl[c] * NJ + ju] = __float2bfloat16_rn(hn[c]);
                    fence_proxy_async_smem();
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if (cellok[c]) *opq[c] = __float2bfloat16_rn(hn[c]);
                }
                if (el == 0 && chain == 0) ASRB_TRACE(7, s);
                named_bar_sync(8 + chain, kR3EpiThreads);
                if (wl == 0) {          // the chain's first epilogue warp publishes
                    if (lane == 0 && chain == 0) ASRB_TRACE(8, s);
                    if (tma_op) {
                        if (lane == 0) {
                            if (!((p.dbg & kR3Sabotage) && blockIdx.x == 0 && s == 5))
                                tma_store_3d(&tmOp, st_op, j0, row0, dir * (T + 2) + t + 1);
                            bulk_commit_group();
                            bulk_wait_group<0>();        // the bulk store has "completed" -- which is NOT "its data is in L2"
                        }
                        __syncwarp();
                        if (readback && row0 + lane < B) {   // lane l: first word of row l of the tile, from L2, until it is there
                            const uint32_t want = *reinterpret_cast<const uint32_t*>(st_op + lane * NJ);
                            const __nv_bfloat16* g = p.hbf + (((size_t)dir * (T + 2) + t + 1) * B + row0 + lane) * p.Hp + j0;
                            uint32_t spins = 0;
                            while (ld_relaxed_gpu_u32(g) != want && ++spins < (1u << 22)) {}
                        }
                        __syncwarp();
                        if (lane == 0) red_relaxed_add_u32(counter, 1u);
                    } else if (lane == 0) {
                        red_release_add_u32(counter, 1u);
                    }
                    if (lane == 0 && chain == 0) ASRB_TRACE(10, s);
                }
                if (!(p.dbg & 1)) {     // (dbg bit 1: timing experiment without these stores, results incomplete)
                    // hold the other stores back until the operand tile has been read: they would sit in the LSU queue ahead of it
                    named_bar_sync(12 + chain, kR3EpiThreads);
                    // (2) the stores nobody waits for
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (cellok[c]) {
                            *hq[c] = hn[c];
                            if constexpr (CELL == ASRB_RNN_LSTM) *cq[c] = cn[c];
#pragma unroll
                            for (int q = 0; q < 4; ++q) svq[c][(size_t)q * sv_gate] = sv[q][c];
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    hq[c] += h_stride;
                    if constexpr (CELL == ASRB_RNN_LSTM) cq[c] += h_stride;
                    svq[c] += sv_stride;
                    opq[c] += op_stride;
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

template <int CELL>
static int rnn3_launch(const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    constexpr int kGates = (CELL == ASRB_RNN_GRU) ? 3 : 4;
    constexpr int NCH = 2, kR3Rows = 64 / NCH;
    const int kpad = pl.kpad_f, nkb = kpad / 64;
    const int B = prm.B;
    if (ceil_div(nkb, kR3Chunk) > kR3MaxChunks || NCH * kR3Rows + kpad / 2 > 512) return ASRB_ERR_UNSUPPORTED;
    prm.P_saved = pl.P;
    prm.P = pl.P;
    prm.kpad = kpad;
    prm.wpack = reinterpret_cast<const float*>(wpack);       // bf16 slices, read once into tensor memory
    prm.out_sum = nullptr;                                   // (the direction sum is a separate kernel: see asrb_rnn_fwd_sum)
    prm.stage_out = 0;
    const size_t chain_bytes = ((size_t)nkb * kR3Rows * 128 + (size_t)2 * kGates * kR3Rows * 64 + (size_t)kR3Rows * 16 * 2 + 256 + 1023) & ~size_t(1023);
    const size_t smem = 1024 + NCH * chain_bytes + 16 + 4 * 16 * 4 + 64;
    if (smem > (size_t)kRnnMaxSmem) return ASRB_ERR_UNSUPPORTED;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return ASRB_ERR_DRIVER;
    // (the sentinel check reads ONE word per 32-byte sector: a sector that is part K padding could show a zero there while its
    // valid columns are stale, so widths that are not whole sectors keep the one-pass release hand-over)
    const bool verified = rnn3_verified_handover() && prm.H % 16 == 0;
    CUtensorMap tmA, tmA2, tmGi, tmOp;
    {   // hbf [2(T+2)][B][Hp] bf16 seen as [slab][K block][row][64 columns]: box = 64 columns x rows of a chain x 4 K blocks,
        // 128-byte swizzle; rows >= B and K blocks >= Hp/64 are zero-filled.  The columns H..Hp of the last K block are inside
        // the tensor: they are zeroed once here (the weights' K padding is zero too, but 0 x NaN is NaN).
        uint64_t d[4] = {64, (uint64_t)B, (uint64_t)prm.Hp / 64, (uint64_t)2 * (prm.T + 2)};
        uint64_t s[3] = {(uint64_t)prm.Hp * 2, 128, (uint64_t)B * prm.Hp * 2};
        uint32_t bx[4] = {64, (uint32_t)kR3Rows, (uint32_t)kR3Chunk, 1};
        int rc = make_tmap_bf16(&tmA, prm.hbf, 4, d, s, bx);
        if (rc) return rc;
        bx[2] = nkb % kR3Chunk ? nkb % kR3Chunk : kR3Chunk;      // box of the first chunk
        rc = make_tmap_bf16(&tmA2, prm.hbf, 4, d, s, bx);
        if (rc) return rc;
        // verified hand-over: every operand slab starts as the sentinel (K padding zero); otherwise only the padding is written
        rc = verified ? rnn3_fill_sentinel(prm.hbf, (size_t)2 * (prm.T + 2) * B, prm.Hp, prm.H, stream)
                      : rnn3_zero_pad_cols(prm.hbf, (size_t)2 * (prm.T + 2) * B, prm.Hp, prm.H, prm.Hp - prm.H, stream);
        if (rc) return rc;
        // the same tensor as plain [slab][row][column] for the epilogue's operand stores: box = 16 units x rows of a chain
        cuuint64_t gdim[3] = {(cuuint64_t)prm.Hp, (cuuint64_t)B, (cuuint64_t)2 * (prm.T + 2)};
        cuuint64_t gstr[2] = {(cuuint64_t)prm.Hp * 2, (cuuint64_t)B * prm.Hp * 2};
        cuuint32_t bxo[3] = {16, (cuuint32_t)kR3Rows, 1}, es[3] = {1, 1, 1};
        if (enc(&tmOp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, prm.hbf, gdim, gstr, bxo, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return ASRB_ERR_TENSORMAP;
    }
    {   // gi [T][B][2G] fp32, box = 16 columns x 32 rows, 64-byte swizzle (conflict-free reads by (row, unit) threads)
        if ((reinterpret_cast<uintptr_t>(prm.gi) & 15) != 0) return ASRB_ERR_ALIGNMENT;
        // dims: column (within [dir*G + gate*H + unit]), batch row, gate (H columns apart), time.  The column extent is 2G - (gates-1)H
        // so that column + gate*H never leaves the row.
        cuuint64_t gdim[4] = {(cuuint64_t)2 * prm.G - (cuuint64_t)(kGates - 1) * prm.H, (cuuint64_t)B, (cuuint64_t)kGates, (cuuint64_t)prm.T};
        cuuint64_t gstr[3] = {(cuuint64_t)2 * prm.G * 4, (cuuint64_t)prm.H * 4, (cuuint64_t)B * 2 * prm.G * 4};
        cuuint32_t bx[4] = {16, (cuuint32_t)kR3Rows, (cuuint32_t)kGates, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmGi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(prm.gi), gdim, gstr, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return ASRB_ERR_TENSORMAP;
    }
    auto kern = rnn_rec3_kernel<CELL>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {   // the step barrier spins: refuse the launch when the device cannot hold the whole grid at once
        int dev = 0, sms = 0, per_sm = 0;
        ASRB_CUDA_OK(cudaGetDevice(&dev));
        ASRB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        ASRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRnnThreads, smem));
        if (2 * pl.P > sms * per_sm) return ASRB_ERR_UNSUPPORTED;
    }
    ASRB_CUDA_OK(cudaMemsetAsync(prm.counters, 0, 128 * sizeof(uint32_t), stream));     // both passes' counters and the flag word
    prm.dbg = rnn3_first_pass_dbg(verified);
    kern<<<dim3(2 * pl.P), dim3(kRnnThreads), smem, stream>>>(tmA, tmA2, tmGi, tmOp, prm);
    ASRB_CUDA_OK(cudaGetLastError());
    if (verified) {      // second pass: the release hand-over, run only if the first pass saw a sentinel (else it leaves at once)
        prm.dbg = rnn3_second_pass_dbg();
        kern<<<dim3(2 * pl.P), dim3(kRnnThreads), smem, stream>>>(tmA, tmA2, tmGi, tmOp, prm);
        ASRB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

// ================================================================================================
// backward recurrence: dh_{t-1} = dgates_t W_hh, weights in tensor memory, K split over a cluster of four CTAs,
// two interleaved half-batch chains
// ================================================================================================
// A cluster of 4 CTAs shares 64 hidden units (16 each).  CTA rank r keeps, for ALL 64 units, the r-th quarter of the gate
// index K (the layout rnn_pack_bwd_split_kernel writes for ks = 4): A = [M = 64 units, K = kpad] in tensor memory.  Per
// step and chain it copies only its quarter of the chain's gate-gradient rows (32 rows x kpad), multiplies, and the four
// 16-unit quarters of the transposed accumulator D[64 units, 32 rows] go where they belong: quarter r stays, the other
// three travel to their owners through distributed shared memory (one 2.5 KB bulk copy each, completing on the owner's
// mbarrier); every CTA then adds the four partial sums of its own 16 units.  The exchange latency (~1.2-2 k cycles)
// sits in the shadow of the other chain like the step hand-over does.
constexpr int kR3XStride = 40;                              // floats per unit row of an exchanged quarter [16 units][32 rows + pad]
constexpr int kR3XBytes = 16 * kR3XStride * 4;              // 2560
constexpr int kR3StageBytes = 32 * 16 * 2;                  // one gate's [32 rows x 16 units] bf16 tile of a staged output

template <int CELL>
__global__ void __launch_bounds__(kRnnThreads, 1)
rnn_rec3_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmDgi,
                    const __grid_constant__ CUtensorMap tmGT, const __grid_constant__ CUtensorMap tmHT,
                    const __grid_constant__ CUtensorMap tmOp, const RnnParams p) {
    constexpr int NCH = 2, KS = 4;
    constexpr int kRows = 64 / NCH;             // batch rows of a chain = N of its MMAs
    constexpr int kEpiWarps = 16 / NCH, kEpiThreads = kEpiWarps * 32;
    constexpr int NJ = 16, NV = NJ / 4;
    constexpr int kGates = (CELL == ASRB_RNN_GRU) ? 3 : 4;
    constexpr int kTmemCols = 512;
    constexpr int KBE = 64;
    constexpr uint32_t kSlotBytes = kRows * 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int T = p.T, B = p.B, H = p.H, G = p.G, P = p.P;
    const int nkb = p.kpad / KBE;
    const int nchunks = ceil_div(nkb, kR3Chunk);
    // per chain: operand tile | own quarter [16][40] | staging [2 parities][4 peers][16][40] | receive [2][4 sources][16][40] | barriers
    const size_t a_bytes = (size_t)nkb * kSlotBytes;
    const size_t x_bytes = (size_t)(1 + 2 * KS + 2 * KS) * kR3XBytes;
    // staged outputs: dgi, dgiT, dghT tiles of every gate; GRU: plus the operand tiles (LSTM: the operand IS the dgi tile)
    constexpr int kStTiles = (CELL == ASRB_RNN_GRU) ? 4 : 3;
    const size_t st_bytes = (size_t)kStTiles * kGates * kR3StageBytes;
    const size_t chain_bytes = (a_bytes + x_bytes + st_bytes + 256 + 1023) & ~size_t(1023);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.x / P, pidx = blockIdx.x % P;
    const int j0 = pidx * NJ;
    const uint32_t crank = cluster_ctarank();                  // = pidx % 4: our K quarter, and the accumulator quarter we own
    const int chain = warp < kRnnCtrlWarps ? warp : (warp - kRnnCtrlWarps) / kEpiWarps;
    const int row0 = chain * kRows;
    const bool chain_on = chain < NCH && row0 < B;
    uint8_t* cs = smem + (size_t)(chain < NCH ? chain : 0) * chain_bytes;
    uint8_t* smem_a = cs;
    float* dt = reinterpret_cast<float*>(cs + a_bytes);                       // our own partial sums [16 units][kR3XStride]
    float* xs = dt + 16 * kR3XStride;                                         // [2][KS][16][kR3XStride] staging, by peer
    float* xr = xs + 2 * KS * 16 * kR3XStride;                                // [2][KS][16][kR3XStride] received, by source
    // staged outputs (TMA stores): [gate][32 rows][16 units] for dgi, [gate][16 units][32 rows] for the transposed copies
    __nv_bfloat16* st_dgi = reinterpret_cast<__nv_bfloat16*>(cs + a_bytes + x_bytes);
    __nv_bfloat16* st_gT = st_dgi + kGates * (kR3StageBytes / 2);
    __nv_bfloat16* st_hT = st_gT + kGates * (kR3StageBytes / 2);
    __nv_bfloat16* st_op = (CELL == ASRB_RNN_GRU) ? st_hT + kGates * (kR3StageBytes / 2) : st_dgi;   // [gate][32 rows][16 units]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cs + a_bytes + x_bytes + st_bytes);
    uint64_t* full_bar = bars;                                // [kR3MaxChunks]
    uint64_t* tfull_bar = bars + kR3MaxChunks;
    uint64_t* x_bar = bars + kR3MaxChunks + 1;                // [2] the peers' quarters of step parity 0 / 1 have arrived
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NCH * chain_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
    uint32_t* mma_lock = tmem_slot + 1;                       // tensor pipe: one chain's MMA sequence at a time

    uint32_t* counter = p.counters + ((p.dbg & kR3Redo) ? kR3RedoCounters : 0) + (dir * NCH + (chain < NCH ? chain : 0)) * kR3CounterStride;
    if (p.dbg & kR3Redo) {       // second pass of a verified launch (see the forward kernel); every CTA of every cluster leaves
        if (ld_relaxed_gpu_u32(p.counters + kR3FlagWord) == 0u) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_rnn3_redos, 1ull);
    }
    const int kcol0 = (int)crank * p.kpad;                    // first gate column of our K quarter
    auto t_of = [&](int s) { return dir == 0 ? (T - 1 - s) : s; };     // the backward pass walks each direction in reverse

    if (warp < NCH && lane == 0) {
        if (warp == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmDgi);
            tma_prefetch_desc(&tmGT);
            tma_prefetch_desc(&tmHT);
            tma_prefetch_desc(&tmOp);
            mbar_init(w_bar, 1);
            *mma_lock = 0u;
        }
        for (int i = 0; i < kR3MaxChunks; ++i) mbar_init(&full_bar[i], 1);
        mbar_init(tfull_bar, 1);
        mbar_init(&x_bar[0], 1);
        mbar_init(&x_bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    cluster_sync_all();                      // the peers' exchange barriers exist before anybody sends
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base + chain * kRows;
    const uint32_t tmem_w = tmem_base + NCH * kRows;

    if (warp < kRnnCtrlWarps) {
        // ===================== control warp of the chain: TMA copies, then the MMAs =====================
        //   D[64 units of the cluster, kRows batch rows] = W_hh^T[our K quarter, tensor memory] x dgates^T[shared memory]
        if (chain_on) {
            constexpr uint32_t idesc = umma_idesc(kFmtBF16, 64, kRows);
            if (elect_one()) {               // one thread runs the chain's control loop (see the forward kernel)
                mbar_wait(w_bar, 0);
                tc_fence_after_sync();
                const uint64_t bdesc0 = umma_desc_sw128(smem_u32(smem_a));
                const int first = nkb % kR3Chunk ? nkb % kR3Chunk : kR3Chunk;     // K blocks of the first chunk
                const bool use_lock = B > kRows && !(p.dbg & 32768);
                for (int s = 1; s < T; ++s) {
                    poll_counter(counter, (uint32_t)P * (uint32_t)s);
                    if (chain == 0) ASRB_TRACE(0, s);
                    fence_proxy_async_global();
                    const int slab = dir * T + t_of(s - 1);
                    for (int c = 0; c < nchunks; ++c) {      // one box per chunk of 4 K blocks (see the forward kernel)
                        const int kb0 = c ? first + (c - 1) * kR3Chunk : 0, nblk = c ? kR3Chunk : first;
                        mbar_arrive_expect_tx(&full_bar[c], (uint32_t)nblk * kSlotBytes);
                        tma_load_4d(smem_a + (size_t)kb0 * kSlotBytes, nblk == kR3Chunk ? &tmA : &tmA2, &full_bar[c], 0, row0, kcol0 / KBE + kb0, slab);
                    }
                    if (chain == 0) ASRB_TRACE(1, s);
                    const uint32_t ph = (uint32_t)((s - 1) & 1);
                    for (int c = 0; c < nchunks; ++c) {
                        const int kb0 = c ? first + (c - 1) * kR3Chunk : 0, nblk = c ? kR3Chunk : first;
                        mbar_wait(&full_bar[c], ph);
                        if (c == 0) {
                            if (chain == 0) ASRB_TRACE(2, s);
                            // one chain's MMA sequence at a time: interleaved in the tensor pipe both chains finish late, and
                            // stay in lockstep; first come first served lets one run ahead until the phases no longer overlap
                            if (use_lock) while (atomicCAS(mma_lock, 0u, 1u) != 0u) {}
                        }
                        if (c == 0) tc_fence_after_sync();
                        const uint64_t cdesc = bdesc0 + (uint64_t)kb0 * (kSlotBytes >> 4);
                        const uint32_t cw = tmem_w + (uint32_t)kb0 * 32;
#pragma unroll
                        for (int i = 0; i < kR3Chunk; ++i) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16_ts_if(tmem_d, cw + (i * 4 + k) * 8, cdesc + (uint64_t)(i * (kSlotBytes >> 4) + 2 * k), idesc, (c | i | k) != 0, i < nblk);
                        }
                    }
                    umma_commit(tfull_bar);
                    if (use_lock) atomicExch(mma_lock, 0u);
                    if (chain == 0) ASRB_TRACE(3, s);
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps of the chain =====================
        // accumulator role: lane quarter q = warp % 4 holds the 16 units of cluster rank q, this warp takes the chain's
        // batch columns 16*half .. +15.  cell role: thread = unit 4*ug + lane%4 of two batch rows of the chain.
        constexpr int kRp16 = kRows / 16;
        const int wl = (warp - kRnnCtrlWarps) % kEpiWarps;
        const int quad = warp & 3, half = wl >> 2;
        const int ug = wl / kRp16, ul = lane & 3;
        const int el = wl * 32 + lane;
        const int ju = 4 * ug + ul, unit = j0 + ju;
        const bool uvalid = unit < H;
        int row[2], len[2];
        bool cellok[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            row[c] = row0 + 16 * (wl % kRp16) + (lane >> 2) + 8 * c;
            cellok[c] = uvalid && row[c] < B;
            len[c] = cellok[c] ? p.lengths[row[c]] : 0;
        }
        const size_t slotHB = (size_t)B * H;

        // ---- once: the weight quarter -> tensor memory ----
        if (chain == 0) {
            if (wl < 4) {
                // row c of the packed slice = unit c of the cluster; TMEM lane 32 * quarter + i holds row 16 * quarter + i
                const int c = 16 * quad + lane;
                const bool have = lane < 16;
                const uint4* wrow = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.wpack) + ((size_t)(dir * P + pidx) * 64 + (have ? c : 0)) * p.kpad);
                for (int k0 = 0; k0 < p.kpad / 16; k0 += 4) {       // kpad is a multiple of 64: four stores per round,
                    uint4 v[8];                                      // their eight loads in flight together
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = have ? __ldg(wrow + 2 * k0 + i) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t r[8] = {v[2 * i].x, v[2 * i].y, v[2 * i].z, v[2 * i].w, v[2 * i + 1].x, v[2 * i + 1].y, v[2 * i + 1].z, v[2 * i + 1].w};
                        tmem_st_32x8(tmem_w + (uint32_t(quad * 32) << 16) + (k0 + i) * 8, r);
                    }
                }
                tmem_st_wait();
            }
            tc_fence_before_sync();
        }
        named_bar_sync(3, kRnnEpiThreads);
        if (warp == kRnnCtrlWarps && lane == 0) mbar_arrive(w_bar);

        if (chain_on) {
            float state_h[2] = {0.f, 0.f}, state_c[2] = {0.f, 0.f};   // direct dh / dc carries
            uint32_t xphase[2] = {0u, 0u};
            // The outputs only later kernels read (gate gradients row-major and transposed: 18 scattered 2-byte stores per
            // thread and step, ~2 k cycles of LSU issue time per chain that the trace showed on the chain's critical path)
            // go through shared-memory tiles and TMA stores when the tiles are whole: 32 batch rows, 16 valid units.
            const bool staged = p.stage_out && j0 + NJ <= H && row0 + kRows <= B;
            const int rl0 = 16 * (wl % kRp16) + (lane >> 2);      // first of our two rows within the chain
            for (int s = 0; s < T; ++s) {
                const int t = t_of(s);
                bool active[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) active[c] = cellok[c] && (t < len[c]);
                if (el == 0 && chain == 0) ASRB_TRACE(4, s);
                if (el == 0 && s > 0) mbar_arrive_expect_tx(&x_bar[s & 1], (uint32_t)((KS - 1) * kR3XBytes));
                // ---- operand prefetch (independent of the recurrent product) ----
                float in[6][2], ct[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) in[q][c] = 0.f;
                    ct[c] = 0.f;
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
                float acc[2] = {0.f, 0.f};
                if (s > 0) {
                    const int par = s & 1;
                    if (p.dbg & kR3Check) {      // verified hand-over: see the forward kernel
                        for (int c = 0; c < nchunks; ++c) mbar_wait(&full_bar[c], (uint32_t)((s - 1) & 1));
                        if (rnn3_tile_has_sentinel(smem_a, nkb, el, kEpiThreads)) atomicOr(p.counters + kR3FlagWord, 1u);
                    }
                    mbar_wait(tfull_bar, (uint32_t)((s - 1) & 1));
                    if (el == 0 && chain == 0) ASRB_TRACE(5, s);
                    tc_fence_after_sync();
                    {
                        float v[8];
                        tmem_ld_16x256b_x2(tmem_d + (uint32_t(quad * 32) << 16) + 16 * half, v);
                        tmem_ld_wait();
                        // quarter `quad` of the accumulator belongs to cluster rank `quad`: ours stays, the others are staged
                        float* dst = ((uint32_t)quad == crank ? dt : xs + ((size_t)par * KS + quad) * 16 * kR3XStride) +
                                     (lane >> 2) * kR3XStride + 16 * half + 2 * ul;
                        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
                        *reinterpret_cast<float2*>(dst + 8 * kR3XStride) = make_float2(v[2], v[3]);
                        *reinterpret_cast<float2*>(dst + 8) = make_float2(v[4], v[5]);
                        *reinterpret_cast<float2*>(dst + 8 * kR3XStride + 8) = make_float2(v[6], v[7]);
                    }
                    tc_fence_before_sync();
                    fence_proxy_async_smem();                  // generic stores -> the bulk copies' async-proxy reads
                    if ((uint32_t)quad != crank) {
                        // the two warps of a quarter sync among themselves (64 threads) and send it: no wait for the other quarters
                        constexpr int kQBar[2][4] = {{1, 2, 6, 7}, {10, 11, 14, 15}};
                        named_bar_sync(kQBar[chain][quad], 64);
                        if (half == 0 && lane == 0) {
                            // lands in slot [par][source = our rank] of the owner's receive buffer of this chain
                            dsmem_bulk_copy(map_to_cta(xr + ((size_t)par * KS + crank) * 16 * kR3XStride, (uint32_t)quad),
                                            xs + ((size_t)par * KS + quad) * 16 * kR3XStride, (uint32_t)kR3XBytes,
                                            map_to_cta(&x_bar[par], (uint32_t)quad));
                        }
                    }
                    if (staged && el == 0) bulk_wait_group_read<0>();   // the previous step's TMA stores have read their tiles
                    named_bar_sync(4 + chain, kEpiThreads);    // our own quarter is in `dt`
                    if (el == 0 && chain == 0) ASRB_TRACE(6, s);
                    mbar_wait_cluster(&x_bar[par], xphase[par]);
                    xphase[par] ^= 1u;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int o = ju * kR3XStride + (row[c] - row0);
                        float a = dt[o];
#pragma unroll
                        for (int q = 1; q < KS; ++q) a += xr[((size_t)par * KS + (crank + q) % KS) * 16 * kR3XStride + o];
                        acc[c] = a;
                    }
                    if (el == 0 && chain == 0) ASRB_TRACE(11, s);
                }

                // ---- cell math ----
                float dg[4][2], eg2[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float carry = acc[c] + state_h[c];
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
                // All 2-byte outputs go out as 4-byte pairs: neighbouring lanes swap one value so that a thread holds two
                // adjacent elements -- lane pairs (ul, ul^1) for the row-major tensors (two units of one row: the even lane
                // takes row 0 of the pair's cells, the odd lane row 1), lane pairs (r, r^1) for the transposed ones (two rows
                // of one unit).  Halves the store instructions of the busiest phase of the step.
                const bool ev_u = (ul & 1) == 0, ev_r = ((lane >> 2) & 1) == 0;
                auto pair_units = [&](float c0, float c1) {          // -> bf16x2 of (unit, unit+1) for our store row
                    const float got = __shfl_xor_sync(0xffffffffu, ev_u ? c1 : c0, 1);
                    const __nv_bfloat162 v = ev_u ? __floats2bfloat162_rn(c0, got) : __floats2bfloat162_rn(got, c1);
                    return *reinterpret_cast<const uint32_t*>(&v);
                };
                auto pair_rows = [&](float c0, float c1) {           // -> bf16x2 of (row, row+1) for our store unit
                    const float got = __shfl_xor_sync(0xffffffffu, ev_r ? c1 : c0, 4);
                    const __nv_bfloat162 v = ev_r ? __floats2bfloat162_rn(c0, got) : __floats2bfloat162_rn(got, c1);
                    return *reinterpret_cast<const uint32_t*>(&v);
                };
                const int cu = ev_u ? 0 : 1;                          // which of our two cells' rows the unit pair is stored for
                const int unit2 = unit & ~1;
                // (1) the next step's MMA operand (hidden-side gate gradients), published through the chain's step counter:
                // with whole tiles, one TMA store + its completion + a relaxed increment instead of stores + a release (see the
                // forward kernel: the release's MEMBAR waits for the other chain's TMA copies)
                const bool tma_op = staged && !(p.dbg & 2);
                const bool readback = (p.dbg & 4) != 0;
                uint32_t vu[kGates];                                   // (unit, unit+1) pairs of the input-side gate gradients
#pragma unroll
                for (int q = 0; q < kGates; ++q) vu[q] = pair_units(dg[q][0], dg[q][1]);
                const uint32_t ve2 = (CELL == ASRB_RNN_GRU) ? pair_units(eg2[0], eg2[1]) : vu[2];   // gate 2 of the hidden side
                const int rl_u = rl0 + 8 * cu;                                  // our row of the unit pairs
                if (tma_op) {
                    // phase A, on the chain's critical path: only the operand tile ([gate][32 rows][16 units]; LSTM: it IS the
                    // dgi tile).  The other output tiles are filled after the hand-over, in the shadow of its latency.
                    uint32_t* s_op = reinterpret_cast<uint32_t*>(st_op);
#pragma unroll
                    for (int q = 0; q < kGates; ++q) s_op[(q * (kR3StageBytes / 2) + rl_u * NJ + (ju & ~1)) / 2] = (q == 2) ? ve2 : vu[q];
                    fence_proxy_async_smem();                  // generic stores -> the TMA store's async-proxy reads
                } else {
                    uint32_t* o = reinterpret_cast<uint32_t*>(p.dghbf + (((size_t)dir * T + t) * B + row[cu]) * p.Gp + unit2);
#pragma unroll
                    for (int q = 0; q < kGates; ++q)
                        if (cellok[cu]) o[(size_t)q * H / 2] = (q == 2) ? ve2 : vu[q];
                }
                if (el == 0 && chain == 0) ASRB_TRACE(7, s);
                named_bar_sync(8 + chain, kEpiThreads);
                if (wl == 0) {          // the chain's first epilogue warp publishes
                    if (lane == 0 && chain == 0) ASRB_TRACE(8, s);
                    if (tma_op) {
                        if (lane == 0) {
                            if (!((p.dbg & kR3Sabotage) && blockIdx.x == 0 && s == 5))
                                tma_store_4d(&tmOp, st_op, j0, row0, 0, dir * T + t);
                            bulk_commit_group();
                            bulk_wait_group<0>();
                        }
                        __syncwarp();
                        if (readback) {      // lane l: first word of row l of every gate's tile, from L2, until they are there
                            const __nv_bfloat16* g = p.dghbf + (((size_t)dir * T + t) * B + row0 + lane) * p.Gp + j0;
                            uint32_t want[kGates];
#pragma unroll
                            for (int q = 0; q < kGates; ++q)
                                want[q] = *reinterpret_cast<const uint32_t*>(st_op + q * (kR3StageBytes / 2) + lane * NJ);
                            uint32_t spins = 0;
                            bool ok = false;
                            while (!ok && ++spins < (1u << 22)) {
                                uint32_t got[kGates];
#pragma unroll
                                for (int q = 0; q < kGates; ++q) got[q] = ld_relaxed_gpu_u32(g + (size_t)q * H);
                                ok = true;
#pragma unroll
                                for (int q = 0; q < kGates; ++q) ok = ok && (got[q] == want[q]);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) red_relaxed_add_u32(counter, 1u);
                    } else if (lane == 0) {
                        red_release_add_u32(counter, 1u);
                    }
                    if (lane == 0 && chain == 0) ASRB_TRACE(10, s);
                }
                if (staged) {
                    // phase B: the tiles only later kernels read
                    const int rl_r = ev_r ? rl0 : rl0 - 1 + 8;                      // first row of our row pair
                    uint32_t* s_dgi = reinterpret_cast<uint32_t*>(st_dgi);
                    uint32_t* s_gT = reinterpret_cast<uint32_t*>(st_gT);
                    uint32_t* s_hT = reinterpret_cast<uint32_t*>(st_hT);
#pragma unroll
                    for (int q = 0; q < kGates; ++q) {
                        if (CELL == ASRB_RNN_GRU || !tma_op) s_dgi[(q * (kR3StageBytes / 2) + rl_u * NJ + (ju & ~1)) / 2] = vu[q];
                        const uint32_t vt = pair_rows(dg[q][0], dg[q][1]);
                        s_gT[(q * (kR3StageBytes / 2) + ju * kRows + rl_r) / 2] = vt;
                        if (p.dghT) s_hT[(q * (kR3StageBytes / 2) + ju * kRows + rl_r) / 2] = (q == 2) ? pair_rows(eg2[0], eg2[1]) : vt;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(12 + chain, kEpiThreads);
                    if (el == 0) {
                        // one store per tensor: the tiles of all gates (the tensor maps walk the gates as an extra dimension)
                        tma_store_4d(&tmDgi, st_dgi, dir * G + j0, row0, 0, t);
                        tma_store_3d(&tmGT, st_gT, t * B + row0, dir * G + j0, 0);
                        if (p.dghT) tma_store_3d(&tmHT, st_hT, t * B + row0, dir * G + j0, 0);
                        bulk_commit_group();
                    }
                }
                if (!staged) {
                    named_bar_sync(12 + chain, kEpiThreads);      // the release's MEMBAR waits for every store in flight: rest after it
                    // (2) outputs only later kernels read
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
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
                                if (hT) hT[(size_t)q * H * p.ldT] = (q == 2) ? __float2bfloat16_rn(eg2[c]) : v;
                            }
                        }
                    }
                }
            }
            if (staged && el == 0) bulk_wait_group<0>();      // the last TMA stores have been written
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
    cluster_sync_all();   // nobody leaves while a peer may still write into its receive buffers
}

template <int CELL>
static int rnn3_bwd_launch(const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    const int kpad = pl.kpad_b, nkb = kpad / 64;
    const int B = prm.B;
    if (pl.ksplit != 4 || pl.npad_b != 64 || ceil_div(nkb, kR3Chunk) > kR3MaxChunks || 64 + kpad / 2 > 512) return ASRB_ERR_UNSUPPORTED;
    prm.P_saved = pl.P;
    prm.P = pl.P_b;
    prm.kpad = kpad;
    prm.wpack = reinterpret_cast<const float*>(wpack);
    constexpr int kGates = (CELL == ASRB_RNN_GRU) ? 3 : 4;
    const size_t chain_bytes = ((size_t)nkb * 32 * 128 + (size_t)(1 + 4 * 4) * kR3XBytes + (size_t)(CELL == ASRB_RNN_GRU ? 4 : 3) * kGates * kR3StageBytes + 256 + 1023) & ~size_t(1023);
    const size_t smem = 1024 + 2 * chain_bytes + 64;
    if (smem > (size_t)kRnnMaxSmem) return ASRB_ERR_UNSUPPORTED;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return ASRB_ERR_DRIVER;
    CUtensorMap tmA, tmA2, tmDgi, tmGT, tmHT, tmOp;
    // staged outputs: whole 32-row x 16-unit tiles only, and 16-byte aligned tile rows in the transposed copies
    prm.stage_out = (B % 32 == 0 && prm.H % 16 == 0 && !(g_rnn_dbg & 2048)) ? 1 : 0;
    bool verified = rnn3_verified_handover();
    if (prm.stage_out) {
        auto plain = [&](CUtensorMap* m, void* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstr, const cuuint32_t* bx) {
            cuuint32_t es[4] = {1, 1, 1, 1};
            if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
            return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        bool ok = true;
        {   // dgi [T][B][2G] bf16 as (column, row, gate, time): box = 16 columns x 32 rows x all gates
            cuuint64_t gdim[4] = {(cuuint64_t)2 * prm.G - (cuuint64_t)(kGates - 1) * prm.H, (cuuint64_t)B, (cuuint64_t)kGates, (cuuint64_t)prm.T};
            cuuint64_t gstr[3] = {(cuuint64_t)2 * prm.G * 2, (cuuint64_t)prm.H * 2, (cuuint64_t)B * 2 * prm.G * 2};
            cuuint32_t bx[4] = {16, 32, (cuuint32_t)kGates, 1};
            ok = ok && plain(&tmDgi, prm.dgi, 4, gdim, gstr, bx);
        }
        {   // dgiT / dghT [2G][ldT] bf16 as (column, unit row, gate): box = 32 columns (batch rows of one time step) x 16 rows x all gates
            cuuint64_t gdim[3] = {(cuuint64_t)prm.ldT, (cuuint64_t)2 * prm.G - (cuuint64_t)(kGates - 1) * prm.H, (cuuint64_t)kGates};
            cuuint64_t gstr[2] = {(cuuint64_t)prm.ldT * 2, (cuuint64_t)prm.H * prm.ldT * 2};
            cuuint32_t bx[3] = {32, 16, (cuuint32_t)kGates};
            ok = ok && plain(&tmGT, prm.dgiT, 3, gdim, gstr, bx);
            ok = ok && plain(&tmHT, prm.dghT ? prm.dghT : prm.dgiT, 3, gdim, gstr, bx);
        }
        {   // dghbf [2 T][B][Gp] bf16 (the recurrence's own operand) as (column, row, gate, slab): box = 16 columns x 32 rows x all gates
            cuuint64_t gdim[4] = {(cuuint64_t)prm.Gp - (cuuint64_t)(kGates - 1) * prm.H, (cuuint64_t)B, (cuuint64_t)kGates, (cuuint64_t)2 * prm.T};
            cuuint64_t gstr[3] = {(cuuint64_t)prm.Gp * 2, (cuuint64_t)prm.H * 2, (cuuint64_t)B * prm.Gp * 2};
            cuuint32_t bx[4] = {16, 32, (cuuint32_t)kGates, 1};
            ok = ok && plain(&tmOp, prm.dghbf, 4, gdim, gstr, bx);
        }
        if (!ok) prm.stage_out = 0;
    }
    if (!prm.stage_out) tmDgi = tmGT = tmHT = tmOp = CUtensorMap{};
    if (!prm.stage_out) verified = false;       // without whole tiles every CTA hands over with a release anyway
    {   // dghbf [2 T][B][Gp] bf16 seen as [slab][K block][row][64 columns]: box = 64 columns x 32 rows (one chain) x 4 K blocks
        uint64_t d[4] = {64, (uint64_t)B, (uint64_t)prm.Gp / 64, (uint64_t)2 * prm.T};
        uint64_t s[3] = {(uint64_t)prm.Gp * 2, 128, (uint64_t)B * prm.Gp * 2};
        uint32_t bx[4] = {64, 32, (uint32_t)kR3Chunk, 1};
        int rc = make_tmap_bf16(&tmA, prm.dghbf, 4, d, s, bx);
        if (rc) return rc;
        bx[2] = nkb % kR3Chunk ? nkb % kR3Chunk : kR3Chunk;      // box of the last chunk of a K quarter
        rc = make_tmap_bf16(&tmA2, prm.dghbf, 4, d, s, bx);
        if (rc) return rc;
        // the columns G..Gp of the last K block are inside the tensor: zero them once (0 x NaN is NaN)
        // (verified hand-over: every operand slab starts as the sentinel, see rnn3_launch)
        rc = verified ? rnn3_fill_sentinel(prm.dghbf, (size_t)2 * prm.T * B, prm.Gp, prm.G, stream)
                      : rnn3_zero_pad_cols(prm.dghbf, (size_t)2 * prm.T * B, prm.Gp, prm.G, prm.Gp - prm.G, stream);
        if (rc) return rc;
    }
    auto kern = rnn_rec3_bwd_kernel<CELL>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ASRB_CUDA_OK(cudaMemsetAsync(prm.counters, 0, 128 * sizeof(uint32_t), stream));     // both passes' counters and the flag word
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pl.P_b);
    cfg.blockDim = dim3(kRnnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 4; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    {   // the step barrier spins: refuse the launch when the device cannot hold the whole grid (in clusters of 4) at once
        int nclusters = 0;
        ASRB_CUDA_OK(cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg));
        if (2 * pl.P_b > 4 * nclusters) return ASRB_ERR_UNSUPPORTED;
    }
    prm.dbg = rnn3_first_pass_dbg(verified);
    ASRB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmA2, tmDgi, tmGT, tmHT, tmOp, prm));
    if (verified) {      // second pass, see rnn3_launch
        prm.dbg = rnn3_second_pass_dbg();
        ASRB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmA2, tmDgi, tmGT, tmHT, tmOp, prm));
    }
    return 0;
}

long long rnn3_redo_count() {
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_rnn3_redos, sizeof(v)) != cudaSuccess) return -1;
    return (long long)v;
}

int rnn3_backward(int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    if (cell == ASRB_RNN_GRU) return rnn3_bwd_launch<ASRB_RNN_GRU>(pl, prm, wpack, stream);
    return rnn3_bwd_launch<ASRB_RNN_LSTM>(pl, prm, wpack, stream);
}

int rnn3_forward(int cell, const RnnPlan& pl, RnnParams& prm, const void* wpack, asrb_stream_t stream) {
    // Two chains of 32 rows.  Measured and dropped: four chains of 16; TMA multicast of the copies over
    // CTA pairs and clusters of 8 (the copies are not bound by L2 read bandwidth); h / c / saved activations staged through
    // shared memory and TMA stores, with the direction sum as a TMA reduce-add (more L2 time than the separate sum kernel).
    if (cell == ASRB_RNN_GRU) return rnn3_launch<ASRB_RNN_GRU>(pl, prm, wpack, stream);
    return rnn3_launch<ASRB_RNN_LSTM>(pl, prm, wpack, stream);
}

}  // namespace asrb
