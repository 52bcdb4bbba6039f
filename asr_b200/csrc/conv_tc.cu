// asr_b200 -- 32->32-channel 2-D convolution (the DeepSpeech2 "conv2": nn.Conv2d(32, 32, (21,11), stride (2,1),
// padding (10,5)), asr_deepspeech/modules/deepspeech.py:64) on the tcgen05 tensor cores, as an implicit GEMM with
// no im2col buffer.  Forward, data gradient and weight gradient; time stride must be 1, channels 32/32.
//
// forward / dgrad ("row" kernel).  Activations are NHWC, so one pixel = 32 channels = 128 bytes = exactly one
// row of the 128B-swizzled K-major operand layout.  For an output row (b, d) and a kernel row kh the A operand of
// EVERY kernel column kw is the same strip of input pixels shifted by kw pixels, i.e. by kw smem rows: the strip
// (128 + KW - 1 pixels) is fetched ONCE per kh by TMA (out-of-image pixels zero-filled = the conv padding) and the
// KW taps are issued as tcgen05.mma with the shared-memory descriptor start address advanced by kw*128 B (matrix
// the swizzle is address-based, so no other descriptor field changes).  M = 128 pixels, N = 32 output channels, K = 32 input channels per
// tap, fp32 accumulation in TMEM over all KH*KW taps; 2 pixel tiles per work item share the per-kh weight tiles.
//   dgrad is the same kernel on dy (NHWC) with the kernel-column shift reversed and only the kh of matching stride
//   parity contributing.
// wgrad.  dW[co,ci,kh,kw] = sum_{b,d,t} dy[b,co,d,t] * x[b,ci,d*SH+kh-PH,t+kw-PW]: K runs over time (contiguous in
// NCHW rows), M = (kw, ci) (KW time-shifted TMA boxes of the same input row; TMA needs 16-byte aligned inner coordinates, so
// the boxes come from 4 copies of x delayed by 0..3 samples), N = co; one CTA per (kh, chunk of rows),
// partial sums combined with fp32 atomics.
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace asrb {

constexpr int kCtThreads = 192;
constexpr int kCtC = 32;                 // channels (in and out)
constexpr int kCtStrip = 18432;          // bytes reserved per A strip (144 rows x 128 B, 1024-aligned)
constexpr int kCtWTile = 4096;           // 32 x 32 fp32 weight tile

struct ConvRowParams {
    int B, Hs, Ws, Ho, Wo, KH, KW, SH, PH, PW, mode;   // mode 0 = forward, 1 = data gradient
    const float* bias;
    const int* lengths;
    float* out;                                        // NCHW [B][32][Ho][Wo]
    int num_items, tpairs;
};

// Operand view that starts `shift_rows` pixels (128-byte rows) into a strip.  The 128B swizzle is a function of
// the absolute shared-memory address bits (TMA writes it that way and tcgen05.mma reads it that way -- measured
// on B200: results are exact with the descriptor's "matrix base offset" field left at 0), so advancing the start
// address is all it takes.
__device__ __forceinline__ uint64_t umma_desc_sw128_shift(uint32_t smem_addr, int shift_rows) {
    return umma_desc_sw128(smem_addr + shift_rows * 128);
}

// source row feeding output row `drow` through kernel row kh, or -1
__device__ __forceinline__ int conv_src_row(const ConvRowParams& p, int drow, int kh) {
    if (p.mode == 0) {
        const int hs = drow * p.SH + kh - p.PH;
        return (hs >= 0 && hs < p.Hs) ? hs : -1;
    }
    const int num = drow + p.PH - kh;
    if (num < 0 || num % p.SH != 0) return -1;
    const int hs = num / p.SH;
    return hs < p.Hs ? hs : -1;
}

__global__ void __launch_bounds__(kCtThreads, 1)
conv_row_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmW,
                   const ConvRowParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = 2 * kCtStrip + p.KW * kCtWTile;
    constexpr int kStages = 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    constexpr int kTmemCols = 128;        // 2 accumulator buffers x 2 pixel tiles x 32 channels

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int strip_rows = 128 + p.KW - 1;
    const uint32_t strip_bytes = (uint32_t)strip_rows * 128u;

    auto decode = [&](int item, int& b, int& drow, int& t0, int& n_mt) {
        const int tp = item % p.tpairs;
        const int r = item / p.tpairs;
        drow = r % p.Ho;
        b = r / p.Ho;
        t0 = tp * 256;
        n_mt = (t0 + 128 < p.Wo) ? 2 : 1;
    };
    auto num_groups = [&](int drow) {
        int n = 0;
        for (int kh = 0; kh < p.KH; ++kh) n += conv_src_row(p, drow, kh) >= 0;
        return n;
    };

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int b, drow, t0, n_mt;
            decode(item, b, drow, t0, n_mt);
            // first source column of the strip: forward x col = t + kw - PW ; dgrad dy col = t + PW - kw
            const int col0 = p.mode == 0 ? t0 - p.PW : t0 + p.PW - (p.KW - 1);
            for (int kh = 0; kh < p.KH; ++kh) {
                const int hs = conv_src_row(p, drow, kh);
                if (hs < 0) continue;
                uint8_t* st = smem + stage * stage_bytes;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], n_mt * strip_bytes + p.KW * kCtWTile);
                    for (int mt = 0; mt < n_mt; ++mt)
                        tma_load_4d(st + mt * kCtStrip, &tmS, &full_bar[stage], 0, col0 + mt * 128, hs, b);
                    for (int kw = 0; kw < p.KW; ++kw)
                        tma_load_2d(st + 2 * kCtStrip + kw * kCtWTile, &tmW, &full_bar[stage], 0, (kh * p.KW + kw) * kCtC);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        constexpr uint32_t idesc = umma_idesc(kFmtTF32, 128, kCtC);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int b, drow, t0, n_mt;
            decode(item, b, drow, t0, n_mt);
            const int ngroups = num_groups(drow);
            if (ngroups == 0) continue;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after_sync();
            int gdone = 0;
            for (int kh = 0; kh < p.KH; ++kh) {
                if (conv_src_row(p, drow, kh) < 0) continue;
                uint8_t* st = smem + stage * stage_bytes;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                if (elect_one()) {
                    for (int kw = 0; kw < p.KW; ++kw) {
                        const int shift = p.mode == 0 ? kw : p.KW - 1 - kw;
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(st + 2 * kCtStrip + kw * kCtWTile));
                        for (int mt = 0; mt < n_mt; ++mt) {
                            const uint64_t adesc = umma_desc_sw128_shift(smem_u32(st + mt * kCtStrip), shift);
                            const uint32_t d_tmem = tmem_base + acc * 64 + mt * 32;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (gdone | kw | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (gdone == ngroups - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                ++gdone;
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..5): thread = pixel, 32 channels in registers =====================
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int b, drow, t0, n_mt;
            decode(item, b, drow, t0, n_mt);
            const bool have = num_groups(drow) > 0;
            if (have) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after_sync();
            }
            const int len = (p.mode == 0 && p.lengths) ? p.lengths[b] : p.Wo;
            for (int mt = 0; mt < n_mt; ++mt) {
                float v[32];
                if (have) {
                    tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * 64 + mt * 32, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = 0.f;
                }
                const int t = t0 + mt * 128 + quad * 32 + lane;
                if (t < p.Wo) {
                    float* o = p.out + (((size_t)b * kCtC) * p.Ho + drow) * p.Wo + t;
                    const size_t cstride = (size_t)p.Ho * p.Wo;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        float r = v[c];
                        if (p.mode == 0) r = t < len ? r + (p.bias ? __ldg(p.bias + c) : 0.f) : 0.f;
                        o[c * cstride] = r;
                    }
                }
            }
            if (have) {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// NR output rows per work item
// ------------------------------------------------------------------------------------------------
// conv_row_tc_kernel above is bound by shared-memory operand reads: every MMA re-reads a 16 KB strip tile for N = 32
// output channels (20 KB per 64 cycles of tensor work against 128 B/clk).  A source row feeds up to KH/SH output rows, through
// different kernel rows: rows d0 + j * rowstep (j < NR) see source row hs through kernel rows kh_j = kh_a -/+ j * SH.  Stacking
// their tap matrices along N gives one MMA of N = 32 * NR per (strip, tap, K step): the strip tile is read once for NR output
// rows (NR = 4: 16 KB + 16 KB per 64-cycle MMA -- tensor pipe and SMEM port balanced).  The stacked tap matrices
// ([entry e = kh_a - kh_lo][kw][32 NR rows][32], zero blocks where kh_j falls outside the kernel) come from
// asrb_conv32_pack_rows.  Two rings: strips (2 pixel tiles per stage) and tap tiles (16 KB each).
struct ConvRowsParams {
    int B, Hs, Ws, Ho, Wo, KH, KW, SH, PH, PW, mode;   // mode 0 = forward, 1 = data gradient
    const float* bias;
    const int* lengths;
    float* out;                                        // NCHW [B][32][Ho][Wo]
    int num_items, tpairs, groups, n_e;
};

// output row 0 of group g, and the source row that pack entry e holds for it (or -1)
__device__ __forceinline__ int conv_rows_d0(const ConvRowsParams& p, int g, int NR) {
    if (p.mode == 0) return g * NR;
    return (g / p.SH) * p.SH * NR + g % p.SH;
}
__device__ __forceinline__ int conv_rows_src(const ConvRowsParams& p, int d0, int e, int NR) {
    if (p.mode == 0) {
        const int hs = d0 * p.SH + e - p.PH;                       // kh_a = e
        return (hs >= 0 && hs < p.Hs) ? hs : -1;
    }
    const int kh_a = e - p.SH * (NR - 1);
    const int num = d0 + p.PH - kh_a;
    if (num < 0 || num % p.SH != 0) return -1;
    const int hs = num / p.SH;
    return hs < p.Hs ? hs : -1;
}

template <int NR>
__global__ void __launch_bounds__(kCtThreads, 1)
conv_rows_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmW, const ConvRowsParams p) {
    constexpr int N = 32 * NR;
    constexpr int kWTile = N * 128;                    // one tap: [N rows][32 tf32]
    constexpr int kSStages = 2;                        // strip ring (2 pixel tiles each)
    constexpr int kWStages = (NR == 4) ? 6 : 10;       // tap-tile ring
    constexpr int kTmemCols = 4 * N;                   // 2 accumulator buffers x 2 pixel tiles x N columns
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_strip = smem;                                        // [kSStages][2][kCtStrip]
    uint8_t* s_w = smem + kSStages * 2 * kCtStrip;                  // [kWStages][kWTile]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + kWStages * kWTile);
    uint64_t* sfull = bars;
    uint64_t* sempty = sfull + kSStages;
    uint64_t* wfull = sempty + kSStages;
    uint64_t* wempty = wfull + kWStages;
    uint64_t* tfull_bar = wempty + kWStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < kSStages; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t strip_bytes = (uint32_t)(128 + p.KW - 1) * 128u;

    auto decode = [&](int item, int& b, int& d0, int& t0, int& n_mt) {
        const int tp = item % p.tpairs;
        const int r = item / p.tpairs;
        d0 = conv_rows_d0(p, r % p.groups, NR);
        b = r / p.groups;
        t0 = tp * 256;
        n_mt = (t0 + 128 < p.Wo) ? 2 : 1;
    };

    if (warp == 0) {
        // ===================== TMA producer: one elected thread =====================
        if (elect_one()) {
            int ss = 0, ws = 0;
            uint32_t sph = 0, wph = 0;
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int b, d0, t0, n_mt;
                decode(item, b, d0, t0, n_mt);
                if (d0 >= p.Ho) continue;
                // first source column of the strip: forward x col = t + kw - PW ; dgrad dy col = t + PW - kw
                const int col0 = p.mode == 0 ? t0 - p.PW : t0 + p.PW - (p.KW - 1);
                for (int e = 0; e < p.n_e; ++e) {
                    const int hs = conv_rows_src(p, d0, e, NR);
                    if (hs < 0) continue;
                    mbar_wait(&sempty[ss], sph ^ 1);
                    mbar_arrive_expect_tx(&sfull[ss], n_mt * strip_bytes);
                    for (int mt = 0; mt < n_mt; ++mt)
                        tma_load_4d(s_strip + (ss * 2 + mt) * kCtStrip, &tmS, &sfull[ss], 0, col0 + mt * 128, hs, b);
                    if (++ss == kSStages) { ss = 0; sph ^= 1; }
                    for (int kw = 0; kw < p.KW; ++kw) {
                        mbar_wait(&wempty[ws], wph ^ 1);
                        mbar_arrive_expect_tx(&wfull[ws], (uint32_t)kWTile);
                        tma_load_2d(s_w + ws * kWTile, &tmW, &wfull[ws], 0, (e * p.KW + kw) * N);
                        if (++ws == kWStages) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: one elected thread =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kFmtTF32, 128, N);
            int ss = 0, ws = 0, acc = 0;
            uint32_t sph = 0, wph = 0, acc_phase = 0;
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int b, d0, t0, n_mt;
                decode(item, b, d0, t0, n_mt);
                if (d0 >= p.Ho) continue;
                int nstrips = 0;
                for (int e = 0; e < p.n_e; ++e) nstrips += conv_rows_src(p, d0, e, NR) >= 0;
                if (nstrips == 0) continue;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * 2 * N;
                int done = 0;
                for (int e = 0; e < p.n_e; ++e) {
                    if (conv_rows_src(p, d0, e, NR) < 0) continue;
                    mbar_wait(&sfull[ss], sph);
                    const uint32_t a0 = smem_u32(s_strip + (ss * 2) * kCtStrip);
                    for (int kw = 0; kw < p.KW; ++kw) {
                        const int shift = p.mode == 0 ? kw : p.KW - 1 - kw;
                        mbar_wait(&wfull[ws], wph);
                        tc_fence_after_sync();
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(s_w + ws * kWTile));
                        for (int mt = 0; mt < n_mt; ++mt) {
                            const uint64_t adesc = umma_desc_sw128(a0 + mt * kCtStrip + shift * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_tf32(d_tmem + mt * N, adesc + 2 * k, bdesc + 2 * k, idesc, (done | kw | k) != 0);
                        }
                        umma_commit(&wempty[ws]);
                        if (++ws == kWStages) { ws = 0; wph ^= 1; }
                    }
                    umma_commit(&sempty[ss]);
                    if (++ss == kSStages) { ss = 0; sph ^= 1; }
                    ++done;
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..5): thread = pixel, 32 channels of one output row at a time =====================
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int rowstep = p.mode == 0 ? 1 : p.SH;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int b, d0, t0, n_mt;
            decode(item, b, d0, t0, n_mt);
            if (d0 >= p.Ho) continue;
            int nstrips = 0;
            for (int e = 0; e < p.n_e; ++e) nstrips += conv_rows_src(p, d0, e, NR) >= 0;
            const bool have = nstrips > 0;
            if (have) {
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after_sync();
            }
            const int len = (p.mode == 0 && p.lengths) ? p.lengths[b] : p.Wo;
            for (int mt = 0; mt < n_mt; ++mt) {
                const int t = t0 + mt * 128 + quad * 32 + lane;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const int drow = d0 + j * rowstep;
                    if (drow >= p.Ho) break;                       // (warp-uniform)
                    float v[32];
                    if (have) {
                        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * 2 * N + mt * N + j * 32, v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = 0.f;
                    }
                    if (t < p.Wo) {
                        float* o = p.out + (((size_t)b * kCtC) * p.Ho + drow) * p.Wo + t;
                        const size_t cstride = (size_t)p.Ho * p.Wo;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            float r = v[c];
                            if (p.mode == 0) r = t < len ? r + (p.bias ? __ldg(p.bias + c) : 0.f) : 0.f;
                            o[c * cstride] = r;
                        }
                    }
                }
            }
            if (have) {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// stacked tap matrices: out[e][kw][j * 32 + n][k] = pack[(kh_j * KW + kw)][n][k] with kh_j = e - j * SH (forward) or
// e - SH (NR - 1) + j * SH (data gradient), zero where kh_j is outside [0, KH)
__global__ void conv_pack_rows_kernel(const float* __restrict__ pack, float* __restrict__ out, int KH, int KW, int SH, int NR, int mode) {
    const int n_e = KH + SH * (NR - 1);
    const size_t total = (size_t)n_e * KW * NR * 1024;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % 32), n = (int)((i / 32) % 32), j = (int)((i / 1024) % NR);
        const int kw = (int)((i / ((size_t)1024 * NR)) % KW), e = (int)(i / ((size_t)1024 * NR * KW));
        const int kh = mode == 0 ? e - j * SH : e - SH * (NR - 1) + j * SH;
        out[i] = (kh >= 0 && kh < KH) ? pack[((size_t)(kh * KW + kw) * 32 + n) * 32 + k] : 0.f;
    }
}

template <int NR>
static int conv_rows_launch(const float* src_nhwc, const float* wpack_rows, ConvRowsParams& p, asrb_stream_t stream) {
    constexpr int N = 32 * NR;
    constexpr int kWStages = (NR == 4) ? 6 : 10;
    CUtensorMap tmS, tmW;
    p.n_e = p.KH + p.SH * (NR - 1);
    {
        uint64_t d[4] = {32, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B};
        uint64_t s[3] = {128, (uint64_t)p.Ws * 128, (uint64_t)p.Hs * p.Ws * 128};
        uint32_t bx[4] = {32, (uint32_t)(128 + p.KW - 1), 1, 1};
        int rc = make_tmap_f32(&tmS, src_nhwc, 4, d, s, bx);
        if (rc) return rc;
    }
    {
        uint64_t d[2] = {32, (uint64_t)p.n_e * p.KW * N}, s[1] = {128};
        uint32_t bx[2] = {32, (uint32_t)N};
        int rc = make_tmap_f32(&tmW, wpack_rows, 2, d, s, bx);
        if (rc) return rc;
    }
    p.tpairs = ceil_div(p.Wo, 256);
    p.groups = p.mode == 0 ? ceil_div(p.Ho, NR) : ceil_div(p.Ho, p.SH * NR) * p.SH;
    p.num_items = p.B * p.groups * p.tpairs;
    const size_t smem = (size_t)2 * 2 * kCtStrip + (size_t)kWStages * N * 128 + 1024 + 512;
    auto kern = conv_rows_tc_kernel<NR>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_items < kNumSMs ? p.num_items : kNumSMs;
    kern<<<grid, kCtThreads, smem, stream>>>(tmS, tmW, p);
    ASRB_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
struct ConvWgradParams {
    int B, Hin, Win, Hout, Wout, KH, KW, SH, PH, PW;
    int rows_per_chunk, m_tiles;
    float* dw;                                          // [32][32][KH][KW]
};

/* BF16: operands are bf16 copies (gradient-only product, like the recurrent stack's backward GEMMs): a 128-byte
 * swizzle row then holds 64 time samples instead of 32, which halves the SMEM bytes (TMA writes + tensor-core operand
 * reads, the resource that bounds this kernel: every x sample is loaded KW times) per sample. */
template <bool BF16>
__global__ void __launch_bounds__(kCtThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                     const ConvWgradParams p) {
    constexpr int KT = BF16 ? 64 : 32;                     // time samples per stage (one 128-byte row)
    constexpr int NQ = BF16 ? 8 : 4;                       // delayed copies: box starts must be 16-byte aligned
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = p.m_tiles * 16384;                 // m_tiles x [128 rows = 4 kernel columns x 32 ci][32 t]
    const int stage_bytes = a_bytes + kCtWTile;            // + dy tile [32 co][32 t]
    const int stages = (200 * 1024) / stage_bytes < 6 ? (200 * 1024) / stage_bytes : 6;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* tfull_bar = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    constexpr int kTmemCols = 128;                         // up to 4 M-tiles x 32 columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kh = blockIdx.x % p.KH, chunk = blockIdx.x / p.KH;
    const int nrows = p.B * p.Hout;
    const int r0 = chunk * p.rows_per_chunk;
    const int r1 = min(nrows, r0 + p.rows_per_chunk);
    const int n_kb = ceil_div(p.Wout, KT);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDy);
        for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // number of (row, k-block) steps this CTA performs (rows whose input row falls outside the image are skipped)
    int total_rows = 0;
    for (int r = r0; r < r1; ++r) {
        const int hi = (r % p.Hout) * p.SH + kh - p.PH;
        total_rows += (hi >= 0 && hi < p.Hin);
    }

    if (warp == 0) {
        // TMA producer: warp-uniform loop, one elected lane issues
        int stage = 0;
        uint32_t phase = 0;
        for (int r = r0; r < r1; ++r) {
            const int b = r / p.Hout, ho = r % p.Hout;
            const int hi = ho * p.SH + kh - p.PH;
            if (hi < 0 || hi >= p.Hin) continue;
            for (int kb = 0; kb < n_kb; ++kb) {
                uint8_t* st = smem + stage * stage_bytes;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(p.m_tiles * 4 + 1) * kCtWTile);
                    for (int j = 0; j < p.m_tiles * 4; ++j) { // kernel column j (columns >= KW are never read back)
                        const int off = j - p.PW;             // time shift; copy q holds x delayed by q samples
                        const int q = (((-off) % NQ) + NQ) % NQ; // so that the box start kb*KT + off + q is 16-byte aligned
                        tma_load_5d(st + j * kCtWTile, &tmX, &full_bar[stage], kb * KT + off + q, hi, 0, b, q);
                    }
                    tma_load_4d(st + a_bytes, &tmDy, &full_bar[stage], kb * KT, ho, 0, b);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform loop, one elected lane issues
        if (total_rows > 0) {
            constexpr uint32_t idesc = umma_idesc(BF16 ? kFmtBF16 : kFmtTF32, 128, kCtC);
            int stage = 0;
            uint32_t phase = 0;
            const int steps = total_rows * n_kb;
            for (int s = 0; s < steps; ++s) {
                uint8_t* st = smem + stage * stage_bytes;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                if (elect_one()) {
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(st + a_bytes));
                    for (int mt = 0; mt < p.m_tiles; ++mt) {
                        const uint64_t adesc = umma_desc_sw128(smem_u32(st + mt * 16384));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if constexpr (BF16) umma_f16(tmem_base + mt * 32, adesc + 2 * k, bdesc + 2 * k, idesc, (s | k) != 0);
                            else umma_tf32(tmem_base + mt * 32, adesc + 2 * k, bdesc + 2 * k, idesc, (s | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (s == steps - 1) umma_commit(tfull_bar);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        if (total_rows > 0) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after_sync();
            for (int mt = 0; mt < p.m_tiles; ++mt) {
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + mt * 32, v);
                tmem_ld_wait();
                const int kw = mt * 4 + quad, ci = lane;     // accumulator row = (kernel column, input channel)
                if (kw < p.KW) {
#pragma unroll
                    for (int co = 0; co < 32; ++co)
                        atomicAdd(p.dw + (((size_t)co * kCtC + ci) * p.KH + kh) * p.KW + kw, v[co]);
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
}

/* Weight gradient, one CTA per (kernel-row class, M tile, chunk of source rows).  conv_wgrad_tc_kernel above fixes kh per
 * CTA, so every x row is fetched KH times (and each of its KW shifted copies once): 23 GB of L2 -> SM traffic per launch at
 * configs[1], the L2 bandwidth cap.  Here a CTA walks the SOURCE rows: x row hi meets the gradient rows y = (hi + PH - kh) / SH
 * of every kh of its residue class c = (hi + PH) mod SH -- kh = c + SH j, up to NKH = ceil(KH / SH) of them -- and those rows
 * are consecutive in y: ONE TMA box [NKH rows of y][32 co][64 t] is the stacked B operand (N = 32 NKH <= 352, zero-filled
 * outside the image), the x taps of an M tile (4 kernel columns x 32 ci) are read once per stage for all of them, and the
 * accumulators of all NKH kernel rows stay in tensor memory (352 columns).  bf16 operands only. */
struct ConvWgradClsParams {
    int B, Hin, Win, Hout, Wout, KH, KW, SH, PH, PW;
    int m_tiles, nkh, chunks;
    float* dw;                                          // [32][32][KH][KW]
};

__global__ void __launch_bounds__(kCtThreads, 1)
conv_wgrad_cls_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                      const ConvWgradClsParams p) {
    constexpr int KT = 64, NQ = 8;
    constexpr int kStages = 3;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int ntot = 32 * p.nkh;                           // stacked N
    const int b_bytes = ntot * 128;
    const int stage_bytes = 16384 + ((b_bytes + 1023) & ~1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    constexpr int kTmemCols = 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = (class * m_tiles + mt) * chunks + chunk
    const int chunk = blockIdx.x % p.chunks;
    const int mt = (blockIdx.x / p.chunks) % p.m_tiles;
    const int c = blockIdx.x / (p.chunks * p.m_tiles);
    const int hi0 = (((c - p.PH) % p.SH) + p.SH) % p.SH;   // first source row of the class
    const int nhi = hi0 < p.Hin ? (p.Hin - hi0 + p.SH - 1) / p.SH : 0;
    const int nrows = p.B * nhi;
    const int per = (nrows + p.chunks - 1) / p.chunks;
    const int r0 = chunk * per, r1 = min(nrows, r0 + per);
    const int n_kb = ceil_div(p.Wout, KT);
    const int steps = max(0, r1 - r0) * n_kb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDy);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int r = r0; r < r1; ++r) {
                const int b = r / nhi, hi = hi0 + (r % nhi) * p.SH;
                const int ybase = (hi + p.PH - c) / p.SH;          // the gradient row that kh = c pairs with x row hi
                for (int kb = 0; kb < n_kb; ++kb) {
                    uint8_t* st = smem + stage * stage_bytes;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(16384 + b_bytes));
                    for (int jw = 0; jw < 4; ++jw) {              // kernel column 4 mt + jw (columns >= KW are never read back)
                        const int off = 4 * mt + jw - p.PW;       // time shift; copy q holds x delayed by q samples
                        const int q = (((-off) % NQ) + NQ) % NQ;  // so that the box start kb*KT + off + q is 16-byte aligned
                        tma_load_5d(st + jw * kCtWTile, &tmX, &full_bar[stage], kb * KT + off + q, hi, 0, b, q);
                    }
                    // rows y = ybase - (nkh - 1) .. ybase, 32 co each: column block jj <-> kh = c + SH (nkh - 1 - jj)
                    tma_load_4d(st + 16384, &tmDy, &full_bar[stage], kb * KT, 0, ybase - (p.nkh - 1), b);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one() && steps > 0) {
            const int n1 = ntot > 256 ? 256 : ntot, n2 = ntot - n1;
            const uint32_t idesc1 = umma_idesc(kFmtBF16, 128, n1), idesc2 = umma_idesc(kFmtBF16, 128, n2 > 0 ? n2 : 16);
            int stage = 0;
            uint32_t phase = 0;
            for (int s = 0; s < steps; ++s) {
                uint8_t* st = smem + stage * stage_bytes;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                const uint64_t adesc = umma_desc_sw128(smem_u32(st));
                const uint64_t bdesc = umma_desc_sw128(smem_u32(st + 16384));
                const uint64_t bdesc2 = umma_desc_sw128(smem_u32(st + 16384 + 256 * 128));     // rows 256.. of the stacked tile
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc1, (s | k) != 0);
                    if (n2 > 0) umma_f16(tmem_base + 256, adesc + 2 * k, bdesc2 + 2 * k, idesc2, (s | k) != 0);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull_bar);
        }
        __syncwarp();
    } else {
        const int quad = warp & 3;
        if (steps > 0) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after_sync();
            const int kw = mt * 4 + quad, ci = lane;           // accumulator row = (kernel column, input channel)
            for (int jj = 0; jj < p.nkh; ++jj) {
                const int kh = c + p.SH * (p.nkh - 1 - jj);
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + jj * 32, v);
                tmem_ld_wait();
                if (kw < p.KW && kh < p.KH) {
#pragma unroll
                    for (int co = 0; co < 32; ++co)
                        atomicAdd(p.dw + (((size_t)co * kCtC + ci) * p.KH + kh) * p.KW + kw, v[co]);
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
}

// bf16 forms: xs[r][row][w] = x[row][w - r], r = 0..7, row stride ldo >= W + 7 (multiple of 8)
__global__ void conv_shift_copies_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xs, long long rows,
                                              int W, int ldo) {
    // one thread = 8 consecutive output samples of a row, for all 8 copies: 15 loads, eight 16-byte stores
    const int gpr = ldo / 8;                                  // groups per row (ldo is a multiple of 8)
    const long long per = rows * ldo, ngroups = rows * gpr;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
        const long long row = g / gpr;
        const int w0 = (int)(g - row * gpr) * 8;
        const float* xr = x + row * W;
        float v[15];                                          // x[w0 - 7 .. w0 + 7]
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            const int w = w0 - 7 + i;
            v[i] = (w >= 0 && w < W) ? __ldg(xr + w) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {                         // copy r, sample w0 + e = x[w0 + e - r] = v[7 + e - r]
            uint4 o;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[7 - r], v[8 - r]), p1 = __floats2bfloat162_rn(v[9 - r], v[10 - r]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[11 - r], v[12 - r]), p3 = __floats2bfloat162_rn(v[13 - r], v[14 - r]);
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(xs + r * per + row * ldo + w0) = o;
        }
    }
}

// out[row][0..ldo) = bf16(in[row][0..W)), zero padded
__global__ void conv_rows_bf16_kernel(const float* __restrict__ in, int ldi, __nv_bfloat16* __restrict__ out, long long rows,
                                      int W, int ldo) {
    const long long per = rows * ldo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / ldo;
        const int w = (int)(i - row * ldo);
        out[i] = __float2bfloat16(w < W ? in[row * ldi + w] : 0.f);
    }
}

// xs[r][row][w] = x[row][w - r] (zero outside the row), r = 0..3, row stride ldo >= W + 3 (multiple of 4)
__global__ void conv_shift_copies_kernel(const float* __restrict__ x, float* __restrict__ xs, long long rows, int W, int ldo) {
    const long long per = rows * ldo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / ldo;
        const int w = (int)(i % ldo);
        const float* xr = x + row * W;
#pragma unroll
        for (int r = 0; r < 4; ++r) xs[r * per + i] = (w - r >= 0 && w - r < W) ? xr[w - r] : 0.f;
    }
}

// w [32][32][KH][KW] -> fwd pack [kh][kw][co][ci], dgrad pack [kh][kw][ci][co]
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ pf, float* __restrict__ pd,
                                         int KH, int KW) {
    const int total = KH * KW * 32 * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kw = i % KW, kh = (i / KW) % KH, ci = (i / (KW * KH)) % 32, co = i / (KW * KH * 32);
        const float v = w[i];
        const size_t tap = (size_t)(kh * KW + kw) * 1024;
        if (pf) pf[tap + co * 32 + ci] = v;
        if (pd) pd[tap + ci * 32 + co] = v;
    }
}

static int conv_row_launch(const float* src_nhwc, const float* wpack, ConvRowParams& p, asrb_stream_t stream) {
    CUtensorMap tmS, tmW;
    {
        uint64_t d[4] = {32, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B};
        uint64_t s[3] = {128, (uint64_t)p.Ws * 128, (uint64_t)p.Hs * p.Ws * 128};
        uint32_t bx[4] = {32, (uint32_t)(128 + p.KW - 1), 1, 1};
        int rc = make_tmap_f32(&tmS, src_nhwc, 4, d, s, bx);
        if (rc) return rc;
    }
    {
        uint64_t d[2] = {32, (uint64_t)p.KH * p.KW * 32}, s[1] = {128};
        uint32_t bx[2] = {32, 32};
        int rc = make_tmap_f32(&tmW, wpack, 2, d, s, bx);
        if (rc) return rc;
    }
    p.tpairs = ceil_div(p.Wo, 256);
    p.num_items = p.B * p.Ho * p.tpairs;
    const size_t smem = 2 * (size_t)(2 * kCtStrip + p.KW * kCtWTile) + 1024 + 256;
    ASRB_CUDA_OK(cudaFuncSetAttribute(conv_row_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_items < kNumSMs ? p.num_items : kNumSMs;
    conv_row_tc_kernel<<<grid, kCtThreads, smem, stream>>>(tmS, tmW, p);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // namespace asrb

using namespace asrb;

extern "C" {

int asrb_conv32_supported(int Cin, int Cout, int KH, int KW, int SH, int SW, int PH, int PW) {
    return Cin == 32 && Cout == 32 && SW == 1 && KW >= 1 && KW <= 16 && KH >= 1 && SH >= 1 && PH >= 0 && PW >= 0 &&
           (2 * (2 * kCtStrip + KW * kCtWTile) + 2048 <= 227 * 1024);
}

int asrb_conv32_pack_weights(const float* w, float* pack_fwd, float* pack_dgrad, int KH, int KW, asrb_stream_t stream) {
    ASRB_REQUIRE(w && (pack_fwd || pack_dgrad) && KH > 0 && KW > 0, ASRB_ERR_BAD_ARG);
    conv_pack_weights_kernel<<<kNumSMs, 256, 0, stream>>>(w, pack_fwd, pack_dgrad, KH, KW);
    ASRB_LAUNCH_OK();
    return 0;
}

/* y[B,32,Hout,Wout] (NCHW) = mask(conv(x) + bias), x given as NHWC [B,Hin,Win,32], weights from asrb_conv32_pack_weights */
int asrb_conv32_fwd(const float* x_nhwc, const float* pack_fwd, const float* bias, const int32_t* lengths, float* y,
                    int B, int Hin, int Win, int Hout, int Wout, int KH, int KW, int SH, int PH, int PW,
                    asrb_stream_t stream) {
    ASRB_REQUIRE(x_nhwc && pack_fwd && y && B > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv32_supported(32, 32, KH, KW, SH, 1, PH, PW), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (Hin + 2 * PH - KH) / SH + 1 && Wout == Win + 2 * PW - KW + 1, ASRB_ERR_BAD_ARG);
    ConvRowParams p = {B, Hin, Win, Hout, Wout, KH, KW, SH, PH, PW, 0, bias, lengths, y, 0, 0};
    return conv_row_launch(x_nhwc, pack_fwd, p, stream);
}

/* dx[B,32,Hin,Win] (NCHW) from dy given as NHWC [B,Hout,Wout,32] (already masked) */
int asrb_conv32_bwd_data(const float* dy_nhwc, const float* pack_dgrad, float* dx, int B, int Hin, int Win, int Hout,
                         int Wout, int KH, int KW, int SH, int PH, int PW, asrb_stream_t stream) {
    ASRB_REQUIRE(dy_nhwc && pack_dgrad && dx && B > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv32_supported(32, 32, KH, KW, SH, 1, PH, PW), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (Hin + 2 * PH - KH) / SH + 1 && Wout == Win + 2 * PW - KW + 1, ASRB_ERR_BAD_ARG);
    ConvRowParams p = {B, Hout, Wout, Hin, Win, KH, KW, SH, PH, PW, 1, nullptr, nullptr, dx, 0, 0};
    return conv_row_launch(dy_nhwc, pack_dgrad, p, stream);
}

/* Row-grouped variants (NR = 2 or 4 output rows per work item share every source strip: see conv_rows_tc_kernel).
 * pack_rows: [KH + SH (NR-1)][KW][32 NR][32] floats, built by asrb_conv32_pack_rows from the plain pack of the same mode. */
int asrb_conv32_pack_rows(const float* pack, float* pack_rows, int KH, int KW, int SH, int NR, int mode, asrb_stream_t stream) {
    ASRB_REQUIRE(pack && pack_rows && KH > 0 && KW > 0 && SH > 0 && (NR == 2 || NR == 4) && (mode == 0 || mode == 1), ASRB_ERR_BAD_ARG);
    conv_pack_rows_kernel<<<kNumSMs, 256, 0, stream>>>(pack, pack_rows, KH, KW, SH, NR, mode);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_conv32_fwd_rows(const float* x_nhwc, const float* pack_rows, const float* bias, const int32_t* lengths, float* y,
                         int B, int Hin, int Win, int Hout, int Wout, int KH, int KW, int SH, int PH, int PW, int NR,
                         asrb_stream_t stream) {
    ASRB_REQUIRE(x_nhwc && pack_rows && y && B > 0 && (NR == 2 || NR == 4), ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv32_supported(32, 32, KH, KW, SH, 1, PH, PW), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (Hin + 2 * PH - KH) / SH + 1 && Wout == Win + 2 * PW - KW + 1, ASRB_ERR_BAD_ARG);
    ConvRowsParams p = {B, Hin, Win, Hout, Wout, KH, KW, SH, PH, PW, 0, bias, lengths, y, 0, 0, 0, 0};
    return NR == 4 ? conv_rows_launch<4>(x_nhwc, pack_rows, p, stream) : conv_rows_launch<2>(x_nhwc, pack_rows, p, stream);
}

int asrb_conv32_bwd_data_rows(const float* dy_nhwc, const float* pack_rows, float* dx, int B, int Hin, int Win, int Hout,
                              int Wout, int KH, int KW, int SH, int PH, int PW, int NR, asrb_stream_t stream) {
    ASRB_REQUIRE(dy_nhwc && pack_rows && dx && B > 0 && (NR == 2 || NR == 4), ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv32_supported(32, 32, KH, KW, SH, 1, PH, PW), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (Hin + 2 * PH - KH) / SH + 1 && Wout == Win + 2 * PW - KW + 1, ASRB_ERR_BAD_ARG);
    ConvRowsParams p = {B, Hout, Wout, Hin, Win, KH, KW, SH, PH, PW, 1, nullptr, nullptr, dx, 0, 0, 0, 0};
    return NR == 4 ? conv_rows_launch<4>(dy_nhwc, pack_rows, p, stream) : conv_rows_launch<2>(dy_nhwc, pack_rows, p, stream);
}

static int g_conv_wgrad_bf16 = 1;
static int g_conv_wgrad_cls = 1;      // bf16 weight gradient: 1 = one CTA per (kernel-row class, M tile, chunk), 0 = per kernel row
/* debug/tuning: the work split of the bf16 weight gradient (see conv_wgrad_cls_kernel); v < 0 queries */
int asrb_debug_conv_wgrad_cls(int v) {
    const int old = g_conv_wgrad_cls;
    if (v >= 0) g_conv_wgrad_cls = v;
    return old;
}
/* debug/tuning: 1 (default) bf16 operand copies for the 32->32 weight gradient, 0 TF32 operands */
int asrb_debug_conv_wgrad_bf16(int v) {
    const int old = g_conv_wgrad_bf16;
    if (v >= 0) g_conv_wgrad_bf16 = v ? 1 : 0;
    return old;
}

size_t asrb_conv32_bwd_weight_workspace_bytes(int B, int Hin, int Win, int Hout, int Wout) {
    const size_t f32 = (size_t)4 * B * 32 * Hin * round_up(Win + 3, 4) * sizeof(float);
    const size_t b16 = ((size_t)8 * B * 32 * Hin * round_up(Win + 7, 8) + (size_t)B * 32 * Hout * round_up(Wout, 8)) * 2;
    return f32 > b16 ? f32 : b16;
}

/* dw[32,32,KH,KW] from x (NCHW, dense) and dy (NCHW, already masked, row stride lddy >= Wout; the TF32 form needs
 * lddy % 4 == 0 for TMA's 16-byte strides: pad odd widths with asrb_copy_rows_padded; the bf16 form converts dy into
 * the workspace and takes any lddy).  ws: asrb_conv32_bwd_weight_workspace_bytes. */
int asrb_conv32_bwd_weight(const float* x, const float* dy, int lddy, float* dw, float* ws, size_t ws_bytes, int B,
                           int Hin, int Win, int Hout, int Wout, int KH, int KW, int SH, int PH, int PW,
                           asrb_stream_t stream) {
    ASRB_REQUIRE(x && dy && dw && ws && B > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv32_supported(32, 32, KH, KW, SH, 1, PH, PW), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (Hin + 2 * PH - KH) / SH + 1 && Wout == Win + 2 * PW - KW + 1, ASRB_ERR_BAD_ARG);
    const bool bf16 = g_conv_wgrad_bf16 != 0;
    ASRB_REQUIRE(lddy >= Wout && (bf16 || lddy % 4 == 0), ASRB_ERR_ALIGNMENT);
    ASRB_REQUIRE(ws_bytes >= asrb_conv32_bwd_weight_workspace_bytes(B, Hin, Win, Hout, Wout), ASRB_ERR_WORKSPACE);
    const int nq = bf16 ? 8 : 4, es = bf16 ? 2 : 4;
    const int ldx = round_up(Win + nq - 1, nq);
    const int ldy = bf16 ? round_up(Wout, 8) : lddy;
    const long long xrows = (long long)B * 32 * Hin, yrows = (long long)B * 32 * Hout;
    const void* dy_src = dy;
    {
        const long long n = bf16 ? xrows * (ldx / 8) : xrows * ldx;      // (bf16: one thread per 8 samples)
        const int g = (int)((n + 255) / 256 < kNumSMs * 8 ? (n + 255) / 256 : kNumSMs * 8);
        if (bf16) {
            __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ws);
            __nv_bfloat16* dyb = xs + (size_t)8 * xrows * ldx;
            conv_shift_copies_bf16_kernel<<<g, 256, 0, stream>>>(x, xs, xrows, Win, ldx);
            ASRB_LAUNCH_OK();
            const long long m = yrows * ldy;
            const int g2 = (int)((m + 255) / 256 < kNumSMs * 8 ? (m + 255) / 256 : kNumSMs * 8);
            conv_rows_bf16_kernel<<<g2, 256, 0, stream>>>(dy, lddy, dyb, yrows, Wout, ldy);
            ASRB_LAUNCH_OK();
            dy_src = dyb;
        } else {
            conv_shift_copies_kernel<<<g, 256, 0, stream>>>(x, ws, xrows, Win, ldx);
            ASRB_LAUNCH_OK();
        }
    }
    CUtensorMap tmX, tmDy;
    {
        uint64_t d[5] = {(uint64_t)Win + nq - 1, (uint64_t)Hin, 32, (uint64_t)B, (uint64_t)nq};
        uint64_t s[4] = {(uint64_t)ldx * es, (uint64_t)Hin * ldx * es, (uint64_t)32 * Hin * ldx * es, (uint64_t)xrows * ldx * es};
        uint32_t bx[5] = {bf16 ? 64u : 32u, 1, 32, 1, 1};
        int rc = bf16 ? make_tmap_bf16(&tmX, ws, 5, d, s, bx) : make_tmap_f32(&tmX, ws, 5, d, s, bx);
        if (rc) return rc;
    }
    {
        uint64_t d[4] = {(uint64_t)Wout, (uint64_t)Hout, 32, (uint64_t)B};
        uint64_t s[3] = {(uint64_t)ldy * es, (uint64_t)Hout * ldy * es, (uint64_t)32 * Hout * ldy * es};
        uint32_t bx[4] = {bf16 ? 64u : 32u, 1, 32, 1};
        int rc = bf16 ? make_tmap_bf16(&tmDy, dy_src, 4, d, s, bx) : make_tmap_f32(&tmDy, dy_src, 4, d, s, bx);
        if (rc) return rc;
    }
    const int nkh = ceil_div(KH, SH);
    if (bf16 && g_conv_wgrad_cls && nkh * 32 <= 352 && nkh <= 256 && 3 * (16384 + round_up(nkh * 32 * 128, 1024)) + 2048 <= 227 * 1024) {
        // one CTA per (kernel-row class, M tile, chunk of source rows): conv_wgrad_cls_kernel
        CUtensorMap tmDy2;
        {   // dy as (t, co, ho, b): a box of nkh consecutive gradient rows is the stacked B operand [ho][co][64 t]
            uint64_t d[4] = {(uint64_t)Wout, 32, (uint64_t)Hout, (uint64_t)B};
            uint64_t s[3] = {(uint64_t)Hout * ldy * es, (uint64_t)ldy * es, (uint64_t)32 * Hout * ldy * es};
            uint32_t bx[4] = {64u, 32, (uint32_t)nkh, 1};
            int rc = make_tmap_bf16(&tmDy2, dy_src, 4, d, s, bx);
            if (rc) return rc;
        }
        ConvWgradClsParams pc = {B, Hin, Win, Hout, Wout, KH, KW, SH, PH, PW, ceil_div(KW, 4), nkh, 1, dw};
        const int combos = SH * pc.m_tiles;
        int chunks = kNumSMs / combos;
        if (chunks < 1) chunks = 1;
        const int max_rows = B * ceil_div(Hin, SH);
        if (chunks > max_rows) chunks = max_rows;
        pc.chunks = chunks;
        ASRB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)32 * 32 * KH * KW * sizeof(float), stream));
        const size_t smem = (size_t)3 * (16384 + round_up(nkh * 32 * 128, 1024)) + 1024 + 256;
        ASRB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_cls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_cls_kernel<<<combos * chunks, kCtThreads, smem, stream>>>(tmX, tmDy2, pc);
        ASRB_LAUNCH_OK();
        return 0;
    }
    ConvWgradParams p = {B, Hin, Win, Hout, Wout, KH, KW, SH, PH, PW, 0, ceil_div(KW, 4), dw};
    int chunks = kNumSMs / KH;
    if (chunks < 1) chunks = 1;
    const int nrows = B * Hout;
    if (chunks > nrows) chunks = nrows;
    p.rows_per_chunk = ceil_div(nrows, chunks);
    chunks = ceil_div(nrows, p.rows_per_chunk);
    ASRB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)32 * 32 * KH * KW * sizeof(float), stream));
    const size_t smem = 200 * 1024 + 1024 + 256;
    if (bf16) {
        ASRB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_tc_kernel<true><<<KH * chunks, kCtThreads, smem, stream>>>(tmX, tmDy, p);
    } else {
        ASRB_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_tc_kernel<false><<<KH * chunks, kCtThreads, smem, stream>>>(tmX, tmDy, p);
    }
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
