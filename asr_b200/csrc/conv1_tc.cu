// asr_b200 -- the DeepSpeech2 "conv1" (nn.Conv2d(1, 32, (41,11), stride (2,2), padding (20,5)),
// asr_deepspeech/modules/deepspeech.py:61) on the tcgen05 tensor cores: forward and weight gradient
// (the spectrogram needs no gradient).  Geometry handled: Cin = 1, Cout = 32, KW = 11, SW = 2, PW = 5, SH = 2,
// PH even, KH <= 48.
//
// Polyphase view of the stride-2 time axis: x[2 t + kw - 5] is phase (kw-5)&1 of x at position t + s,
// s = floor((kw-5)/2) in {-3..2}.  So conv1 is a stride-1, 6-tap convolution over "channels" = (input row, phase):
//   y[b,co,d,t] = sum_{s} sum_{kh,ph} W'[s][co][(kh,ph)] * XT[b][t+s][(2d-20+kh, ph)]
// forward : XT is a time-major copy of the spectrogram, [B][T'][324] (161 rows x 2 phases, padded), so the 82 channels
//           an output row needs are CONTIGUOUS and start at a multiple of 4 floats: the same shifted-descriptor
//           implicit GEMM as conv2 (one TMA strip of 128+5 time positions per 32-channel block, 6 taps = 6 descriptor
//           shifts), with all weights (72 KB) resident in shared memory.
// wgrad   : K = time (contiguous in the phase-split rows XP[q][b][ph][row][t], 4 copies delayed by q samples for
//           16-byte aligned TMA coordinates), M = (tap, kh, phase) = 492 rows -> 4 accumulator tiles, N = co; every
//           (b, d) output row accumulates into the SAME tiles, one CTA per chunk of rows, fp32 atomics at the end.
#include "ptx.cuh"

namespace asrb {

constexpr int kC1Threads = 192;
constexpr int kC1Taps = 6;                 // s = -3..2
constexpr int kC1Kb = 3;                   // 32-channel blocks per output row (covers 2*KH <= 96 channels)
constexpr int kC1Ld = 324;                 // channels per time step in XT (161 rows x 2 phases, padded to 16 bytes)
constexpr int kC1Strip = 17408;            // 136 rows x 128 B reserved per strip (133 used)
constexpr int kC1WBytes = kC1Taps * kC1Kb * 4096;

struct Conv1Params {
    int B, F, T, Tp, Hout, Wout, KH, PH;   // Tp = phases' length = ceil(T/2); Hout/Wout = output rows/cols
    const float* bias;
    const int* lengths;
    float* out;                            // [B][32][Hout][Wout]
    float* dw;                             // [32][1][KH][11]  (wgrad)
    int num_items, ttiles, rows_per_chunk;
};

// XT[b][tau][row*2+ph] = x[b][row][2 tau + ph]   (zero beyond T / beyond 161 rows)
__global__ void conv1_make_xt_kernel(const float* __restrict__ x, float* __restrict__ xt, int B, int F, int T, int Tp) {
    __shared__ float tile[32][65];
    const int b = blockIdx.z, r0 = blockIdx.y * 32, tau0 = blockIdx.x * 32;
    // read 32 rows x 64 consecutive samples (coalesced along time)
    for (int i = threadIdx.y; i < 32; i += 8)
        for (int j = threadIdx.x; j < 64; j += 32) {
            const int row = r0 + i, t = 2 * tau0 + j;
            tile[i][j] = (row < F && t < T) ? x[((size_t)b * F + row) * T + t] : 0.f;
        }
    __syncthreads();
    // write: for each tau, 32 rows x 2 phases = 64 consecutive channels (coalesced along channels)
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int tau = tau0 + i;
        if (tau >= Tp) continue;
        for (int c = threadIdx.x; c < 64; c += 32) {
            const int row = r0 + (c >> 1), ph = c & 1;
            if (row * 2 + ph < kC1Ld && row < F + 1) {
                const float v = (row < F) ? tile[c >> 1][2 * i + ph] : 0.f;
                if (row * 2 + ph < kC1Ld) xt[((size_t)b * Tp + tau) * kC1Ld + row * 2 + ph] = v;
            }
        }
    }
}

// XP[q][b][ph][row][tau] = x[b][row][2 (tau - q) + ph]  (zero outside), row stride ldp (multiple of 4, >= Tp + 3)
__global__ void conv1_make_xp_kernel(const float* __restrict__ x, float* __restrict__ xp, int B, int F, int T, int Tp,
                                     int ldp) {
    const long long per = (long long)B * 2 * F * ldp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const int tau = (int)(i % ldp);
        long long r = i / ldp;
        const int row = (int)(r % F);
        r /= F;
        const int ph = (int)(r % 2), b = (int)(r / 2);
        const float* xr = x + ((size_t)b * F + row) * T;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = 2 * (tau - q) + ph;
            xp[q * per + i] = (tau - q >= 0 && t < T) ? xr[t] : 0.f;
        }
    }
}

// W'[s][kb][co][ch]: ch = kb*32 + (kh*2 + ph) ; kw = 2s+5 (even phase) / 2s+6 (odd phase), s = tap-3 ; zero elsewhere
__global__ void conv1_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int KH) {
    const int total = kC1Taps * kC1Kb * 32 * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % 32, co = (i / 32) % 32, kb = (i / 1024) % kC1Kb, tap = i / (1024 * kC1Kb);
        const int ch = kb * 32 + c, kh = ch >> 1, ph = ch & 1, s = tap - 3;
        const int kw = ph ? 2 * s + 6 : 2 * s + 5;
        float v = 0.f;
        if (kh < KH && kw >= 0 && kw < 11) v = w[((size_t)co * KH + kh) * 11 + kw];
        wp[i] = v;
    }
}

__global__ void __launch_bounds__(kC1Threads, 1)
conv1_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const Conv1Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kStages = 2;
    constexpr int kStageBytes = kC1Kb * kC1Strip;
    uint8_t* smem_w = smem;                               // [tap][kb][32 co][32 ch]
    uint8_t* smem_a = smem + kC1WBytes;                   // stages x kb x strip
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + kStages * kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* w_bar = bars + 2 * kStages;
    uint64_t* tfull_bar = w_bar + 1;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    constexpr int kTmemCols = 64;                         // 2 accumulator buffers x 32 channels

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(w_bar, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kStripBytes = (128 + kC1Taps - 1) * 128;

    auto decode = [&](int item, int& b, int& d, int& t0) {
        const int tt = item % p.ttiles;
        const int r = item / p.ttiles;
        d = r % p.Hout;
        b = r / p.Hout;
        t0 = tt * 128;
    };

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, kC1WBytes);
            for (int i = 0; i < kC1Taps * kC1Kb; ++i) tma_load_2d(smem_w + i * 4096, &tmW, w_bar, 0, i * 32);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int b, d, t0;
            decode(item, b, d, t0);
            const int c0 = (d * 2 - p.PH) * 2;            // first channel (row, phase) this output row reads
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[stage], kC1Kb * kStripBytes);
                for (int kb = 0; kb < kC1Kb; ++kb)
                    tma_load_3d(smem_a + stage * kStageBytes + kb * kC1Strip, &tmX, &full_bar[stage], c0 + kb * 32, t0 - 3, b);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(kFmtTF32, 128, 32);
        mbar_wait(w_bar, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(smem_a + stage * kStageBytes);
                const uint32_t w0 = smem_u32(smem_w);
                const uint32_t d_tmem = tmem_base + acc * 32;
                for (int tap = 0; tap < kC1Taps; ++tap)
                    for (int kb = 0; kb < kC1Kb; ++kb) {
                        const uint64_t adesc = umma_desc_sw128(a0 + kb * kC1Strip + tap * 128);   // shift by `tap` time steps
                        const uint64_t bdesc = umma_desc_sw128(w0 + (tap * kC1Kb + kb) * 4096);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | kb | k) != 0);
                    }
                umma_commit(&empty_bar[stage]);
                umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int b, d, t0;
            decode(item, b, d, t0);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after_sync();
            float v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * 32, v);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            const int len = p.lengths ? p.lengths[b] : p.Wout;
            const int t = t0 + quad * 32 + lane;
            if (t < p.Wout) {
                float* o = p.out + (((size_t)b * 32) * p.Hout + d) * p.Wout + t;
                const size_t cstride = (size_t)p.Hout * p.Wout;
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c * cstride] = t < len ? v[c] + (p.bias ? __ldg(p.bias + c) : 0.f) : 0.f;
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
// weight gradient
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kC1Threads, 1)
conv1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmXP, const __grid_constant__ CUtensorMap tmDy,
                      const Conv1Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // A stage: 12 boxes (tap, phase) of KH rows x 128 B laid back to back = (12*KH) rows -> 4 M tiles of 128 rows
    constexpr int kABytes = 4 * 128 * 128;                 // 64 KB
    constexpr int kStageBytes = kABytes + 4096;            // + dy tile [32 co][32 t]
    constexpr int kStages = 3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    constexpr int kTmemCols = 128;                         // 4 M tiles x 32 channels

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrows = p.B * p.Hout;
    const int r0 = blockIdx.x * p.rows_per_chunk, r1 = min(nrows, r0 + p.rows_per_chunk);
    const int n_kb = ceil_div(p.Wout, 32);
    const int steps = (r1 - r0) * n_kb;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmXP);
        tma_prefetch_desc(&tmDy);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    // rows of the A stage that no TMA box ever writes must not hold NaN patterns (their products land in unused
    // accumulator rows, but keep them finite anyway)
    for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kC1Threads)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t box_bytes = (uint32_t)p.KH * 128u;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int r = r0; r < r1; ++r) {
            const int b = r / p.Hout, d = r % p.Hout;
            const int row0 = d * 2 - p.PH;
            for (int kb = 0; kb < n_kb; ++kb) {
                uint8_t* st = smem + stage * kStageBytes;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], 12 * box_bytes + 4096);
                    for (int tap = 0; tap < kC1Taps; ++tap) {
                        const int s = tap - 3;
                        const int q = (((-s) % 4) + 4) % 4;      // delayed copy that makes the time coordinate 16-byte aligned
                        for (int ph = 0; ph < 2; ++ph)
                            tma_load_5d(st + (size_t)(tap * 2 + ph) * box_bytes, &tmXP, &full_bar[stage], kb * 32 + s + q, row0, ph, b, q);
                    }
                    tma_load_4d(st + kABytes, &tmDy, &full_bar[stage], kb * 32, d, 0, b);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(kFmtTF32, 128, 32);
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s < steps; ++s) {
            uint8_t* st = smem + stage * kStageBytes;
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            if (elect_one()) {
                const uint64_t bdesc = umma_desc_sw128(smem_u32(st + kABytes));
                for (int mt = 0; mt < 4; ++mt) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(st + mt * 16384));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(tmem_base + mt * 32, adesc + 2 * k, bdesc + 2 * k, idesc, (s | k) != 0);
                }
                umma_commit(&empty_bar[stage]);
                if (s == steps - 1) umma_commit(tfull_bar);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else if (steps > 0) {
        const int quad = warp & 3;
        mbar_wait(tfull_bar, 0);
        tc_fence_after_sync();
        for (int mt = 0; mt < 4; ++mt) {
            float v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + mt * 32, v);
            tmem_ld_wait();
            const int m = mt * 128 + quad * 32 + lane;       // accumulator row = ((tap, phase), kh)
            const int box = m / p.KH, kh = m % p.KH;
            if (box < 12) {
                const int s = box / 2 - 3, ph = box & 1;
                const int kw = ph ? 2 * s + 6 : 2 * s + 5;
                if (kw >= 0 && kw < 11) {
#pragma unroll
                    for (int co = 0; co < 32; ++co) atomicAdd(p.dw + ((size_t)co * p.KH + kh) * 11 + kw, v[co]);
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

static int conv1_geometry_ok(int Cin, int Cout, int KH, int KW, int SH, int SW, int PH, int PW) {
    return Cin == 1 && Cout == 32 && KW == 11 && SW == 2 && PW == 5 && SH == 2 && PH % 2 == 0 && KH >= 1 &&
           2 * KH <= 32 * kC1Kb && 12 * KH <= 512;
}

}  // namespace asrb

using namespace asrb;

extern "C" {

int asrb_conv1_supported(int Cin, int Cout, int F, int KH, int KW, int SH, int SW, int PH, int PW) {
    return conv1_geometry_ok(Cin, Cout, KH, KW, SH, SW, PH, PW) && 2 * F <= kC1Ld;
}

size_t asrb_conv1_workspace_bytes(int B, int F, int T, int for_wgrad) {
    const int Tp = (T + 1) / 2;
    if (!for_wgrad) return (size_t)B * Tp * kC1Ld * 4 + (size_t)kC1WBytes;
    return (size_t)4 * B * 2 * F * round_up(Tp + 3, 4) * 4;
}

/* y[B,32,Hout,Wout] = mask(conv1(x) + bias), x [B,1,F,T] */
int asrb_conv1_fwd(const float* x, const float* w, const float* bias, const int32_t* lengths, float* y, float* ws,
                   size_t ws_bytes, int B, int F, int T, int Hout, int Wout, int KH, int PH, asrb_stream_t stream) {
    ASRB_REQUIRE(x && w && y && ws && B > 0 && B <= 65535, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv1_supported(1, 32, F, KH, 11, 2, 2, PH, 5), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (F + 2 * PH - KH) / 2 + 1 && Wout == (T + 10 - 11) / 2 + 1, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(ws_bytes >= asrb_conv1_workspace_bytes(B, F, T, 0), ASRB_ERR_WORKSPACE);
    const int Tp = (T + 1) / 2;
    float* xt = ws;
    float* wp = ws + (size_t)B * Tp * kC1Ld;
    conv1_make_xt_kernel<<<dim3(ceil_div(Tp, 32), ceil_div(kC1Ld / 2, 32), B), dim3(32, 8), 0, stream>>>(x, xt, B, F, T, Tp);
    ASRB_LAUNCH_OK();
    conv1_pack_weights_kernel<<<72, 256, 0, stream>>>(w, wp, KH);
    ASRB_LAUNCH_OK();
    CUtensorMap tmX, tmW;
    {
        uint64_t d[3] = {(uint64_t)2 * F, (uint64_t)Tp, (uint64_t)B};
        uint64_t s[2] = {(uint64_t)kC1Ld * 4, (uint64_t)Tp * kC1Ld * 4};
        uint32_t bx[3] = {32, 128 + kC1Taps - 1, 1};
        int rc = make_tmap_f32(&tmX, xt, 3, d, s, bx);
        if (rc) return rc;
    }
    {
        uint64_t d[2] = {32, (uint64_t)kC1Taps * kC1Kb * 32}, s[1] = {128};
        uint32_t bx[2] = {32, 32};
        int rc = make_tmap_f32(&tmW, wp, 2, d, s, bx);
        if (rc) return rc;
    }
    Conv1Params p = {};
    p.B = B; p.F = F; p.T = T; p.Tp = Tp; p.Hout = Hout; p.Wout = Wout; p.KH = KH; p.PH = PH;
    p.bias = bias; p.lengths = lengths; p.out = y;
    p.ttiles = ceil_div(Wout, 128);
    p.num_items = B * Hout * p.ttiles;
    const size_t smem = (size_t)kC1WBytes + 2 * kC1Kb * kC1Strip + 1024 + 256;
    ASRB_CUDA_OK(cudaFuncSetAttribute(conv1_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_items < kNumSMs ? p.num_items : kNumSMs;
    conv1_fwd_tc_kernel<<<grid, kC1Threads, smem, stream>>>(tmX, tmW, p);
    ASRB_LAUNCH_OK();
    return 0;
}

/* dw[32,1,KH,11] from x [B,1,F,T] and dy [B,32,Hout,Wout] (already masked; row stride lddy multiple of 4) */
int asrb_conv1_bwd_weight(const float* x, const float* dy, int lddy, float* dw, float* ws, size_t ws_bytes, int B,
                          int F, int T, int Hout, int Wout, int KH, int PH, asrb_stream_t stream) {
    ASRB_REQUIRE(x && dy && dw && ws && B > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(asrb_conv1_supported(1, 32, F, KH, 11, 2, 2, PH, 5), ASRB_ERR_UNSUPPORTED);
    ASRB_REQUIRE(Hout == (F + 2 * PH - KH) / 2 + 1 && Wout == (T + 10 - 11) / 2 + 1, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(lddy >= Wout && lddy % 4 == 0, ASRB_ERR_ALIGNMENT);
    ASRB_REQUIRE(ws_bytes >= asrb_conv1_workspace_bytes(B, F, T, 1), ASRB_ERR_WORKSPACE);
    const int Tp = (T + 1) / 2, ldp = round_up(Tp + 3, 4);
    const long long per = (long long)B * 2 * F * ldp;
    {
        const int g = (int)((per + 255) / 256 < kNumSMs * 8 ? (per + 255) / 256 : kNumSMs * 8);
        conv1_make_xp_kernel<<<g, 256, 0, stream>>>(x, ws, B, F, T, Tp, ldp);
        ASRB_LAUNCH_OK();
    }
    CUtensorMap tmXP, tmDy;
    {
        uint64_t d[5] = {(uint64_t)Tp + 3, (uint64_t)F, 2, (uint64_t)B, 4};
        uint64_t s[4] = {(uint64_t)ldp * 4, (uint64_t)F * ldp * 4, (uint64_t)2 * F * ldp * 4, (uint64_t)per * 4};
        uint32_t bx[5] = {32, (uint32_t)KH, 1, 1, 1};
        int rc = make_tmap_f32(&tmXP, ws, 5, d, s, bx);
        if (rc) return rc;
    }
    {
        uint64_t d[4] = {(uint64_t)Wout, (uint64_t)Hout, 32, (uint64_t)B};
        uint64_t s[3] = {(uint64_t)lddy * 4, (uint64_t)Hout * lddy * 4, (uint64_t)32 * Hout * lddy * 4};
        uint32_t bx[4] = {32, 1, 32, 1};
        int rc = make_tmap_f32(&tmDy, dy, 4, d, s, bx);
        if (rc) return rc;
    }
    Conv1Params p = {};
    p.B = B; p.F = F; p.T = T; p.Tp = Tp; p.Hout = Hout; p.Wout = Wout; p.KH = KH; p.PH = PH; p.dw = dw;
    const int nrows = B * Hout;
    int chunks = nrows < kNumSMs ? nrows : kNumSMs;
    p.rows_per_chunk = ceil_div(nrows, chunks);
    chunks = ceil_div(nrows, p.rows_per_chunk);
    ASRB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)32 * KH * 11 * sizeof(float), stream));
    const size_t smem = 3 * (size_t)(4 * 128 * 128 + 4096) + 1024 + 256;
    ASRB_CUDA_OK(cudaFuncSetAttribute(conv1_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_wgrad_tc_kernel<<<chunks, kC1Threads, smem, stream>>>(tmXP, tmDy, p);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
