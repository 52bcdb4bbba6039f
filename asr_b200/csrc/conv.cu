// asr_b200 -- MaskConv building blocks (asr_deepspeech/modules/blocks.py:42-56 over the conv stack of
// modules/deepspeech.py:59-68), NCHW = (batch, channel, frequency, time) like the reference.
//
//   conv2d_mask_fwd / bwd_data / bwd_weight : nn.Conv2d (cross-correlation, zero padding, dilation 1) with the
//       reference's time mask fused in: y[b,:,:,t >= len[b]] = 0 (and, in the backward, masked positions carry no
//       gradient).  Direct shared-memory-tiled CUDA-core kernels (round-1 baseline; conv2 moves to a tcgen05
//       implicit GEMM next).
//   bn_act_mask_* : nn.BatchNorm2d (batch statistics INCLUDING the zeroed tail, as the reference computes them)
//       -> mask -> nn.Hardtanh(lo,hi) -> mask, forward and backward, statistics via two-stage deterministic
//       reductions.
//   transpose_batched : the [B,C,D,T] -> [T,B,C*D] layout change of deepspeech.py:135-137 and its inverse.
#include "common.cuh"

namespace asrb {

constexpr int kConvTW = 128;   // output time positions per block
constexpr int kConvCO = 32;    // output channels per pass

struct ConvDims {
    int B, Cin, Hin, Win, Cout, Hout, Wout, KH, KW, SH, SW, PH, PW;
};

// ------------------------------------------------------------------------------------------------
// forward: block = (time tile, output row, batch); thread = 2 time positions x 8 output channels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                const int* __restrict__ lengths, float* __restrict__ y, ConvDims d) {
    extern __shared__ float sm[];
    const int seg = (kConvTW - 1) * d.SW + d.KW;
    float* xs = sm;                       // [KH][seg]
    float* ws = sm + d.KH * seg;          // [32][KH*KW]
    const int t0 = blockIdx.x * kConvTW, ho = blockIdx.y, b = blockIdx.z;
    const int tx = threadIdx.x & 63, cg = threadIdx.x >> 6;
    const int khkw = d.KH * d.KW;
    const int len = lengths ? lengths[b] : d.Wout;
    const int win0 = t0 * d.SW - d.PW;

    for (int co0 = 0; co0 < d.Cout; co0 += kConvCO) {
        float acc0[8], acc1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc0[c] = acc1[c] = 0.f;
        for (int ci = 0; ci < d.Cin; ++ci) {
            __syncthreads();
            for (int i = threadIdx.x; i < d.KH * seg; i += 256) {
                const int kh = i / seg, o = i % seg;
                const int hi = ho * d.SH + kh - d.PH, wi = win0 + o;
                float v = 0.f;
                if (hi >= 0 && hi < d.Hin && wi >= 0 && wi < d.Win)
                    v = x[(((size_t)b * d.Cin + ci) * d.Hin + hi) * d.Win + wi];
                xs[i] = v;
            }
            for (int i = threadIdx.x; i < kConvCO * khkw; i += 256) {
                const int c = i / khkw, k = i % khkw;
                ws[i] = (co0 + c < d.Cout) ? w[((size_t)(co0 + c) * d.Cin + ci) * khkw + k] : 0.f;
            }
            __syncthreads();
            const float* wrow = ws + (cg * 8) * khkw;
            for (int kh = 0; kh < d.KH; ++kh) {
                const float* xr = xs + kh * seg;
                for (int kw = 0; kw < d.KW; ++kw) {
                    const float x0 = xr[tx * d.SW + kw], x1 = xr[(tx + 64) * d.SW + kw];
                    const int k = kh * d.KW + kw;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float wv = wrow[c * khkw + k];
                        acc0[c] = fmaf(x0, wv, acc0[c]);
                        acc1[c] = fmaf(x1, wv, acc1[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int co = co0 + cg * 8 + c;
            if (co < d.Cout) {
                const float bv = bias ? bias[co] : 0.f;
                float* yr = y + (((size_t)b * d.Cout + co) * d.Hout + ho) * d.Wout;
                const int ta = t0 + tx, tb = t0 + tx + 64;
                if (ta < d.Wout) yr[ta] = ta < len ? acc0[c] + bv : 0.f;
                if (tb < d.Wout) yr[tb] = tb < len ? acc1[c] + bv : 0.f;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. input (gather form): block = (input time tile, input row, batch);
// thread = 2 input time positions x 8 input channels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, const int* __restrict__ lengths,
                     float* __restrict__ dx, ConvDims d) {
    extern __shared__ float sm[];
    const int seg = (kConvTW - 1 + d.KW) / d.SW + 2;
    float* ds = sm;                       // [KH][seg]   (masked dy rows)
    float* ws = sm + d.KH * seg;          // [32 ci][KH*KW]
    const int w0 = blockIdx.x * kConvTW, hi = blockIdx.y, b = blockIdx.z;
    const int tx = threadIdx.x & 63, cg = threadIdx.x >> 6;
    const int khkw = d.KH * d.KW;
    const int len = lengths ? lengths[b] : d.Wout;
    // first output column that can touch this tile: t = ceil((w0 + PW - (KW-1)) / SW)
    int tlo = w0 + d.PW - (d.KW - 1);
    tlo = tlo >= 0 ? (tlo + d.SW - 1) / d.SW : -((-tlo) / d.SW);

    for (int ci0 = 0; ci0 < d.Cin; ci0 += kConvCO) {
        float acc0[8], acc1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc0[c] = acc1[c] = 0.f;
        for (int co = 0; co < d.Cout; ++co) {
            __syncthreads();
            for (int i = threadIdx.x; i < d.KH * seg; i += 256) {
                const int kh = i / seg, o = i % seg;
                const int num = hi + d.PH - kh;
                float v = 0.f;
                if (num >= 0 && num % d.SH == 0) {
                    const int ho = num / d.SH, t = tlo + o;
                    if (ho < d.Hout && t >= 0 && t < d.Wout && t < len)
                        v = dy[(((size_t)b * d.Cout + co) * d.Hout + ho) * d.Wout + t];
                }
                ds[i] = v;
            }
            for (int i = threadIdx.x; i < kConvCO * khkw; i += 256) {
                const int c = i / khkw, k = i % khkw;
                ws[i] = (ci0 + c < d.Cin) ? w[((size_t)co * d.Cin + ci0 + c) * khkw + k] : 0.f;
            }
            __syncthreads();
            const float* wrow = ws + (cg * 8) * khkw;
            for (int kh = 0; kh < d.KH; ++kh) {
                const int num = hi + d.PH - kh;
                if (num < 0 || num % d.SH != 0 || num / d.SH >= d.Hout) continue;
                const float* dr = ds + kh * seg;
                for (int kw = 0; kw < d.KW; ++kw) {
                    const int na = w0 + tx + d.PW - kw, nb = na + 64;
                    float v0 = 0.f, v1 = 0.f;
                    if (na >= 0 && na % d.SW == 0) { const int o = na / d.SW - tlo; if (o >= 0 && o < seg) v0 = dr[o]; }
                    if (nb >= 0 && nb % d.SW == 0) { const int o = nb / d.SW - tlo; if (o >= 0 && o < seg) v1 = dr[o]; }
                    const int k = kh * d.KW + kw;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float wv = wrow[c * khkw + k];
                        acc0[c] = fmaf(v0, wv, acc0[c]);
                        acc1[c] = fmaf(v1, wv, acc1[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ci = ci0 + cg * 8 + c;
            if (ci < d.Cin) {
                float* xr = dx + (((size_t)b * d.Cin + ci) * d.Hin + hi) * d.Win;
                const int wa = w0 + tx, wb = wa + 64;
                if (wa < d.Win) xr[wa] = acc0[c];
                if (wb < d.Win) xr[wb] = acc1[c];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. weights: block = ((ci, kh), chunk of (b, ho) rows); thread = (co, 16-wide time strip),
// 16 kernel columns accumulated in registers; partial sums combined with atomicAdd.
// ------------------------------------------------------------------------------------------------
template <int SW>
__global__ void __launch_bounds__(256)
conv_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x, const int* __restrict__ lengths,
                       float* __restrict__ dw, ConvDims d, int rows_per_block) {
    constexpr int kSeg = (kConvTW - 1) * SW + 16;
    __shared__ float dys[kConvCO][kConvTW + 1];
    __shared__ float xs[kSeg + 16];
    const int ci = blockIdx.x / d.KH, kh = blockIdx.x % d.KH;
    const int co_l = threadIdx.x >> 3, tq = threadIdx.x & 7;
    const int nrows = d.B * d.Hout;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(nrows, r0 + rows_per_block);

    for (int co0 = 0; co0 < d.Cout; co0 += kConvCO) {
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.f;
        for (int r = r0; r < r1; ++r) {
            const int b = r / d.Hout, ho = r % d.Hout;
            const int hi = ho * d.SH + kh - d.PH;
            if (hi < 0 || hi >= d.Hin) continue;
            const int len = lengths ? min(lengths[b], d.Wout) : d.Wout;
            for (int t0 = 0; t0 < len; t0 += kConvTW) {
                __syncthreads();
                for (int i = threadIdx.x; i < kConvCO * kConvTW; i += 256) {
                    const int c = i / kConvTW, t = i % kConvTW;
                    float v = 0.f;
                    if (co0 + c < d.Cout && t0 + t < len)
                        v = dy[(((size_t)b * d.Cout + co0 + c) * d.Hout + ho) * d.Wout + t0 + t];
                    dys[c][t] = v;
                }
                const int win0 = t0 * SW - d.PW;
                for (int i = threadIdx.x; i < kSeg + 16; i += 256) {
                    const int wi = win0 + i;
                    xs[i] = (wi >= 0 && wi < d.Win) ? x[(((size_t)b * d.Cin + ci) * d.Hin + hi) * d.Win + wi] : 0.f;
                }
                __syncthreads();
                float xw[15 * SW + 16];
#pragma unroll
                for (int i = 0; i < 15 * SW + 16; ++i) xw[i] = xs[tq * 16 * SW + i];
#pragma unroll
                for (int tt = 0; tt < 16; ++tt) {
                    const float a = dys[co_l][tq * 16 + tt];
#pragma unroll
                    for (int k = 0; k < 16; ++k) acc[k] = fmaf(a, xw[tt * SW + k], acc[k]);
                }
            }
        }
        // reduce the 8 time strips (adjacent lanes) and accumulate
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float v = acc[k];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (tq == 0 && k < d.KW && co0 + co_l < d.Cout && v != 0.f)
                atomicAdd(dw + (((size_t)(co0 + co_l) * d.Cin + ci) * d.KH + kh) * d.KW + k, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-channel reductions over NCHW (two-stage, deterministic)
//   MODE 0: (sum x, sum x^2)                                     -- batch-norm statistics
//   MODE 1: (sum g, sum g*xhat), g = dz*[t<len]*[lo<yhat<hi]     -- batch-norm backward
//   MODE 2: (sum dy*[t<len], 0)                                  -- conv bias gradient
// ------------------------------------------------------------------------------------------------
struct BnAct {
    const float* mean; const float* invstd; const float* gamma; const float* beta;
    float lo, hi; int has_bn, has_act;
};

__device__ __forceinline__ float bn_hat(const BnAct& p, int c, float y, float* xhat) {
    if (p.has_bn) {
        const float xh = (y - p.mean[c]) * p.invstd[c];
        *xhat = xh;
        return xh * p.gamma[c] + p.beta[c];
    }
    *xhat = y;
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(256)
nchw_reduce_kernel(const float* __restrict__ a, const float* __restrict__ yraw, const int* __restrict__ lengths,
                   BnAct p, double* __restrict__ partial, int B, int C, int HW, int W, int nsplit, unsigned wmagic) {
    const int c = blockIdx.x, b = blockIdx.y / nsplit, sp = blockIdx.y % nsplit;
    const int len = lengths ? lengths[b] : W;
    const size_t base = ((size_t)b * C + c) * HW;
    const int chunk = ceil_div(HW, nsplit);
    const int i0 = sp * chunk, i1 = min(HW, i0 + chunk);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
        const float v = a[base + i];
        const int w = i - (int)__umulhi((unsigned)i, wmagic) * W;
        if (MODE == 0) { s0 += v; s1 += v * v; }
        else if (MODE == 2) { if (w < len) s0 += v; }
        else {
            float xh;
            const float yh = bn_hat(p, c, yraw[base + i], &xh);
            const bool pass = (w < len) && (!p.has_act || (yh > p.lo && yh < p.hi));
            if (pass) { s0 += v; s1 += v * xh; }
        }
    }
    __shared__ float r0[8], r1[8];
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s0; r1[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0;
        for (int i = 0; i < 8; ++i) { t0 += r0[i]; t1 += r1[i]; }
        partial[((size_t)c * gridDim.y + blockIdx.y) * 2 + 0] = t0;
        partial[((size_t)c * gridDim.y + blockIdx.y) * 2 + 1] = t1;
    }
}

// combine partials -> out0[c], out1[c]; kind 0: mean / invstd (+ running stats), kind 1: raw sums
// one warp per channel: the lanes stride over the parts, fixed-order shuffle tree in double
__global__ void __launch_bounds__(128)
reduce_finalize_kernel(const double* __restrict__ partial, int nparts, int C, double count, int kind,
                       float eps, float momentum, float* __restrict__ out0, float* __restrict__ out1,
                       float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s0 = 0, s1 = 0;
    for (int i = lane; i < nparts; i += 32) { s0 += partial[((size_t)c * nparts + i) * 2]; s1 += partial[((size_t)c * nparts + i) * 2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane != 0) return;
    if (kind == 0) {
        const double mean = s0 / count;
        double var = s1 / count - mean * mean;
        if (var < 0) var = 0;
        out0[c] = (float)mean;
        out1[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) {
            const double unb = count > 1 ? var * count / (count - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
        }
    } else {
        if (out0) out0[c] = (float)s0;
        if (out1) out1[c] = (float)s1;
    }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { mean[c] = rm[c]; invstd[c] = rsqrtf(rv[c] + eps); }
}

// Elementwise NCHW kernels: blockIdx.y = (b, c) plane, blockIdx.x = chunk of the plane, kEwPerThread coalesced
// elements per thread.  The time index of element i of a plane is i - W * floor(i / W) with the quotient taken as
// __umulhi(i, ceil(2^32 / W)) (exact for i * W < 2^32, checked on the host) -- the former per-element 64-bit
// division was the whole cost of these HBM-bound kernels.
constexpr int kEwPerThread = 8;

// z = mask(act(mask(bn(y))))
__global__ void __launch_bounds__(256)
bn_act_mask_fwd_kernel(const float* __restrict__ y, const int* __restrict__ lengths, BnAct p, float* __restrict__ z, int C,
                       int HW, int W, unsigned wmagic) {
    const int plane = blockIdx.y, c = plane % C, b = plane / C;
    const int len = lengths ? lengths[b] : W;
    const size_t base = (size_t)plane * HW;
    float mu = 0.f, is = 1.f, ga = 1.f, sh = 0.f;
    if (p.has_bn) { mu = p.mean[c]; is = p.invstd[c]; ga = p.gamma[c]; sh = p.beta[c]; }
    const int i0 = blockIdx.x * (256 * kEwPerThread) + threadIdx.x;
    float v[kEwPerThread];
#pragma unroll
    for (int u = 0; u < kEwPerThread; ++u) {
        const int i = i0 + u * 256;
        v[u] = i < HW ? __ldcs(y + base + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kEwPerThread; ++u) {
        const int i = i0 + u * 256;
        if (i < HW) {
            const int w = i - (int)__umulhi((unsigned)i, wmagic) * W;
            float o = 0.f;
            if (w < len) {
                o = p.has_bn ? ((v[u] - mu) * is) * ga + sh : v[u];   // same operation order as bn_hat
                if (p.has_act) o = fminf(fmaxf(o, p.lo), p.hi);
            }
            z[base + i] = o;
        }
    }
}

// dy = mask * gamma*invstd*(g - s0/R - xhat*s1/R)   (training)  |  mask * gamma*invstd*g  (eval / no stats)
__global__ void __launch_bounds__(256)
bn_act_mask_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ y, const int* __restrict__ lengths, BnAct p,
                       const float* __restrict__ s0, const float* __restrict__ s1, float inv_count, int training,
                       float* __restrict__ dy, int C, int HW, int W, unsigned wmagic) {
    const int plane = blockIdx.y, c = plane % C, b = plane / C;
    const int len = lengths ? lengths[b] : W;
    const size_t base = (size_t)plane * HW;
    float a0 = 0.f, a1 = 0.f, gi = 1.f;
    if (p.has_bn) {
        gi = p.gamma[c] * p.invstd[c];
        if (training) { a0 = s0[c] * inv_count; a1 = s1[c] * inv_count; }
    }
    const int i0 = blockIdx.x * (256 * kEwPerThread) + threadIdx.x;
    float g[kEwPerThread], yv[kEwPerThread];
#pragma unroll
    for (int u = 0; u < kEwPerThread; ++u) {
        const int i = i0 + u * 256;
        g[u] = i < HW ? __ldcs(dz + base + i) : 0.f;
        yv[u] = i < HW ? __ldcs(y + base + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kEwPerThread; ++u) {
        const int i = i0 + u * 256;
        if (i < HW) {
            const int w = i - (int)__umulhi((unsigned)i, wmagic) * W;
            float out = 0.f;
            if (w < len) {
                float xh;
                const float yh = bn_hat(p, c, yv[u], &xh);
                float gg = g[u];
                if (p.has_act && !(yh > p.lo && yh < p.hi)) gg = 0.f;
                if (p.has_bn) {
                    if (training) gg = gg - a0 - xh * a1;
                    out = gg * gi;
                } else {
                    out = gg;
                }
            }
            dy[base + i] = out;
        }
    }
}

// batched tiled transpose: out[n][c][r] = in[n][r][c] ; 64 x 64 tiles, 16-byte accesses where the strides allow
template <bool VEC_IN, bool VEC_OUT>
__global__ void __launch_bounds__(256)
transpose_batched_kernel(const float* __restrict__ in, int rows, int cols, long long ld_in, long long bs_in,
                         float* __restrict__ out, long long ld_out, long long bs_out) {
    __shared__ float tile[64][65];
    in += (size_t)blockIdx.z * bs_in;
    out += (size_t)blockIdx.z * bs_out;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k, r = idx >> 4, c = (idx & 15) * 4;
        if (r0 + r < rows) {
            const float* src = in + (size_t)(r0 + r) * ld_in + c0 + c;
            if (VEC_IN && c0 + c + 3 < cols) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src));
                tile[r][c] = v.x; tile[r][c + 1] = v.y; tile[r][c + 2] = v.z; tile[r][c + 3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) tile[r][c + e] = (c0 + c + e < cols) ? src[e] : 0.f;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k, c = idx >> 4, r = (idx & 15) * 4;
        if (c0 + c < cols) {
            float* dst = out + (size_t)(c0 + c) * ld_out + r0 + r;
            if (VEC_OUT && r0 + r + 3 < rows) {
                *reinterpret_cast<float4*>(dst) = make_float4(tile[r][c], tile[r + 1][c], tile[r + 2][c], tile[r + 3][c]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (r0 + r + e < rows) dst[e] = tile[r + e][c];
            }
        }
    }
}

// out[r][0..cols) = in[r][0..cols), out[r][cols..ld_out) = 0
__global__ void copy_rows_padded_kernel(const float* __restrict__ in, long long ld_in, float* __restrict__ out,
                                        long long ld_out, long long rows, int cols) {
    const long long total = rows * ld_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ld_out;
        const int c = (int)(i % ld_out);
        out[i] = c < cols ? in[r * ld_in + c] : 0.f;
    }
}

static int conv_dims_ok(const ConvDims& d) {
    if (d.B <= 0 || d.Cin <= 0 || d.Cout <= 0 || d.KH <= 0 || d.KW <= 0 || d.SH <= 0 || d.SW <= 0) return 0;
    if (d.Hout != (d.Hin + 2 * d.PH - d.KH) / d.SH + 1 || d.Wout != (d.Win + 2 * d.PW - d.KW) / d.SW + 1) return 0;
    return d.Hout > 0 && d.Wout > 0;
}

static inline int ew_grid(long long n) {
    long long g = (n + 255) / 256;
    return (int)(g < kNumSMs * 8 ? (g > 0 ? g : 1) : kNumSMs * 8);
}

}  // namespace asrb

using namespace asrb;

extern "C" {

int asrb_conv2d_mask_fwd(const float* x, const float* w, const float* bias, const int32_t* lengths, float* y, int B,
                         int Cin, int Hin, int Win, int Cout, int Hout, int Wout, int KH, int KW, int SH, int SW,
                         int PH, int PW, asrb_stream_t stream) {
    ConvDims d = {B, Cin, Hin, Win, Cout, Hout, Wout, KH, KW, SH, SW, PH, PW};
    ASRB_REQUIRE(x && w && y && conv_dims_ok(d), ASRB_ERR_BAD_ARG);
    const int seg = (kConvTW - 1) * SW + KW;
    const size_t smem = ((size_t)KH * seg + (size_t)kConvCO * KH * KW) * 4;
    ASRB_REQUIRE(smem <= 200 * 1024 && Hout <= 65535 && B <= 65535, ASRB_ERR_UNSUPPORTED);
    ASRB_CUDA_OK(cudaFuncSetAttribute(conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(Wout, kConvTW), Hout, B);
    conv_fwd_kernel<<<grid, 256, smem, stream>>>(x, w, bias, lengths, y, d);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_conv2d_mask_bwd_data(const float* dy, const float* w, const int32_t* lengths, float* dx, int B, int Cin,
                              int Hin, int Win, int Cout, int Hout, int Wout, int KH, int KW, int SH, int SW, int PH,
                              int PW, asrb_stream_t stream) {
    ConvDims d = {B, Cin, Hin, Win, Cout, Hout, Wout, KH, KW, SH, SW, PH, PW};
    ASRB_REQUIRE(dy && w && dx && conv_dims_ok(d), ASRB_ERR_BAD_ARG);
    const int seg = (kConvTW - 1 + KW) / SW + 2;
    const size_t smem = ((size_t)KH * seg + (size_t)kConvCO * KH * KW) * 4;
    ASRB_REQUIRE(smem <= 200 * 1024 && Hin <= 65535 && B <= 65535, ASRB_ERR_UNSUPPORTED);
    ASRB_CUDA_OK(cudaFuncSetAttribute(conv_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(Win, kConvTW), Hin, B);
    conv_bwd_data_kernel<<<grid, 256, smem, stream>>>(dy, w, lengths, dx, d);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_conv2d_mask_bwd_weight(const float* dy, const float* x, const int32_t* lengths, float* dw, float* dbias,
                                double* ws, size_t ws_bytes, int B, int Cin, int Hin, int Win, int Cout, int Hout,
                                int Wout, int KH, int KW, int SH, int SW, int PH, int PW, asrb_stream_t stream) {
    ConvDims d = {B, Cin, Hin, Win, Cout, Hout, Wout, KH, KW, SH, SW, PH, PW};
    ASRB_REQUIRE(dy && x && dw && conv_dims_ok(d), ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(KW <= 16 && (SW == 1 || SW == 2), ASRB_ERR_UNSUPPORTED);
    ASRB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * KH * KW * sizeof(float), stream));
    const int nrows = B * Hout;
    int chunks = ceil_div(kNumSMs * 4, Cin * KH);
    if (chunks > nrows) chunks = nrows;
    const int rpb = ceil_div(nrows, chunks);
    dim3 grid(Cin * KH, ceil_div(nrows, rpb));
    if (SW == 1) conv_bwd_weight_kernel<1><<<grid, 256, 0, stream>>>(dy, x, lengths, dw, d, rpb);
    else         conv_bwd_weight_kernel<2><<<grid, 256, 0, stream>>>(dy, x, lengths, dw, d, rpb);
    ASRB_LAUNCH_OK();
    if (dbias) {
        const int HW = Hout * Wout;
        const int nsplit = HW >= 4096 ? 4 : 1;
        ASRB_REQUIRE(ws && ws_bytes >= (size_t)Cout * B * nsplit * 2 * sizeof(double), ASRB_ERR_WORKSPACE);
        BnAct p = {};
        nchw_reduce_kernel<2><<<dim3(Cout, B * nsplit), 256, 0, stream>>>(dy, nullptr, lengths, p, ws, B, Cout, HW, Wout, nsplit, (unsigned)(((1ULL << 32) + Wout - 1) / Wout));
        ASRB_LAUNCH_OK();
        reduce_finalize_kernel<<<ceil_div(Cout, 4), 128, 0, stream>>>(ws, B * nsplit, Cout, 1.0, 1, 0.f, 0.f, dbias, nullptr, nullptr, nullptr);
        ASRB_LAUNCH_OK();
    }
    return 0;
}

int asrb_copy_rows_padded(const float* in, long long ld_in, float* out, long long ld_out, long long rows, int cols,
                          asrb_stream_t stream) {
    ASRB_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= cols, ASRB_ERR_BAD_ARG);
    copy_rows_padded_kernel<<<ew_grid(rows * ld_out), 256, 0, stream>>>(in, ld_in, out, ld_out, rows, cols);
    ASRB_LAUNCH_OK();
    return 0;
}

/* out[c] = sum over (b, h, w < lengths[b]) of a[b,c,h,w] */
int asrb_nchw_channel_sums(const float* a, const int32_t* lengths, float* out, double* ws, size_t ws_bytes, int B, int C,
                           int H, int W, asrb_stream_t stream) {
    ASRB_REQUIRE(a && out && ws && B > 0 && C > 0 && H > 0 && W > 0, ASRB_ERR_BAD_ARG);
    const int HW = H * W, nsplit = HW >= 4096 ? 4 : 1;
    ASRB_REQUIRE(ws_bytes >= (size_t)C * B * nsplit * 2 * sizeof(double), ASRB_ERR_WORKSPACE);
    BnAct p = {};
    nchw_reduce_kernel<2><<<dim3(C, B * nsplit), 256, 0, stream>>>(a, nullptr, lengths, p, ws, B, C, HW, W, nsplit, (unsigned)(((1ULL << 32) + W - 1) / W));
    ASRB_LAUNCH_OK();
    reduce_finalize_kernel<<<ceil_div(C, 4), 128, 0, stream>>>(ws, B * nsplit, C, 1.0, 1, 0.f, 0.f, out, nullptr, nullptr, nullptr);
    ASRB_LAUNCH_OK();
    return 0;
}

size_t asrb_nchw_reduce_workspace_bytes(int B, int C, int H, int W) {
    const int nsplit = (H * W) >= 4096 ? 4 : 1;
    return (size_t)C * B * nsplit * 2 * sizeof(double);
}

/* Batch statistics of y[B,C,H,W] (all positions): mean[C], invstd[C]; updates running stats when given. */
int asrb_bn2d_stats(const float* y, float* mean, float* invstd, float* running_mean, float* running_var,
                    float momentum, float eps, double* ws, size_t ws_bytes, int B, int C, int H, int W,
                    asrb_stream_t stream) {
    ASRB_REQUIRE(y && mean && invstd && ws && B > 0 && C > 0 && H > 0 && W > 0, ASRB_ERR_BAD_ARG);
    const int HW = H * W, nsplit = HW >= 4096 ? 4 : 1;
    ASRB_REQUIRE(ws_bytes >= (size_t)C * B * nsplit * 2 * sizeof(double), ASRB_ERR_WORKSPACE);
    BnAct p = {};
    nchw_reduce_kernel<0><<<dim3(C, B * nsplit), 256, 0, stream>>>(y, nullptr, nullptr, p, ws, B, C, HW, W, nsplit, (unsigned)(((1ULL << 32) + W - 1) / W));
    ASRB_LAUNCH_OK();
    reduce_finalize_kernel<<<ceil_div(C, 4), 128, 0, stream>>>(ws, B * nsplit, C, (double)B * HW, 0, eps, momentum, mean, invstd, running_mean, running_var);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_bn_eval_stats(const float* running_mean, const float* running_var, float eps, int C, float* mean,
                       float* invstd, asrb_stream_t stream) {
    ASRB_REQUIRE(running_mean && running_var && mean && invstd && C > 0, ASRB_ERR_BAD_ARG);
    bn_eval_stats_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(running_mean, running_var, eps, C, mean, invstd);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_bn_act_mask_fwd(const float* y, const int32_t* lengths, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, int has_bn, int has_act, float lo, float hi, float* z,
                         int B, int C, int H, int W, asrb_stream_t stream) {
    ASRB_REQUIRE(y && z && B > 0 && C > 0 && H > 0 && W > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(!has_bn || (mean && invstd && gamma && beta), ASRB_ERR_BAD_ARG);
    BnAct p = {mean, invstd, gamma, beta, lo, hi, has_bn, has_act};
    const long long total = (long long)B * C * H * W;
    (void)total;
    ASRB_REQUIRE((long long)H * W * W < (1LL << 32) && (long long)B * C <= 65535, ASRB_ERR_UNSUPPORTED);
    const unsigned wmagic = (unsigned)(((1ULL << 32) + W - 1) / W);
    bn_act_mask_fwd_kernel<<<dim3(ceil_div(H * W, 256 * kEwPerThread), B * C), 256, 0, stream>>>(y, lengths, p, z, C, H * W, W, wmagic);
    ASRB_LAUNCH_OK();
    return 0;
}

/* dz -> dy (gradient w.r.t. the masked conv output), plus dgamma[C], dbeta[C] when has_bn. */
int asrb_bn_act_mask_bwd(const float* dz, const float* y, const int32_t* lengths, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, int has_bn, int has_act, float lo,
                         float hi, int training, float* dy, float* dgamma, float* dbeta, double* ws, size_t ws_bytes,
                         int B, int C, int H, int W, asrb_stream_t stream) {
    ASRB_REQUIRE(dz && y && dy && B > 0 && C > 0 && H > 0 && W > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(!has_bn || (mean && invstd && gamma && beta && dgamma && dbeta && ws), ASRB_ERR_BAD_ARG);
    BnAct p = {mean, invstd, gamma, beta, lo, hi, has_bn, has_act};
    const int HW = H * W, nsplit = HW >= 4096 ? 4 : 1;
    if (has_bn) {
        ASRB_REQUIRE(ws_bytes >= (size_t)C * B * nsplit * 2 * sizeof(double), ASRB_ERR_WORKSPACE);
        nchw_reduce_kernel<1><<<dim3(C, B * nsplit), 256, 0, stream>>>(dz, y, lengths, p, ws, B, C, HW, W, nsplit, (unsigned)(((1ULL << 32) + W - 1) / W));
        ASRB_LAUNCH_OK();
        reduce_finalize_kernel<<<ceil_div(C, 4), 128, 0, stream>>>(ws, B * nsplit, C, 1.0, 1, 0.f, 0.f, dbeta, dgamma, nullptr, nullptr);
        ASRB_LAUNCH_OK();
    }
    const long long total = (long long)B * C * HW;
    (void)total;
    ASRB_REQUIRE((long long)HW * W < (1LL << 32) && (long long)B * C <= 65535, ASRB_ERR_UNSUPPORTED);
    const unsigned wmagic = (unsigned)(((1ULL << 32) + W - 1) / W);
    bn_act_mask_bwd_kernel<<<dim3(ceil_div(HW, 256 * kEwPerThread), B * C), 256, 0, stream>>>(dz, y, lengths, p, dbeta, dgamma, 1.0f / ((float)B * HW), training, dy, C, HW, W, wmagic);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_transpose_batched(const float* in, int rows, int cols, long long ld_in, long long batch_stride_in, float* out,
                           long long ld_out, long long batch_stride_out, int nbatch, asrb_stream_t stream) {
    ASRB_REQUIRE(in && out && rows > 0 && cols > 0 && nbatch > 0 && nbatch <= 65535, ASRB_ERR_BAD_ARG);
    dim3 grid(ceil_div(cols, 64), ceil_div(rows, 64), nbatch);
    ASRB_REQUIRE(grid.y <= 65535, ASRB_ERR_UNSUPPORTED);
    const bool vin = ld_in % 4 == 0 && batch_stride_in % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    const bool vout = ld_out % 4 == 0 && batch_stride_out % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (vin && vout)  transpose_batched_kernel<true, true><<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, batch_stride_in, out, ld_out, batch_stride_out);
    else if (vin)     transpose_batched_kernel<true, false><<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, batch_stride_in, out, ld_out, batch_stride_out);
    else if (vout)    transpose_batched_kernel<false, true><<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, batch_stride_in, out, ld_out, batch_stride_out);
    else              transpose_batched_kernel<false, false><<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, batch_stride_in, out, ld_out, batch_stride_out);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
