// asr_b200 -- row-matrix kernels on [R = T*N, H] activations:
//   * SequenceWise(BatchNorm1d) forward / backward (asr_deepspeech/modules/blocks.py:16-21,85-86 and the FC head's
//     BatchNorm1d, modules/deepspeech.py:104): batch statistics over ALL T*N rows, padded (zero) rows included,
//     exactly as the reference computes them.
//   * column sums (bias gradients of the recurrent layers)
//   * log_softmax / softmax / argmax over the class dimension (trainers/deepspeech_trainer.py:110,
//     blocks.py:62, decoders/greedy_decoder.py:61) and the log_softmax backward.
#include <cuda_bf16.h>

#include "common.cuh"

namespace asrb {

constexpr int kRowChunks = 296;  // 2 x 148 SMs worth of row chunks for the two-stage column reductions

// partial[chunk][2][cols]: MODE 0 (sum x, sum x^2) ; MODE 1 (sum dy, sum dy*xhat) ; MODE 2 (sum x, -)
template <int MODE>
__global__ void __launch_bounds__(256)
rows_reduce_kernel(const float* __restrict__ a, int lda, const float* __restrict__ x, int ldx,
                   const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ partial,
                   long long R, int cols, int rows_per_chunk) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    const long long r1 = r0 + rows_per_chunk < R ? r0 + rows_per_chunk : R;
    float s0 = 0.f, s1 = 0.f;
    if (col < cols) {
        float mu = 0.f, is = 1.f;
        if (MODE == 1) { mu = mean[col]; is = invstd[col]; }
        for (long long r = r0 + ty; r < r1; r += 8) {
            const float v = a[r * lda + col];
            if (MODE == 0) { s0 += v; s1 += v * v; }
            else if (MODE == 2) { s0 += v; }
            else { s0 += v; s1 += v * (x[r * ldx + col] - mu) * is; }
        }
    }
    __shared__ float sh0[8][33], sh1[8][33];
    sh0[ty][tx] = s0; sh1[ty][tx] = s1;
    __syncthreads();
    if (ty == 0 && col < cols) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { s0 += sh0[i][tx]; s1 += sh1[i][tx]; }
        partial[((size_t)blockIdx.y * 2 + 0) * cols + col] = s0;
        partial[((size_t)blockIdx.y * 2 + 1) * cols + col] = s1;
    }
}

// kind 0: mean/invstd (+running stats) ; kind 1: raw sums
// block = 32 columns x 8 chunk groups: the chunk loop is 8x shorter than with one thread per column (the kernel is pure
// latency: 296 dependent-free but serially issued iterations cost 50 us), partials combined in a fixed order
__global__ void __launch_bounds__(256)
rows_finalize_kernel(const float* __restrict__ partial, int nchunks, int cols, double count, int kind,
                     float eps, float momentum, float* __restrict__ out0, float* __restrict__ out1,
                     float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    __shared__ double sh0[8][33], sh1[8][33];
    double s0 = 0, s1 = 0;
    if (c < cols) {
#pragma unroll 4
        for (int i = ty; i < nchunks; i += 8) {
            s0 += partial[((size_t)i * 2 + 0) * cols + c];
            s1 += partial[((size_t)i * 2 + 1) * cols + c];
        }
    }
    sh0[ty][tx] = s0; sh1[ty][tx] = s1;
    __syncthreads();
    if (ty != 0 || c >= cols) return;
#pragma unroll
    for (int i = 1; i < 8; ++i) { s0 += sh0[i][tx]; s1 += sh1[i][tx]; }
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

// Row-matrix elementwise kernels: blockIdx.y = chunk of rows, threads stride over the columns in 16-byte vectors
// (cols % 4 == 0) or scalars; per-column constants are loaded once per thread and reused for every row of the chunk.
constexpr int kRowsPerEwBlock = 32;

template <bool VEC>
__global__ void __launch_bounds__(256)
bn_rows_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, long long R, int cols) {
    constexpr int V = VEC ? 4 : 1;
    const int c = (blockIdx.x * 256 + threadIdx.x) * V;
    if (c >= cols) return;
    float mu[V], is[V], ga[V], be[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { mu[e] = mean[c + e]; is[e] = invstd[c + e]; ga[e] = gamma[c + e]; be[e] = beta[c + e]; }
    const long long r0 = (long long)blockIdx.y * kRowsPerEwBlock;
    const long long r1 = r0 + kRowsPerEwBlock < R ? r0 + kRowsPerEwBlock : R;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        if constexpr (VEC) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(x + r * cols + c));
            float4 o;
            o.x = (v.x - mu[0]) * is[0] * ga[0] + be[0]; o.y = (v.y - mu[1]) * is[1] * ga[1] + be[1];
            o.z = (v.z - mu[2]) * is[2] * ga[2] + be[2]; o.w = (v.w - mu[3]) * is[3] * ga[3] + be[3];
            *reinterpret_cast<float4*>(y + r * cols + c) = o;
        } else {
            y[r * cols + c] = (x[r * cols + c] - mu[0]) * is[0] * ga[0] + be[0];
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
bn_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ s0,
                   const float* __restrict__ s1, float inv_count, int training, float* __restrict__ dx, long long R, int cols) {
    constexpr int V = VEC ? 4 : 1;
    const int c = (blockIdx.x * 256 + threadIdx.x) * V;
    if (c >= cols) return;
    float mu[V], is[V], gi[V], a0[V], a1[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mu[e] = mean[c + e]; is[e] = invstd[c + e]; gi[e] = gamma[c + e] * invstd[c + e];
        a0[e] = training ? s0[c + e] * inv_count : 0.f;
        a1[e] = training ? s1[c + e] * inv_count : 0.f;
    }
    const long long r0 = (long long)blockIdx.y * kRowsPerEwBlock;
    const long long r1 = r0 + kRowsPerEwBlock < R ? r0 + kRowsPerEwBlock : R;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        if constexpr (VEC) {
            const float4 g = __ldcs(reinterpret_cast<const float4*>(dy + r * cols + c));
            const float4 v = __ldcs(reinterpret_cast<const float4*>(x + r * cols + c));
            float4 o;
            o.x = (training ? g.x - a0[0] - (v.x - mu[0]) * is[0] * a1[0] : g.x) * gi[0];
            o.y = (training ? g.y - a0[1] - (v.y - mu[1]) * is[1] * a1[1] : g.y) * gi[1];
            o.z = (training ? g.z - a0[2] - (v.z - mu[2]) * is[2] * a1[2] : g.z) * gi[2];
            o.w = (training ? g.w - a0[3] - (v.w - mu[3]) * is[3] * a1[3] : g.w) * gi[3];
            *reinterpret_cast<float4*>(dx + r * cols + c) = o;
        } else {
            const float g = dy[r * cols + c];
            dx[r * cols + c] = (training ? g - a0[0] - (x[r * cols + c] - mu[0]) * is[0] * a1[0] : g) * gi[0];
        }
    }
}

// out[r] = sum_c a[r*ld + c]   (one block per row, fixed reduction order)
__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256)
row_sums_kernel(const T* __restrict__ a, long long ld, float* __restrict__ out, long long cols) {
    const T* row = a + (size_t)blockIdx.x * ld;
    constexpr int V = 16 / sizeof(T);                  // elements per 16-byte load
    float s = 0.f;
    long long c0 = 0;
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        const long long nv = cols / V;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};           // four independent accumulators per thread
#pragma unroll 4
        for (long long i = threadIdx.x; i < nv; i += 256) {
            const uint4 u = __ldcs(reinterpret_cast<const uint4*>(row) + i);
            if constexpr (sizeof(T) == 4) {
                acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y);
                acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
            } else {   // bf16 pairs: the low half is the first element
                acc[0] += __uint_as_float(u.x << 16) + __uint_as_float(u.x & 0xffff0000u);
                acc[1] += __uint_as_float(u.y << 16) + __uint_as_float(u.y & 0xffff0000u);
                acc[2] += __uint_as_float(u.z << 16) + __uint_as_float(u.z & 0xffff0000u);
                acc[3] += __uint_as_float(u.w << 16) + __uint_as_float(u.w & 0xffff0000u);
            }
        }
        s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        c0 = nv * V;
    }
    for (long long c = c0 + threadIdx.x; c < cols; c += 256) s += ld_as_float(row + c);
    __shared__ float red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        out[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// class-dimension kernels: one warp per row
// ------------------------------------------------------------------------------------------------
// lp = log_softmax(logits) ; optional probs = softmax ; optional idx = argmax (first maximum)
__global__ void __launch_bounds__(256)
log_softmax_kernel(const float* __restrict__ logits, int ld, float* __restrict__ lp, float* __restrict__ probs,
                   long long* __restrict__ idx, long long R, int C) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    const float* in = logits + row * ld;
    float m = -INFINITY;
    int mi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float v = in[c];
        if (v > m) { m = v; mi = c; }   // strict > keeps the first maximum within a lane's stride
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(in[c] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    for (int c = lane; c < C; c += 32) {
        const float v = in[c] - lse;
        if (lp) lp[row * C + c] = v;
        if (probs) probs[row * C + c] = expf(v);
    }
    if (idx && lane == 0) idx[row] = mi;
}

// dlogits = g - exp(lp) * sum_c g      (log_softmax backward)
__global__ void __launch_bounds__(256)
log_softmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ lp, float* __restrict__ dlogits, int ld,
                       long long R, int C) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += g[row * C + c];
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) dlogits[row * ld + c] = g[row * C + c] - expf(lp[row * C + c]) * s;
}


template <int MODE>
static int rows_reduce(const float* a, int lda, const float* x, int ldx, const float* mean, const float* invstd,
                       float* ws, size_t ws_bytes, long long R, int cols, int* nchunks_out, asrb_stream_t stream) {
    int nchunks = (int)(R < kRowChunks ? R : kRowChunks);
    const int rpc = (int)((R + nchunks - 1) / nchunks);
    nchunks = (int)((R + rpc - 1) / rpc);
    if (ws_bytes < (size_t)nchunks * 2 * cols * sizeof(float)) return ASRB_ERR_WORKSPACE;
    rows_reduce_kernel<MODE><<<dim3(ceil_div(cols, 32), nchunks), 256, 0, stream>>>(a, lda, x, ldx, mean, invstd, ws, R, cols, rpc);
    ASRB_LAUNCH_OK();
    *nchunks_out = nchunks;
    return 0;
}

}  // namespace asrb

using namespace asrb;

extern "C" {

size_t asrb_rows_workspace_bytes(int cols) { return (size_t)kRowChunks * 2 * cols * sizeof(float); }

/* y = BatchNorm1d(x) over x[R, cols].  training: batch stats -> mean/invstd (saved for backward) and running stats
 * update (momentum, unbiased variance); eval: mean/invstd from running stats.  y may alias x. */
int asrb_bn_rows_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     int training, float momentum, float eps, float* mean, float* invstd, float* y, float* ws,
                     size_t ws_bytes, long long R, int cols, asrb_stream_t stream) {
    ASRB_REQUIRE(x && gamma && beta && mean && invstd && y && R > 0 && cols > 0, ASRB_ERR_BAD_ARG);
    if (training) {
        ASRB_REQUIRE(ws, ASRB_ERR_WORKSPACE);
        int nchunks = 0;
        int rc = rows_reduce<0>(x, cols, nullptr, 0, nullptr, nullptr, ws, ws_bytes, R, cols, &nchunks, stream);
        if (rc) return rc;
        rows_finalize_kernel<<<ceil_div(cols, 32), 256, 0, stream>>>(ws, nchunks, cols, (double)R, 0, eps, momentum, mean, invstd, running_mean, running_var);
        ASRB_LAUNCH_OK();
    } else {
        ASRB_REQUIRE(running_mean && running_var, ASRB_ERR_BAD_ARG);
        int rc = asrb_bn_eval_stats(running_mean, running_var, eps, cols, mean, invstd, stream);
        if (rc) return rc;
    }
    const long long total = R * cols;
    (void)total;
    {
        const bool vec = cols % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
        const dim3 grid(ceil_div(vec ? cols / 4 : cols, 256), (unsigned)ceil_div64(R, kRowsPerEwBlock));
        if (vec) bn_rows_apply_kernel<true><<<grid, 256, 0, stream>>>(x, mean, invstd, gamma, beta, y, R, cols);
        else     bn_rows_apply_kernel<false><<<grid, 256, 0, stream>>>(x, mean, invstd, gamma, beta, y, R, cols);
    }
    ASRB_LAUNCH_OK();
    return 0;
}

/* dx, dgamma, dbeta from dy and the saved x / mean / invstd.  dx may alias dy. */
int asrb_bn_rows_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                     int training, float* dx, float* dgamma, float* dbeta, float* ws, size_t ws_bytes, long long R,
                     int cols, asrb_stream_t stream) {
    ASRB_REQUIRE(dy && x && mean && invstd && gamma && dx && dgamma && dbeta && ws && R > 0 && cols > 0, ASRB_ERR_BAD_ARG);
    int nchunks = 0;
    int rc = rows_reduce<1>(dy, cols, x, cols, mean, invstd, ws, ws_bytes, R, cols, &nchunks, stream);
    if (rc) return rc;
    rows_finalize_kernel<<<ceil_div(cols, 32), 256, 0, stream>>>(ws, nchunks, cols, 1.0, 1, 0.f, 0.f, dbeta, dgamma, nullptr, nullptr);
    ASRB_LAUNCH_OK();
    const long long total = R * cols;
    (void)total;
    {
        const bool vec = cols % 4 == 0 &&
                         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
        const dim3 grid(ceil_div(vec ? cols / 4 : cols, 256), (unsigned)ceil_div64(R, kRowsPerEwBlock));
        if (vec) bn_rows_bwd_kernel<true><<<grid, 256, 0, stream>>>(dy, x, mean, invstd, gamma, dbeta, dgamma, 1.0f / (float)R, training, dx, R, cols);
        else     bn_rows_bwd_kernel<false><<<grid, 256, 0, stream>>>(dy, x, mean, invstd, gamma, dbeta, dgamma, 1.0f / (float)R, training, dx, R, cols);
    }
    ASRB_LAUNCH_OK();
    return 0;
}

/* out[c] = sum_r a[r*lda + c] */
int asrb_col_sums(const float* a, int lda, float* out, float* ws, size_t ws_bytes, long long R, int cols,
                  asrb_stream_t stream) {
    ASRB_REQUIRE(a && out && ws && R > 0 && cols > 0 && lda >= cols, ASRB_ERR_BAD_ARG);
    int nchunks = 0;
    int rc = rows_reduce<2>(a, lda, nullptr, 0, nullptr, nullptr, ws, ws_bytes, R, cols, &nchunks, stream);
    if (rc) return rc;
    rows_finalize_kernel<<<ceil_div(cols, 32), 256, 0, stream>>>(ws, nchunks, cols, 1.0, 1, 0.f, 0.f, out, nullptr, nullptr, nullptr);
    ASRB_LAUNCH_OK();
    return 0;
}

/* out[r] = sum_c a[r*ld + c] */
int asrb_row_sums(const float* a, long long ld, float* out, int rows, long long cols, asrb_stream_t stream) {
    ASRB_REQUIRE(a && out && rows > 0 && cols > 0 && ld >= cols, ASRB_ERR_BAD_ARG);
    row_sums_kernel<float><<<rows, 256, 0, stream>>>(a, ld, out, cols);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_row_sums_bf16(const void* a, long long ld, float* out, int rows, long long cols, asrb_stream_t stream) {
    ASRB_REQUIRE(a && out && rows > 0 && cols > 0 && ld >= cols, ASRB_ERR_BAD_ARG);
    row_sums_kernel<__nv_bfloat16><<<rows, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a), ld, out, cols);
    ASRB_LAUNCH_OK();
    return 0;
}

/* logits[R, ld] (ld >= C) -> log_probs[R, C] (may be NULL), probs[R, C] (may be NULL), argmax int64[R] (may be NULL) */
int asrb_log_softmax_fwd(const float* logits, int ld, float* log_probs, float* probs, long long* argmax, long long R,
                         int C, asrb_stream_t stream) {
    ASRB_REQUIRE(logits && R > 0 && C > 0 && ld >= C, ASRB_ERR_BAD_ARG);
    log_softmax_kernel<<<(unsigned)((R + 7) / 8), 256, 0, stream>>>(logits, ld, log_probs, probs, argmax, R, C);
    ASRB_LAUNCH_OK();
    return 0;
}

/* dlogits[R, ld] = g - exp(log_probs) * rowsum(g) */
int asrb_log_softmax_bwd(const float* g, const float* log_probs, float* dlogits, int ld, long long R, int C,
                         asrb_stream_t stream) {
    ASRB_REQUIRE(g && log_probs && dlogits && R > 0 && C > 0 && ld >= C, ASRB_ERR_BAD_ARG);
    log_softmax_bwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, stream>>>(g, log_probs, dlogits, ld, R, C);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// greedy CTC collapse: decoders/greedy_decoder.py:27-46 `process_string` with
// remove_repetitions=True on the frame-wise argmax -- keep frame t iff its class is not the blank and differs from the
// class of frame t-1.  One warp per utterance, ballot/popc stream compaction in frame order.
// ------------------------------------------------------------------------------------------------
namespace asrb {
__global__ void __launch_bounds__(128)
greedy_collapse_kernel(const long long* __restrict__ idx, const int* __restrict__ sizes, int N, int T, int blank,
                       int* __restrict__ labels, int* __restrict__ offsets, int* __restrict__ counts) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const int len = sizes ? min(max(sizes[n], 0), T) : T;
    const long long* row = idx + (size_t)n * T;
    int out = 0;
    for (int t0 = 0; t0 < len; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        int c = blank;
        if (t < len) {
            c = (int)row[t];
            keep = c != blank && (t == 0 || c != (int)row[t - 1]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = out + __popc(m & ((1u << lane) - 1u));
            labels[(size_t)n * T + pos] = c;
            offsets[(size_t)n * T + pos] = t;
        }
        out += __popc(m);
    }
    if (lane == 0) counts[n] = out;
}
}  // namespace asrb

extern "C" int asrb_greedy_collapse(const long long* argmax, const int32_t* sizes, int N, int T, int blank, int32_t* labels,
                                    int32_t* offsets, int32_t* counts, asrb_stream_t stream) {
    ASRB_REQUIRE(argmax && labels && offsets && counts && N > 0 && T > 0, ASRB_ERR_BAD_ARG);
    asrb::greedy_collapse_kernel<<<asrb::ceil_div(N, 4), 128, 0, stream>>>(argmax, sizes, N, T, blank, labels, offsets, counts);
    ASRB_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Lookahead convolution: asr_deepspeech/modules/blocks.py:96-121 -- a depthwise Conv1d over
// time with `context` taps looking AHEAD, zero padded at the end: y[t,n,f] = sum_k w[f,k] x[t+k,n,f], on [T,N,H]
// activations, with the Hardtanh(0,20) that follows it in the unidirectional model (deepspeech.py:94-101) fused.
// HBM-bound: x is read once from DRAM (the `context` re-reads of a 256-column strip hit L1/L2), y written once.
// ------------------------------------------------------------------------------------------------
namespace asrb {
constexpr int kLaChunk = 32;   // time steps per block

__global__ void __launch_bounds__(256)
lookahead_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int T, int NH, int H,
                     int ctx, int has_act, float lo, float hi) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= NH) return;
    const float* wf = w + (size_t)(c % H) * ctx;
    const int t1 = min(T, (int)(blockIdx.y + 1) * kLaChunk);
    for (int t = blockIdx.y * kLaChunk; t < t1; ++t) {
        float acc = 0.f;
        const int kmax = min(ctx, T - t);
        for (int k = 0; k < kmax; ++k) acc = fmaf(__ldg(wf + k), __ldg(x + (size_t)(t + k) * NH + c), acc);
        if (has_act) acc = fminf(fmaxf(acc, lo), hi);
        y[(size_t)t * NH + c] = acc;
    }
}

// gradient entering the convolution: dy where the clamp was inactive (hardtanh_backward: 0 where the input was <= lo or
// >= hi, i.e. where the OUTPUT sits on a bound)
__device__ __forceinline__ float la_gate(const float* __restrict__ dy, const float* __restrict__ y, size_t i, int has_act,
                                         float lo, float hi) {
    const float g = __ldg(dy + i);
    if (!has_act) return g;
    const float v = __ldg(y + i);
    return (v > lo && v < hi) ? g : 0.f;
}

// dx[t,c] = sum_k w[f,k] g[t-k,c]
__global__ void __launch_bounds__(256)
lookahead_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ w,
                          float* __restrict__ dx, int T, int NH, int H, int ctx, int has_act, float lo, float hi) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= NH) return;
    const float* wf = w + (size_t)(c % H) * ctx;
    const int t1 = min(T, (int)(blockIdx.y + 1) * kLaChunk);
    for (int t = blockIdx.y * kLaChunk; t < t1; ++t) {
        float acc = 0.f;
        const int kmax = min(ctx, t + 1);
        for (int k = 0; k < kmax; ++k) acc = fmaf(__ldg(wf + k), la_gate(dy, y, (size_t)(t - k) * NH + c, has_act, lo, hi), acc);
        dx[(size_t)t * NH + c] = acc;
    }
}

// dw[f,k] = sum_{t,n} g[t,n,f] x[t+k,n,f].  Block = 32 features x 8 warps striding the batch, one time chunk; the 8
// partial sums of a (feature, tap) meet in shared memory and ONE atomic per block goes to dw.
__global__ void __launch_bounds__(256)
lookahead_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                            float* __restrict__ dw, int T, int N, int H, int ctx, int has_act, float lo, float hi) {
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + lane;
    const int t0 = blockIdx.y * kLaChunk, t1 = min(T, t0 + kLaChunk);
    const size_t NH = (size_t)N * H;
    for (int k = 0; k < ctx; ++k) {
        float s = 0.f;
        if (f < H) {
            for (int n = warp; n < N; n += 8) {
                const size_t col = (size_t)n * H + f;
                for (int t = t0; t < t1 && t + k < T; ++t)
                    s = fmaf(la_gate(dy, y, (size_t)t * NH + col, has_act, lo, hi), __ldg(x + (size_t)(t + k) * NH + col), s);
            }
        }
        part[warp][lane] = s;
        __syncthreads();
        if (warp == 0 && f < H) {
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) tot += part[i][lane];
            atomicAdd(dw + (size_t)f * ctx + k, tot);
        }
        __syncthreads();
    }
}
}  // namespace asrb

extern "C" int asrb_lookahead_fwd(const float* x, const float* w, float* y, int T, int N, int H, int context, int has_act,
                                  float lo, float hi, asrb_stream_t stream) {
    ASRB_REQUIRE(x && w && y && T > 0 && N > 0 && H > 0 && context > 0, ASRB_ERR_BAD_ARG);
    const long long NH = (long long)N * H;
    ASRB_REQUIRE(NH < (1LL << 31) && asrb::ceil_div(T, asrb::kLaChunk) <= 65535, ASRB_ERR_UNSUPPORTED);
    dim3 grid((unsigned)asrb::ceil_div((int)NH, 256), (unsigned)asrb::ceil_div(T, asrb::kLaChunk));
    asrb::lookahead_fwd_kernel<<<grid, 256, 0, stream>>>(x, w, y, T, (int)NH, H, context, has_act, lo, hi);
    ASRB_LAUNCH_OK();
    return 0;
}

/* y: the forward OUTPUT (only read when has_act); dx and/or dw may be NULL; dw [H, context] is overwritten */
extern "C" int asrb_lookahead_bwd(const float* dy, const float* x, const float* y, const float* w, float* dx, float* dw, int T,
                                  int N, int H, int context, int has_act, float lo, float hi, asrb_stream_t stream) {
    ASRB_REQUIRE(dy && x && w && (y || !has_act) && T > 0 && N > 0 && H > 0 && context > 0, ASRB_ERR_BAD_ARG);
    const long long NH = (long long)N * H;
    ASRB_REQUIRE(NH < (1LL << 31) && asrb::ceil_div(T, asrb::kLaChunk) <= 65535, ASRB_ERR_UNSUPPORTED);
    const unsigned chunks = (unsigned)asrb::ceil_div(T, asrb::kLaChunk);
    if (dx) {
        dim3 grid((unsigned)asrb::ceil_div((int)NH, 256), chunks);
        asrb::lookahead_bwd_data_kernel<<<grid, 256, 0, stream>>>(dy, y, w, dx, T, (int)NH, H, context, has_act, lo, hi);
        ASRB_LAUNCH_OK();
    }
    if (dw) {
        ASRB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)H * context * sizeof(float), stream));
        dim3 grid((unsigned)asrb::ceil_div(H, 32), chunks);
        asrb::lookahead_bwd_weight_kernel<<<grid, 256, 0, stream>>>(dy, y, x, dw, T, N, H, context, has_act, lo, hi);
        ASRB_LAUNCH_OK();
    }
    return 0;
}
