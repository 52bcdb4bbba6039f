// asr_b200 -- CTC loss forward + backward (log-space alpha/beta, Graves et al. 2006) with the semantics of
// torch.nn.CTCLoss(blank=0, reduction="sum", zero_infinity=False) as constructed at
// asr_deepspeech/trainers/__main__.py:53 and called at trainers/deepspeech_trainer.py:111:
//   log_probs [T,N,C] fp32 (output of log_softmax), targets int32 concatenated in batch order
//   (asr_deepspeech/functional.py:30-31), input/target lengths int32[N].
//   forward : nll[n] = -log sum_paths ;  backward: grad[t,n,c] = g * (exp(lp) - exp(logsumexp_{s: l'_s = c}(alpha+beta) + nll - lp))
//   (the tensor torch returns for d/dlog_probs, i.e. the gradient w.r.t. the logits), zero for t >= input_len[n].
//
// One CTA per utterance (the DP is sequential in t, parallel in the 2U+1 extended-label positions, which live in
// shared memory).  alpha is spilled to an HBM workspace [N, T, Smax]; the beta kernel walks t backwards, forms the
// per-class posterior in a C-sized shared-memory row (only the <= U+1 classes of this utterance are ever non-zero
// in it) and streams the dense gradient row: read lp once, write grad once -- the 2*T*N*C*4 algorithmic bytes.
#include "common.cuh"

namespace asrb {

constexpr int kCtcThreads = 256;

__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__device__ __forceinline__ int ctc_setup_labels(const int* __restrict__ targets, const int* __restrict__ tgt_len, int n,
                                                int blank, int* ext, int* U_out) {
    // offset of this utterance inside the concatenated target vector
    __shared__ int s_off;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int i = 0; i < n; ++i) off += tgt_len[i];
        s_off = off;
    }
    __syncthreads();
    const int U = tgt_len[n];
    const int S = 2 * U + 1;
    for (int s = threadIdx.x; s < S; s += blockDim.x) ext[s] = (s & 1) ? targets[s_off + (s >> 1)] : blank;
    __syncthreads();
    *U_out = U;
    return S;
}

__global__ void __launch_bounds__(kCtcThreads)
ctc_alpha_kernel(const float* __restrict__ lp, const int* __restrict__ targets, const int* __restrict__ in_len,
                 const int* __restrict__ tgt_len, float* __restrict__ alpha, float* __restrict__ nll, int T, int N, int C,
                 int Smax, int blank) {
    extern __shared__ int smi[];
    int* ext = smi;                                   // [Smax]
    float* a0 = reinterpret_cast<float*>(smi + Smax); // [Smax]
    float* a1 = a0 + Smax;
    const int n = blockIdx.x;
    int U;
    const int S = ctc_setup_labels(targets, tgt_len, n, blank, ext, &U);
    const int Tn = min(in_len[n], T);
    float* al = alpha + (size_t)n * T * Smax;
    if (Tn <= 0) {
        if (threadIdx.x == 0) nll[n] = (U == 0) ? 0.f : INFINITY;
        return;
    }
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float v = (s < 2) ? lp[(size_t)n * C + ext[s]] : -INFINITY;
        a0[s] = v;
        al[s] = v;
    }
    __syncthreads();
    float* prev = a0;
    float* cur = a1;
    for (int t = 1; t < Tn; ++t) {
        const float* row = lp + ((size_t)t * N + n) * C;
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            const int l = ext[s];
            const float x0 = prev[s];
            const float x1 = s >= 1 ? prev[s - 1] : -INFINITY;
            const float x2 = (s >= 2 && l != blank && l != ext[s - 2]) ? prev[s - 2] : -INFINITY;
            const float v = lse3(x0, x1, x2) + row[l];
            cur[s] = v;
            al[(size_t)t * Smax + s] = v;
        }
        __syncthreads();
        float* tmp = prev; prev = cur; cur = tmp;
    }
    if (threadIdx.x == 0) {
        const float l1 = prev[S - 1], l2 = S > 1 ? prev[S - 2] : -INFINITY;
        nll[n] = -lse3(l1, l2, -INFINITY);
    }
}

__global__ void __launch_bounds__(kCtcThreads)
ctc_beta_grad_kernel(const float* __restrict__ lp, const int* __restrict__ targets, const int* __restrict__ in_len,
                     const int* __restrict__ tgt_len, const float* __restrict__ alpha, const float* __restrict__ nll,
                     const float* __restrict__ gscale, float* __restrict__ grad, int T, int N, int C, int Smax, int blank) {
    extern __shared__ int smi[];
    int* ext = smi;                                    // [Smax]
    int* nxt = ext + Smax;                             // [Smax] next position with the same class (or -1)
    int* lead = nxt + Smax;                            // [Smax] 1 if first position of its class
    float* b0 = reinterpret_cast<float*>(lead + Smax); // [Smax]
    float* b1 = b0 + Smax;
    float* ab = b1 + Smax;                             // [Smax] alpha+beta
    float* post = ab + Smax;                           // [C]
    const int n = blockIdx.x;
    int U;
    const int S = ctc_setup_labels(targets, tgt_len, n, blank, ext, &U);
    const int Tn = min(in_len[n], T);
    const float g = gscale ? gscale[0] : 1.f;
    const float nl = nll[n];
    const float* al = alpha + (size_t)n * T * Smax;

    for (int c = threadIdx.x; c < C; c += blockDim.x) post[c] = 0.f;
    // class chains over the label positions (odd s).  Every position whose class is the blank symbol (all even s,
    // plus any label that equals `blank`) is reduced by the whole block instead, see below.
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int l = ext[s];
        int nx = -1, first = 0;
        if ((s & 1) && l != blank) {
            first = 1;
            for (int q = s + 2; q < S; q += 2) if (ext[q] == l) { nx = q; break; }
            for (int q = s - 2; q >= 1; q -= 2) if (ext[q] == l) { first = 0; break; }
        }
        nxt[s] = nx;
        lead[s] = first;
    }
    __shared__ float wpm[kCtcThreads / 32], wps[kCtcThreads / 32];
    __syncthreads();

    float* prev = b0;
    float* cur = b1;
    for (int t = Tn - 1; t >= 0; --t) {
        const float* row = lp + ((size_t)t * N + n) * C;
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            const int l = ext[s];
            float v;
            if (t == Tn - 1) {
                v = (s >= S - 2) ? row[l] : -INFINITY;
            } else {
                const float x0 = prev[s];
                const float x1 = s + 1 < S ? prev[s + 1] : -INFINITY;
                const float x2 = (s + 2 < S && ext[s + 2] != blank && ext[s + 2] != l) ? prev[s + 2] : -INFINITY;
                v = lse3(x0, x1, x2) + row[l];
            }
            cur[s] = v;
            ab[s] = al[(size_t)t * Smax + s] + v;
        }
        __syncthreads();
        // (a) label classes: the first position of each class folds its (short) chain
        // (b) blank class: online (max, sum) over all blank positions, reduced across the block
        float bm = -INFINITY, bs = 0.f;
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            const int l = ext[s];
            if (l == blank) {
                const float v = ab[s];
                if (v > bm) { bs = bs * expf(bm - v) + 1.f; bm = v; }
                else if (v != -INFINITY) bs += expf(v - bm);
            } else if (lead[s]) {
                float m = -INFINITY;
                for (int q = s; q >= 0; q = nxt[q]) m = fmaxf(m, ab[q]);
                float acc = 0.f;
                if (m != -INFINITY)
                    for (int q = s; q >= 0; q = nxt[q]) acc += expf(ab[q] - m);
                const float lcab = (m == -INFINITY) ? -INFINITY : m + logf(acc);
                post[l] = expf(lcab + nl - row[l]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, bm, o);
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const float m = fmaxf(bm, om);
            if (m != -INFINITY) bs = bs * expf(bm - m) + os * expf(om - m);
            bm = m;
        }
        if ((threadIdx.x & 31) == 0) { wpm[threadIdx.x >> 5] = bm; wps[threadIdx.x >> 5] = bs; }
        __syncthreads();
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < kCtcThreads / 32; ++i) m = fmaxf(m, wpm[i]);
        float acc = 0.f;
        if (m != -INFINITY) {
#pragma unroll
            for (int i = 0; i < kCtcThreads / 32; ++i) acc += wps[i] * expf(wpm[i] - m);
        }
        const float post_blank = expf(((m == -INFINITY) ? -INFINITY : m + logf(acc)) + nl - row[blank]);
        float* gr = grad + ((size_t)t * N + n) * C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) gr[c] = g * (expf(row[c]) - (c == blank ? post_blank : post[c]));
        __syncthreads();
        float* tmp = prev; prev = cur; cur = tmp;
    }
    for (int t = max(Tn, 0); t < T; ++t) {
        float* gr = grad + ((size_t)t * N + n) * C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) gr[c] = 0.f;
    }
}

__global__ void ctc_sum_kernel(const float* __restrict__ nll, int N, float* __restrict__ loss) {
    // N <= a few thousand: one warp, fixed order (deterministic)
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += 32) s += nll[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) loss[0] = s;
}

}  // namespace asrb

using namespace asrb;

extern "C" {

size_t asrb_ctc_workspace_bytes(int T, int N, int max_target_len) {
    return (size_t)N * T * (2 * max_target_len + 1) * sizeof(float);
}

/* Forward: nll[N] per utterance and loss[1] = sum_n nll[n]; alpha_ws keeps alpha for the backward. */
int asrb_ctc_fwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, size_t ws_bytes, float* nll, float* loss, int T, int N,
                 int C, int max_target_len, int blank, asrb_stream_t stream) {
    ASRB_REQUIRE(log_probs && targets && input_lengths && target_lengths && alpha_ws && nll && loss, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(T > 0 && N > 0 && C > 0 && max_target_len >= 0 && blank >= 0 && blank < C, ASRB_ERR_BAD_ARG);
    const int Smax = 2 * max_target_len + 1;
    ASRB_REQUIRE(ws_bytes >= asrb_ctc_workspace_bytes(T, N, max_target_len), ASRB_ERR_WORKSPACE);
    const size_t smem = (size_t)Smax * 3 * 4;
    ASRB_REQUIRE(smem <= 200 * 1024, ASRB_ERR_UNSUPPORTED);
    ASRB_CUDA_OK(cudaFuncSetAttribute(ctc_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_alpha_kernel<<<N, kCtcThreads, smem, stream>>>(log_probs, targets, input_lengths, target_lengths, alpha_ws, nll, T, N, C, Smax, blank);
    ASRB_LAUNCH_OK();
    ctc_sum_kernel<<<1, 32, 0, stream>>>(nll, N, loss);
    ASRB_LAUNCH_OK();
    return 0;
}

/* Backward: grad[T,N,C] = grad_scale[0] * d(sum nll)/d(logits); grad_scale is a DEVICE scalar (NULL = 1). */
int asrb_ctc_bwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, const float* alpha_ws, const float* nll, const float* grad_scale,
                 float* grad, int T, int N, int C, int max_target_len, int blank, asrb_stream_t stream) {
    ASRB_REQUIRE(log_probs && targets && input_lengths && target_lengths && alpha_ws && nll && grad, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(T > 0 && N > 0 && C > 0 && max_target_len >= 0 && blank >= 0 && blank < C, ASRB_ERR_BAD_ARG);
    const int Smax = 2 * max_target_len + 1;
    const size_t smem = (size_t)Smax * 6 * 4 + (size_t)C * 4;
    ASRB_REQUIRE(smem <= 200 * 1024, ASRB_ERR_UNSUPPORTED);
    ASRB_CUDA_OK(cudaFuncSetAttribute(ctc_beta_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_beta_grad_kernel<<<N, kCtcThreads, smem, stream>>>(log_probs, targets, input_lengths, target_lengths, alpha_ws, nll, grad_scale, grad, T, N, C, Smax, blank);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
