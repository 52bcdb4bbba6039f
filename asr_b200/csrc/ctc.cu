// asr_b200 -- CTC loss forward + backward (log-space alpha/beta, Graves et al. 2006) with the semantics of
// torch.nn.CTCLoss(blank=0, reduction="sum", zero_infinity=False) as constructed at
// asr_deepspeech/trainers/__main__.py:53 and called at trainers/deepspeech_trainer.py:111:
//   log_probs [T,N,C] fp32 (output of log_softmax), targets int32 concatenated in batch order
//   (asr_deepspeech/functional.py:30-31), input/target lengths int32[N].
//   forward : nll[n] = -log sum_paths ;  backward: grad[t,n,c] = g * (exp(lp) - exp(logsumexp_{s: l'_s = c}(alpha+beta) + nll - lp))
//   (the tensor torch returns for d/dlog_probs, i.e. the gradient w.r.t. the logits), zero for t >= input_len[n].
//
// Two kinds of work with opposite bottlenecks:
//   * the DP over the blank-extended label sequence (2U+1 states) is strictly sequential in t and touches only the
//     2U+1 label columns of each log-prob row: latency-bound.  One CTA per utterance, states in shared memory, and the
//     gathered log-probs (and, for beta, the alpha rows) are PREFETCHED with cp.async into a shared-memory ring several
//     time steps ahead, so no DRAM round trip sits on the step-to-step chain.
//   * the dense gradient exp(lp) - posterior is a pure stream over T*N*C elements: HBM-bound (the 2*T*N*C*4 algorithmic
//     bytes).  It runs in the SAME launch as the beta DP, on extra CTAs that follow the DP's progress through a
//     per-utterance release/acquire flag, scatter that row's <= U+1 class posteriors into a C-sized shared row and
//     stream the row once: read lp once, write grad once.
// alpha is spilled to an HBM workspace [N, T, Smax]; beta overwrites each alpha row with the row's class posteriors.
#include "common.cuh"

namespace asrb {

constexpr int kCtcThreads = 256;
constexpr int kCtcRing = 8;              // prefetch ring depth (time steps)

__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    if (m == -INFINITY) return -INFINITY;
    return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// blank-extended label sequence of utterance n into ext[0..S) ; returns S = 2U+1
__device__ __forceinline__ int ctc_setup_labels(const int* __restrict__ targets, const int* __restrict__ tgt_len, int n,
                                                int blank, int* ext, int* U_out) {
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

// ------------------------------------------------------------------------------------------------
// forward: alpha
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtcThreads)
ctc_alpha_kernel(const float* __restrict__ lp, const int* __restrict__ targets, const int* __restrict__ in_len,
                 const int* __restrict__ tgt_len, float* __restrict__ alpha, float* __restrict__ nll, int T, int N, int C,
                 int Smax, int blank) {
    extern __shared__ int smi[];
    int* ext = smi;                                    // [Smax]
    float* a0 = reinterpret_cast<float*>(smi + Smax);  // [2][Smax]
    float* ring = a0 + 2 * Smax;                       // [kCtcRing][Smax] gathered lp[t, n, ext[s]]
    const int n = blockIdx.x;
    int U;
    const int S = ctc_setup_labels(targets, tgt_len, n, blank, ext, &U);
    const int Tn = min(in_len[n], T);
    float* al = alpha + (size_t)n * T * Smax;
    if (Tn <= 0) {
        if (threadIdx.x == 0) nll[n] = (U == 0) ? 0.f : INFINITY;
        return;
    }
    auto gather = [&](int t) {   // one cp.async group per time step (possibly empty)
        if (t < Tn) {
            const float* row = lp + ((size_t)t * N + n) * C;
            float* dst = ring + (t % kCtcRing) * Smax;
            for (int s = threadIdx.x; s < S; s += kCtcThreads) cp_async4(dst + s, row + ext[s]);
        }
        cp_async_commit();
    };
    for (int t = 0; t < kCtcRing - 1; ++t) gather(t);
    for (int t = 0; t < Tn; ++t) {
        cp_async_wait<kCtcRing - 2>();                 // this thread's copies for step t have landed
        __syncthreads();                               // everybody's copies + previous step's states visible
        const float* g = ring + (t % kCtcRing) * Smax;
        const float* prev = a0 + ((t + 1) & 1) * Smax;
        float* cur = a0 + (t & 1) * Smax;
        for (int s = threadIdx.x; s < S; s += kCtcThreads) {
            float v;
            if (t == 0) {
                v = (s < 2) ? g[s] : -INFINITY;
            } else {
                const int l = ext[s];
                const float x0 = prev[s];
                const float x1 = s >= 1 ? prev[s - 1] : -INFINITY;
                const float x2 = (s >= 2 && l != blank && l != ext[s - 2]) ? prev[s - 2] : -INFINITY;
                v = lse3(x0, x1, x2) + g[s];
            }
            cur[s] = v;
            al[(size_t)t * Smax + s] = v;
        }
        gather(t + kCtcRing - 1);                      // refills the slot consumed in the previous iteration
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float* last = a0 + ((Tn - 1) & 1) * Smax;
        const float l1 = last[S - 1], l2 = S > 1 ? last[S - 2] : -INFINITY;
        nll[n] = -lse3(l1, l2, -INFINITY);
    }
}

// ------------------------------------------------------------------------------------------------
// backward: beta DP CTAs (blockIdx < N) + dense streaming CTAs (blockIdx >= N) in one launch
// ------------------------------------------------------------------------------------------------
struct CtcBwdParams {
    const float* lp; const int* targets; const int* in_len; const int* tgt_len;
    float* alpha;            // in: alpha rows ; out: per-row class posteriors at the first position of each class
    const float* nll; const float* gscale; float* grad;
    int* progress;           // [N] lowest time step whose posteriors are published (starts at T)
    int T, N, C, Smax, blank, n_dp, fused;
};

__device__ void ctc_beta_role(const CtcBwdParams& p, int n, int* smi) {
    const int Smax = p.Smax, N = p.N, C = p.C, T = p.T, blank = p.blank;
    int* ext = smi;                                     // [Smax]
    int* nxt = ext + Smax;                              // [Smax] next position with the same class (or -1)
    int* lead = nxt + Smax;                             // [Smax] 1 if first position of its (non-blank) class
    float* b0 = reinterpret_cast<float*>(lead + Smax);  // [2][Smax]
    float* ab = b0 + 2 * Smax;                          // [Smax] alpha+beta
    float* ring_lp = ab + Smax;                         // [kCtcRing][Smax]
    float* ring_al = ring_lp + kCtcRing * Smax;         // [kCtcRing][Smax]
    __shared__ float wpm[kCtcThreads / 32], wps[kCtcThreads / 32];
    int U;
    const int S = ctc_setup_labels(p.targets, p.tgt_len, n, blank, ext, &U);
    const int Tn = min(p.in_len[n], T);
    const float nl = p.nll[n];
    float* al = p.alpha + (size_t)n * T * Smax;

    // class chains over the label positions (odd s); every position whose class is the blank symbol (all even s,
    // plus any label equal to `blank`) is reduced by the whole block instead
    for (int s = threadIdx.x; s < S; s += kCtcThreads) {
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
    if (Tn <= 0) {
        if (threadIdx.x == 0) st_release_s32(p.progress + n, 0);
        return;
    }
    auto gather = [&](int i) {   // i = iteration index, t = Tn-1-i
        const int t = Tn - 1 - i;
        if (t >= 0) {
            const float* row = p.lp + ((size_t)t * N + n) * C;
            float* d0 = ring_lp + (i % kCtcRing) * Smax;
            float* d1 = ring_al + (i % kCtcRing) * Smax;
            const float* ar = al + (size_t)t * Smax;
            for (int s = threadIdx.x; s < S; s += kCtcThreads) {
                cp_async4(d0 + s, row + ext[s]);
                cp_async4(d1 + s, ar + s);
            }
        }
        cp_async_commit();
    };
    for (int i = 0; i < kCtcRing - 1; ++i) gather(i);
    for (int i = 0; i < Tn; ++i) {
        const int t = Tn - 1 - i;
        cp_async_wait<kCtcRing - 2>();
        __syncthreads();                                           // (A) ring data + previous beta visible; previous row's
        if (threadIdx.x == 0 && i > 0) st_release_s32(p.progress + n, t + 1);  //     posteriors are complete -> publish them
        const float* g = ring_lp + (i % kCtcRing) * Smax;
        const float* a = ring_al + (i % kCtcRing) * Smax;
        const float* prev = b0 + ((i + 1) & 1) * Smax;
        float* cur = b0 + (i & 1) * Smax;
        for (int s = threadIdx.x; s < S; s += kCtcThreads) {
            const int l = ext[s];
            float v;
            if (i == 0) {
                v = (s >= S - 2) ? g[s] : -INFINITY;
            } else {
                const float x0 = prev[s];
                const float x1 = s + 1 < S ? prev[s + 1] : -INFINITY;
                const float x2 = (s + 2 < S && ext[s + 2] != blank && ext[s + 2] != l) ? prev[s + 2] : -INFINITY;
                v = lse3(x0, x1, x2) + g[s];
            }
            cur[s] = v;
            ab[s] = a[s] + v;
        }
        __syncthreads();                                           // (B) alpha+beta complete
        float bm = -INFINITY, bs = 0.f;
        float* out = al + (size_t)t * Smax;                        // the alpha row is dead now: reuse it for the posteriors
        for (int s = threadIdx.x; s < S; s += kCtcThreads) {
            const int l = ext[s];
            float post = 0.f;
            if (l == blank) {
                const float v = ab[s];
                if (v > bm) { bs = bs * __expf(bm - v) + 1.f; bm = v; }
                else if (v != -INFINITY) bs += __expf(v - bm);
            } else if (lead[s]) {
                float m = -INFINITY;
                for (int q = s; q >= 0; q = nxt[q]) m = fmaxf(m, ab[q]);
                float acc = 0.f;
                if (m != -INFINITY)
                    for (int q = s; q >= 0; q = nxt[q]) acc += __expf(ab[q] - m);
                const float lcab = (m == -INFINITY) ? -INFINITY : m + __logf(acc);
                post = __expf(lcab + nl - g[s]);
            }
            if (s > 0) out[s] = post;                              // position 0 (blank class) is written below
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, bm, o);
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const float m = fmaxf(bm, om);
            if (m != -INFINITY) bs = bs * __expf(bm - m) + os * __expf(om - m);
            bm = m;
        }
        if ((threadIdx.x & 31) == 0) { wpm[threadIdx.x >> 5] = bm; wps[threadIdx.x >> 5] = bs; }
        __syncthreads();                                           // (C) blank partials complete
        if (threadIdx.x == 0) {
            float m = -INFINITY;
#pragma unroll
            for (int w = 0; w < kCtcThreads / 32; ++w) m = fmaxf(m, wpm[w]);
            float acc = 0.f;
            if (m != -INFINITY) {
#pragma unroll
                for (int w = 0; w < kCtcThreads / 32; ++w) acc += wps[w] * __expf(wpm[w] - m);
            }
            out[0] = __expf(((m == -INFINITY) ? -INFINITY : m + __logf(acc)) + nl - g[0]);   // ext[0] == blank
        }
        gather(i + kCtcRing - 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_s32(p.progress + n, 0);
}

__device__ void ctc_dense_role(const CtcBwdParams& p, int j, int M, int* smi) {
    const int Smax = p.Smax, N = p.N, C = p.C, T = p.T, blank = p.blank;
    float* sm = reinterpret_cast<float*>(smi);           // [C] class posteriors of the current row (zero elsewhere)
    int* ext = smi + C;                                  // [Smax] labels of the current utterance
    float* pr = reinterpret_cast<float*>(ext + Smax);    // [Smax] posteriors of the current row
    __shared__ int s_off;
    const float g = p.gscale ? p.gscale[0] : 1.f;
    for (int c = threadIdx.x; c < C; c += kCtcThreads) sm[c] = 0.f;
    const bool vec = (C % 4 == 0);
    int cur_n = -1, S = 0;
    const long long rows = (long long)T * N;
    for (long long r = j; r < rows; r += M) {
        const int t = T - 1 - (int)(r / N), n = (int)(r % N);
        const int Tn = min(p.in_len[n], T);
        float* gr = p.grad + ((size_t)t * N + n) * C;
        if (t >= Tn) {                                    // beyond the utterance: zero gradient
            if (vec) for (int c = threadIdx.x; c < C / 4; c += kCtcThreads) reinterpret_cast<float4*>(gr)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            else     for (int c = threadIdx.x; c < C; c += kCtcThreads) gr[c] = 0.f;
            continue;
        }
        if (n != cur_n) {                                 // labels of this utterance (uniform branch)
            __syncthreads();
            if (threadIdx.x == 0) {
                int off = 0;
                for (int i = 0; i < n; ++i) off += p.tgt_len[i];
                s_off = off;
            }
            __syncthreads();
            S = 2 * p.tgt_len[n] + 1;
            for (int s = threadIdx.x; s < S; s += kCtcThreads) ext[s] = (s & 1) ? p.targets[s_off + (s >> 1)] : blank;
            cur_n = n;
        }
        if (p.fused && threadIdx.x == 0) {
            while (ld_acquire_s32(p.progress + n) > t) {
            }
        }
        __syncthreads();
        const float* post = p.alpha + ((size_t)n * T + t) * Smax;
        for (int s = threadIdx.x; s < S; s += kCtcThreads) {
            const float v = __ldcg(post + s);             // written by another CTA during this launch: bypass L1
            pr[s] = v;
            if (v != 0.f) sm[ext[s]] = v;                 // one non-zero position per class
        }
        __syncthreads();
        const float* row = p.lp + ((size_t)t * N + n) * C;
        if (vec) {
            for (int c = threadIdx.x; c < C / 4; c += kCtcThreads) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(row) + c);
                const float4 q = reinterpret_cast<const float4*>(sm)[c];
                float4 o;
                o.x = g * (__expf(x.x) - q.x); o.y = g * (__expf(x.y) - q.y);
                o.z = g * (__expf(x.z) - q.z); o.w = g * (__expf(x.w) - q.w);
                reinterpret_cast<float4*>(gr)[c] = o;
            }
        } else {
            for (int c = threadIdx.x; c < C; c += kCtcThreads) gr[c] = g * (__expf(__ldg(row + c)) - sm[c]);
        }
        __syncthreads();
        for (int s = threadIdx.x; s < S; s += kCtcThreads)
            if (pr[s] != 0.f) sm[ext[s]] = 0.f;           // restore the all-zero row
    }
}

__global__ void __launch_bounds__(kCtcThreads)
ctc_bwd_kernel(const CtcBwdParams p) {
    extern __shared__ int smi[];
    if ((int)blockIdx.x < p.n_dp) ctc_beta_role(p, blockIdx.x, smi);
    else ctc_dense_role(p, blockIdx.x - p.n_dp, gridDim.x - p.n_dp, smi);
}

__global__ void ctc_fill_progress_kernel(int* progress, int N, int T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) progress[i] = T;
}

__global__ void ctc_sum_kernel(const float* __restrict__ nll, int N, float* __restrict__ loss) {
    // N <= a few thousand: one warp, fixed order (deterministic)
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += 32) s += nll[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) loss[0] = s;
}

static size_t ctc_alpha_floats(int T, int N, int max_target_len) { return (size_t)N * T * (2 * max_target_len + 1); }

}  // namespace asrb

using namespace asrb;

extern "C" {

size_t asrb_ctc_workspace_bytes(int T, int N, int max_target_len) {
    return ctc_alpha_floats(T, N, max_target_len) * sizeof(float) + (size_t)N * sizeof(int) + 64;
}

/* Forward: nll[N] per utterance and loss[1] = sum_n nll[n]; ws keeps alpha for the backward. */
int asrb_ctc_fwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, size_t ws_bytes, float* nll, float* loss, int T, int N,
                 int C, int max_target_len, int blank, asrb_stream_t stream) {
    ASRB_REQUIRE(log_probs && targets && input_lengths && target_lengths && alpha_ws && nll && loss, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(T > 0 && N > 0 && C > 0 && max_target_len >= 0 && blank >= 0 && blank < C, ASRB_ERR_BAD_ARG);
    const int Smax = 2 * max_target_len + 1;
    ASRB_REQUIRE(ws_bytes >= asrb_ctc_workspace_bytes(T, N, max_target_len), ASRB_ERR_WORKSPACE);
    const size_t smem = (size_t)Smax * (3 + kCtcRing) * 4;
    ASRB_REQUIRE(smem <= 200 * 1024, ASRB_ERR_UNSUPPORTED);
    ASRB_CUDA_OK(cudaFuncSetAttribute(ctc_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_alpha_kernel<<<N, kCtcThreads, smem, stream>>>(log_probs, targets, input_lengths, target_lengths, alpha_ws, nll, T, N, C, Smax, blank);
    ASRB_LAUNCH_OK();
    ctc_sum_kernel<<<1, 32, 0, stream>>>(nll, N, loss);
    ASRB_LAUNCH_OK();
    return 0;
}

/* Backward: grad[T,N,C] = grad_scale[0] * d(sum nll)/d(logits); grad_scale is a DEVICE scalar (NULL = 1).
 * alpha_ws is the workspace asrb_ctc_fwd filled (it is consumed: a second backward needs a new forward). */
int asrb_ctc_bwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, const float* nll, const float* grad_scale,
                 float* grad, int T, int N, int C, int max_target_len, int blank, asrb_stream_t stream) {
    ASRB_REQUIRE(log_probs && targets && input_lengths && target_lengths && alpha_ws && nll && grad, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(T > 0 && N > 0 && C > 0 && max_target_len >= 0 && blank >= 0 && blank < C, ASRB_ERR_BAD_ARG);
    const int Smax = 2 * max_target_len + 1;
    const size_t smem_dp = (size_t)Smax * (6 + 2 * kCtcRing) * 4;
    const size_t smem_dense = (size_t)C * 4 + (size_t)Smax * 8;
    const size_t smem = smem_dp > smem_dense ? smem_dp : smem_dense;
    ASRB_REQUIRE(smem <= 100 * 1024, ASRB_ERR_UNSUPPORTED);
    int* progress = reinterpret_cast<int*>(alpha_ws + ctc_alpha_floats(T, N, max_target_len));
    ctc_fill_progress_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(progress, N, T);
    ASRB_LAUNCH_OK();
    ASRB_CUDA_OK(cudaFuncSetAttribute(ctc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CtcBwdParams p = {log_probs, targets, input_lengths, target_lengths, alpha_ws, nll, grad_scale, grad, progress,
                      T, N, C, Smax, blank, N, 1};
    // Fused launch: the DP CTAs have the lowest block indices, so they are all resident before any streaming CTA
    // starts waiting on them -- provided they fit in one wave.  Otherwise run the two roles back to back.
    int per_sm = 0;
    ASRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctc_bwd_kernel, kCtcThreads, smem));
    const long long rows = (long long)T * N;
    int dense = kNumSMs * (per_sm > 2 ? 2 : (per_sm > 0 ? per_sm : 1));
    if (dense > rows) dense = (int)rows;
    if (per_sm > 0 && N + dense <= kNumSMs * per_sm) {
        ctc_bwd_kernel<<<N + dense, kCtcThreads, smem, stream>>>(p);
        ASRB_LAUNCH_OK();
    } else {
        p.fused = 0;
        p.n_dp = N;
        ctc_bwd_kernel<<<N, kCtcThreads, smem, stream>>>(p);           // beta DP only
        ASRB_LAUNCH_OK();
        p.n_dp = 0;
        ctc_bwd_kernel<<<dense, kCtcThreads, smem, stream>>>(p);       // dense stream only
        ASRB_LAUNCH_OK();
    }
    return 0;
}

}  // extern "C"
