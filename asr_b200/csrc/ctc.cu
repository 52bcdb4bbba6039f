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
//     2U+1 label columns of each log-prob row: latency-bound.  ONE WARP per utterance: lane i keeps KPL consecutive
//     states in registers (2U+1 <= 32*KPL <= 1024), the two neighbour states of the previous lane arrive by shuffle, so a
//     time step has no block barrier at all; the gathered log-probs (and, for beta, the alpha rows) are PREFETCHED with
//     cp.async into a per-warp shared-memory ring several time steps ahead, so no DRAM round trip sits on the
//     step-to-step chain either.  (Longer label sequences fall back to one CTA per utterance with the states in
//     shared memory.)
//   * the dense gradient exp(lp) - posterior is a pure stream over T*N*C elements: HBM-bound (the 2*T*N*C*4 algorithmic
//     bytes).  Only the <= U+1 classes that occur in an utterance's labels have a non-zero posterior, so the stream
//     writes g*exp(lp) for every OTHER class (a C-bit class mask per utterance) and depends on nothing, while the DP
//     writes the gradient of the label classes itself.  Both run in the SAME launch (DP blocks first, streaming
//     blocks after) without any flag between them: the stream runs at HBM speed underneath the latency-bound DP.
// alpha is spilled to an HBM workspace [N, T, row]; the forward also keeps the gathered log-probs lp[t,n,l'_s] there
// so that the backward DP reads two contiguous rows per step instead of gathering 4-byte elements from DRAM again.
#include "ptx.cuh"

namespace asrb {

constexpr int kCtcThreads = 256;
int g_ctc_dbg = 0;   // DEBUG timing: 1 = no dense stream, 2 = no DP (progress pre-set to 0)
int g_ctc_min_smem = 0, g_ctc_blocks_per_sm = 5;   // tuning (asrb_debug_ctc_tuning)
constexpr int kCtcRing = 8;              // prefetch ring depth (time steps)

// branch-free (a divergent early-out per state keeps the compiler from interleaving the independent states of a lane):
// all operands -inf -> 0 + log(0) = -inf
__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    const float ms = (m == -INFINITY) ? 0.f : m;
    return ms + __logf(__expf(a - ms) + __expf(b - ms) + __expf(c - ms));
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
__device__ __forceinline__ int ctc_setup_labels(const int* __restrict__ targets, const int* __restrict__ tgt_off,
                                                const int* __restrict__ tgt_len, int n, int blank, int* ext, int* U_out) {
    const int U = tgt_len[n], off = tgt_off[n];
    const int S = 2 * U + 1;
    for (int s = threadIdx.x; s < S; s += blockDim.x) ext[s] = (s & 1) ? targets[off + (s >> 1)] : blank;
    __syncthreads();
    *U_out = U;
    return S;
}

// tgt_off[n] = sum of the target lengths before utterance n
__global__ void ctc_prepare_kernel(const int* __restrict__ tgt_len, int* __restrict__ tgt_off, int N) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 32) {           // one warp, inclusive scan by shuffles
        const int i = base + threadIdx.x;
        const int v = i < N ? tgt_len[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)threadIdx.x >= o) x += y;
        }
        if (i < N) tgt_off[i] = carry + x - v;
        __syncwarp();
        if (threadIdx.x == 31) carry += x;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// warp-per-utterance DP
// ------------------------------------------------------------------------------------------------
// Row layout of the alpha workspace for KPL states per lane: state s = lane*KPL + k lives at column k*32 + lane, so
// that every one of the KPL loads/stores of a warp covers 128 contiguous bytes.
__host__ __device__ inline int ctc_col(int s, int kpl) { return kpl ? (s % kpl) * 32 + s / kpl : s; }

// The warp DP uses the flush-to-zero MUFU forms directly: __expf/__logf spend ~10 extra FMUL/FSETP per state on
// denormal scaling (measured: 655 instructions per step at 13 states per lane).  Values stay in natural-log units --
// log2 units would save three more multiplies per state, but adding lp*log2(e) every step rounds differently from
// the reference's lp additions and the gradients drift away from torch's (measured 3e-5 vs 1e-6 on the golden cases).
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float exp_ftz(float x) { return ex2_ftz(x * kLog2e); }
// log(e^a + e^b + e^c), branch-free; all operands -inf -> fma(log2(0), ln2, -1e30) = -inf
__device__ __forceinline__ float lse3_2(float a, float b, float c) {
    const float ms = fmaxf(fmaxf(fmaxf(a, b), c), -1e30f);
    return fmaf(lg2_ftz(exp_ftz(a - ms) + exp_ftz(b - ms) + exp_ftz(c - ms)), kLn2, ms);
}

template <int KPL> struct CtcRing { static constexpr int R = KPL <= 8 ? 16 : (KPL <= 16 ? 8 : 4); };

__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b);
    const float ms = (m == -INFINITY) ? 0.f : m;
    return ms + __logf(__expf(a - ms) + __expf(b - ms));
}

template <int KPL>
__global__ void __launch_bounds__(32)
ctc_alpha_warp_kernel(const float* __restrict__ lp, const int* __restrict__ targets, const int* __restrict__ tgt_off,
                      const int* __restrict__ in_len, const int* __restrict__ tgt_len, float* __restrict__ alpha,
                      float* __restrict__ gathered, float* __restrict__ nll, int T, int N, int C, int blank) {
    constexpr int R = CtcRing<KPL>::R, RS = KPL * 32;
    __shared__ float ring[R][RS];
    const int n = blockIdx.x, lane = threadIdx.x;
    const int U = tgt_len[n], S = 2 * U + 1, off = tgt_off[n];
    const int Tn = min(in_len[n], T);
    if (Tn <= 0) {
        if (lane == 0) nll[n] = (U == 0) ? 0.f : INFINITY;
        return;
    }
    // States beyond S are dummies: they read lp[blank], their values never reach a real state (alpha only flows
    // upwards) and they land in the padding columns of the workspace rows -- so the loop carries no s < S predicate.
    int ext[KPL];
    float pen[KPL];                                  // 0 if the s-2 -> s transition exists, else -inf
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
        const int s = lane * KPL + k;
        int l = blank;
        bool sk = false;
        if (s < S && (s & 1)) {
            l = targets[off + (s >> 1)];
            sk = s >= 3 && l != blank && l != targets[off + (s >> 1) - 1];
        }
        ext[k] = l;
        pen[k] = sk ? 0.f : -INFINITY;
    }
    float* al = alpha + (size_t)n * T * RS + lane;
    float* gw = gathered + (size_t)n * T * RS + lane;
    const float* row_next = lp + (size_t)n * C;      // row of the next time step to prefetch
    const size_t row_stride = (size_t)N * C;
    auto gather = [&](int t) {   // one cp.async group per time step (possibly empty); every lane fetches its own states
        if (t < Tn) {
#pragma unroll
            for (int k = 0; k < KPL; ++k) cp_async4(&ring[t % R][k * 32 + lane], row_next + ext[k]);
            row_next += row_stride;
        }
        cp_async_commit();
    };
    for (int t = 0; t < R - 1; ++t) gather(t);
    float prev[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) prev[k] = -INFINITY;
    for (int t = 0; t < Tn; ++t) {
        cp_async_wait<R - 2>();                        // this lane's copies for step t have landed
        float g[KPL];
#pragma unroll
        for (int k = 0; k < KPL; ++k) g[k] = ring[t % R][k * 32 + lane];
        // states s-1 and s-2 of this lane's first states live in the previous lane
        float p1 = __shfl_up_sync(0xffffffffu, prev[KPL - 1], 1);
        float p2 = KPL >= 2 ? __shfl_up_sync(0xffffffffu, prev[KPL >= 2 ? KPL - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, prev[0], 2);
        if (lane == 0) p1 = -INFINITY;
        if (lane < (KPL >= 2 ? 1 : 2)) p2 = -INFINITY;
        float cur[KPL];
        if (t == 0) {                                  // uniform branch
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const int s = lane * KPL + k;
                cur[k] = (s < 2 && s < S) ? g[k] : -INFINITY;
            }
        } else {
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const float x0 = prev[k];
                const float x1 = k >= 1 ? prev[k >= 1 ? k - 1 : 0] : p1;
                const float x2 = (k >= 2 ? prev[k >= 2 ? k - 2 : 0] : (k == 1 ? p1 : p2)) + pen[k];
                cur[k] = lse3_2(x0, x1, x2) + g[k];
            }
        }
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            al[k * 32] = cur[k];
            gw[k * 32] = g[k];
            prev[k] = cur[k];
        }
        al += RS;
        gw += RS;
        gather(t + R - 1);                             // refills the slot consumed in the previous iteration
    }
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KPL; ++k) ring[0][k * 32 + lane] = prev[k];
    __syncwarp();
    if (lane == 0) {
        const float l1 = ring[0][ctc_col(S - 1, KPL)], l2 = S > 1 ? ring[0][ctc_col(S - 2, KPL)] : -INFINITY;
        nll[n] = -lse3_2(l1, l2, -INFINITY);
    }
}

struct CtcBwdParams {
    const float* lp; const int* targets; const int* tgt_off; const int* in_len; const int* tgt_len;
    const float* alpha;      // [N, T, RS] alpha rows of the forward
    float* gathered;         // [N, T, RS] lp[t, n, l'_s] rows of the forward (warp DP only); the backward overwrites
                             // a consumed row with the row's class posteriors (at the column of each class's first state)
    int* progress;           // [N] lowest time step whose posterior row is published (starts at T)
    const float* nll; const float* gscale; float* grad;
    int T, N, C, Smax, RS, kpl, blank, n_dp, nparts, dbg;
};

constexpr int kDenseDepth = 8;         // 16-byte loads in flight per lane of a streaming warp (16 would cost a block per SM in registers)
constexpr int kCtcPublishEvery = 16;   // rows between two progress publications (each is a MEMBAR.GPU)
constexpr int kCtcXRing = 4;           // rows of alpha+beta in flight between the recurrence warp and the posterior warp
constexpr int kCtcBetaRing = 4;        // prefetch depth of the beta warp (contiguous rows: three steps ahead is plenty)

// shared memory of a DP block (floats): cp.async rings of lp and alpha, the hand-over ring, two scratch rows, labels
template <int KPL>
__host__ __device__ constexpr size_t ctc_dp_smem_bytes() {
    return (size_t)(2 * kCtcBetaRing + kCtcXRing + 3) * KPL * 32 * 4 + 2 * kCtcXRing * 8 + 64;
}

// Backward DP of utterance n by TWO warps of a DP block:
//   warp 0 runs the beta recurrence (the only sequential chain) and hands row after row of
//          x[s] = alpha_t(s) + beta_t(s) - lp_t(l'_s) + nll  to warp 1 through a shared-memory ring;
//   warp 1 turns a row into class posteriors, post(c) = sum_{s: l'_s = c} exp(x[s]), publishes them as a compact row
//          and every 16 rows bumps the utterance's progress word, which the streaming warps follow.  These terms are probabilities
//          (they sum to 1 over the whole row), so the class sums are taken in the linear domain, in a fixed order:
//          the blank class by a warp reduction, the label classes as differences of ONE prefix sum over the label
//          positions sorted by class (no pointer chasing along per-class chains, no atomics).
__device__ __forceinline__ void named_bar_sync_ctc(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int KPL>
__device__ void ctc_beta_warp_role(const CtcBwdParams& p, int n, int* smi) {
    constexpr int R = kCtcBetaRing, RS = KPL * 32, XR = kCtcXRing;
    constexpr int KP2 = (KPL + 1) / 2;                  // sorted label slots per lane (32*KP2 >= U)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, blank = p.blank;
    float* ring_lp = reinterpret_cast<float*>(smi);     // [R][RS]
    float* ring_al = ring_lp + R * RS;                  // [R][RS]
    float* xr = ring_al + R * RS;                       // [XR][RS] alpha+beta-lp+nll
    float* se = xr + XR * RS;                           // [RS] exp(x) by column
    float* sP = se + RS;                                // [RS] prefix sums over the sorted label slots
    int* sext = reinterpret_cast<int*>(sP + RS);        // [RS] labels by column (setup only)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sext + RS);   // [XR]
    uint64_t* empty_bar = full_bar + XR;                            // [XR]
    const int U = p.tgt_len[n], S = 2 * U + 1, off = p.tgt_off[n];
    const int Tn = min(p.in_len[n], T);
    if (threadIdx.x == 0) {
        for (int i = 0; i < XR; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_fence_init();
    }
    named_bar_sync_ctc(1, 64);                           // the two DP warps only
    if (Tn <= 0) return;
    const float* al = p.alpha + (size_t)n * T * RS;
    const float* gw = p.gathered + (size_t)n * T * RS;   // read 3 steps ahead of the row being handed over
    int ext[KPL];
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
        const int s = lane * KPL + k;
        ext[k] = (s < S && (s & 1)) ? p.targets[off + (s >> 1)] : blank;
    }

    if (warp == 0) {
        // ============================== beta recurrence ==============================
        const float nl = p.nll[n];
        float pen[KPL];                                  // 0 if the s -> s+2 transition exists, else -inf
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            const int s = lane * KPL + k;
            bool sk = false;
            if (s < S && (s & 1) && s + 2 < S) {
                const int l2 = p.targets[off + (s >> 1) + 1];
                sk = l2 != blank && l2 != ext[k];
            }
            pen[k] = sk ? 0.f : -INFINITY;
        }
        // Dummy states (s >= S) stay at -inf by themselves: beta only flows downwards from S-1, S-2.
        const float* gr_next = gw + (size_t)(Tn - 1) * RS + lane;
        const float* ar_next = al + (size_t)(Tn - 1) * RS + lane;
        auto gather = [&](int i) {   // i = iteration index, t = Tn-1-i
            if (i < Tn) {
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    cp_async4(ring_lp + (i % R) * RS + k * 32 + lane, gr_next + k * 32);
                    cp_async4(ring_al + (i % R) * RS + k * 32 + lane, ar_next + k * 32);
                }
                gr_next -= RS;
                ar_next -= RS;
            }
            cp_async_commit();
        };
        for (int i = 0; i < R - 1; ++i) gather(i);
        float prev[KPL];
#pragma unroll
        for (int k = 0; k < KPL; ++k) prev[k] = -INFINITY;
        for (int i = 0; i < Tn; ++i) {
            cp_async_wait<R - 2>();
            float g[KPL], a[KPL];
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                g[k] = ring_lp[(i % R) * RS + k * 32 + lane];
                a[k] = ring_al[(i % R) * RS + k * 32 + lane];
            }
            // states s+1 and s+2 of this lane's last states live in the next lane
            float n1 = __shfl_down_sync(0xffffffffu, prev[0], 1);
            float n2 = KPL >= 2 ? __shfl_down_sync(0xffffffffu, prev[KPL >= 2 ? 1 : 0], 1) : __shfl_down_sync(0xffffffffu, prev[0], 2);
            if (lane == 31) n1 = -INFINITY;
            if (lane > (KPL >= 2 ? 30 : 29)) n2 = -INFINITY;
            float cur[KPL];
            if (i == 0) {                                  // uniform branch
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const int s = lane * KPL + k;
                    cur[k] = (s >= S - 2 && s < S) ? g[k] : -INFINITY;
                }
            } else {
#pragma unroll
                for (int k = 0; k < KPL; ++k) {
                    const float x0 = prev[k];
                    const float x1 = k + 1 < KPL ? prev[k + 1 < KPL ? k + 1 : 0] : n1;
                    const float x2 = (k + 2 < KPL ? prev[k + 2 < KPL ? k + 2 : 0] : (k + 2 == KPL ? n1 : n2)) + pen[k];
                    cur[k] = lse3_2(x0, x1, x2) + g[k];
                }
            }
            // hand the row over: x = alpha + beta - lp + nll
            const int slot = i % XR;
            mbar_wait(&empty_bar[slot], ((uint32_t)(i / XR) & 1u) ^ 1u);
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                xr[slot * RS + k * 32 + lane] = (a[k] + cur[k]) - g[k] + nl;
                prev[k] = cur[k];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[slot]);
            gather(i + R - 1);
        }
    } else {
        // ============================== class posteriors ==============================
        bool isblank[KPL], lead[KPL];
        int seg_a[KPL], seg_b[KPL];                      // sorted-slot range of the class a lead position stands for
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            const int s = lane * KPL + k;
            isblank[k] = s < S && ext[k] == blank;
            sext[k * 32 + lane] = s < S ? ext[k] : -1;
            lead[k] = false;
            seg_a[k] = seg_b[k] = 0;
        }
        __syncwarp();
        int* sperm = reinterpret_cast<int*>(se);          // setup scratch: sorted slot -> column
        for (int i = lane; i < RS; i += 32) sperm[i] = -1;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            const int s = lane * KPL + k;
            if (s < S && (s & 1)) {
                const int l = ext[k], j = s >> 1;
                int less = 0, same_before = 0, same = 0;
                for (int i = 0; i < U; ++i) {
                    const int li = sext[ctc_col(2 * i + 1, KPL)];
                    less += li < l;
                    same += li == l;
                    same_before += (li == l) && (i < j);
                }
                sperm[less + same_before] = k * 32 + lane;
                if (l != blank && same_before == 0) {
                    lead[k] = true;
                    seg_a[k] = less;
                    seg_b[k] = less + same - 1;
                }
            }
        }
        __syncwarp();
        int pcol[KP2];
#pragma unroll
        for (int q = 0; q < KP2; ++q) {
            const int slot = lane * KP2 + q;
            pcol[q] = slot < U ? sperm[slot] : -1;
        }
        __syncwarp();
        for (int i = 0; i < Tn; ++i) {
            const int t = Tn - 1 - i;
            const int slot = i % XR;
            mbar_wait(&full_bar[slot], (uint32_t)(i / XR) & 1u);
            float e[KPL];
#pragma unroll
            for (int k = 0; k < KPL; ++k) e[k] = (lane * KPL + k < S) ? exp_ftz(xr[slot * RS + k * 32 + lane]) : 0.f;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            float be = 0.f;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                if (isblank[k]) be += e[k];
                se[k * 32 + lane] = e[k];
            }
            be = warp_sum(be);
            __syncwarp();
            float v[KP2], run = 0.f;
#pragma unroll
            for (int q = 0; q < KP2; ++q) {
                run += pcol[q] >= 0 ? se[pcol[q]] : 0.f;
                v[q] = run;
            }
            float incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const float excl = incl - run;
#pragma unroll
            for (int q = 0; q < KP2; ++q) sP[lane * KP2 + q] = v[q] + excl;
            __syncwarp();
            // posterior row, by state column: the class's first state carries the class posterior, state 0 the blank's.
            // It replaces the gathered row t, which the recurrence warp has already consumed.
            float* out = p.gathered + ((size_t)n * T + t) * RS + lane;
#pragma unroll
            for (int k = 0; k < KPL; ++k) {
                const int s = lane * KPL + k;
                float post = 0.f;
                if (lead[k]) post = sP[seg_b[k]] - (seg_a[k] > 0 ? sP[seg_a[k] - 1] : 0.f);
                if (s == 0) post = be;
                out[k * 32] = post;
            }
            if ((i % kCtcPublishEvery) == kCtcPublishEvery - 1 || i == Tn - 1) {
                __syncwarp();
                if (lane == 0) st_release_s32(p.progress + n, i == Tn - 1 ? 0 : t);   // rows >= t are complete
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward: alpha
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtcThreads)
ctc_alpha_kernel(const float* __restrict__ lp, const int* __restrict__ targets, const int* __restrict__ tgt_off,
                 const int* __restrict__ in_len, const int* __restrict__ tgt_len, float* __restrict__ alpha,
                 float* __restrict__ nll, int T, int N, int C, int Smax, int blank) {
    extern __shared__ int smi[];
    int* ext = smi;                                    // [Smax]
    float* a0 = reinterpret_cast<float*>(smi + Smax);  // [2][Smax]
    float* ring = a0 + 2 * Smax;                       // [kCtcRing][Smax] gathered lp[t, n, ext[s]]
    const int n = blockIdx.x;
    int U;
    const int S = ctc_setup_labels(targets, tgt_off, tgt_len, n, blank, ext, &U);
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
    const int S = ctc_setup_labels(p.targets, p.tgt_off, p.tgt_len, n, blank, ext, &U);
    const int Tn = min(p.in_len[n], T);
    const float nl = p.nll[n];
    const float* al = p.alpha + (size_t)n * T * p.RS;
    const float gs = p.gscale ? p.gscale[0] : 1.f;

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
    if (Tn <= 0) return;
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
        __syncthreads();                                           // (A) ring data + previous beta visible
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
        float* grow = p.grad + ((size_t)t * N + n) * C;            // the label classes' gradients are written here
        for (int s = threadIdx.x; s < S; s += kCtcThreads) {
            const int l = ext[s];
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
                grow[l] = gs * (__expf(g[s]) - __expf(lcab + nl - g[s]));
            }
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
            grow[blank] = gs * (__expf(g[0]) - __expf(((m == -INFINITY) ? -INFINITY : m + __logf(acc)) + nl - g[0]));   // ext[0] == blank
        }
        gather(i + kCtcRing - 1);
    }
}

// Dense gradient stream.  Every block works for ONE utterance n = block % N; the utterance's streaming warps -- warps
// 2..7 of its DP block plus all 8 warps of its nparts-1 further blocks -- deal the rows out warp by warp, latest time
// step first (the order in which the DP publishes posteriors): no block barrier, eight 16-byte loads in flight per lane.
//   POST = true  (warp DP): full rows grad = g*(exp(lp) - post).  Only the <= U+1 classes of the utterance's labels have a
//           posterior: a C-bit mask marks them, first_state[c] says where in the DP's compact posterior row the value is,
//           and the warp stages that row (RS floats) in shared memory once it is published.  Every store is a whole
//           16-byte vector: no partial-sector read-modify-write in DRAM (the version in which the DP scattered the label
//           classes' gradients itself moved 30.7 GB instead of 22 GB at the isolation shape).
//   POST = false (block-wide DP, long label sequences): the DP writes the label classes itself, the stream skips them.
template <bool POST>
__device__ void ctc_dense_role(const CtcBwdParams& p, int n, int wi, int nw, const uint32_t* mask, const int* first_state,
                               float* wpost) {
    const int N = p.N, C = p.C, T = p.T, RS = p.RS;
    const int Tn = min(p.in_len[n], T);
    const float g = p.gscale ? p.gscale[0] : 1.f;
    const int lane = threadIdx.x & 31;
    const bool vec = C % 4 == 0;
    const int C4 = C / 4;
    // rows beyond the utterance: zeros
    for (int t = Tn + wi; t < T; t += nw) {
        float* gr = p.grad + ((size_t)t * N + n) * C;
        if (vec) for (int c = lane; c < C4; c += 32) __stcs(reinterpret_cast<float4*>(gr) + c, make_float4(0.f, 0.f, 0.f, 0.f));
        else     for (int c = lane; c < C; c += 32) gr[c] = 0.f;
    }
    int seen = T;                                        // last progress value this warp has observed
    for (int t = Tn - 1 - wi; t >= 0; t -= nw) {
        if constexpr (POST) {
            if (seen > t) {
                if (lane == 0) {
                    do { seen = ld_acquire_s32(p.progress + n); } while (seen > t);
                }
                seen = __shfl_sync(0xffffffffu, seen, 0);
            }
            const float* post = p.gathered + ((size_t)n * T + t) * RS;
            __syncwarp();                                // the previous row's readers are done with wpost
            for (int i = lane; i < RS; i += 32) wpost[i] = __ldcg(post + i);   // written during this launch: bypass L1
            __syncwarp();
        }
        const float* row = p.lp + ((size_t)t * N + n) * C;
        float* gr = p.grad + ((size_t)t * N + n) * C;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            float4* gr4 = reinterpret_cast<float4*>(gr);
            // kDenseDepth 16-byte loads in flight per lane: the kernel's register count (set by the DP warps) allows
            // only two blocks per SM, so each streaming warp has to cover >= 8 KB of the HBM latency by itself
            for (int c0 = lane; c0 < C4; c0 += 32 * kDenseDepth) {
                float4 x[kDenseDepth];
#pragma unroll
                for (int u = 0; u < kDenseDepth; ++u)
                    if (c0 + 32 * u < C4) x[u] = __ldcs(row4 + c0 + 32 * u);
#pragma unroll
                for (int u = 0; u < kDenseDepth; ++u) {
                    const int c = c0 + 32 * u;
                    if (c < C4) {
                        float o[4] = {exp_ftz(x[u].x), exp_ftz(x[u].y), exp_ftz(x[u].z), exp_ftz(x[u].w)};
                        const uint32_t bits = (mask[(4 * c) >> 5] >> ((4 * c) & 31)) & 0xFu;
                        if (bits == 0u) {
                            __stcs(gr4 + c, make_float4(g * o[0], g * o[1], g * o[2], g * o[3]));
                        } else if constexpr (POST) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (bits & (1u << e)) o[e] -= wpost[-first_state[4 * c + e] - 1];   // two LDS: the slow path is divergent
                            __stcs(gr4 + c, make_float4(g * o[0], g * o[1], g * o[2], g * o[3]));
                        } else {
                            float* gs = reinterpret_cast<float*>(gr4 + c);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (!(bits & (1u << e))) gs[e] = g * o[e];
                        }
                    }
                }
            }
        } else {
            for (int c = lane; c < C; c += 32) {
                const bool m = (mask[c >> 5] >> (c & 31)) & 1u;
                if (!m) gr[c] = g * exp_ftz(__ldg(row + c));
                else if constexpr (POST) gr[c] = g * (exp_ftz(__ldg(row + c)) - wpost[-first_state[c] - 1]);
            }
        }
    }
}

// bit c of mask set: class c occurs in utterance n's labels (or is the blank); first_state[c] (POST only) = column, in
// the DP's posterior row, of the first state of the extended label sequence that carries class c (0 for the blank)
__device__ void ctc_build_mask(const CtcBwdParams& p, int n, uint32_t* mask, int* first_state) {
    const int words = (p.C + 31) / 32;
    for (int w = threadIdx.x; w < words; w += kCtcThreads) mask[w] = 0u;
    const bool live = min(p.in_len[n], p.T) > 0;
    const int U = p.tgt_len[n], off = p.tgt_off[n];
    if (first_state && live)
        for (int i = threadIdx.x; i < U; i += kCtcThreads) first_state[p.targets[off + i]] = 0x7fffffff;
    __syncthreads();
    if (live) {
        for (int i = threadIdx.x; i < U; i += kCtcThreads) {
            const int l = p.targets[off + i];
            atomicOr(&mask[l >> 5], 1u << (l & 31));
            if (first_state) atomicMin(&first_state[l], 2 * i + 1);
        }
        if (threadIdx.x == 0) atomicOr(&mask[p.blank >> 5], 1u << (p.blank & 31));
    }
    __syncthreads();
    // state index -> column of the DP's posterior row (done once here: the lookup sits on the streaming warps' divergent
    // path, where an integer division per element cost 8x the whole loop body)
    if (first_state && live) {
        for (int i = threadIdx.x; i < U; i += kCtcThreads) {
            const int l = p.targets[off + i];
            // stored as -(column + 1): a converted entry must never look like the state index of a later occurrence
            if (first_state[l] == 2 * i + 1) first_state[l] = -(ctc_col(2 * i + 1, p.kpl) + 1);
        }
    }
    __syncthreads();
    if (first_state && live && threadIdx.x == 0) first_state[p.blank] = -1;   // column 0 (state 0)
    __syncthreads();
}

__global__ void ctc_reset_progress_kernel(int* progress, int N, int T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) progress[i] = T;
}

// KPL > 0: warp DP -- block (n, part 0) runs the DP on warps 0,1 and streams on warps 2..7, blocks (n, part >= 1) only
// stream.  KPL == 0 (long label sequences): block-wide DP blocks first, then streaming blocks.
template <int KPL>
__global__ void __launch_bounds__(kCtcThreads)
ctc_bwd_kernel(const CtcBwdParams p) {
    extern __shared__ int smi[];
    const int warp = threadIdx.x >> 5;
    const int mask_words = ((p.C + 31) / 32 + 31) / 32 * 32;     // keep what follows 128-byte aligned
    uint32_t* mask = reinterpret_cast<uint32_t*>(smi);
    if constexpr (KPL > 0) {
        const int n = blockIdx.x % p.N, part = blockIdx.x / p.N;
        int* first_state = smi + mask_words;                      // [C]
        float* wpost = reinterpret_cast<float*>(first_state + (p.C + 31) / 32 * 32);   // [8][RS]
        int* dp_smem = reinterpret_cast<int*>(wpost + 8 * p.RS);
        ctc_build_mask(p, n, mask, first_state);
        const int own = (p.dbg & 4) ? 0 : 6;                      // streaming warps inside the DP block
        const int nw = 8 * (p.nparts - 1) + own;
        if (part == 0) {
            if (warp < 2) { if (!(p.dbg & 2)) ctc_beta_warp_role<KPL>(p, n, dp_smem); }
            else if (!(p.dbg & 1) && own) ctc_dense_role<true>(p, n, warp - 2, nw, mask, first_state, wpost + warp * p.RS);
        } else if (!(p.dbg & 1)) {
            ctc_dense_role<true>(p, n, own + 8 * (part - 1) + warp, nw, mask, first_state, wpost + warp * p.RS);
        }
    } else {
        if ((int)blockIdx.x < p.N) {
            ctc_beta_role(p, blockIdx.x, smi);
        } else {
            const int j = blockIdx.x - p.N, n = j % p.N, part = j / p.N;
            ctc_build_mask(p, n, mask, nullptr);
            ctc_dense_role<false>(p, n, 8 * part + warp, 8 * p.nparts, mask, nullptr, nullptr);
        }
    }
}

__global__ void ctc_sum_kernel(const float* __restrict__ nll, int N, float* __restrict__ loss) {
    // N <= a few thousand: one warp, fixed order (deterministic)
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += 32) s += nll[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) loss[0] = s;
}

// states per lane of the warp DP for 2U+1 = S states (0: too long, block-wide fallback)
static int ctc_pick_kpl(int S) {
    const int cand[14] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 13, 16, 20, 26, 32};
    for (int i = 0; i < 14; ++i)
        if (cand[i] * 32 >= S) return cand[i];
    return 0;
}
static int ctc_row_stride(int Smax) {
    const int kpl = ctc_pick_kpl(Smax);
    return kpl ? kpl * 32 : Smax;
}

static size_t ctc_alpha_floats(int T, int N, int max_target_len) {
    return (size_t)N * T * ctc_row_stride(2 * max_target_len + 1);
}

#define ASRB_CTC_KPL_SWITCH(kpl, CASE)                                                                        \
    switch (kpl) {                                                                                            \
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(10) CASE(13) CASE(16) CASE(20)   \
        CASE(26) CASE(32)                                                                                     \
        default: return ASRB_ERR_UNSUPPORTED;                                                                 \
    }

template <int KPL>
static int ctc_launch_bwd(CtcBwdParams& p, size_t smem_dp, asrb_stream_t stream) {
    const size_t smem_mask = (size_t)(((p.C + 31) / 32 + 31) / 32 * 32) * 4;
    const size_t smem_post = (size_t)((p.C + 31) / 32 * 32) * 4 + (size_t)8 * p.RS * 4;   // first_state[C] + 8 posterior rows
    size_t smem = KPL > 0 ? smem_mask + smem_post + smem_dp : (smem_dp > smem_mask ? smem_dp : smem_mask);
    // The DP warps are latency chains that share their SM with streaming warps: fewer resident blocks per SM (a larger
    // shared-memory request) leave them more issue slots while 2-3 x 8 streaming warps still saturate HBM.
    if (KPL > 0 && g_ctc_min_smem > 0 && smem < (size_t)g_ctc_min_smem) smem = (size_t)g_ctc_min_smem;
    ASRB_REQUIRE(smem <= 200 * 1024, ASRB_ERR_UNSUPPORTED);
    auto kern = ctc_bwd_kernel<KPL>;
    ASRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // One launch, about five blocks per SM.  Nothing waits on anything, so residency does not matter; the blocks that
    // carry a DP have the lowest indices and start first (the DP is the long pole).
    const int N = p.N;
    int nparts = (kNumSMs * g_ctc_blocks_per_sm + N - 1) / N;
    if (KPL > 0) {
        if (nparts < 1) nparts = 1;
        p.nparts = nparts;
        p.dbg = g_ctc_dbg;
        ctc_reset_progress_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(p.progress, N, (g_ctc_dbg & 2) ? 0 : p.T);
        ASRB_LAUNCH_OK();
        kern<<<N * nparts, kCtcThreads, smem, stream>>>(p);
    } else {
        if (nparts < 2) nparts = 2;
        p.nparts = nparts - 1;
        kern<<<N + N * (nparts - 1), kCtcThreads, smem, stream>>>(p);
    }
    ASRB_LAUNCH_OK();
    return 0;
}

}  // namespace asrb

using namespace asrb;

extern "C" {

/* DEBUG / tuning: minimum dynamic shared memory of the backward blocks (limits blocks per SM), target blocks per SM */
int asrb_debug_ctc_dbg(int bits) { g_ctc_dbg = bits; return 0; }
int asrb_debug_ctc_tuning(int min_smem_bytes, int blocks_per_sm) {
    ASRB_REQUIRE(min_smem_bytes >= 0 && blocks_per_sm >= 1 && blocks_per_sm <= 16, ASRB_ERR_BAD_ARG);
    g_ctc_min_smem = min_smem_bytes;
    g_ctc_blocks_per_sm = blocks_per_sm;
    return 0;
}

/* alpha rows [N, T, row stride], the gathered log-prob rows (same shape), target offsets[N], progress[N] */
size_t asrb_ctc_workspace_bytes(int T, int N, int max_target_len) {
    return 2 * ctc_alpha_floats(T, N, max_target_len) * sizeof(float) + (size_t)2 * N * sizeof(int) + 64;
}

/* Forward: nll[N] per utterance and loss[1] = sum_n nll[n]; ws keeps alpha (and the gathered log-probs) for the backward. */
int asrb_ctc_fwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, size_t ws_bytes, float* nll, float* loss, int T, int N,
                 int C, int max_target_len, int blank, asrb_stream_t stream) {
    ASRB_REQUIRE(log_probs && (targets || max_target_len == 0) && input_lengths && target_lengths && alpha_ws && nll && loss, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(T > 0 && N > 0 && C > 0 && max_target_len >= 0 && blank >= 0 && blank < C, ASRB_ERR_BAD_ARG);
    const int Smax = 2 * max_target_len + 1;
    ASRB_REQUIRE(ws_bytes >= asrb_ctc_workspace_bytes(T, N, max_target_len), ASRB_ERR_WORKSPACE);
    const size_t rows = ctc_alpha_floats(T, N, max_target_len);
    float* gathered = alpha_ws + rows;
    int* tgt_off = reinterpret_cast<int*>(alpha_ws + 2 * rows);
    ctc_prepare_kernel<<<1, 32, 0, stream>>>(target_lengths, tgt_off, N);
    ASRB_LAUNCH_OK();
    const int kpl = ctc_pick_kpl(Smax);
    if (kpl) {
#define ASRB_CTC_CASE(K)                                                                                              \
    case K:                                                                                                           \
        ctc_alpha_warp_kernel<K><<<N, 32, 0, stream>>>(log_probs, targets, tgt_off, input_lengths, target_lengths,    \
                                                        alpha_ws, gathered, nll, T, N, C, blank);                     \
        break;
        ASRB_CTC_KPL_SWITCH(kpl, ASRB_CTC_CASE)
#undef ASRB_CTC_CASE
    } else {
        const size_t smem = (size_t)Smax * (3 + kCtcRing) * 4;
        ASRB_REQUIRE(smem <= 200 * 1024, ASRB_ERR_UNSUPPORTED);
        ASRB_CUDA_OK(cudaFuncSetAttribute(ctc_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctc_alpha_kernel<<<N, kCtcThreads, smem, stream>>>(log_probs, targets, tgt_off, input_lengths, target_lengths, alpha_ws, nll, T, N, C, Smax, blank);
    }
    ASRB_LAUNCH_OK();
    ctc_sum_kernel<<<1, 32, 0, stream>>>(nll, N, loss);
    ASRB_LAUNCH_OK();
    return 0;
}

/* Backward: grad[T,N,C] = grad_scale[0] * d(sum nll)/d(logits); grad_scale is a DEVICE scalar (NULL = 1).
 * alpha_ws is the workspace asrb_ctc_fwd filled; its gathered-log-prob rows are consumed (a second backward needs a
 * new forward), alpha stays intact. */
int asrb_ctc_bwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, const float* nll, const float* grad_scale,
                 float* grad, int T, int N, int C, int max_target_len, int blank, asrb_stream_t stream) {
    ASRB_REQUIRE(log_probs && (targets || max_target_len == 0) && input_lengths && target_lengths && alpha_ws && nll && grad, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(T > 0 && N > 0 && C > 0 && max_target_len >= 0 && blank >= 0 && blank < C, ASRB_ERR_BAD_ARG);
    const int Smax = 2 * max_target_len + 1;
    const size_t rows = ctc_alpha_floats(T, N, max_target_len);
    int* tgt_off = reinterpret_cast<int*>(alpha_ws + 2 * rows);   // filled by asrb_ctc_fwd
    const int kpl = ctc_pick_kpl(Smax);
    CtcBwdParams p = {log_probs, targets, tgt_off, input_lengths, target_lengths, alpha_ws, alpha_ws + rows, tgt_off + N, nll,
                      grad_scale, grad, T, N, C, Smax, ctc_row_stride(Smax), kpl, blank, N, 1};
    if (kpl) {
#define ASRB_CTC_CASE(K)                                                                                   \
    case K:                                                                                                \
        return ctc_launch_bwd<K>(p, ctc_dp_smem_bytes<K>(), stream);
        ASRB_CTC_KPL_SWITCH(kpl, ASRB_CTC_CASE)
#undef ASRB_CTC_CASE
    }
    return ctc_launch_bwd<0>(p, (size_t)Smax * (6 + 2 * kCtcRing) * 4, stream);
}

}  // extern "C"
