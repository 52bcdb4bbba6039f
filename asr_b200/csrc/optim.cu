// asr_b200 -- fused AdamW step on the gradient bucket.
//
// Replaces torch.optim.AdamW(lr, betas, eps, weight_decay) as the reference constructs it
// (asr_deepspeech/trainers/__main__.py:41-47; config.yml:41-47: lr 1.5e-4, betas (0.9, 0.999), eps 1e-8, wd 1e-5) and
// steps it (trainers/deepspeech_trainer.py:86-95), with the unscale of torch.amp.GradScaler folded in:
//   g   = grad * inv_scale                     (inv_scale: device scalar, NULL = 1)
//   p  *= 1 - lr * wd                          (decoupled weight decay)
//   m   = m + (1 - b1) (g - m) ;  v = b2 v + (1 - b2) g^2
//   p  -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Pure HBM stream: read p, g, m, v and write p, m, v = 28 bytes per parameter, one pass, 16-byte vectors.
#include "common.cuh"

namespace asrb {

struct AdamWArgs {
    float decay, b2, one_minus_b1, one_minus_b2, eps, step_size, bc2_sqrt;   // complements taken in double on the host, like torch
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a) {
    p *= a.decay;
    m = fmaf(a.one_minus_b1, g - m, m);
    v = fmaf(a.one_minus_b2, g * g, v * a.b2);
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= a.step_size * (m / denom);
}

template <bool VEC>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, AdamWArgs a, const float* __restrict__ inv_scale) {
    const float gs = inv_scale ? inv_scale[0] : 1.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC) {
        const long long n4 = n / 4;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
            const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
            adamw_one(pp.x, gg.x * gs, mm.x, vv.x, a);
            adamw_one(pp.y, gg.y * gs, mm.y, vv.y, a);
            adamw_one(pp.z, gg.z * gs, mm.z, vv.z, a);
            adamw_one(pp.w, gg.w * gs, mm.w, vv.w, a);
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
        }
        for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
            adamw_one(p[i], g[i] * gs, m[i], v[i], a);
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
            adamw_one(p[i], g[i] * gs, m[i], v[i], a);
    }
}

}  // namespace asrb

using namespace asrb;

extern "C" {

/* One AdamW step (step >= 1 is the 1-based step count) on n parameters; exp_avg / exp_avg_sq are updated in place.
 * inv_scale: optional DEVICE scalar multiplied into the gradient (GradScaler unscale), NULL = 1. */
int asrb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
                    double beta2, double eps, double weight_decay, int step, const float* inv_scale, asrb_stream_t stream) {
    ASRB_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0. && lr >= 0., ASRB_ERR_BAD_ARG);
    // hyper-parameters arrive as doubles (what Python holds): torch forms 1 - beta in double before rounding to fp32
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamWArgs a = {(float)(1.0 - lr * weight_decay), (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                   (float)(lr / bc1), (float)sqrt(bc2)};
    const bool vec = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
    long long blocks = (n / (vec ? 4 : 1) + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    if (vec) adamw_kernel<true><<<(int)blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, a, inv_scale);
    else     adamw_kernel<false><<<(int)blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, a, inv_scale);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
