// asr_b200 -- log-magnitude spectrogram on the GPU, replacing the CPU path of
// asr_deepspeech/data/parsers/spectrogram_parser.py:45-60:
//   librosa.stft(y, n_fft, hop, win_length=n_fft, window) [center=True, zero padding of n_fft/2, 1 + len//hop frames]
//   -> magphase -> |D| -> log1p -> float32 -> (x - mean) / std   (torch unbiased std over the whole [F, T] map)
//
// The DFT of every frame is one tensor-core GEMM: frames[B*T, n_fft] x basis[2F, n_fft]^T (cos rows then sin rows).
// To keep fp32-level accuracy on TF32 tensor cores both operands are expanded hi/lo ("3xTF32"): the frame matrix is
// written as [hi | lo | hi] and the basis as [hi | hi | lo], so one K = 3*n_fft GEMM yields
// hi*hi + lo*hi + hi*lo with fp32 accumulation (the dropped lo*lo term is ~2^-22 relative).
#include "common.cuh"

namespace asrb {

__device__ __forceinline__ void split_tf32(float v, float* hi, float* lo) {
    const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
    *hi = h;
    *lo = v - h;
}

// basis_cat [2F, 3*n_fft]: row f = cos(2 pi f n / N), row F+f = sin(2 pi f n / N), each as [hi | hi | lo]
__global__ void dft_basis_kernel(float* __restrict__ basis, int n_fft, int F) {
    const int total = 2 * F * n_fft;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / n_fft, n = i % n_fft;
        const int f = r < F ? r : r - F;
        const int ph = (int)(((long long)f * n) % n_fft);
        const double ang = 2.0 * 3.14159265358979323846 * (double)ph / (double)n_fft;
        const float v = (float)(r < F ? cos(ang) : sin(ang));
        float hi, lo;
        split_tf32(v, &hi, &lo);
        float* o = basis + (size_t)r * 3 * n_fft;
        o[n] = hi; o[n_fft + n] = hi; o[2 * n_fft + n] = lo;
    }
}

// frames [B*Tmax, 3*n_fft] as [hi | lo | hi] of window[n] * y[t*hop + n - n_fft/2] (zero outside the utterance)
__global__ void stft_frames_kernel(const float* __restrict__ wav, long long wav_ld, const int* __restrict__ n_samples,
                                   const float* __restrict__ window, float* __restrict__ frames, int B, int Tmax,
                                   int n_fft, int hop) {
    const long long total = (long long)B * Tmax * n_fft;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % n_fft);
        const long long row = i / n_fft;
        const int t = (int)(row % Tmax), b = (int)(row / Tmax);
        const int ns = n_samples[b];
        const int nframes = 1 + ns / hop;
        float v = 0.f;
        if (t < nframes) {
            const int idx = t * hop + n - n_fft / 2;
            if (idx >= 0 && idx < ns) v = wav[(size_t)b * wav_ld + idx] * window[n];
        }
        float hi, lo;
        split_tf32(v, &hi, &lo);
        float* o = frames + row * 3 * n_fft;
        o[n] = hi; o[n_fft + n] = lo; o[2 * n_fft + n] = hi;
    }
}

// spec[b][f][t] = log1p(|re + i im|) for t < frames(b), else 0 ; accumulates per-utterance sum / sum of squares
__global__ void __launch_bounds__(256)
stft_mag_kernel(const float* __restrict__ reim, int ld, const int* __restrict__ n_samples, float* __restrict__ spec,
                double* __restrict__ stats, int Tmax, int F, int hop) {
    __shared__ float tile[32][33];
    __shared__ double red[2][8];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int nframes = 1 + n_samples[b] / hop;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int t = t0 + i, f = f0 + tx;
        float v = 0.f;
        if (t < Tmax && t < nframes && f < F) {
            const float* r = reim + ((size_t)b * Tmax + t) * ld;
            const float re = r[f], im = r[F + f];
            v = log1pf(sqrtf(re * re + im * im));
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    double s0 = 0, s1 = 0;
    for (int i = ty; i < 32; i += 8) {
        const int f = f0 + i, t = t0 + tx;
        if (f < F && t < Tmax) {
            const float v = tile[tx][i];
            spec[((size_t)b * F + f) * Tmax + t] = v;
            s0 += v; s1 += (double)v * v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (tx == 0) { red[0][ty] = s0; red[1][ty] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
        for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
        atomicAdd(stats + 2 * b, a);
        atomicAdd(stats + 2 * b + 1, c);
    }
}

__global__ void stft_normalize_kernel(float* __restrict__ spec, const int* __restrict__ n_samples,
                                      const double* __restrict__ stats, int Tmax, int F, int hop, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % Tmax);
        const int b = (int)(i / ((long long)F * Tmax));
        const int nframes = 1 + n_samples[b] / hop;
        if (t >= nframes) continue;  // padding stays zero (asr_deepspeech/functional.py:18,27)
        const double cnt = (double)F * nframes;
        const double mean = stats[2 * b] / cnt;
        double var = (stats[2 * b + 1] - cnt * mean * mean) / (cnt - 1.0);
        if (var < 0) var = 0;
        spec[i] = (float)(((double)spec[i] - mean) / sqrt(var));
    }
}

}  // namespace asrb

using namespace asrb;

extern "C" {

size_t asrb_spectrogram_workspace_bytes(int B, int max_samples, int n_fft, int hop) {
    const size_t rows = (size_t)B * (1 + max_samples / hop);
    const size_t ld = (size_t)round_up(2 * (n_fft / 2 + 1), 4);
    return rows * 3 * n_fft * 4 + rows * ld * 4 + (size_t)B * 2 * sizeof(double) + 256;
}

int asrb_dft_basis(float* basis_cat, int n_fft, asrb_stream_t stream) {
    ASRB_REQUIRE(basis_cat && n_fft >= 8 && n_fft % 4 == 0, ASRB_ERR_BAD_ARG);
    dft_basis_kernel<<<kNumSMs, 256, 0, stream>>>(basis_cat, n_fft, n_fft / 2 + 1);
    ASRB_LAUNCH_OK();
    return 0;
}

/* wav [B, wav_ld] (zero padded), n_samples int32[B], window [n_fft], basis_cat from asrb_dft_basis;
 * out spec [B, 1, F = n_fft/2+1, Tmax = 1 + max_samples/hop] */
int asrb_spectrogram(const float* wav, long long wav_ld, const int32_t* n_samples, const float* window,
                     const float* basis_cat, float* spec, int normalize, void* ws, size_t ws_bytes, int B,
                     int max_samples, int n_fft, int hop, asrb_stream_t stream) {
    ASRB_REQUIRE(wav && n_samples && window && basis_cat && spec && ws, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(B > 0 && B <= 65535 && max_samples > 0 && n_fft >= 8 && n_fft % 4 == 0 && hop > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(ws_bytes >= asrb_spectrogram_workspace_bytes(B, max_samples, n_fft, hop), ASRB_ERR_WORKSPACE);
    const int F = n_fft / 2 + 1, Tmax = 1 + max_samples / hop;
    const int ld = round_up(2 * F, 4);
    const size_t rows = (size_t)B * Tmax;
    float* frames = reinterpret_cast<float*>(ws);
    float* reim = frames + rows * 3 * n_fft;
    double* stats = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(reim + rows * ld) + 15) & ~uintptr_t(15));
    ASRB_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)B * 2 * sizeof(double), stream));
    const long long nf = (long long)rows * n_fft;
    const int g = (int)((nf + 255) / 256 < kNumSMs * 8 ? (nf + 255) / 256 : kNumSMs * 8);
    stft_frames_kernel<<<g, 256, 0, stream>>>(wav, wav_ld, n_samples, window, frames, B, Tmax, n_fft, hop);
    ASRB_LAUNCH_OK();
    int rc = asrb_gemm_tn(frames, 3 * n_fft, basis_cat, 3 * n_fft, reim, ld, nullptr, (int)rows, 2 * F, 3 * n_fft, 0, stream);
    if (rc) return rc;
    stft_mag_kernel<<<dim3(ceil_div(Tmax, 32), ceil_div(F, 32), B), 256, 0, stream>>>(reim, ld, n_samples, spec, stats, Tmax, F, hop);
    ASRB_LAUNCH_OK();
    if (normalize) {
        const long long total = (long long)B * F * Tmax;
        const int g2 = (int)((total + 255) / 256 < kNumSMs * 8 ? (total + 255) / 256 : kNumSMs * 8);
        stft_normalize_kernel<<<g2, 256, 0, stream>>>(spec, n_samples, stats, Tmax, F, hop, total);
        ASRB_LAUNCH_OK();
    }
    return 0;
}

}  // extern "C"
