/* asr_b200 -- C ABI of the B200-native DeepSpeech2 training-step kernels (libasr_b200.so).
 *
 * The reference (zakuro-ai/asr, package asr_deepspeech v0.4.10) is pure Python on top of torch; it has NO
 * FFI of its own.  Its seam for this path is Python dependency injection (trainers/__main__.py:23,51-58:
 * a `model` and a `criterion` are handed to DeepSpeechTrainer).  The entry points below are what the
 * Python drop-in modules in asr_b200/ bind with ctypes -- one group per reference call site:
 *
 *   asrb_spectrogram*        data/parsers/spectrogram_parser.py:45-60  (librosa.stft -> |.| -> log1p -> normalise)
 *   asrb_conv2d_mask_*       modules/blocks.py:48-55 on nn.Conv2d      (deepspeech.py:61,64)
 *   asrb_bn_act_mask_*       modules/blocks.py:48-55 on nn.BatchNorm2d + nn.Hardtanh (deepspeech.py:62-63,65-66)
 *   asrb_nchw_to_tnf*        modules/deepspeech.py:135-137              (view/transpose/contiguous)
 *   asrb_bn_rows_*           modules/blocks.py:16-21,85-86 SequenceWise(BatchNorm1d) ; deepspeech.py:104
 *   asrb_gemm_tn             every Linear-shaped product: GRU/LSTM input projection (blocks.py:88), FC head
 *                            (deepspeech.py:105) and all of their dgrad / wgrad products
 *   asrb_rnn_*               modules/blocks.py:87-92 pack -> nn.GRU / nn.LSTM (bidirectional) -> pad -> sum(2)
 *   asrb_log_softmax_*       trainers/deepspeech_trainer.py:110 ; blocks.py:62 (softmax) ; greedy_decoder.py:61 (argmax)
 *   asrb_ctc_*               trainers/deepspeech_trainer.py:111 with torch.nn.CTCLoss(reduction="sum") (trainers/__main__.py:53)
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; every pointer is DEVICE memory of the current CUDA device unless a name ends in
 *     `_host`; fp32 unless stated; tensors are dense row-major with the shape given in the comment.
 *   - the caller owns every buffer including workspaces; the library allocates nothing and keeps no state
 *     besides the process-wide debug flag word.
 *   - asynchronous: work is enqueued on `stream`; no host synchronisation inside.
 *   - return value: 0 = ok, negative = ASRB_ERR_* (bad argument / unsupported shape; nothing was launched),
 *     positive = a cudaError_t.  Never throws, never falls back to another implementation.
 */
#ifndef ASR_B200_H_
#define ASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* asrb_stream_t; /* == cudaStream_t */

#define ASRB_OK 0
#define ASRB_ERR_BAD_ARG (-1)
#define ASRB_ERR_ALIGNMENT (-2)
#define ASRB_ERR_UNSUPPORTED (-3)
#define ASRB_ERR_WORKSPACE (-4)
#define ASRB_ERR_DRIVER (-5)
#define ASRB_ERR_TENSORMAP (-6)

#define ASRB_GEMM_ACCUMULATE 1 /* C += A*B^T instead of C = A*B^T */

#define ASRB_RNN_GRU 0
#define ASRB_RNN_LSTM 1

/* Debug switches (tests only; default 0 = the tensor-core product path). */
#define ASRB_DEBUG_SIMT_GEMM 1u /* asrb_gemm_tn runs a plain fp32 CUDA-core kernel */
#define ASRB_DEBUG_SIMT_RNN 2u  /* asrb_rnn_{fwd,bwd} compute the recurrent product with fp32 CUDA-core dot products */

int asrb_version(void);
const char* asrb_strerror(int code);
int asrb_set_debug_flags(unsigned flags);
int asrb_debug_gemm_tile(int force_bn, int gain_pct);
/* Caps the persistent CTAs of the GEMM launches that follow (0 = no cap, n < 0 queries); returns the old value.  For
 * work issued on a second stream beside a kernel that needs its own SMs (asr_b200/functional.py, weight gradients). */
int asrb_gemm_cta_limit(int n);
/* epilogue of the tcgen05 GEMM: 1 (default) staged in shared memory and written by TMA stores (needs ldc % 4 == 0 and a
 * 16-byte aligned C, else the direct form is used), 0 direct row-per-lane stores; v < 0 queries.  Returns the old value. */
int asrb_debug_gemm_tma_store(int v);

/* ---------------------------------------------------------------- GEMM (tcgen05, TF32 operands, fp32 accumulate)
 * C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N]).  lda/ldb/ldc are row strides in elements; lda, ldb multiples of 4. */
int asrb_gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias, int M,
                 int N, int K, int flags, asrb_stream_t stream);
/* Same product with bf16 operands (lda, ldb in bf16 elements, multiples of 8); fp32 accumulate and C. */
int asrb_gemm_tn_bf16(const void* A, int lda, const void* B, int ldb, float* C, int ldc, const float* bias, int M, int N,
                      int K, int flags, asrb_stream_t stream);
/* out[c, r] = in[r, c] */
int asrb_transpose(const float* in, long long rows, int cols, int ld_in, float* out, int ld_out, asrb_stream_t stream);
/* out[c, r] = bf16(in[r, c]), ld_out in bf16 elements */
int asrb_transpose_bf16(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out, asrb_stream_t stream);
/* 3xTF32 operand expansion: mode 0 -> [x_hi | x_lo | x_hi], mode 1 -> [x_hi | x_hi | x_lo]  (out: [rows, 3*cols]) */
int asrb_split3(const float* in, long long rows, int cols, int ld_in, float* out, int ld_out, int mode,
                asrb_stream_t stream);

/* ---------------------------------------------------------------- recurrent layers
 * `bf16` selects the operand precision of the recurrent product h W_hh^T (1: bf16 weights + bf16 copy of the state,
 * needs H % 8 == 0; 0: tf32 straight from the fp32 state, needs H % 4 == 0).  Accumulation, gate math, stored
 * states and gradients are fp32 in both modes.  B <= 128.
 * Geometry of the persistent kernel (nj = hidden units per CTA, P = CTAs per direction) and the sizes in BYTES of
 * the two packed-weight buffers: */
int asrb_rnn_plan(int cell, int H, int B, int bf16, int* nj, int* P, size_t* wpack_fwd_bytes, size_t* wpack_bwd_bytes);
/* w_hh_*: [gates*H, H] (torch weight_hh_l0 / weight_hh_l0_reverse).  Either output may be NULL. */
int asrb_rnn_pack_weights(int cell, int H, int B, int bf16, const float* w_hh_fwd, const float* w_hh_rev,
                          void* wpack_fwd, void* wpack_bwd, asrb_stream_t stream);
/* gi [T,B,2,G] = x W_ih^T + b_ih for both directions; b_hh [2,G]; lengths int32[B];
 * out: hseq [2,T+2,B,H] (slot t+1 = step t; slots 0,T+1 zero), hseq_bf16 [2,T+2,B,Hp] bf16 with Hp = H rounded up to 64 (bf16 mode, else NULL),
 * cseq like hseq (LSTM only, else NULL), saved: asrb_rnn_saved_floats floats (GRU: r,z,n,W_hn h+b_hn ; LSTM: i,f,g,o);
 * counters: uint32[128] scratch (step counters per direction and chain, 64 bytes apart, for two passes; one flag word). */
int asrb_rnn_fwd(int cell, int bf16, const float* gi, const float* b_hh, const void* wpack_fwd, const int32_t* lengths,
                 float* hseq, void* hseq_bf16, float* cseq, float* saved, uint32_t* counters, int T, int B, int H,
                 asrb_stream_t stream);
/* The same with the direction sum of blocks.py:92 folded in: out_sum [T,B,H] = hseq[0][1..T] + hseq[1][1..T] (NULL: not
 * wanted).  The tensor-memory kernel adds both directions' tiles into it with TMA reduce-adds; the other kernels run
 * asrb_rnn_sum_dirs behind the recurrence. */
int asrb_rnn_fwd_sum(int cell, int bf16, const float* gi, const float* b_hh, const void* wpack_fwd, const int32_t* lengths,
                     float* hseq, void* hseq_bf16, float* cseq, float* saved, float* out_sum, uint32_t* counters, int T, int B,
                     int H, asrb_stream_t stream);
/* dout [T,B,H] (gradient of the direction-summed output).  out: dgi [T,B,2,G] (for the input-gradient GEMM);
 * dgiT [2G, ldT] = its transpose (row = dir*G + gate*H + unit, column = t*B + b; ldT >= T*B, multiple of 4) and, for
 * GRU, dghT [2G, ldT] = the transposed hidden-side gate gradients (they differ from dgiT in the n gate; LSTM: NULL)
 * -- the K-major operands of the weight-gradient GEMMs, written directly so no transpose pass is needed.  dgi, dgiT
 * and dghT are fp32 in tf32 mode and BF16 in bf16 mode (ldT then a multiple of 8; consumed by asrb_gemm_tn_bf16); the
 * recurrent operand of the next step is
 * dgh_bf16 [2,T,B,Gp] with Gp = G rounded up to 64 (bf16 mode, dgh may be NULL) or dgh [2,T,B,G] fp32 (tf32 mode, dgh_bf16 may be NULL). */
int asrb_rnn_bwd(int cell, int bf16, const float* dout, const void* wpack_bwd, const int32_t* lengths,
                 const float* hseq, const float* cseq, const float* saved, void* dgi, float* dgh, void* dgh_bf16,
                 void* dgiT, void* dghT, long long ldT, uint32_t* counters, int T, int B, int H,
                 asrb_stream_t stream);
/* floats in the saved-gates buffer (slice-major layout private to asrb_rnn_fwd / asrb_rnn_bwd) */
size_t asrb_rnn_saved_floats(int cell, int H, int B, int bf16, int T);
int asrb_debug_rnn_trace(long long* trace);
int asrb_debug_rnn_chunk(int blocks);
int asrb_debug_rnn_ksplit(int on);
int asrb_debug_rnn_dbg(int bits);
/* launches of the tensor-memory recurrent kernels whose second (release) pass had to run since the library was loaded
 * (verified hand-over, asrb_debug_rnn_dbg bit 4096); synchronises the device; -1 on a CUDA error */
int asrb_debug_rnn_redos(void);
/* out[T,B,H] = hseq[0][1..T] + hseq[1][1..T] */
int asrb_rnn_sum_dirs(const float* hseq, float* out, int T, int B, int H, asrb_stream_t stream);

/* ---------------------------------------------------------------- MaskConv pieces (NCHW; lengths int32[B] or NULL)
 * y = mask(conv2d(x, w) + bias): zero for time index >= lengths[b].  Hout/Wout must equal the conv output size. */
int asrb_conv2d_mask_fwd(const float* x, const float* w, const float* bias, const int32_t* lengths, float* y, int B,
                         int Cin, int Hin, int Win, int Cout, int Hout, int Wout, int KH, int KW, int SH, int SW,
                         int PH, int PW, asrb_stream_t stream);
/* dx = conv2d^T(mask(dy), w) */
int asrb_conv2d_mask_bwd_data(const float* dy, const float* w, const int32_t* lengths, float* dx, int B, int Cin,
                              int Hin, int Win, int Cout, int Hout, int Wout, int KH, int KW, int SH, int SW, int PH,
                              int PW, asrb_stream_t stream);
/* dw[Cout,Cin,KH,KW], dbias[Cout] (may be NULL) from mask(dy) and x.  ws: asrb_nchw_reduce_workspace_bytes(B,Cout,Hout,Wout). */
int asrb_conv2d_mask_bwd_weight(const float* dy, const float* x, const int32_t* lengths, float* dw, float* dbias,
                                double* ws, size_t ws_bytes, int B, int Cin, int Hin, int Win, int Cout, int Hout,
                                int Wout, int KH, int KW, int SH, int SW, int PH, int PW, asrb_stream_t stream);
size_t asrb_nchw_reduce_workspace_bytes(int B, int C, int H, int W);
/* BatchNorm2d batch statistics over ALL positions of y[B,C,H,W]; running stats updated when non-NULL. */
int asrb_bn2d_stats(const float* y, float* mean, float* invstd, float* running_mean, float* running_var,
                    float momentum, float eps, double* ws, size_t ws_bytes, int B, int C, int H, int W,
                    asrb_stream_t stream);
/* eval-mode statistics: mean = running_mean, invstd = rsqrt(running_var + eps) */
int asrb_bn_eval_stats(const float* running_mean, const float* running_var, float eps, int C, float* mean,
                       float* invstd, asrb_stream_t stream);
/* z = mask(hardtanh(mask(bn(y)), lo, hi)); has_bn / has_act switch the two stages off (plain mask when both 0). */
int asrb_bn_act_mask_fwd(const float* y, const int32_t* lengths, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, int has_bn, int has_act, float lo, float hi, float* z,
                         int B, int C, int H, int W, asrb_stream_t stream);
int asrb_bn_act_mask_bwd(const float* dz, const float* y, const int32_t* lengths, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, int has_bn, int has_act, float lo,
                         float hi, int training, float* dy, float* dgamma, float* dbeta, double* ws, size_t ws_bytes,
                         int B, int C, int H, int W, asrb_stream_t stream);
/* out[n][c][r] = in[n][r][c] with explicit leading dimensions / batch strides (elements) */
int asrb_transpose_batched(const float* in, int rows, int cols, long long ld_in, long long batch_stride_in, float* out,
                           long long ld_out, long long batch_stride_out, int nbatch, asrb_stream_t stream);

int asrb_copy_rows_padded(const float* in, long long ld_in, float* out, long long ld_out, long long rows, int cols,
                          asrb_stream_t stream);
int asrb_nchw_channel_sums(const float* a, const int32_t* lengths, float* out, double* ws, size_t ws_bytes, int B, int C,
                           int H, int W, asrb_stream_t stream);

/* ---------------------------------------------------------------- 32->32 channel conv on tcgen05 (implicit GEMM, TF32)
 * (deepspeech.py:64 "conv2").  Time stride 1, KW <= 16.  Activations NHWC for fwd/dgrad sources, NCHW elsewhere. */
int asrb_conv32_supported(int Cin, int Cout, int KH, int KW, int SH, int SW, int PH, int PW);
int asrb_conv32_pack_weights(const float* w, float* pack_fwd, float* pack_dgrad, int KH, int KW, asrb_stream_t stream);
int asrb_conv32_fwd(const float* x_nhwc, const float* pack_fwd, const float* bias, const int32_t* lengths, float* y,
                    int B, int Hin, int Win, int Hout, int Wout, int KH, int KW, int SH, int PH, int PW,
                    asrb_stream_t stream);
int asrb_conv32_bwd_data(const float* dy_nhwc, const float* pack_dgrad, float* dx, int B, int Hin, int Win, int Hout,
                         int Wout, int KH, int KW, int SH, int PH, int PW, asrb_stream_t stream);
/* Row-grouped variants: NR = 2 or 4 output rows (forward: consecutive; data gradient: SH apart) share every source strip
 * in one work item, their tap matrices stacked along N = 32 NR, so the strip tile is read from shared memory once for NR
 * rows.  pack_rows holds [KH + SH (NR-1)][KW][32 NR][32] floats, built from the plain pack of the same mode (0 forward,
 * 1 data gradient). */
int asrb_conv32_pack_rows(const float* pack, float* pack_rows, int KH, int KW, int SH, int NR, int mode, asrb_stream_t stream);
int asrb_conv32_fwd_rows(const float* x_nhwc, const float* pack_rows, const float* bias, const int32_t* lengths, float* y,
                         int B, int Hin, int Win, int Hout, int Wout, int KH, int KW, int SH, int PH, int PW, int NR,
                         asrb_stream_t stream);
int asrb_conv32_bwd_data_rows(const float* dy_nhwc, const float* pack_rows, float* dx, int B, int Hin, int Win, int Hout,
                              int Wout, int KH, int KW, int SH, int PH, int PW, int NR, asrb_stream_t stream);
size_t asrb_conv32_bwd_weight_workspace_bytes(int B, int Hin, int Win, int Hout, int Wout);
/* operands of the weight-gradient product: 1 (default) bf16 copies, 0 TF32 (then lddy % 4 == 0); v < 0 queries */
int asrb_debug_conv_wgrad_bf16(int v);
/* work split of the bf16 weight gradient: 1 (default) one CTA per (kernel-row class, M tile, chunk of source rows), 0 one
 * CTA per (kernel row, chunk); v < 0 queries */
int asrb_debug_conv_wgrad_cls(int v);
int asrb_conv32_bwd_weight(const float* x, const float* dy, int lddy, float* dw, float* ws, size_t ws_bytes, int B,
                           int Hin, int Win, int Hout, int Wout, int KH, int KW, int SH, int PH, int PW,
                           asrb_stream_t stream);

/* ---------------------------------------------------------------- 1->32 channel, time-stride-2 conv on tcgen05
 * (deepspeech.py:61 "conv1": kernel (KH,11), stride (2,2), padding (PH,5)); forward and weight gradient. */
int asrb_conv1_supported(int Cin, int Cout, int F, int KH, int KW, int SH, int SW, int PH, int PW);
size_t asrb_conv1_workspace_bytes(int B, int F, int T, int for_wgrad);
int asrb_conv1_fwd(const float* x, const float* w, const float* bias, const int32_t* lengths, float* y, float* ws,
                   size_t ws_bytes, int B, int F, int T, int Hout, int Wout, int KH, int PH, asrb_stream_t stream);
int asrb_conv1_bwd_weight(const float* x, const float* dy, int lddy, float* dw, float* ws, size_t ws_bytes, int B,
                          int F, int T, int Hout, int Wout, int KH, int PH, asrb_stream_t stream);

/* ---------------------------------------------------------------- row-matrix kernels, x[R = T*N, cols] */
size_t asrb_rows_workspace_bytes(int cols);
int asrb_bn_rows_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     int training, float momentum, float eps, float* mean, float* invstd, float* y, float* ws,
                     size_t ws_bytes, long long R, int cols, asrb_stream_t stream);
int asrb_bn_rows_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                     int training, float* dx, float* dgamma, float* dbeta, float* ws, size_t ws_bytes, long long R,
                     int cols, asrb_stream_t stream);
int asrb_col_sums(const float* a, int lda, float* out, float* ws, size_t ws_bytes, long long R, int cols,
                  asrb_stream_t stream);
int asrb_row_sums(const float* a, long long ld, float* out, int rows, long long cols, asrb_stream_t stream);
int asrb_row_sums_bf16(const void* a, long long ld, float* out, int rows, long long cols, asrb_stream_t stream);
/* logits[R, ld] -> log_probs[R,C] / probs[R,C] / argmax int64[R] (each optional) */
int asrb_log_softmax_fwd(const float* logits, int ld, float* log_probs, float* probs, long long* argmax, long long R,
                         int C, asrb_stream_t stream);
int asrb_log_softmax_bwd(const float* g, const float* log_probs, float* dlogits, int ld, long long R, int C,
                         asrb_stream_t stream);

/* Greedy CTC collapse of the frame-wise argmax (decoders/greedy_decoder.py:27-46, remove_repetitions=True): for each
 * utterance n the kept classes / their frame indices are written densely to labels[n, 0..counts[n]) / offsets[n, ...)
 * (row stride T); sizes (int32[N], may be NULL = T) bounds the frames read. */
int asrb_greedy_collapse(const long long* argmax, const int32_t* sizes, int N, int T, int blank, int32_t* labels,
                         int32_t* offsets, int32_t* counts, asrb_stream_t stream);

/* ---------------------------------------------------------------- CTC (blank-extended alpha/beta, reduction = sum) */
size_t asrb_ctc_workspace_bytes(int T, int N, int max_target_len);
int asrb_debug_ctc_tuning(int min_smem_bytes, int blocks_per_sm);
int asrb_debug_ctc_dbg(int bits);
int asrb_ctc_fwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, size_t ws_bytes, float* nll, float* loss, int T, int N,
                 int C, int max_target_len, int blank, asrb_stream_t stream);
int asrb_ctc_bwd(const float* log_probs, const int32_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha_ws, const float* nll, const float* grad_scale,
                 float* grad, int T, int N, int C, int max_target_len, int blank, asrb_stream_t stream);

/* ---------------------------------------------------------------- optimizer
 * torch.optim.AdamW step of trainers/__main__.py:41-47 / deepspeech_trainer.py:86-95 on n contiguous parameters, with the
 * GradScaler unscale folded in (inv_scale: optional device scalar).  step is the 1-based step count. */
int asrb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
                    double beta2, double eps, double weight_decay, int step, const float* inv_scale, asrb_stream_t stream);

/* ---------------------------------------------------------------- Lookahead convolution
 * modules/blocks.py:96-121: depthwise Conv1d over time with `context` taps looking ahead (zero padding at the end) on
 * [T,N,H] activations: y[t,n,f] = sum_k w[f,k] x[t+k,n,f]; w is conv.weight [H,1,context].  has_act fuses the
 * Hardtanh(lo,hi) that follows it in the unidirectional model (modules/deepspeech.py:94-101).  Backward: y is the
 * forward output (read only when has_act); dx and/or dw may be NULL; dw [H,context] is overwritten. */
int asrb_lookahead_fwd(const float* x, const float* w, float* y, int T, int N, int H, int context, int has_act,
                       float lo, float hi, asrb_stream_t stream);
int asrb_lookahead_bwd(const float* dy, const float* x, const float* y, const float* w, float* dx, float* dw, int T,
                       int N, int H, int context, int has_act, float lo, float hi, asrb_stream_t stream);

/* ---------------------------------------------------------------- spectrogram (STFT -> |.| -> log1p -> normalise) */
size_t asrb_spectrogram_workspace_bytes(int B, int max_samples, int n_fft, int hop);
int asrb_dft_basis(float* basis_cat /* [2*(n_fft/2+1), 3*n_fft] */, int n_fft, asrb_stream_t stream);
int asrb_spectrogram(const float* wav, long long wav_ld, const int32_t* n_samples, const float* window,
                     const float* basis_cat, float* spec, int normalize, void* ws, size_t ws_bytes, int B,
                     int max_samples, int n_fft, int hop, asrb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H_ */
