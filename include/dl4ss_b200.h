/*
 * dl4ss_b200 -- C ABI of the B200 (sm_100a) separation hot path.
 *
 * The reference (shincling/DL4SS) is pure Python: it has NO FFI / operator interface.  The
 * boundary of its hot path is the set of Python calls listed per function below; this header
 * is the C surface a binding for those calls uses (ctypes stub in INTEGRATION.md; the shipped
 * binding is dl4ss_b200/_lib.py).
 *
 * Conventions
 *   - every function returns 0 on success, a negative DL4SS_E* code on failure;
 *     dl4ss_last_error() returns a thread-local message.  Nothing throws across the ABI.
 *   - all tensor arguments are caller-owned DEVICE pointers, contiguous, fp32 unless stated.
 *     No function allocates device memory; scratch is passed in by the caller
 *     (size from the matching *_workspace_bytes query).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - layouts are the reference's: features [B,T,F], complex spectra [B,T,F,2] (re,im) --
 *     `convert2`, TDAA_beta/predata_fromList_cRM_123.py:37-41 --, masks [B,S,T,F] or
 *     [B,S,T,F,2], waveforms [B,L] / [B,S,hop*(T-1)], RNN weights in torch.nn.LSTM/GRU
 *     state-dict order (gate order LSTM i,f,g,o ; GRU r,z,n).
 */
#ifndef DL4SS_B200_H
#define DL4SS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DL4SS_OK            0
#define DL4SS_EINVAL       -1   /* bad argument (shape, enum, null pointer)            */
#define DL4SS_EUNSUPPORTED -2   /* valid request this build has no kernel for          */
#define DL4SS_ECUDA        -3   /* CUDA runtime / launch error (message has the cause) */
#define DL4SS_EWORKSPACE   -4   /* workspace too small                                  */

#define DL4SS_FEAT_NONE 0       /* only the complex spectrum                            */
#define DL4SS_FEAT_ABS  1       /* |X|          (TDAA_beta/predata_fromList.py:194)     */
#define DL4SS_FEAT_LOG  2       /* log(|X|+eps) (TDAA_beta/predata_fromList.py:188-192) */

#define DL4SS_WAV_F32 0
#define DL4SS_WAV_F64 1         /* the reference's mix_wav is float64                   */

#define DL4SS_MASK_NONE    0    /* spec is already per source [B,S,T,F,2]               */
#define DL4SS_MASK_REAL    1    /* mask [B,S,T,F]   x mixture (a9)                      */
#define DL4SS_MASK_COMPLEX 2    /* mask [B,S,T,F,2] x mixture, complex multiply (a10)   */

#define DL4SS_CELL_LSTM 0
#define DL4SS_CELL_GRU  1

#define DL4SS_ACT_NONE    0
#define DL4SS_ACT_TANH    1
#define DL4SS_ACT_SIGMOID 2

#define DL4SS_ATT_DOT      0    /* sigmoid(<emb,q>)          EvalVer.py:216-226          */
#define DL4SS_ATT_DOT_CRM  1    /* K*tanh(<emb,q1|q2>) + decompress  cRM_EvalVer.py:260-271,512 */

int         dl4ss_version(void);
const char *dl4ss_last_error(void);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
uint64_t    dl4ss_launch_count(void);

/* ---- K1: frame + window + FFT + |.|/log  ------------------------------------------------
 * Replaces librosa.core.spectrum.stft(y, n_fft, hop) + np.abs / np.log / convert2 at
 *   TDAA_beta/predata_fromList.py:166,179,188-200 ; Torch_multi/predata_multiAims.py:168,185,194-205 ;
 *   TDAA_beta/predata_fromList_cRM_123.py:215-218,232-235,250-255.
 * wav [B,L] (f32 or f64), centre=True reflect padding, T = 1 + L/hop, F = 1 + n_fft/2.
 * window: n_fft fp32 taps on the device.  feat_out [B,T,F] (may be NULL when feat_mode NONE),
 * cplx_out [B,T,F,2] (may be NULL).  n_fft must be 256 in this build.  conj!=0 conjugates the
 * spectrum (librosa <= 0.5 flavour). */
int dl4ss_stft_feat(const void *wav, int wav_dtype, int B, int L, int n_fft, int hop,
                    const float *window, int feat_mode, float eps, int conj,
                    float *feat_out, float *cplx_out, void *stream);

/* ---- K6: mask x mixture + iFFT + window + overlap-add + normalise -----------------------
 * Replaces `pred = mask*mix_feas` ... `pred*exp(1j*angle(mix))` ... librosa.istft(spec.T, hop):
 *   TDAA_beta/main_run_sstune_EvalVer.py:466-470,55-65 ; cRM: ...cRM_EvalVer.py:545-553,96-99.
 * spec: mixture [B,T,F,2] (MASK_REAL / MASK_COMPLEX) or per-source [B,S,T,F,2] (MASK_NONE).
 * wav_out [B,S,hop*(T-1)].  (mask*|X|)*exp(j*angle X) == mask*X, so the phase never
 * materialises.  Imaginary parts of the DC / Nyquist bins are ignored, as in irfft. */
int dl4ss_mask_istft(const float *mask, int mask_kind, const float *spec, int B, int S, int T,
                     int n_fft, int hop, const float *window, float *wav_out, void *stream);

/* ---- dense projection  C[M,N] = act(A[M,K] * W[N,K]^T + bias[N])  (fp32) ------------------
 * The nn.Linear contractions on the path: RNN input projections and MIX_SPEECH.Linear
 * (TDAA_beta/main_run_sstune_EvalVer.py:290,298-299) when the embedding is materialised,
 * ADDJUST.layer (:369,375).  lda/ldw/ldc are row pitches in elements.  bias may be NULL. */
int dl4ss_linear_fwd(const float *A, int lda, const float *W, int ldw, const float *bias,
                     float *C, int ldc, int M, int N, int K, int act, void *stream);

/* ---- K3: bidirectional recurrent layer (persistent kernel) -------------------------------
 * Replaces one layer of nn.LSTM / nn.GRU(batch_first, bidirectional)
 *   (TDAA_beta/main_run_sstune_EvalVer.py:282-289,293 ; ...cRM_EvalVer.py:345-351,356).
 * xproj [B,T,2,G*H]: x*W_ih^T + b_ih (+ b_hh for every LSTM gate and the GRU r,z gates),
 *   direction-major inside a frame, G = 4 (LSTM i,f,g,o) / 3 (GRU r,z,n).
 * whh [2,G*H,H] ; bhn [2,H] (GRU b_hn; NULL for LSTM) ; y [B,T,2H] (fwd | reverse halves).
 * Zero initial state.  workspace: dl4ss_rnn_workspace_bytes() bytes, zero-filled by callee.
 * Optional training outputs (NULL in inference): gates_save [B,T,2,G*H] post-activation gates
 * and cell_save [B,T,2,H] (LSTM c_t). */
size_t dl4ss_rnn_workspace_bytes(int B, int T, int H, int cell);
int dl4ss_rnn_layer_fwd(int cell, const float *xproj, const float *whh, const float *bhn,
                        float *y, int B, int T, int H, float *gates_save, float *cell_save,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- K3 on warp-level tensor cores for large hidden sizes (bf16x3 mma.sync.m16n8k16) ---------------------
 * Same contract as dl4ss_rnn_layer_fwd.  For hidden sizes the TMEM-resident tcgen05 form cannot hold -- the speaker
 * classifier's BLSTM 3x600 (MIX_SPEECH_classifier, TDAA_beta/main_run_sstune_EvalVer.py:305-326): W_hh resident ONCE
 * on the chip (a CTA owns a direction and 10 hidden units and walks every 16-utterance tile of the batch each step).
 * Supported when dl4ss_rnn_mma_supported(H, cell) != 0 (H a multiple of 10, 2*H/10 <= number of SMs, slice fits
 * shared memory: H <= 740 on B200).  workspace: dl4ss_rnn_mma_workspace_bytes() bytes, 256-byte aligned, zero-filled
 * by the callee (release counters + the L2-resident bf16 exchange buffer). */
int    dl4ss_rnn_mma_supported(int H, int cell);
size_t dl4ss_rnn_mma_workspace_bytes(int B, int T, int H, int cell);
int    dl4ss_rnn_layer_mma_fwd(int cell, const float *xproj, const float *whh, const float *bhn,
                               float *y, int B, int T, int H, float *gates_save, float *cell_save,
                               void *workspace, size_t workspace_bytes, void *stream);

/* ---- K3 on the tensor cores (tcgen05, bf16x3 split of h and W_hh, fp32 TMEM accumulation) -----------
 * Same contract as dl4ss_rnn_layer_fwd except that W_hh arrives pre-packed by dl4ss_rnn_tc_pack_whh:
 * whh [2,G*H,H] fp32 -> bf16 hi/lo planes [2][2 dir * 4H rows][Kp], row dir*4H + 4*u + g = gate g of
 * unit u (the unit-major order the kernel's epilogue transposes inside 4-lane groups; GRU's 4th row is
 * zero), dl4ss_rnn_tc_whh_bytes(H) bytes, 16-byte aligned.  Pack once per weight update.
 * Supported when dl4ss_rnn_tc_supported(H, cell) != 0 (H a multiple of 4, <= 320 -- every reference
 * config uses 300); otherwise DL4SS_EUNSUPPORTED and the caller uses dl4ss_rnn_layer_fwd.
 * workspace: dl4ss_rnn_tc_workspace_bytes() bytes, 256-byte aligned, zero-filled by the callee
 * (release counters + the L2-resident bf16 exchange buffer the CTAs pass h_t through). */
int    dl4ss_rnn_tc_supported(int H, int cell);
/* profiling hook: device buffer of steps*16 int64 that receives CTA 0's per-phase clock64() stamps of
 * subsequent dl4ss_rnn_layer_tc_fwd launches (NULL switches it off; off by default) */
void   dl4ss_rnn_tc_set_trace(void *dev_buf, int steps);
/* utterance tiles (32 utterances) a CTA of dl4ss_rnn_layer_tc_fwd interleaves: 0 = as few as the co-resident CTAs allow
 * (shortest launch), 1..3 = at least that many (fewer CTAs per launch, so that launches on different streams run side by
 * side).  Process-wide; initial value from the environment variable DL4SS_RNN_TILES_PER_CTA. */
void   dl4ss_rnn_tc_set_tiles_per_cta(int tiles);
/* persistent CTAs a tcgen05 projection launch (dl4ss_linear_tc_fwd, dl4ss_emb_attn_mask_tc_fwd, split-K form) may
 * occupy: 0 = one per SM; n = at most n, so that two pipelined batches on different streams share the SMs (a recurrent
 * launch of one batch next to a projection launch of the other).  Process-wide; initial value from DL4SS_GEMM_MAX_CTAS. */
void   dl4ss_gemm_tc_set_max_ctas(int ctas);
/* 2-CTA (tcgen05.mma.cta_group::2) form of the plain projections: 0 = never, 1 = when a launch owns the GPU (no CTA cap;
 * default), 2 = also under a cap (cap / 2 CTA pairs).  Initial value from DL4SS_GEMM_2CTA. */
void   dl4ss_gemm_tc_set_two_cta(int mode);
/* launch dl4ss_rnn_layer_tc_fwd as 2-CTA clusters (placement only: its CTAs then fill whole TPCs, leaving whole TPCs to the
 * 2-CTA projections of another in-flight batch).  Initial value from DL4SS_RNN_CLUSTER_PAIRS (default 0). */
void   dl4ss_rnn_tc_set_cluster_pairs(int on);
size_t dl4ss_rnn_tc_whh_bytes(int H);
int    dl4ss_rnn_tc_pack_whh(int cell, const float *whh, int H, void *planes, void *stream);
size_t dl4ss_rnn_tc_workspace_bytes(int B, int T, int H, int cell);
/* Optional fused outputs (NULL to skip): y_planes = y already split into bf16 hi/lo planes
 * [2][B*T][Kp], Kp = 2H rounded up to 64, for the next dl4ss_linear_tc_fwd / dl4ss_emb_attn_mask_tc_fwd (the
 * kernel writes columns [0,2H); the caller keeps the padding columns zero); hmean_out [B,2H] = mean over T of
 * y, the ADDJUST input (pass it to dl4ss_speaker_query_fwd as h with T = 1).  * y may be NULL when y_planes is given: the layer's output then leaves as bf16 hi/lo planes only (the operand form of the next
 * projection), half of the output bytes -- the inference pipeline reads nothing else.
 */
int dl4ss_rnn_layer_tc_fwd(int cell, const float *xproj, const void *whh_planes, const float *bhn,
                           float *y, int B, int T, int H, float *gates_save, float *cell_save,
                           void *y_planes, float *hmean_out,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ---- K4: Linear + tanh + speaker attention + mask, fused ---------------------------------
 * Replaces MIX_SPEECH.Linear+tanh (EvalVer.py:298-301), the S-fold expand().contiguous()
 * (:453-455), ATTENTION 'dot' (:216-226) and, for cRM, K*tanh + decompression
 * (cRM_EvalVer.py:260-271,512), without materialising emb[B,T,F,E].
 * h [B*T, K] ; W [F*E, K] ; bias [F*E] ; q [B,S,E] (DOT) or [B,S,2E] (DOT_CRM) ;
 * mask_out [B,S,T,F] or [B,S,T,F,2].  crm_k / crm_c: K=10, C=0.1 in the reference
 * (crm_c <= 0 skips the decompression and returns K*tanh).  workspace: at least one
 * utterance of T*F*E floats; dl4ss_emb_attn_mask_workspace_bytes() returns the preferred size. */
size_t dl4ss_emb_attn_mask_workspace_bytes(int B, int T, int F, int E);
int dl4ss_emb_attn_mask_fwd(const float *h, const float *W, const float *bias, const float *q,
                            int B, int T, int F, int E, int K, int S, int mode,
                            float crm_k, float crm_c, float *mask_out,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---- tensor-core projections (tcgen05, bf16x3 split: fp32-grade results) ------------------
 * Same contractions as dl4ss_linear_fwd / dl4ss_emb_attn_mask_fwd on the 5th-gen tensor cores.
 * Operands are first split into two bf16 planes (x = hi + lo) laid out [2][R][Kp],
 * Kp = K rounded up to 64, zero padded; three MMAs per k-block (hi*hi + hi*lo + lo*hi) accumulate
 * in fp32 TMEM.  `planes` buffers are caller-owned, 16-byte aligned, dl4ss_split_bf16_bytes()
 * bytes.  The fused K4 epilogue is built for E = 50 (every reference config) and S <= 4;
 * anything else returns DL4SS_EUNSUPPORTED and the caller uses dl4ss_emb_attn_mask_fwd. */
size_t dl4ss_split_bf16_bytes(long long R, int K);
int dl4ss_split_bf16(const float *x, int ld, long long R, int K, void *planes, void *stream);
/* transposing split: x [R,C] fp32 -> planes [2][C][Rp] of x^T (Rp = R rounded up to 64, zero padded,
 * dl4ss_split_bf16_bytes(C, R) bytes): the operand form for contractions over the row dimension (dW = dY^T X). */
int dl4ss_split_bf16_t(const float *x, long long ld, int R, int C, void *planes, void *stream);
int dl4ss_linear_tc_fwd(const void *a_planes, const void *w_planes, const float *bias, float *C,
                        int ldc, int M, int N, int K, int act, void *stream);
/* dl4ss_linear_tc_fwd (act none) whose A planes have row pitch lda >= K elements (a multiple of 8) instead of K rounded up to
 * 64: [2][M][lda]; columns [K, lda) are never read (the k tail of the last block reads as zeros).  Lets a producer's planes
 * (e.g. the BPTT kernel's gate-gradient planes, pitch 2*G*H) feed the projection as they lie. */
int dl4ss_linear_tc_lda_fwd(const void *a_planes, int lda, const void *w_planes, const float *bias, float *C, int ldc,
                            int M, int N, int K, void *stream);
/* Same contraction without an activation, for outputs of only a few 128x256 tiles and a long K (the backward
 * contractions over B*T rows, dW = dY^T X): the k-blocks of every tile are cut into as many splits as fill the
 * SMs and the partial tiles meet in C through fp32 atomic adds (C is zeroed by the callee; the summation order of
 * the splits is not fixed, so results can differ in the last bits from run to run).  Falls back to the single-pass
 * launch when the tiles already fill the machine. */
int dl4ss_linear_tc_splitk_fwd(const void *a_planes, const void *w_planes, const float *bias, float *C,
                               int ldc, int M, int N, int K, void *stream);
/* C[m][n] = sum over (b,t) of A[b][t + shift_a][col0_a + m] * B[b][t + shift_b][col0_b + n]   (m < M, n < N), fp32 results of
 * bf16x3 products -- the weight gradients dW = dY^T X of the training step (TDAA_beta/main_run_sstune_EvalVer.py:673
 * loss.backward()) straight from ROW-MAJOR bf16 hi/lo planes [2][B][T][ld] (wa / wb valid columns per row, pitch lda / ldb a
 * multiple of 8 elements, 16-byte aligned; col0_a / col0_b multiples of 8): both UMMA operands are MN-major, so no transposed
 * copy is made.  Frames outside
 * [0,T) read as zeros, which is how the recurrent weight gradients pair dgates[t] with h[t-1] (shift_b = -1) or h[t+1] (+1)
 * without a shifted copy.  The contraction is cut into splits that fill the SMs; C (pitch ldc) is zeroed by the callee. */
int dl4ss_linear_tc_tn_splitk_fwd(const void *a_planes, int lda, int wa, int col0_a, int shift_a,
                                  const void *b_planes, int ldb, int wb, int col0_b, int shift_b,
                                  float *C, int ldc, int M, int N, int B, int T, void *stream);
int dl4ss_emb_attn_mask_tc_fwd(const void *h_planes, const void *w_planes, const float *bias,
                               const float *q, int B, int T, int F, int E, int K, int S, int mode,
                               float crm_k, float crm_c, float *mask_out, void *stream);

/* ---- attention over a MATERIALISED embedding (module-level drop-in) ----------------------
 * ATTENTION.forward(mix_hidden[N,T,F,E], query[N,E|2E]) (EvalVer.py:210-226).
 * emb [Nb,TF,E]; emb_batch_stride = 0 shares one embedding across the S queries of an
 * utterance (the reference's expand without the copy). */
int dl4ss_attn_dot_fwd(const float *emb, long long emb_utt_stride, const float *q, int B, int S,
                       int TF, int E, int mode, float crm_k, float crm_c, float *mask_out,
                       void *stream);

/* ---- speaker queries: embedding gather (+ ADDJUST self-tune) --------------------------------
 * e = table[idx[b,s]] (idx NULL: `table` is e[B,S,EQ] itself) ;
 * q[b,s,:] = residual*e + Wadj * [mean_t h[b,t,:] ; e]          (Wadj NULL: q = e, gather only)
 *   SPEECH_EMBEDDING.forward TDAA_beta/main_run_sstune_EvalVer.py:357-361 (2E wide for cRM);
 *   ADDJUST.forward + residual :371-377,445-446.
 * h [B,T,C] ; table [num_spk,EQ] ; idx int64 [B,S] ; Wadj [EQ,C+EQ] ; q [B,S,EQ] ;
 * hmean_out [B,C] optional (kept for backward) ; err_flag: device int set to 1 on an
 * out-of-range speaker id (nn.Embedding raises; the caller checks it). */
int dl4ss_speaker_query_fwd(const float *h, int B, int T, int C, const float *table, int num_spk,
                            int EQ, const long long *idx, int S, const float *Wadj, int residual,
                            float *q, float *hmean_out, int *err_flag, void *stream);

/* ---- K5: mask x mixture + MSE losses ------------------------------------------------------
 * real: loss_main = mean((mask*feas - y)^2), loss_sum = mean((sum_s mask - 1)^2)
 *       (EvalVer.py:470,487-492) ; cRM: mean((Re)^2), mean((Im)^2) (cRM_EvalVer.py:545-568).
 * partial sums are accumulated in double into loss_out[2] (caller zeroes it; the means are
 * taken by the caller so shards of one global batch can be all-reduced first). */
int dl4ss_mask_loss_fwd(const float *mask, int mask_kind, const float *mix, const float *target,
                        int B, int S, int TF, double *loss_out, void *stream);
/* Permutation-invariant form: pair_out[b][s][s'] (double, [B,S,S], zeroed by the callee) = sum over T*F of
 * |mask[b,s] x mix[b] - target[b,s']|^2 (cRM: real and imaginary squared errors summed), every prediction against every
 * target in one pass; the caller searches the S! permutations (dl4ss_b200.pit_mask_loss).  The reference pairs sources
 * by sorted speaker index (TDAA_beta/main_run_sstune_EvalVer.py:632-639); PIT is the north-star's extension.  S <= 4. */
int dl4ss_mask_pair_loss_fwd(const float *mask, int mask_kind, const float *mix, const float *target,
                             int B, int S, int TF, double *pair_out, void *stream);

/* ---- a1: batched mixture synthesis ---------------------------------------------------------
 * The per-source preprocessing of the reference generators (TDAA_beta/predata_fromList.py:140-177):
 * crop to lengths[b,s] (NULL: L), subtract the mean, divide by max|.|, zero-pad to L, gain 10^(dB/20),
 * mixture = sum over sources.  src [B,S,L] fp32 ; gains_db [B,S] ; src_out [B,S,L] (may be NULL, may alias
 * src) ; mix_out [B,L].  S <= 16. */
int dl4ss_premix_fwd(const float *src, const int *lengths, const float *gains_db, int B, int S, int L,
                     float *src_out, float *mix_out, void *stream);
/* The same with the AUGMENT_DATA circular shift of the training generators (TDAA_beta/predata_fromList.py:150-153,
 * `signal = np.append(signal[shift:], signal[:shift])`, applied to the normalised signal BEFORE the zero padding):
 * shifts [B,S] int (NULL: none), source (b,s) is rotated left by shifts[b,s] mod lengths[b,s] samples.  With shifts
 * src_out must not alias src. */
int dl4ss_premix_shift_fwd(const float *src, const int *lengths, const int *shifts, const float *gains_db,
                           int B, int S, int L, float *src_out, float *mix_out, void *stream);

/* ---- n2: short-lag cross-correlations for on-device BSS-Eval -----------------------------------
 * Replaces the FFT correlations inside mir_eval.separation.bss_eval_sources (`_compute_reference_correlations`,
 * `_compute_projection_filters`; called at Torch_multi/bss_test.py:55 on wav files read back from disk).
 *   out[b, i, j, k] = sum_m x[b, i, m] * y[b, j, m + lag0 + k],  k in [0, nlags),  y taken as 0 outside [0, N)
 * x [B,Sx,N], y [B,Sy,N] fp32 ; out [B,Sx,Sy,nlags] fp64 (products and sums in double).  The Gram matrix of the
 * 512-tap projection is G[(i,a),(j,c)] = out_ref_ref[b,i,j, flen-1 + a - c] (lag0 = -(flen-1), nlags = 2*flen-1)
 * and its right-hand side D[(i,a)] = out_ref_est[b,i,e,a] (lag0 = 0, nlags = flen). */
int dl4ss_xcorr_f64(const float *x, const float *y, int B, int Sx, int Sy, int N, int nlags, int lag0,
                    double *out, void *stream);

/* ---- n4: Discriminator forward (replaces cuDNN conv + cuBLAS of TDAA_beta/main_run_sstune_EvalVer.py:328-346) --------
 * y[n,co,oy,ox] = relu(b[co] + sum_{ci,ky,kx} w[co,ci,ky,kx] * x[n,ci,2*oy+ky,2*ox+kx]): 3x3 kernel, stride 2, no padding,
 * NCHW fp32, OH = (IH-3)/2+1, OW = (IW-3)/2+1; b may be NULL.  Supported: Cin = 1 (any Cout), or Cin a multiple of 16 with
 * Cout = 64 and OW <= 32 (the reference's layers: 1->64 on [313,129], 64->64 on [156,64], 64->64 on [77,31]).
 * N <= 65535 samples per call. */
int dl4ss_conv3x3s2_relu_fwd(const float *x, const float *w, const float *b, float *y, int N, int Cin, int IH, int IW,
                             int Cout, void *stream);
/* out[n] = sigmoid(<x[n, 0:K], w> + b[0]) (the Linear(36480 -> 1) + sigmoid score, :336,345); b may be NULL */
int dl4ss_rowdot_sigmoid_fwd(const float *x, const float *w, const float *b, float *out, int N, int K, void *stream);

/* ---- training step, backward side -----------------------------------------------------------
 * loss = l0 + 0.5*l1 (real; EvalVer.py:641,659-666) or l_re + l_im (cRM; cRM_EvalVer.py:741-743).
 * dl4ss_mask_loss_bwd: dmask = d(loss)/d(mask), same layout as mask.
 *   real: c0 = 2*g/N0, c1 = g/N1 with N0 = B*S*T*F, N1 = B*T*F of the GLOBAL batch, g = upstream grad;
 *   cRM : c0 = c1 = 2*g/N0.
 * dl4ss_attn_dot_bwd: emb [B,TF,E] = tanh(z) ; q [B,S,E|2E] ; mask/dmask [B,S,TF(,2)] ->
 *   dz [B,TF,E] = d(loss)/dz (may alias emb) and dq [B,S,E|2E] (zeroed by the callee).
 * dl4ss_rnn_bwd_step: BPTT gate arithmetic of backward step s (s = 0 is the LAST forward step of each
 *   direction).  dy [B,T,2H]; dh_rec [2,B,H] = dg_cur(previous call) x W_hh (host GEMM; unused at s = 0);
 *   gates_save / cell_save / y from dl4ss_rnn_layer_*fwd; carry [2,B,H] state (LSTM dc, GRU dh*z);
 *   dgx [B,T,2,G*H] = d/d(xproj) ; dgh (GRU only; NULL for LSTM) = d/d(W_hh h + b_hh) ;
 *   dg_cur [2,B,G*H] = this step's recurrent-side gate gradients, contiguous for the next GEMM. */
int dl4ss_mask_loss_bwd(const float *mask, int mask_kind, const float *mix, const float *target, int B,
                        int S, int TF, float c0, float c1, float *dmask, void *stream);
int dl4ss_attn_dot_bwd(const float *emb, const float *q, const float *mask, const float *dmask, int B,
                       int S, int TF, int E, int mode, float crm_k, float crm_c, float *dz, float *dq,
                       void *stream);
/* the same with dz emitted as bf16 hi/lo planes [2][B*T][ldp] (row (b,t), column f*E + e; ldp >= F*E, a multiple of 8; pad
 * columns are left untouched: zero them once) instead of fp32 -- the operand form of the GEMMs that consume dz
 * (dl4ss_linear_tc_tn_splitk_fwd for dW_lin, dl4ss_linear_tc_lda_fwd for dh); emb is only read */
int dl4ss_attn_dot_bwd_planes(const float *emb, const float *q, const float *mask, const float *dmask, int B,
                              int S, int T, int F, int E, int mode, float crm_k, float crm_c, void *dz_planes,
                              int ldp, float *dq, void *stream);
int dl4ss_rnn_bwd_step(int cell, int s, const float *dy, const float *dh_rec, const float *gates_save,
                       const float *cell_save, const float *y, float *carry, float *dgx, float *dgh,
                       float *dg_cur, int B, int T, int H, void *stream);

/* ---- the whole BPTT chain of one bidirectional layer as ONE persistent kernel -------------------------
 * Replaces autograd's T sequential LSTM / GRU backward steps under `loss.backward()`
 * (TDAA_beta/main_run_sstune_EvalVer.py:673 ; TDAA_beta/main_run_sstune_cRM_EvalVer.py:751).
 * dy [B,T,2H] = d(loss)/d(layer output); whh [2,G*H,H] fp32 (the nn.LSTM / nn.GRU weight_hh of both directions);
 * gates_save / cell_save / y as written by dl4ss_rnn_layer_*fwd.  Outputs: dgx [B,T,2,G*H] = d/d(xproj) and,
 * GRU only, dgh [B,T,2,G*H] = d/d(W_hh h + b_hh) (NULL for LSTM, where it equals dgx); both 16-byte aligned.
 * The weight / input gradients are GEMMs over these arrays (dl4ss_linear_tc_fwd on transposed operands).
 * Supported when dl4ss_rnn_bwd_supported(H, cell) != 0 (H a multiple of 20 whose 20-unit W_hh slice fits
 * shared memory: every reference config, H = 300); otherwise DL4SS_EUNSUPPORTED and the caller walks the
 * chain with dl4ss_rnn_bwd_step.  workspace: dl4ss_rnn_bwd_workspace_bytes() bytes, 256-byte aligned,
 * zero-filled by the callee (release counters). */
int    dl4ss_rnn_bwd_supported(int H, int cell);
/* profiling hook: device buffer of steps*8 int64 that receives CTA 0's per-phase clock64() stamps of subsequent
 * dl4ss_rnn_layer_bwd launches (NULL switches it off; off by default) */
void   dl4ss_rnn_bwd_set_trace(void *dev_buf, int steps);
size_t dl4ss_rnn_bwd_workspace_bytes(int B, int T, int H, int cell);
int    dl4ss_rnn_layer_bwd(int cell, const float *dy, const float *whh, const float *gates_save,
                           const float *cell_save, const float *y, float *dgx, float *dgh, int B, int T, int H,
                           void *workspace, size_t workspace_bytes, void *stream);
/* Tensor-core form (bf16x3 on warp-level mma.sync.m16n8k16, fp32 accumulation): the per-step product has K = G*H
 * against a 16 x 20 output, which suits many warps issuing small MMAs, not tcgen05's single-thread 128-row UMMAs.
 * Same contract plus `xplanes`: caller-owned bf16 [2 (hi,lo)][B*T][2][GHg] (GHg = G*H rounded up to 8,
 * dl4ss_rnn_bwd_tc_xplanes_bytes() bytes, 16-byte aligned) through which the CTAs exchange the recurrent-side gate
 * gradients; the caller zero-fills it ONCE (the kernel never writes the GHg - G*H pad columns, which must be zero).
 * LSTM: `dgx` may be NULL -- the planes then are the only copy of the gate gradients (weight, input and bias gradients
 * can all be taken from them); the GRU needs dgx and dgh. */
int    dl4ss_rnn_bwd_tc_supported(int H, int cell);
size_t dl4ss_rnn_bwd_tc_xplanes_bytes(int B, int T, int H, int cell);
int    dl4ss_rnn_layer_bwd_tc(int cell, const float *dy, const float *whh, const float *gates_save,
                              const float *cell_save, const float *y, float *dgx, float *dgh, void *xplanes,
                              int B, int T, int H, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DL4SS_B200_H */
