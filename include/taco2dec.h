/*
 * taco2dec.h -- C ABI of the B200-native Tacotron2 dual-stream mel decoder.
 *
 * Drop-in boundary for ONE hot path of PhucNguyenAH/tacotron2_subword: the per-frame
 * loop of model.Decoder.  The reference has no native boundary of its own (it is pure
 * Python/PyTorch); the seam this library replaces is the nn.Module method:
 *
 *   taco2dec_forward_teacher_forced  <->  Decoder.forward   (/root/reference/model.py:392-428)
 *   taco2dec_infer                   <->  Decoder.inference (/root/reference/model.py:430-492)
 *   taco2dec_set_weights             <->  Decoder.__init__ parameters / load_state_dict
 *                                         (model.py:142-207, attention.py:7-37, 305-322)
 *   taco2dec_config                  <->  hparams.py:55,67,72-90 fields read by Decoder
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless its name ends in
 *     _host; tensors are dense, row-major, fp32 unless stated; sizes are element counts;
 *   - the caller owns all memory (inputs, outputs, weights, workspace); the library borrows
 *     pointers for the duration of the enqueued work.  Weights passed to
 *     taco2dec_set_weights must stay alive and unchanged until the next set_weights / destroy;
 *   - calls enqueue on the given cudaStream_t (pass the caller's current stream) and return
 *     without synchronising, except where documented;
 *   - return value: 0 = ok, <0 = error (see TACO2DEC_E_*), text via taco2dec_last_error();
 *   - the device must be compute capability 10.x (B200, sm_100a); there is NO CPU fallback;
 *   - one handle per (process, device); a handle is not re-entrant.
 */
#ifndef TACO2DEC_H_
#define TACO2DEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TACO2DEC_ABI_VERSION 8

#define TACO2DEC_ATTN_SMA 0 /* StepwiseMonotonicAttention   (attention.py:291-398, hparams default) */
#define TACO2DEC_ATTN_LSA 1 /* LocationSensitiveAttention   (attention.py:7-85)                      */

#define TACO2DEC_E_ARG (-1)     /* bad argument / shape                      */
#define TACO2DEC_E_ARCH (-2)    /* device is not sm_100                      */
#define TACO2DEC_E_CUDA (-3)    /* CUDA runtime error                        */
#define TACO2DEC_E_STATE (-4)   /* weights not set, workspace too small, ... */
#define TACO2DEC_E_ABORTED (-5) /* in-kernel watchdog fired (grid barrier timeout) */

/* Kernel family selection (taco2dec_set_mode).  AUTO = latency path when eligible (B == 1, SMA, default decoder
 * dims); tensor path for 2 <= B <= 128 with default dims unless taco2dec_set_batched_precision(h, 1) asked for
 * fp32-exact batched arithmetic; else the generic any-shape path. */
#define TACO2DEC_PATH_AUTO 0
#define TACO2DEC_PATH_GENERIC 1 /* grid-barrier persistent kernel, fp32 weights in PyTorch layout, any B / SMA+LSA */
#define TACO2DEC_PATH_LATENCY 2 /* role-specialised persistent kernel, packed weights resident/streamed (batch 1) */
#define TACO2DEC_PATH_TENSOR 3  /* batched (2 <= B <= 128): LSTM / query products on tcgen05, fp16 operands, fp32 accumulate;
                                   one persistent launch for all frames when the shape allows, else the per-frame graph */
#define TACO2DEC_PATH_TENSOR_GRAPH 4 /* the same arithmetic as a CUDA graph of 6-9 kernels replayed per frame (round-1
                                        implementation; kept for A/B measurements and as the shape fallback) */
#define TACO2DEC_W_FP32 0       /* packed LSTM weights stay fp32 (parity ~1e-6)                           */
#define TACO2DEC_W_FP16 1       /* the three LSTM matrices stored fp16, fp32 accumulate (parity ~5e-5)   */

typedef struct taco2dec_handle taco2dec_handle;

/* hparams.py fields the decoder reads (model.py:131-140, 158-207). */
typedef struct taco2dec_config {
  int n_mel;        /* n_mel_channels * n_frames_per_step (n_frames_per_step must be 1) */
  int enc_dim;      /* encoder_embedding_dim            512  */
  int attn_rnn_dim; /* attention_rnn_dim                1024 */
  int dec_rnn_dim;  /* decoder_rnn_dim                  1024 */
  int prenet_dim;   /* prenet_dim                       256  */
  int attn_dim;     /* attention_dim                    128  */
  int loc_filters;  /* attention_location_n_filters     32   (LSA only) */
  int loc_kernel;   /* attention_location_kernel_size   31   (LSA only) */
  int attention;    /* TACO2DEC_ATTN_*                            */
  int n_streams;    /* 2 = BERT_Tacotron2 (char + sub-word), 1 = Tacotron2 compat */
  float p_attention_dropout; /* 0.1, applied to h AND c of the attention LSTMs (model.py:341-346) */
  float p_decoder_dropout;   /* 0.1, applied to h AND c of the decoder LSTM   (model.py:372-373) */
} taco2dec_config;

/* Per-stream weights; index 0 = char/phoneme stream, 1 = sub-word ("bert") stream.
 * PyTorch layouts exactly as in the reference state_dict. */
typedef struct taco2dec_stream_weights {
  const float* prenet_w0;    /* prenet{,_bert}.layers.0.linear_layer.weight   [prenet, n_mel]   */
  const float* prenet_w1;    /* prenet{,_bert}.layers.1.linear_layer.weight   [prenet, prenet]  */
  const float* arnn_w_ih;    /* attention_rnn{,_bert}.weight_ih               [4H, prenet+enc]  */
  const float* arnn_w_hh;    /* attention_rnn{,_bert}.weight_hh               [4H, H]           */
  const float* arnn_b_ih;    /* attention_rnn{,_bert}.bias_ih                 [4H]              */
  const float* arnn_b_hh;    /* attention_rnn{,_bert}.bias_hh                 [4H]              */
  const float* query_w;      /* attention_layer{,_bert}.query_layer.linear_layer.weight  [A, H]   */
  const float* memory_w;     /* attention_layer{,_bert}.memory_layer.linear_layer.weight [A, enc] */
  const float* v;            /* SMA: ...v.weight   LSA: ...v.linear_layer.weight          [1, A]   */
  const float* loc_conv_w;   /* LSA: ...location_layer.location_conv.conv.weight  [F, 2, K]  (else NULL) */
  const float* loc_dense_w;  /* LSA: ...location_layer.location_dense.linear_layer.weight [A, F] (else NULL) */
} taco2dec_stream_weights;

typedef struct taco2dec_weights {
  taco2dec_stream_weights stream[2];
  const float* drnn_w_ih; /* decoder_rnn.weight_ih  [4D, n_streams*(H+enc)]  input order h,ctx,h_bert,ctx_bert (model.py:362) */
  const float* drnn_w_hh; /* decoder_rnn.weight_hh  [4D, D]  */
  const float* drnn_b_ih; /* [4D] */
  const float* drnn_b_hh; /* [4D] */
  const float* proj_w;    /* linear_projection.linear_layer.weight [n_mel, D + n_streams*enc]  input order h2,ctx,ctx_bert (model.py:382) */
  const float* proj_b;    /* [n_mel] */
  const float* gate_w;    /* gate_layer.linear_layer.weight [1, D + n_streams*enc] */
  const float* gate_b;    /* [1] */
  /* decoder_rnn_bert.* is dead in decode() (model.py:375-378): never passed, never read. */
} taco2dec_weights;

/* Dropout / noise source.  Any pointer may be NULL: the kernel then draws that mask from
 * Philox4x32-10 keyed by `seed` (production mode).  Non-NULL = replay of externally drawn
 * masks (parity mode).  uint8 keep-masks: 1 = keep (scaled by 1/(1-p)). */
typedef struct taco2dec_rng {
  uint64_t seed;
  const uint8_t* prenet_keep[2][2]; /* [stream][layer] -> [rows, B, prenet]; teacher-forced rows = T+1
                                       (model.py:412-413), free-running rows = max_decoder_steps
                                       (model.py:449-450, 470-471).  Prenet dropout is ALWAYS on (model.py:23). */
  const uint8_t* lstm_keep;         /* training only: [T, 6, B, H] order attn_h, attn_c, attn_h_bert,
                                       attn_c_bert, dec_h, dec_c (model.py:341-346, 372-373) */
  const float* sma_noise[2];        /* training + SMA only: [T, B, T_stream] N(0,1) (attention.py:346-348) */
} taco2dec_rng;

/* Decoder.forward (teacher-forced).  Outputs are written in final layout. */
typedef struct taco2dec_tf_args {
  int B, T, T_in, T_sub;
  const float* memory;           /* [B, T_in, enc]   encoder outputs (char stream)            */
  const float* embeddings;       /* [B, T_sub, enc]  sub-word stream (NULL when n_streams==1) */
  const float* decoder_inputs;   /* [B, n_mel, T]    teacher-forcing targets (model.py:283-287 reads this layout) */
  const int64_t* memory_lengths; /* [B] or NULL (= all T_in).  max must equal T_in (utils.py:11) */
  const int64_t* bert_lengths;   /* [B] or NULL */
  int training;                  /* 1 = LSTM-state dropout + SMA noise on (module.training) */
  taco2dec_rng rng;
  float* mel;        /* [B, T, n_mel]  (== the storage behind the reference's [B, n_mel, T] transposed view, model.py:314-318) */
  float* gate;       /* [B, T]          gate energies (logits) */
  float* align;      /* [B, T, T_in]    */
  float* align_bert; /* [B, T, T_sub]   (NULL when n_streams==1) */
  void* workspace;
  size_t workspace_bytes;
  int independent;    /* 0 = the reference's batched semantics (Decoder.forward under training: SMA lets alignment mass
                         leak onto padded memory positions, model.py:414 + attention.py:330-338).  1 = every utterance is
                         its own sequence, exactly as if it had been run alone at batch 1 on its un-padded memory (what
                         GTA.py:35-61 does one utterance at a time): positions >= length do not exist */
  void* saved;        /* NULL, or a buffer of taco2dec_saved_layout_query().total bytes: the call keeps the activations
                         taco2dec_backward needs (tensor path only: 2 <= B <= 128, default dims) */
  size_t saved_bytes;
  const float* processed_memory[2]; /* NULL, or [B, T_s, attn_dim] = memory_layer(memory) of stream s already computed by
                         taco2dec_memprep_forward (model.py:258-261); NULL = computed inside the call */
} taco2dec_tf_args;

/* Byte offsets of the arrays inside the `saved` buffer (fp32).  State arrays have T+1 slots: slot 0 = the zero
 * initial state (model.py:237-256), slot t+1 = frame t.  s = stream index. */
typedef struct taco2dec_saved_layout {
  size_t gates1;   /* [T][S][5][H][B]   attention LSTMs: i, f, g, o (activated), new cell state before dropout */
  size_t gates2;   /* [T][5][D][B]      decoder LSTM */
  size_t h1;       /* [T+1][S][B][H]    attention-LSTM hidden state after dropout  (model.py:340-346) */
  size_t ctx;      /* [T+1][S][B][enc]  attention contexts                          (attention.py:395) */
  size_t h2;       /* [T+1][B][D]       decoder-LSTM hidden state after dropout    (model.py:371-373) */
  size_t q;        /* [T][S][B][A]      processed queries                           (attention.py:368) */
  size_t p[2];     /* [T][B][T_s]       SMA selection probabilities sigmoid(e)      (attention.py:352) */
  size_t pm[2];    /* [B][T_s][A]       processed memory                            (model.py:258-261) */
  size_t pre[2];   /* [T+1][B][prenet]  prenet output                               (model.py:412-413) */
  size_t pre0[2];  /* [T+1][B][prenet]  prenet layer-0 output (after dropout) */
  size_t total;
} taco2dec_saved_layout;

/* Byte offsets inside the `grads` buffer taco2dec_backward fills (fp32).  G = 4H gate rows, order i,f,g,o.
 * The per-frame rows are contracted with the saved activations by the caller (plain GEMMs):
 *   dW_ih[s] = dg1[s]^T . [pre[s][t] | ctx[s][t-1]],  dW_hh[s] = dg1[s]^T . h1[s][t-1],  db = sum dg1[s]
 *   dWd_ih = dg2^T . [h1[0][t] | ctx[0][t] | h1[1][t] | ctx[1][t]],  dWd_hh = dg2^T . h2[t-1]
 *   dWq[s] = dq[s]^T . h1[s][t],  dv[s] = sum_b dv[s][b],  dWm[s] = dpm[s]^T . memory[s]
 *   dmemory[s][b] = align[s][b]^T . dctx[s][:, b] + dpm[s][b] . Wm[s]
 *   prenet: dpre[s] back through the two bias-free ReLU/dropout layers (model.py:13-24). */
typedef struct taco2dec_grad_layout {
  size_t dg1;      /* [S][T][B][G]    attention-LSTM gate pre-activation gradients */
  size_t dg2;      /* [T][B][G]       decoder-LSTM gate pre-activation gradients   */
  size_t dq;       /* [S][T][B][A]    */
  size_t dctx;     /* [S][T][B][enc]  total gradient of each frame's context       */
  size_t dpre;     /* [S][T][B][prenet] gradient of the prenet output fed to frame t */
  size_t dv;       /* [S][B][A]       */
  size_t dpm[2];   /* [B][T_s][A]     gradient of the processed memory             */
  size_t dloc_dense; /* LSA: [S][B][A][F]     per-utterance sums, location_dense weight gradient (attention.py:16-17) */
  size_t dloc_conv;  /* LSA: [S][B][F][2][K]  per-utterance sums, location_conv weight gradient  (attention.py:12-15) */
  size_t scratch;  /* internal carries (d alignment state, d cell states) */
  size_t total;
} taco2dec_grad_layout;

/* Back-propagation through the teacher-forced pass (the reference obtains it from autograd over
 * Decoder.forward, model.py:392-428, under train.py's loss.backward()).  Must follow a
 * taco2dec_forward_teacher_forced call with the same shapes, rng and `saved` buffer. */
typedef struct taco2dec_bwd_args {
  int B, T, T_in, T_sub;
  const float* memory;           /* as in the forward call */
  const float* embeddings;
  const int64_t* memory_lengths;
  const int64_t* bert_lengths;
  int training;
  taco2dec_rng rng;              /* same masks / seed as the forward call (LSTM-state dropout is re-drawn) */
  const float* align;            /* [B, T, T_in]   forward outputs */
  const float* align_bert;       /* [B, T, T_sub]  */
  const float* d_mel;            /* [B, T, n_mel]  gradient of the mel output (same layout as the forward output) */
  const float* d_gate;           /* [B, T] */
  const float* d_align;          /* [B, T, T_in]  or NULL */
  const float* d_align_bert;     /* [B, T, T_sub] or NULL */
  int independent;               /* as in the forward call */
  const void* saved;
  size_t saved_bytes;
  void* grads;
  size_t grads_bytes;
} taco2dec_bwd_args;

/* Decoder.inference (free-running).  The reference is batch-1 only (model.py:461,480);
 * B > 1 here means B independent utterances, each defined as its own batch-1 run on its
 * un-padded memory (positions >= length do not exist). */
typedef struct taco2dec_infer_args {
  int B, T_in, T_sub;
  int max_decoder_steps;         /* hparams.max_decoder_steps; also the T capacity of the outputs */
  float gate_threshold;          /* stop when sigmoid(gate) > gate_threshold (strict, model.py:480) */
  const float* memory;           /* [B, T_in, enc]  */
  const float* embeddings;       /* [B, T_sub, enc] */
  const int64_t* memory_lengths; /* [B] or NULL */
  const int64_t* bert_lengths;   /* [B] or NULL */
  taco2dec_rng rng;
  float* mel;        /* [B, max_decoder_steps, n_mel]; rows >= n_frames[b] are unspecified */
  float* gate;       /* [B, max_decoder_steps]        */
  float* align;      /* [B, max_decoder_steps, T_in]  */
  float* align_bert; /* [B, max_decoder_steps, T_sub] */
  int32_t* n_frames; /* [B] frames produced (stop frame included, model.py:475-481)  */
  int32_t* reached_max; /* [B] 1 = hit max_decoder_steps without the gate firing (INFER_FLAG = False, model.py:482-485) */
  void* workspace;
  size_t workspace_bytes;
  const float* processed_memory[2]; /* as in taco2dec_tf_args */
} taco2dec_infer_args;

int taco2dec_abi_version(void);
const char* taco2dec_last_error(void);

/* device = CUDA ordinal.  Fails with TACO2DEC_E_ARCH unless the device is sm_100. */
int taco2dec_create(const taco2dec_config* cfg, int device, taco2dec_handle** out);
int taco2dec_destroy(taco2dec_handle* h);

/* Borrow the fp32 weights (the generic path reads them in place).  The latency path and the tensor path also
 * keep library-owned re-layouts (per-CTA streams, fp16 UMMA tiles) built here: call again after the weights change. */
int taco2dec_set_weights(taco2dec_handle* h, const taco2dec_weights* w, void* cuda_stream);

/* Scratch the caller must provide (state, processed memory, prenet activations). */
size_t taco2dec_workspace_bytes(const taco2dec_handle* h, int B, int T_in, int T_sub, int T_or_max_steps,
                                int teacher_forced);

int taco2dec_forward_teacher_forced(taco2dec_handle* h, const taco2dec_tf_args* a, void* cuda_stream);
int taco2dec_saved_layout_query(const taco2dec_handle* h, int B, int T_in, int T_sub, int T, taco2dec_saved_layout* out);
int taco2dec_grad_layout_query(const taco2dec_handle* h, int B, int T_in, int T_sub, int T, taco2dec_grad_layout* out);
int taco2dec_backward(taco2dec_handle* h, const taco2dec_bwd_args* a, void* cuda_stream);
int taco2dec_infer(taco2dec_handle* h, const taco2dec_infer_args* a, void* cuda_stream);

/* Weight-gradient contraction after taco2dec_backward: C[m][n] (+)= sum over rows k = (t, b) of
 * Y[k][m] * X[k][n], K = T*B.  Row (t, b) of Y starts at Y + t*y_stride_t + b*y_stride_b (elements), likewise X: the per-frame
 * gradient rows of taco2dec_grad_layout and the saved activations of taco2dec_saved_layout are addressed in place.  Runs
 * on the tcgen05 GEMM (fp16 operands, Y pre-scaled by a power of two from its absolute maximum, fp32 accumulation).
 * reuse_y = 1: the previous call on this workspace had the same Y, M, T, B -- its packed operand is used again. */
size_t taco2dec_wgrad_workspace_bytes(const taco2dec_handle* h, int M, int N, int T, int B);
int taco2dec_wgrad_gemm(taco2dec_handle* h, const float* Y, int64_t y_stride_t, int64_t y_stride_b, int M, const float* X,
                        int64_t x_stride_t, int64_t x_stride_b, int N, int T, int B, float* C, int64_t ldc, int accumulate,
                        int reuse_y, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* The two small backward contractions that are not sums over (frame, utterance) rows, fp32 on the CUDA cores:
 * sgemm_nn: C[r][n] (+)= sum_k A[r][k] B[k][n], optionally gated by the prenet's ReLU/dropout mask (x2 where mask > 0, else 0;
 * model.py:23); bmm_tn: C[b][m][n] = sum_t A[b][t][m] B[t][b][n] (d memory = alignments^T . d context, attention.py:395). */
int taco2dec_sgemm_nn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int R, int N, int K,
                      const float* mask, int64_t ldm, int accumulate, void* cuda_stream);
int taco2dec_bmm_tn(const float* A, int64_t a_stride_b, int64_t a_stride_t, const float* B, int64_t b_stride_t, int64_t b_stride_b,
                    float* C, int64_t c_stride_b, int64_t ldc, int batch, int M, int N, int T, void* cuda_stream);

/* Synchronises the stream and reports TACO2DEC_E_ABORTED if the in-kernel watchdog fired
 * during any call since the last check (asynchronous CUDA faults surface here too).  The abort word is sticky:
 * once set every later kernel of the handle bails out at once, until a check has reported and cleared it. */
int taco2dec_check(taco2dec_handle* h, void* cuda_stream);
/* The same test without waiting for the caller's stream (reads the sticky abort word on a private stream): for
 * pipelined callers that already know the call of interest has finished (e.g. after an event wait). */
int taco2dec_poll_abort(taco2dec_handle* h);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t taco2dec_launch_count(const taco2dec_handle* h);

/* The Philox keep-mask the kernels draw when a replay pointer is NULL: out[row*n + i] for
 * mask_id (prenet: stream*2+layer; lstm: 4+k) -- lets tests replay production masks. */
int taco2dec_philox_keep_mask(uint64_t seed, int mask_id, int rows, int n, float p_drop, uint8_t* out,
                              void* cuda_stream);

/* Test hook for the tcgen05 GEMM building block of the batched path: out[M][N] = A[M][K] . X[N][K]^T with
 * fp16-rounded operands and fp32 accumulation (device pointers, fp32 row-major).  M % 128 == 0,
 * K % (64*splits) == 0, 1 <= N <= 128.  Synchronises the stream. */
int taco2dec_test_gemm(int M, int N, int K, int splits, const float* A, const float* X, float* out, void* cuda_stream);

/* Select the kernel family and the storage type of the packed LSTM weights (latency path). */
int taco2dec_set_mode(taco2dec_handle* h, int path, int weight_dtype);
/* fp32_exact = 0 (default): AUTO sends 2 <= B <= 128 to the tensor path (fp16 operands, fp32 accumulation; stated
 * bound vs the fp32 reference: mel / gate <= 1e-3, alignments <= 2e-4).  fp32_exact = 1: AUTO keeps batched calls on
 * the generic fp32 kernel; saving activations for backward is then refused (there is no fp32 BPTT). */
int taco2dec_set_batched_precision(taco2dec_handle* h, int fp32_exact);
/* TACO2DEC_PATH_GENERIC / _LATENCY / _TENSOR / _TENSOR_GRAPH: the path the most recent call actually took. */
int taco2dec_last_path(const taco2dec_handle* h);

/* Optional: record CUDA events on the launching stream around the persistent decoder kernel
 * (only that kernel, not the one-off prologue kernels); taco2dec_last_kernel_ms waits for the
 * most recent profiled launch and returns its device duration. */
int taco2dec_set_profiling(taco2dec_handle* h, int on);
int taco2dec_last_kernel_ms(taco2dec_handle* h, float* ms);

/* Diagnostics: SM-clock cycles CTA 0 spent in each phase of the most recent persistent launch.
 * out16_host[2k] = work of phase k, [2k+1] = the grid barrier after it, k = P0a,P0b,A,Q,B,C,D;
 * [15] = state init + first barrier.  Synchronises the stream. */
int taco2dec_read_phase_clocks(taco2dec_handle* h, void* cuda_stream, long long* out16_host);

/* Diagnostics of the persistent batched kernel (enabled by the environment variable TACO2DEC_PB_DEBUG=<frame>,
 * TACO2DEC_PB_DEBUG_CTA=<cta>): 256 SM-clock stamps of that CTA during that frame -- [0,32) activation tile requested,
 * [32,64) operands landed, [64,96) MMAs issued, [96,128) weight tile requested, [128,144) compute-warp phase marks. */
int taco2dec_read_debug_stamps(taco2dec_handle* h, void* cuda_stream, long long* out256_host);

/* Machine probes behind bench.py's roofline for kernels whose weights live on-chip / in L2: the L2 -> SM read rate of
 * this device (GB/s, 32 MiB resident buffer read by every SM) and the latency of one cross-CTA exchange through L2 (ns:
 * one 8-byte {value, tag} store by one CTA until a polling CTA on another SM has seen it).  Synchronises the stream. */
int taco2dec_measure_machine(taco2dec_handle* h, void* cuda_stream, double* l2_read_gbs, double* hop_ns);

/* Persistent-kernel launch geometry actually used (for DESIGN.md / bench bookkeeping). */
int taco2dec_launch_geometry(const taco2dec_handle* h, int B, int* grid, int* block, int* smem_bytes);

/* ---------------------------------------------------------------------------------------------------------
 * Postnet, eval mode: replaces Postnet.forward (model.py:27-70) + the residual add
 * (model.py:557-558) + the output mask (model.py:531-541) when the module is not training.  Five Conv1d(k=5) +
 * BatchNorm1d (running statistics, folded into the weights at set_weights) + tanh on tcgen05, fp32 accumulation, operands
 * as split fp16 pairs by default (see taco2dec_postnet_set_precision).  Training mode: the building blocks further down.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct taco2dec_postnet taco2dec_postnet;

typedef struct taco2dec_postnet_layer {
  const float* conv_w;   /* postnet.convolutions.<i>.0.conv.weight  [C_out, C_in, 5] */
  const float* conv_b;   /* postnet.convolutions.<i>.0.conv.bias    [C_out] */
  const float* bn_weight;/* postnet.convolutions.<i>.1.weight       [C_out] */
  const float* bn_bias;  /* postnet.convolutions.<i>.1.bias         [C_out] */
  const float* bn_mean;  /* postnet.convolutions.<i>.1.running_mean [C_out] */
  const float* bn_var;   /* postnet.convolutions.<i>.1.running_var  [C_out] */
} taco2dec_postnet_layer;

typedef struct taco2dec_postnet_weights {
  int n_layers;                        /* hparams.postnet_n_convolutions (5) */
  float bn_eps;                        /* nn.BatchNorm1d eps (1e-5) */
  taco2dec_postnet_layer layer[8];
} taco2dec_postnet_weights;

/* n_mel <= 128, embed_dim a multiple of 128, kernel_size must be 5 (hparams.py:49), 2 <= n_layers <= 8. */
int taco2dec_postnet_create(int n_mel, int embed_dim, int kernel_size, int n_layers, int device, taco2dec_postnet** out);
int taco2dec_postnet_destroy(taco2dec_postnet* h);
/* fp16_only = 0 (default): every operand is split into two fp16 terms and the convolutions are evaluated as
 * hi.hi + hi.lo + lo.hi (fp32-grade mel_postnet, ~2.5x the tensor work); fp16_only = 1: plain fp16 operands (relative error
 * ~8e-4 of the output scale).  Takes effect at the next taco2dec_postnet_set_weights. */
int taco2dec_postnet_set_precision(taco2dec_postnet* h, int fp16_only);
/* Folds BatchNorm and packs the GEMM tiles (library-owned copies): call again when the module's tensors change. */
int taco2dec_postnet_set_weights(taco2dec_postnet* h, const taco2dec_postnet_weights* w, void* cuda_stream);
size_t taco2dec_postnet_workspace_bytes(const taco2dec_postnet* h, int B, int T);
/* mel: element (b, c, t) at mel[b*stride_b + c*stride_c + t*stride_t] (the decoder's storage is [B, T, n_mel]);
 * mel_postnet: contiguous [B, n_mel, T] = mel + postnet(mel), zero where t >= output_lengths[b] (NULL = no mask).
 * independent = 0: the reference's batched behaviour (the convolutions run over the padded frames, model.py:557);
 * independent = 1: frames >= output_lengths[b] do not exist (they are the convolutions' zero padding in every layer), so
 * row b equals the batch-1 result on its first output_lengths[b] frames -- for batched free-running synthesis. */
int taco2dec_postnet_forward(taco2dec_postnet* h, const float* mel, int64_t stride_b, int64_t stride_c, int64_t stride_t,
                             int B, int T, const int64_t* output_lengths, int independent, float* mel_postnet,
                             void* workspace, size_t workspace_bytes, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Postnet, training mode (model.py:27-70 under model.train(): BatchNorm1d with BATCH statistics, tanh, F.dropout(0.5) after
 * every layer; the reference gets the backward pass from autograd).  Activations are kept channel-last with a two-frame
 * zero halo per utterance, x_pad[b][t + 2][c]: the im2col row of frame (b, t) is then the contiguous run of 5*C floats at
 * x_pad[b][t][0], and every contraction of a layer is a "rows x weights^T" product over overlapping rows on tcgen05
 * (fp16 operands, fp32 accumulation).  The caller (tacotron2_subword_b200.model._PostnetTrain) chains these calls per layer:
 *   forward : rows_gemm(x_pad, W'[co][(k,ci)], bias, stats) -> y[n][co] and the per-channel sums of y and y^2
 *             bn_act_forward(y, mean, rstd, gamma, beta) -> next layer's x_pad interior (or the [B, C, T] output)
 *   backward: bn_act_backward(d, y, ...) -> dz in place + per-channel sums of dz and dz.zhat
 *             bn_backward_input(dz, y, ...) -> dy_pad interior;  postnet_wgrad per tap (dW);  rows_gemm(dy_pad, W''[ci][(k,co)],
 *             scale_x = 1) -> gradient w.r.t. the layer input.
 * Dropout masks come from Philox (seed, mask_id) or from a replayed uint8 array [n_rows][C] (parity tests).
 * --------------------------------------------------------------------------------------------------------- */
size_t taco2dec_postnet_rows_gemm_workspace_bytes(const taco2dec_postnet* h, int M, int K, int n_rows);
/* out[n][m] = sum_k X_n[k] W[m][k] (+ bias[m]);  row n = (b, t) is the K contiguous floats at X + b*x_stride_b + t*x_stride_t
 * (rows may overlap); W is row-major [M][K]; out has leading dimension ldo >= M.  scale_x = 1 scales the rows by a power
 * of two taken from their absolute maximum before the fp16 conversion (gradient rows).  stats (optional, double[2*M], must be
 * zeroed by the caller): += per-channel sums of out and out^2 over all rows. */
int taco2dec_postnet_rows_gemm(taco2dec_postnet* h, const float* X, int64_t x_stride_b, int64_t x_stride_t, int B, int T, int K,
                               const float* W, int M, const float* bias, int scale_x, float* out, int64_t ldo, double* stats,
                               void* workspace, size_t workspace_bytes, void* cuda_stream);
/* o = dropout(act(gamma (y - mean) rstd + beta)), act = tanh or identity; y is [B*T][C]; element (b, t, c) of the output goes to
 * out[b*out_stride_b + t*out_stride_t + c*out_stride_c].  C % 4 == 0, C <= 1024. */
int taco2dec_postnet_bn_act_forward(taco2dec_postnet* h, const float* y, int B, int T, int C, const float* mean, const float* rstd,
                                    const float* gamma, const float* beta, int use_tanh, uint64_t seed, int mask_id, float p_drop,
                                    const uint8_t* keep, float* out, int64_t out_stride_b, int64_t out_stride_t, int64_t out_stride_c,
                                    void* cuda_stream);
/* d [n_rows][C]: gradient w.r.t. the layer output, overwritten with dz (gradient w.r.t. the BatchNorm output);
 * sums (double[2*C], zeroed by the caller) += sum_n dz, sum_n dz . zhat. */
int taco2dec_postnet_bn_act_backward(taco2dec_postnet* h, float* d, const float* y, int n_rows, int C, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, int use_tanh, uint64_t seed, int mask_id, float p_drop,
                                     const uint8_t* keep, double* sums, void* cuda_stream);
/* dy = gamma rstd (dz - mean_dz - zhat mean_dz_zhat) into the interior of dy_pad [B][T + 4][C] (halo rows stay untouched). */
int taco2dec_postnet_bn_backward_input(taco2dec_postnet* h, const float* dz, const float* y, int B, int T, int C, const float* mean,
                                       const float* rstd, const float* gamma, const float* mean_dz, const float* mean_dz_zhat,
                                       float* dy_pad, void* cuda_stream);
/* Same contract as taco2dec_wgrad_gemm / taco2dec_wgrad_workspace_bytes, on a postnet handle. */
size_t taco2dec_postnet_wgrad_workspace_bytes(const taco2dec_postnet* h, int M, int N, int T, int B);
int taco2dec_postnet_wgrad(taco2dec_postnet* h, const float* Y, int64_t y_stride_t, int64_t y_stride_b, int M, const float* X,
                           int64_t x_stride_t, int64_t x_stride_b, int N, int T, int B, float* C, int64_t ldc, int accumulate,
                           int reuse_y, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Decoder inputs: memory = linear_converter(cat(encoder_outputs, cls_embeddings)) (model.py:548-549,
 * 553-554) and processed_memory = memory_layer(memory) (model.py:258-261), one handle per stream.  Two chained tcgen05
 * GEMMs over all (utterance, position) rows; every fp32 operand is split into two fp16 terms (hi + lo) and the product is
 * evaluated as hi.hi + hi.lo + lo.hi with fp32 accumulation, so the results agree with the fp32 reference to ~1e-6
 * relative (they feed the whole recurrence).  Inference / no-grad only.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct taco2dec_memprep taco2dec_memprep;
/* enc_dim (512) and attn_dim (128) multiples of 128; enc_dim + cls_dim (1280) a multiple of 64. */
int taco2dec_memprep_create(int enc_dim, int cls_dim, int attn_dim, int device, taco2dec_memprep** out);
int taco2dec_memprep_destroy(taco2dec_memprep* h);
/* converter_w [enc_dim, enc_dim + cls_dim], converter_b [enc_dim] (linear_converter{,_sub}.linear_layer.*),
 * memory_w [attn_dim, enc_dim] (decoder.attention_layer{,_bert}.memory_layer.linear_layer.weight).  converter_w may be
 * NULL: the handle then only projects an existing memory (taco2dec_memprep_project). */
int taco2dec_memprep_set_weights(taco2dec_memprep* h, const float* converter_w, const float* converter_b, const float* memory_w,
                                 void* cuda_stream);
size_t taco2dec_memprep_workspace_bytes(const taco2dec_memprep* h, int n_rows);
/* encoder_outputs [n_rows, enc_dim], cls_embeddings [n_rows, cls_dim] -> memory [n_rows, enc_dim], processed_memory
 * [n_rows, attn_dim]; n_rows = B * T (rows of padded positions are computed like any other, as in the reference). */
int taco2dec_memprep_forward(taco2dec_memprep* h, const float* encoder_outputs, const float* cls_embeddings, int n_rows,
                             float* memory, float* processed_memory, void* workspace, size_t workspace_bytes, void* cuda_stream);
/* memory [n_rows, enc_dim] -> processed_memory [n_rows, attn_dim] only. */
int taco2dec_memprep_project(taco2dec_memprep* h, const float* memory, int n_rows, float* processed_memory, void* workspace,
                             size_t workspace_bytes, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------------
 * Tacotron2Loss + the output gradients that seed the backward pass, one sweep (reference
 * loss_function.py:12-66): mel_loss = MSE(mel, target) + MSE(mel_postnet, target), gate_loss = BCEWithLogits(gate, target),
 * optional alignment L2 terms (alignloss == "L2").  d_mel is written in [B, T, n_mel] order = taco2dec_bwd_args.d_mel.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct taco2dec_loss_args {
  int B, n_mel, T;
  const float* mel;                       /* decoder mel: element (b, c, t) at mel[b*stride_b + c*stride_c + t*stride_t] */
  int64_t mel_stride_b, mel_stride_c, mel_stride_t;
  const float* mel_postnet;               /* [B, n_mel, T] */
  const float* gate;                      /* [B, T] logits */
  const float* mel_target;                /* [B, n_mel, T] */
  const float* gate_target;               /* [B, T] */
  const float* align[2];                  /* NULL, or [B, T, T_align[s]] (char / sub-word stream) */
  const float* align_target[2];
  int T_align[2];
  float* d_mel;                           /* [B, T, n_mel]   d loss / d mel (the direct MSE term only) */
  float* d_mel_postnet;                   /* [B, n_mel, T] */
  float* d_gate;                          /* [B, T] */
  float* d_align[2];                      /* [B, T, T_align[s]] when align[s] is given */
  float* losses;                          /* device [5]: total, mel_loss, gate_loss, align_loss, align_bert_loss */
  void* workspace;
  size_t workspace_bytes;
} taco2dec_loss_args;
size_t taco2dec_loss_workspace_bytes(int B, int n_mel, int T);
int taco2dec_loss_forward(const taco2dec_loss_args* a, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* TACO2DEC_H_ */
