// backward.cuh -- back-propagation through time for the teacher-forced decoder (tensor path), included by
// taco2dec.cu after batched.cuh.
//
// The reference gets this pass from autograd over the per-frame Python loop (model.py:392-428 run under
// train.py:245-256: loss.backward()).  Here the recurrent part is a reverse-time frame loop, again one CUDA graph
// replayed per frame (frame index counts down in device memory):
//
//   bw_pointwise2   dh2[t] (from frame t+1's GEMM + the projection rows) -> decoder-LSTM gate gradients dG2[t]
//   GEMM            dX2 = [Wd_ih | Wd_hh]^T . dG2^T        tcgen05, bf16 operands, fp32 accumulate  (model.py:362-371)
//   bw_attention    d ctx[t] -> d alpha' -> stepwise-monotonic recurrence -> d energies -> dq, dv, d processed_memory
//                                                                                     (attention.py:330-398)
//   (bw_attention_lsa for location-sensitive attention, attention.py:7-85)
//   bw_pointwise1   dh1[t] (frame t+1's GEMM + dX2 + Wq^T dq) -> attention-LSTM gate gradients dG1[t]
//   GEMM            dX1 = [W_ih | W_hh]^T . dG1^T          tcgen05                           (model.py:337-346)
// (d prenet[t], rows of dX1, is saved by the next frame's first kernel; the frame counter is moved by bw_pointwise1.)
//
// Everything that is a plain sum over (frame, utterance) -- all weight gradients, d memory, the hoisted prenet -- is
// left as saved per-frame gradient rows (dG1, dG2, dq, d ctx, d prenet) that the host wrapper contracts with the saved
// activations in a handful of large library GEMMs after the loop.
//
// Gate gradients are handed to the tensor cores in bf16 (fp32 range: no loss scaling), weights in bf16 as well.
#pragma once

#include <cuda_bf16.h>

namespace bw {

using bt::A;
using bt::E;
using bt::H;
using bt::M;
using bt::P;
using bt::K1;
constexpr int G = 4 * H;                  // gate rows = K of the backward GEMMs
constexpr int SPLITSB1 = 4, SPLITSB2 = 4;

struct Bufs {
  unsigned char* a1t;   // [S][K1/128][G/64] tiles: rows = [prenet | ctx | h1] input features, K = gate rows
  unsigned char* a2t;   // [K2/128][G/64]
  unsigned char* dg1;   // [S][G/64][NPAD x 64] bf16 tiles of dG1[t]
  unsigned char* dg2;   // [G/64][NPAD x 64]
  float* dx1;           // [S][SPLITSB1][K1][NPAD]
  float* dx2;           // [SPLITSB2][K2][NPAD]
  int NPAD, K2;
};

struct Grads {
  bt::Saved sv;
  const float* p_saved[2];   // [T][B][Ts]
  const float* align[2];     // [B][T][Ts] forward outputs
  const float* d_align[2];   // [B][T][Ts] or null
  const float *d_mel, *d_gate;   // [B][T][M], [B][T]
  float* dg1;      // [S][T][B][G]
  float* dg2;      // [T][B][G]
  float* dq;       // [S][T][B][A]
  float* dctx;     // [S][T][B][E]
  float* dpre;     // [S][T][B][P]
  float* dv;       // [S][B][A]      per-utterance partial sums
  float* dpm[2];   // [B][Ts][A]     d processed_memory
  float* dalpha[2];// [B][Ts]        carry: d alignment state
  float *dc1, *dc2;// [S][B][H], [B][H] carry: d cell state
  float* dyh;      // [T][H][B]      projection backwards, decoder-LSTM hidden columns   (bw_dy_all)
  float* dyc;      // [T][B][S*E]    projection backwards, context columns
  // location-sensitive attention only
  float* dcum[2];  // [B][Ts]        carry: d cumulative attention weights
  float* dz[2];    // [B][Ts][A]     scratch: d tanh-argument of the current frame
  float* dwd;      // [S][B][A*LF]   per-utterance partial sums, location_dense weight gradient [a][f]
  float* dwc;      // [S][B][LF*2*LK] per-utterance partial sums, location_conv weight gradient [f][c][k]
};

// transposed weights: dst rows r = feature of the concatenation [src0 cols | src1 cols], K = source row (gate)
__global__ void pack_concat_tiles_T_kernel(const float* __restrict__ src0, int K0, const float* __restrict__ src1, int K1c,
                                           int gates, unsigned char* __restrict__ dst) {
  const int R = K0 + K1c, kb_total = gates / tc::kBlockK;
  const size_t total = (size_t)R * gates;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % R), k = (int)(i / R);       // r fastest: coalesced source reads
    const float v = r < K0 ? src0[(size_t)k * K0 + r] : src1[(size_t)k * K1c + (r - K0)];
    const size_t tile = ((size_t)(r / 128) * kb_total + (k / tc::kBlockK)) * tc::kATileBytes;
    *reinterpret_cast<__nv_bfloat16*>(dst + tile + tc::tile_offset_bytes(128, r % 128, k % tc::kBlockK)) = __float2bfloat16(v);
  }
}

// one LSTM cell backwards (nn.LSTMCell + the dropout on h and c, model.py:340-346 / 371-373).
//   dh: gradient w.r.t. the post-dropout hidden state, *dc_carry: w.r.t. the post-dropout cell state (updated
//   in place to the gradient w.r.t. the previous frame's post-dropout cell state).
__device__ __forceinline__ void lstm_cell_backward(const float* sv, size_t gs, float c_prev, float mh, float mc, float dh,
                                                   float* dc_carry, float (&dgate)[4]) {
  const float gi = sv[0], gf = sv[gs], gg = sv[2 * gs], go = sv[3 * gs], cn = sv[4 * gs];
  const float tcn = tanhf(cn);
  const float dhp = dh * mh;
  const float dcn = *dc_carry * mc + dhp * go * (1.0f - tcn * tcn);
  *dc_carry = dcn * gf;
  dgate[0] = dcn * gg * gi * (1.0f - gi);
  dgate[1] = dcn * c_prev * gf * (1.0f - gf);
  dgate[2] = dcn * gi * (1.0f - gg * gg);
  dgate[3] = dhp * tcn * go * (1.0f - go);
}

// ---- LSTM pointwise kernels: block = 32 utterances x 8 hidden units, thread (bl = tid & 31, jl = tid >> 5) ----------
// Lanes run along the batch, so the split-K partials ([row][NPAD]) and the saved gates ([unit][B]) are read coalesced;
// the results are transposed through shared memory so that every (utterance, gate) writes its 8 consecutive units as
// one 16-byte bf16 core-matrix row of the GEMM operand tile and 32 contiguous bytes of the fp32 gradient rows.
constexpr int kPwB = 32, kPwJ = 8;

// Projection backwards for ALL frames before the loop (model.py:382-388):
//   dY[(t,b)][k] = sum_row dmel[b][t][row] Wp[row][k] + dgate[b][t] wg[k]
// K is only 81, so this is an output-bound pass: block = 64 (t,b) rows x 64 columns, thread = 8 rows x 2 columns.
// The decoder-LSTM columns are written [t][unit][b] and the context columns [t][b][column]: the layouts in which the
// frame loop's kernels read them coalesced.
constexpr int kDyRows = 64, kDyCols = 64;
__global__ void __launch_bounds__(256) bw_dy_all(Params p, Grads g) {
  __shared__ float dm_s[kDyRows][M + 2];
  __shared__ float w_s[M + 1][kDyCols];
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  const int KD = H + p.S * E, R = p.T * p.B;
  const int r0 = blockIdx.x * kDyRows, c0 = blockIdx.y * kDyCols;
  for (int i = tid; i < kDyRows * (M + 1); i += 256) {
    const int rl = i / (M + 1), k = i - rl * (M + 1), r = r0 + rl;
    float v = 0.f;
    if (r < R) {
      const int t = r / p.B, b = r - t * p.B;
      v = k < M ? g.d_mel[((size_t)b * p.T + t) * M + k] : g.d_gate[(size_t)b * p.T + t];
    }
    dm_s[rl][k] = v;
  }
  for (int i = tid; i < (M + 1) * kDyCols; i += 256) {
    const int k = i / kDyCols, c = i - k * kDyCols;
    w_s[k][c] = k < M ? __ldg(p.proj_w + (size_t)k * KD + c0 + c) : __ldg(p.gate_w + c0 + c);
  }
  __syncthreads();
  float acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = 0.f; }
#pragma unroll 3
  for (int k = 0; k <= M; ++k) {
    const float w0 = w_s[k][lane], w1 = w_s[k][lane + 32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = dm_s[rg * 8 + i][k];
      acc[i][0] = fmaf(d, w0, acc[i][0]); acc[i][1] = fmaf(d, w1, acc[i][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + rg * 8 + i;
    if (r >= R) continue;
    const int t = r / p.B, b = r - t * p.B;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = c0 + lane + 32 * q;
      if (c < H) g.dyh[((size_t)t * H + c) * p.B + b] = acc[i][q];
      else g.dyc[((size_t)t * p.B + b) * (p.S * E) + (c - H)] = acc[i][q];
    }
  }
}

// gate gradients of the block -> bf16 operand tile + fp32 rows [.., b, G]
__device__ __forceinline__ void store_gate_grads(const float (*out_s)[kPwB][kPwJ + 1], int NPAD, int B, int b0, int j0,
                                                 unsigned char* tiles, float* rows /* + t*B*G */) {
  const int tid = threadIdx.x;
  if (tid < 4 * kPwB) {
    const int q = tid >> 5, bl = tid & 31, b = b0 + bl;
    if (b < B) {
      const float* v = out_s[q][bl];
      const int k = q * H + j0;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
      uint4 pk;
      pk.x = *reinterpret_cast<unsigned*>(&h0); pk.y = *reinterpret_cast<unsigned*>(&h1);
      pk.z = *reinterpret_cast<unsigned*>(&h2); pk.w = *reinterpret_cast<unsigned*>(&h3);
      *reinterpret_cast<uint4*>(tiles + (size_t)(k >> 6) * ((size_t)NPAD * 128) + tc::tile_offset_bytes(NPAD, b, k & 63)) = pk;
      float4* dst = reinterpret_cast<float4*>(rows + (size_t)b * G + k);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// d prenet[t] = rows [0, P) of dX1[t] (sum of the split-K partials)
__device__ __forceinline__ void save_dpre(const Params& p, const Bufs& bb, const Grads& g, int t, int s, int k, int b) {
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < SPLITSB1; ++q) acc += bb.dx1[(((size_t)s * SPLITSB1 + q) * K1 + k) * bb.NPAD + b];
  g.dpre[(((size_t)s * p.T + t) * p.B + b) * P + k] = acc;
}

// First kernel of a backward frame.  Besides the decoder-LSTM cell it copies the frame index to t_ptr[1] (read by
// bw_pointwise1, which then moves t_ptr[0] to t-1) and saves d prenet[t+1] out of the previous frame's dX1 before this
// frame's GEMM overwrites it, so the frame needs no separate save / counter kernels.
__global__ void __launch_bounds__(256) bw_pointwise2(Params p, Bufs bb, Grads g, int* t_ptr) {
  __shared__ float out_s[4][kPwB][kPwJ + 1];
  const int t = *t_ptr;
  if (blockIdx.x == 0 && threadIdx.x == 0) t_ptr[1] = t;
  const int nbt = (p.B + kPwB - 1) / kPwB;
  const int b0 = (blockIdx.x % nbt) * kPwB, j0 = (blockIdx.x / nbt) * kPwJ;
  const int bl = threadIdx.x & 31, jl = threadIdx.x >> 5, b = b0 + bl, j = j0 + jl;
  if (b < p.B && j < p.S * P && t + 1 < p.T) save_dpre(p, bb, g, t + 1, j / P, j % P, b);
  if (b < p.B) {
    const size_t idx = (size_t)b * H + j;
    float dh = g.dyh[((size_t)t * H + j) * p.B + b];
#pragma unroll
    for (int k = 0; k < SPLITSB2; ++k) dh += bb.dx2[((size_t)k * bb.K2 + p.S * (H + E) + j) * bb.NPAD + b];
    const size_t gs = (size_t)H * p.B;
    const float* sv = g.sv.gates2 + ((size_t)t * 5 * H + j) * p.B + b;
    const float cn_prev = t > 0 ? (sv - 5 * gs)[4 * gs] : 0.f;
    float mh = 1.f, mc = 1.f, mc_prev = 1.f;
    if (p.training) {
      const float sc = 1.0f / (1.0f - p.p_dec);
      const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 4) * p.B * H : nullptr;
      const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 5) * p.B * H : nullptr;
      mh = keep_mult(kh, idx, p.seed, 8, t, (int)idx, p.thresh_dec, sc);
      mc = keep_mult(kc, idx, p.seed, 9, t, (int)idx, p.thresh_dec, sc);
      if (t > 0) {
        const uint8_t* kcp = p.lstm_keep ? p.lstm_keep + ((size_t)(t - 1) * 6 + 5) * p.B * H : nullptr;
        mc_prev = keep_mult(kcp, idx, p.seed, 9, t - 1, (int)idx, p.thresh_dec, sc);
      }
    }
    float dgate[4];
    lstm_cell_backward(sv, gs, mc_prev * cn_prev, mh, mc, dh, g.dc2 + idx, dgate);
#pragma unroll
    for (int q = 0; q < 4; ++q) out_s[q][bl][jl] = dgate[q];
  }
  __syncthreads();
  store_gate_grads(out_s, bb.NPAD, p.B, b0, j0, bb.dg2, g.dg2 + (size_t)t * p.B * G);
}

__global__ void __launch_bounds__(256) bw_pointwise1(Params p, Bufs bb, Grads g, int* t_ptr) {
  __shared__ float dq_s[A][kPwB + 1];
  __shared__ float wq_s[A][kPwJ];
  __shared__ float out_s[4][kPwB][kPwJ + 1];
  const int t = t_ptr[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) t_ptr[0] = t - 1;   // last t-dependent kernel of the frame; nobody reads t_ptr[0] until the next frame
  const int nbt = (p.B + kPwB - 1) / kPwB, njt = H / kPwJ;
  const int s = blockIdx.x / (nbt * njt), rem = blockIdx.x - s * nbt * njt;
  const int b0 = (rem % nbt) * kPwB, j0 = (rem / nbt) * kPwJ;
  const int bl = threadIdx.x & 31, jl = threadIdx.x >> 5, b = b0 + bl, j = j0 + jl;
  // query layer operands of the block: dq of its 32 utterances, Wq columns of its 8 units (attention.py:56, 368)
  for (int i = threadIdx.x; i < kPwB * A; i += blockDim.x) {
    const int bb_ = i / A, a = i - bb_ * A;
    dq_s[a][bb_] = (b0 + bb_ < p.B) ? g.dq[(((size_t)s * p.T + t) * p.B + b0 + bb_) * A + a] : 0.f;
  }
  for (int i = threadIdx.x; i < A * kPwJ; i += blockDim.x) {
    const int a = i / kPwJ, jj = i - a * kPwJ;
    wq_s[a][jj] = __ldg(p.st[s].wq + (size_t)a * H + j0 + jj);
  }
  float dh = 0.f, cn_prev = 0.f;
  const size_t gs = (size_t)H * p.B;
  const size_t idx = (size_t)min(b, p.B - 1) * H + j;
  const float* sv = g.sv.gates1 + (((size_t)t * p.S + s) * 5 * H + j) * p.B + min(b, p.B - 1);
  if (b < p.B) {
#pragma unroll
    for (int k = 0; k < SPLITSB1; ++k) dh += bb.dx1[(((size_t)s * SPLITSB1 + k) * K1 + P + E + j) * bb.NPAD + b];
#pragma unroll
    for (int k = 0; k < SPLITSB2; ++k) dh += bb.dx2[((size_t)k * bb.K2 + s * (H + E) + j) * bb.NPAD + b];
    if (t > 0) cn_prev = (sv - (size_t)p.S * 5 * gs)[4 * gs];
  }
  __syncthreads();
  if (b < p.B) {
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 16
    for (int a = 0; a < A; a += 2) {
      acc0 = fmaf(dq_s[a][bl], wq_s[a][jl], acc0);
      acc1 = fmaf(dq_s[a + 1][bl], wq_s[a + 1][jl], acc1);
    }
    dh += acc0 + acc1;
    float mh = 1.f, mc = 1.f, mc_prev = 1.f;
    if (p.training) {
      const float sc = 1.0f / (1.0f - p.p_att);
      const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * s) * p.B * H : nullptr;
      const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * s + 1) * p.B * H : nullptr;
      mh = keep_mult(kh, idx, p.seed, 4 + 2 * s, t, (int)idx, p.thresh_att, sc);
      mc = keep_mult(kc, idx, p.seed, 5 + 2 * s, t, (int)idx, p.thresh_att, sc);
      if (t > 0) {
        const uint8_t* kcp = p.lstm_keep ? p.lstm_keep + ((size_t)(t - 1) * 6 + 2 * s + 1) * p.B * H : nullptr;
        mc_prev = keep_mult(kcp, idx, p.seed, 5 + 2 * s, t - 1, (int)idx, p.thresh_att, sc);
      }
    }
    float dgate[4];
    lstm_cell_backward(sv, gs, mc_prev * cn_prev, mh, mc, dh, g.dc1 + (size_t)s * p.B * H + idx, dgate);
#pragma unroll
    for (int q = 0; q < 4; ++q) out_s[q][bl][jl] = dgate[q];
  }
  __syncthreads();
  store_gate_grads(out_s, bb.NPAD, p.B, b0, j0, bb.dg1 + (size_t)s * (G / 64) * bb.NPAD * 128,
                   g.dg1 + ((size_t)s * p.T + t) * p.B * G);
}

// stepwise monotonic attention backwards, one CTA per (utterance, stream)
constexpr int kBwThreads = 512;
__host__ __device__ inline size_t bw_attention_smem_floats(int Ts) { return (size_t)E + 4 * A + 4 * (size_t)(Ts + 4) + M + 4; }

__global__ void __launch_bounds__(kBwThreads, 1) bw_attention(Params p, Bufs bb, Grads g, const int* t_ptr) {
  extern __shared__ __align__(16) float sm[];
  const int t = *t_ptr;
  const int s = blockIdx.x % p.S, b = blockIdx.x / p.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kW = kBwThreads / 32;
  const StreamParams& sp = p.st[s];
  const int Ts = sp.Ts;
  const int len = sp.len ? (int)sp.len[b] : Ts;
  float* dctx_s = sm;                 // E
  float* q_s = dctx_s + E;            // A
  float* v_s = q_s + A;               // A
  float* dq_s = v_s + A;              // A
  float* dv_s = dq_s + A;             // A
  float* p_s = dv_s + A;              // Ts+4
  float* ap_s = p_s + Ts + 4;         // Ts+4   alignment state entering frame t
  float* dan_s = ap_s + Ts + 4;       // Ts+4   d alpha'[t]
  float* de_s = dan_s + Ts + 4;       // Ts+4

  // ---- early requests: nothing below depends on the incoming gradients, so these loads overlap the d ctx chain ----
  const float* mem_b = sp.mem + (size_t)b * Ts * E;
  const float* pm_b = sp.pm + (size_t)b * Ts * A;
  float* dpm_b = g.dpm[s] + (size_t)b * Ts * A;
  float4 mrow[2][4];
  float pmr[2][4], dpr[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = warp + kW * r;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mrow[r][q] = j < Ts ? __ldg(reinterpret_cast<const float4*>(mem_b + (size_t)j * E) + lane + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      pmr[r][q] = j < len ? __ldg(pm_b + (size_t)j * A + lane + 32 * q) : 0.f;
      dpr[r][q] = j < len ? dpm_b[(size_t)j * A + lane + 32 * q] : 0.f;
    }
  }
  for (int a = tid; a < A; a += kBwThreads) {
    q_s[a] = g.sv.q[(((size_t)t * p.S + s) * p.B + b) * A + a];
    v_s[a] = sp.v[a];
    dq_s[a] = 0.f;
    dv_s[a] = 0.f;
  }
  for (int j = tid; j < Ts; j += kBwThreads) {
    p_s[j] = g.p_saved[s][((size_t)t * p.B + b) * Ts + j];
    ap_s[j] = t > 0 ? g.align[s][((size_t)b * p.T + (t - 1)) * Ts + j] : (j == 0 ? 1.f : 0.f);   // attention.py:324-328
    float d = g.dalpha[s][(size_t)b * Ts + j];                      // carry from frame t+1
    if (g.d_align[s]) d += g.d_align[s][((size_t)b * p.T + t) * Ts + j];   // external alignment gradient
    dan_s[j] = d;
  }
  __syncthreads();
  // d ctx[t] = next frame's attention-LSTM input gradient + this frame's decoder-LSTM input and projection gradients
  for (int d = tid; d < E; d += kBwThreads) {
    float acc = g.dyc[((size_t)t * p.B + b) * (p.S * E) + s * E + d];
#pragma unroll
    for (int k = 0; k < SPLITSB1; ++k) acc += bb.dx1[(((size_t)s * SPLITSB1 + k) * K1 + P + d) * bb.NPAD + b];
#pragma unroll
    for (int k = 0; k < SPLITSB2; ++k) acc += bb.dx2[((size_t)k * bb.K2 + s * (H + E) + H + d) * bb.NPAD + b];
    dctx_s[d] = acc;
    g.dctx[(((size_t)s * p.T + t) * p.B + b) * E + d] = acc;
  }
  __syncthreads();

  // d alpha'_j += d ctx . memory_j   (attention.py:395); one warp per position, two positions in flight
  {
    const float4* dc4 = reinterpret_cast<const float4*>(dctx_s);
    auto dot_row = [&](int j, const float4 (&m)[4]) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 dv = dc4[lane + 32 * q];
        acc = fmaf(m[q].x, dv.x, acc); acc = fmaf(m[q].y, dv.y, acc); acc = fmaf(m[q].z, dv.z, acc); acc = fmaf(m[q].w, dv.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) dan_s[j] += acc;
    };
    if (warp < Ts) dot_row(warp, mrow[0]);
    if (warp + kW < Ts) dot_row(warp + kW, mrow[1]);
    for (int j = warp + 2 * kW; j < Ts; j += 2 * kW) {
      const int j1 = j + kW;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        mrow[0][q] = __ldg(reinterpret_cast<const float4*>(mem_b + (size_t)j * E) + lane + 32 * q);
        mrow[1][q] = j1 < Ts ? __ldg(reinterpret_cast<const float4*>(mem_b + (size_t)j1 * E) + lane + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      dot_row(j, mrow[0]);
      if (j1 < Ts) dot_row(j1, mrow[1]);
    }
  }
  __syncthreads();
  // alpha'_j = alpha_j p_j + alpha_{j-1} (1 - p_{j-1})   (attention.py:330-338), p = sigmoid(e)
  if (p.independent) {   // alpha'_j was forced to 0 beyond the utterance's length: no gradient flows through it
    for (int j = len + tid; j < Ts; j += kBwThreads) dan_s[j] = 0.f;
    __syncthreads();
  }
  for (int j = tid; j < Ts; j += kBwThreads) {
    const float dn = dan_s[j], dn1 = j + 1 < Ts ? dan_s[j + 1] : 0.f, pj = p_s[j];
    g.dalpha[s][(size_t)b * Ts + j] = dn * pj + dn1 * (1.0f - pj);
    de_s[j] = ap_s[j] * (dn - dn1) * pj * (1.0f - pj);
  }
  __syncthreads();
  // e_j = v . tanh(q + pm_j)   (attention.py:340-345); positions >= len have p = 0, hence de = 0
  {
    float dq_acc[4] = {0.f, 0.f, 0.f, 0.f}, dv_acc[4] = {0.f, 0.f, 0.f, 0.f};
    float qv[4], vv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { qv[q] = q_s[lane + 32 * q]; vv[q] = v_s[lane + 32 * q]; }
    auto energy_bw = [&](int j, const float (&pm)[4], const float (&dp)[4]) {
      const float de = de_s[j];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float u = lat::fast_tanh(qv[q] + pm[q]);
        const float dz = de * vv[q] * (1.0f - u * u);
        dq_acc[q] += dz;
        dv_acc[q] = fmaf(de, u, dv_acc[q]);
        dpm_b[(size_t)j * A + lane + 32 * q] = dp[q] + dz;
      }
    };
    if (warp < len) energy_bw(warp, pmr[0], dpr[0]);
    if (warp + kW < len) energy_bw(warp + kW, pmr[1], dpr[1]);
    for (int j = warp + 2 * kW; j < len; j += 2 * kW) {
      const int j1 = j + kW;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pmr[0][q] = __ldg(pm_b + (size_t)j * A + lane + 32 * q);
        dpr[0][q] = dpm_b[(size_t)j * A + lane + 32 * q];
        pmr[1][q] = j1 < len ? __ldg(pm_b + (size_t)j1 * A + lane + 32 * q) : 0.f;
        dpr[1][q] = j1 < len ? dpm_b[(size_t)j1 * A + lane + 32 * q] : 0.f;
      }
      energy_bw(j, pmr[0], dpr[0]);
      if (j1 < len) energy_bw(j1, pmr[1], dpr[1]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(&dq_s[lane + 32 * q], dq_acc[q]);
      atomicAdd(&dv_s[lane + 32 * q], dv_acc[q]);
    }
  }
  __syncthreads();
  for (int a = tid; a < A; a += kBwThreads) {
    g.dq[(((size_t)s * p.T + t) * p.B + b) * A + a] = dq_s[a];
    g.dv[((size_t)s * p.B + b) * A + a] += dv_s[a];
  }
}

// location-sensitive attention backwards (attention.py:7-85), one CTA per (utterance, stream):
//   feat = conv1d(cat(alpha[t-1], cum[t-1]))  ->  loc = dense(feat)  ->  e = v . tanh(q + pm + loc)  ->  alpha = softmax(e)
//   ctx = alpha . memory,  cum[t] = cum[t-1] + alpha[t]                                            (model.py:358-359)
// alpha[t] receives gradient from the context, from frame t+1's convolution input (carry dalpha) and, through the
// cumulative weights, from the convolution inputs of every later frame (carry dcum, a running sum).
constexpr int kLF = 32, kLK = 31, kLPad = 15;
__host__ __device__ inline size_t bw_lsa_smem_floats(int Ts) {
  return (size_t)E + 4 * A + (size_t)kLF * A + 2 * kLK * kLF + 4 * (size_t)(Ts + 4) + 2 * (size_t)(Ts + 2 * kLPad + 2) +
         2 * (size_t)Ts * kLF + (kBwThreads / 32) * A + 64;
}
__global__ void __launch_bounds__(kBwThreads, 1) bw_attention_lsa(Params p, Bufs bb, Grads g, const int* t_ptr) {
  extern __shared__ __align__(16) float sm[];
  const int t = *t_ptr;
  const int s = blockIdx.x % p.S, b = blockIdx.x / p.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kW = kBwThreads / 32;
  const StreamParams& sp = p.st[s];
  const int Ts = sp.Ts, Tp = Ts + 2 * kLPad;
  const int len = sp.len ? (int)sp.len[b] : Ts;
  float* dctx_s = sm;                       // E
  float* q_s = dctx_s + E;                  // A
  float* v_s = q_s + A;                     // A
  float* dq_s = v_s + A;                    // A
  float* dv_s = dq_s + A;                   // A
  float* wd_s = dv_s + A;                   // LF*A  location_dense^T [f][a]
  float* wc_s = wd_s + kLF * A;             // 2*LK*LF  location_conv [c][k][f]
  float* al_s = wc_s + 2 * kLK * kLF;       // Ts+4  alpha[t]
  float* dan_s = al_s + Ts + 4;             // Ts+4  d alpha[t]
  float* de_s = dan_s + Ts + 4;             // Ts+4
  float* red_s = de_s + Ts + 4;             // Ts+4 (only 32 used)
  float* in0_s = red_s + Ts + 4;            // Tp+2  padded alpha[t-1]
  float* in1_s = in0_s + Tp + 2;            // Tp+2  padded cum[t-1]
  float* feat_s = in1_s + Tp + 2;           // Ts*LF
  float* dfeat_s = feat_s + (size_t)Ts * kLF;   // Ts*LF
  float* dzw_s = dfeat_s + (size_t)Ts * kLF;    // kW*A per-warp staging of dz

  for (int d = tid; d < E; d += kBwThreads) {
    float acc = g.dyc[((size_t)t * p.B + b) * (p.S * E) + s * E + d];
#pragma unroll
    for (int k = 0; k < SPLITSB1; ++k) acc += bb.dx1[(((size_t)s * SPLITSB1 + k) * K1 + P + d) * bb.NPAD + b];
#pragma unroll
    for (int k = 0; k < SPLITSB2; ++k) acc += bb.dx2[((size_t)k * bb.K2 + s * (H + E) + H + d) * bb.NPAD + b];
    dctx_s[d] = acc;
    g.dctx[(((size_t)s * p.T + t) * p.B + b) * E + d] = acc;
  }
  for (int a = tid; a < A; a += kBwThreads) {
    q_s[a] = g.sv.q[(((size_t)t * p.S + s) * p.B + b) * A + a];
    v_s[a] = sp.v[a];
    dq_s[a] = 0.f;
    dv_s[a] = 0.f;
  }
  for (int i = tid; i < kLF * A; i += kBwThreads) { const int f = i / A, a = i - f * A; wd_s[i] = sp.loc_dense[(size_t)a * kLF + f]; }
  for (int i = tid; i < 2 * kLK * kLF; i += kBwThreads) { const int f = i % kLF, ck = i / kLF; wc_s[i] = sp.loc_conv[(size_t)f * 2 * kLK + ck]; }
  for (int i = tid; i < Tp; i += kBwThreads) {
    const int j = i - kLPad;
    const bool in = j >= 0 && j < Ts;
    in0_s[i] = (in && t > 0) ? g.align[s][((size_t)b * p.T + (t - 1)) * Ts + j] : 0.f;
    in1_s[i] = in ? g.p_saved[s][((size_t)t * p.B + b) * Ts + j] : 0.f;
  }
  for (int j = tid; j < Ts; j += kBwThreads) {
    al_s[j] = g.align[s][((size_t)b * p.T + t) * Ts + j];
    float d = g.dalpha[s][(size_t)b * Ts + j] + g.dcum[s][(size_t)b * Ts + j];
    if (g.d_align[s]) d += g.d_align[s][((size_t)b * p.T + t) * Ts + j];
    dan_s[j] = d;
  }
  __syncthreads();
  // d alpha_j += d ctx . memory_j
  {
    const float* mem_b = sp.mem + (size_t)b * Ts * E;
    const float4* dc4 = reinterpret_cast<const float4*>(dctx_s);
    for (int j = warp; j < Ts; j += kW) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mem_b + (size_t)j * E) + lane + 32 * q);
        const float4 dv = dc4[lane + 32 * q];
        acc = fmaf(m.x, dv.x, acc); acc = fmaf(m.y, dv.y, acc); acc = fmaf(m.z, dv.z, acc); acc = fmaf(m.w, dv.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) dan_s[j] += acc;
    }
  }
  // location features of this frame (recomputed): feat[j][f] = sum_k Wc[f][0][k] a_prev[j+k-15] + Wc[f][1][k] cum[j+k-15]
  for (int i = tid; i < Ts * kLF; i += kBwThreads) {
    const int j = i / kLF, f = i - j * kLF;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kLK; ++k) {
      acc = fmaf(wc_s[k * kLF + f], in0_s[j + k], acc);
      acc = fmaf(wc_s[(kLK + k) * kLF + f], in1_s[j + k], acc);
    }
    feat_s[i] = acc;
    dfeat_s[i] = 0.f;
  }
  __syncthreads();
  // softmax backwards: de_j = alpha_j (d alpha_j - sum_k alpha_k d alpha_k)
  {
    float part = 0.f;
    for (int j = tid; j < Ts; j += kBwThreads) part = fmaf(al_s[j], dan_s[j], part);
    part = warp_sum(part);
    if (lane == 0) red_s[warp] = part;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kW; ++w) tot += red_s[w];
    for (int j = tid; j < Ts; j += kBwThreads) de_s[j] = al_s[j] * (dan_s[j] - tot);
  }
  __syncthreads();
  // energies backwards, one warp per position
  {
    const float* pm_b = sp.pm + (size_t)b * Ts * A;
    float* dpm_b = g.dpm[s] + (size_t)b * Ts * A;
    float* dz_b = g.dz[s] + (size_t)b * Ts * A;
    float* dzw = dzw_s + warp * A;
    float dq_acc[4] = {0.f, 0.f, 0.f, 0.f}, dv_acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = warp; j < len; j += kW) {
      const float de = de_s[j];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int a = lane + 32 * q;
        float loc = 0.f;
#pragma unroll 8
        for (int f = 0; f < kLF; ++f) loc = fmaf(wd_s[f * A + a], feat_s[j * kLF + f], loc);
        const float u = tanhf(q_s[a] + __ldg(pm_b + (size_t)j * A + a) + loc);
        const float dz = de * v_s[a] * (1.0f - u * u);
        dq_acc[q] += dz;
        dv_acc[q] = fmaf(de, u, dv_acc[q]);
        dpm_b[(size_t)j * A + a] += dz;
        dz_b[(size_t)j * A + a] = dz;
        dzw[a] = dz;
      }
      __syncwarp();
      {  // d feat[j][f] = sum_a dz[a] Wd[a][f]   (lane = f)
        float acc = 0.f;
#pragma unroll 8
        for (int a = 0; a < A; ++a) acc = fmaf(dzw[a], wd_s[lane * A + a], acc);
        dfeat_s[j * kLF + lane] = acc;
      }
      __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(&dq_s[lane + 32 * q], dq_acc[q]);
      atomicAdd(&dv_s[lane + 32 * q], dv_acc[q]);
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int a = tid; a < A; a += kBwThreads) {
    g.dq[(((size_t)s * p.T + t) * p.B + b) * A + a] = dq_s[a];
    g.dv[((size_t)s * p.B + b) * A + a] += dv_s[a];
  }
  // location_dense weight gradient: dWd[a][f] += sum_j dz[j][a] feat[j][f]   (thread = a, 8 filters)
  {
    const int a = tid & (A - 1), f0 = (tid / A) * 8;
    const float* dz_b = g.dz[s] + (size_t)b * Ts * A;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < len; ++j) {
      const float dz = dz_b[(size_t)j * A + a];
#pragma unroll
      for (int f = 0; f < 8; ++f) acc[f] = fmaf(dz, feat_s[j * kLF + f0 + f], acc[f]);
    }
    float* dst = g.dwd + ((size_t)s * p.B + b) * (A * kLF) + (size_t)a * kLF + f0;
#pragma unroll
    for (int f = 0; f < 8; ++f) dst[f] += acc[f];
  }
  // convolution backwards: inputs (-> carries for frame t-1) and weights
  for (int i = tid; i < 2 * Ts; i += kBwThreads) {
    const int c = i / Ts, pos = i - c * Ts;
    float acc = 0.f;
    for (int k = 0; k < kLK; ++k) {
      const int j = pos - k + kLPad;                 // output position that read input `pos` through tap k
      if (j < 0 || j >= Ts) continue;
      const float* df = dfeat_s + j * kLF;
      const float* w = wc_s + (c * kLK + k) * kLF;
#pragma unroll 8
      for (int f = 0; f < kLF; ++f) acc = fmaf(df[f], w[f], acc);
    }
    if (c == 0) g.dalpha[s][(size_t)b * Ts + pos] = acc;            // d alpha[t-1] through the "previous weights" channel
    else g.dcum[s][(size_t)b * Ts + pos] += acc;                    // running sum over all later frames
  }
  for (int i = tid; i < kLF * 2 * kLK; i += kBwThreads) {
    const int f = i / (2 * kLK), ck = i - f * 2 * kLK, c = ck / kLK, k = ck - c * kLK;
    const float* in = c == 0 ? in0_s : in1_s;
    float acc = 0.f;
    for (int j = 0; j < len; ++j) acc = fmaf(dfeat_s[j * kLF + f], in[j + k], acc);
    g.dwc[((size_t)s * p.B + b) * (kLF * 2 * kLK) + i] += acc;
  }
}

// after the loop: d prenet[0] is still in dX1
__global__ void bw_save_dpre0(Params p, Bufs bb, Grads g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.S * p.B * P) return;
  const int s = i / (p.B * P), r = i - s * p.B * P, k = r / p.B, b = r - k * p.B;
  save_dpre(p, bb, g, 0, s, k, b);
}

}  // namespace bw
