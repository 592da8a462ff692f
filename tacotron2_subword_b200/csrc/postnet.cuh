// postnet.cuh -- eval-mode Postnet on tcgen05, included by taco2dec.cu.
//
// Reference: /root/reference/model.py:27-70 (five Conv1d(k=5, 'same') + BatchNorm1d, tanh on all but the last; dropout is
// the identity in eval), model.py:557-558 (mel_postnet = mel + postnet(mel)) and model.py:531-541 (zero beyond
// output_lengths).  Each layer is one dense contraction over (tap, input channel):
//   out[co][n] = sum_{kk, ci} W'[co][kk * Cin + ci] . X[n][kk * Cin + ci],   X[n][kk * Cin + ci] = act[n + kk - 2][ci]
// with BatchNorm folded into W' and the bias, and runs on the same tcgen05 GEMM as the LSTM gates (gemm_tc.cuh):
// M = output channels (128 per CTA), N = 128 positions per CTA, K = 5 * Cin, fp16 operands, fp32 accumulation in TMEM.
// The im2col operand X is never materialised by a separate pass: the pointwise kernel of layer l (bias + tanh) writes
// its activations straight into the five tap slots of layer l+1's operand tiles, destination-driven so that every slot
// -- including the zero padding at utterance boundaries -- is written exactly once (no memset, no stale data).
#pragma once

namespace pn {

constexpr int kTaps = 5, kHalo = 2, kNP = 128;   // kernel size, (k-1)/2, positions per GEMM tile
constexpr int kMaxLayers = 8;

struct Layer {
  int cin, cout, cin_pad, cout_pad, K;           // K = kTaps * cin_pad
  unsigned char* a_tiles;                        // [cout_pad/128][K/64] fp16 tiles, BatchNorm folded
  float* bias;                                   // [cout_pad] folded bias
};

// Precision: segs = 1 -- plain fp16 operands (relative error ~8e-4 of the output scale after five layers); segs = 3 -- every
// fp32 operand is split into two fp16 terms (hi + lo) and the product is evaluated as hi.hi + hi.lo + lo.hi: K segments
// [W_hi | W_hi | W_lo] x [X_hi | X_lo | X_hi], fp32 accumulation -> fp32-grade mel_postnet (the default; the tensor the
// vocoder consumes is held to the same 1e-3 bar as the decoder mel).
__device__ __forceinline__ void pn_split(float v, __half& hi, __half& lo) {
  hi = __float2half(v);
  lo = __float2half(v - __half2float(hi));
}

// conv weight [cout][cin][taps] + BatchNorm (eval) -> fp16 GEMM tiles with K = tap * cin_pad + ci, folded bias
__global__ void pn_pack_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, int cout, int cin, int cout_pad, int cin_pad, unsigned char* __restrict__ tiles,
                               float* __restrict__ bias_out, int segs) {
  const int K = kTaps * cin_pad, kb_total = segs * K / tc::kBlockK;
  const size_t total = (size_t)cout_pad * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), ro = (int)(i / K);
    const int kk = k / cin_pad, ci = k - kk * cin_pad;
    float v = 0.f;
    if (ro < cout && ci < cin) v = w[((size_t)ro * cin + ci) * kTaps + kk] * (gamma[ro] * rsqrtf(var[ro] + eps));
    __half hi, lo;
    pn_split(v, hi, lo);
    for (int seg = 0; seg < segs; ++seg) {
      const int kk = seg * K + k;
      const size_t tile = ((size_t)(ro / 128) * kb_total + (kk / tc::kBlockK)) * tc::kATileBytes;
      *reinterpret_cast<__half*>(tiles + tile + tc::tile_offset_bytes(128, ro % 128, kk % tc::kBlockK)) = seg == 2 ? lo : hi;
    }
    if (k == 0) {
      float bo = 0.f;
      if (ro < cout) { const float sc = gamma[ro] * rsqrtf(var[ro] + eps); bo = (b[ro] - mean[ro]) * sc + beta[ro]; }
      bias_out[ro] = bo;
    }
  }
}

// 16 bytes = 8 consecutive channels of tap kk at position n of the operand (group = n / 128)
__device__ __forceinline__ void x_chunk_store(unsigned char* X, int K, int n, int k8, const float* v, int segs) {
  const int g = n / kNP, nl = n - g * kNP, k = k8 * 8;
  __align__(16) __half hi[8], lo[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) pn_split(v[q], hi[q], lo[q]);
  for (int seg = 0; seg < segs; ++seg) {      // operand segments [hi | lo | hi]
    const int kk = seg * K + k;
    unsigned char* tile = X + ((size_t)g * (segs * K / tc::kBlockK) + (kk >> 6)) * ((size_t)kNP * 128);
    *reinterpret_cast<uint4*>(tile + tc::tile_offset_bytes(kNP, nl, kk & 63)) = *reinterpret_cast<const uint4*>(seg == 1 ? lo : hi);
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 r;
  r.x = *reinterpret_cast<unsigned*>(&h0); r.y = *reinterpret_cast<unsigned*>(&h1);
  r.z = *reinterpret_cast<unsigned*>(&h2); r.w = *reinterpret_cast<unsigned*>(&h3);
  return r;
}

// layer-0 operand from the mel input (arbitrary strides: the decoder's storage is [B, T, n_mel]).  One thread per
// destination chunk (position n', tap kk, 8 channels): source frame t = t' + kk - 2 of the same utterance, else zeros.
// `seq_len` (independent utterances): frames >= seq_len[b] do not exist -- they read as the convolution's zero padding in
// every layer, so row b equals the batch-1 result on its first seq_len[b] frames.
__global__ void pn_input_kernel(const float* __restrict__ mel, long long sb, long long sc, long long st, int B, int T,
                                int cin, int cin_pad, int n_pad /* groups * 128 */, const long long* __restrict__ seq_len,
                                unsigned char* __restrict__ X, int segs) {
  const int chunks = cin_pad / 8, K = kTaps * cin_pad;
  const size_t total = (size_t)n_pad * kTaps * chunks;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % chunks), kk = (int)((i / chunks) % kTaps), n = (int)(i / ((size_t)chunks * kTaps));
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < B * T) {
      const int b = n / T, t = n - b * T + kk - kHalo;
      const int Tb = seq_len ? min((int)seq_len[b], T) : T;
      if (t >= 0 && t < Tb) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int c = c8 * 8 + q; if (c < cin) v[q] = mel[b * sb + c * sc + t * st]; }
      }
    }
    x_chunk_store(X, K, n, kk * chunks + c8, v, segs);
  }
}

// bias + tanh of one layer's GEMM output, written into the next layer's operand (block = one 128-position tile x 8
// output channels; the tile's activations plus a 2-position halo from the neighbouring tiles are staged in shared
// memory, then every (position, tap) destination gets its 8 channels as one 16-byte store).
__global__ void __launch_bounds__(256) pn_pointwise_kernel(const float* __restrict__ part, int splits, int cout_pad,
                                                           const float* __restrict__ bias, int B, int T, int groups,
                                                           const long long* __restrict__ seq_len,
                                                           unsigned char* __restrict__ Xn, int Kn /* 5 * cout_pad */, int segs) {
  __shared__ float act_s[8][kNP + 2 * kHalo + 1];
  const int g = blockIdx.x, c0 = blockIdx.y * 8, tid = threadIdx.x;
  const int NT = B * T, n0 = g * kNP;
  for (int i = tid; i < 8 * (kNP + 2 * kHalo); i += 256) {
    const int cl = i / (kNP + 2 * kHalo), pl = i - cl * (kNP + 2 * kHalo);
    const int n = n0 - kHalo + pl;
    float a = 0.f;
    bool exists = n >= 0 && n < NT;
    if (exists && seq_len) { const int b = n / T; exists = (n - b * T) < (int)seq_len[b]; }
    if (exists) {
      const int gg = n / kNP, nl = n - gg * kNP;
      float acc = bias[c0 + cl];
      for (int k = 0; k < splits; ++k) acc += part[(((size_t)gg * splits + k) * cout_pad + c0 + cl) * kNP + nl];
      a = tanhf(acc);
    }
    act_s[cl][pl] = a;
  }
  __syncthreads();
  for (int i = tid; i < kNP * kTaps; i += 256) {
    const int kk = i / kNP, nl = i - kk * kNP, n = n0 + nl;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < NT) {
      const int t = n % T + kk - kHalo;                 // source frame inside the same utterance?
      if (t >= 0 && t < T) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = act_s[q][nl + kk];
      }
    }
    x_chunk_store(Xn, Kn, n, kk * (cout_pad / 8) + blockIdx.y, v, segs);
  }
  (void)groups;
}

// last layer: mel_postnet[b][c][t] = mel + conv + bias, zero beyond the utterance's length (model.py:531-541, 557-558)
__global__ void pn_output_kernel(const float* __restrict__ part, int splits, int cout_pad, const float* __restrict__ bias,
                                 const float* __restrict__ mel, long long sb, long long sc, long long st, int B, int T, int cout,
                                 const long long* __restrict__ lengths, float* __restrict__ out) {
  const size_t total = (size_t)B * cout * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T), c = (int)((i / T) % cout), b = (int)(i / ((size_t)T * cout));
    const int n = b * T + t, g = n / kNP, nl = n - g * kNP;
    float acc = bias[c];
    for (int k = 0; k < splits; ++k) acc += part[(((size_t)g * splits + k) * cout_pad + c) * kNP + nl];
    float y = mel[b * sb + c * sc + t * st] + acc;
    if (lengths && t >= (int)lengths[b]) y = 0.f;
    out[i] = y;
  }
}

}  // namespace pn
