// loss.cuh -- Tacotron2Loss and the gradients that seed the backward pass, in one sweep, included by
// taco2dec.cu.
//
// Reference: /root/reference/loss_function.py:12-66
//   mel_loss  = MSE(mel_out, mel_target) + MSE(mel_out_postnet, mel_target)           (means over all B * n_mel * T elements)
//   gate_loss = BCEWithLogits(gate_out.view(-1, 1), gate_target.view(-1, 1))          (mean over B * T)
//   alignloss == "L2" (iterations < 40000): + MSE(align_out, align_target) + MSE(align_bert_out, align_target)
// The reference makes one pass per term for the forward value and autograd makes another one per term backwards; here a single
// pass reads every tensor once and writes the loss terms AND d loss / d output for each output, the mel gradient directly
// in the [B, T, n_mel] storage order the decoder's BPTT consumes (taco2dec_bwd_args.d_mel).  Sums are reduced in a fixed
// order (per-block partials, then one block in double precision): bit-reproducible run to run.
#pragma once

namespace ls {

constexpr int kTT = 32;      // frames per block of the mel kernel

__device__ __forceinline__ float block_sum(float v, float* red_s) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red_s[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red_s[w];
  return t;      // valid in thread 0
}

// block = (utterance b, tile of kTT frames): mel in its own strides (the decoder writes [B, T, n_mel]), post / target [B, n_mel, T]
__global__ void __launch_bounds__(256) ls_mel_kernel(const float* __restrict__ mel, long long sb, long long sc, long long st,
                                                     const float* __restrict__ post, const float* __restrict__ target, int B, int C,
                                                     int T, float inv_n, float* __restrict__ d_mel, float* __restrict__ d_post,
                                                     float* __restrict__ part /* [blocks][2] */) {
  extern __shared__ float tile_s[];          // [C][kTT + 1] d_mel tile, transposed on the way out
  __shared__ float red_s[8];
  const int tiles = (T + kTT - 1) / kTT;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x - b * tiles) * kTT;
  float s_mel = 0.f, s_post = 0.f;
  for (int i = threadIdx.x; i < C * kTT; i += blockDim.x) {
    const int c = i / kTT, tl = i - c * kTT, t = t0 + tl;
    float gm = 0.f;
    if (t < T) {
      const size_t o = ((size_t)b * C + c) * T + t;
      const float y = target[o];
      const float em = mel[b * sb + c * sc + t * st] - y, ep = post[o] - y;
      s_mel = fmaf(em, em, s_mel);
      s_post = fmaf(ep, ep, s_post);
      gm = 2.0f * em * inv_n;
      d_post[o] = 2.0f * ep * inv_n;
    }
    tile_s[c * (kTT + 1) + tl] = gm;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * kTT; i += blockDim.x) {       // [B, T, C]: channels contiguous
    const int tl = i / C, c = i - tl * C, t = t0 + tl;
    if (t < T) d_mel[((size_t)b * T + t) * C + c] = tile_s[c * (kTT + 1) + tl];
  }
  const float a = block_sum(s_mel, red_s);
  const float p2 = block_sum(s_post, red_s);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = p2; }
}

// BCE with logits, numerically stable form: max(x, 0) - x y + log(1 + exp(-|x|))
__global__ void __launch_bounds__(256) ls_gate_kernel(const float* __restrict__ gate, const float* __restrict__ target, int n, float inv_n,
                                                      float* __restrict__ d_gate, float* __restrict__ part /* [blocks] */) {
  __shared__ float red_s[8];
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = gate[i], y = target[i];
    s += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
    d_gate[i] = (1.0f / (1.0f + expf(-x)) - y) * inv_n;
  }
  const float t = block_sum(s, red_s);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) ls_mse_kernel(const float* __restrict__ x, const float* __restrict__ target, size_t n, float inv_n,
                                                     float* __restrict__ dx, float* __restrict__ part) {
  __shared__ float red_s[8];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float e = x[i] - target[i];
    s = fmaf(e, e, s);
    dx[i] = 2.0f * e * inv_n;
  }
  const float t = block_sum(s, red_s);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// one block: fixed-order sums of the partials in double precision -> losses[0..5] = total, mel, gate, align, align_bert
struct ReduceArgs {
  const float *mel_part, *gate_part, *al_part, *alb_part;
  int n_mel_blocks, n_gate_blocks, n_al_blocks, n_alb_blocks;
  double inv_mel, inv_gate, inv_al, inv_alb;
  float* losses;
};
__global__ void __launch_bounds__(256) ls_reduce_kernel(ReduceArgs a) {
  __shared__ double acc_s[5][256];
  double v[5] = {0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < a.n_mel_blocks; i += 256) { v[0] += a.mel_part[2 * i]; v[1] += a.mel_part[2 * i + 1]; }
  for (int i = threadIdx.x; i < a.n_gate_blocks; i += 256) v[2] += a.gate_part[i];
  for (int i = threadIdx.x; i < a.n_al_blocks; i += 256) v[3] += a.al_part[i];
  for (int i = threadIdx.x; i < a.n_alb_blocks; i += 256) v[4] += a.alb_part[i];
  for (int k = 0; k < 5; ++k) acc_s[k][threadIdx.x] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 5; ++k)
      for (int i = 0; i < 256; ++i) t[k] += acc_s[k][i];
    const double mel_loss = t[0] * a.inv_mel + t[1] * a.inv_mel, gate_loss = t[2] * a.inv_gate;
    const double al = t[3] * a.inv_al, alb = t[4] * a.inv_alb;
    a.losses[0] = (float)(mel_loss + gate_loss + al + alb);
    a.losses[1] = (float)mel_loss; a.losses[2] = (float)gate_loss; a.losses[3] = (float)al; a.losses[4] = (float)alb;
  }
}

}  // namespace ls
