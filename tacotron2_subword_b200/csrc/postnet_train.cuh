// postnet_train.cuh -- training-mode Postnet (forward AND backward) on tcgen05, included by taco2dec.cu after wgrad.cuh.
//
// Reference: /root/reference/model.py:27-70 in model.train(): five Conv1d(k=5, 'same') + BatchNorm1d with BATCH statistics,
// tanh on all but the last layer, F.dropout(p=0.5) after every layer; gradients from autograd.
//
// Activations live channel-last with a two-frame zero halo per utterance, x_pad[b][t + 2][c] (fp32).  In that layout the
// im2col row of frame (b, t) -- the five taps of every input channel -- is ONE contiguous run of 5*C floats starting at
// x_pad[b][t][0], so every contraction of the layer is a plain "rows x weights^T" product over overlapping rows:
//   forward   y[n][co]  = sum_{k,ci} x_pad[n + k][ci] . W[co][ci][k]                 (rows of x_pad,  W'  [co][(k, ci)])
//   d input   dx[n][ci] = sum_{k,co} dy_pad[n + k][co] . W[co][ci][4 - k]            (rows of dy_pad, W'' [ci][(k, co)])
//   d weight  dW[co][(k,ci)] = sum_n dy[n][co] . x_pad[n + k][ci]                    (wgrad.cuh over the same overlapping rows)
// The first two run through pt_pack_rows_kernel -> tc::gemm_f16_tn_kernel -> pt_finish_kernel (fp16 operands, fp32
// accumulation in TMEM; gradient rows are scaled by a power of two taken from their absolute maximum so that they survive
// fp16, exactly as in wgrad.cuh); the finishing kernel also accumulates the per-channel sums BatchNorm needs.  Everything
// between the contractions -- normalisation, tanh, dropout and their derivatives -- is three fused element-wise kernels.
#pragma once

namespace pt {

constexpr int kNP = 128;

// fp32 rows -> fp16 operand tiles [group][Kpad/64][128 x 64].  Row n = (b, t) starts at src + b*sb + t*st and has K contiguous
// elements (rows may overlap); rows >= B*T and columns >= K read as zero.  Optional power-of-two scale.
__global__ void __launch_bounds__(256) pt_pack_rows_kernel(const float* __restrict__ src, long long sb, long long st, int B, int T, int K,
                                                           int Kpad, const float* __restrict__ scale, unsigned char* __restrict__ dst) {
  const int kb = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const float sc = scale ? scale[0] : 1.0f;
  const bool vec = (sb & 3) == 0 && (st & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  unsigned char* tile = dst + ((size_t)g * (Kpad / 64) + kb) * tc::kATileBytes;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = tid + it * 256, r = i >> 3, k8 = i & 7;
    const int n = g * kNP + r, k = kb * 64 + k8 * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < B * T && k < K) {
      const int b = n / T, t = n - b * T;
      const float* p = src + b * sb + t * st + k;
      if (vec && k + 8 <= K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) if (k + q < K) v[q] = __ldg(p + q);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] *= sc;
    }
    *reinterpret_cast<uint4*>(tile + tc::tile_offset_bytes(kNP, r, k8 * 8)) = pn::pack8(v);
  }
}

// weights [M][K] fp32 row-major -> A tiles [Mpad/128][Kpad/64], zero padded
__global__ void pt_pack_w_kernel(const float* __restrict__ w, int M, int K, int Mpad, int Kpad, unsigned char* __restrict__ dst) {
  const size_t total = (size_t)Mpad * (Kpad / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k8 = (int)(i % (Kpad / 8)), r = (int)(i / (Kpad / 8)), k = k8 * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < M)
#pragma unroll
      for (int q = 0; q < 8; ++q) if (k + q < K) v[q] = __ldg(w + (size_t)r * K + k + q);
    unsigned char* tile = dst + ((size_t)(r / 128) * (Kpad / 64) + (k >> 6)) * tc::kATileBytes;
    *reinterpret_cast<uint4*>(tile + tc::tile_offset_bytes(128, r % 128, k & 63)) = pn::pack8(v);
  }
}

// partials [group][split][Mpad][128] -> out[n][m] = (sum of the splits) / scale + bias[m]; optionally the per-channel sums of
// out and out^2 over all rows (BatchNorm batch statistics), accumulated in double precision.  Block = 128 rows x 64 channels.
__global__ void __launch_bounds__(256) pt_finish_kernel(const float* __restrict__ part, int splits, int Mpad, int M, int n_rows,
                                                        const float* __restrict__ bias, const float* __restrict__ scale2,
                                                        float* __restrict__ out, long long ldo, double* __restrict__ stats) {
  constexpr int kMC = 64;
  __shared__ float t_s[kMC][kNP + 1];
  const int g = blockIdx.x, m0 = blockIdx.y * kMC, tid = threadIdx.x;
  const float inv = scale2 ? scale2[1] : 1.0f;
  for (int i = tid; i < kMC * kNP; i += 256) {
    const int ml = i >> 7, nl = i & 127, m = m0 + ml;
    float acc = 0.f;
    for (int k = 0; k < splits; ++k) acc += part[(((size_t)g * splits + k) * Mpad + m) * kNP + nl];
    t_s[ml][nl] = acc * inv + ((bias && m < M) ? bias[m] : 0.f);
  }
  __syncthreads();
  for (int i = tid; i < kNP * kMC; i += 256) {
    const int nl = i >> 6, ml = i & 63, n = g * kNP + nl, m = m0 + ml;
    if (n < n_rows && m < M) out[(size_t)n * ldo + m] = t_s[ml][nl];
  }
  if (stats) {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ml = warp * 8 + q, m = m0 + ml;
      float s1 = 0.f, s2 = 0.f;
      for (int nl = lane; nl < kNP; nl += 32)
        if (g * kNP + nl < n_rows) { const float v = t_s[ml][nl]; s1 += v; s2 = fmaf(v, v, s2); }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0 && m < M) { atomicAdd(stats + m, (double)s1); atomicAdd(stats + M + m, (double)s2); }
    }
  }
}

// dropout keep decisions of 4 consecutive channels of one row from ONE Philox draw (replay array if given)
__device__ __forceinline__ void keep4(const uint8_t* replay, size_t elem0, unsigned long long seed, int mask_id, unsigned thresh, bool (&k)[4]) {
  if (replay) {
#pragma unroll
    for (int q = 0; q < 4; ++q) k[q] = replay[elem0 + q] != 0;
  } else {
    unsigned o[4];
    const unsigned long long i4 = elem0 >> 2;
    philox4x32_10((unsigned)i4, (unsigned)(i4 >> 32), (unsigned)mask_id, 0x504e5431u, (unsigned)seed, (unsigned)(seed >> 32), o);
#pragma unroll
    for (int q = 0; q < 4; ++q) k[q] = o[q] >= thresh;
  }
}

struct BnArgs {
  const float *mean, *rstd, *gamma, *beta;   // [C]
  int use_tanh;
  unsigned long long seed;
  int mask_id;
  unsigned thresh;                            // keep when the Philox word >= thresh
  float keep_scale;                           // 1 / (1 - p)
  const uint8_t* keep;                        // optional replayed mask [N][C]
};

// o[n][c] = dropout(act(gamma (y - mean) rstd + beta)) written through (b, t, c) strides: the interior of the next layer's
// halo-padded buffer, or the module's [B, C, T] output for the last layer
__global__ void __launch_bounds__(256) pt_bn_act_fwd_kernel(const float* __restrict__ y, int B, int T, int C, BnArgs a,
                                                            float* __restrict__ out, long long osb, long long ost, long long osc) {
  const int C4 = C >> 2;
  const size_t total = (size_t)B * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t n = i / C4;
    const int b = (int)(n / T), t = (int)(n - (size_t)b * T);
    const float4 yv = *reinterpret_cast<const float4*>(y + n * C + c);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    bool kp[4];
    keep4(a.keep, n * C + c, a.seed, a.mask_id, a.thresh, kp);
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float z = a.gamma[c + q] * ((yy[q] - a.mean[c + q]) * a.rstd[c + q]) + a.beta[c + q];
      const float act = a.use_tanh ? tanhf(z) : z;
      o[q] = kp[q] ? act * a.keep_scale : 0.f;
    }
    float* dst = out + b * osb + t * ost + c * osc;
    if (osc == 1) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    else {
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q * osc] = o[q];
    }
  }
}

// backward, first half: d[n][c] (gradient w.r.t. the layer output) -> dz = d . keep . scale . act'(z) in place, and the
// per-channel sums  sum_n dz  and  sum_n dz . zhat  (zhat = (y - mean) rstd) in double precision: sums[0..C), sums[C..2C).
// A block sweeps whole rows with every thread pinned to one channel quad, so the sums stay in registers until the end.
__global__ void __launch_bounds__(256) pt_bn_act_bwd1_kernel(float* __restrict__ d, const float* __restrict__ y, int N, int C, BnArgs a,
                                                             double* __restrict__ sums) {
  __shared__ float red_s[256][8];
  const int C4 = C >> 2, tid = threadIdx.x;
  const int lanes = 256 / C4 > 0 ? 256 / C4 : 1;          // rows in flight per block (C4 <= 256)
  const int c4 = tid % C4, rl = tid / C4, c = c4 * 4;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < lanes) {
    float mu[4], rs[4], ga[4], be[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { mu[q] = a.mean[c + q]; rs[q] = a.rstd[c + q]; ga[q] = a.gamma[c + q]; be[q] = a.beta[c + q]; }
    for (size_t n = (size_t)blockIdx.x * lanes + rl; n < (size_t)N; n += (size_t)gridDim.x * lanes) {
      const float4 yv = *reinterpret_cast<const float4*>(y + n * C + c);
      float4 dv = *reinterpret_cast<const float4*>(d + n * C + c);
      const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
      float dd[4] = {dv.x, dv.y, dv.z, dv.w};
      bool kp[4];
      keep4(a.keep, n * C + c, a.seed, a.mask_id, a.thresh, kp);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float zh = (yy[q] - mu[q]) * rs[q];
        float g = kp[q] ? dd[q] * a.keep_scale : 0.f;
        if (a.use_tanh) { const float u = tanhf(ga[q] * zh + be[q]); g *= 1.0f - u * u; }
        dd[q] = g;
        s[q] += g;
        s[4 + q] = fmaf(g, zh, s[4 + q]);
      }
      *reinterpret_cast<float4*>(d + n * C + c) = make_float4(dd[0], dd[1], dd[2], dd[3]);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) red_s[tid][q] = s[q];
  __syncthreads();
  if (tid < C4) {
    float t8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < lanes; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q) t8[q] += red_s[r * C4 + tid][q];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(sums + tid * 4 + q, (double)t8[q]);
      atomicAdd(sums + C + tid * 4 + q, (double)t8[4 + q]);
    }
  }
}

// backward, second half (BatchNorm in training mode): dy = gamma rstd (dz - mean(dz) - zhat mean(dz zhat)), written into the
// interior of the halo-padded gradient buffer [B][T + 4][C] that both the data-gradient and the weight-gradient products read
__global__ void __launch_bounds__(256) pt_bn_bwd2_kernel(const float* __restrict__ dz, const float* __restrict__ y, int B, int T, int C,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ m_dz,
                                                         const float* __restrict__ m_dzz, float* __restrict__ dy_pad) {
  const int C4 = C >> 2;
  const size_t total = (size_t)B * T * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t n = i / C4;
    const int b = (int)(n / T), t = (int)(n - (size_t)b * T);
    const float4 yv = *reinterpret_cast<const float4*>(y + n * C + c);
    const float4 dv = *reinterpret_cast<const float4*>(dz + n * C + c);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float zh = (yy[q] - mean[c + q]) * rstd[c + q];
      o[q] = gamma[c + q] * rstd[c + q] * (dd[q] - m_dz[c + q] - zh * m_dzz[c + q]);
    }
    *reinterpret_cast<float4*>(dy_pad + ((size_t)b * (T + 4) + t + 2) * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

struct RowsPlan { int Mpad, Kpad, groups, splits; size_t a_off, a_bytes, x_off, x_bytes, part_off, part_bytes, total; };
inline RowsPlan rows_plan(int M, int K, int n_rows, int num_sms) {
  RowsPlan p;
  p.Mpad = (M + 127) / 128 * 128;
  p.Kpad = (K + 63) / 64 * 64;
  p.groups = (n_rows + kNP - 1) / kNP;
  const int base = (p.Mpad / 128) * p.groups, kbs = p.Kpad / 64;
  int splits = 1;
  while (base * splits * 2 <= num_sms && kbs % (splits * 2) == 0 && kbs / (splits * 2) >= 4) splits *= 2;
  p.splits = splits;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  p.a_off = 256;
  p.a_bytes = up((size_t)p.Mpad * p.Kpad * 2);
  p.x_off = p.a_off + p.a_bytes;
  p.x_bytes = up((size_t)p.groups * kNP * p.Kpad * 2);
  p.part_off = p.x_off + p.x_bytes;
  p.part_bytes = up((size_t)p.groups * splits * p.Mpad * kNP * sizeof(float));
  p.total = p.part_off + p.part_bytes;
  return p;
}

}  // namespace pt
