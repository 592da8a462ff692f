// wgrad.cuh -- weight-gradient contractions of the decoder's backward pass on tcgen05, included by taco2dec.cu.
//
// After the reverse-time loop every parameter gradient is a plain sum over (frame, utterance) rows:
//     C[m][n] = sum_k Y[k][m] . X[k][n],    k = (t, b),  K = T*B  (51,200 for BASELINE cfg 5)
// with Y = per-frame gradient rows (dG1, dG2, dq, d mel, ...) and X = saved activations (prenet, context, h1, h2, ...).  Both
// live row-major with k as the ROW index, the tcgen05 GEMM of gemm_tc.cuh wants both operands K-major: a pack kernel
// transposes 64-row slabs through shared memory into the core-matrix tile layout (fp16).  Gradient rows span many orders of
// magnitude, so Y is scaled by a power of two taken from its absolute maximum before the fp16 conversion (exact, undone in
// the finishing kernel): fp16 keeps 11 significant bits -- the TF32 grade these products had on the library path.
#pragma once

namespace wg {

constexpr int kNP = 128;

// one warp per (t, b) row, 16-byte loads when the row is aligned for them
__global__ void wg_absmax_kernel(const float* __restrict__ src, long long st, long long sb, int T, int B, int cols, unsigned* amax_bits) {
  float m = 0.f;
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5, K = T * B;
  const bool vec = (cols & 3) == 0 && (st & 3) == 0 && (sb & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < K; k += warps) {
    const int t = k / B, b = k - t * B;
    const float* row = src + t * st + b * sb;
    if (vec) {
      for (int c = lane * 4; c < cols; c += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + c));
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else {
      for (int c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(row[c]));
    }
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  if (lane == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));      // non-negative floats order like their bit patterns
}

// scale = 2^e with amax * scale in [2^13, 2^14): far from fp16 overflow (65504) even after rounding, and values down to
// amax * 2^-27 stay representable (subnormals); {scale, 1/scale} for the pack / finish kernels
__global__ void wg_scale_kernel(const unsigned* amax_bits, float* scale2) {
  const float amax = __uint_as_float(*amax_bits);
  float s = 1.0f;
  if (amax > 0.f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);                   // amax = f * 2^e, f in [0.5, 1)
    s = ldexpf(1.0f, 14 - e);
  }
  scale2[0] = s;
  scale2[1] = 1.0f / s;
}

// src rows k = (t, b) at src + t*st + b*sb, `cols` columns -> tiles [cols_pad/128][Kpad/64][128 x 64] (row = column of src,
// K = row of src), fp16, times *scale (null = 1).  Block = 64 k-rows x 128 columns, transposed through shared memory.
__global__ void __launch_bounds__(256) wg_pack_T_kernel(const float* __restrict__ src, long long st, long long sb, int T, int B, int cols,
                                                        int Kpad, const float* __restrict__ scale, unsigned char* __restrict__ dst) {
  __shared__ __align__(16) float t_s[64][kNP + 4];      // row stride 132 words: 16-byte stores and column reads both conflict-free
  __shared__ long long row_s[64];          // element offset of each of the slab's k-rows, -1 past the end
  const int kb = blockIdx.x, ct = blockIdx.y, K = T * B, tid = threadIdx.x;
  const float sc = scale ? scale[0] : 1.0f;
  if (tid < 64) {
    const int k = kb * 64 + tid, t = k / B, b = k - t * B;
    row_s[tid] = k < K ? t * st + b * sb : -1;
  }
  __syncthreads();
  const int c0 = ct * kNP;
  const bool vec = (st & 3) == 0 && (sb & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && c0 + kNP <= cols;
  if (vec) {
    const int cl = (tid & 31) * 4;
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {          // all eight loads in flight before the first use
      const long long off = row_s[(tid >> 5) + 8 * j];
      v[j] = off >= 0 ? __ldg(reinterpret_cast<const float4*>(src + off + c0 + cl)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float4*>(&t_s[(tid >> 5) + 8 * j][cl]) = make_float4(v[j].x * sc, v[j].y * sc, v[j].z * sc, v[j].w * sc);
    }
  } else {
    for (int i = tid; i < 64 * kNP; i += 256) {
      const int kl = i / kNP, cl = i - kl * kNP, c = c0 + cl;
      const long long off = row_s[kl];
      t_s[kl][cl] = (off >= 0 && c < cols) ? src[off + c] * sc : 0.f;
    }
  }
  __syncthreads();
  unsigned char* tile = dst + ((size_t)ct * (Kpad / 64) + kb) * tc::kATileBytes;
  for (int i = tid; i < kNP * 8; i += 256) {          // (column, group of 8 k) -> one 16-byte core-matrix row
    const int cl = i & 127, k8 = i >> 7;               // a warp writes 32 consecutive rows of one core-matrix column: 512 contiguous bytes
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = t_s[k8 * 8 + q][cl];
    *reinterpret_cast<uint4*>(tile + tc::tile_offset_bytes(kNP, cl, k8 * 8)) = pn::pack8(v);
  }
}

// partials [group][split][Mpad][128] -> C[m][group*128 + n] (* inv scale), optional accumulate
__global__ void __launch_bounds__(256) wg_finish_kernel(const float* __restrict__ part, int splits, int Mpad, int M, int N,
                                                        const float* __restrict__ scale2, float* __restrict__ C, long long ldc,
                                                        int accumulate) {
  const int g = blockIdx.y;
  const float inv = scale2 ? scale2[1] : 1.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * kNP; i += gridDim.x * blockDim.x) {
    const int m = i / kNP, nl = i - m * kNP, n = g * kNP + nl;
    if (n >= N) continue;
    float acc = 0.f;
    for (int k = 0; k < splits; ++k) acc += part[(((size_t)g * splits + k) * Mpad + m) * kNP + nl];
    float* dst = C + (size_t)m * ldc + n;
    *dst = accumulate ? *dst + acc * inv : acc * inv;
  }
}

// workspace: [scale block][A tiles][X tiles][partials].  K is padded to a multiple of 2048 so that the A image does not depend
// on N (any power-of-two split of its k-blocks is legal): consecutive products with the same Y reuse the packed A operand.
struct Plan { int Mpad, groups, splits, Kpad; size_t a_off, a_bytes, x_off, x_bytes, part_off, part_bytes, total; };
inline Plan plan(int M, int N, int K, int num_sms) {
  Plan p;
  p.Mpad = (M + 127) / 128 * 128;
  p.groups = (N + kNP - 1) / kNP;
  p.Kpad = (K + 2047) / 2048 * 2048;
  const int base_ctas = (p.Mpad / 128) * p.groups;
  int splits = 1;
  while (base_ctas * splits * 2 <= num_sms && splits < 32 && p.Kpad / 64 / (splits * 2) >= 8) splits *= 2;
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

}  // namespace wg

namespace wg {

// ---- the two small contractions that are NOT sums over (frame, utterance) rows: fp32 on the CUDA cores ----
// C[r][n] (+)= sum_k A[r][k] * Bm[k][n]  (* 2 where mask[r][n] > 0, else 0, when mask is given: the ReLU + dropout(0.5) of the
// prenet, model.py:23).  64 x 64 tile per block, 4 x 4 outputs per thread.
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bm, long long ldb,
                                                       float* __restrict__ C, long long ldc, int R, int N, int K, const float* __restrict__ mask,
                                                       long long ldm, int accumulate) {
  __shared__ float a_s[16][64 + 1], b_s[16][64 + 1];
  const int r0 = blockIdx.y * 64, n0 = blockIdx.x * 64, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rl = i >> 4, kl = i & 15;
      a_s[kl][rl] = (r0 + rl < R && k0 + kl < K) ? A[(size_t)(r0 + rl) * lda + k0 + kl] : 0.f;
      const int kl2 = i >> 6, nl = i & 63;
      b_s[kl2][nl] = (k0 + kl2 < K && n0 + nl < N) ? Bm[(size_t)(k0 + kl2) * ldb + n0 + nl] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kl = 0; kl < 16; ++kl) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = a_s[kl][ty * 4 + i]; bv[i] = b_s[kl][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (r >= R || n >= N) continue;
      float v = acc[i][j];
      if (mask) v = mask[(size_t)r * ldm + n] > 0.f ? 2.0f * v : 0.f;
      float* dst = C + (size_t)r * ldc + n;
      *dst = accumulate ? *dst + v : v;
    }
}

// per-utterance: C[b][m][n] = sum_t A[b][t][m] * Bm[t][b][n]   (d memory = alignments^T . d context, attention.py:395 backwards)
__global__ void __launch_bounds__(256) bmm_tn_kernel(const float* __restrict__ A, long long a_sb, long long a_st, const float* __restrict__ Bm,
                                                     long long b_st, long long b_sb, float* __restrict__ C, long long c_sb, long long ldc,
                                                     int M, int N, int T) {
  __shared__ float a_s[16][64 + 1], b_s[16][64 + 1];
  const int b = blockIdx.z, m0 = blockIdx.y * 64, n0 = blockIdx.x * 64, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* Ab = A + (size_t)b * a_sb;
  const float* Bb = Bm + (size_t)b * b_sb;
  float acc[4][4] = {};
  for (int t0 = 0; t0 < T; t0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int tl = i >> 6, l = i & 63;
      a_s[tl][l] = (t0 + tl < T && m0 + l < M) ? Ab[(size_t)(t0 + tl) * a_st + m0 + l] : 0.f;
      b_s[tl][l] = (t0 + tl < T && n0 + l < N) ? Bb[(size_t)(t0 + tl) * b_st + n0 + l] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int tl = 0; tl < 16; ++tl) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = a_s[tl][ty * 4 + i]; bv[i] = b_s[tl][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(size_t)b * c_sb + (size_t)m * ldc + n] = acc[i][j];
    }
}

}  // namespace wg
