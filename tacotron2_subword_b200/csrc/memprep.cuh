// memprep.cuh -- the step immediately upstream of the decoder, included by taco2dec.cu:
//
//   memory           = linear_converter(cat(encoder_outputs, cls_embeddings))    model.py:548-549 / 553-554  (1280 -> 512, bias)
//   processed_memory = attention_layer.memory_layer(memory)                      model.py:258-261            ( 512 -> 128, no bias)
//
// Both are dense contractions over all (utterance, position) rows and run on the tcgen05 GEMM of gemm_tc.cuh, chained
// without a pass in between: the pointwise kernel that finishes GEMM 1 (bias, transpose to [row][512]) also writes GEMM 2's
// operand tiles.  These tensors are INPUTS of the whole decoder recurrence (context vectors, attention energies), so a plain
// fp16-operand product (relative error ~3e-4) would move every downstream number; instead each fp32 operand is split into
// two fp16 terms, x = hi + lo, and the product is evaluated as  hi.hi + hi.lo + lo.hi  (the dropped lo.lo term is ~2^-22):
// three K segments of one GEMM, K' = 3K, fp32 accumulation in TMEM -- fp32-grade results (measured ~1e-6 relative) at
// tensor-core speed.  `lo` terms of magnitude < 6e-5 are fp16 subnormals, which tcgen05 handles exactly.
#pragma once

namespace mp {

constexpr int kNP = 128;   // rows ((utterance, position) pairs) per GEMM tile

__device__ __forceinline__ void split_hi_lo(float v, __half& hi, __half& lo) {
  hi = __float2half(v);
  lo = __float2half(v - __half2float(hi));
}

// weight [rows][K] fp32 -> A tiles [rows_pad/128][3K/64]: K segments [hi | hi | lo]   (pairs with X segments [hi | lo | hi])
__global__ void mp_pack_weights(const float* __restrict__ w, int rows, int K, int rows_pad, unsigned char* __restrict__ tiles) {
  const int kb_total = 3 * K / tc::kBlockK;
  const size_t total = (size_t)rows_pad * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), r = (int)(i / K);
    __half hi = __float2half(0.f), lo = hi;
    if (r < rows) split_hi_lo(w[(size_t)r * K + k], hi, lo);
#pragma unroll
    for (int seg = 0; seg < 3; ++seg) {
      const int kk = seg * K + k;
      const size_t tile = ((size_t)(r / 128) * kb_total + (kk / tc::kBlockK)) * tc::kATileBytes;
      *reinterpret_cast<__half*>(tiles + tile + tc::tile_offset_bytes(128, r % 128, kk % tc::kBlockK)) = seg == 2 ? lo : hi;
    }
  }
}

// 8 consecutive fp32 values of one row -> the three operand segments [hi | lo | hi] of that row's group tile
__device__ __forceinline__ void x_store_split8(unsigned char* X, int K, int n, int k0, const float* v) {
  __align__(16) __half hi[8], lo[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) split_hi_lo(v[q], hi[q], lo[q]);
  const int g = n / kNP, nl = n - g * kNP, kb_total = 3 * K / tc::kBlockK;
#pragma unroll
  for (int seg = 0; seg < 3; ++seg) {
    const int kk = seg * K + k0;
    unsigned char* tile = X + ((size_t)g * kb_total + (kk >> 6)) * ((size_t)kNP * 128);
    *reinterpret_cast<uint4*>(tile + tc::tile_offset_bytes(kNP, nl, kk & 63)) = *reinterpret_cast<const uint4*>(seg == 1 ? lo : hi);
  }
}

// GEMM-1 operand from the two inputs (the concatenation of model.py:548 is never materialised); rows >= n_rows are zero
__global__ void mp_input_kernel(const float* __restrict__ enc, int Ke, const float* __restrict__ cls, int Kc, int n_rows, int n_pad,
                                unsigned char* __restrict__ X) {
  const int K = Ke + Kc, chunks = K / 8;
  const size_t total = (size_t)n_pad * chunks;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % chunks), n = (int)(i / chunks), k0 = c8 * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < n_rows) {
      const float* src = k0 < Ke ? enc + (size_t)n * Ke + k0 : cls + (size_t)n * Kc + (k0 - Ke);
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    x_store_split8(X, K, n, k0, v);
  }
}

// split-K partials [group][split][M][128] (+ bias) -> out[row][M] (fp32, row-major) and, optionally, the next GEMM's operand.
// Block = one group x 32 output features: the [32 x 128] slab is transposed through shared memory.
__global__ void __launch_bounds__(256) mp_finish_kernel(const float* __restrict__ part, int splits, int M, const float* __restrict__ bias,
                                                        int n_rows, float* __restrict__ out, unsigned char* __restrict__ Xn) {
  __shared__ float t_s[32][kNP + 1];
  const int g = blockIdx.x, m0 = blockIdx.y * 32, tid = threadIdx.x;
  for (int i = tid; i < 32 * kNP; i += 256) {
    const int ml = i / kNP, nl = i - ml * kNP;
    float acc = bias ? bias[m0 + ml] : 0.f;
    for (int k = 0; k < splits; ++k) acc += part[(((size_t)g * splits + k) * M + m0 + ml) * kNP + nl];
    t_s[ml][nl] = acc;
  }
  __syncthreads();
  for (int i = tid; i < kNP * 4; i += 256) {          // (row, 8 features)
    const int nl = i >> 2, q8 = i & 3, n = g * kNP + nl;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = t_s[q8 * 8 + q][nl];
    if (n < n_rows) {
      float* dst = out + (size_t)n * M + m0 + q8 * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = 0.f;
    }
    if (Xn) x_store_split8(Xn, M, n, m0 + q8 * 8, v);
  }
}

// stand-alone operand builder for GEMM 2 when only processed_memory is wanted (the decoder's own call path)
__global__ void mp_memory_operand_kernel(const float* __restrict__ mem, int K, int n_rows, int n_pad, unsigned char* __restrict__ X) {
  const int chunks = K / 8;
  const size_t total = (size_t)n_pad * chunks;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % chunks), n = (int)(i / chunks);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < n_rows) {
      const float* src = mem + (size_t)n * K + c8 * 8;
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    x_store_split8(X, K, n, c8 * 8, v);
  }
}

inline int pick_splits(int ctas_per_split, int kb_total, int num_sms) {
  const int want = std::max(1, num_sms / std::max(1, ctas_per_split));
  int best = 1;
  for (int s = 2; s <= 12; ++s)
    if (s <= want && kb_total % s == 0) best = s;
  return best;
}

}  // namespace mp
