// gemm_tc.cuh -- tcgen05 (5th-gen tensor core) GEMM for the batched decoder path on sm_100a.
//
//   D[split][M][NPAD] (fp32) = A[M][K] (fp16, weights)  x  X[NPAD][K]^T (fp16, activations)
//
// "Swap-AB" shape of an LSTM gate matmul: the 4096 gate rows are the MMA M dimension (128 per CTA), the
// batch is the MMA N dimension (16..128), K is split across CTAs so that ~128 CTAs pull the weights out of
// L2/HBM concurrently (the op is bandwidth-bound: every weight is used once per batch column).
//
// Both operands are PRE-TILED in global memory in exactly the shared-memory image tcgen05 wants for a
// K-major, non-swizzled ("interleaved") operand: 8-row x 16-byte core matrices, 128 contiguous bytes each,
// ordered [k-core][row-core] inside a (rows x 64) tile.  A tile is therefore one contiguous block and is
// staged with a single TMA bulk copy (cp.async.bulk, SASS UBLKCP) -- no tensor maps, no swizzle bookkeeping.
//   shared-memory descriptor: SBO (next 8-row group) = 128 B, LBO (next 8 K-elements) = rows/8 * 128 B.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lanes 32w..32w+31 -> registers -> global), warp 4 TMA
// producer, warp 5 TMEM allocation + single-thread MMA issue.  6/8-stage mbarrier pipeline; every barrier is
// waited on by exactly one thread that observes every phase in order.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // fp16 elements per k-block (8 core matrices of 8)
// pipeline depth: bytes in flight per CTA decide the achieved L2/HBM bandwidth of this bandwidth-bound GEMM
__host__ __device__ constexpr int stages_for(int npad) { return npad <= 64 ? 8 : 6; }
constexpr int kThreads = 192;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB

__host__ __device__ inline size_t tile_offset_bytes(int rows_in_tile, int r, int k) {
  // byte offset of element (r, k) inside a (rows_in_tile x 64) fp16 tile in core-matrix order
  return ((size_t)(k >> 3) * (rows_in_tile >> 3) + (r >> 3)) * 128 + (size_t)(r & 7) * 16 + (size_t)(k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug must fail the launch (trap), never hang the GPU
  for (unsigned spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // version
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// kind::f16 instruction descriptor: fp16 A/B (K-major), fp32 accumulate, M x N
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// operand format fields of the kind::f16 descriptor: a_format bits [7,10), b_format bits [10,13); 0 = f16, 1 = bf16
constexpr uint32_t kFmtF16 = 0u;
constexpr uint32_t kFmtBF16 = (1u << 7) | (1u << 10);
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

struct GemmParams {
  const unsigned char* a_tiles;   // [groups][M/128][K/64] tiles of 16 KB
  const unsigned char* x_tiles;   // per group: k-block tiles of NPAD*128 B
  float* out;                     // [groups][splits][M][NPAD]
  int M, K, splits, groups;       // K per group; each split covers K/splits columns (multiple of 64)
  long long x_group_stride;       // bytes between the X operands of consecutive groups (0 = shared buffer)
  int x_kb_base, x_kb_group_step; // first k-block of X inside its buffer = x_kb_base + group * x_kb_group_step
  const int* done;                // optional early-exit: skip when *done >= done_target (free-running decode)
  int done_target;
  uint32_t fmt;                   // operand formats OR-ed into the instruction descriptor (kFmtF16 / kFmtBF16)
  int stages;                     // 0 = stages_for(NPAD); experiments may ask for fewer
  int a_shared;                   // 1 = every group multiplies the same A tiles (e.g. a convolution's weights)
  unsigned long long* dbg;        // optional [CTAs][8] globaltimer stamps (ns): entry, setup done, first tile landed,
                                  // accumulator complete, epilogue stored, exit
};
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int NPAD>
__global__ void __launch_bounds__(kThreads, 1) gemm_f16_tn_kernel(const GemmParams p) {
  constexpr int kXTileBytes = NPAD * kBlockK * 2;
  constexpr int kMaxStages = stages_for(NPAD);
  const int kStages = (p.stages > 0 && p.stages < kMaxStages) ? p.stages : kMaxStages;
  constexpr int kTmemCols = NPAD < 32 ? 32 : NPAD;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_s = smem;                                   // [stages][16 KB]
  unsigned char* x_s = smem + kStages * kATileBytes;           // [stages][NPAD*128]
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  if (p.done != nullptr && __ldcg(p.done) >= p.done_target) return;   // uniform: nothing in this launch changes it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, split = blockIdx.y, group = blockIdx.z;
  unsigned long long* dbg = p.dbg ? p.dbg + ((size_t)(group * gridDim.y + split) * gridDim.x + mt) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = gtime_ns();
  const int m_tiles = p.M / kBlockM, kb_total = p.K / kBlockK, kb_per_split = kb_total / p.splits;
  const int kb0 = split * kb_per_split;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {   // TMEM allocation (one warp, same warp frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (dbg && threadIdx.x == 0) dbg[1] = gtime_ns();

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      const unsigned char* a_src = p.a_tiles + ((size_t)(p.a_shared ? 0 : group) * m_tiles + mt) * kb_total * kATileBytes;
      const unsigned char* x_src = p.x_tiles + (size_t)group * p.x_group_stride +
                                   (size_t)(p.x_kb_base + group * p.x_kb_group_step) * kXTileBytes;
      for (int i = 0; i < kb_per_split; ++i) {
        const int s = i % kStages;
        mbar_wait(&empty_bar[s], (uint32_t)(((i / kStages) & 1) ^ 1));
        mbar_expect_tx(&full_bar[s], kATileBytes + kXTileBytes);
        tma_load_1d(a_s + (size_t)s * kATileBytes, a_src + (size_t)(kb0 + i) * kATileBytes, kATileBytes, &full_bar[s]);
        tma_load_1d(x_s + (size_t)s * kXTileBytes, x_src + (size_t)(kb0 + i) * kXTileBytes, kXTileBytes, &full_bar[s]);
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBlockM, NPAD) | p.fmt;
      constexpr uint32_t lbo_a = (kBlockM / 8) * 128, lbo_x = (NPAD / 8) * 128, sbo = 128;
      for (int i = 0; i < kb_per_split; ++i) {
        const int s = i % kStages;
        mbar_wait(&full_bar[s], (uint32_t)((i / kStages) & 1));
        if (dbg && i == 0) dbg[2] = gtime_ns();
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_s + (size_t)s * kATileBytes);
        const uint32_t x_addr = smem_u32(x_s + (size_t)s * kXTileBytes);
#pragma unroll
        for (int j = 0; j < kBlockK / 16; ++j) {   // one UMMA consumes K = 16 = two 8-element core columns
          const uint64_t da = make_smem_desc(a_addr + j * 2 * lbo_a, lbo_a, sbo);
          const uint64_t dx = make_smem_desc(x_addr + j * 2 * lbo_x, lbo_x, sbo);
          umma_f16(tmem_base, da, dx, idesc, (i > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);            // frees the smem stage when these MMAs have read it
      }
      umma_commit(&tmem_full_bar);             // accumulator complete
    }
  } else {
    // ===== epilogue: warps 0-3 own TMEM lanes 32w .. 32w+31 (= output rows) =====
    mbar_wait(&tmem_full_bar, 0);
    if (dbg && threadIdx.x == 0) dbg[3] = gtime_ns();
    tc_fence_after();
    const int row = mt * kBlockM + warp * 32 + lane;
    float* dst = p.out + (((size_t)group * p.splits + split) * p.M + row) * NPAD;
#pragma unroll
    for (int c0 = 0; c0 < NPAD; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q)
        reinterpret_cast<float4*>(dst + c0)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                              __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
  }
  if (dbg && threadIdx.x == 0) dbg[4] = gtime_ns();
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[5] = gtime_ns();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// fp32 row-major [rows][K] -> fp16 tiles ([row_tiles][K/64] tiles of rows_per_tile x 64), rows >= n_rows are zero
__global__ void pack_tiles_kernel(const float* __restrict__ src, int n_rows, int K, int rows_per_tile, int n_row_tiles,
                                  unsigned char* __restrict__ dst, int as_bf16 = 0) {
  const size_t total = (size_t)n_row_tiles * rows_per_tile * K;
  const int kb_total = K / kBlockK;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int r = (int)(i / K);
    const int rt = r / rows_per_tile, rr = r - rt * rows_per_tile;
    const float v = r < n_rows ? src[(size_t)r * K + k] : 0.f;
    const size_t tile = ((size_t)rt * kb_total + (k / kBlockK)) * ((size_t)rows_per_tile * kBlockK * 2);
    unsigned char* at = dst + tile + tile_offset_bytes(rows_per_tile, rr, k % kBlockK);
    if (as_bf16) *reinterpret_cast<__nv_bfloat16*>(at) = __float2bfloat16(v);
    else *reinterpret_cast<__half*>(at) = __float2half(v);
  }
}

template <int NPAD>
inline cudaError_t prepare_gemm() {   // once per process / NPAD (not a stream operation)
  const size_t smem = (size_t)stages_for(NPAD) * (kATileBytes + NPAD * kBlockK * 2);
  return cudaFuncSetAttribute(gemm_f16_tn_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int NPAD>
inline cudaError_t launch_gemm(const GemmParams& p, cudaStream_t st) {
  const int stages = (p.stages > 0 && p.stages < stages_for(NPAD)) ? p.stages : stages_for(NPAD);
  const size_t smem = (size_t)stages * (kATileBytes + NPAD * kBlockK * 2);
  dim3 grid(p.M / kBlockM, p.splits, p.groups);
  gemm_f16_tn_kernel<NPAD><<<grid, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace tc
