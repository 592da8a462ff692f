// persist.cuh -- persistent batched decoder (2 <= B <= 128, SMA): ONE cooperative launch runs every frame.
//
// Same arithmetic as the per-frame graph of batched.cuh (fp16 operands on tcgen05, fp32 accumulation in TMEM, fp32
// pointwise / attention / projection) but without kernel boundaries: 128 CTAs stay resident, each owns one
// (128-row tile, K split) of BOTH LSTM gate products for the whole utterance and keeps its cell states in shared memory.
//
//   warp 16 (one thread)   TMA producer: walks the frame's tile program -- 14 k-blocks of the attention-LSTM product, 16 of
//                          the decoder-LSTM product -- ordered by when their inputs become available (h[t-1] first, then
//                          context, then the prenet / this frame's h1 and context), polls the producers' counters, and stages
//                          weight tile + activation tile per k-block with cp.async.bulk into an mbarrier ring
//   warp 17 (one thread)   issues tcgen05.mma (M = 128 gate rows, N = padded batch, K = 16) into two TMEM accumulators
//   warps 0-15             epilogue + everything pointwise: TMEM -> split-K partials -> (partner CTAs) -> LSTM cell update for
//                          a column share of the tile -> fp16 operand tiles of the consumers + query partials; then one
//                          attention task (utterance, stream) per CTA; free-running: projection partials, mel sum + stop
//                          test, and the two prenet layers, each spread over all 128 CTAs
//
// CTAs exchange through L2: data with plain stores, then ONE release-increment of a counter per CTA and phase; consumers
// acquire-poll the counter (no grid-wide barrier: a consumer only waits for the counters it depends on, and the producer
// warp prefetches the weight tiles of the next product while the compute warps are still in the previous phase).
//
// Reference arithmetic: model.py:322-390 (decode), attention.py:330-398 (SMA), model.py:13-24 (prenet), :480-485 (stop).
#pragma once

namespace pb {

using bt::A;
using bt::E;
using bt::H;
using bt::M;
using bt::P;

constexpr int kCtas = 128;                    // LSTM CTAs = 32 row tiles x 4 splits (decoder) = 2 x 32 x 2 (attention LSTMs)
constexpr int kCT = 512;                      // compute threads (warps 0-15)
constexpr int kThreads = kCT + 96;            // + producer warp + one MMA-issuing warp per gate product
constexpr int kTiles1 = 14;                   // k-blocks per CTA and frame, attention-LSTM product (K1 = 1792, split 2)
constexpr int kMaxTiles = 30;                 // + 16 of the decoder-LSTM product (K2 = 4096, split 4)
constexpr int kMelPad = 96;                   // 80 mel rows + gate, padded
constexpr int kSlices = 16;                   // context slices of the projection (64 features each)
constexpr long long kTimeoutClocks = 6000000000LL;

// counters (one 128-byte line each)
enum { F_P1 = 0, F_P2 = 64, F_H1 = 96, F_CTX = 98, F_H2 = 100, F_MEL = 101, F_L0 = 102, F_PRE = 104, F_DONE = 106, F_COUNT = 108 };
constexpr int kFlagStride = 32;               // words

struct PbParams {
  const unsigned char* wt;      // [kCtas][kMaxTiles] fp16 weight tiles (16 KB) in consumption order
  unsigned char* x1;            // [2 parities][S][28 kb][NPAD x 64] fp16   [prenet | ctx | h1]
  unsigned char* x2;            // [2 parities][K2/64 kb][NPAD x 64] fp16   [h1_0 | ctx_0 | (h1_1 | ctx_1) | h2]
  const unsigned char* xpre;    // teacher-forced: [T][S][4 kb][NPAD x 64] hoisted prenet tiles
  float* part1;                 // [S][32][2][128][NPAD]
  float* part2;                 // [32][4][128][NPAD]
  float* qpart;                 // [S][32][NPAD][A]
  float* melp;                  // [32][NPAD][kMelPad]   projection partials over the CTA's 32 h2 units
  float* ctxp;                  // [kSlices][NPAD][kMelPad] projection partials over a 64-feature context slice
  float* melx;                  // [NPAD][M]             mel frame fed back to the prenet
  float* l0x;                   // [S][NPAD][P]          prenet layer-0 activations
  unsigned* flags;              // [F_COUNT][kFlagStride]
  bt::Saved sv;
  int stages_a, stages_x;       // ring depths: streamed weight tiles / activation tiles
  int n_res;                    // weight tiles resident in shared memory
  int n_tm;                     // weight tiles resident in tensor memory (A operand read from TMEM)
  long long* dbg;               // optional [256] SM-clock stamps of CTA dbg_cta during frame dbg_frame (diagnostics):
  int dbg_frame, dbg_cta;       //   [0,32) activation tile requested, [32,64) operands landed (MMA thread), [64,96) MMAs issued,
};                              //   [96,128) weight tile requested, [128,160) compute-warp phase marks

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA bulk copy with an explicit L2 eviction policy: the 63 MB of weight tiles are re-read every frame and must stay in
// L2 (evict_last); per-frame teacher-forced prenet tiles are read exactly once (evict_first)
__device__ __forceinline__ void tma_load_1d_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                 unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(tc::smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(tc::smem_u32(bar)), "l"(policy)
               : "memory");
}
// tcgen05.mma with the A operand in tensor memory (128 lanes = rows, 8 columns = 16 fp16 K elements per instruction)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// non-blocking probe (try_wait may suspend the thread for a hardware time-out when the phase is not complete: fatal inside
// the producer's event loop, where a full ring on one cursor must not delay the other)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

struct Ctl {
  int* abort_flag;
  volatile int* s_exit;
};
// every wait in the kernel goes through one of these two: a protocol bug or a dead peer becomes an abort word all CTAs see
__device__ __noinline__ bool poll_ge(const unsigned* f, unsigned target, const Ctl& c) {
  if (target == 0u) return true;
  const long long t0 = clock64();
  unsigned spins = 0;
  while ((int)(ld_acquire_u32(f) - target) < 0) {
    if ((++spins & 63u) == 0u) {
      if (*c.s_exit) return false;
      if (*((volatile int*)c.abort_flag) != 0) { *c.s_exit = 1; return false; }
      if ((spins & 4095u) == 0u && clock64() - t0 > kTimeoutClocks) { atomicExch(c.abort_flag, 1); *c.s_exit = 1; return false; }
    }
  }
  return true;
}
__device__ __noinline__ bool mbar_wait_ab(uint64_t* bar, uint32_t parity, const Ctl& c) {
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 63u) == 0u) {
      if (*c.s_exit) return false;
      if (*((volatile int*)c.abort_flag) != 0) { *c.s_exit = 1; return false; }
      if ((spins & 1023u) == 0u && clock64() - t0 > kTimeoutClocks) { atomicExch(c.abort_flag, 1); *c.s_exit = 1; return false; }
    }
  }
  return true;
}
// LSTM gate activations with ex2-based forms (abs error ~1e-6, far below the fp16 operand rounding of this path)
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * x)); }
__device__ __forceinline__ float ftanh(float x) { return lat::fast_tanh(x); }
// one thread, after a compute-warp barrier: publish this CTA's stores of the phase
__device__ __forceinline__ void signal(unsigned* f) {
  red_release_add(f, 1u);      // release at gpu scope: cumulative over the stores the barrier before it ordered
}

// ------------------------------------------------------------------------------------------------------------------
// The frame's tile program.  Logical k-block numbering follows the operand buffers:
//   X1[s]: pre 0-3 | ctx 4-11 | h1 12-27            X2 (S=2): h1_0 0-15 | ctx_0 16-23 | h1_1 24-39 | ctx_1 40-47 | h2 48-63
//                                                   X2 (S=1): h1_0 0-15 | ctx_0 16-23 | h2 24-39
// ------------------------------------------------------------------------------------------------------------------
enum Dep { DEP_NONE = 0, DEP_H1_PREV, DEP_CTX_PREV, DEP_PRE, DEP_H2_PREV, DEP_H1_0, DEP_H1_1, DEP_CTX_0, DEP_CTX_1 };
struct TileInfo {
  int gemm;    // 0 = attention LSTM, 1 = decoder LSTM
  int kb;      // logical k-block inside the operand buffer
  int dep;
  int first, last;   // first / last tile of its product
};
__host__ __device__ inline int tiles2_of(int S) { return S == 2 ? 16 : 10; }
__host__ __device__ inline TileInfo tile_info(int i, bool has_g1, int S, int sig1, int sig2) {
  TileInfo ti;
  if (has_g1 && i < kTiles1) {
    ti.gemm = 0; ti.first = i == 0; ti.last = i == kTiles1 - 1;
    if (i < 8) { ti.kb = 12 + 8 * sig1 + i; ti.dep = DEP_H1_PREV; }
    else if (i < 12) { ti.kb = 4 + 4 * sig1 + (i - 8); ti.dep = DEP_CTX_PREV; }
    else { ti.kb = 2 * sig1 + (i - 12); ti.dep = DEP_PRE; }
    return ti;
  }
  const int j = has_g1 ? i - kTiles1 : i;
  const int n2 = tiles2_of(S);
  ti.gemm = 1; ti.first = j == 0; ti.last = j == n2 - 1;
  const int kb_h2 = S == 2 ? 48 : 24;
  if (j < 4) { ti.kb = kb_h2 + 4 * sig2 + j; ti.dep = DEP_H2_PREV; }
  else if (j < 8) { ti.kb = 4 * sig2 + (j - 4); ti.dep = DEP_H1_0; }
  else if (S == 2 && j < 12) { ti.kb = 24 + 4 * sig2 + (j - 8); ti.dep = DEP_H1_1; }
  else if (S == 2 && j < 14) { ti.kb = 16 + 2 * sig2 + (j - 12); ti.dep = DEP_CTX_0; }
  else if (S == 2) { ti.kb = 40 + 2 * sig2 + (j - 14); ti.dep = DEP_CTX_1; }
  else { ti.kb = 16 + 2 * sig2 + (j - 8); ti.dep = DEP_CTX_0; }
  return ti;
}

// Where a weight tile of the CTA's program lives: the later its activations arrive inside a frame, the more it pays to have
// it on-chip (its product sits on the critical path).  rank 0 = most critical.
__host__ __device__ inline int dep_rank(int dep) {
  switch (dep) {
    case DEP_PRE: return 0;
    case DEP_CTX_PREV: return 1;
    case DEP_CTX_0: case DEP_CTX_1: return 2;
    case DEP_H1_0: case DEP_H1_1: return 3;
    case DEP_H1_PREV: return 4;
    default: return 5;
  }
}
// per-tile record of the frame program, built once per CTA in shared memory (the producer / MMA threads are single
// threads on the critical path: their per-tile work has to be a table lookup)
struct TileRec {
  int kb, gemm, first, last, loc;
  int flag;        // counter the activation tile depends on (index into the flag block), -1 = none
  int mul, add;    // it is ready when counter >= mul * (frame + add)
  int xkind;       // 0: X1 of stream s1, 1: X2, 2: hoisted prenet tiles (teacher-forced)
};
constexpr int LOC_STREAM = -1, LOC_SMEM = 64;     // loc < 0: streamed; 0..63: tensor-memory slot; >= 64: shared-memory slot + 64

// weights -> per-CTA fp16 tiles in consumption order; tile rows = [gate g][unit u] of the CTA's 32 hidden units
struct PackSrc {
  const float* w_ih[2];
  const float* w_hh[2];
  const float* d_w_ih;
  const float* d_w_hh;
};
__global__ void pb_pack_weights(PackSrc src, int S, unsigned char* __restrict__ wt) {
  const int c = blockIdx.x;
  const bool has_g1 = c < S * 64;
  const int s1 = c / 64, mt1 = (c % 64) / 2, sig1 = c % 2, mt2 = c / 4, sig2 = c % 4;
  const int n_tiles = (has_g1 ? kTiles1 : 0) + tiles2_of(S);
  const int K2x = S * (H + E);
  for (int i = 0; i < n_tiles; ++i) {
    const TileInfo ti = tile_info(i, has_g1, S, sig1, sig2);
    unsigned char* tile = wt + ((size_t)c * kMaxTiles + i) * tc::kATileBytes;
    for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
      const int kk = e & 63, r = e >> 6, g = r >> 5, u = r & 31;
      const int k = ti.kb * 64 + kk;
      float v;
      if (ti.gemm == 0) {
        const size_t row = (size_t)g * H + mt1 * 32 + u;
        v = k < P + E ? src.w_ih[s1][row * (P + E) + k] : src.w_hh[s1][row * H + (k - (P + E))];
      } else {
        const size_t row = (size_t)g * H + mt2 * 32 + u;
        v = k < K2x ? src.d_w_ih[row * K2x + k] : src.d_w_hh[row * H + (k - K2x)];
      }
      *reinterpret_cast<__half*>(tile + tc::tile_offset_bytes(128, r, kk)) = __float2half(v);
    }
  }
}

// teacher-forced: hoisted prenet output [T+1][B][P] fp32 -> per-frame operand tiles [T][4 kb][NPAD x 64] fp16 of one stream
__global__ void pb_prenet_tiles(const float* __restrict__ pre, int T, int B, int NPAD, int S, int s, unsigned char* __restrict__ xpre) {
  const size_t total = (size_t)T * B * (P / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % (P / 8)), b = (int)((i / (P / 8)) % B), t = (int)(i / ((size_t)(P / 8) * B));
    const float* src = pre + ((size_t)t * B + b) * P + c8 * 8;
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const int k = c8 * 8;
    unsigned char* dst = xpre + (((size_t)t * S + s) * 4 + (k >> 6)) * ((size_t)NPAD * 128) + tc::tile_offset_bytes(NPAD, b, k & 63);
    *reinterpret_cast<uint4*>(dst) = pn::pack8(v);
  }
}

__device__ __forceinline__ unsigned char* x_chunk_ptr(unsigned char* xbase, int NPAD, int b, int k) {   // k % 8 == 0
  return xbase + (size_t)(k >> 6) * ((size_t)NPAD * 128) + tc::tile_offset_bytes(NPAD, b, k & 63);
}

// shared-memory plan (bytes); everything after the ring is fixed-size except the attention scratch
struct Smem {
  size_t aring, xring, res, c1, c2, hs, wq, wph, wpc, w0, w1, bias, att, red, total;
};
__host__ __device__ inline Smem smem_plan(int NPAD, int stages_a, int stages_x, int n_res, int max_ts, int fr) {
  Smem s;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 127) & ~(size_t)127; return o; };
  s.aring = take((size_t)2 * stages_a * tc::kATileBytes);
  s.xring = take((size_t)2 * stages_x * (size_t)NPAD * 128);
  s.res = take((size_t)n_res * tc::kATileBytes);
  s.c1 = take((size_t)32 * (NPAD / 2) * 4);
  s.c2 = take((size_t)32 * (NPAD / 4) * 4);
  s.hs = take((size_t)(NPAD / 2) * 36 * 4);
  s.wq = take((size_t)A * 33 * 4);
  s.wph = take(fr ? (size_t)(M + 1) * 33 * 4 : 0);
  s.wpc = take(fr ? (size_t)(M + 1) * 65 * 4 : 0);
  s.w0 = take(0);
  s.w1 = take(0);
  s.bias = take((size_t)2 * 128 * 4);
  s.att = take(bt::sma_smem_floats(max_ts) * 4);
  s.red = take((size_t)8 * kMelPad * 4 + 16 * 64 * 4);
  s.total = off;
  return s;
}

template <int NPAD>
__global__ void __launch_bounds__(kThreads, 1) decoder_batched_persistent(const __grid_constant__ Params p,
                                                                           const __grid_constant__ PbParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // one weight ring and one activation ring PER gate product: the early tiles of the next frame's attention-LSTM product
  // must not queue behind the decoder-LSTM tiles that are still waiting for this frame's context
  __shared__ __align__(8) uint64_t full_a[2][8], empty_a[2][8], full_x[2][8], empty_x[2][8], acc_full[2], acc_empty[2], res_bar;
  __shared__ int s_loc[32];
  __shared__ TileRec s_tile[32];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int s_exit;
  // rounds filled so far per ring slot: lanes / issuers may be several rounds away from a slot, and an mbarrier parity test
  // only distinguishes neighbouring phases -- nobody tests a slot's barrier before the round before its own has been filled
  __shared__ volatile unsigned s_xfill[2][8], s_afill[2][8];
  __shared__ volatile int s_go, s_stop_at;      // frames the compute warps have decided to run / first frame that does not run
  __shared__ volatile int s_ok[2];
  __shared__ long long s_ph[16];

  constexpr int kXTileBytes = NPAD * 128;
  constexpr int kTmemCols = 512;               // accumulators in columns [0, 2 NPAD), resident weight tiles above
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const int S = p.S, B = p.B, fr = p.free_running;
  const bool has_g1 = c < S * 64;
  const int s1 = c / 64, mt1 = (c % 64) / 2, sig1 = c % 2, mt2 = c / 4, sig2 = c % 4;
  const int n_tiles = (has_g1 ? kTiles1 : 0) + tiles2_of(S);
  const int NSA = q.stages_a, NSX = q.stages_x;
  const int n_steps = fr ? p.max_steps : p.T;
  constexpr int n_h1 = 64;                     // CTAs publishing h1 of one stream (32 row tiles x 2 splits)
  const size_t x1_stream = (size_t)28 * kXTileBytes, x1_par = (size_t)S * x1_stream;
  const size_t x2_par = (size_t)(S == 2 ? 64 : 40) * kXTileBytes;
  unsigned* const F = q.flags;
  auto flag = [&](int id) { return F + (size_t)id * kFlagStride; };

  int max_ts = 0;
  for (int s = 0; s < S; ++s) max_ts = max(max_ts, p.st[s].Ts);
  const Smem sp = smem_plan(NPAD, NSA, NSX, q.n_res, max_ts, fr);
  unsigned char* aring = smem + sp.aring;
  unsigned char* xring = smem + sp.xring;
  unsigned char* res_s = smem + sp.res;
  float* c1_s = (float*)(smem + sp.c1);
  float* c2_s = (float*)(smem + sp.c2);
  float* hs_s = (float*)(smem + sp.hs);
  float* wq_s = (float*)(smem + sp.wq);
  float* wph_s = (float*)(smem + sp.wph);
  float* wpc_s = (float*)(smem + sp.wpc);
  float* bias_s = (float*)(smem + sp.bias);     // [0,128): attention LSTM [g*32+u], [128,256): decoder LSTM
  float* att_s = (float*)(smem + sp.att);
  float* red_s = (float*)(smem + sp.red);

  // ---- one-off setup ---------------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int g = 0; g < 2; ++g)
      for (int i = 0; i < 8; ++i) {
        tc::mbar_init(&full_a[g][i], 1); tc::mbar_init(&empty_a[g][i], 1); tc::mbar_init(&full_x[g][i], 1); tc::mbar_init(&empty_x[g][i], 1);
      }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 1); }
    tc::mbar_init(&res_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_exit = 0; s_ok[0] = 1; s_ok[1] = 1;
    s_go = fr ? 2 : n_steps; s_stop_at = 0x7fffffff;
    for (int g = 0; g < 2; ++g) for (int i = 0; i < 8; ++i) { s_xfill[g][i] = 0; s_afill[g][i] = 0; }
    // placement of the program's weight tiles: tensor memory first, then shared memory, in order of criticality
    int n_tm = q.n_tm, n_sm = q.n_res, tm = 0, sm = 0;
    for (int i = 0; i < 32; ++i) s_loc[i] = LOC_STREAM;
    for (int rank = 0; rank < 6; ++rank)
      for (int i = 0; i < n_tiles; ++i) {
        if (dep_rank(tile_info(i, has_g1, S, sig1, sig2).dep) != rank) continue;
        if (tm < n_tm) s_loc[i] = tm++;
        else if (sm < n_sm) s_loc[i] = LOC_SMEM + sm++;
      }
    for (int i = 0; i < n_tiles; ++i) {
      const TileInfo ti = tile_info(i, has_g1, S, sig1, sig2);
      TileRec r;
      r.kb = ti.kb; r.gemm = ti.gemm; r.first = ti.first; r.last = ti.last; r.loc = s_loc[i];
      r.xkind = ti.gemm == 0 ? ((ti.dep == DEP_PRE && !fr) ? 2 : 0) : 1;
      switch (ti.dep) {
        case DEP_H1_PREV: r.flag = F_H1 + s1; r.mul = n_h1; r.add = 0; break;
        case DEP_CTX_PREV: r.flag = F_CTX + s1; r.mul = B; r.add = 0; break;
        case DEP_PRE: r.flag = fr ? F_PRE + s1 : -1; r.mul = 4 * ((B + 3) / 4); r.add = 0; break;       // prenet tasks per stream: 4 row blocks x groups of 4 utterances
        case DEP_H2_PREV: r.flag = F_H2; r.mul = kCtas; r.add = 0; break;
        case DEP_H1_0: r.flag = F_H1 + 0; r.mul = n_h1; r.add = 1; break;
        case DEP_H1_1: r.flag = F_H1 + 1; r.mul = n_h1; r.add = 1; break;
        case DEP_CTX_0: r.flag = F_CTX + 0; r.mul = B; r.add = 1; break;
        default: r.flag = F_CTX + 1; r.mul = B; r.add = 1; break;
      }
      s_tile[i] = r;
    }
  }
  if (tid < 16) s_ph[tid] = 0;
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < kCT) {
    for (int i = tid; i < 32 * (NPAD / 2); i += kCT) c1_s[i] = 0.f;
    for (int i = tid; i < 32 * (NPAD / 4); i += kCT) c2_s[i] = 0.f;
    for (int i = tid; i < 256; i += kCT) {
      const int which = i >> 7, r = i & 127, g = r >> 5, u = r & 31;
      float v = 0.f;
      if (which == 0) { if (has_g1) v = p.st[s1].b_ih[g * H + mt1 * 32 + u] + p.st[s1].b_hh[g * H + mt1 * 32 + u]; }
      else v = p.d_b_ih[g * H + mt2 * 32 + u] + p.d_b_hh[g * H + mt2 * 32 + u];
      bias_s[i] = v;
    }
    if (has_g1)
      for (int i = tid; i < A * 32; i += kCT) { const int a = i >> 5, u = i & 31; wq_s[a * 33 + u] = p.st[s1].wq[(size_t)a * H + mt1 * 32 + u]; }
    if (fr) {
      const int KD = H + S * E;
      for (int i = tid; i < (M + 1) * 32; i += kCT) {
        const int r = i >> 5, u = i & 31;
        wph_s[r * 33 + u] = r < M ? p.proj_w[(size_t)r * KD + mt2 * 32 + u] : p.gate_w[mt2 * 32 + u];
      }
      const int x = c & 15;
      if (x < S * 8)
        for (int i = tid; i < (M + 1) * 64; i += kCT) {
          const int r = i >> 6, k = i & 63;
          wpc_s[r * 65 + k] = r < M ? p.proj_w[(size_t)r * KD + H + x * 64 + k] : p.gate_w[H + x * 64 + k];
        }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t acc_addr[2] = {tmem_base, tmem_base + (uint32_t)NPAD};
  const uint32_t tm_w_col = (uint32_t)(2 * NPAD);      // first tensor-memory column of the resident weight tiles (32 per tile)
  Ctl ctl{p.abort_flag, &s_exit};
  const unsigned char* my_wt = q.wt + (size_t)c * kMaxTiles * tc::kATileBytes;

  // ---- weight tiles resident in tensor memory: copied in once (tcgen05.st), read by tcgen05.mma as its A operand for the
  //      whole utterance.  A [128 x 64] fp16 tile = 128 lanes x 32 columns (two K elements per 32-bit column); warp w fills
  //      lane quarter w & 3 of the slots w >> 2, w >> 2 + 4, ...; a thread owns one row: its eight 16-byte K groups. ----
  if (tid < kCT && q.n_tm > 0) {
    const int quarter = warp & 3, r = quarter * 32 + lane;
    for (int i = 0; i < n_tiles; ++i) {
      const int loc = s_loc[i];
      if (loc < 0 || loc >= LOC_SMEM || (loc & 3) != (warp >> 2)) continue;
      const unsigned char* src = my_wt + (size_t)i * tc::kATileBytes + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16;
      uint32_t wv[32];
#pragma unroll
      for (int kg = 0; kg < 8; ++kg) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + (size_t)kg * 16 * 128);
        wv[4 * kg] = u.x; wv[4 * kg + 1] = u.y; wv[4 * kg + 2] = u.z; wv[4 * kg + 3] = u.w;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + tm_w_col + (uint32_t)(loc * 32);
      lat::tmem_st16(taddr, wv);
      lat::tmem_st16(taddr + 16, wv + 16);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  // free-running: every role takes the same decision at the top of frame t: the utterances were all finished by the end
  // of frame t-2 (that fact is published before the prenet rows every role has -- transitively -- waited for)
  auto frame_runs = [&](int t) -> bool {
    if (!fr || t < 2) return true;
    const unsigned d = ld_relaxed_u32(flag(F_DONE));
    return !(d != 0u && (int)d <= t - 1);
  };

  // free-running: the compute warps decide at the top of every frame whether it runs (frame_runs) and publish the decision
  // in shared memory; the producer lanes and the MMA threads never work on a frame that has not been released
  auto wait_frame = [&](int t) -> bool {
    if (t < s_go) return true;
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
      if (t < s_go) return true;
      if (s_stop_at <= t || s_exit) return false;
      if ((++spins & 255u) == 0u) {
        if (*((volatile int*)p.abort_flag) != 0) { s_exit = 1; return false; }
        if ((spins & 4095u) == 0u && clock64() - t0 > kTimeoutClocks) { atomicExch(p.abort_flag, 1); s_exit = 1; return false; }
      }
    }
  };

  auto wait_fill = [&](volatile unsigned* f, unsigned want) -> bool {     // the ring slot has been handed to round want-1
    if ((int)(*f - want) >= 0) return true;
    const long long t0 = clock64();
    unsigned spins = 0;
    while ((int)(*f - want) < 0) {
      if ((++spins & 255u) == 0u) {
        if (s_exit) return false;
        if (*((volatile int*)p.abort_flag) != 0) { s_exit = 1; return false; }
        if ((spins & 4095u) == 0u && clock64() - t0 > kTimeoutClocks) { atomicExch(p.abort_flag, 1); s_exit = 1; return false; }
      }
    }
    return true;
  };

  if (warp == 16) {
    // =========================== TMA producer: one lane per tile of the frame program ===========================
    // Lane i owns tile i of every frame: it requests the weight tile as soon as its ring slot is free (weights depend on
    // nothing) and the activation tile as soon as the counter it depends on has reached the frame's target.  Lanes are
    // independent, so a dependency that is not ready yet never delays the tiles behind it.
    const unsigned long long pol_keep = lat::l2_policy_evict_last(), pol_once = lat::l2_policy_evict_first();
    if (lane == 0) {      // shared-memory resident tiles: one-off bulk copies
      unsigned bytes = 0;
      for (int i = 0; i < n_tiles; ++i) if (s_loc[i] >= LOC_SMEM) bytes += (unsigned)tc::kATileBytes;
      if (bytes) {
        tc::mbar_expect_tx(&res_bar, bytes);
        for (int i = 0; i < n_tiles; ++i)
          if (s_loc[i] >= LOC_SMEM)
            tma_load_1d_hint(res_s + (size_t)(s_loc[i] - LOC_SMEM) * tc::kATileBytes, my_wt + (size_t)i * tc::kATileBytes,
                             tc::kATileBytes, &res_bar, pol_once);
      } else {
        mbar_arrive(&res_bar);
      }
    }
    if (lane < n_tiles) {
      const int i = lane;
      const TileRec tr = s_tile[i];
      const bool streamed = tr.loc == LOC_STREAM;
      const int g = tr.gemm;
      const int g_lo = g == 0 ? 0 : (has_g1 ? kTiles1 : 0), g_hi = g == 0 ? kTiles1 : n_tiles, n_g = g_hi - g_lo;
      int arank = 0, n_streamed = 0;       // streamed weight tiles of this product: before this tile / per frame
      for (int j = g_lo; j < g_hi; ++j) { const int st_ = s_loc[j] == LOC_STREAM; n_streamed += st_; if (j < i) arank += st_; }
      unsigned char* const my_aring = aring + (size_t)g * NSA * tc::kATileBytes;
      unsigned char* const my_xring = xring + (size_t)g * NSX * kXTileBytes;
      const unsigned* fptr = tr.flag >= 0 ? flag(tr.flag) : nullptr;
      unsigned seen = 0;
      bool ok = true;
      for (int t = 0; t < n_steps && ok; ++t) {
        if (fr && !wait_frame(t)) break;
        bool need_a = streamed, need_x = true;
        const unsigned ga = (unsigned)t * (unsigned)n_streamed + (unsigned)arank, gx = (unsigned)t * (unsigned)n_g + (unsigned)(i - g_lo);
        const unsigned target = fptr ? (unsigned)tr.mul * (unsigned)(t + tr.add) : 0u;
        const unsigned char* xsrc = tr.xkind == 2 ? q.xpre + (((size_t)t * S + s1) * 4 + tr.kb) * kXTileBytes
                                    : tr.xkind == 0 ? q.x1 + (size_t)(t & 1) * x1_par + (size_t)s1 * x1_stream + (size_t)tr.kb * kXTileBytes
                                                    : q.x2 + (size_t)(t & 1) * x2_par + (size_t)tr.kb * kXTileBytes;
        unsigned spins = 0;
        long long t0 = 0;
        for (;;) {
          if (need_a) {
            const int slot = (int)(ga % (unsigned)NSA);
            const unsigned round = ga / (unsigned)NSA;
            if (s_afill[g][slot] == round && (round == 0u || mbar_test(&empty_a[g][slot], (round & 1u) ^ 1u))) {
              tc::mbar_expect_tx(&full_a[g][slot], (unsigned)tc::kATileBytes);
              tma_load_1d_hint(my_aring + (size_t)slot * tc::kATileBytes, my_wt + (size_t)i * tc::kATileBytes, tc::kATileBytes,
                               &full_a[g][slot], pol_keep);
              s_afill[g][slot] = round + 1u;
              if (q.dbg && c == q.dbg_cta && t == q.dbg_frame) q.dbg[96 + i] = clock64();
              need_a = false;
            }
          }
          if (need_x) {
            bool ready = fptr == nullptr || (int)(seen - target) >= 0;
            if (!ready) {
              seen = ld_acquire_u32(fptr);
              ready = (int)(seen - target) >= 0;
              if (ready) fence_proxy_async();      // the tiles were written through the generic proxy by other SMs
            }
            if (ready) {
              const int slot = (int)(gx % (unsigned)NSX);
              const unsigned round = gx / (unsigned)NSX;
              if (s_xfill[g][slot] == round && (round == 0u || mbar_test(&empty_x[g][slot], (round & 1u) ^ 1u))) {
                tc::mbar_expect_tx(&full_x[g][slot], (unsigned)kXTileBytes);
                if (tr.xkind == 2) tma_load_1d_hint(my_xring + (size_t)slot * kXTileBytes, xsrc, kXTileBytes, &full_x[g][slot], pol_once);
                else tc::tma_load_1d(my_xring + (size_t)slot * kXTileBytes, xsrc, kXTileBytes, &full_x[g][slot]);
                s_xfill[g][slot] = round + 1u;
                if (q.dbg && c == q.dbg_cta && t == q.dbg_frame) q.dbg[i] = clock64();
                need_x = false;
              }
            }
          }
          if (!need_a && !need_x) break;
          if ((++spins & 63u) == 0u) {
            if (s_exit) { ok = false; break; }
            if (*((volatile int*)p.abort_flag) != 0) { s_exit = 1; ok = false; break; }
            if (spins == 4096u) t0 = clock64();
            else if ((spins & 4095u) == 0u && clock64() - t0 > kTimeoutClocks) { atomicExch(p.abort_flag, 1); s_exit = 1; ok = false; break; }
          }
        }
      }
    }
  } else if (warp == 17 || warp == 18) {
    // =========================== MMA issuers: warp 17 = attention-LSTM product, warp 18 = decoder-LSTM product ==========
    const int my_gemm = warp - 17;
    if (lane == 0 && (my_gemm == 1 || has_g1)) {
      const uint32_t idesc = tc::make_idesc_f16(128, NPAD);
      constexpr uint32_t lbo_a = (128 / 8) * 128, lbo_x = (NPAD / 8) * 128, sbo = 128;
      const int i_lo = my_gemm == 0 ? 0 : (has_g1 ? kTiles1 : 0), i_hi = my_gemm == 0 ? kTiles1 : n_tiles;
      const int g = my_gemm;
      unsigned char* const my_aring = aring + (size_t)g * NSA * tc::kATileBytes;
      unsigned char* const my_xring = xring + (size_t)g * NSX * kXTileBytes;
      // descriptor templates (K-major, no swizzle): everything but the 14-bit start-address field is constant
      const uint64_t desc_hi_a = tc::make_smem_desc(0u, lbo_a, sbo), desc_hi_x = tc::make_smem_desc(0u, lbo_x, sbo);
      const uint32_t a_ring_addr = tc::smem_u32(my_aring), x_ring_addr = tc::smem_u32(my_xring), res_addr = tc::smem_u32(res_s);
      const uint32_t acc = acc_addr[my_gemm];
      bool ok = mbar_wait_ab(&res_bar, 0u, ctl);
      int sx = 0, sa = 0;
      uint32_t px = 0, pa = 0;
      for (int t = 0; t < n_steps && ok; ++t) {
        if (fr && !wait_frame(t)) break;
        if (t > 0) {     // the epilogue of the previous frame must have drained this accumulator
          ok = mbar_wait_ab(&acc_empty[my_gemm], (uint32_t)((t - 1) & 1), ctl);
          if (!ok) break;
        }
        const bool dbg_on = q.dbg && c == q.dbg_cta && t == q.dbg_frame;
        for (int i = i_lo; i < i_hi; ++i) {
          // this thread is the critical path of the product: per tile a couple of barrier probes, eight descriptor adds,
          // four MMAs and the commits -- nothing else (ring slots and parities are carried incrementally: no divisions)
          const int loc = s_loc[i];
          if (!mbar_try(&full_x[g][sx], px)) ok = mbar_wait_ab(&full_x[g][sx], px, ctl);
          if (ok && loc == LOC_STREAM && !mbar_try(&full_a[g][sa], pa)) ok = mbar_wait_ab(&full_a[g][sa], pa, ctl);
          if (!ok) break;
          if (dbg_on) q.dbg[32 + i] = clock64();
          tc::tc_fence_after();
          const uint64_t dx0 = desc_hi_x | (uint64_t)(((x_ring_addr + (uint32_t)sx * (uint32_t)kXTileBytes) >> 4) & 0x3fffu);
          const uint32_t first = (i == i_lo) ? 0u : 1u;
          if (loc >= 0 && loc < LOC_SMEM) {
            const uint32_t a_tm = tmem_base + tm_w_col + (uint32_t)(loc * 32);
            umma_f16_ts(acc, a_tm, dx0, idesc, first);
            umma_f16_ts(acc, a_tm + 8u, dx0 + (uint64_t)((2 * lbo_x) >> 4), idesc, 1u);
            umma_f16_ts(acc, a_tm + 16u, dx0 + (uint64_t)((4 * lbo_x) >> 4), idesc, 1u);
            umma_f16_ts(acc, a_tm + 24u, dx0 + (uint64_t)((6 * lbo_x) >> 4), idesc, 1u);
          } else {
            const uint32_t a_addr = loc == LOC_STREAM ? a_ring_addr + (uint32_t)sa * (uint32_t)tc::kATileBytes
                                                      : res_addr + (uint32_t)(loc - LOC_SMEM) * (uint32_t)tc::kATileBytes;
            const uint64_t da0 = desc_hi_a | (uint64_t)((a_addr >> 4) & 0x3fffu);
            tc::umma_f16(acc, da0, dx0, idesc, first);
            tc::umma_f16(acc, da0 + (uint64_t)((2 * lbo_a) >> 4), dx0 + (uint64_t)((2 * lbo_x) >> 4), idesc, 1u);
            tc::umma_f16(acc, da0 + (uint64_t)((4 * lbo_a) >> 4), dx0 + (uint64_t)((4 * lbo_x) >> 4), idesc, 1u);
            tc::umma_f16(acc, da0 + (uint64_t)((6 * lbo_a) >> 4), dx0 + (uint64_t)((6 * lbo_x) >> 4), idesc, 1u);
          }
          tc::umma_commit(&empty_x[g][sx]);
          if (++sx == NSX) { sx = 0; px ^= 1u; }
          if (loc == LOC_STREAM) {
            tc::umma_commit(&empty_a[g][sa]);
            if (++sa == NSA) { sa = 0; pa ^= 1u; }
          }
          if (i == i_hi - 1) tc::umma_commit(&acc_full[my_gemm]);
          if (dbg_on) q.dbg[64 + i] = clock64();
        }
      }
    }
  } else {
    // =========================== compute warps ===========================
    long long ph_t = clock64();
#define PB_PH(slot)                                       \
    if (c == 0 && tid == 0) {                             \
      const long long n_ = clock64();                     \
      s_ph[slot] += n_ - ph_t;                            \
      ph_t = n_;                                          \
    }                                                     \
    if (q.dbg && c == q.dbg_cta && tid == 0 && t == q.dbg_frame) q.dbg[128 + slot] = clock64();
    int wn = 0;      // wait counter: alternates the broadcast slot
    // thread 0 waits, everybody learns the outcome
#define PB_WAIT_FLAG(fptr, target)                                            \
    {                                                                         \
      if (tid == 0) s_ok[wn & 1] = poll_ge((fptr), (target), ctl) ? 1 : 0;    \
      bar_compute();                                                          \
      const int ok_ = s_ok[wn & 1];                                           \
      ++wn;                                                                   \
      if (!ok_) goto pb_done;                                                 \
    }
#define PB_WAIT_MBAR(bar, parity)                                             \
    {                                                                         \
      if (tid == 0) s_ok[wn & 1] = mbar_wait_ab((bar), (parity), ctl) ? 1 : 0;\
      bar_compute();                                                          \
      const int ok_ = s_ok[wn & 1];                                           \
      ++wn;                                                                   \
      if (!ok_) goto pb_done;                                                 \
    }
    const float sc_att = 1.0f / (1.0f - p.p_att), sc_dec = 1.0f / (1.0f - p.p_dec);
    const int row_ep = (warp & 3) * 32 + lane;
    const int n_grp = (B + 3) / 4;              // free-running prenet tasks: groups of 4 utterances

    for (int t = 0; t < n_steps; ++t) {
      if (!frame_runs(t)) break;
      // the fate of frame t+1 is already final here (every stop decision of frame t-1 was published before the prenet
      // rows this CTA waited for at the end of frame t-1): release it to the producer lanes / MMA threads one frame ahead,
      // so that their early tiles (h[t], context[t]) overlap the tail of this frame
      // (free-running: the stop decisions of frame t-1 are taken by the phoneme-stream prenet tasks just before they publish their
      // rows -- every CTA acquires that counter here, so that all of them read the same verdict about frame t+1)
      if (fr && t > 0) PB_WAIT_FLAG(flag(F_PRE + 0), (unsigned)(4 * n_grp) * (unsigned)t)
      if (fr && tid == 0) { if (frame_runs(t + 1)) s_go = t + 2; else s_stop_at = t + 1; }
      unsigned char* x1_next = q.x1 + (size_t)((t + 1) & 1) * x1_par;
      unsigned char* x2_cur = q.x2 + (size_t)(t & 1) * x2_par;
      unsigned char* x2_next = q.x2 + (size_t)((t + 1) & 1) * x2_par;

      // ---------------- attention LSTM: epilogue -> partials -> cell update for NPAD/2 columns ----------------
      if (has_g1) {
        PB_WAIT_MBAR(&acc_full[0], (uint32_t)(t & 1))
        PB_PH(0)
        tc::tc_fence_after();
        float* part_mine = q.part1 + (((size_t)(s1 * 32 + mt1) * 2 + sig1) * 128) * NPAD;
        for (int cg = warp >> 2; cg < NPAD / 16; cg += 4) {
          uint32_t v[16];
          const uint32_t taddr = acc_addr[0] + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cg * 16);
          lat::tmem_ld16(taddr, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float4* dst = reinterpret_cast<float4*>(part_mine + (size_t)row_ep * NPAD + cg * 16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            dst[k4] = make_float4(__uint_as_float(v[4 * k4]), __uint_as_float(v[4 * k4 + 1]), __uint_as_float(v[4 * k4 + 2]),
                                  __uint_as_float(v[4 * k4 + 3]));
        }
        tc::tc_fence_before();
        bar_compute();
        unsigned* fp1 = flag(F_P1 + s1 * 32 + mt1);
        if (tid == 0) { mbar_arrive(&acc_empty[0]); signal(fp1); }
        PB_PH(1)
        PB_WAIT_FLAG(fp1, 2u * (unsigned)(t + 1))
        PB_PH(2)
        constexpr int NC = NPAD / 2;
        const int col0 = sig1 * NC;
        const float* part_tile = q.part1 + ((size_t)(s1 * 32 + mt1) * 2 * 128) * NPAD;
#define PB_DBG_C(k) if (q.dbg && c == q.dbg_cta && tid == 0 && t == q.dbg_frame) q.dbg[176 + (k)] = clock64();
        PB_DBG_C(0)
        constexpr int CPT1 = (32 * NC + kCT - 1) / kCT;      // cells per thread: all partial loads are issued before any is used
        float pr1[CPT1][4];
#pragma unroll
        for (int ci = 0; ci < CPT1; ++ci) {
          const int e = tid + ci * kCT, bl = e % NC, u = (e / NC) & 31, b = col0 + bl;
          const bool live = e < 32 * NC && b < B;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float* pa = part_tile + (size_t)(g * 32 + u) * NPAD + (live ? b : 0);
            pr1[ci][g] = __ldcg(pa) + __ldcg(pa + (size_t)128 * NPAD);
          }
        }
        if (q.dbg && c == q.dbg_cta && tid == 0 && t == q.dbg_frame) { volatile float sink = pr1[0][0] + pr1[CPT1 - 1][3]; (void)sink; q.dbg[177] = clock64(); }
#pragma unroll
        for (int ci = 0; ci < CPT1; ++ci) {
          const int e = tid + ci * kCT;
          if (e >= 32 * NC) break;
          const int bl = e % NC, u = e / NC, b = col0 + bl;
          float hn = 0.f;
          if (b < B) {
            const int j = mt1 * 32 + u;
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) pre[g] = bias_s[g * 32 + u] + pr1[ci][g];
            const float gi = fsig(pre[0]), gf = fsig(pre[1]), gg = ftanh(pre[2]), go = fsig(pre[3]);
            float cn = gf * c1_s[u * NC + bl] + gi * gg;
            hn = go * ftanh(cn);
            if (q.sv.gates1) {
              float* sv = q.sv.gates1 + (((size_t)t * S + s1) * 5 * H + j) * B + b;
              const size_t gs = (size_t)H * B;
              __stcs(sv, gi); __stcs(sv + gs, gf); __stcs(sv + 2 * gs, gg); __stcs(sv + 3 * gs, go); __stcs(sv + 4 * gs, cn);   // written once, read by backward: keep them out of L2
            }
            if (p.training) {
              const size_t idx = (size_t)b * H + j;
              const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * s1) * B * H : nullptr;
              const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * s1 + 1) * B * H : nullptr;
              hn *= keep_mult(kh, idx, p.seed, 4 + 2 * s1, t, (int)idx, p.thresh_att, sc_att);
              cn *= keep_mult(kc, idx, p.seed, 5 + 2 * s1, t, (int)idx, p.thresh_att, sc_att);
            }
            c1_s[u * NC + bl] = cn;
          }
          hs_s[bl * 36 + u] = hn;
        }
        PB_DBG_C(2)
        bar_compute();
        PB_DBG_C(3)
        // fp16 operand chunks (8 units = 16 bytes): next frame's attention-LSTM input, this frame's decoder-LSTM input
        for (int e = tid; e < NC * 4; e += kCT) {
          const int bl = e >> 2, k8 = e & 3, b = col0 + bl;
          if (b >= B) continue;
          const float* hv = hs_s + bl * 36 + k8 * 8;
          const float v[8] = {hv[0], hv[1], hv[2], hv[3], hv[4], hv[5], hv[6], hv[7]};
          const uint4 pk = pn::pack8(v);
          const int j = mt1 * 32 + k8 * 8;
          *reinterpret_cast<uint4*>(x_chunk_ptr(x1_next + (size_t)s1 * x1_stream, NPAD, b, P + E + j)) = pk;
          *reinterpret_cast<uint4*>(x_chunk_ptr(x2_cur, NPAD, b, s1 * (H + E) + j)) = pk;
          if (q.sv.h1) {
            float* d = q.sv.h1 + (((size_t)(t + 1) * S + s1) * B + b) * H + j;
            __stcs(reinterpret_cast<float4*>(d), make_float4(v[0], v[1], v[2], v[3]));
            __stcs(reinterpret_cast<float4*>(d + 4), make_float4(v[4], v[5], v[6], v[7]));
          }
        }
        PB_DBG_C(4)
        // query partials over the CTA's 32 units: q_part[b][a] = sum_u Wq[a][u] h1[u][b]   (attention.py:56, 368)
        {
          const int a = tid & (A - 1), grp = tid >> 7;
          const float* wrow = wq_s + a * 33;
          float wr[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) wr[u] = wrow[u];
          float* qdst = q.qpart + ((size_t)(s1 * 32 + mt1) * NPAD) * A + a;
          for (int bl = grp; bl < NC; bl += 4) {
            const int b = col0 + bl;
            if (b >= B) continue;
            const float4* hv4 = reinterpret_cast<const float4*>(hs_s + bl * 36);
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int u4 = 0; u4 < 8; ++u4) {
              const float4 hq = hv4[u4];
              acc0 = fmaf(wr[4 * u4], hq.x, acc0); acc1 = fmaf(wr[4 * u4 + 1], hq.y, acc1);
              acc0 = fmaf(wr[4 * u4 + 2], hq.z, acc0); acc1 = fmaf(wr[4 * u4 + 3], hq.w, acc1);
            }
            qdst[(size_t)b * A] = acc0 + acc1;
          }
        }
        PB_DBG_C(5)
        bar_compute();
        PB_DBG_C(6)
        if (tid == 0) signal(flag(F_H1 + s1));
        PB_DBG_C(7)
        PB_PH(3)
      }

      // ---------------- attention: one (utterance, stream) task per CTA and round (attention.py:330-398) ----------------
      for (int tau = c; tau < S * B; tau += kCtas) {
        const int s = tau / B, b = tau - s * B;
        const StreamParams& sa = p.st[s];
        const int Ts = sa.Ts;
        const int len = sa.len ? (int)sa.len[b] : Ts;
        // independent utterances (free-running, GTA): positions >= len do not exist -- their alignment is exactly zero, so
        // neither their energies nor their memory rows are touched
        const int Tc = (fr || p.independent) ? min(len, Ts) : Ts;
        float* ctxr_s = att_s;                   // 4*E context partials
        float* q_s = ctxr_s + 4 * E;             // A
        float* v_s = q_s + A;                    // A
        float* e_s = v_s + A;                    // Ts+4
        float* ap_s = e_s + Ts + 4;              // Ts+4
        float* an_s = ap_s + Ts + 4;             // Ts+4
        // requests that do not depend on the query go out BEFORE the wait for h1: the first 8 memory rows of every context
        // thread and the processed-memory rows of the warp's first two energy rounds
        const int jg = tid >> 7, d4 = tid & 127;
        const float4* mem4 = reinterpret_cast<const float4*>(sa.mem + (size_t)b * Ts * E) + d4;
        float4 pf[bt::kCtxPF];
#pragma unroll
        for (int i = 0; i < bt::kCtxPF; ++i) {
          const int j = jg + 4 * i;
          pf[i] = j < Tc ? __ldg(mem4 + (size_t)j * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* pm_b = sa.pm + (size_t)b * Ts * A;
        float pmv[2][4][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const float* row = pm_b + (size_t)min(warp * 4 + r * 64 + pp, Tc - 1) * A;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) pmv[r][pp][cc] = __ldg(row + lane + 32 * cc);
          }
        for (int j = tid; j < Ts; j += kCT) ap_s[1 + j] = sa.a_prev[(size_t)b * Ts + j];
        if (tid == 0) ap_s[0] = 0.f;
        if (tid < A) v_s[tid] = sa.v[tid];
        PB_WAIT_FLAG(flag(F_H1 + s), (unsigned)n_h1 * (unsigned)(t + 1))
        PB_PH(4)
        if (tid < A) {
          const float* qs = q.qpart + ((size_t)(s * 32) * NPAD + b) * A + tid;
          float qa[32];
#pragma unroll
          for (int m = 0; m < 32; ++m) qa[m] = __ldcg(qs + (size_t)m * NPAD * A);
          float qv = 0.f;
#pragma unroll
          for (int m = 0; m < 32; ++m) qv += qa[m];
          if (q.sv.q) __stcs(q.sv.q + (((size_t)t * S + s) * B + b) * A + tid, qv);
          q_s[tid] = qv;
        }
        bar_compute();
        {
          const float q0 = q_s[lane], q1 = q_s[lane + 32], q2 = q_s[lane + 64], q3 = q_s[lane + 96];
          const float v0 = v_s[lane], v1 = v_s[lane + 32], v2 = v_s[lane + 64], v3 = v_s[lane + 96];
          auto round_of = [&](int j0, const float (&x)[4][4]) {
            float e[4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
              e[pp] = v0 * lat::fast_tanh(q0 + x[pp][0]) + v1 * lat::fast_tanh(q1 + x[pp][1]) +
                      v2 * lat::fast_tanh(q2 + x[pp][2]) + v3 * lat::fast_tanh(q3 + x[pp][3]);
            const float ev = lat::butterfly4(e[0], e[1], e[2], e[3], lane);
            const int j = j0 + (lane >> 3);
            if ((lane & 7) == 0 && j < Tc) e_s[j] = (j >= len) ? -INFINITY : ev;
          };
          for (int j = Tc + tid; j < Ts; j += kCT) e_s[j] = -INFINITY;
          if (warp * 4 < Tc) round_of(warp * 4, pmv[0]);
          if (warp * 4 + 64 < Tc) round_of(warp * 4 + 64, pmv[1]);
          for (int j0 = warp * 4 + 128; j0 < Tc; j0 += 64) {
            float x[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              const float* row = pm_b + (size_t)min(j0 + pp, Tc - 1) * A;
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) x[pp][cc] = __ldg(row + lane + 32 * cc);
            }
            round_of(j0, x);
          }
        }
        bar_compute();
        for (int j = tid; j < Ts; j += kCT) {
          float e = e_s[j];
          if (p.training) {
            const size_t ni = ((size_t)t * B + b) * Ts + j;
            const float nz = sa.noise ? sa.noise[ni] : philox_normal(p.seed, 10 + s, t, b * Ts + j);
            e = e + nz * 2.0f;
          }
          const float pj = sigmoidf_(e);
          e_s[j] = pj;
          if (sa.p_save) __stcs(sa.p_save + ((size_t)t * B + b) * Ts + j, pj);
        }
        bar_compute();
        {
          float* align_out = sa.align + ((size_t)b * p.Tcap + t) * Ts;
          for (int j = tid; j < Ts; j += kCT) {
            float a = ap_s[1 + j] * e_s[j];
            if (j > 0) a += ap_s[j] * (1.0f - e_s[j - 1]);
            if ((fr || p.independent) && j >= len) a = 0.0f;
            an_s[j] = a;
            sa.a_prev[(size_t)b * Ts + j] = a;
            align_out[j] = a;
          }
        }
        bar_compute();
        {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < bt::kCtxPF; ++i) {
            const int j = jg + 4 * i;
            const float a = j < Tc ? an_s[j] : 0.f;
            acc.x = fmaf(a, pf[i].x, acc.x); acc.y = fmaf(a, pf[i].y, acc.y); acc.z = fmaf(a, pf[i].z, acc.z); acc.w = fmaf(a, pf[i].w, acc.w);
          }
          for (int jb = jg + 4 * bt::kCtxPF; jb < Tc; jb += 4 * bt::kCtxPF) {
#pragma unroll
            for (int i = 0; i < bt::kCtxPF; ++i) {
              const int j = jb + 4 * i;
              pf[i] = j < Tc ? __ldg(mem4 + (size_t)j * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < bt::kCtxPF; ++i) {
              const int j = jb + 4 * i;
              const float a = j < Tc ? an_s[j] : 0.f;
              acc.x = fmaf(a, pf[i].x, acc.x); acc.y = fmaf(a, pf[i].y, acc.y); acc.z = fmaf(a, pf[i].z, acc.z); acc.w = fmaf(a, pf[i].w, acc.w);
            }
          }
          reinterpret_cast<float4*>(ctxr_s)[jg * (E / 4) + d4] = acc;
        }
        bar_compute();
        if (tid < E / 8) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int d = tid * 8 + k;
            v[k] = (ctxr_s[d] + ctxr_s[E + d]) + (ctxr_s[2 * E + d] + ctxr_s[3 * E + d]);
          }
          const int d0 = tid * 8;
          float* cdst = p.ctx + ((size_t)s * B + b) * E + d0;
          *reinterpret_cast<float4*>(cdst) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(cdst + 4) = make_float4(v[4], v[5], v[6], v[7]);
          if (q.sv.ctx) {
            float* d = q.sv.ctx + (((size_t)(t + 1) * S + s) * B + b) * E + d0;
            __stcs(reinterpret_cast<float4*>(d), make_float4(v[0], v[1], v[2], v[3]));
            __stcs(reinterpret_cast<float4*>(d + 4), make_float4(v[4], v[5], v[6], v[7]));
          }
          const uint4 pk = pn::pack8(v);
          *reinterpret_cast<uint4*>(x_chunk_ptr(x1_next + (size_t)s * x1_stream, NPAD, b, P + d0)) = pk;
          *reinterpret_cast<uint4*>(x_chunk_ptr(x2_cur, NPAD, b, s * (H + E) + H + d0)) = pk;
        }
        bar_compute();
        if (tid == 0) signal(flag(F_CTX + s));
        PB_PH(5)
      }

      // ---------------- free-running: projection partial over a 64-feature context slice (model.py:382-388) ----------------
      if (fr) {
        const int x = c & 15;
        if (x < S * 8) {
          const int sx = x >> 3, dx0 = (x & 7) * 64, grp = c >> 4;
          const int NB = (B + 7) / 8, b0 = grp * NB;
          PB_WAIT_FLAG(flag(F_CTX + sx), (unsigned)B * (unsigned)(t + 1))
          float* cs = red_s + 8 * kMelPad;        // [16][64]
          for (int e = tid; e < NB * 64; e += kCT) {
            const int bl = e >> 6, k = e & 63, b = b0 + bl;
            cs[e] = b < B ? __ldcg(p.ctx + ((size_t)sx * B + b) * E + dx0 + k) : 0.f;
          }
          bar_compute();
          if (tid < (M + 1) * 6) {
            const int r = tid % (M + 1), g6 = tid / (M + 1);
            float acc[3] = {0.f, 0.f, 0.f};             // NB <= 16 utterances -> at most 3 per thread
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float wr[32];
#pragma unroll
              for (int k = 0; k < 32; ++k) wr[k] = wpc_s[r * 65 + half * 32 + k];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const int bl = g6 + 6 * i;
                if (bl >= NB) break;
                const float4* xv4 = reinterpret_cast<const float4*>(cs + bl * 64 + half * 32);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                  const float4 xq = xv4[k4];
                  acc[i] = fmaf(wr[4 * k4], xq.x, acc[i]); acc[i] = fmaf(wr[4 * k4 + 1], xq.y, acc[i]);
                  acc[i] = fmaf(wr[4 * k4 + 2], xq.z, acc[i]); acc[i] = fmaf(wr[4 * k4 + 3], xq.w, acc[i]);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int bl = g6 + 6 * i, b = b0 + bl;
              if (bl < NB && b < B) q.ctxp[((size_t)x * NPAD + b) * kMelPad + r] = acc[i];
            }
          }
        }
        PB_PH(6)
      }

      // ---------------- decoder LSTM: epilogue -> partials -> cell update for NPAD/4 columns ----------------
      {
        PB_WAIT_MBAR(&acc_full[1], (uint32_t)(t & 1))
        PB_PH(7)
        tc::tc_fence_after();
        float* part_mine = q.part2 + (((size_t)mt2 * 4 + sig2) * 128) * NPAD;
        for (int cg = warp >> 2; cg < NPAD / 16; cg += 4) {
          uint32_t v[16];
          const uint32_t taddr = acc_addr[1] + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cg * 16);
          lat::tmem_ld16(taddr, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float4* dst = reinterpret_cast<float4*>(part_mine + (size_t)row_ep * NPAD + cg * 16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            dst[k4] = make_float4(__uint_as_float(v[4 * k4]), __uint_as_float(v[4 * k4 + 1]), __uint_as_float(v[4 * k4 + 2]),
                                  __uint_as_float(v[4 * k4 + 3]));
        }
        tc::tc_fence_before();
        bar_compute();
        unsigned* fp2 = flag(F_P2 + mt2);
        if (tid == 0) { mbar_arrive(&acc_empty[1]); signal(fp2); }
        PB_PH(8)
        PB_WAIT_FLAG(fp2, 4u * (unsigned)(t + 1))
        PB_PH(9)
        constexpr int NC = NPAD / 4;
        const int col0 = sig2 * NC;
        const float* part_tile = q.part2 + ((size_t)mt2 * 4 * 128) * NPAD;
        constexpr int CPT2 = (32 * NC + kCT - 1) / kCT;
        float pr2[CPT2][4];
#pragma unroll
        for (int ci = 0; ci < CPT2; ++ci) {
          const int e = tid + ci * kCT, bl = e % NC, u = (e / NC) & 31, b = col0 + bl;
          const bool live = e < 32 * NC && b < B;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float* pa = part_tile + (size_t)(g * 32 + u) * NPAD + (live ? b : 0);
            pr2[ci][g] = (__ldcg(pa) + __ldcg(pa + (size_t)128 * NPAD)) + (__ldcg(pa + (size_t)256 * NPAD) + __ldcg(pa + (size_t)384 * NPAD));
          }
        }
#pragma unroll
        for (int ci = 0; ci < CPT2; ++ci) {
          const int e = tid + ci * kCT;
          if (e >= 32 * NC) break;
          const int bl = e % NC, u = e / NC, b = col0 + bl;
          float hn = 0.f;
          if (b < B) {
            const int j = mt2 * 32 + u;
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) pre[g] = bias_s[128 + g * 32 + u] + pr2[ci][g];
            const float gi = fsig(pre[0]), gf = fsig(pre[1]), gg = ftanh(pre[2]), go = fsig(pre[3]);
            float cn = gf * c2_s[u * NC + bl] + gi * gg;
            hn = go * ftanh(cn);
            if (q.sv.gates2) {
              float* sv = q.sv.gates2 + ((size_t)t * 5 * H + j) * B + b;
              const size_t gs = (size_t)H * B;
              __stcs(sv, gi); __stcs(sv + gs, gf); __stcs(sv + 2 * gs, gg); __stcs(sv + 3 * gs, go); __stcs(sv + 4 * gs, cn);   // written once, read by backward: keep them out of L2
            }
            if (p.training) {
              const size_t idx = (size_t)b * H + j;
              const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 4) * B * H : nullptr;
              const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 5) * B * H : nullptr;
              hn *= keep_mult(kh, idx, p.seed, 8, t, (int)idx, p.thresh_dec, sc_dec);
              cn *= keep_mult(kc, idx, p.seed, 9, t, (int)idx, p.thresh_dec, sc_dec);
            }
            c2_s[u * NC + bl] = cn;
          }
          hs_s[bl * 36 + u] = hn;
        }
        bar_compute();
        for (int e = tid; e < NC * 4; e += kCT) {
          const int bl = e >> 2, k8 = e & 3, b = col0 + bl;
          if (b >= B) continue;
          const float* hv = hs_s + bl * 36 + k8 * 8;
          const float v[8] = {hv[0], hv[1], hv[2], hv[3], hv[4], hv[5], hv[6], hv[7]};
          const int j = mt2 * 32 + k8 * 8;
          *reinterpret_cast<uint4*>(x_chunk_ptr(x2_next, NPAD, b, S * (H + E) + j)) = pn::pack8(v);
          if (q.sv.h2) {
            float* d = q.sv.h2 + ((size_t)(t + 1) * B + b) * H + j;
            __stcs(reinterpret_cast<float4*>(d), make_float4(v[0], v[1], v[2], v[3]));
            __stcs(reinterpret_cast<float4*>(d + 4), make_float4(v[4], v[5], v[6], v[7]));
          }
        }
        if (fr) {   // projection partial over the CTA's 32 h2 units (model.py:382-388)
          if (tid < (M + 1) * 6) {
            const int r = tid % (M + 1), g6 = tid / (M + 1);
            float wr[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) wr[u] = wph_s[r * 33 + u];
            for (int bl = g6; bl < NC; bl += 6) {
              const int b = col0 + bl;
              if (b >= B) continue;
              const float4* hv4 = reinterpret_cast<const float4*>(hs_s + bl * 36);
              float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
              for (int u4 = 0; u4 < 8; ++u4) {
                const float4 hq = hv4[u4];
                acc0 = fmaf(wr[4 * u4], hq.x, acc0); acc1 = fmaf(wr[4 * u4 + 1], hq.y, acc1);
                acc0 = fmaf(wr[4 * u4 + 2], hq.z, acc0); acc1 = fmaf(wr[4 * u4 + 3], hq.w, acc1);
              }
              q.melp[((size_t)mt2 * NPAD + b) * kMelPad + r] = acc0 + acc1;
            }
          }
        }
        bar_compute();
        if (tid == 0) signal(flag(F_H2));
        PB_PH(10)
      }

      if (fr) {
        // ---------------- free-running feedback: mel / gate sum + stop test (model.py:382-388, 480-485) and both prenet layers of
        // frame t+1 (model.py:13-24, 470-471).  Round 2 first spread every step over all 128 CTAs (mel sum -> prenet L0 ->
        // prenet L1: three all-to-all exchanges through L2, 22 kcyc of a 88 kcyc frame, almost all of it exchange latency).  Now a
        // task = (stream, 64-row block of layer 1, group of 4 utterances) re-sums the projection partials of its utterances,
        // computes ALL of layer 0 for them (redundantly in the four row-block tasks: 80 KB of weights, read once for the four
        // utterances) and its 64 rows of layer 1 (64 KB): no exchange inside the prenet, 206 KB through the SM's L2 port per task,
        // and the only counter left is the one the attention-LSTM product waits for.  (One task per (utterance, stream) with all
        // 256 rows of layer 1 -- 336 KB per task -- measured 16 kcyc.) ----
        const int tt = t + 1;
        float* sum_s = red_s;                       // [3][4][kMelPad]  partial sums / later reused
        float* mel_s = att_s;                       // [4][kMelPad]     the attention scratch is free in this phase
        float* l0_s = mel_s + 4 * kMelPad;          // [4][P]
        float* pre_s = l0_s + 4 * P;                // [4][64]
        for (int id = c; id < S * 4 * n_grp; id += kCtas) {
          const int s = id / (4 * n_grp), rem = id - s * 4 * n_grp, rb = rem & 3, b0 = (rem >> 2) * 4;
          const StreamParams& sq = p.st[s];
          // layer-0 weight rows of the first pass (independent of the data): requested before the wait
          float4 w0v[8];
          {
            const int r0 = warp * 16;
#pragma unroll
            for (int r = 0; r < 8; ++r)
              w0v[r] = lane < M / 4 ? __ldg(reinterpret_cast<const float4*>(sq.pre_w0 + (size_t)(r0 + r) * M) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          // dropout multipliers: layer 0 -- (utterance u, row o) for o = tid & 255, u = 2 (tid >> 8) + {0, 1}; layer 1 -- (u, 64 rb + o)
          float pm0[2], pm1 = 0.f;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int u = 2 * (tid >> 8) + k, o = tid & (P - 1), b = b0 + u;
            pm0[k] = b < B ? keep_mult(sq.keep0, ((size_t)tt * B + b) * P + o, p.seed, s * 2 + 0, tt, b * P + o, p.thresh_pre, 2.0f) : 0.f;
          }
          if (tid < 4 * 64) {
            const int u = tid >> 6, o = rb * 64 + (tid & 63), b = b0 + u;
            pm1 = b < B ? keep_mult(sq.keep1, ((size_t)tt * B + b) * P + o, p.seed, s * 2 + 1, tt, b * P + o, p.thresh_pre, 2.0f) : 0.f;
          }
          PB_WAIT_FLAG(flag(F_H2), (unsigned)kCtas * (unsigned)(t + 1))
          PB_PH(11)
          // mel / gate of the four utterances: thread = (utterance u, row r < 81), 48 partials in two batches of 24 (every batch
          // is one round trip to L2 lines other SMs have just written: ~2 kcyc)
          {
            const int u = tid >> 7, r = tid & 127, b = b0 + u;
            const int n_part = 32 + S * 8;
            if (r <= M && b < B) {
              float acc = r < M ? p.proj_b[r] : p.gate_b[0];
#pragma unroll
              for (int k0 = 0; k0 < 48; k0 += 24) {
                float pv[24];
#pragma unroll
                for (int k = 0; k < 24; ++k) {
                  const int pi = k0 + k;
                  pv[k] = pi >= n_part ? 0.f
                          : pi < 32 ? __ldcg(q.melp + ((size_t)pi * NPAD + b) * kMelPad + r)
                                    : __ldcg(q.ctxp + ((size_t)(pi - 32) * NPAD + b) * kMelPad + r);
                }
                float a24 = 0.f;
#pragma unroll
                for (int k = 0; k < 24; ++k) a24 += pv[k];
                acc += a24;
              }
              mel_s[u * kMelPad + r] = acc;
              if (s == 0 && rb == 0) {
                if (r < M) {
                  p.mel[((size_t)b * p.Tcap + t) * M + r] = acc;
                } else {
                  p.gate[(size_t)b * p.Tcap + t] = acc;
                  if (p.n_frames[b] == 0) {
                    bool fin = false;
                    if (sigmoidf_(acc) > p.gate_thr) { p.n_frames[b] = t + 1; fin = true; }
                    else if (t + 1 == p.max_steps) { p.n_frames[b] = t + 1; p.reached_max[b] = 1; fin = true; }
                    if (fin && atomicAdd(p.done_count, 1) + 1 == B) atomicExch(flag(F_DONE), (unsigned)(t + 1));
                  }
                }
              }
            } else if (r <= M) {
              mel_s[u * kMelPad + r] = 0.f;
            }
          }
          bar_compute();
          PB_PH(12)
          // layer 0, all 256 rows for the four utterances: warp w owns rows [16w, 16w + 16), two passes of 8 rows; a weight row is
          // read once and used for the four utterances
          {
            float4 xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = lane < M / 4 ? reinterpret_cast<const float4*>(mel_s + u * kMelPad)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const int r0 = warp * 16 + pass * 8;
              if (pass == 1) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                  w0v[r] = lane < M / 4 ? __ldg(reinterpret_cast<const float4*>(sq.pre_w0 + (size_t)(r0 + r) * M) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int r4 = 0; r4 < 8; r4 += 4)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float acc[4];
#pragma unroll
                  for (int r = 0; r < 4; ++r)
                    acc[r] = w0v[r4 + r].x * xv[u].x + w0v[r4 + r].y * xv[u].y + w0v[r4 + r].z * xv[u].z + w0v[r4 + r].w * xv[u].w;
                  const float v = lat::butterfly4(acc[0], acc[1], acc[2], acc[3], lane);
                  if ((lane & 7) == 0) l0_s[u * P + r0 + r4 + (lane >> 3)] = v;
                }
            }
          }
          // layer-1 weight rows of this warp (4 of the task's 64): requested before the barrier
          float4 wa[4], wb[4];
          {
            const int r0 = rb * 64 + warp * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float4* wr = reinterpret_cast<const float4*>(sq.pre_w1 + (size_t)(r0 + r) * P) + lane * 2;
              wa[r] = __ldg(wr); wb[r] = __ldg(wr + 1);
            }
          }
          bar_compute();
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int u = 2 * (tid >> 8) + k, o = tid & (P - 1);
            l0_s[u * P + o] = fmaxf(l0_s[u * P + o], 0.f) * pm0[k];
          }
          bar_compute();
          PB_PH(13)
          {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 xa = reinterpret_cast<const float4*>(l0_s + u * P)[lane * 2], xb = reinterpret_cast<const float4*>(l0_s + u * P)[lane * 2 + 1];
              float acc[4];
#pragma unroll
              for (int r = 0; r < 4; ++r)
                acc[r] = (wa[r].x * xa.x + wa[r].y * xa.y + wa[r].z * xa.z + wa[r].w * xa.w) +
                         (wb[r].x * xb.x + wb[r].y * xb.y + wb[r].z * xb.z + wb[r].w * xb.w);
              const float v = lat::butterfly4(acc[0], acc[1], acc[2], acc[3], lane);
              if ((lane & 7) == 0) pre_s[u * 64 + warp * 4 + (lane >> 3)] = v;
            }
          }
          bar_compute();
          if (tid < 4 * 64) pre_s[tid] = fmaxf(pre_s[tid], 0.f) * pm1;
          bar_compute();
          if (tid < 4 * 8) {       // fp16 operand chunks of the attention-LSTM product of frame t+1: (utterance u, 8 rows)
            const int u = tid >> 3, k8 = tid & 7, b = b0 + u;
            if (b < B) {
              const float* pv = pre_s + u * 64 + k8 * 8;
              const float v[8] = {pv[0], pv[1], pv[2], pv[3], pv[4], pv[5], pv[6], pv[7]};
              *reinterpret_cast<uint4*>(x_chunk_ptr(x1_next + (size_t)s * x1_stream, NPAD, b, rb * 64 + k8 * 8)) = pn::pack8(v);
            }
          }
          bar_compute();
          if (tid == 0) signal(flag(F_PRE + s));
        }
        if (c >= S * 4 * n_grp) {      // CTAs without a task still follow the frame
          PB_WAIT_FLAG(flag(F_H2), (unsigned)kCtas * (unsigned)(t + 1))
        }
        PB_PH(14)
      }
    }
  pb_done:;
#undef PB_WAIT_FLAG
#undef PB_WAIT_MBAR
#undef PB_PH
    if (c == 0 && tid == 0 && p.phase_clocks)
      for (int i = 0; i < 16; ++i) p.phase_clocks[i] = s_ph[i];
  }

  // ---- teardown: every thread of every role ends up here (normally, or through the abort word) ----
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

}  // namespace pb
