// taco2dec.cu -- B200 (sm_100a) persistent Tacotron2 dual-stream mel decoder, fp32 exact path.
//
// One cooperative launch runs ALL frames of Decoder.forward / Decoder.inference
// (/root/reference/model.py:392-492).  148 CTAs stay resident; every frame is a fixed
// sequence of phases separated by a hand-written grid barrier:
//
//   [free-running only]  P0a prenet layer 0 -> P0b prenet layer 1          (model.py:13-24)
//   A  attention LSTM cells, both streams                                  (model.py:337-346)
//   Q  query projection  q = Wq h                                          (attention.py:56, 368)
//   B  energies -> (sigmoid + stepwise-monotonic | location conv + softmax) -> context
//                                                                          (attention.py:330-398, 7-85)
//   C  decoder LSTM cell                                                   (model.py:362-373)
//   D  mel / gate projection (+ stop test when free-running)               (model.py:382-388, 480-485)
//
// GEMV phases: one warp owns one hidden unit (its 4 gate rows), lanes stride over K with
// 16-byte streaming loads (ld.global.nc.L1::no_allocate), the activation vectors are staged
// once per CTA in shared memory, warp-shuffle reductions, LSTM pointwise fused in the warp.
// This is the bandwidth-bound small-batch path; weights are read in the reference's own
// PyTorch layouts (no repack) in fp32, so results agree with the fp32 oracle to ~1e-6.
//
// No CPU fallback, no multi-backend dispatch: the host API refuses non-sm_100 devices.

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "taco2dec.h"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxStreams = 2;
constexpr long long kBarrierTimeoutClocks = 6000000000LL;  // ~3 s at 1.9 GHz: watchdog, never hit in a healthy run

// ------------------------------------------------------------------------------------------
// Kernel parameter block
// ------------------------------------------------------------------------------------------
struct StreamParams {
  const float *pre_w0, *pre_w1, *w_ih, *w_hh, *b_ih, *b_hh, *wq, *wm, *v, *loc_conv, *loc_dense;
  const float* mem;         // [B, Ts, E]
  const long long* len;     // [B] or null
  const float* noise;       // [T, B, Ts] or null
  const uint8_t* keep0;     // prenet keep masks [rows, B, P] or null
  const uint8_t* keep1;
  float* pm;                // [B, Ts, A] processed memory
  float* pre;               // TF: [T+1, B, P] prenet output; FR: [B, P]
  float* pre0;              // FR: [B, P] prenet layer-0 output
  float* a_prev;            // [B, Ts]  SMA: alignment state; LSA: previous attention weights
  float* a_cum;             // [B, Ts]  LSA cumulative weights
  float* align;             // [B, Tcap, Ts] output
  float* p_save;            // [T, B, Ts] kept for the backward pass, or null: SMA selection probabilities /
                            // LSA cumulative weights entering the frame
  int Ts;
};

struct Params {
  int B, T, S, H, D, E, P, A, M, LF, LK;
  int attention, free_running, training, max_steps, Tcap;
  int independent;           // utterances are independent sequences: positions >= length do not exist (always so when free-running)
  float gate_thr, p_att, p_dec;
  unsigned thresh_pre, thresh_att, thresh_dec;  // Philox keep thresholds (u32 >= thresh -> keep)
  unsigned long long seed;
  StreamParams st[kMaxStreams];
  const float *d_w_ih, *d_w_hh, *d_b_ih, *d_b_hh, *proj_w, *proj_b, *gate_w, *gate_b;
  const float* dec_in;       // TF targets [B, M, T]
  const uint8_t* lstm_keep;  // [T, 6, B, H] or null
  float *h1, *c1;            // h1: [2 bufs][S][B][H], c1: [S][B][H]
  float *h2, *c2;            // h2: [2 bufs][B][D],    c2: [B][D]
  float* ctx;                // [S][B][E]
  float* q;                  // [S][B][A]
  float *mel, *gate;         // [B, Tcap, M], [B, Tcap]
  int *n_frames, *reached_max;  // [B] (free-running)
  unsigned* sync_ctr;        // grid barrier counter (zeroed by the host before launch)
  int* abort_flag;           // watchdog
  int* done_count;           // free-running: utterances that have stopped
  long long* phase_clocks;   // [16] SM-clock cycles CTA 0 spent per phase / per barrier (diagnostics)
};

// ------------------------------------------------------------------------------------------
// Small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Philox4x32-10, counter = (idx, row, mask_id, 0), key = seed.
__host__ __device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                                       unsigned k0, unsigned k1, unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned long long p0 = 0xD2511F53ull * c0;
    unsigned long long p1 = 0xCD9E8D57ull * c2;
    unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
    unsigned n1 = (unsigned)p1;
    unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    unsigned n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ bool philox_keep(unsigned long long seed, int mask_id, int row, int idx, unsigned thresh) {
  unsigned o[4];
  philox4x32_10((unsigned)idx, (unsigned)row, (unsigned)mask_id, 0u, (unsigned)seed, (unsigned)(seed >> 32), o);
  return o[0] >= thresh;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, int mask_id, int row, int idx) {
  unsigned o[4];
  philox4x32_10((unsigned)idx, (unsigned)row, (unsigned)mask_id, 0u, (unsigned)seed, (unsigned)(seed >> 32), o);
  float u1 = ((float)o[0] + 0.5f) * 2.3283064365386963e-10f;
  float u2 = ((float)o[1] + 0.5f) * 2.3283064365386963e-10f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// keep-mask lookup: replay array if given, else Philox.  Returns the multiplier (0 or scale).
__device__ __forceinline__ float keep_mult(const uint8_t* replay, size_t replay_index, unsigned long long seed,
                                           int mask_id, int row, int idx, unsigned thresh, float scale) {
  bool k = replay ? (replay[replay_index] != 0) : philox_keep(seed, mask_id, row, idx, thresh);
  return k ? scale : 0.0f;
}

}  // namespace
namespace {
#include "latency.cuh"
#include "gemm_tc.cuh"
}  // namespace
namespace {

// ------------------------------------------------------------------------------------------
// Grid barrier (all CTAs co-resident: cooperative launch).  Monotonic counter; thread 0 of
// each CTA arrives with a release-add and spins with acquire loads.  A clock watchdog turns
// a would-be hang into an abort flag every CTA observes.
// ------------------------------------------------------------------------------------------
struct GridBarrier {
  unsigned* ctr;
  int* abort_flag;
  unsigned target;
  unsigned nblocks;
};

__device__ __forceinline__ bool grid_sync(GridBarrier& gb, int* s_flag) {
  __syncthreads();
  gb.target += gb.nblocks;
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_add(gb.ctr, 1u);
    const long long t0 = clock64();
    int aborted = 0;
    unsigned spins = 0;
    while ((int)(ld_acquire_u32(gb.ctr) - gb.target) < 0) {
      if ((++spins & 1023u) == 0u) {
        if (*((volatile int*)gb.abort_flag) != 0) { aborted = 1; break; }
        if (clock64() - t0 > kBarrierTimeoutClocks) { atomicExch(gb.abort_flag, 1); aborted = 1; break; }
      }
    }
    __threadfence();
    *s_flag = aborted;
  }
  __syncthreads();
  return *s_flag == 0;
}

// ------------------------------------------------------------------------------------------
// Warp-level multi-row dot product: acc[r][b] += sum_k rows[r][k] * xs[b*ldx + k]
// rows: global fp32 (streamed, 16-byte loads), xs: shared (or global) fp32, K4 = K/4.
// ------------------------------------------------------------------------------------------
template <int R, int BT>
__device__ __forceinline__ void warp_dot(const float* const (&rows)[R], const float* xs, int ldx, int K4, int lane,
                                         float (&acc)[R][BT]) {
#pragma unroll 4
  for (int k4 = lane; k4 < K4; k4 += 32) {
    float4 w[R];
#pragma unroll
    for (int r = 0; r < R; ++r) w[r] = ld_stream4(rows[r] + 4 * k4);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(xs + (size_t)b * ldx + 4 * k4);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][b] = fmaf(w[r].x, x.x, acc[r][b]);
        acc[r][b] = fmaf(w[r].y, x.y, acc[r][b]);
        acc[r][b] = fmaf(w[r].z, x.z, acc[r][b]);
        acc[r][b] = fmaf(w[r].w, x.w, acc[r][b]);
      }
    }
  }
}

template <int R, int BT>
__device__ __forceinline__ void warp_reduce_all(float (&acc)[R][BT]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[r][b] = warp_sum(acc[r][b]);
}

// ------------------------------------------------------------------------------------------
// LSTM cell phase: tasks are hidden units; warp computes the 4 gate rows for a batch tile.
//   n_cells cells, each with its own weights / x layout in shared memory.
// ------------------------------------------------------------------------------------------
struct CellDesc {
  const float *w_ih, *w_hh, *b_ih, *b_hh;
  int Kx, Kh, Hn;            // input width, hidden width (recurrent), number of hidden units
  float* c;                  // [B][Hn]
  float* h_out;              // [B][Hn]
  const uint8_t* keep_h;     // replay [B][Hn] slices for this frame or null
  const uint8_t* keep_c;
  int mask_id_h, mask_id_c;  // Philox ids
  unsigned thresh;
  float scale;
  int xs_off;                // offset of this cell's x block in shared: [BT][Kx+Kh]
};

template <int BT>
__device__ __forceinline__ void lstm_phase(const CellDesc* cells, int n_cells, const float* xs, int b0, int nb,
                                           int training, unsigned long long seed, int t, int B) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int total = 0;
  for (int c = 0; c < n_cells; ++c) total += cells[c].Hn;
  // task i -> CTA (i % grid), warp slot (i / grid): spreads units evenly over SMs
  for (int slot = warp;; slot += kWarps) {
    const int task = slot * gridDim.x + blockIdx.x;
    if (task >= total) break;
    int ci = 0, j = task;
    while (j >= cells[ci].Hn) { j -= cells[ci].Hn; ++ci; }
    const CellDesc& cd = cells[ci];
    const int ldx = cd.Kx + cd.Kh;
    const float* x = xs + cd.xs_off;
    float acc[4][BT];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[g][b] = 0.0f;
    {
      const float* rows[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) rows[g] = cd.w_ih + (size_t)(g * cd.Hn + j) * cd.Kx;
      warp_dot<4, BT>(rows, x, ldx, cd.Kx >> 2, lane, acc);
#pragma unroll
      for (int g = 0; g < 4; ++g) rows[g] = cd.w_hh + (size_t)(g * cd.Hn + j) * cd.Kh;
      warp_dot<4, BT>(rows, x + cd.Kx, ldx, cd.Kh >> 2, lane, acc);
    }
    warp_reduce_all<4, BT>(acc);
    float bias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = cd.b_ih[g * cd.Hn + j] + cd.b_hh[g * cd.Hn + j];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      if (lane == b && b < nb) {
        const int bb = b0 + b;
        // nn.LSTMCell: gate order i, f, g, o
        const float ig = sigmoidf_(acc[0][b] + bias[0]);
        const float fg = sigmoidf_(acc[1][b] + bias[1]);
        const float gg = tanhf(acc[2][b] + bias[2]);
        const float og = sigmoidf_(acc[3][b] + bias[3]);
        const size_t idx = (size_t)bb * cd.Hn + j;
        float cn = fg * cd.c[idx] + ig * gg;
        float hn = og * tanhf(cn);
        if (training) {  // dropout on h AND c (model.py:341-346, 372-373); the dropped c feeds back
          hn *= keep_mult(cd.keep_h, idx, seed, cd.mask_id_h, t, (int)idx, cd.thresh, cd.scale);
          cn *= keep_mult(cd.keep_c, idx, seed, cd.mask_id_c, t, (int)idx, cd.thresh, cd.scale);
        }
        cd.c[idx] = cn;
        cd.h_out[idx] = hn;
      }
    }
  }
}

// generic "rows x batch-tile" linear phase: out(row, b) = dot(W[row], x_b); epilogue by functor
template <int BT, typename Epi>
__device__ __forceinline__ void linear_phase(const float* W, int n_rows, int K, const float* xs, int ldx, int task_base,
                                             int task_total, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int slot = warp;; slot += kWarps) {
    const int task = slot * gridDim.x + blockIdx.x;  // global task id across the whole phase
    if (task >= task_total) break;
    const int row = task - task_base;
    if (row < 0 || row >= n_rows) continue;
    float acc[1][BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[0][b] = 0.0f;
    const float* rows[1] = {W + (size_t)row * K};
    warp_dot<1, BT>(rows, xs, ldx, K >> 2, lane, acc);
    warp_reduce_all<1, BT>(acc);
#pragma unroll
    for (int b = 0; b < BT; ++b)
      if (lane == b) epi(row, b, acc[0][b]);
  }
}

// ------------------------------------------------------------------------------------------
// Attention task (one CTA per (batch, stream)): energies, probabilities, context.
// ------------------------------------------------------------------------------------------
template <bool kFast>   // kFast: ex2-based tanh + 4 positions per warp (batched tensor path; error ~1e-6)
__device__ void attention_task(const Params& p, int s, int b, int t, float* sm) {
  const StreamParams& sp = p.st[s];
  const int Ts = sp.Ts, A = p.A, E = p.E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  const int len = sp.len ? (int)sp.len[b] : Ts;
  const int pad = (p.attention == TACO2DEC_ATTN_LSA) ? (p.LK - 1) / 2 : 1;
  const int Tp = Ts + 2 * pad;
  // shared layout: 16-byte-aligned fixed-size blocks first, T-dependent scalar arrays last
  const int E4_ = E >> 2;
  const int NJ_ = kThreads / E4_ > 0 ? kThreads / E4_ : 1;
  float* red_s = sm;                                  // max(NJ*E, 64) partial contexts / block-reduce scratch
  float* q_s = red_s + (NJ_ * E > 64 ? NJ_ * E : 64); // A
  float* v_s = q_s + A;                               // A
  float* wd_s = v_s + A;                              // LF*A    : location_dense^T [f][a]   (LSA)
  float* wc_s = wd_s + p.LF * A;                      // 2*LK*LF : location_conv [c][k][f]   (LSA)
  float* e_s = wc_s + 2 * p.LK * p.LF;                // Ts : energies -> probabilities
  float* an_s = e_s + Ts;                             // Ts : new alignment
  float* ap_s = an_s + Ts;                            // Tp : padded previous alignment
  float* ac_s = ap_s + Tp;                            // Tp : padded cumulative (LSA)

  for (int i = tid; i < A; i += kThreads) {
    q_s[i] = __ldcg(p.q + ((size_t)s * p.B + b) * A + i);
    v_s[i] = sp.v[i];
  }
  for (int i = tid; i < Tp; i += kThreads) {
    const int j = i - pad;
    const bool in = (j >= 0 && j < Ts);
    ap_s[i] = in ? __ldcg(sp.a_prev + (size_t)b * Ts + j) : 0.0f;
    if (p.attention == TACO2DEC_ATTN_LSA) ac_s[i] = in ? __ldcg(sp.a_cum + (size_t)b * Ts + j) : 0.0f;
  }
  if (p.attention == TACO2DEC_ATTN_LSA) {
    for (int i = tid; i < p.LF * A; i += kThreads) {  // dense [A][LF] -> [LF][A]
      const int f = i / A, a = i - f * A;
      wd_s[i] = sp.loc_dense[(size_t)a * p.LF + f];
    }
    for (int i = tid; i < 2 * p.LK * p.LF; i += kThreads) {  // conv [LF][2][LK] -> [2][LK][LF]
      const int f = i % p.LF, ck = i / p.LF;
      wc_s[i] = sp.loc_conv[(size_t)f * 2 * p.LK + ck];
    }
  }
  __syncthreads();

  // ---- energies: one warp per position ----------------------------------------------
  const float* pm_b = sp.pm + (size_t)b * Ts * A;
  if (kFast && p.attention == TACO2DEC_ATTN_SMA && A == 128) {
    const float q0 = q_s[lane], q1 = q_s[lane + 32], q2 = q_s[lane + 64], q3 = q_s[lane + 96];
    const float v0 = v_s[lane], v1 = v_s[lane + 32], v2 = v_s[lane + 64], v3 = v_s[lane + 96];
    for (int j0 = warp * 4; j0 < Ts; j0 += kWarps * 4) {
      float e[4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const float* r = pm_b + (size_t)min(j0 + pp, Ts - 1) * A;
        e[pp] = v0 * lat::fast_tanh(q0 + __ldg(r + lane)) + v1 * lat::fast_tanh(q1 + __ldg(r + lane + 32)) +
                v2 * lat::fast_tanh(q2 + __ldg(r + lane + 64)) + v3 * lat::fast_tanh(q3 + __ldg(r + lane + 96));
      }
      const float ev = lat::butterfly4(e[0], e[1], e[2], e[3], lane);
      const int j = j0 + (lane >> 3);
      if ((lane & 7) == 0 && j < Ts) e_s[j] = (j >= len) ? -INFINITY : ev;
    }
  } else
  for (int j = warp; j < Ts; j += kWarps) {
    float feat = 0.0f;
    if (p.attention == TACO2DEC_ATTN_LSA) {
      // location conv (attention.py:12-15,20): lane f computes filter f at position j
      if (lane < p.LF) {
        for (int k = 0; k < p.LK; ++k) {
          feat = fmaf(wc_s[(0 * p.LK + k) * p.LF + lane], ap_s[j + k], feat);
          feat = fmaf(wc_s[(1 * p.LK + k) * p.LF + lane], ac_s[j + k], feat);
        }
      }
    }
    float part = 0.0f;
    for (int a = lane; a < A; a += 32) {
      float z = q_s[a] + __ldg(pm_b + (size_t)j * A + a);
      if (p.attention == TACO2DEC_ATTN_LSA) {
        float loc = 0.0f;  // location dense (attention.py:16-17,22)
        for (int f = 0; f < p.LF; ++f) loc = fmaf(wd_s[f * A + a], __shfl_sync(0xffffffffu, feat, f), loc);
        z += loc;
      }
      part = fmaf(v_s[a], kFast ? lat::fast_tanh(z) : tanhf(z), part);
    }
    part = warp_sum(part);
    if (lane == 0) e_s[j] = (j >= len) ? -INFINITY : part;  // masked_fill_(mask, -inf), attention.py:79,389
  }
  __syncthreads();

  float* align_out = sp.align + ((size_t)b * p.Tcap + t) * Ts;
  if (p.attention == TACO2DEC_ATTN_SMA) {
    // p_j = sigmoid(e_j [+ 2 N(0,1)])  (attention.py:340-352)
    for (int j = tid; j < Ts; j += kThreads) {
      float e = e_s[j];
      if (p.training) {
        const size_t ni = ((size_t)t * p.B + b) * Ts + j;
        const float nz = sp.noise ? sp.noise[ni] : philox_normal(p.seed, 10 + s, t, b * Ts + j);
        e = e + nz * 2.0f;
      }
      e_s[j] = sigmoidf_(e);
      if (sp.p_save) sp.p_save[((size_t)t * p.B + b) * Ts + j] = e_s[j];
    }
    __syncthreads();
    // alpha'_j = alpha_j p_j + alpha_{j-1} (1 - p_{j-1})  (attention.py:330-338)
    for (int j = tid; j < Ts; j += kThreads) {
      float a = ap_s[pad + j] * e_s[j];
      if (j > 0) a += ap_s[pad + j - 1] * (1.0f - e_s[j - 1]);
      if ((p.free_running || p.independent) && j >= len) a = 0.0f;  // independent utterances: padded positions do not exist
      an_s[j] = a;
      sp.a_prev[(size_t)b * Ts + j] = a;
      align_out[j] = a;
    }
  } else {
    // softmax over positions (attention.py:81)
    float m = -INFINITY;
    for (int j = tid; j < Ts; j += kThreads) m = fmaxf(m, e_s[j]);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if (lane == 0) red_s[warp] = m;
    __syncthreads();
    m = red_s[0];
    for (int w = 1; w < kWarps; ++w) m = fmaxf(m, red_s[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int j = tid; j < Ts; j += kThreads) {
      const float ex = expf(e_s[j] - m);
      e_s[j] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    if (lane == 0) red_s[warp] = sum;
    __syncthreads();
    sum = 0.0f;
    for (int w = 0; w < kWarps; ++w) sum += red_s[w];
    __syncthreads();
    for (int j = tid; j < Ts; j += kThreads) {
      const float a = e_s[j] / sum;
      an_s[j] = a;
      sp.a_prev[(size_t)b * Ts + j] = a;
      sp.a_cum[(size_t)b * Ts + j] = ac_s[pad + j] + a;  // model.py:358-359
      if (sp.p_save) sp.p_save[((size_t)t * p.B + b) * Ts + j] = ac_s[pad + j];   // cumulative weights entering the frame
      align_out[j] = a;
    }
  }
  __syncthreads();

  // ---- context = alpha' . memory  (attention.py:82, 395) ------------------------------
  // thread (jg, d4): float4 feature column d4, positions j = jg (mod NJ); partials reduced in smem
  const int E4 = E >> 2;
  const int NJ = kThreads / E4 > 0 ? kThreads / E4 : 1;  // 4 for E=512
  const float* mem_b = sp.mem + (size_t)b * Ts * E;
  if (tid < NJ * E4) {
    const int jg = tid / E4, d4 = tid - jg * E4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = jg; j < Ts; j += NJ) {
      const float a = an_s[j];
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(mem_b + (size_t)j * E) + d4);
      acc.x = fmaf(a, m4.x, acc.x);
      acc.y = fmaf(a, m4.y, acc.y);
      acc.z = fmaf(a, m4.z, acc.z);
      acc.w = fmaf(a, m4.w, acc.w);
    }
    reinterpret_cast<float4*>(red_s)[jg * E4 + d4] = acc;
  }
  __syncthreads();
  for (int d = tid; d < E; d += kThreads) {
    float c = 0.0f;
    for (int jg = 0; jg < NJ; ++jg) c += red_s[jg * E + d];
    p.ctx[((size_t)s * p.B + b) * E + d] = c;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// The persistent decoder kernel
// ------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(kThreads, 1) decoder_persistent(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_flag;
  __shared__ CellDesc s_cells[3];
  __shared__ long long s_ph[16];
  long long ph_t = clock64();
  if (threadIdx.x < 16) s_ph[threadIdx.x] = 0;
  // slot 2k = work of phase k, 2k+1 = the grid barrier that follows it (P0a,P0b,A,Q,B,C,D = k 0..6)
#define PH_MARK(slot)                                             \
  if (blockIdx.x == 0 && threadIdx.x == 0) {                      \
    const long long n_ = clock64();                               \
    s_ph[slot] += n_ - ph_t;                                      \
    ph_t = n_;                                                    \
  }
  GridBarrier gb{p.sync_ctr, p.abort_flag, 0u, gridDim.x};
  const int tid = threadIdx.x;
  const int gtid = blockIdx.x * kThreads + tid, gthreads = gridDim.x * kThreads;
  const int B = p.B, S = p.S, H = p.H, D = p.D, E = p.E, P = p.P, A = p.A, M = p.M;
  const int KA = P + E + H;            // attention-LSTM x block: [prenet | ctx | h]
  const int KC = S * (H + E) + D;      // decoder-LSTM x block:   [h, ctx, h_bert, ctx_bert | h2]
  const int KD = D + S * E;            // projection x block:     [h2, ctx, ctx_bert]
  const float sc_att = 1.0f / (1.0f - p.p_att), sc_dec = 1.0f / (1.0f - p.p_dec);

  // ---- init state (model.py:223-270): zeros; SMA alignment one-hot at 0 ------------------
  for (int i = gtid; i < 2 * S * B * H; i += gthreads) p.h1[i] = 0.0f;
  for (int i = gtid; i < S * B * H; i += gthreads) p.c1[i] = 0.0f;
  for (int i = gtid; i < 2 * B * D; i += gthreads) p.h2[i] = 0.0f;
  for (int i = gtid; i < B * D; i += gthreads) p.c2[i] = 0.0f;
  for (int i = gtid; i < S * B * E; i += gthreads) p.ctx[i] = 0.0f;
  for (int s = 0; s < S; ++s) {
    const int Ts = p.st[s].Ts;
    for (int i = gtid; i < B * Ts; i += gthreads) {
      p.st[s].a_prev[i] = (p.attention == TACO2DEC_ATTN_SMA && (i % Ts) == 0) ? 1.0f : 0.0f;
      p.st[s].a_cum[i] = 0.0f;
    }
  }
  if (p.free_running) {
    for (int i = gtid; i < B; i += gthreads) { p.n_frames[i] = 0; p.reached_max[i] = 0; }
    if (gtid == 0) *p.done_count = 0;
  }
  if (!grid_sync(gb, &s_flag)) return;
  PH_MARK(15)

  const int n_steps = p.free_running ? p.max_steps : p.T;
  int cur = 0;  // h1/h2 buffer holding the previous frame's hidden state
  for (int t = 0; t <= n_steps; ++t) {
    const bool last_tf_tail = (!p.free_running && t == n_steps);  // only the projection of frame T-1 is left
    if (p.free_running && t == n_steps) break;
    const int nxt = cur ^ 1;

    // ================= P0 (free-running): prenet on the previous mel (model.py:449-450, 470-471) =====
    if (p.free_running) {
      for (int layer = 0; layer < 2; ++layer) {
        const int K = layer == 0 ? M : P;
        for (int b0 = 0; b0 < B; b0 += BT) {
          const int nb = min(BT, B - b0);
          // x block: layer 0: [BT][M] shared by both streams; layer 1: [S][BT][P]
          if (layer == 0) {
            for (int i = tid; i < BT * M; i += kThreads) {
              const int b = i / M, m = i - b * M;
              smem[i] = (b < nb && t > 0) ? __ldcg(p.mel + ((size_t)(b0 + b) * p.Tcap + (t - 1)) * M + m) : 0.0f;
            }
          } else {
            for (int i = tid; i < S * BT * P; i += kThreads) {
              const int s = i / (BT * P), r = i - s * BT * P, b = r / P, k = r - b * P;
              smem[i] = (b < nb) ? __ldcg(p.st[s].pre0 + (size_t)(b0 + b) * P + k) : 0.0f;
            }
          }
          __syncthreads();
          for (int s = 0; s < S; ++s) {
            const StreamParams& sp = p.st[s];
            const float* W = layer == 0 ? sp.pre_w0 : sp.pre_w1;
            const uint8_t* keep = layer == 0 ? sp.keep0 : sp.keep1;
            float* out = layer == 0 ? sp.pre0 : sp.pre;
            const float* xs = layer == 0 ? smem : smem + (size_t)s * BT * P;
            auto epi = [&](int row, int b, float v) {
              if (b >= nb) return;
              const int bb = b0 + b;
              const size_t ki = ((size_t)t * B + bb) * P + row;
              const float mult = keep_mult(keep, ki, p.seed, s * 2 + layer, t, bb * P + row, p.thresh_pre, 2.0f);
              out[(size_t)bb * P + row] = fmaxf(v, 0.0f) * mult;  // dropout(relu(.), p=0.5, training=True)
            };
            linear_phase<BT>(W, P, K, xs, K, s * P, S * P, epi);
          }
          __syncthreads();
        }
        PH_MARK(2 * layer)
        if (!grid_sync(gb, &s_flag)) return;
        PH_MARK(2 * layer + 1)
      }
    }

    // ================= A: attention LSTM cells (+ D of the previous frame when teacher-forced) ========
    if (!last_tf_tail) {
      for (int b0 = 0; b0 < B; b0 += BT) {
        const int nb = min(BT, B - b0);
        for (int i = tid; i < S * BT * KA; i += kThreads) {
          const int s = i / (BT * KA), r = i - s * BT * KA, b = r / KA, k = r - b * KA;
          float v = 0.0f;
          if (b < nb) {
            const int bb = b0 + b;
            if (k < P) {
              const size_t row = p.free_running ? 0 : (size_t)t * B;
              v = __ldcg(p.st[s].pre + (row + bb) * P + k);
            } else if (k < P + E) {
              v = __ldcg(p.ctx + ((size_t)s * B + bb) * E + (k - P));
            } else {
              v = __ldcg(p.h1 + (((size_t)cur * S + s) * B + bb) * H + (k - P - E));
            }
          }
          smem[i] = v;
        }
        if (tid < S) {
          const StreamParams& sp = p.st[tid];
          CellDesc cd;
          cd.w_ih = sp.w_ih; cd.w_hh = sp.w_hh; cd.b_ih = sp.b_ih; cd.b_hh = sp.b_hh;
          cd.Kx = P + E; cd.Kh = H; cd.Hn = H;
          cd.c = p.c1 + (size_t)tid * B * H;
          cd.h_out = p.h1 + ((size_t)nxt * S + tid) * B * H;
          cd.keep_h = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * tid) * B * H : nullptr;
          cd.keep_c = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * tid + 1) * B * H : nullptr;
          cd.mask_id_h = 4 + 2 * tid; cd.mask_id_c = 5 + 2 * tid;
          cd.thresh = p.thresh_att; cd.scale = sc_att;
          cd.xs_off = tid * BT * KA;
          s_cells[tid] = cd;
        }
        __syncthreads();
        lstm_phase<BT>(s_cells, S, smem, b0, nb, p.training, p.seed, t, B);
        __syncthreads();
      }
    }
    if (!p.free_running && t > 0) {
      // D(t-1): mel / gate projection of the previous frame rides along (no dependency on phase A)
      for (int b0 = 0; b0 < B; b0 += BT) {
        const int nb = min(BT, B - b0);
        for (int i = tid; i < BT * KD; i += kThreads) {
          const int b = i / KD, k = i - b * KD;
          float v = 0.0f;
          if (b < nb) {
            const int bb = b0 + b;
            if (k < D) v = __ldcg(p.h2 + ((size_t)cur * B + bb) * D + k);
            else { const int s = (k - D) / E, d = (k - D) - s * E; v = __ldcg(p.ctx + ((size_t)s * B + bb) * E + d); }
          }
          smem[i] = v;
        }
        __syncthreads();
        auto epi_mel = [&](int row, int b, float v) {
          if (b < nb) p.mel[((size_t)(b0 + b) * p.Tcap + (t - 1)) * M + row] = v + p.proj_b[row];
        };
        linear_phase<BT>(p.proj_w, M, KD, smem, KD, 0, M + 1, epi_mel);
        auto epi_gate = [&](int row, int b, float v) {
          if (b < nb) p.gate[(size_t)(b0 + b) * p.Tcap + (t - 1)] = v + p.gate_b[0];
        };
        linear_phase<BT>(p.gate_w, 1, KD, smem, KD, M, M + 1, epi_gate);
        __syncthreads();
      }
    }
    if (last_tf_tail) break;
    PH_MARK(4)
    if (!grid_sync(gb, &s_flag)) return;
    PH_MARK(5)

    // ================= Q: query projections (attention.py:56, 368) =================================
    for (int b0 = 0; b0 < B; b0 += BT) {
      const int nb = min(BT, B - b0);
      for (int i = tid; i < S * BT * H; i += kThreads) {
        const int s = i / (BT * H), r = i - s * BT * H, b = r / H, k = r - b * H;
        smem[i] = (b < nb) ? __ldcg(p.h1 + (((size_t)nxt * S + s) * B + b0 + b) * H + k) : 0.0f;
      }
      __syncthreads();
      for (int s = 0; s < S; ++s) {
        auto epi = [&](int row, int b, float v) {
          if (b < nb) p.q[((size_t)s * B + b0 + b) * A + row] = v;
        };
        linear_phase<BT>(p.st[s].wq, A, H, smem + (size_t)s * BT * H, H, s * A, S * A, epi);
      }
      __syncthreads();
    }
    PH_MARK(6)
    if (!grid_sync(gb, &s_flag)) return;
    PH_MARK(7)

    // ================= B: attention (one CTA per (batch, stream)) ==================================
    for (int task = blockIdx.x; task < S * B; task += gridDim.x) attention_task<false>(p, task % S, task / S, t, smem);
    PH_MARK(8)
    if (!grid_sync(gb, &s_flag)) return;
    PH_MARK(9)

    // ================= C: decoder LSTM cell (model.py:362-373) =====================================
    for (int b0 = 0; b0 < B; b0 += BT) {
      const int nb = min(BT, B - b0);
      for (int i = tid; i < BT * KC; i += kThreads) {
        const int b = i / KC, k = i - b * KC;
        float v = 0.0f;
        if (b < nb) {
          const int bb = b0 + b;
          if (k < S * (H + E)) {
            const int s = k / (H + E), r = k - s * (H + E);
            v = (r < H) ? __ldcg(p.h1 + (((size_t)nxt * S + s) * B + bb) * H + r)
                        : __ldcg(p.ctx + ((size_t)s * B + bb) * E + (r - H));
          } else {
            v = __ldcg(p.h2 + ((size_t)cur * B + bb) * D + (k - S * (H + E)));
          }
        }
        smem[i] = v;
      }
      if (tid == 0) {
        CellDesc cd;
        cd.w_ih = p.d_w_ih; cd.w_hh = p.d_w_hh; cd.b_ih = p.d_b_ih; cd.b_hh = p.d_b_hh;
        cd.Kx = S * (H + E); cd.Kh = D; cd.Hn = D;
        cd.c = p.c2;
        cd.h_out = p.h2 + (size_t)nxt * B * D;
        cd.keep_h = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 4) * B * H : nullptr;
        cd.keep_c = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 5) * B * H : nullptr;
        cd.mask_id_h = 8; cd.mask_id_c = 9;
        cd.thresh = p.thresh_dec; cd.scale = sc_dec;
        cd.xs_off = 0;
        s_cells[2] = cd;
      }
      __syncthreads();
      lstm_phase<BT>(&s_cells[2], 1, smem, b0, nb, p.training, p.seed, t, B);
      __syncthreads();
    }
    PH_MARK(10)
    if (!grid_sync(gb, &s_flag)) return;
    PH_MARK(11)
    cur = nxt;

    // ================= D (free-running): projection + stop test (model.py:382-388, 480-485) ========
    if (p.free_running) {
      for (int b0 = 0; b0 < B; b0 += BT) {
        const int nb = min(BT, B - b0);
        for (int i = tid; i < BT * KD; i += kThreads) {
          const int b = i / KD, k = i - b * KD;
          float v = 0.0f;
          if (b < nb) {
            const int bb = b0 + b;
            if (k < D) v = __ldcg(p.h2 + ((size_t)cur * B + bb) * D + k);
            else { const int s = (k - D) / E, d = (k - D) - s * E; v = __ldcg(p.ctx + ((size_t)s * B + bb) * E + d); }
          }
          smem[i] = v;
        }
        __syncthreads();
        auto epi_mel = [&](int row, int b, float v) {
          if (b < nb) p.mel[((size_t)(b0 + b) * p.Tcap + t) * M + row] = v + p.proj_b[row];
        };
        linear_phase<BT>(p.proj_w, M, KD, smem, KD, 0, M + 1, epi_mel);
        auto epi_gate = [&](int row, int b, float v) {
          if (b >= nb) return;
          const int bb = b0 + b;
          const float g = v + p.gate_b[0];
          p.gate[(size_t)bb * p.Tcap + t] = g;
          if (p.n_frames[bb] == 0) {
            if (sigmoidf_(g) > p.gate_thr) {            // strict >, stop frame kept (model.py:480-481)
              p.n_frames[bb] = t + 1;
              atomicAdd(p.done_count, 1);
            } else if (t + 1 == p.max_steps) {          // model.py:482-485 -> INFER_FLAG = False
              p.n_frames[bb] = t + 1;
              p.reached_max[bb] = 1;
              atomicAdd(p.done_count, 1);
            }
          }
        };
        linear_phase<BT>(p.gate_w, 1, KD, smem, KD, M, M + 1, epi_gate);
        __syncthreads();
      }
      PH_MARK(12)
      if (!grid_sync(gb, &s_flag)) return;
      PH_MARK(13)
      if (__ldcg(p.done_count) >= B) break;  // uniform: every CTA reads it after the same barrier
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int i = 0; i < 16; ++i) p.phase_clocks[i] = s_ph[i];
#undef PH_MARK
}

#include "batched.cuh"
#include "backward.cuh"
#include "postnet.cuh"
#include "persist.cuh"
#include "persist_bwd.cuh"
#include "memprep.cuh"
#include "loss.cuh"
#include "wgrad.cuh"
#include "postnet_train.cuh"

// ------------------------------------------------------------------------------------------
// One-off kernels: processed memory (model.py:258-261) and the hoisted teacher-forced prenet
// (model.py:412-413).  Warp per output row, same warp_dot as the persistent kernel.
// ------------------------------------------------------------------------------------------
// A work item is (memory row n, column chunk c of n_chunk): small calls (one utterance = 150 rows) are cut by output column so
// that they fill the machine instead of running 32 dependent dot products per warp; large calls use n_chunk = 1 (a warp
// walks all A columns of its row, the warps of a block share the weight rows through L1).  The summation order of an
// output element does not depend on n_chunk.
__global__ void __launch_bounds__(256) processed_memory_kernel(const float* __restrict__ mem, const float* __restrict__ wm,
                                                               float* __restrict__ pm, int n_rows, int E, int A, int n_chunk) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int chunk_cols = (((A + 3) / 4 + n_chunk - 1) / n_chunk) * 4;
  for (int item = gw; item < n_rows * n_chunk; item += nw) {
    const int n = item / n_chunk, c = item - n * n_chunk;
    const float* x = mem + (size_t)n * E;
    const int a_end = min(A, (c + 1) * chunk_cols);
    for (int a = c * chunk_cols; a < a_end; a += 4) {
      float acc[4][1] = {{0.f}, {0.f}, {0.f}, {0.f}};
      const float* rows[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) rows[r] = wm + (size_t)min(a + r, A - 1) * E;
      warp_dot<4, 1>(rows, x, E, E >> 2, lane, acc);
      warp_reduce_all<4, 1>(acc);
      if (lane < 4 && a + lane < A) {
        float v = lane == 0 ? acc[0][0] : lane == 1 ? acc[1][0] : lane == 2 ? acc[2][0] : acc[3][0];
        pm[(size_t)n * A + a + lane] = v;
      }
    }
  }
}

// rows = (T+1)*B frames; frame 0 is the all-zero go-frame (model.py:407-411); frame r>0 reads
// decoder_inputs[b, :, r-1] (the [B, M, T] layout the reference transposes at model.py:283-287).
__global__ void __launch_bounds__(256) prenet_tf_kernel(const float* __restrict__ dec_in, const float* __restrict__ w0,
                                                        const float* __restrict__ w1, const uint8_t* keep0,
                                                        const uint8_t* keep1, float* __restrict__ out,
                                                        float* __restrict__ out0, int B, int T,
                                                        int M, int P, unsigned long long seed, int stream_id,
                                                        unsigned thresh) {
  extern __shared__ __align__(16) float sm[];  // per warp: x[M4] + h0[P]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int Mp = (M + 3) & ~3;
  float* x = sm + (size_t)warp * (Mp + P);
  float* h0 = x + Mp;
  const int n_rows = (T + 1) * B;
  for (int n = blockIdx.x * wpb + warp; n < n_rows; n += gridDim.x * wpb) {
    const int r = n / B, b = n - r * B;
    for (int m = lane; m < Mp; m += 32)
      x[m] = (r > 0 && m < M) ? dec_in[((size_t)b * M + m) * T + (r - 1)] : 0.0f;
    __syncwarp();
    for (int layer = 0; layer < 2; ++layer) {
      const float* W = layer == 0 ? w0 : w1;
      const float* in = layer == 0 ? x : h0;
      const int K = layer == 0 ? M : P;
      const uint8_t* keep = layer == 0 ? keep0 : keep1;
      for (int i = 0; i < P; i += 4) {
        float acc[4][1] = {{0.f}, {0.f}, {0.f}, {0.f}};
        const float* rows[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rows[q] = W + (size_t)min(i + q, P - 1) * K;
        warp_dot<4, 1>(rows, in, K, K >> 2, lane, acc);
        warp_reduce_all<4, 1>(acc);
        if (lane < 4 && i + lane < P) {
          float v = lane == 0 ? acc[0][0] : lane == 1 ? acc[1][0] : lane == 2 ? acc[2][0] : acc[3][0];
          const int col = i + lane;
          const float mult = keep_mult(keep, (size_t)n * P + col, seed, stream_id * 2 + layer, r, b * P + col, thresh, 2.0f);
          v = fmaxf(v, 0.0f) * mult;
          if (layer == 0) { h0[col] = v; if (out0) out0[(size_t)n * P + col] = v; }
          else out[(size_t)n * P + col] = v;
        }
      }
      __syncwarp();
    }
  }
}

// Same computation for the full-size prenet (P = 256), tiled: one block = 64 frame rows x all 256 features, both layers, the
// weights streamed through shared memory once per block instead of once per row (the per-row kernel above moves
// rows x 344 KB through L2: 2.6 ms per stream at 51k rows; this one is bound by its fp32 FMAs).  fp32 throughout.
constexpr int kPtRows = 64, kPtKC = 16, kPtP = 256;
constexpr size_t kPtSmem = ((size_t)kPtP * (kPtRows + 4) + (size_t)kPtKC * (kPtP + 4)) * sizeof(float);
__global__ void __launch_bounds__(256, 2) prenet_tf_tiled_kernel(const float* __restrict__ dec_in, const float* __restrict__ w0,
                                                                 const float* __restrict__ w1, const uint8_t* keep0, const uint8_t* keep1,
                                                                 float* __restrict__ out, float* __restrict__ out0, int B, int T, int M,
                                                                 unsigned long long seed, int stream_id, unsigned thresh) {
  extern __shared__ __align__(16) float sm[];
  float (*in_s)[kPtRows + 4] = reinterpret_cast<float (*)[kPtRows + 4]>(sm);                       // [k][row]: x, then layer-0 output
  float (*w_s)[kPtP + 4] = reinterpret_cast<float (*)[kPtP + 4]>(sm + (size_t)kPtP * (kPtRows + 4));   // [k in chunk][feature]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int n_rows = (T + 1) * B, n0 = blockIdx.x * kPtRows;
  const int Mp = (M + kPtKC - 1) / kPtKC * kPtKC;
  // teacher-forced input rows, transposed: row n = (frame r, utterance b); r = 0 is the go-frame
  for (int i = tid; i < Mp * kPtRows; i += 256) {
    const int m = i / kPtRows, row = i - m * kPtRows, n = n0 + row;
    float v = 0.f;
    if (n < n_rows && m < M) {
      const int r = n / B, b = n - r * B;
      if (r > 0) v = dec_in[((size_t)b * M + m) * T + (r - 1)];
    }
    in_s[m][row] = v;
  }
  for (int layer = 0; layer < 2; ++layer) {
    const float* W = layer == 0 ? w0 : w1;
    const int K = layer == 0 ? M : kPtP, Kp = layer == 0 ? Mp : kPtP;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < Kp; k0 += kPtKC) {
      __syncthreads();                                   // previous chunk consumed (first pass: in_s complete)
      for (int i = tid; i < kPtKC * kPtP; i += 256) {   // W[feature][k0 .. k0+16) -> w_s[k][feature]
        const int f = i >> 4, kl = i & 15;
        w_s[kl][f] = (k0 + kl < K) ? W[(size_t)f * K + k0 + kl] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kl = 0; kl < kPtKC; ++kl) {
        const float4 xa = *reinterpret_cast<const float4*>(&in_s[k0 + kl][ty * 8]);
        const float4 xb = *reinterpret_cast<const float4*>(&in_s[k0 + kl][ty * 8 + 4]);
        const float4 wa = *reinterpret_cast<const float4*>(&w_s[kl][tx * 4]);
        const float4 wb = *reinterpret_cast<const float4*>(&w_s[kl][128 + tx * 4]);
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
    }
    __syncthreads();                                     // every thread is done reading in_s before layer 0 overwrites it
    const uint8_t* keep = layer == 0 ? keep0 : keep1;
    float* dst = layer == 0 ? out0 : out;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = ty * 8 + i, n = n0 + row;
      if (n >= n_rows) {
        if (layer == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) in_s[(j < 4 ? 0 : 128) + tx * 4 + (j & 3)][row] = 0.f;
        }
        continue;
      }
      const int r = n / B, b = n - r * B;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = (j < 4 ? 0 : 128) + tx * 4 + (j & 3);
        const float mult = keep_mult(keep, (size_t)n * kPtP + col, seed, stream_id * 2 + layer, r, b * kPtP + col, thresh, 2.0f);
        v[j] = fmaxf(acc[i][j], 0.0f) * mult;
        if (layer == 0) in_s[col][row] = v[j];
      }
      if (dst) {
        *reinterpret_cast<float4*>(dst + (size_t)n * kPtP + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + (size_t)n * kPtP + 128 + tx * 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
}

__global__ void philox_mask_kernel(unsigned long long seed, int mask_id, int rows, int n, unsigned thresh, uint8_t* out) {
  const size_t total = (size_t)rows * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / n), idx = (int)(i - (size_t)row * n);
    out[i] = philox_keep(seed, mask_id, row, idx, thresh) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int env_int(const char* name, int dflt);
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess)                                                                       \
      return fail(TACO2DEC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));             \
  } while (0)

unsigned keep_threshold(double p_drop) {
  double v = p_drop * 4294967296.0;
  if (v < 0) v = 0;
  if (v > 4294967295.0) v = 4294967295.0;
  return (unsigned)v;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WorkspaceLayout {
  size_t pm[2], pre[2], pre0[2], a_prev[2], a_cum[2], h1, c1, h2, c2, ctx, q, ctl, h2_all, ctx_all, total;
};

}  // namespace

struct taco2dec_handle {
  taco2dec_config cfg;
  int device;
  int num_sms;
  int max_smem_optin;
  bool have_weights;
  taco2dec_weights w;
  int64_t launches;
  int* abort_dev;        // sticky watchdog word (library-owned); every kernel of the handle polls / sets it
  cudaStream_t aux_stream;  // private non-blocking stream: reads of abort_dev that must not wait for the caller's stream
  int batched_fp32;      // 1 = AUTO keeps 2 <= B <= 128 on the generic fp32 kernel
  long long* last_phase_clocks;
  int path_mode;         // TACO2DEC_PATH_*
  int weight_dtype;      // TACO2DEC_W_*
  unsigned char* packed; // latency path: per-LSTM-CTA packed weight streams (library-owned)
  size_t packed_bytes;
  int packed_wbytes;     // 0 = not packed yet
  unsigned long long* packed_off;  // device [NL+1]
  unsigned long long* ll_buf;      // latency path LL exchange region (library-owned)
  size_t ll_bytes;
  int last_path;         // path actually taken by the most recent call (1 generic, 2 latency, 3 tensor)
  bt::Bufs bt_bufs;      // batched tensor path: library-owned tiled operands / partials / state
  bool bt_alloc, bt_tiles_valid;
  bt::Saved cur_sv;      // where the current teacher-forced call keeps activations for backward (null = nowhere)
  pb::PbParams pb;       // persistent batched kernel: library-owned weight tiles, operand buffers, partials, counters
  bool pb_alloc, pb_tiles_valid;
  unsigned char* pb_xpre; size_t pb_xpre_bytes;
  long long* pb_dbg;
  const float* pm_given[2];   // processed memory supplied by the caller for the current call (else null)
  bw::Bufs bw_bufs;      // backward pass: transposed bf16 weight tiles, gate-gradient tiles, GEMM partials
  bool bw_alloc, bw_tiles_valid;
  int* bw_ctl;           // device word: frame counter of the backward graph
  pbw::PbwParams pbw;    // persistent backward kernel: parity-buffered split-K partials, counters
  bool pbw_alloc;
  __half* pbw_h16;       // fp16 copies of memory / processed memory for the attention tasks (grown on demand)
  size_t pbw_h16_elems;
  cudaStream_t cap_stream;   // private stream used only to capture the per-frame CUDA graph
  bool profiling;        // record CUDA events around the persistent launch
  cudaEvent_t ev0, ev1;
  bool ev_valid;
};

namespace {

WorkspaceLayout plan_workspace(const taco2dec_config& c, int B, int T_in, int T_sub, int T, bool tf) {
  WorkspaceLayout L;
  size_t off = 0;
  auto take = [&](size_t n_floats) { size_t o = off; off = align_up(off + n_floats * sizeof(float), 256); return o; };
  const int Ts[2] = {T_in, T_sub};
  for (int s = 0; s < 2; ++s) {
    const bool on = s < c.n_streams;
    L.pm[s] = take(on ? (size_t)B * Ts[s] * c.attn_dim : 0);
    L.pre[s] = take(on ? (tf ? (size_t)(T + 1) * B * c.prenet_dim : (size_t)B * c.prenet_dim) : 0);
    L.pre0[s] = take(on ? (size_t)B * c.prenet_dim : 0);
    L.a_prev[s] = take(on ? (size_t)B * Ts[s] : 0);
    L.a_cum[s] = take(on ? (size_t)B * Ts[s] : 0);
  }
  L.h1 = take((size_t)2 * c.n_streams * B * c.attn_rnn_dim);
  L.c1 = take((size_t)c.n_streams * B * c.attn_rnn_dim);
  L.h2 = take((size_t)2 * B * c.dec_rnn_dim);
  L.c2 = take((size_t)B * c.dec_rnn_dim);
  L.ctx = take((size_t)c.n_streams * B * c.enc_dim);
  L.q = take((size_t)c.n_streams * B * c.attn_dim);
  L.ctl = take(64);
  // tensor path, teacher-forced: h2 / context rows of every frame for the hoisted projection (bt_proj_all)
  const bool rows = tf && B >= 2 && B <= 128;
  L.h2_all = take(rows ? (size_t)(T + 1) * B * c.dec_rnn_dim : 0);
  L.ctx_all = take(rows ? (size_t)(T + 1) * c.n_streams * B * c.enc_dim : 0);
  L.total = off;
  return L;
}

int pick_bt(int B) { return B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : 8; }

size_t persistent_smem_bytes(const taco2dec_config& c, int BT, int T_in, int T_sub) {
  const int S = c.n_streams;
  size_t ka = (size_t)S * BT * (c.prenet_dim + c.enc_dim + c.attn_rnn_dim);
  size_t kc = (size_t)BT * (S * (c.attn_rnn_dim + c.enc_dim) + c.dec_rnn_dim);
  size_t kq = (size_t)S * BT * c.attn_rnn_dim;
  size_t kd = (size_t)BT * (c.dec_rnn_dim + S * c.enc_dim);
  size_t kp = (size_t)S * BT * std::max(c.prenet_dim, c.n_mel);
  const int Tm = std::max(T_in, T_sub);
  const int pad = c.attention == TACO2DEC_ATTN_LSA ? (c.loc_kernel - 1) / 2 : 1;
  size_t att = (size_t)2 * c.attn_dim + 2 * (size_t)Tm + 2 * (size_t)(Tm + 2 * pad) +
               (size_t)c.loc_filters * c.attn_dim + 2 * (size_t)c.loc_kernel * c.loc_filters +
               std::max<size_t>((size_t)(kThreads / (c.enc_dim / 4) > 0 ? kThreads / (c.enc_dim / 4) : 1) * c.enc_dim, 64) + 64;
  size_t m = std::max({ka, kc, kq, kd, kp, att});
  return align_up(m * sizeof(float), 16);
}

template <int BT>
int launch_persistent(taco2dec_handle* h, const Params& p, size_t smem, cudaStream_t st) {
  auto kern = decoder_persistent<BT>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) return fail(TACO2DEC_E_STATE, "persistent kernel does not fit on an SM");
  void* args[] = {(void*)&p};
  if (h->profiling) CUDA_TRY(cudaEventRecord(h->ev0, st));
  CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(h->num_sms), dim3(kThreads), args, smem, st));
  if (h->profiling) { CUDA_TRY(cudaEventRecord(h->ev1, st)); h->ev_valid = true; }
  h->launches++;
  return 0;
}

int check_cfg(const taco2dec_config& c) {
  if (c.n_streams < 1 || c.n_streams > 2) return fail(TACO2DEC_E_ARG, "n_streams must be 1 or 2");
  if (c.attention != TACO2DEC_ATTN_SMA && c.attention != TACO2DEC_ATTN_LSA)
    return fail(TACO2DEC_E_ARG, "attention must be SMA or LSA");
  const int dims4[] = {c.n_mel, c.enc_dim, c.attn_rnn_dim, c.dec_rnn_dim, c.prenet_dim, c.attn_dim};
  for (int d : dims4)
    if (d <= 0 || d % 4) return fail(TACO2DEC_E_ARG, "all feature dims must be positive multiples of 4");
  if (c.attn_rnn_dim != c.dec_rnn_dim)
    return fail(TACO2DEC_E_ARG, "attention_rnn_dim must equal decoder_rnn_dim (lstm_keep layout)");
  if (c.attention == TACO2DEC_ATTN_LSA && (c.loc_filters < 1 || c.loc_filters > 32 || c.loc_kernel % 2 == 0))
    return fail(TACO2DEC_E_ARG, "LSA needs 1..32 location filters and an odd kernel size");
  if (c.enc_dim / 4 > kThreads) return fail(TACO2DEC_E_ARG, "encoder_embedding_dim too large");
  if (c.attn_dim % 32) return fail(TACO2DEC_E_ARG, "attention_dim must be a multiple of 32");
  if (c.p_attention_dropout < 0 || c.p_attention_dropout >= 1 || c.p_decoder_dropout < 0 || c.p_decoder_dropout >= 1)
    return fail(TACO2DEC_E_ARG, "dropout probabilities must be in [0,1)");
  return 0;
}

int fill_common(taco2dec_handle* h, Params& p, int B, int T_in, int T_sub, const float* memory, const float* embeddings,
                const int64_t* mlen, const int64_t* blen, const taco2dec_rng& rng, char* ws, const WorkspaceLayout& L) {
  const taco2dec_config& c = h->cfg;
  memset(&p, 0, sizeof(p));
  p.B = B; p.S = c.n_streams; p.H = c.attn_rnn_dim; p.D = c.dec_rnn_dim; p.E = c.enc_dim; p.P = c.prenet_dim;
  p.A = c.attn_dim; p.M = c.n_mel; p.LF = c.attention == TACO2DEC_ATTN_LSA ? c.loc_filters : 0;
  p.LK = c.attention == TACO2DEC_ATTN_LSA ? c.loc_kernel : 1;
  p.attention = c.attention;
  p.p_att = c.p_attention_dropout; p.p_dec = c.p_decoder_dropout;
  p.thresh_pre = keep_threshold(0.5);
  p.thresh_att = keep_threshold(c.p_attention_dropout);
  p.thresh_dec = keep_threshold(c.p_decoder_dropout);
  p.seed = rng.seed;
  const int Ts[2] = {T_in, T_sub};
  const float* mems[2] = {memory, embeddings};
  const int64_t* lens[2] = {mlen, blen};
  for (int s = 0; s < c.n_streams; ++s) {
    const taco2dec_stream_weights& sw = h->w.stream[s];
    StreamParams& sp = p.st[s];
    sp.pre_w0 = sw.prenet_w0; sp.pre_w1 = sw.prenet_w1; sp.w_ih = sw.arnn_w_ih; sp.w_hh = sw.arnn_w_hh;
    sp.b_ih = sw.arnn_b_ih; sp.b_hh = sw.arnn_b_hh; sp.wq = sw.query_w; sp.wm = sw.memory_w; sp.v = sw.v;
    sp.loc_conv = sw.loc_conv_w; sp.loc_dense = sw.loc_dense_w;
    sp.mem = mems[s]; sp.len = (const long long*)lens[s]; sp.noise = rng.sma_noise[s];
    sp.keep0 = rng.prenet_keep[s][0]; sp.keep1 = rng.prenet_keep[s][1];
    sp.pm = (float*)(ws + L.pm[s]); sp.pre = (float*)(ws + L.pre[s]); sp.pre0 = (float*)(ws + L.pre0[s]);
    sp.a_prev = (float*)(ws + L.a_prev[s]); sp.a_cum = (float*)(ws + L.a_cum[s]);
    sp.Ts = Ts[s];
  }
  p.d_w_ih = h->w.drnn_w_ih; p.d_w_hh = h->w.drnn_w_hh; p.d_b_ih = h->w.drnn_b_ih; p.d_b_hh = h->w.drnn_b_hh;
  p.proj_w = h->w.proj_w; p.proj_b = h->w.proj_b; p.gate_w = h->w.gate_w; p.gate_b = h->w.gate_b;
  p.lstm_keep = rng.lstm_keep;
  p.h1 = (float*)(ws + L.h1); p.c1 = (float*)(ws + L.c1); p.h2 = (float*)(ws + L.h2); p.c2 = (float*)(ws + L.c2);
  p.ctx = (float*)(ws + L.ctx); p.q = (float*)(ws + L.q);
  p.sync_ctr = (unsigned*)(ws + L.ctl);
  p.abort_flag = h->abort_dev;
  p.done_count = (int*)(ws + L.ctl) + 2;
  p.phase_clocks = (long long*)(ws + L.ctl + 64);
  return 0;
}

// ------------------------------------------------------------------------------------------
// Latency path (latency.cuh): eligibility, weight packing, launch
// ------------------------------------------------------------------------------------------
struct LatGeometry {
  int NL, NL1, wbytes, res_budget, na[2];
  size_t smem;
};

constexpr size_t kLatFixedSmem = 31744;   // activation segments, query slice, accumulators, barriers
constexpr size_t kLatDynSmem = 231424;    // 226 KB of the 227 KB opt-in maximum

LatGeometry lat_geometry(const taco2dec_handle* h, int wbytes) {
  LatGeometry g;
  const int S = h->cfg.n_streams;
  // attention CTAs: 8 feature slices for the phoneme stream, 4 for the (about 3x shorter) sub-word stream
  g.na[0] = 8; g.na[1] = S == 2 ? 4 : 0;
  int NL = h->num_sms - g.na[0] - g.na[1] - lat::kAux;
  if (NL > 128) NL = 128;               // 1024 hidden units / 128 CTAs = 8 (decoder) and 16 (attention) per CTA
  NL = NL / (2 * S) * (2 * S);
  g.NL = NL; g.NL1 = NL / S; g.wbytes = wbytes;
  g.res_budget = (int)((kLatDynSmem - kLatFixedSmem) / 128 * 128);
  g.smem = kLatDynSmem;
  return g;
}

bool lat_shape_ok(const taco2dec_handle* h, int B, int T_in, int T_sub) {
  const taco2dec_config& c = h->cfg;
  if (B != 1) return false;
  const bool lsa = c.attention == TACO2DEC_ATTN_LSA;
  if (lsa && (c.loc_filters < 1 || c.loc_filters > 32 || c.loc_kernel % 2 == 0 || c.loc_kernel > 63)) return false;
  if (c.attn_rnn_dim != lat::H || c.dec_rnn_dim != lat::H || c.enc_dim != lat::E || c.prenet_dim != lat::P ||
      c.attn_dim != lat::A || c.n_mel != lat::M)
    return false;
  if (h->num_sms < 84 || (size_t)h->max_smem_optin < kLatDynSmem) return false;
  const LatGeometry g = lat_geometry(h, 4);
  const int Ts[2] = {T_in, T_sub};
  for (int s = 0; s < c.n_streams; ++s) {
    const size_t fs = lat::E / g.na[s], as = lat::A / g.na[s];
    size_t att = (size_t)Ts[s] * (as + fs) + 4 * lat::kThreads + 8 * 32 + 64 + lat::kThreads + 3 * (size_t)Ts[s] + 8;
    if (lsa)    // location term, folded conv . dense weights, zero-padded previous and cumulative weights
      att += (size_t)Ts[s] * as + 2 * (size_t)c.loc_kernel * as + 2 * ((size_t)Ts[s] + c.loc_kernel + 5);
    if (att * sizeof(float) > kLatDynSmem || Ts[s] > lat::kLatTsCap) return false;
  }
  return true;
}

size_t lat_ll_words(const taco2dec_handle* h) {
  const size_t d = lat::kLLDepth, r = lat::kRep;
  return r * (d * 2 * lat::H + d * lat::H + d * 2 * lat::E + d * 2 * (lat::P + 8)) + d * 2 * (size_t)h->num_sms * lat::A +
         d * (lat::M + 16) + d * lat::kAux * (2 * lat::P + 8) + d * 2 * 8 * (size_t)lat::kLatTsCap + 64;
}

int lat_pack_weights(taco2dec_handle* h, const LatGeometry& g, cudaStream_t st) {
  if (h->packed_wbytes == g.wbytes) return 0;
  const int S = h->cfg.n_streams;
  std::vector<unsigned long long> off(g.NL + 1, 0);
  for (int lc = 0; lc < g.NL; ++lc) {
    const int i1 = lc % g.NL1;
    const long long nu1 = (long long)(i1 + 1) * lat::H / g.NL1 - (long long)i1 * lat::H / g.NL1;
    const long long nu2 = (long long)(lc + 1) * lat::H / g.NL - (long long)lc * lat::H / g.NL;
    if (nu1 > lat::kMaxU1 || nu2 > lat::kMaxU2) return fail(TACO2DEC_E_STATE, "too few SMs for the latency path");
    const long long elems = nu1 * 4 * (lat::H + lat::E + lat::P) + nu2 * 4 * ((long long)lat::H + S * lat::H + S * lat::E);
    off[lc + 1] = off[lc] + (unsigned long long)elems * g.wbytes;
  }
  const size_t total = off[g.NL];
  if (h->packed_bytes < total) {
    if (h->packed) CUDA_TRY(cudaFree(h->packed));
    h->packed = nullptr; h->packed_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->packed, total));
    h->packed_bytes = total;
  }
  if (!h->packed_off) CUDA_TRY(cudaMalloc(&h->packed_off, sizeof(unsigned long long) * (h->num_sms + 1)));
  CUDA_TRY(cudaMemcpyAsync(h->packed_off, off.data(), sizeof(unsigned long long) * (g.NL + 1), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // `off` is a host temporary
  lat::PackSrc src;
  for (int s = 0; s < 2; ++s) { src.w_ih[s] = h->w.stream[s].arnn_w_ih; src.w_hh[s] = h->w.stream[s].arnn_w_hh; }
  src.d_w_ih = h->w.drnn_w_ih; src.d_w_hh = h->w.drnn_w_hh;
  if (g.wbytes == 4)
    lat::pack_weights_kernel<float><<<g.NL, 512, 0, st>>>(src, S, g.NL, g.NL1, h->packed_off, h->packed);
  else
    lat::pack_weights_kernel<__half><<<g.NL, 512, 0, st>>>(src, S, g.NL, g.NL1, h->packed_off, h->packed);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  h->packed_wbytes = g.wbytes;
  return 0;
}

int run_latency(taco2dec_handle* h, const Params& gp, cudaStream_t st) {
  const taco2dec_config& c = h->cfg;
  const LatGeometry g = lat_geometry(h, h->weight_dtype == TACO2DEC_W_FP16 ? 2 : 4);
  if (int rc = lat_pack_weights(h, g, st)) return rc;
  const size_t words = lat_ll_words(h);
  if (!h->ll_buf) {
    CUDA_TRY(cudaMalloc(&h->ll_buf, words * sizeof(unsigned long long)));
    h->ll_bytes = words * sizeof(unsigned long long);
  }
  CUDA_TRY(cudaMemsetAsync(h->ll_buf, 0, h->ll_bytes, st));
  lat::LatParams p;
  memset(&p, 0, sizeof(p));
  p.S = c.n_streams; p.NL = g.NL; p.NL1 = g.NL1;
  p.free_running = gp.free_running; p.training = gp.training; p.n_steps = gp.free_running ? gp.max_steps : gp.T;
  p.Tcap = gp.Tcap; p.gate_thr = gp.gate_thr; p.p_att = gp.p_att; p.p_dec = gp.p_dec;
  p.lsa = c.attention == TACO2DEC_ATTN_LSA ? 1 : 0; p.LF = p.lsa ? c.loc_filters : 0; p.LK = p.lsa ? c.loc_kernel : 1;
  p.thresh_pre = gp.thresh_pre; p.thresh_att = gp.thresh_att; p.thresh_dec = gp.thresh_dec; p.seed = gp.seed;
  p.wbytes = g.wbytes; p.packed = h->packed; p.packed_off = h->packed_off;
  p.res_budget = g.res_budget;
  // L2 policy per segment (a b c d e0 e1 f).  fp16: the whole streamed set (37 MB) fits in L2 -> keep all.
  // fp32: 98 MB are streamed per frame; keep a, e0, e1 (65 MB; they sit on the critical path) and let b, c
  // (33 MB; they run while the aux CTAs compute the next prenet) stream through from HBM.  Measured:
  // 29.3 us/frame with no hint, 26.4 keep-all, 25.4 keep a/b/c, 22.2 keep a/e0/e1.
  // fp32 storage: segments e0/e1 (on the critical path right after the h1 exchange) live in tensor memory, a and c are
  // kept in L2 (evict_last), b streams from HBM (evict_first).  Measured sweep on one box, us/frame: TMEM off + keep
  // a/e0/e1 21.6; TMEM(e) + keep a,c 18.5; keep a,b 19.1; keep a 19.5; keep a,b,c 20.6; keep c 23.0; TMEM(b,c) 22.7.
  // default placement: dual-stream fp32 -> f and e1 in tensor memory, e0 in shared memory (3): with the early weight requests
  // the LSTM CTAs have slack in the attention window and f sits on the chain between the context and h2 (13.5 -> 13.0 us/frame
  // on one box); everything else -> e0, e1 (and a when it fits) (1)
  p.use_tmem = (g.wbytes == 4 && c.n_streams == 2) ? 3 : 1;
  { const char* e = getenv("TACO2DEC_TMEM"); if (e) p.use_tmem = atoi(e); }
  p.stream_prefetch = env_int("TACO2DEC_LAT_PREFETCH", 1);
  p.l2_keep_mask = g.wbytes == 2 ? 0x7f : (p.use_tmem == 1 || p.use_tmem == 3 ? 0x05 : p.use_tmem == 4 ? 0x25 : 0x31);
  { const char* e = getenv("TACO2DEC_L2_KEEP_MASK"); if (e) p.l2_keep_mask = (int)strtol(e, nullptr, 0); }
  for (int s = 0; s < c.n_streams; ++s) {
    const StreamParams& sp = gp.st[s];
    lat::LatStream& ls = p.st[s];
    ls.b_ih = sp.b_ih; ls.b_hh = sp.b_hh; ls.wq = sp.wq; ls.v = sp.v; ls.pre_w0 = sp.pre_w0; ls.pre_w1 = sp.pre_w1;
    ls.loc_conv = sp.loc_conv; ls.loc_dense = sp.loc_dense;
    ls.mem = sp.mem; ls.pm = sp.pm; ls.pre_tf = sp.pre; ls.noise = sp.noise; ls.keep0 = sp.keep0; ls.keep1 = sp.keep1;
    ls.align = sp.align; ls.Ts = sp.Ts; ls.len = sp.len; ls.na = g.na[s];
  }
  p.d_b_ih = gp.d_b_ih; p.d_b_hh = gp.d_b_hh; p.proj_w = gp.proj_w; p.proj_b = gp.proj_b; p.gate_w = gp.gate_w;
  p.gate_b = gp.gate_b; p.lstm_keep = gp.lstm_keep; p.mel = gp.mel; p.gate = gp.gate; p.n_frames = gp.n_frames;
  p.reached_max = gp.reached_max;
  unsigned long long* w = h->ll_buf;
  const size_t d = lat::kLLDepth;
  const size_t r = lat::kRep;
  p.ll_h1 = w; w += r * d * 2 * lat::H;
  p.ll_h2 = w; w += r * d * lat::H;
  p.ll_ctx = w; w += r * d * 2 * lat::E;
  p.ll_q = w; w += d * 2 * (size_t)h->num_sms * lat::A;
  p.ll_pre = w; w += r * d * 2 * (lat::P + 8);
  p.ll_mel = w; w += d * (lat::M + 16);
  p.ll_l0 = w; w += d * lat::kAux * (2 * lat::P + 8);
  p.ll_e = w; w += d * 2 * 8 * (size_t)lat::kLatTsCap;
  p.aux_done = (unsigned*)w;
  p.abort_flag = gp.abort_flag;
  p.phase_clocks = gp.phase_clocks;
  p.dbg = nullptr;
  if (env_int("TACO2DEC_LAT_DEBUG", 0)) {        // diagnostics: per-role cycle sums, read with taco2dec_read_debug_stamps
    if (!h->pb_dbg) CUDA_TRY(cudaMalloc(&h->pb_dbg, 256 * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(h->pb_dbg, 0, 256 * sizeof(long long), st));
    p.dbg = h->pb_dbg;
  }
  void* args[] = {(void*)&p};
  // early weight requests for the streamed steps b and c (fp32 storage): only if every LSTM CTA's plan qualifies
  bool pre = p.stream_prefetch != 0 && g.wbytes == 4;
  for (int lc = 0; lc < p.NL && pre; ++lc) {
    lat::StepPlan plan[lat::kSteps];
    lat::lat_build_plan(p, lc, g.wbytes, plan);
    pre = lat::stream_prefetchable(plan[1], lat::kPreK) && lat::stream_prefetchable(plan[2], lat::kPreK) &&
          lat::stream_prefetchable(plan[0], 2 * lat::kPreK) && plan[0].ksplit == 1;      // a: two passes per warp
  }
  void* kern = g.wbytes == 4 ? (pre ? (void*)lat::decoder_latency<4, true> : (void*)lat::decoder_latency<4, false>)
                             : (void*)lat::decoder_latency<2, false>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, lat::kThreads, g.smem));
  if (per_sm < 1) return fail(TACO2DEC_E_STATE, "latency kernel does not fit on an SM");
  if (h->profiling) CUDA_TRY(cudaEventRecord(h->ev0, st));
  CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(h->num_sms), dim3(lat::kThreads), args, g.smem, st));
  if (h->profiling) { CUDA_TRY(cudaEventRecord(h->ev1, st)); h->ev_valid = true; }
  h->launches++;
  h->last_path = TACO2DEC_PATH_LATENCY;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Batched tensor-core path (batched.cuh + gemm_tc.cuh): eligibility, buffers, per-frame launch sequence
// ------------------------------------------------------------------------------------------
size_t attention_smem_bytes(const taco2dec_config& c, int T_in, int T_sub) {
  const int Tm = std::max(T_in, T_sub);
  const int pad = c.attention == TACO2DEC_ATTN_LSA ? (c.loc_kernel - 1) / 2 : 1;
  const size_t nj = (size_t)(kThreads / (c.enc_dim / 4) > 0 ? kThreads / (c.enc_dim / 4) : 1);
  const size_t att = (size_t)2 * c.attn_dim + 2 * (size_t)Tm + 2 * (size_t)(Tm + 2 * pad) + (size_t)c.loc_filters * c.attn_dim +
                     2 * (size_t)c.loc_kernel * c.loc_filters + std::max<size_t>(nj * c.enc_dim, 64) + 64;
  return align_up(att * sizeof(float), 16);
}

bool bt_shape_ok(const taco2dec_handle* h, int B, int T_in, int T_sub) {
  const taco2dec_config& c = h->cfg;
  if (B < 2 || B > 128) return false;   // B = 1 belongs to the latency path; columns up to the next multiple of 16 are padding
  if (c.attn_rnn_dim != bt::H || c.dec_rnn_dim != bt::H || c.enc_dim != bt::E || c.prenet_dim != bt::P ||
      c.attn_dim != bt::A || c.n_mel != bt::M)
    return false;
  return attention_smem_bytes(c, T_in, T_sub) <= (size_t)h->max_smem_optin;
}

int bt_prepare(taco2dec_handle* h, cudaStream_t st) {
  const int S = h->cfg.n_streams;
  bt::Bufs& b = h->bt_bufs;
  const int K2 = S * (bt::H + bt::E) + bt::H;
  if (!h->bt_alloc) {
    const size_t NP = 128;
    CUDA_TRY(cudaMalloc(&b.a1, (size_t)S * 32 * (bt::K1 / 64) * tc::kATileBytes));
    CUDA_TRY(cudaMalloc(&b.a2, (size_t)32 * (K2 / 64) * tc::kATileBytes));
    CUDA_TRY(cudaMalloc(&b.aq, (size_t)S * (bt::H / 64) * tc::kATileBytes));
    CUDA_TRY(cudaMalloc(&b.x1, (size_t)S * (bt::K1 / 64) * NP * 128));
    CUDA_TRY(cudaMalloc(&b.x2, (size_t)(K2 / 64) * NP * 128));
    CUDA_TRY(cudaMalloc(&b.g1, (size_t)S * bt::SPLITS1 * 4 * bt::H * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.g2, (size_t)bt::SPLITS2 * 4 * bt::H * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.gq, (size_t)S * bt::SPLITSQ * 128 * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.c1, (size_t)S * NP * bt::H * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.c2, (size_t)NP * bt::H * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.h2f, (size_t)NP * bt::H * sizeof(float)));
    for (int s = 0; s < S; ++s) {
      CUDA_TRY(cudaMalloc(&b.w0t[s], (size_t)bt::M * bt::P * sizeof(float)));
      CUDA_TRY(cudaMalloc(&b.w1t[s], (size_t)bt::P * bt::P * sizeof(float)));
    }
    h->bt_alloc = true;
  }
  b.K2 = K2;
  if (!h->bt_tiles_valid) {
    for (int s = 0; s < S; ++s) {
      bt::pack_concat_tiles_kernel<<<2048, 256, 0, st>>>(h->w.stream[s].arnn_w_ih, bt::P + bt::E, h->w.stream[s].arnn_w_hh, bt::H,
                                                         4 * bt::H, b.a1 + (size_t)s * 32 * (bt::K1 / 64) * tc::kATileBytes);
      bt::pack_concat_tiles_kernel<<<256, 256, 0, st>>>(h->w.stream[s].query_w, bt::H, nullptr, 0, bt::A,
                                                        b.aq + (size_t)s * (bt::H / 64) * tc::kATileBytes);
    }
    for (int s = 0; s < S; ++s) {
      bt::transpose_kernel<<<(bt::P * bt::M + 255) / 256, 256, 0, st>>>(h->w.stream[s].prenet_w0, bt::P, bt::M, b.w0t[s]);
      bt::transpose_kernel<<<(bt::P * bt::P + 255) / 256, 256, 0, st>>>(h->w.stream[s].prenet_w1, bt::P, bt::P, b.w1t[s]);
    }
    bt::pack_concat_tiles_kernel<<<4096, 256, 0, st>>>(h->w.drnn_w_ih, S * (bt::H + bt::E), h->w.drnn_w_hh, bt::H, 4 * bt::H, b.a2);
    CUDA_TRY(cudaGetLastError());
    h->launches += 2 * S + 1;
    h->bt_tiles_valid = true;
  }
  return 0;
}

template <int NPAD>
int bt_run_frames(taco2dec_handle* h, const Params& p, size_t att_smem, cudaStream_t st) {
  bt::Bufs bf = h->bt_bufs;
  bf.NPAD = NPAD;
  bf.sv = h->cur_sv;
  const int S = p.S, B = p.B;
  const int n_steps = p.free_running ? p.max_steps : p.T;
  // slot 0 of the stored state arrays = the zero initial state (model.py:237-256)
  if (bf.sv.h1) CUDA_TRY(cudaMemsetAsync(bf.sv.h1, 0, (size_t)S * B * bt::H * sizeof(float), st));
  if (bf.sv.ctx) CUDA_TRY(cudaMemsetAsync(bf.sv.ctx, 0, (size_t)S * B * bt::E * sizeof(float), st));
  if (bf.sv.h2) CUDA_TRY(cudaMemsetAsync(bf.sv.h2, 0, (size_t)B * bt::H * sizeof(float), st));
  const bool hoist_proj = !p.free_running && bf.sv.h2 && bf.sv.ctx && !getenv("TACO2DEC_NO_HOIST");
  CUDA_TRY(cudaMemsetAsync(bf.x1, 0, (size_t)S * (bt::K1 / 64) * NPAD * 128, st));
  CUDA_TRY(cudaMemsetAsync(bf.x2, 0, (size_t)(bf.K2 / 64) * NPAD * 128, st));
  bt::bt_init_kernel<<<h->num_sms, 256, 0, st>>>(p, bf);
  CUDA_TRY(cudaFuncSetAttribute(bt::bt_attention<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
  int max_ts = 0;
  for (int s = 0; s < S; ++s) max_ts = std::max(max_ts, p.st[s].Ts);
  const size_t sma_smem = bt::sma_smem_floats(max_ts) * sizeof(float);
  CUDA_TRY(cudaFuncSetAttribute(bt::bt_attention_sma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sma_smem));
  const int* done = p.free_running ? p.done_count : nullptr;
  tc::GemmParams g1{bf.a1, bf.x1, bf.g1, 4 * bt::H, bt::K1, bt::SPLITS1, S, (long long)(bt::K1 / 64) * NPAD * 128, 0, 0, done, B};
  tc::GemmParams gq{bf.aq, bf.x2, bf.gq, 128, bt::H, bt::SPLITSQ, S, 0, 0, (bt::H + bt::E) / 64, done, B};
  tc::GemmParams g2{bf.a2, bf.x2, bf.g2, 4 * bt::H, bf.K2, bt::SPLITS2, 1, 0, 0, 0, done, B};
  CUDA_TRY(tc::prepare_gemm<NPAD>());
  // One frame = 9 kernels free-running, 6 teacher-forced; the frame index is read from device memory, so the sequence
  // is captured ONCE into a CUDA graph and replayed n_steps times (one graph launch per frame).
  // Capture happens on a private stream (the caller's stream may be the legacy default stream, which cannot
  // be captured); the instantiated graph is launched on the caller's stream.
  int* t_ptr = p.done_count + 1;                      // control block word, zeroed by run_common
  if (!h->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  cudaStream_t cs = h->cap_stream;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  CUDA_TRY(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
  if (hoist_proj) {   // frame 0's prenet rows; later frames get theirs from bt_pointwise1 of the frame before
    bt::bt_prenet_tf_to_x1<<<(S * B * bt::P + 255) / 256, 256, 0, st>>>(p, bf, t_ptr);
    h->launches++;
  }
  // teacher-forced + hoisted projection: 6 kernels per frame (the LSTM pointwise kernels also move the prenet rows and
  // the frame counter); otherwise 9
  const int tf_merge = hoist_proj ? 1 : 0;
  if (p.free_running) bt::bt_prenet_fr<<<S * ((B + bt::kPreBT - 1) / bt::kPreBT), 1024, 0, cs>>>(p, bf, t_ptr);
  else if (!tf_merge) bt::bt_prenet_tf_to_x1<<<(S * B * bt::P + 255) / 256, 256, 0, cs>>>(p, bf, t_ptr);
  cudaError_t ce = tc::launch_gemm<NPAD>(g1, cs);
  bt::bt_pointwise1<<<(S * B * bt::H + 255) / 256, 256, 0, cs>>>(p, bf, t_ptr, tf_merge);
  if (ce == cudaSuccess) ce = tc::launch_gemm<NPAD>(gq, cs);
  if (p.attention == TACO2DEC_ATTN_SMA) bt::bt_attention_sma<<<S * B, kThreads, sma_smem, cs>>>(p, bf, t_ptr);
  else bt::bt_attention<true><<<S * B, kThreads, att_smem, cs>>>(p, bf, t_ptr);
  if (ce == cudaSuccess) ce = tc::launch_gemm<NPAD>(g2, cs);
  bt::bt_pointwise2<<<(B * bt::H + 255) / 256, 256, 0, cs>>>(p, bf, t_ptr, tf_merge);
  if (!hoist_proj) bt::bt_proj<<<((bt::M + 1) * B * 32 + 255) / 256, 256, 0, cs>>>(p, bf, t_ptr);
  if (!tf_merge) bt::bt_advance<<<1, 1, 0, cs>>>(t_ptr);
  const cudaError_t ee = cudaStreamEndCapture(cs, &graph);
  if (ce != cudaSuccess || ee != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    return fail(TACO2DEC_E_CUDA, std::string("graph capture of the frame sequence failed: ") +
                                     cudaGetErrorString(ce != cudaSuccess ? ce : ee));
  }
  CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
  for (int t = 0; t < n_steps; ++t) {
    const cudaError_t le = cudaGraphLaunch(exec, st);
    if (le != cudaSuccess) {
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
      return fail(TACO2DEC_E_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(le));
    }
  }
  h->launches += (hoist_proj ? 6LL : 9LL) * n_steps;
  // the executable graph may be destroyed once its launches are enqueued; CUDA keeps it alive until they finish
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  if (hoist_proj) {
    bt::bt_proj_all<<<(p.T * B + bt::kPaRows - 1) / bt::kPaRows, 256, 0, st>>>(p, bf);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  h->last_path = TACO2DEC_PATH_TENSOR_GRAPH;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Persistent batched kernel (persist.cuh): eligibility, buffers, launch
// ------------------------------------------------------------------------------------------
struct PbGeometry { int npad, stages_a, stages_x, n_res, n_tm; size_t smem; };

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

bool pb_geometry(const taco2dec_handle* h, int B, int T_in, int T_sub, int fr, PbGeometry* out) {
  const int npad = B <= 16 ? 16 : B <= 32 ? 32 : B <= 64 ? 64 : 128;
  const int max_ts = std::max(T_in, h->cfg.n_streams == 2 ? T_sub : 0);
  const size_t budget = (size_t)h->max_smem_optin - 2048;          // static shared: barriers, phase stamps, placement table
  const size_t fixed = pb::smem_plan(npad, 0, 0, 0, max_ts, fr).total;
  const size_t xt = (size_t)npad * 128, at = tc::kATileBytes;
  // tensor memory: 512 columns, 2 * npad of them accumulators, 32 per resident weight tile
  int n_tm = std::min(env_int("TACO2DEC_PB_TMEM", 64), (512 - 2 * npad) / 32);
  n_tm = std::max(0, std::min(n_tm, pb::kMaxTiles));
  // ring depths are PER gate product (two weight rings, two activation rings)
  int stages_x = std::max(2, std::min(8, env_int("TACO2DEC_PB_STAGES_X", npad <= 32 ? 8 : npad == 64 ? 4 : 3)));
  int stages_a = std::max(2, std::min(8, env_int("TACO2DEC_PB_STAGES_A", 2)));
  while (fixed + 2 * (stages_x * xt + stages_a * at) + 1024 > budget && stages_x > 2) --stages_x;
  if (fixed + 2 * (stages_x * xt + stages_a * at) + 1024 > budget) return false;
  int n_res = (int)((budget - fixed - 2 * (stages_x * xt + stages_a * at) - 1024) / at);
  n_res = std::max(0, std::min(std::min(n_res, env_int("TACO2DEC_PB_RES", 64)), pb::kMaxTiles - n_tm));
  out->npad = npad; out->stages_a = stages_a; out->stages_x = stages_x; out->n_res = n_res; out->n_tm = n_tm;
  out->smem = pb::smem_plan(npad, stages_a, stages_x, n_res, max_ts, fr).total;
  return out->smem <= budget;
}

bool pb_shape_ok(const taco2dec_handle* h, int B, int T_in, int T_sub, int fr) {
  if (!bt_shape_ok(h, B, T_in, T_sub)) return false;
  if (h->cfg.attention != TACO2DEC_ATTN_SMA || h->num_sms < pb::kCtas) return false;
  if (getenv("TACO2DEC_NO_PERSIST")) return false;
  PbGeometry g;
  return pb_geometry(h, B, T_in, T_sub, fr, &g);
}

int pb_prepare(taco2dec_handle* h, cudaStream_t st) {
  const int S = h->cfg.n_streams;
  pb::PbParams& b = h->pb;
  if (!h->pb_alloc) {
    const size_t NP = 128;
    unsigned char* wt = nullptr;
    CUDA_TRY(cudaMalloc(&wt, (size_t)pb::kCtas * pb::kMaxTiles * tc::kATileBytes));
    b.wt = wt;
    CUDA_TRY(cudaMalloc(&b.x1, (size_t)2 * 2 * 28 * NP * 128));
    CUDA_TRY(cudaMalloc(&b.x2, (size_t)2 * 64 * NP * 128));
    CUDA_TRY(cudaMalloc(&b.part1, (size_t)2 * 32 * 2 * 128 * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.part2, (size_t)32 * 4 * 128 * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.qpart, (size_t)2 * 32 * NP * pb::A * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.melp, (size_t)32 * NP * pb::kMelPad * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.ctxp, (size_t)pb::kSlices * NP * pb::kMelPad * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.melx, (size_t)NP * pb::M * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.l0x, (size_t)2 * NP * pb::P * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.flags, (size_t)pb::F_COUNT * pb::kFlagStride * sizeof(unsigned)));
    h->pb_alloc = true;
  }
  if (!h->pb_tiles_valid) {
    pb::PackSrc src;
    for (int s = 0; s < 2; ++s) { src.w_ih[s] = h->w.stream[s].arnn_w_ih; src.w_hh[s] = h->w.stream[s].arnn_w_hh; }
    src.d_w_ih = h->w.drnn_w_ih; src.d_w_hh = h->w.drnn_w_hh;
    pb::pb_pack_weights<<<pb::kCtas, 512, 0, st>>>(src, S, const_cast<unsigned char*>(b.wt));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    h->pb_tiles_valid = true;
  }
  return 0;
}

__global__ void pb_init_kernel(Params p) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
  for (int i = gtid; i < p.S * p.B * p.E; i += n) p.ctx[i] = 0.f;
  for (int s = 0; s < p.S; ++s) {
    const int Ts = p.st[s].Ts;
    for (int i = gtid; i < p.B * Ts; i += n) p.st[s].a_prev[i] = (i % Ts) == 0 ? 1.0f : 0.0f;    // attention.py:324-328
  }
  if (p.free_running) {
    for (int i = gtid; i < p.B; i += n) { p.n_frames[i] = 0; p.reached_max[i] = 0; }
    if (gtid == 0) *p.done_count = 0;
  }
}

template <int NPAD>
int pb_launch(taco2dec_handle* h, const Params& p, const PbGeometry& g, cudaStream_t st) {
  const int S = p.S, B = p.B;
  pb::PbParams q = h->pb;
  q.sv = h->cur_sv;
  q.stages_a = g.stages_a; q.stages_x = g.stages_x; q.n_res = g.n_res; q.n_tm = g.n_tm;
  q.dbg = nullptr;
  if (getenv("TACO2DEC_PB_DEBUG")) {     // diagnostics: clock stamps of one CTA during one frame (tools/pb_timeline.py)
    if (!h->pb_dbg) CUDA_TRY(cudaMalloc(&h->pb_dbg, 256 * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(h->pb_dbg, 0, 256 * sizeof(long long), st));
    q.dbg = h->pb_dbg; q.dbg_frame = env_int("TACO2DEC_PB_DEBUG", 20); q.dbg_cta = env_int("TACO2DEC_PB_DEBUG_CTA", 0);
  }
  const size_t xt = (size_t)NPAD * 128;
  CUDA_TRY(cudaMemsetAsync(q.flags, 0, (size_t)pb::F_COUNT * pb::kFlagStride * sizeof(unsigned), st));
  CUDA_TRY(cudaMemsetAsync(q.x1, 0, (size_t)2 * S * 28 * xt, st));
  CUDA_TRY(cudaMemsetAsync(q.x2, 0, (size_t)2 * (S == 2 ? 64 : 40) * xt, st));
  if (q.sv.h1) CUDA_TRY(cudaMemsetAsync(q.sv.h1, 0, (size_t)S * B * bt::H * sizeof(float), st));
  if (q.sv.ctx) CUDA_TRY(cudaMemsetAsync(q.sv.ctx, 0, (size_t)S * B * bt::E * sizeof(float), st));
  if (q.sv.h2) CUDA_TRY(cudaMemsetAsync(q.sv.h2, 0, (size_t)B * bt::H * sizeof(float), st));
  pb_init_kernel<<<h->num_sms, 256, 0, st>>>(p);
  h->launches++;
  if (!p.free_running) {      // hoisted prenet rows -> per-frame fp16 operand tiles
    const size_t need = (size_t)p.T * S * 4 * xt;
    if (h->pb_xpre_bytes < need) {
      CUDA_TRY(cudaStreamSynchronize(st));
      if (h->pb_xpre) CUDA_TRY(cudaFree(h->pb_xpre));
      h->pb_xpre = nullptr; h->pb_xpre_bytes = 0;
      CUDA_TRY(cudaMalloc(&h->pb_xpre, need));
      h->pb_xpre_bytes = need;
    }
    if (B < NPAD) CUDA_TRY(cudaMemsetAsync(h->pb_xpre, 0, need, st));
    for (int s = 0; s < S; ++s) {
      const size_t total = (size_t)p.T * B * (bt::P / 8);
      pb::pb_prenet_tiles<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(p.st[s].pre, p.T, B, NPAD, S, s, h->pb_xpre);
      h->launches++;
    }
    q.xpre = h->pb_xpre;
  }
  auto kern = pb::decoder_batched_persistent<NPAD>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pb::kThreads, g.smem));
  if (per_sm < 1) return fail(TACO2DEC_E_STATE, "persistent batched kernel does not fit on an SM");
  void* args[] = {(void*)&p, (void*)&q};
  if (h->profiling) CUDA_TRY(cudaEventRecord(h->ev0, st));
  CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(pb::kCtas), dim3(pb::kThreads), args, g.smem, st));
  if (h->profiling) { CUDA_TRY(cudaEventRecord(h->ev1, st)); h->ev_valid = true; }
  h->launches++;
  if (!p.free_running) {      // mel / gate of all frames in one pass over the stored h2 / context rows (model.py:382-388)
    bt::Bufs bf;
    memset(&bf, 0, sizeof(bf));
    bf.sv = h->cur_sv;
    bt::bt_proj_all<<<(p.T * B + bt::kPaRows - 1) / bt::kPaRows, 256, 0, st>>>(p, bf);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  h->last_path = TACO2DEC_PATH_TENSOR;
  return 0;
}

int run_persistent_batched(taco2dec_handle* h, const Params& p, int T_in, int T_sub, cudaStream_t st) {
  PbGeometry g;
  if (!pb_geometry(h, p.B, T_in, T_sub, p.free_running, &g)) return fail(TACO2DEC_E_STATE, "persistent batched kernel: shared-memory plan does not fit");
  if (int rc = pb_prepare(h, st)) return rc;
  switch (g.npad) {
    case 16: return pb_launch<16>(h, p, g, st);
    case 32: return pb_launch<32>(h, p, g, st);
    case 64: return pb_launch<64>(h, p, g, st);
    default: return pb_launch<128>(h, p, g, st);
  }
}

int run_batched(taco2dec_handle* h, const Params& p, int T_in, int T_sub, cudaStream_t st) {
  if (h->path_mode != TACO2DEC_PATH_TENSOR_GRAPH && pb_shape_ok(h, p.B, T_in, T_sub, p.free_running))
    return run_persistent_batched(h, p, T_in, T_sub, st);
  if (int rc = bt_prepare(h, st)) return rc;
  const size_t att = attention_smem_bytes(h->cfg, T_in, T_sub);
  if (p.B <= 16) return bt_run_frames<16>(h, p, att, st);
  if (p.B <= 32) return bt_run_frames<32>(h, p, att, st);
  if (p.B <= 64) return bt_run_frames<64>(h, p, att, st);
  return bt_run_frames<128>(h, p, att, st);
}

// ------------------------------------------------------------------------------------------
// Backward pass (backward.cuh): buffer layouts, transposed weight tiles, reverse-time frame graph
// ------------------------------------------------------------------------------------------
taco2dec_saved_layout plan_saved(const taco2dec_config& c, int B, int T_in, int T_sub, int T) {
  taco2dec_saved_layout L;
  memset(&L, 0, sizeof(L));
  size_t off = 0;
  auto take = [&](size_t n_floats) { size_t o = off; off = align_up(off + n_floats * sizeof(float), 256); return o; };
  const int S = c.n_streams, H = c.attn_rnn_dim, D = c.dec_rnn_dim;
  const int Ts[2] = {T_in, T_sub};
  L.gates1 = take((size_t)T * S * 5 * H * B);
  L.gates2 = take((size_t)T * 5 * D * B);
  L.h1 = take((size_t)(T + 1) * S * B * H);
  L.ctx = take((size_t)(T + 1) * S * B * c.enc_dim);
  L.h2 = take((size_t)(T + 1) * B * D);
  L.q = take((size_t)T * S * B * c.attn_dim);
  for (int s = 0; s < 2; ++s) {
    const bool on = s < S;
    L.p[s] = take(on ? (size_t)T * B * Ts[s] : 0);
    L.pm[s] = take(on ? (size_t)B * Ts[s] * c.attn_dim : 0);
    L.pre[s] = take(on ? (size_t)(T + 1) * B * c.prenet_dim : 0);
    L.pre0[s] = take(on ? (size_t)(T + 1) * B * c.prenet_dim : 0);
  }
  L.total = off;
  return L;
}

struct GradScratch { size_t dalpha[2], dcum[2], dz[2], dc1, dc2, zero_begin, zero_end, dyh, dyc; };

taco2dec_grad_layout plan_grads(const taco2dec_config& c, int B, int T_in, int T_sub, int T, GradScratch* gs) {
  taco2dec_grad_layout L;
  memset(&L, 0, sizeof(L));
  size_t off = 0;
  auto take = [&](size_t n_floats) { size_t o = off; off = align_up(off + n_floats * sizeof(float), 256); return o; };
  const int S = c.n_streams, H = c.attn_rnn_dim, D = c.dec_rnn_dim;
  const int Ts[2] = {T_in, T_sub};
  L.dg1 = take((size_t)S * T * B * 4 * H);
  L.dg2 = take((size_t)T * B * 4 * D);
  L.dq = take((size_t)S * T * B * c.attn_dim);
  L.dctx = take((size_t)S * T * B * c.enc_dim);
  L.dpre = take((size_t)S * T * B * c.prenet_dim);
  // accumulators and carries: zeroed by taco2dec_backward before the frame loop
  const size_t zero_begin = off;
  L.dv = take((size_t)S * B * c.attn_dim);
  for (int s = 0; s < 2; ++s) L.dpm[s] = take(s < S ? (size_t)B * Ts[s] * c.attn_dim : 0);
  const bool lsa = c.attention == TACO2DEC_ATTN_LSA;
  L.dloc_dense = take(lsa ? (size_t)S * B * c.attn_dim * c.loc_filters : 0);
  L.dloc_conv = take(lsa ? (size_t)S * B * c.loc_filters * 2 * c.loc_kernel : 0);
  L.scratch = off;
  GradScratch g;
  for (int s = 0; s < 2; ++s) g.dalpha[s] = take(s < S ? (size_t)B * Ts[s] : 0);
  for (int s = 0; s < 2; ++s) g.dcum[s] = take((lsa && s < S) ? (size_t)B * Ts[s] : 0);
  g.dc1 = take((size_t)S * B * H);
  g.dc2 = take((size_t)B * D);
  g.zero_begin = zero_begin;
  g.zero_end = off;
  g.dyh = take((size_t)T * D * B);
  g.dyc = take((size_t)T * B * S * c.enc_dim);
  for (int s = 0; s < 2; ++s) g.dz[s] = take((lsa && s < S) ? (size_t)B * Ts[s] * c.attn_dim : 0);
  L.total = off;
  if (gs) *gs = g;
  return L;
}

bool bw_shape_ok(const taco2dec_handle* h, int B, int T_in, int T_sub) {
  if (!bt_shape_ok(h, B, T_in, T_sub)) return false;
  if (h->cfg.attention == TACO2DEC_ATTN_SMA) return true;
  // LSA backward: compiled for the default location layer; its shared-memory plan grows with the memory length
  return h->cfg.loc_filters == bw::kLF && h->cfg.loc_kernel == bw::kLK &&
         bw::bw_lsa_smem_floats(std::max(T_in, T_sub)) * sizeof(float) <= (size_t)h->max_smem_optin;
}

int bw_prepare(taco2dec_handle* h, cudaStream_t st) {
  const int S = h->cfg.n_streams;
  bw::Bufs& b = h->bw_bufs;
  const int K2 = S * (bt::H + bt::E) + bt::H;
  if (!h->bw_alloc) {
    const size_t NP = 128;
    CUDA_TRY(cudaMalloc(&b.a1t, (size_t)S * (bt::K1 / 128) * (bw::G / 64) * tc::kATileBytes));
    CUDA_TRY(cudaMalloc(&b.a2t, (size_t)(K2 / 128) * (bw::G / 64) * tc::kATileBytes));
    CUDA_TRY(cudaMalloc(&b.dg1, (size_t)S * (bw::G / 64) * NP * 128));
    CUDA_TRY(cudaMalloc(&b.dg2, (size_t)(bw::G / 64) * NP * 128));
    CUDA_TRY(cudaMalloc(&b.dx1, (size_t)S * bw::SPLITSB1 * bt::K1 * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&b.dx2, (size_t)bw::SPLITSB2 * K2 * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->bw_ctl, 64));
    h->bw_alloc = true;
  }
  b.K2 = K2;
  if (!h->bw_tiles_valid) {
    for (int s = 0; s < S; ++s)
      bw::pack_concat_tiles_T_kernel<<<2048, 256, 0, st>>>(h->w.stream[s].arnn_w_ih, bt::P + bt::E, h->w.stream[s].arnn_w_hh, bt::H,
                                                           bw::G, b.a1t + (size_t)s * (bt::K1 / 128) * (bw::G / 64) * tc::kATileBytes);
    bw::pack_concat_tiles_T_kernel<<<4096, 256, 0, st>>>(h->w.drnn_w_ih, S * (bt::H + bt::E), h->w.drnn_w_hh, bt::H, bw::G, b.a2t);
    CUDA_TRY(cudaGetLastError());
    h->launches += S + 1;
    h->bw_tiles_valid = true;
  }
  return 0;
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

template <int NPAD>
int bw_run_frames(taco2dec_handle* h, const Params& p, const bw::Grads& g, cudaStream_t st) {
  bw::Bufs bb = h->bw_bufs;
  bb.NPAD = NPAD;
  const int S = p.S, B = p.B;
  // frame T-1 has no successor: every carry starts at zero
  CUDA_TRY(cudaMemsetAsync(bb.dg1, 0, (size_t)S * (bw::G / 64) * NPAD * 128, st));
  CUDA_TRY(cudaMemsetAsync(bb.dg2, 0, (size_t)(bw::G / 64) * NPAD * 128, st));
  CUDA_TRY(cudaMemsetAsync(bb.dx1, 0, (size_t)S * bw::SPLITSB1 * bt::K1 * NPAD * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(bb.dx2, 0, (size_t)bw::SPLITSB2 * bb.K2 * NPAD * sizeof(float), st));
  int* t_ptr = h->bw_ctl;
  set_int_kernel<<<1, 1, 0, st>>>(t_ptr, p.T - 1);
  bw::bw_dy_all<<<dim3((p.T * B + bw::kDyRows - 1) / bw::kDyRows, (bt::H + S * bt::E) / bw::kDyCols), 256, 0, st>>>(p, g);
  int max_ts = 0;
  for (int s = 0; s < S; ++s) max_ts = std::max(max_ts, p.st[s].Ts);
  const bool lsa = p.attention == TACO2DEC_ATTN_LSA;
  const size_t att_smem = (lsa ? bw::bw_lsa_smem_floats(max_ts) : bw::bw_attention_smem_floats(max_ts)) * sizeof(float);
  if (lsa) CUDA_TRY(cudaFuncSetAttribute(bw::bw_attention_lsa, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
  else CUDA_TRY(cudaFuncSetAttribute(bw::bw_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
  tc::GemmParams gb2{bb.a2t, bb.dg2, bb.dx2, bb.K2, bw::G, bw::SPLITSB2, 1, 0, 0, 0, nullptr, 0, tc::kFmtBF16};
  tc::GemmParams gb1{bb.a1t, bb.dg1, bb.dx1, bt::K1, bw::G, bw::SPLITSB1, S, (long long)(bw::G / 64) * NPAD * 128, 0, 0,
                     nullptr, 0, tc::kFmtBF16};
  CUDA_TRY(tc::prepare_gemm<NPAD>());
  if (!h->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  cudaStream_t cs = h->cap_stream;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  CUDA_TRY(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
  const int pw_blocks = ((B + bw::kPwB - 1) / bw::kPwB) * (bt::H / bw::kPwJ);
  bw::bw_pointwise2<<<pw_blocks, 256, 0, cs>>>(p, bb, g, t_ptr);
  cudaError_t ce = tc::launch_gemm<NPAD>(gb2, cs);
  if (lsa) bw::bw_attention_lsa<<<S * B, bw::kBwThreads, att_smem, cs>>>(p, bb, g, t_ptr);
  else bw::bw_attention<<<S * B, bw::kBwThreads, att_smem, cs>>>(p, bb, g, t_ptr);
  bw::bw_pointwise1<<<S * pw_blocks, 256, 0, cs>>>(p, bb, g, t_ptr);
  if (ce == cudaSuccess) ce = tc::launch_gemm<NPAD>(gb1, cs);
  const cudaError_t ee = cudaStreamEndCapture(cs, &graph);
  if (ce != cudaSuccess || ee != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    return fail(TACO2DEC_E_CUDA, std::string("graph capture of the backward frame sequence failed: ") +
                                     cudaGetErrorString(ce != cudaSuccess ? ce : ee));
  }
  CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
  for (int t = 0; t < p.T; ++t) {
    const cudaError_t le = cudaGraphLaunch(exec, st);
    if (le != cudaSuccess) {
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
      return fail(TACO2DEC_E_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(le));
    }
  }
  h->launches += 5LL * p.T + 3;
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  bw::bw_save_dpre0<<<(S * B * bt::P + 255) / 256, 256, 0, st>>>(p, bb, g);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Persistent backward kernel (persist_bwd.cuh): eligibility, buffers, launch
// ------------------------------------------------------------------------------------------
struct PbwGeometry { int npad, stages_a0, stages_a1, stages_x0, stages_x1, n_res, n_tm; size_t smem; };

bool pbw_geometry(const taco2dec_handle* h, int B, int T_in, int T_sub, PbwGeometry* out) {
  const int npad = B <= 16 ? 16 : B <= 32 ? 32 : 64;
  const int max_ts = std::max(T_in, h->cfg.n_streams == 2 ? T_sub : 0);
  const size_t budget = (size_t)h->max_smem_optin - 2048;
  const size_t fixed = pbw::smem_plan(npad, 0, 0, 0, max_ts).total;
  const size_t xt = (size_t)npad * 128, at = tc::kATileBytes;
  int n_tm = std::min(env_int("TACO2DEC_PBW_TMEM", 64), (512 - 2 * npad) / 32);
  n_tm = std::max(0, std::min(n_tm, 32));
  // The attention-LSTM product (0) is the one on the critical chain.  Measured on its timeline (tools/pbw_phases.py with
  // TACO2DEC_PBW_DEBUG): one k-block costs ~0.47 kcyc of MMA issue and ~1.3 kcyc of TMA latency, so with two activation slots the
  // product ran at ~1 kcyc per k-block.  Plan: all 16 weight tiles of product 0 resident (tensor memory first, then shared memory)
  // -> no weight ring for it; the shared memory saved goes to its activation ring; product 1 (a frame of slack) streams everything
  // through two slots each.
  int n_res = std::max(0, pbw::kKb - n_tm);
  int sa0 = 0, sa1 = 2, sx0 = std::max(2, std::min(8, env_int("TACO2DEC_PBW_STAGES_X0", 8))), sx1 = 2;
  auto need = [&]() { return fixed + (size_t)(sx0 + sx1) * xt + (size_t)(sa0 + sa1 + n_res) * at + 1024; };
  while (need() > budget && sx0 > 2) --sx0;
  while (need() > budget && n_res > 0) { --n_res; sa0 = 2; }      // product 0 then streams its last tiles
  if (need() > budget) return false;
  while (need() + at <= budget && n_tm + n_res < 32) ++n_res;     // leftovers: resident tiles of product 1
  out->npad = npad; out->stages_a0 = sa0; out->stages_a1 = sa1; out->stages_x0 = sx0; out->stages_x1 = sx1; out->n_res = n_res; out->n_tm = n_tm;
  out->smem = pbw::smem_plan(npad, sa0 + sa1, sx0 + sx1, n_res, max_ts).total;
  return out->smem <= budget;
}

bool pbw_shape_ok(const taco2dec_handle* h, int B, int T_in, int T_sub) {
  if (h->cfg.attention != TACO2DEC_ATTN_SMA || h->num_sms < pbw::kCtas || B < 2 || B > 64) return false;
  if (getenv("TACO2DEC_NO_PERSIST_BWD")) return false;
  PbwGeometry g;
  return pbw_geometry(h, B, T_in, T_sub, &g);
}

template <int NPAD>
int pbw_run(taco2dec_handle* h, const Params& p, const bw::Grads& g, const PbwGeometry& geo, cudaStream_t st) {
  const int S = p.S, B = p.B;
  const bw::Bufs& bb = h->bw_bufs;
  if (!h->pbw_alloc) {
    const size_t NP = 128;
    CUDA_TRY(cudaMalloc(&h->pbw.dx1, (size_t)2 * 2 * pbw::kSplits * bt::K1 * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->pbw.dx2, (size_t)2 * pbw::kSplits * (2 * (bt::H + bt::E) + bt::H) * NP * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->pbw.flags, (size_t)pbw::F_COUNT * pbw::kFlagStride * sizeof(unsigned) + 16 * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&h->pbw.dxc1, (size_t)2 * 2 * pbw::kSplits * NP * bt::E * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->pbw.dxc2, (size_t)2 * pbw::kSplits * 2 * NP * bt::E * sizeof(float)));
    h->pbw_alloc = true;
  }
  pbw::PbwParams q = h->pbw;
  q.a1t = bb.a1t; q.a2t = bb.a2t; q.dg1t = bb.dg1; q.dg2t = bb.dg2; q.K2 = bb.K2;
  {   // attention sub-tasks: whole tasks by default.  Cutting the long memories into position chunks (TACO2DEC_PBW_CHUNK=<positions>)
      // balances the streams but was slower when measured (53 vs 45 us/frame): a task is a chain of dependent load phases, not a
      // byte stream, and every sub-task pays the chain again.
    int max_ts = 0;
    for (int s = 0; s < S; ++s) max_ts = std::max(max_ts, p.st[s].Ts);
    q.att_chunk = std::min(480, std::max(16, env_int("TACO2DEC_PBW_CHUNK", max_ts)));     // a warp owns <= 30 consecutive positions
  }
  {   // fp16 copies of the attention operands (read every frame by the attention tasks)
    size_t need = 0;
    for (int s = 0; s < S; ++s) need += (size_t)B * p.st[s].Ts * (bt::E + bt::A);
    if (h->pbw_h16_elems < need) {
      CUDA_TRY(cudaStreamSynchronize(st));
      if (h->pbw_h16) CUDA_TRY(cudaFree(h->pbw_h16));
      h->pbw_h16 = nullptr; h->pbw_h16_elems = 0;
      CUDA_TRY(cudaMalloc(&h->pbw_h16, need * sizeof(__half)));
      h->pbw_h16_elems = need;
    }
    __half* w16 = h->pbw_h16;
    for (int s = 0; s < S; ++s) {
      const size_t nm = (size_t)B * p.st[s].Ts * bt::E, np = (size_t)B * p.st[s].Ts * bt::A;
      pbw::to_half_kernel<<<(unsigned)std::min<size_t>((nm / 4 + 255) / 256, 2048), 256, 0, st>>>(p.st[s].mem, w16, nm);
      q.mem16[s] = w16; w16 += nm;
      pbw::to_half_kernel<<<(unsigned)std::min<size_t>((np / 4 + 255) / 256, 2048), 256, 0, st>>>(p.st[s].pm, w16, np);
      q.pm16[s] = w16; w16 += np;
    }
    h->launches += 2 * S;
  }
  q.dbg = nullptr; q.dbg_step = 0;
  if (getenv("TACO2DEC_PBW_DEBUG")) {      // diagnostics: clock stamps of CTA 0 around the attention-LSTM product of one step
    if (!h->pb_dbg) CUDA_TRY(cudaMalloc(&h->pb_dbg, 256 * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(h->pb_dbg, 0, 256 * sizeof(long long), st));
    q.dbg = h->pb_dbg; q.dbg_step = env_int("TACO2DEC_PBW_DEBUG", 20);
  }
  q.w2_stream = env_int("TACO2DEC_PBW_W2_STREAM", 1);     // measured: 40.5 -> 39.1 us/frame (the working set no longer fits in L2 otherwise)
  CUDA_TRY(cudaMemsetAsync(g.dq, 0, (size_t)S * p.T * B * bt::A * sizeof(float), st));     // sub-tasks add their partial dq rows
  q.stages_a0 = geo.stages_a0; q.stages_a1 = geo.stages_a1; q.stages_x0 = geo.stages_x0; q.stages_x1 = geo.stages_x1; q.n_res = geo.n_res; q.n_tm = geo.n_tm;
  CUDA_TRY(cudaMemsetAsync(q.flags, 0, (size_t)pbw::F_COUNT * pbw::kFlagStride * sizeof(unsigned), st));
  CUDA_TRY(cudaMemsetAsync(bb.dg1, 0, (size_t)S * (bw::G / 64) * NPAD * 128, st));      // utterance columns >= B of the operand tiles
  CUDA_TRY(cudaMemsetAsync(bb.dg2, 0, (size_t)(bw::G / 64) * NPAD * 128, st));
  bw::bw_dy_all<<<dim3((p.T * B + bw::kDyRows - 1) / bw::kDyRows, (bt::H + S * bt::E) / bw::kDyCols), 256, 0, st>>>(p, g);
  auto kern = pbw::decoder_backward_persistent<NPAD>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pbw::kThreads, geo.smem));
  if (per_sm < 1) return fail(TACO2DEC_E_STATE, "persistent backward kernel does not fit on an SM");
  void* args[] = {(void*)&p, (void*)&g, (void*)&q};
  if (h->profiling) CUDA_TRY(cudaEventRecord(h->ev0, st));
  CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(pbw::kCtas), dim3(pbw::kThreads), args, geo.smem, st));
  if (h->profiling) { CUDA_TRY(cudaEventRecord(h->ev1, st)); h->ev_valid = true; }
  CUDA_TRY(cudaGetLastError());
  h->launches += 2;
  return 0;
}

int run_common(taco2dec_handle* h, Params& p, int T_in, int T_sub, char* ws, const WorkspaceLayout& L, cudaStream_t st) {
  const taco2dec_config& c = h->cfg;
  // control words: barrier counter, watchdog flag, done counter
  CUDA_TRY(cudaMemsetAsync(ws + L.ctl, 0, 64 * sizeof(float), st));
  h->last_phase_clocks = p.phase_clocks;
  // processed memory, once per call (model.py:258-261)
  for (int s = 0; s < c.n_streams; ++s) {
    const int n_rows = p.B * p.st[s].Ts;
    if (h->pm_given[s]) {      // already computed upstream (taco2dec_memprep_forward): one copy into the call's layout
      CUDA_TRY(cudaMemcpyAsync(p.st[s].pm, h->pm_given[s], (size_t)n_rows * c.attn_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
      continue;
    }
    const int groups = (c.attn_dim + 3) / 4, want_warps = h->num_sms * 16;
    const int n_chunk = std::max(1, std::min(groups, want_warps / std::max(n_rows, 1)));
    const int blocks = std::min((n_rows * n_chunk + 7) / 8, h->num_sms * 8);
    processed_memory_kernel<<<blocks, 256, 0, st>>>(p.st[s].mem, p.st[s].wm, p.st[s].pm, n_rows, c.enc_dim, c.attn_dim, n_chunk);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  const bool tensor_ok = bt_shape_ok(h, p.B, T_in, T_sub);
  const bool tensor_mode = h->path_mode == TACO2DEC_PATH_TENSOR || h->path_mode == TACO2DEC_PATH_TENSOR_GRAPH;
  if (h->cur_sv.gates1) {   // activations are kept for backward: only the tensor path produces them
    if (!bw_shape_ok(h, p.B, T_in, T_sub) || (h->path_mode != TACO2DEC_PATH_AUTO && !tensor_mode) ||
        (h->path_mode == TACO2DEC_PATH_AUTO && h->batched_fp32))
      return fail(TACO2DEC_E_ARG, "saving activations for backward needs the tensor path: 2 <= B <= 128, default decoder "
                                  "dims, batched precision fp16");
    return run_batched(h, p, T_in, T_sub, st);
  }
  const bool want_lat = (h->path_mode == TACO2DEC_PATH_AUTO || h->path_mode == TACO2DEC_PATH_LATENCY) &&
                        lat_shape_ok(h, p.B, T_in, T_sub);
  if (h->path_mode == TACO2DEC_PATH_LATENCY && !want_lat)
    return fail(TACO2DEC_E_ARG, "latency path needs B=1, default decoder dims and a short enough memory");
  if (want_lat) return run_latency(h, p, st);
  if (tensor_mode && !tensor_ok)
    return fail(TACO2DEC_E_ARG, "tensor path needs 2 <= B <= 128 and default decoder dims");
  // AUTO: batched calls take the tensor-core path (weights AND x/h operands rounded to fp16, fp32 accumulation)
  // unless the caller asked for fp32-exact batched arithmetic (taco2dec_set_batched_precision)
  if (tensor_ok && (tensor_mode || (h->path_mode == TACO2DEC_PATH_AUTO && !h->batched_fp32)))
    return run_batched(h, p, T_in, T_sub, st);
  h->last_path = TACO2DEC_PATH_GENERIC;
  const int BT = pick_bt(p.B);
  const size_t smem = persistent_smem_bytes(c, BT, T_in, T_sub);
  if ((int)smem > h->max_smem_optin)
    return fail(TACO2DEC_E_ARG, "sequence too long for the attention shared-memory plan");
  switch (BT) {
    case 1: return launch_persistent<1>(h, p, smem, st);
    case 2: return launch_persistent<2>(h, p, smem, st);
    case 4: return launch_persistent<4>(h, p, smem, st);
    default: return launch_persistent<8>(h, p, smem, st);
  }
}

}  // namespace

namespace {
__global__ void sum_splits_kernel(const float* part, int splits, int M, int NPAD, int N, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[((size_t)k * M + m) * NPAD + n];
  out[i] = s;
}
template <int NPAD>
int run_test_gemm(int M, int N, int K, int splits, const float* A, const float* X, float* out, cudaStream_t st) {
  unsigned char *a_t = nullptr, *x_t = nullptr;
  float* part = nullptr;
  CUDA_TRY(cudaMalloc(&a_t, (size_t)M * K * 2));
  CUDA_TRY(cudaMalloc(&x_t, (size_t)NPAD * K * 2));
  CUDA_TRY(cudaMalloc(&part, (size_t)splits * M * NPAD * sizeof(float)));
  // operand formats: default fp16 x fp16; TACO2DEC_GEMM_FMT = "bf16" (both) or "mixed" (fp16 weights x bf16 activations)
  const char* fmt_env = getenv("TACO2DEC_GEMM_FMT");
  const std::string fmt = fmt_env ? fmt_env : "";
  const int a_bf16 = fmt == "bf16", x_bf16 = (fmt == "bf16" || fmt == "mixed");
  tc::pack_tiles_kernel<<<1024, 256, 0, st>>>(A, M, K, tc::kBlockM, M / tc::kBlockM, a_t, a_bf16);
  tc::pack_tiles_kernel<<<256, 256, 0, st>>>(X, N, K, NPAD, 1, x_t, x_bf16);
  tc::GemmParams gp{a_t, x_t, part, M, K, splits, 1, 0, 0, 0, nullptr, 0};
  gp.fmt = ((unsigned)a_bf16 << 7) | ((unsigned)x_bf16 << 10);
  unsigned long long* dbg = nullptr;
  const int n_cta = (M / tc::kBlockM) * splits;
  if (getenv("TACO2DEC_GEMM_STAMPS")) {
    CUDA_TRY(cudaMalloc(&dbg, (size_t)n_cta * 8 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(dbg, 0, (size_t)n_cta * 8 * sizeof(unsigned long long)));
    gp.dbg = dbg;
  }
  CUDA_TRY(tc::prepare_gemm<NPAD>());
  if (dbg) {   // warm launches, then a timed one whose stamps are printed (ns relative to the earliest CTA entry)
    for (int i = 0; i < 3; ++i) CUDA_TRY(tc::launch_gemm<NPAD>(gp, st));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    CUDA_TRY(tc::launch_gemm<NPAD>(gp, st));
    cudaEventRecord(e1, st);
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> hst((size_t)n_cta * 8);
    CUDA_TRY(cudaMemcpy(hst.data(), dbg, hst.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull, tend = 0;
    for (int c = 0; c < n_cta; ++c) { t0 = std::min(t0, hst[c * 8]); tend = std::max(tend, hst[c * 8 + 5]); }
    fprintf(stderr, "[gemm stamps] M=%d N=%d K=%d splits=%d ctas=%d event %.2f us, first entry -> last exit %.2f us\n", M, N, K,
            splits, n_cta, ms * 1e3, (tend - t0) * 1e-3);
    for (int c = 0; c < n_cta; c += std::max(1, n_cta / 4))
      fprintf(stderr, "  cta %3d: entry +%.2f  setup %.2f  first tile %.2f  acc done %.2f  stored %.2f  exit %.2f us\n", c,
              (hst[c * 8] - t0) * 1e-3, (hst[c * 8 + 1] - t0) * 1e-3, (hst[c * 8 + 2] - t0) * 1e-3, (hst[c * 8 + 3] - t0) * 1e-3,
              (hst[c * 8 + 4] - t0) * 1e-3, (hst[c * 8 + 5] - t0) * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dbg);
  }
  CUDA_TRY(tc::launch_gemm<NPAD>(gp, st));
  sum_splits_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(part, splits, M, NPAD, N, out);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  cudaFree(a_t); cudaFree(x_t); cudaFree(part);
  return 0;
}
}  // namespace


extern "C" {

int taco2dec_abi_version(void) { return TACO2DEC_ABI_VERSION; }
const char* taco2dec_last_error(void) { return g_err.c_str(); }

int taco2dec_create(const taco2dec_config* cfg, int device, taco2dec_handle** out) {
  if (!cfg || !out) return fail(TACO2DEC_E_ARG, "null argument");
  if (int rc = check_cfg(*cfg)) return rc;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(TACO2DEC_E_ARCH, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                     "; this library is built for sm_100a (B200) only and has no fallback");
  int coop = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) return fail(TACO2DEC_E_ARCH, "device does not support cooperative launch");
  taco2dec_handle* h = new (std::nothrow) taco2dec_handle();
  if (!h) return fail(TACO2DEC_E_STATE, "out of host memory");
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->have_weights = false;
  h->launches = 0;
  h->abort_dev = nullptr; h->aux_stream = nullptr; h->batched_fp32 = 0;
  h->last_phase_clocks = nullptr;
  h->path_mode = TACO2DEC_PATH_AUTO;
  h->weight_dtype = TACO2DEC_W_FP32;
  h->packed = nullptr; h->packed_bytes = 0; h->packed_wbytes = 0; h->packed_off = nullptr;
  h->ll_buf = nullptr; h->ll_bytes = 0; h->last_path = 0;
  memset(&h->bt_bufs, 0, sizeof(h->bt_bufs)); h->bt_alloc = false; h->bt_tiles_valid = false; h->cap_stream = nullptr;
  memset(&h->cur_sv, 0, sizeof(h->cur_sv));
  memset(&h->pb, 0, sizeof(h->pb)); h->pb_alloc = false; h->pb_tiles_valid = false; h->pb_xpre = nullptr; h->pb_xpre_bytes = 0; h->pb_dbg = nullptr; h->pm_given[0] = h->pm_given[1] = nullptr;
  memset(&h->bw_bufs, 0, sizeof(h->bw_bufs)); h->bw_alloc = false; h->bw_tiles_valid = false; h->bw_ctl = nullptr;
  memset(&h->pbw, 0, sizeof(h->pbw)); h->pbw_alloc = false; h->pbw_h16 = nullptr; h->pbw_h16_elems = 0;
  h->profiling = false;
  h->ev_valid = false;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaEventCreate(&h->ev0));
  CUDA_TRY(cudaEventCreate(&h->ev1));
  CUDA_TRY(cudaMalloc(&h->abort_dev, 64));
  CUDA_TRY(cudaMemset(h->abort_dev, 0, 64));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
  *out = h;
  return 0;
}

int taco2dec_destroy(taco2dec_handle* h) {
  if (h) {
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    if (h->abort_dev) cudaFree(h->abort_dev);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->packed) cudaFree(h->packed);
    if (h->packed_off) cudaFree(h->packed_off);
    if (h->ll_buf) cudaFree(h->ll_buf);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->bt_alloc) {
      bt::Bufs& b = h->bt_bufs;
      void* ptrs[] = {b.a1, b.a2, b.aq, b.x1, b.x2, b.g1, b.g2, b.gq, b.c1, b.c2, b.h2f, b.w0t[0], b.w0t[1], b.w1t[0], b.w1t[1]};
      for (void* q : ptrs) if (q) cudaFree(q);
    }
    if (h->pb_alloc) {
      pb::PbParams& b = h->pb;
      void* ptrs[] = {(void*)b.wt, b.x1, b.x2, b.part1, b.part2, b.qpart, b.melp, b.ctxp, b.melx, b.l0x, b.flags, h->pb_xpre};
      for (void* q : ptrs) if (q) cudaFree(q);
    }
    if (h->bw_alloc) {
      bw::Bufs& b = h->bw_bufs;
      void* ptrs[] = {b.a1t, b.a2t, b.dg1, b.dg2, b.dx1, b.dx2, h->bw_ctl};
      for (void* q : ptrs) if (q) cudaFree(q);
    }
    if (h->pbw_alloc) {
      void* ptrs[] = {h->pbw.dx1, h->pbw.dx2, h->pbw.flags, h->pbw.dxc1, h->pbw.dxc2, h->pbw_h16};
      for (void* q : ptrs) if (q) cudaFree(q);
    }
  }
  delete h;
  return 0;
}

int taco2dec_read_debug_stamps(taco2dec_handle* h, void* cuda_stream, long long* out256_host) {
  if (!h || !out256_host) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->pb_dbg) return fail(TACO2DEC_E_STATE, "no debug stamps recorded (set TACO2DEC_PB_DEBUG=<frame>)");
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  CUDA_TRY(cudaMemcpy(out256_host, h->pb_dbg, 256 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

int taco2dec_read_phase_clocks(taco2dec_handle* h, void* cuda_stream, long long* out16_host) {
  if (!h || !out16_host) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->last_phase_clocks) return fail(TACO2DEC_E_STATE, "no launch recorded");
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  CUDA_TRY(cudaMemcpy(out16_host, h->last_phase_clocks, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

/* Test hook for the tcgen05 GEMM building block: out[M][N] = A[M][K] . X[N][K]^T with fp16-rounded operands,
 * fp32 accumulation.  M % 128 == 0, K % (64*splits) == 0, 1 <= N <= 128.  Synchronises the stream. */
int taco2dec_test_gemm(int M, int N, int K, int splits, const float* A, const float* X, float* out, void* cuda_stream) {
  if (!A || !X || !out || M < 128 || M % 128 || N < 1 || N > 128 || splits < 1 || K % (64 * splits))
    return fail(TACO2DEC_E_ARG, "bad GEMM shape");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (N <= 16) return run_test_gemm<16>(M, N, K, splits, A, X, out, st);
  if (N <= 32) return run_test_gemm<32>(M, N, K, splits, A, X, out, st);
  if (N <= 64) return run_test_gemm<64>(M, N, K, splits, A, X, out, st);
  return run_test_gemm<128>(M, N, K, splits, A, X, out, st);
}

int taco2dec_set_mode(taco2dec_handle* h, int path, int weight_dtype) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  if (path < TACO2DEC_PATH_AUTO || path > TACO2DEC_PATH_TENSOR_GRAPH) return fail(TACO2DEC_E_ARG, "bad path");
  if (weight_dtype != TACO2DEC_W_FP32 && weight_dtype != TACO2DEC_W_FP16) return fail(TACO2DEC_E_ARG, "bad weight dtype");
  h->path_mode = path;
  h->weight_dtype = weight_dtype;
  return 0;
}

int taco2dec_last_path(const taco2dec_handle* h) { return h ? h->last_path : 0; }

int taco2dec_set_profiling(taco2dec_handle* h, int on) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  h->profiling = on != 0;
  return 0;
}

int taco2dec_last_kernel_ms(taco2dec_handle* h, float* ms) {
  if (!h || !ms) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->ev_valid) return fail(TACO2DEC_E_STATE, "no profiled launch recorded");
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return 0;
}

int taco2dec_set_weights(taco2dec_handle* h, const taco2dec_weights* w, void* /*cuda_stream*/) {
  if (!h || !w) return fail(TACO2DEC_E_ARG, "null argument");
  auto bad = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  for (int s = 0; s < h->cfg.n_streams; ++s) {
    const taco2dec_stream_weights& sw = w->stream[s];
    const void* req[] = {sw.prenet_w0, sw.prenet_w1, sw.arnn_w_ih, sw.arnn_w_hh, sw.query_w, sw.memory_w};
    for (const void* p : req)
      if (bad(p)) return fail(TACO2DEC_E_ARG, "stream weight pointer is null or not 16-byte aligned");
    if (!sw.arnn_b_ih || !sw.arnn_b_hh || !sw.v) return fail(TACO2DEC_E_ARG, "stream bias / v pointer is null");
    if (h->cfg.attention == TACO2DEC_ATTN_LSA && (!sw.loc_conv_w || !sw.loc_dense_w))
      return fail(TACO2DEC_E_ARG, "LSA needs location conv / dense weights");
  }
  const void* req[] = {w->drnn_w_ih, w->drnn_w_hh, w->proj_w, w->gate_w};
  for (const void* p : req)
    if (bad(p)) return fail(TACO2DEC_E_ARG, "decoder weight pointer is null or not 16-byte aligned");
  if (!w->drnn_b_ih || !w->drnn_b_hh || !w->proj_b || !w->gate_b) return fail(TACO2DEC_E_ARG, "bias pointer is null");
  h->w = *w;
  h->have_weights = true;
  h->packed_wbytes = 0;  // latency-path weight streams are re-packed on next use
  h->bt_tiles_valid = false;
  h->bw_tiles_valid = false;
  h->pb_tiles_valid = false;
  return 0;
}

size_t taco2dec_workspace_bytes(const taco2dec_handle* h, int B, int T_in, int T_sub, int T, int teacher_forced) {
  if (!h || B < 1 || T_in < 1 || T < 1) return 0;
  return plan_workspace(h->cfg, B, T_in, std::max(T_sub, 1), T, teacher_forced != 0).total;
}

int taco2dec_forward_teacher_forced(taco2dec_handle* h, const taco2dec_tf_args* a, void* cuda_stream) {
  if (!h || !a) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->have_weights) return fail(TACO2DEC_E_STATE, "weights not set");
  const taco2dec_config& c = h->cfg;
  if (a->B < 1 || a->T < 1 || a->T_in < 1 || (c.n_streams == 2 && a->T_sub < 1))
    return fail(TACO2DEC_E_ARG, "B, T, T_in, T_sub must be >= 1");
  if (!a->memory || !a->decoder_inputs || !a->mel || !a->gate || !a->align || !a->workspace)
    return fail(TACO2DEC_E_ARG, "null tensor pointer");
  if (c.n_streams == 2 && (!a->embeddings || !a->align_bert)) return fail(TACO2DEC_E_ARG, "sub-word stream tensors missing");
  if ((reinterpret_cast<uintptr_t>(a->memory) & 15u) || (c.n_streams == 2 && (reinterpret_cast<uintptr_t>(a->embeddings) & 15u)))
    return fail(TACO2DEC_E_ARG, "memory / embeddings must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  const WorkspaceLayout L = plan_workspace(c, a->B, a->T_in, std::max(a->T_sub, 1), a->T, true);
  if (a->workspace_bytes < L.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  Params p;
  char* ws = (char*)a->workspace;
  fill_common(h, p, a->B, a->T_in, a->T_sub, a->memory, a->embeddings, a->memory_lengths, a->bert_lengths, a->rng, ws, L);
  p.T = a->T; p.Tcap = a->T; p.free_running = 0; p.training = a->training ? 1 : 0; p.max_steps = a->T;
  p.independent = a->independent ? 1 : 0;
  p.dec_in = a->decoder_inputs;
  p.mel = a->mel; p.gate = a->gate;
  p.st[0].align = a->align;
  if (c.n_streams == 2) p.st[1].align = a->align_bert;
  memset(&h->cur_sv, 0, sizeof(h->cur_sv));
  float* pre0_save[2] = {nullptr, nullptr};
  if (a->saved) {
    const taco2dec_saved_layout SL = plan_saved(c, a->B, a->T_in, std::max(a->T_sub, 1), a->T);
    if (a->saved_bytes < SL.total) return fail(TACO2DEC_E_STATE, "saved-activation buffer too small");
    if (reinterpret_cast<uintptr_t>(a->saved) & 255u) return fail(TACO2DEC_E_ARG, "saved buffer must be 256-byte aligned");
    char* sb = (char*)a->saved;
    h->cur_sv.gates1 = (float*)(sb + SL.gates1); h->cur_sv.gates2 = (float*)(sb + SL.gates2);
    h->cur_sv.h1 = (float*)(sb + SL.h1); h->cur_sv.ctx = (float*)(sb + SL.ctx); h->cur_sv.h2 = (float*)(sb + SL.h2);
    h->cur_sv.q = (float*)(sb + SL.q);
    for (int s = 0; s < c.n_streams; ++s) {
      p.st[s].p_save = (float*)(sb + SL.p[s]);
      p.st[s].pm = (float*)(sb + SL.pm[s]);
      p.st[s].pre = (float*)(sb + SL.pre[s]);
      pre0_save[s] = (float*)(sb + SL.pre0[s]);
    }
  } else if (a->B >= 2 && a->B <= 128) {
    h->cur_sv.h2 = (float*)(ws + L.h2_all);
    h->cur_sv.ctx = (float*)(ws + L.ctx_all);
  }
  // hoisted prenet over the go-frame + all T targets (model.py:407-413)
  for (int s = 0; s < c.n_streams; ++s) {
    const int n_rows = (a->T + 1) * a->B;
    const int wpb = 8;
    const int blocks = std::min((n_rows + wpb - 1) / wpb, h->num_sms * 8);
    const size_t sm = (size_t)wpb * (((c.n_mel + 3) & ~3) + c.prenet_dim) * sizeof(float);
    if (c.prenet_dim == kPtP && c.n_mel <= kPtP && n_rows >= 4 * kPtRows) {
      CUDA_TRY(cudaFuncSetAttribute(prenet_tf_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPtSmem));
      prenet_tf_tiled_kernel<<<(n_rows + kPtRows - 1) / kPtRows, 256, kPtSmem, st>>>(
          a->decoder_inputs, h->w.stream[s].prenet_w0, h->w.stream[s].prenet_w1, a->rng.prenet_keep[s][0], a->rng.prenet_keep[s][1],
          p.st[s].pre, pre0_save[s], a->B, a->T, c.n_mel, a->rng.seed, s, p.thresh_pre);
    } else {
      prenet_tf_kernel<<<blocks, wpb * 32, sm, st>>>(a->decoder_inputs, h->w.stream[s].prenet_w0, h->w.stream[s].prenet_w1,
                                                     a->rng.prenet_keep[s][0], a->rng.prenet_keep[s][1], p.st[s].pre,
                                                     pre0_save[s], a->B, a->T, c.n_mel, c.prenet_dim, a->rng.seed, s,
                                                     p.thresh_pre);
    }
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  h->pm_given[0] = a->processed_memory[0]; h->pm_given[1] = c.n_streams == 2 ? a->processed_memory[1] : nullptr;
  const int rc = run_common(h, p, a->T_in, std::max(a->T_sub, 1), ws, L, st);
  h->pm_given[0] = h->pm_given[1] = nullptr;
  memset(&h->cur_sv, 0, sizeof(h->cur_sv));
  return rc;
}

int taco2dec_saved_layout_query(const taco2dec_handle* h, int B, int T_in, int T_sub, int T, taco2dec_saved_layout* out) {
  if (!h || !out || B < 1 || T_in < 1 || T < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  *out = plan_saved(h->cfg, B, T_in, std::max(T_sub, 1), T);
  return 0;
}

int taco2dec_grad_layout_query(const taco2dec_handle* h, int B, int T_in, int T_sub, int T, taco2dec_grad_layout* out) {
  if (!h || !out || B < 1 || T_in < 1 || T < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  *out = plan_grads(h->cfg, B, T_in, std::max(T_sub, 1), T, nullptr);
  return 0;
}

int taco2dec_backward(taco2dec_handle* h, const taco2dec_bwd_args* a, void* cuda_stream) {
  if (!h || !a) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->have_weights) return fail(TACO2DEC_E_STATE, "weights not set");
  const taco2dec_config& c = h->cfg;
  if (a->B < 1 || a->T < 1 || a->T_in < 1 || (c.n_streams == 2 && a->T_sub < 1))
    return fail(TACO2DEC_E_ARG, "B, T, T_in, T_sub must be >= 1");
  if (!a->memory || !a->align || !a->d_mel || !a->d_gate || !a->saved || !a->grads)
    return fail(TACO2DEC_E_ARG, "null tensor pointer");
  if (c.n_streams == 2 && (!a->embeddings || !a->align_bert)) return fail(TACO2DEC_E_ARG, "sub-word stream tensors missing");
  const int T_sub = std::max(a->T_sub, 1);
  if (!bw_shape_ok(h, a->B, a->T_in, T_sub))
    return fail(TACO2DEC_E_ARG, "backward needs the tensor path: 2 <= B <= 128, default decoder dims");
  if ((reinterpret_cast<uintptr_t>(a->saved) & 255u) || (reinterpret_cast<uintptr_t>(a->grads) & 255u))
    return fail(TACO2DEC_E_ARG, "saved / grads buffers must be 256-byte aligned");
  const taco2dec_saved_layout SL = plan_saved(c, a->B, a->T_in, T_sub, a->T);
  GradScratch gs;
  const taco2dec_grad_layout GL = plan_grads(c, a->B, a->T_in, T_sub, a->T, &gs);
  if (a->saved_bytes < SL.total) return fail(TACO2DEC_E_STATE, "saved-activation buffer too small");
  if (a->grads_bytes < GL.total) return fail(TACO2DEC_E_STATE, "gradient buffer too small");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  Params p;
  WorkspaceLayout L0;
  memset(&L0, 0, sizeof(L0));
  fill_common(h, p, a->B, a->T_in, a->T_sub, a->memory, a->embeddings, a->memory_lengths, a->bert_lengths, a->rng, nullptr, L0);
  p.h1 = p.c1 = p.h2 = p.c2 = p.ctx = p.q = nullptr;
  p.sync_ctr = nullptr; p.abort_flag = nullptr; p.done_count = nullptr; p.phase_clocks = nullptr;
  p.T = a->T; p.Tcap = a->T; p.free_running = 0; p.training = a->training ? 1 : 0; p.max_steps = a->T;
  p.independent = a->independent ? 1 : 0;
  char* sb = (char*)const_cast<void*>(a->saved);
  char* gb = (char*)a->grads;
  bw::Grads g;
  memset(&g, 0, sizeof(g));
  g.sv.gates1 = (float*)(sb + SL.gates1); g.sv.gates2 = (float*)(sb + SL.gates2);
  g.sv.h1 = (float*)(sb + SL.h1); g.sv.ctx = (float*)(sb + SL.ctx); g.sv.h2 = (float*)(sb + SL.h2);
  g.sv.q = (float*)(sb + SL.q);
  const float* aligns[2] = {a->align, a->align_bert};
  const float* d_aligns[2] = {a->d_align, a->d_align_bert};
  for (int s = 0; s < c.n_streams; ++s) {
    p.st[s].pm = (float*)(sb + SL.pm[s]);
    p.st[s].pre = nullptr; p.st[s].pre0 = nullptr; p.st[s].a_prev = nullptr; p.st[s].a_cum = nullptr;
    g.p_saved[s] = (const float*)(sb + SL.p[s]);
    g.align[s] = aligns[s];
    g.d_align[s] = d_aligns[s];
    g.dpm[s] = (float*)(gb + GL.dpm[s]);
    g.dalpha[s] = (float*)(gb + gs.dalpha[s]);
    g.dcum[s] = (float*)(gb + gs.dcum[s]);
    g.dz[s] = (float*)(gb + gs.dz[s]);
  }
  g.dwd = (float*)(gb + GL.dloc_dense); g.dwc = (float*)(gb + GL.dloc_conv);
  g.d_mel = a->d_mel; g.d_gate = a->d_gate;
  g.dg1 = (float*)(gb + GL.dg1); g.dg2 = (float*)(gb + GL.dg2); g.dq = (float*)(gb + GL.dq);
  g.dctx = (float*)(gb + GL.dctx); g.dpre = (float*)(gb + GL.dpre); g.dv = (float*)(gb + GL.dv);
  g.dc1 = (float*)(gb + gs.dc1); g.dc2 = (float*)(gb + gs.dc2);
  g.dyh = (float*)(gb + gs.dyh); g.dyc = (float*)(gb + gs.dyc);
  CUDA_TRY(cudaMemsetAsync(gb + gs.zero_begin, 0, gs.zero_end - gs.zero_begin, st));
  if (int rc = bw_prepare(h, st)) return rc;
  if (pbw_shape_ok(h, a->B, a->T_in, T_sub)) {       // one persistent launch for the whole reverse-time loop
    PbwGeometry geo;
    pbw_geometry(h, a->B, a->T_in, T_sub, &geo);
    p.abort_flag = h->abort_dev;
    if (!h->pbw_alloc) {       // first call: pbw_run allocates; the phase clocks live behind the counters
      const size_t NP = 128;
      CUDA_TRY(cudaMalloc(&h->pbw.dx1, (size_t)2 * 2 * pbw::kSplits * bt::K1 * NP * sizeof(float)));
      CUDA_TRY(cudaMalloc(&h->pbw.dx2, (size_t)2 * pbw::kSplits * (2 * (bt::H + bt::E) + bt::H) * NP * sizeof(float)));
      CUDA_TRY(cudaMalloc(&h->pbw.flags, (size_t)pbw::F_COUNT * pbw::kFlagStride * sizeof(unsigned) + 16 * sizeof(long long)));
      CUDA_TRY(cudaMalloc(&h->pbw.dxc1, (size_t)2 * 2 * pbw::kSplits * NP * bt::E * sizeof(float)));
      CUDA_TRY(cudaMalloc(&h->pbw.dxc2, (size_t)2 * pbw::kSplits * 2 * NP * bt::E * sizeof(float)));
      h->pbw_alloc = true;
    }
    p.phase_clocks = (long long*)(h->pbw.flags + (size_t)pbw::F_COUNT * pbw::kFlagStride);
    h->last_phase_clocks = p.phase_clocks;
    switch (geo.npad) {
      case 16: return pbw_run<16>(h, p, g, geo, st);
      case 32: return pbw_run<32>(h, p, g, geo, st);
      default: return pbw_run<64>(h, p, g, geo, st);
    }
  }
  if (a->B <= 16) return bw_run_frames<16>(h, p, g, st);
  if (a->B <= 32) return bw_run_frames<32>(h, p, g, st);
  if (a->B <= 64) return bw_run_frames<64>(h, p, g, st);
  return bw_run_frames<128>(h, p, g, st);
}

int taco2dec_infer(taco2dec_handle* h, const taco2dec_infer_args* a, void* cuda_stream) {
  if (!h || !a) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->have_weights) return fail(TACO2DEC_E_STATE, "weights not set");
  const taco2dec_config& c = h->cfg;
  if (a->B < 1 || a->max_decoder_steps < 1 || a->T_in < 1 || (c.n_streams == 2 && a->T_sub < 1))
    return fail(TACO2DEC_E_ARG, "B, max_decoder_steps, T_in, T_sub must be >= 1");
  if (!a->memory || !a->mel || !a->gate || !a->align || !a->n_frames || !a->reached_max || !a->workspace)
    return fail(TACO2DEC_E_ARG, "null tensor pointer");
  if (c.n_streams == 2 && (!a->embeddings || !a->align_bert)) return fail(TACO2DEC_E_ARG, "sub-word stream tensors missing");
  if ((reinterpret_cast<uintptr_t>(a->memory) & 15u) || (c.n_streams == 2 && (reinterpret_cast<uintptr_t>(a->embeddings) & 15u)))
    return fail(TACO2DEC_E_ARG, "memory / embeddings must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  const WorkspaceLayout L = plan_workspace(c, a->B, a->T_in, std::max(a->T_sub, 1), a->max_decoder_steps, false);
  if (a->workspace_bytes < L.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  Params p;
  char* ws = (char*)a->workspace;
  fill_common(h, p, a->B, a->T_in, a->T_sub, a->memory, a->embeddings, a->memory_lengths, a->bert_lengths, a->rng, ws, L);
  p.T = a->max_decoder_steps; p.Tcap = a->max_decoder_steps; p.free_running = 1; p.training = 0;
  p.max_steps = a->max_decoder_steps; p.gate_thr = a->gate_threshold;
  p.mel = a->mel; p.gate = a->gate;
  p.st[0].align = a->align;
  if (c.n_streams == 2) p.st[1].align = a->align_bert;
  p.n_frames = a->n_frames; p.reached_max = a->reached_max;
  h->pm_given[0] = a->processed_memory[0]; h->pm_given[1] = c.n_streams == 2 ? a->processed_memory[1] : nullptr;
  const int rc = run_common(h, p, a->T_in, std::max(a->T_sub, 1), ws, L, st);
  h->pm_given[0] = h->pm_given[1] = nullptr;
  return rc;
}

int taco2dec_check(taco2dec_handle* h, void* cuda_stream) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  CUDA_TRY(cudaGetLastError());
  return taco2dec_poll_abort(h);
}

int taco2dec_poll_abort(taco2dec_handle* h) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, h->abort_dev, sizeof(int), cudaMemcpyDeviceToHost, h->aux_stream));
  CUDA_TRY(cudaStreamSynchronize(h->aux_stream));
  if (flag) {
    CUDA_TRY(cudaMemsetAsync(h->abort_dev, 0, sizeof(int), h->aux_stream));
    CUDA_TRY(cudaStreamSynchronize(h->aux_stream));
    return fail(TACO2DEC_E_ABORTED, "persistent decoder kernel aborted: an in-kernel watchdog fired (outputs of the calls "
                                    "since the last check are invalid)");
  }
  return 0;
}

int taco2dec_set_batched_precision(taco2dec_handle* h, int fp32_exact) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  h->batched_fp32 = fp32_exact ? 1 : 0;
  return 0;
}

int64_t taco2dec_launch_count(const taco2dec_handle* h) { return h ? h->launches : 0; }

int taco2dec_philox_keep_mask(uint64_t seed, int mask_id, int rows, int n, float p_drop, uint8_t* out, void* cuda_stream) {
  if (!out || rows < 1 || n < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  philox_mask_kernel<<<256, 256, 0, (cudaStream_t)cuda_stream>>>(seed, mask_id, rows, n, keep_threshold(p_drop), out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int taco2dec_launch_geometry(const taco2dec_handle* h, int B, int* grid, int* block, int* smem_bytes) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  if (grid) *grid = h->num_sms;
  if (block) *block = kThreads;
  if (smem_bytes) *smem_bytes = (int)persistent_smem_bytes(h->cfg, pick_bt(B), 256, 128);
  return 0;
}

}  // extern "C"


// ------------------------------------------------------------------------------------------
// Machine probes for bench.py's roofline: what the decoder kernels are actually bound by once the weights are on-chip /
// L2-resident is (a) the L2 -> SM read rate and (b) the latency of one cross-CTA exchange through L2
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(512) l2_read_kernel(const float4* __restrict__ buf, size_t n_vec, int iters, float* sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += stride * 4) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (i + k * stride < n_vec) ? __ldcg(buf + i + k * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
  if (acc == 1.2345e-30f) *sink = acc;
}
// two CTAs bounce one 64-bit {value, tag} word: one round trip = two exchanges
__global__ void hop_pingpong_kernel(unsigned long long* words, int rounds, int* abort_flag) {
  if (threadIdx.x != 0) return;
  lat::Watch wd{abort_flag, 0, 0};
  float v;
  for (int k = 1; k <= rounds; ++k) {
    if (blockIdx.x == 0) {
      lat::ll_store(words, (float)k, (unsigned)k);
      if (!lat::ll_wait(words + 16, (unsigned)k, v, wd)) return;
    } else {
      if (!lat::ll_wait(words, (unsigned)k, v, wd)) return;
      lat::ll_store(words + 16, v, (unsigned)k);
    }
  }
}
}  // namespace

extern "C" int taco2dec_measure_machine(taco2dec_handle* h, void* cuda_stream, double* l2_read_gbs, double* hop_ns) {
  if (!h || !l2_read_gbs || !hop_ns) return fail(TACO2DEC_E_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t bytes = (size_t)32 << 20;      // 32 MiB: resident in either half of the L2
  float4* buf = nullptr;
  float* sink = nullptr;
  unsigned long long* words = nullptr;
  CUDA_TRY(cudaMalloc(&buf, bytes));
  CUDA_TRY(cudaMalloc(&sink, 256));
  CUDA_TRY(cudaMalloc(&words, 256));
  CUDA_TRY(cudaMemsetAsync(buf, 0, bytes, st));
  CUDA_TRY(cudaMemsetAsync(words, 0, 256, st));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int iters = 40;
  l2_read_kernel<<<h->num_sms * 4, 512, 0, st>>>(buf, bytes / 16, 2, sink);          // warm: pull the buffer into L2
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, st));
    l2_read_kernel<<<h->num_sms * 4, 512, 0, st>>>(buf, bytes / 16, iters, sink);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  *l2_read_gbs = (double)bytes * iters / (best * 1e-3) / 1e9;
  const int rounds = 2000;
  hop_pingpong_kernel<<<2, 32, 0, st>>>(words, 50, h->abort_dev);
  CUDA_TRY(cudaMemsetAsync(words, 0, 256, st));
  CUDA_TRY(cudaEventRecord(e0, st));
  hop_pingpong_kernel<<<2, 32, 0, st>>>(words, rounds, h->abort_dev);
  CUDA_TRY(cudaEventRecord(e1, st));
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *hop_ns = (double)ms * 1e6 / (2.0 * rounds);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink); cudaFree(words);
  h->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------
// Decoder inputs (memprep.cuh)
// ------------------------------------------------------------------------------------------
struct taco2dec_memprep {
  int device, num_sms, enc, cls, attn;
  bool have_conv, have_mem;
  unsigned char *a1, *a2;     // split-fp16 weight tiles: converter [enc/128][3 (enc+cls)/64], memory layer [attn/128][3 enc/64]
  float* bias;                // [enc]
  int64_t launches;
};

namespace {
struct MpPlan { int groups, n_pad, sp1, sp2; size_t x1, part1, x2, part2, total; };
MpPlan mp_plan(const taco2dec_memprep* h, int n_rows) {
  MpPlan pl;
  pl.groups = (n_rows + mp::kNP - 1) / mp::kNP;
  pl.n_pad = pl.groups * mp::kNP;
  const size_t k1 = 3 * (size_t)(h->enc + h->cls), k2 = 3 * (size_t)h->enc;
  // split K only as far as it takes to give every SM a CTA (small inputs: one utterance is 2 groups)
  pl.sp1 = mp::pick_splits(h->enc / tc::kBlockM * pl.groups, (int)(k1 / tc::kBlockK), h->num_sms);
  pl.sp2 = mp::pick_splits(h->attn / tc::kBlockM * pl.groups, (int)(k2 / tc::kBlockK), h->num_sms);
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  pl.x1 = take((size_t)pl.n_pad * k1 * 2);
  pl.part1 = take((size_t)pl.groups * pl.sp1 * h->enc * mp::kNP * sizeof(float));
  pl.x2 = take((size_t)pl.n_pad * k2 * 2);
  pl.part2 = take((size_t)pl.groups * pl.sp2 * h->attn * mp::kNP * sizeof(float));
  pl.total = off;
  return pl;
}
int mp_project_impl(taco2dec_memprep* h, const MpPlan& pl, int n_rows, float* pm, char* ws, cudaStream_t st) {
  const int kb2 = 3 * h->enc / tc::kBlockK;
  const int sp2 = pl.sp2;
  float* part2 = (float*)(ws + pl.part2);
  tc::GemmParams g2{h->a2, (unsigned char*)(ws + pl.x2), part2, h->attn, 3 * h->enc, sp2, pl.groups, (long long)kb2 * mp::kNP * 128, 0, 0, nullptr, 0};
  g2.a_shared = 1;
  CUDA_TRY(tc::launch_gemm<mp::kNP>(g2, st));
  mp::mp_finish_kernel<<<dim3(pl.groups, h->attn / 32), 256, 0, st>>>(part2, sp2, h->attn, nullptr, n_rows, pm, nullptr);
  h->launches += 2;
  return 0;
}
}  // namespace

extern "C" {

int taco2dec_memprep_create(int enc_dim, int cls_dim, int attn_dim, int device, taco2dec_memprep** out) {
  if (!out) return fail(TACO2DEC_E_ARG, "null argument");
  if (enc_dim < 128 || enc_dim % 128 || attn_dim < 128 || attn_dim % 128 || cls_dim < 0 || (enc_dim + cls_dim) % 64)
    return fail(TACO2DEC_E_ARG, "memprep needs enc_dim and attn_dim multiples of 128 and enc_dim + cls_dim a multiple of 64");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(TACO2DEC_E_ARCH, "this library is built for sm_100a (B200) only and has no fallback");
  taco2dec_memprep* h = new (std::nothrow) taco2dec_memprep();
  if (!h) return fail(TACO2DEC_E_STATE, "out of host memory");
  h->device = device; h->num_sms = prop.multiProcessorCount; h->enc = enc_dim; h->cls = cls_dim; h->attn = attn_dim;
  h->have_conv = h->have_mem = false; h->launches = 0;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaMalloc(&h->a1, (size_t)enc_dim * 3 * (enc_dim + cls_dim) * 2));
  CUDA_TRY(cudaMalloc(&h->a2, (size_t)attn_dim * 3 * enc_dim * 2));
  CUDA_TRY(cudaMalloc(&h->bias, (size_t)enc_dim * sizeof(float)));
  *out = h;
  return 0;
}

int taco2dec_memprep_destroy(taco2dec_memprep* h) {
  if (h) { cudaFree(h->a1); cudaFree(h->a2); cudaFree(h->bias); }
  delete h;
  return 0;
}

int taco2dec_memprep_set_weights(taco2dec_memprep* h, const float* converter_w, const float* converter_b, const float* memory_w,
                                 void* cuda_stream) {
  if (!h || !memory_w) return fail(TACO2DEC_E_ARG, "null argument");
  if ((converter_w == nullptr) != (converter_b == nullptr)) return fail(TACO2DEC_E_ARG, "converter weight and bias go together");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  if (converter_w) {
    mp::mp_pack_weights<<<1024, 256, 0, st>>>(converter_w, h->enc, h->enc + h->cls, h->enc, h->a1);
    CUDA_TRY(cudaMemcpyAsync(h->bias, converter_b, (size_t)h->enc * sizeof(float), cudaMemcpyDeviceToDevice, st));
    h->launches++;
  }
  mp::mp_pack_weights<<<256, 256, 0, st>>>(memory_w, h->attn, h->enc, h->attn, h->a2);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  h->have_conv = converter_w != nullptr;
  h->have_mem = true;
  return 0;
}

size_t taco2dec_memprep_workspace_bytes(const taco2dec_memprep* h, int n_rows) {
  if (!h || n_rows < 1) return 0;
  return mp_plan(h, n_rows).total;
}

int taco2dec_memprep_forward(taco2dec_memprep* h, const float* encoder_outputs, const float* cls_embeddings, int n_rows,
                             float* memory, float* processed_memory, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!h || !encoder_outputs || !memory || !processed_memory || !workspace) return fail(TACO2DEC_E_ARG, "null argument");
  if (h->cls > 0 && !cls_embeddings) return fail(TACO2DEC_E_ARG, "cls_embeddings missing");
  if (!h->have_conv || !h->have_mem) return fail(TACO2DEC_E_STATE, "weights not set");
  if (n_rows < 1) return fail(TACO2DEC_E_ARG, "n_rows must be >= 1");
  if ((reinterpret_cast<uintptr_t>(encoder_outputs) | reinterpret_cast<uintptr_t>(cls_embeddings) | reinterpret_cast<uintptr_t>(memory) |
       reinterpret_cast<uintptr_t>(processed_memory)) & 15u)
    return fail(TACO2DEC_E_ARG, "tensors must be 16-byte aligned");
  const MpPlan pl = mp_plan(h, n_rows);
  if (workspace_bytes < pl.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(TACO2DEC_E_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  char* ws = (char*)workspace;
  CUDA_TRY(tc::prepare_gemm<mp::kNP>());
  const int K1 = h->enc + h->cls, kb1 = 3 * K1 / tc::kBlockK;
  {
    const size_t total = (size_t)pl.n_pad * (K1 / 8);
    mp::mp_input_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 65535), 256, 0, st>>>(encoder_outputs, h->enc, cls_embeddings, h->cls,
                                                                                               n_rows, pl.n_pad, (unsigned char*)(ws + pl.x1));
  }
  const int sp1 = pl.sp1;
  float* part1 = (float*)(ws + pl.part1);
  tc::GemmParams g1{h->a1, (unsigned char*)(ws + pl.x1), part1, h->enc, 3 * K1, sp1, pl.groups, (long long)kb1 * mp::kNP * 128, 0, 0, nullptr, 0};
  g1.a_shared = 1;
  CUDA_TRY(tc::launch_gemm<mp::kNP>(g1, st));
  mp::mp_finish_kernel<<<dim3(pl.groups, h->enc / 32), 256, 0, st>>>(part1, sp1, h->enc, h->bias, n_rows, memory, (unsigned char*)(ws + pl.x2));
  h->launches += 3;
  if (int rc = mp_project_impl(h, pl, n_rows, processed_memory, ws, st)) return rc;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int taco2dec_memprep_project(taco2dec_memprep* h, const float* memory, int n_rows, float* processed_memory, void* workspace,
                             size_t workspace_bytes, void* cuda_stream) {
  if (!h || !memory || !processed_memory || !workspace) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->have_mem) return fail(TACO2DEC_E_STATE, "weights not set");
  if (n_rows < 1) return fail(TACO2DEC_E_ARG, "n_rows must be >= 1");
  const MpPlan pl = mp_plan(h, n_rows);
  if (workspace_bytes < pl.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(TACO2DEC_E_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  char* ws = (char*)workspace;
  CUDA_TRY(tc::prepare_gemm<mp::kNP>());
  const size_t total = (size_t)pl.n_pad * (h->enc / 8);
  mp::mp_memory_operand_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 65535), 256, 0, st>>>(memory, h->enc, n_rows, pl.n_pad,
                                                                                                      (unsigned char*)(ws + pl.x2));
  h->launches++;
  if (int rc = mp_project_impl(h, pl, n_rows, processed_memory, ws, st)) return rc;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"


// ------------------------------------------------------------------------------------------
// Loss + gradient seeding (loss.cuh)
// ------------------------------------------------------------------------------------------
extern "C" size_t taco2dec_loss_workspace_bytes(int B, int n_mel, int T) {
  if (B < 1 || n_mel < 1 || T < 1) return 0;
  const size_t mel_blocks = (size_t)B * ((T + ls::kTT - 1) / ls::kTT);
  return align_up((2 * mel_blocks + 3 * 1024) * sizeof(float), 256);
}

extern "C" int taco2dec_loss_forward(const taco2dec_loss_args* a, void* cuda_stream) {
  if (!a) return fail(TACO2DEC_E_ARG, "null argument");
  if (a->B < 1 || a->n_mel < 1 || a->T < 1) return fail(TACO2DEC_E_ARG, "B, n_mel, T must be >= 1");
  if (!a->mel || !a->mel_postnet || !a->gate || !a->mel_target || !a->gate_target || !a->d_mel || !a->d_mel_postnet || !a->d_gate ||
      !a->losses || !a->workspace)
    return fail(TACO2DEC_E_ARG, "null tensor pointer");
  for (int s = 0; s < 2; ++s)
    if (a->align[s] && (!a->align_target[s] || !a->d_align[s] || a->T_align[s] < 1))
      return fail(TACO2DEC_E_ARG, "alignment loss needs a target, a gradient buffer and its width");
  if (a->workspace_bytes < taco2dec_loss_workspace_bytes(a->B, a->n_mel, a->T)) return fail(TACO2DEC_E_STATE, "workspace too small");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int mel_blocks = a->B * ((a->T + ls::kTT - 1) / ls::kTT);
  float* part_mel = (float*)a->workspace;
  float* part_gate = part_mel + 2 * (size_t)mel_blocks;
  float* part_al[2] = {part_gate + 1024, part_gate + 2048};
  const double n_mel_el = (double)a->B * a->n_mel * a->T, n_gate_el = (double)a->B * a->T;
  const size_t smem = (size_t)a->n_mel * (ls::kTT + 1) * sizeof(float);
  ls::ls_mel_kernel<<<mel_blocks, 256, smem, st>>>(a->mel, a->mel_stride_b, a->mel_stride_c, a->mel_stride_t, a->mel_postnet, a->mel_target,
                                                   a->B, a->n_mel, a->T, (float)(1.0 / n_mel_el), a->d_mel, a->d_mel_postnet, part_mel);
  const int gate_blocks = (int)std::min<size_t>(1024, ((size_t)a->B * a->T + 255) / 256);
  ls::ls_gate_kernel<<<gate_blocks, 256, 0, st>>>(a->gate, a->gate_target, a->B * a->T, (float)(1.0 / n_gate_el), a->d_gate, part_gate);
  ls::ReduceArgs r;
  memset(&r, 0, sizeof(r));
  int al_blocks[2] = {0, 0};
  double inv_al[2] = {0, 0};
  for (int s = 0; s < 2; ++s) {
    if (!a->align[s]) continue;
    const size_t n = (size_t)a->B * a->T * a->T_align[s];
    al_blocks[s] = (int)std::min<size_t>(1024, (n + 255) / 256);
    inv_al[s] = 1.0 / (double)n;
    ls::ls_mse_kernel<<<al_blocks[s], 256, 0, st>>>(a->align[s], a->align_target[s], n, (float)inv_al[s], a->d_align[s], part_al[s]);
  }
  r.mel_part = part_mel; r.gate_part = part_gate; r.al_part = part_al[0]; r.alb_part = part_al[1];
  r.n_mel_blocks = mel_blocks; r.n_gate_blocks = gate_blocks; r.n_al_blocks = al_blocks[0]; r.n_alb_blocks = al_blocks[1];
  r.inv_mel = 1.0 / n_mel_el; r.inv_gate = 1.0 / n_gate_el; r.inv_al = inv_al[0]; r.inv_alb = inv_al[1];
  r.losses = a->losses;
  ls::ls_reduce_kernel<<<1, 256, 0, st>>>(r);
  CUDA_TRY(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------
// Weight-gradient contractions (wgrad.cuh)
// ------------------------------------------------------------------------------------------
extern "C" size_t taco2dec_wgrad_workspace_bytes(const taco2dec_handle* h, int M, int N, int T, int B) {
  if (!h || M < 1 || N < 1 || T < 1 || B < 1) return 0;
  return wg::plan(M, N, T * B, h->num_sms).total;
}

static int wgrad_impl(int device, int num_sms, int64_t* launches, const float* Y, int64_t y_stride_t, int64_t y_stride_b, int M,
                      const float* X, int64_t x_stride_t, int64_t x_stride_b, int N, int T, int B, float* Cout, int64_t ldc,
                      int accumulate, int reuse_y, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!Y || !X || !Cout || !workspace) return fail(TACO2DEC_E_ARG, "null argument");
  if (M < 1 || N < 1 || T < 1 || B < 1 || ldc < N) return fail(TACO2DEC_E_ARG, "bad shape");
  const wg::Plan pl = wg::plan(M, N, T * B, num_sms);
  if (workspace_bytes < pl.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(TACO2DEC_E_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(device));
  char* ws = (char*)workspace;
  unsigned* amax = (unsigned*)ws;
  float* scale2 = (float*)(ws + 64);
  unsigned char* a_t = (unsigned char*)(ws + pl.a_off);
  unsigned char* x_t = (unsigned char*)(ws + pl.x_off);
  float* part = (float*)(ws + pl.part_off);
  const int kbs = pl.Kpad / 64;
  if (!reuse_y) {       // the packed, scaled image of Y at the head of the workspace is still valid otherwise
    CUDA_TRY(cudaMemsetAsync(amax, 0, 64, st));
    wg::wg_absmax_kernel<<<num_sms * 4, 256, 0, st>>>(Y, y_stride_t, y_stride_b, T, B, M, amax);
    wg::wg_scale_kernel<<<1, 1, 0, st>>>(amax, scale2);
    wg::wg_pack_T_kernel<<<dim3(kbs, pl.Mpad / 128), 256, 0, st>>>(Y, y_stride_t, y_stride_b, T, B, M, pl.Kpad, scale2, a_t);
    *launches += 3;
  }
  wg::wg_pack_T_kernel<<<dim3(kbs, pl.groups), 256, 0, st>>>(X, x_stride_t, x_stride_b, T, B, N, pl.Kpad, nullptr, x_t);
  CUDA_TRY(tc::prepare_gemm<wg::kNP>());
  tc::GemmParams gp{a_t, x_t, part, pl.Mpad, pl.Kpad, pl.splits, pl.groups, (long long)kbs * wg::kNP * 128, 0, 0, nullptr, 0};
  gp.a_shared = 1;
  CUDA_TRY(tc::launch_gemm<wg::kNP>(gp, st));
  wg::wg_finish_kernel<<<dim3(std::min(64, (M * wg::kNP + 255) / 256), pl.groups), 256, 0, st>>>(part, pl.splits, pl.Mpad, M, N, scale2, Cout, ldc,
                                                                                                accumulate);
  CUDA_TRY(cudaGetLastError());
  *launches += 3;
  return 0;
}

extern "C" int taco2dec_wgrad_gemm(taco2dec_handle* h, const float* Y, int64_t y_stride_t, int64_t y_stride_b, int M, const float* X,
                                   int64_t x_stride_t, int64_t x_stride_b, int N, int T, int B, float* Cout, int64_t ldc, int accumulate,
                                   int reuse_y, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!h) return fail(TACO2DEC_E_ARG, "null argument");
  return wgrad_impl(h->device, h->num_sms, &h->launches, Y, y_stride_t, y_stride_b, M, X, x_stride_t, x_stride_b, N, T, B, Cout, ldc,
                    accumulate, reuse_y, workspace, workspace_bytes, cuda_stream);
}

extern "C" int taco2dec_sgemm_nn(const float* A, int64_t lda, const float* Bm, int64_t ldb, float* Cout, int64_t ldc, int R, int N, int K,
                                 const float* mask, int64_t ldm, int accumulate, void* cuda_stream) {
  if (!A || !Bm || !Cout || R < 1 || N < 1 || K < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  wg::sgemm_nn_kernel<<<dim3((N + 63) / 64, (R + 63) / 64), 256, 0, (cudaStream_t)cuda_stream>>>(A, lda, Bm, ldb, Cout, ldc, R, N, K, mask, ldm,
                                                                                                accumulate);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int taco2dec_bmm_tn(const float* A, int64_t a_stride_b, int64_t a_stride_t, const float* Bm, int64_t b_stride_t, int64_t b_stride_b,
                               float* Cout, int64_t c_stride_b, int64_t ldc, int batch, int M, int N, int T, void* cuda_stream) {
  if (!A || !Bm || !Cout || batch < 1 || M < 1 || N < 1 || T < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  wg::bmm_tn_kernel<<<dim3((N + 63) / 64, (M + 63) / 64, batch), 256, 0, (cudaStream_t)cuda_stream>>>(A, a_stride_b, a_stride_t, Bm, b_stride_t,
                                                                                                     b_stride_b, Cout, c_stride_b, ldc, M, N, T);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Postnet (postnet.cuh)
// ------------------------------------------------------------------------------------------
struct taco2dec_postnet {
  int device, num_sms, n_mel, dim, n_layers;
  int segs;              // 3 = split-fp16 operands (fp32-grade, default), 1 = plain fp16 operands
  bool have_weights;
  pn::Layer layer[pn::kMaxLayers];
  int64_t launches;
};

namespace {
struct PnPlan { int groups; size_t x_bytes, part_bytes, total; };
PnPlan pn_plan(const taco2dec_postnet* h, int B, int T) {
  PnPlan pl;
  pl.groups = (B * T + pn::kNP - 1) / pn::kNP;
  int kmax = 0, cmax = 0;
  for (int l = 0; l < h->n_layers; ++l) { kmax = std::max(kmax, h->layer[l].K); cmax = std::max(cmax, h->layer[l].cout_pad); }
  pl.x_bytes = align_up((size_t)pl.groups * (h->segs * kmax / tc::kBlockK) * pn::kNP * 128, 256);
  pl.part_bytes = align_up((size_t)pl.groups * 4 * cmax * pn::kNP * sizeof(float), 256);     // up to 4 K-splits
  pl.total = 2 * pl.x_bytes + pl.part_bytes;
  return pl;
}
int pn_splits(int ctas_per_split, int kb_total, int num_sms) {
  const int want = std::max(1, num_sms / std::max(1, ctas_per_split));
  int best = 1;
  for (int s : {2, 4}) if (s <= want && kb_total % s == 0) best = s;
  return best;
}
}  // namespace

extern "C" {

int taco2dec_postnet_create(int n_mel, int embed_dim, int kernel_size, int n_layers, int device, taco2dec_postnet** out) {
  if (!out) return fail(TACO2DEC_E_ARG, "null argument");
  if (kernel_size != pn::kTaps) return fail(TACO2DEC_E_ARG, "postnet_kernel_size must be 5");
  if (n_mel < 1 || n_mel > 128 || embed_dim < 128 || embed_dim % 128 || n_layers < 2 || n_layers > pn::kMaxLayers)
    return fail(TACO2DEC_E_ARG, "postnet needs n_mel <= 128, embed_dim a multiple of 128, 2..8 layers");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(TACO2DEC_E_ARCH, "this library is built for sm_100a (B200) only and has no fallback");
  taco2dec_postnet* h = new (std::nothrow) taco2dec_postnet();
  if (!h) return fail(TACO2DEC_E_STATE, "out of host memory");
  h->device = device; h->num_sms = prop.multiProcessorCount; h->n_mel = n_mel; h->dim = embed_dim; h->n_layers = n_layers;
  h->have_weights = false; h->launches = 0; h->segs = 3;
  CUDA_TRY(cudaSetDevice(device));
  for (int l = 0; l < n_layers; ++l) {
    pn::Layer& L = h->layer[l];
    L.cin = l == 0 ? n_mel : embed_dim;
    L.cout = l == n_layers - 1 ? n_mel : embed_dim;
    L.cin_pad = (L.cin + 63) / 64 * 64;
    L.cout_pad = (L.cout + 127) / 128 * 128;
    L.K = pn::kTaps * L.cin_pad;
    CUDA_TRY(cudaMalloc(&L.a_tiles, (size_t)L.cout_pad * 3 * L.K * 2));
    CUDA_TRY(cudaMalloc(&L.bias, (size_t)L.cout_pad * sizeof(float)));
  }
  *out = h;
  return 0;
}

int taco2dec_postnet_set_precision(taco2dec_postnet* h, int fp16_only) {
  if (!h) return fail(TACO2DEC_E_ARG, "null handle");
  const int segs = fp16_only ? 1 : 3;
  if (segs != h->segs) { h->segs = segs; h->have_weights = false; }      // the weight tiles depend on it: set_weights again
  return 0;
}

int taco2dec_postnet_destroy(taco2dec_postnet* h) {
  if (h) for (int l = 0; l < h->n_layers; ++l) { cudaFree(h->layer[l].a_tiles); cudaFree(h->layer[l].bias); }
  delete h;
  return 0;
}

int taco2dec_postnet_set_weights(taco2dec_postnet* h, const taco2dec_postnet_weights* w, void* cuda_stream) {
  if (!h || !w) return fail(TACO2DEC_E_ARG, "null argument");
  if (w->n_layers != h->n_layers) return fail(TACO2DEC_E_ARG, "layer count differs from create()");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  for (int l = 0; l < h->n_layers; ++l) {
    const taco2dec_postnet_layer& s = w->layer[l];
    if (!s.conv_w || !s.conv_b || !s.bn_weight || !s.bn_bias || !s.bn_mean || !s.bn_var)
      return fail(TACO2DEC_E_ARG, "null postnet weight pointer");
    const pn::Layer& L = h->layer[l];
    pn::pn_pack_kernel<<<1024, 256, 0, st>>>(s.conv_w, s.conv_b, s.bn_weight, s.bn_bias, s.bn_mean, s.bn_var, w->bn_eps, L.cout,
                                             L.cin, L.cout_pad, L.cin_pad, L.a_tiles, L.bias, h->segs);
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += h->n_layers;
  h->have_weights = true;
  return 0;
}

size_t taco2dec_postnet_workspace_bytes(const taco2dec_postnet* h, int B, int T) {
  if (!h || B < 1 || T < 1) return 0;
  return pn_plan(h, B, T).total;
}

int taco2dec_postnet_forward(taco2dec_postnet* h, const float* mel, int64_t stride_b, int64_t stride_c, int64_t stride_t,
                             int B, int T, const int64_t* output_lengths, int independent, float* mel_postnet,
                             void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!h || !mel || !mel_postnet || !workspace) return fail(TACO2DEC_E_ARG, "null argument");
  if (!h->have_weights) return fail(TACO2DEC_E_STATE, "weights not set");
  if (B < 1 || T < 1) return fail(TACO2DEC_E_ARG, "B and T must be >= 1");
  if (independent && !output_lengths) return fail(TACO2DEC_E_ARG, "independent utterances need output_lengths");
  const long long* seq_len = independent ? (const long long*)output_lengths : nullptr;
  const PnPlan pl = pn_plan(h, B, T);
  if (workspace_bytes < pl.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(TACO2DEC_E_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  unsigned char* X[2] = {(unsigned char*)workspace, (unsigned char*)workspace + pl.x_bytes};
  float* part = (float*)((char*)workspace + 2 * pl.x_bytes);
  const int n_pad = pl.groups * pn::kNP;
  CUDA_TRY(tc::prepare_gemm<pn::kNP>());
  {
    const pn::Layer& L0 = h->layer[0];
    const size_t total = (size_t)n_pad * pn::kTaps * (L0.cin_pad / 8);
    pn::pn_input_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 65535), 256, 0, st>>>(mel, stride_b, stride_c, stride_t, B, T,
                                                                                               L0.cin, L0.cin_pad, n_pad, seq_len, X[0], h->segs);
  }
  for (int l = 0; l < h->n_layers; ++l) {
    const pn::Layer& L = h->layer[l];
    const int m_tiles = L.cout_pad / tc::kBlockM, kb_total = h->segs * L.K / tc::kBlockK;
    const int splits = pn_splits(m_tiles * pl.groups, kb_total, h->num_sms);
    tc::GemmParams gp{L.a_tiles, X[l & 1], part, L.cout_pad, h->segs * L.K, splits, pl.groups, (long long)kb_total * pn::kNP * 128, 0, 0, nullptr, 0};
    gp.a_shared = 1;
    CUDA_TRY(tc::launch_gemm<pn::kNP>(gp, st));
    if (l + 1 < h->n_layers) {
      const pn::Layer& Ln = h->layer[l + 1];
      pn::pn_pointwise_kernel<<<dim3(pl.groups, L.cout_pad / 8), 256, 0, st>>>(part, splits, L.cout_pad, L.bias, B, T, pl.groups,
                                                                             seq_len, X[(l + 1) & 1], Ln.K, h->segs);
    } else {
      const size_t total = (size_t)B * L.cout * T;
      pn::pn_output_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 65535), 256, 0, st>>>(
          part, splits, L.cout_pad, L.bias, mel, stride_b, stride_c, stride_t, B, T, L.cout, (const long long*)output_lengths, mel_postnet);
    }
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += 1 + 2 * h->n_layers;
  return 0;
}

// ---- training-mode Postnet (postnet_train.cuh): contraction + element-wise building blocks, orchestrated by the caller ----
size_t taco2dec_postnet_rows_gemm_workspace_bytes(const taco2dec_postnet* h, int M, int K, int n_rows) {
  if (!h || M < 1 || K < 1 || n_rows < 1) return 0;
  return pt::rows_plan(M, K, n_rows, h->num_sms).total;
}

int taco2dec_postnet_rows_gemm(taco2dec_postnet* h, const float* X, int64_t x_stride_b, int64_t x_stride_t, int B, int T, int K,
                               const float* W, int M, const float* bias, int scale_x, float* out, int64_t ldo, double* stats,
                               void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!h || !X || !W || !out || !workspace) return fail(TACO2DEC_E_ARG, "null argument");
  if (B < 1 || T < 1 || K < 1 || M < 1 || ldo < M) return fail(TACO2DEC_E_ARG, "bad shape");
  const pt::RowsPlan pl = pt::rows_plan(M, K, B * T, h->num_sms);
  if (workspace_bytes < pl.total) return fail(TACO2DEC_E_STATE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(TACO2DEC_E_ARG, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_TRY(cudaSetDevice(h->device));
  char* ws = (char*)workspace;
  unsigned* amax = (unsigned*)ws;
  float* scale2 = (float*)(ws + 64);
  unsigned char* a_t = (unsigned char*)(ws + pl.a_off);
  unsigned char* x_t = (unsigned char*)(ws + pl.x_off);
  float* part = (float*)(ws + pl.part_off);
  const int kbs = pl.Kpad / 64;
  if (scale_x) {        // gradient rows: power-of-two scale from the absolute maximum so that they survive fp16
    CUDA_TRY(cudaMemsetAsync(amax, 0, 64, st));
    wg::wg_absmax_kernel<<<h->num_sms * 4, 256, 0, st>>>(X, x_stride_t, x_stride_b, T, B, K, amax);
    wg::wg_scale_kernel<<<1, 1, 0, st>>>(amax, scale2);
  }
  pt::pt_pack_w_kernel<<<(unsigned)std::min<size_t>(((size_t)pl.Mpad * (pl.Kpad / 8) + 255) / 256, 4096), 256, 0, st>>>(W, M, K, pl.Mpad, pl.Kpad, a_t);
  pt::pt_pack_rows_kernel<<<dim3(kbs, pl.groups), 256, 0, st>>>(X, x_stride_b, x_stride_t, B, T, K, pl.Kpad, scale_x ? scale2 : nullptr, x_t);
  CUDA_TRY(tc::prepare_gemm<pt::kNP>());
  tc::GemmParams gp{a_t, x_t, part, pl.Mpad, pl.Kpad, pl.splits, pl.groups, (long long)kbs * pt::kNP * 128, 0, 0, nullptr, 0};
  gp.a_shared = 1;
  CUDA_TRY(tc::launch_gemm<pt::kNP>(gp, st));
  pt::pt_finish_kernel<<<dim3(pl.groups, pl.Mpad / 64), 256, 0, st>>>(part, pl.splits, pl.Mpad, M, B * T, bias, scale_x ? scale2 : nullptr, out,
                                                                      ldo, stats);
  CUDA_TRY(cudaGetLastError());
  h->launches += scale_x ? 6 : 4;
  return 0;
}

static int pt_bn_args(const float* mean, const float* rstd, const float* gamma, const float* beta, int use_tanh, uint64_t seed,
                      int mask_id, float p_drop, const uint8_t* keep, int C, pt::BnArgs* a) {
  if (!mean || !rstd || !gamma || !beta) return fail(TACO2DEC_E_ARG, "null argument");
  if (C < 4 || (C & 3) || C > 1024) return fail(TACO2DEC_E_ARG, "channel count must be a multiple of 4, at most 1024");
  if (!(p_drop >= 0.f && p_drop < 1.f)) return fail(TACO2DEC_E_ARG, "bad dropout probability");
  a->mean = mean; a->rstd = rstd; a->gamma = gamma; a->beta = beta; a->use_tanh = use_tanh ? 1 : 0; a->seed = seed; a->mask_id = mask_id;
  a->thresh = keep_threshold(p_drop); a->keep_scale = 1.0f / (1.0f - p_drop); a->keep = keep;
  return 0;
}

int taco2dec_postnet_bn_act_forward(taco2dec_postnet* h, const float* y, int B, int T, int C, const float* mean, const float* rstd,
                                    const float* gamma, const float* beta, int use_tanh, uint64_t seed, int mask_id, float p_drop,
                                    const uint8_t* keep, float* out, int64_t out_stride_b, int64_t out_stride_t, int64_t out_stride_c,
                                    void* cuda_stream) {
  if (!h || !y || !out || B < 1 || T < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  pt::BnArgs a;
  if (int rc = pt_bn_args(mean, rstd, gamma, beta, use_tanh, seed, mask_id, p_drop, keep, C, &a)) return rc;
  if (out_stride_c == 1 && ((out_stride_b | out_stride_t) & 3)) return fail(TACO2DEC_E_ARG, "output strides must keep 16-byte alignment");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t total = (size_t)B * T * (C / 4);
  pt::pt_bn_act_fwd_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 16), 256, 0, (cudaStream_t)cuda_stream>>>(
      y, B, T, C, a, out, out_stride_b, out_stride_t, out_stride_c);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return 0;
}

int taco2dec_postnet_bn_act_backward(taco2dec_postnet* h, float* d, const float* y, int n_rows, int C, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, int use_tanh, uint64_t seed, int mask_id, float p_drop,
                                     const uint8_t* keep, double* sums, void* cuda_stream) {
  if (!h || !d || !y || !sums || n_rows < 1) return fail(TACO2DEC_E_ARG, "bad argument");
  pt::BnArgs a;
  if (int rc = pt_bn_args(mean, rstd, gamma, beta, use_tanh, seed, mask_id, p_drop, keep, C, &a)) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  pt::pt_bn_act_bwd1_kernel<<<h->num_sms * 4, 256, 0, (cudaStream_t)cuda_stream>>>(d, y, n_rows, C, a, sums);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return 0;
}

int taco2dec_postnet_bn_backward_input(taco2dec_postnet* h, const float* dz, const float* y, int B, int T, int C, const float* mean,
                                       const float* rstd, const float* gamma, const float* mean_dz, const float* mean_dz_zhat,
                                       float* dy_pad, void* cuda_stream) {
  if (!h || !dz || !y || !mean || !rstd || !gamma || !mean_dz || !mean_dz_zhat || !dy_pad || B < 1 || T < 1 || C < 4 || (C & 3))
    return fail(TACO2DEC_E_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t total = (size_t)B * T * (C / 4);
  pt::pt_bn_bwd2_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 16), 256, 0, (cudaStream_t)cuda_stream>>>(
      dz, y, B, T, C, mean, rstd, gamma, mean_dz, mean_dz_zhat, dy_pad);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return 0;
}

size_t taco2dec_postnet_wgrad_workspace_bytes(const taco2dec_postnet* h, int M, int N, int T, int B) {
  if (!h || M < 1 || N < 1 || T < 1 || B < 1) return 0;
  return wg::plan(M, N, T * B, h->num_sms).total;
}

int taco2dec_postnet_wgrad(taco2dec_postnet* h, const float* Y, int64_t y_stride_t, int64_t y_stride_b, int M, const float* X,
                           int64_t x_stride_t, int64_t x_stride_b, int N, int T, int B, float* Cout, int64_t ldc, int accumulate,
                           int reuse_y, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!h) return fail(TACO2DEC_E_ARG, "null argument");
  return wgrad_impl(h->device, h->num_sms, &h->launches, Y, y_stride_t, y_stride_b, M, X, x_stride_t, x_stride_b, N, T, B, Cout, ldc,
                    accumulate, reuse_y, workspace, workspace_bytes, cuda_stream);
}

}  // extern "C"
