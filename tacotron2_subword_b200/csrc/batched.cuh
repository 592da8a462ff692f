// batched.cuh -- tensor-core path for batched decoding (B >= 16), included by taco2dec.cu.
//
// Once batch x hidden makes the LSTM gate products genuine dense contractions they run on tcgen05
// (gemm_tc.cuh): gates[4096, B] = [W_ih | W_hh][4096, K] x [x | h][B, K]^T with fp16 operands and fp32
// accumulation in TMEM, 128 CTAs per GEMM (32 row tiles x split-K).  A frame is a short sequence of
// kernels on one stream, captured once as a CUDA graph (no host synchronisation inside the loop):
//
//   x1 <- prenet[t]            bt_prenet_fr (free-running); teacher-forced: copied by bt_pointwise1 of frame t-1
//   G1  = A1 . X1^T            tcgen05 GEMM, 2 streams x split-K 2        (model.py:337-344)
//   pointwise LSTM 1           -> c1, h1 (fp32) + fp16 tiles of X1', X2   (model.py:340-346)
//   Q   = Wq . h1^T            tcgen05 GEMM (128 rows), split-K 4         (attention.py:56, 368)
//   attention                  bt_attention_sma (SMA) / generic attention_task (LSA), one CTA per (utterance, stream)
//   G2  = A2 . X2^T            tcgen05 GEMM, split-K 4                    (model.py:362-371)
//   pointwise LSTM 2           -> c2, h2; teacher-forced: also advances the frame counter   (model.py:371-373)
//   projection + stop test     free-running only; teacher-forced frames defer it to bt_proj_all after the loop
//   frame counter              free-running only
//
// Every producer writes its fp16 output straight into the pre-tiled operand buffers of the GEMMs that
// consume it (core-matrix layout of gemm_tc.cuh), so no separate packing pass exists.
#pragma once

namespace bt {

constexpr int H = 1024, E = 512, P = 256, A = 128, M = 80;
constexpr int K1 = P + E + H;            // 1792: [prenet | ctx | h1]
constexpr int SPLITS1 = 2, SPLITS2 = 4, SPLITSQ = 4;

// Activations a training-mode forward keeps for the backward pass (caller-provided "saved" buffer; all null
// when nothing is saved).  State arrays have T+1 slots: slot 0 = the zero initial state, slot t+1 = frame t.
struct Saved {
  float* gates1;   // [T][S][5][H][B]  i, f, g, o (activated) and the new cell state before dropout
  float* gates2;   // [T][5][H][B]
  float* h1;       // [T+1][S][B][H]   post-dropout attention-LSTM hidden
  float* ctx;      // [T+1][S][B][E]
  float* h2;       // [T+1][B][H]      post-dropout decoder-LSTM hidden
  float* q;        // [T][S][B][A]
};

struct Bufs {
  // tiled fp16 operands
  unsigned char* a1;    // [S][32 m-tiles][28 kb] attention LSTM weights [W_ih | W_hh]
  unsigned char* a2;    // [32][K2/64]            decoder LSTM weights   [W_ih | W_hh]
  unsigned char* aq;    // [S][1][16]             query weights
  unsigned char* x1;    // [S][28 kb][NPAD x 64]
  unsigned char* x2;    // [K2/64][NPAD x 64]     [h1_0 | ctx_0 | h1_1 | ctx_1 | h2]
  // GEMM partial outputs (fp32)
  float* g1;            // [S][SPLITS1][4096][NPAD]
  float* g2;            // [SPLITS2][4096][NPAD]
  float* gq;            // [S][SPLITSQ][128][NPAD]
  // fp32 state
  float *c1, *c2, *h2f; // [S][B][H], [B][H], [B][H]
  float *w0t[2], *w1t[2]; // transposed prenet weights [M][P], [P][P] (coalesced column reads)
  int NPAD, K2;
  Saved sv;
};

__device__ __forceinline__ void x_store(unsigned char* xbase, int NPAD, int b, int k, float v) {
  const size_t off = (size_t)(k >> 6) * ((size_t)NPAD * 128) + tc::tile_offset_bytes(NPAD, b, k & 63);
  *reinterpret_cast<__half*>(xbase + off) = __float2half(v);
}

// weights: rows x (K0 + K1c) from two row-major fp32 sources -> [row tiles of 128][kb] fp16 tiles
__global__ void pack_concat_tiles_kernel(const float* __restrict__ src0, int K0, const float* __restrict__ src1, int K1c,
                                         int rows, unsigned char* __restrict__ dst) {
  const int K = K0 + K1c, kb_total = K / tc::kBlockK;
  const int rows_pad = (rows + 127) / 128 * 128;
  const size_t total = (size_t)rows_pad * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), r = (int)(i / K);
    float v = 0.f;
    if (r < rows) v = k < K0 ? src0[(size_t)r * K0 + k] : src1[(size_t)r * K1c + (k - K0)];
    const size_t tile = ((size_t)(r / 128) * kb_total + (k / tc::kBlockK)) * tc::kATileBytes;
    *reinterpret_cast<__half*>(dst + tile + tc::tile_offset_bytes(128, r % 128, k % tc::kBlockK)) = __float2half(v);
  }
}

__global__ void bt_init_kernel(Params p, Bufs bf) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
  for (int i = gtid; i < p.S * p.B * H; i += n) bf.c1[i] = 0.f;
  for (int i = gtid; i < p.B * H; i += n) { bf.c2[i] = 0.f; bf.h2f[i] = 0.f; }
  for (int i = gtid; i < p.S * p.B * E; i += n) p.ctx[i] = 0.f;
  for (int s = 0; s < p.S; ++s) {
    const int Ts = p.st[s].Ts;
    for (int i = gtid; i < p.B * Ts; i += n) {
      p.st[s].a_prev[i] = (p.attention == TACO2DEC_ATTN_SMA && (i % Ts) == 0) ? 1.0f : 0.0f;
      p.st[s].a_cum[i] = 0.f;
    }
  }
  if (p.free_running) {
    for (int i = gtid; i < p.B; i += n) { p.n_frames[i] = 0; p.reached_max[i] = 0; }
    if (gtid == 0) *p.done_count = 0;
  }
}

// teacher-forced: hoisted prenet output of frame t -> X1 columns [0, P)
__global__ void bt_prenet_tf_to_x1(Params p, Bufs bf, const int* t_ptr) {
  const int t = *t_ptr;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.S * p.B * P) return;
  const int s = i / (p.B * P), r = i - s * p.B * P, b = r / P, k = r - b * P;
  const float v = p.st[s].pre[((size_t)t * p.B + b) * P + k];
  x_store(bf.x1 + (size_t)s * (K1 / 64) * bf.NPAD * 128, bf.NPAD, b, k, v);
}

__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) { const int r = i / cols, c = i - r * cols; dst[(size_t)c * rows + r] = src[i]; }
}

// free-running: prenet(mel[t-1]) for both streams.  Block = (stream, tile of kPreBT utterances), 1024 threads =
// 4 k-quarters x 256 outputs; weights are read transposed (coalesced across outputs) and shared by the tile.
constexpr int kPreBT = 1;
__global__ void __launch_bounds__(1024) bt_prenet_fr(Params p, Bufs bf, const int* t_ptr) {
  __shared__ float x_s[kPreBT][M], h_s[kPreBT][P], part_s[4][kPreBT][P];
  const int t = *t_ptr;
  if (__ldcg(p.done_count) >= p.B) return;
  const int tiles = (p.B + kPreBT - 1) / kPreBT;
  const int s = blockIdx.x / tiles, b0 = (blockIdx.x - s * tiles) * kPreBT, tid = threadIdx.x;
  const int kq = tid >> 8, o = tid & 255;
  const StreamParams& sp = p.st[s];
  for (int i = tid; i < kPreBT * M; i += 1024) {
    const int bb = i / M, m = i - bb * M, b = b0 + bb;
    x_s[bb][m] = (t > 0 && b < p.B) ? p.mel[((size_t)b * p.Tcap + (t - 1)) * M + m] : 0.f;
  }
  __syncthreads();
  {
    float acc[kPreBT] = {};
    const float* w = bf.w0t[s] + o;
#pragma unroll 10
    for (int k = kq * (M / 4); k < (kq + 1) * (M / 4); ++k) {
      const float wv = __ldg(w + (size_t)k * P);
#pragma unroll
      for (int bb = 0; bb < kPreBT; ++bb) acc[bb] = fmaf(wv, x_s[bb][k], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < kPreBT; ++bb) part_s[kq][bb][o] = acc[bb];
  }
  __syncthreads();
  if (tid < kPreBT * P) {
    const int bb = tid >> 8, b = b0 + bb;
    const float v = (part_s[0][bb][o] + part_s[1][bb][o]) + (part_s[2][bb][o] + part_s[3][bb][o]);
    float mult = 0.f;
    if (b < p.B) mult = keep_mult(sp.keep0, ((size_t)t * p.B + b) * P + o, p.seed, s * 2 + 0, t, b * P + o, p.thresh_pre, 2.0f);
    h_s[bb][o] = fmaxf(v, 0.f) * mult;
  }
  __syncthreads();
  {
    float acc[kPreBT] = {};
    const float* w = bf.w1t[s] + o;
#pragma unroll 16
    for (int k = kq * (P / 4); k < (kq + 1) * (P / 4); ++k) {
      const float wv = __ldg(w + (size_t)k * P);
#pragma unroll
      for (int bb = 0; bb < kPreBT; ++bb) acc[bb] = fmaf(wv, h_s[bb][k], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < kPreBT; ++bb) part_s[kq][bb][o] = acc[bb];
  }
  __syncthreads();
  if (tid < kPreBT * P) {
    const int bb = tid >> 8, b = b0 + bb;
    if (b < p.B) {
      const float v = (part_s[0][bb][o] + part_s[1][bb][o]) + (part_s[2][bb][o] + part_s[3][bb][o]);
      const float mult = keep_mult(sp.keep1, ((size_t)t * p.B + b) * P + o, p.seed, s * 2 + 1, t, b * P + o, p.thresh_pre, 2.0f);
      x_store(bf.x1 + (size_t)s * (K1 / 64) * bf.NPAD * 128, bf.NPAD, b, o, fmaxf(v, 0.f) * mult);
    }
  }
}

// attention-LSTM pointwise: sums the split-K partials, gate order i,f,g,o (nn.LSTMCell), dropout on h and c
// tf_merge (teacher-forced frames): this kernel also (a) copies the frame index to t_ptr[1] for bt_pointwise2, which
// then advances t_ptr[0] itself, and (b) moves the hoisted prenet output of frame t+1 into X1 -- GEMM 1 of frame t has
// already consumed X1 -- so the frame needs neither a prenet-copy nor a counter kernel.
__global__ void bt_pointwise1(Params p, Bufs bf, int* t_ptr, int tf_merge) {
  const int t = *t_ptr;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (tf_merge && i == 0) t_ptr[1] = t;
  if (i >= p.S * p.B * H) return;
  if (p.free_running && __ldcg(p.done_count) >= p.B) return;
  const int s = i / (p.B * H), r = i - s * p.B * H, j = r / p.B, b = r - j * p.B;   // b fastest: coalesced partial reads
  const StreamParams& sp = p.st[s];
  if (tf_merge && j < P && t + 1 < p.T)
    x_store(bf.x1 + (size_t)s * (K1 / 64) * bf.NPAD * 128, bf.NPAD, b, j, sp.pre[((size_t)(t + 1) * p.B + b) * P + j]);
  float pre[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float a = sp.b_ih[g * H + j] + sp.b_hh[g * H + j];
#pragma unroll
    for (int k = 0; k < SPLITS1; ++k) a += bf.g1[(((size_t)s * SPLITS1 + k) * 4 * H + g * H + j) * bf.NPAD + b];
    pre[g] = a;
  }
  const size_t idx = (size_t)b * H + j;
  const float gi = sigmoidf_(pre[0]), gf = sigmoidf_(pre[1]), gg = tanhf(pre[2]), go = sigmoidf_(pre[3]);
  float cn = gf * bf.c1[(size_t)s * p.B * H + idx] + gi * gg;
  float hn = go * tanhf(cn);
  if (bf.sv.gates1) {
    float* sv = bf.sv.gates1 + (((size_t)t * p.S + s) * 5 * H + j) * p.B + b;
    const size_t gs = (size_t)H * p.B;
    sv[0] = gi; sv[gs] = gf; sv[2 * gs] = gg; sv[3 * gs] = go; sv[4 * gs] = cn;
  }
  if (p.training) {
    const float sc = 1.0f / (1.0f - p.p_att);
    const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * s) * p.B * H : nullptr;
    const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * s + 1) * p.B * H : nullptr;
    hn *= keep_mult(kh, idx, p.seed, 4 + 2 * s, t, (int)idx, p.thresh_att, sc);
    cn *= keep_mult(kc, idx, p.seed, 5 + 2 * s, t, (int)idx, p.thresh_att, sc);
  }
  bf.c1[(size_t)s * p.B * H + idx] = cn;
  if (bf.sv.h1) bf.sv.h1[(((size_t)(t + 1) * p.S + s) * p.B) * H + idx] = hn;
  x_store(bf.x1 + (size_t)s * (K1 / 64) * bf.NPAD * 128, bf.NPAD, b, P + E + j, hn);     // next frame's LSTM-1 input
  x_store(bf.x2, bf.NPAD, b, s * (H + E) + j, hn);                                        // this frame's LSTM-2 / query input
}

// attention: q = sum of the query-GEMM partials, then the generic attention task; context -> fp32 + tiles
template <bool kDummy>
__global__ void __launch_bounds__(kThreads, 1) bt_attention(Params p, Bufs bf, const int* t_ptr) {
  extern __shared__ __align__(16) float att_smem[];
  const int t = *t_ptr;
  if (p.free_running && __ldcg(p.done_count) >= p.B) return;
  const int s = blockIdx.x % p.S, b = blockIdx.x / p.S, tid = threadIdx.x;
  if (tid < A) {
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < SPLITSQ; ++k) q += bf.gq[(((size_t)s * SPLITSQ + k) * 128 + tid) * bf.NPAD + b];
    p.q[((size_t)s * p.B + b) * A + tid] = q;
    if (bf.sv.q) bf.sv.q[(((size_t)t * p.S + s) * p.B + b) * A + tid] = q;
  }
  __syncthreads();
  attention_task<true>(p, s, b, t, att_smem);
  for (int d = tid; d < E; d += kThreads) {
    const float c = __ldcg(p.ctx + ((size_t)s * p.B + b) * E + d);
    if (bf.sv.ctx) bf.sv.ctx[(((size_t)(t + 1) * p.S + s) * p.B + b) * E + d] = c;
    x_store(bf.x1 + (size_t)s * (K1 / 64) * bf.NPAD * 128, bf.NPAD, b, P + d, c);          // next frame's LSTM-1 input
    x_store(bf.x2, bf.NPAD, b, s * (H + E) + H + d, c);                                     // this frame's LSTM-2 input
  }
}

// Stepwise-monotonic attention specialised for this path (attention.py:330-398): same arithmetic as
// attention_task<true>, restructured around memory latency -- one CTA sees only ~0.4 MB per frame, so the frame time
// is the length of its dependent-load chain, not bandwidth.  The first 8 memory rows of every context thread and the
// first two rounds of processed-memory rows of every energy warp are requested before anything else (neither depends
// on the query), and the context loop keeps 8 independent 16-byte loads in flight per thread.
constexpr int kCtxPF = 8;
__host__ __device__ inline size_t sma_smem_floats(int Ts) { return 2 * (size_t)A + 3 * (size_t)(Ts + 4) + 4 * (size_t)E; }

__global__ void __launch_bounds__(kThreads, 1) bt_attention_sma(Params p, Bufs bf, const int* t_ptr) {
  extern __shared__ __align__(16) float att_smem[];
  const int t = *t_ptr;
  if (p.free_running && __ldcg(p.done_count) >= p.B) return;
  const int s = blockIdx.x % p.S, b = blockIdx.x / p.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const StreamParams& sp = p.st[s];
  const int Ts = sp.Ts;
  const int len = sp.len ? (int)sp.len[b] : Ts;
  float* red_s = att_smem;                 // 4*E context partials (16-byte aligned)
  float* q_s = red_s + 4 * E;              // A
  float* v_s = q_s + A;                    // A
  float* e_s = v_s + A;                    // Ts+4: energies -> selection probabilities
  float* ap_s = e_s + Ts + 4;              // Ts+4: ap_s[0] = 0, ap_s[1+j] = alpha_j entering the frame
  float* an_s = ap_s + Ts + 4;             // Ts+4: alpha'_j

  // ---- early requests: context rows (thread = position group jg, float4 column d4) ----
  const int jg = tid >> 7, d4 = tid & 127;
  const float4* mem4 = reinterpret_cast<const float4*>(sp.mem + (size_t)b * Ts * E) + d4;
  float4 pf[kCtxPF];
#pragma unroll
  for (int i = 0; i < kCtxPF; ++i) {
    const int j = jg + 4 * i;
    pf[i] = j < Ts ? __ldg(mem4 + (size_t)j * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // ---- early requests: processed memory of the warp's first two rounds (4 positions x 4 columns per lane each) ----
  const float* pm_b = sp.pm + (size_t)b * Ts * A;
  float pmv[2][4][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      const float* row = pm_b + (size_t)min(warp * 4 + r * (kWarps * 4) + pp, Ts - 1) * A;
#pragma unroll
      for (int c = 0; c < 4; ++c) pmv[r][pp][c] = __ldg(row + lane + 32 * c);
    }

  if (tid < A) {
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < SPLITSQ; ++k) q += bf.gq[(((size_t)s * SPLITSQ + k) * 128 + tid) * bf.NPAD + b];
    p.q[((size_t)s * p.B + b) * A + tid] = q;
    if (bf.sv.q) bf.sv.q[(((size_t)t * p.S + s) * p.B + b) * A + tid] = q;
    q_s[tid] = q;
    v_s[tid] = sp.v[tid];
  }
  for (int j = tid; j < Ts; j += kThreads) ap_s[1 + j] = __ldcg(sp.a_prev + (size_t)b * Ts + j);
  if (tid == 0) ap_s[0] = 0.f;
  __syncthreads();

  // ---- energies e_j = v . tanh(q + pm_j), masked (attention.py:340-345, 389) ----
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
      if ((lane & 7) == 0 && j < Ts) e_s[j] = (j >= len) ? -INFINITY : ev;
    };
    if (warp * 4 < Ts) round_of(warp * 4, pmv[0]);
    if (warp * 4 + kWarps * 4 < Ts) round_of(warp * 4 + kWarps * 4, pmv[1]);
    for (int j0 = warp * 4 + 2 * kWarps * 4; j0 < Ts; j0 += kWarps * 4) {
      float x[4][4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const float* row = pm_b + (size_t)min(j0 + pp, Ts - 1) * A;
#pragma unroll
        for (int c = 0; c < 4; ++c) x[pp][c] = __ldg(row + lane + 32 * c);
      }
      round_of(j0, x);
    }
  }
  __syncthreads();
  // ---- p_j = sigmoid(e_j [+ 2 N(0,1)])  (attention.py:346-352) ----
  for (int j = tid; j < Ts; j += kThreads) {
    float e = e_s[j];
    if (p.training) {
      const size_t ni = ((size_t)t * p.B + b) * Ts + j;
      const float nz = sp.noise ? sp.noise[ni] : philox_normal(p.seed, 10 + s, t, b * Ts + j);
      e = e + nz * 2.0f;
    }
    const float pj = sigmoidf_(e);
    e_s[j] = pj;
    if (sp.p_save) sp.p_save[((size_t)t * p.B + b) * Ts + j] = pj;
  }
  __syncthreads();
  // ---- alpha'_j = alpha_j p_j + alpha_{j-1} (1 - p_{j-1})  (attention.py:330-338) ----
  {
    float* align_out = sp.align + ((size_t)b * p.Tcap + t) * Ts;
    for (int j = tid; j < Ts; j += kThreads) {
      float a = ap_s[1 + j] * e_s[j];
      if (j > 0) a += ap_s[j] * (1.0f - e_s[j - 1]);
      if ((p.free_running || p.independent) && j >= len) a = 0.0f;   // independent utterances: padded positions do not exist
      an_s[j] = a;
      sp.a_prev[(size_t)b * Ts + j] = a;
      align_out[j] = a;
    }
  }
  __syncthreads();
  // ---- context = alpha' . memory  (attention.py:395) ----
  {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kCtxPF; ++i) {
      const int j = jg + 4 * i;
      const float a = j < Ts ? an_s[j] : 0.f;
      acc.x = fmaf(a, pf[i].x, acc.x); acc.y = fmaf(a, pf[i].y, acc.y); acc.z = fmaf(a, pf[i].z, acc.z); acc.w = fmaf(a, pf[i].w, acc.w);
    }
    for (int jb = jg + 4 * kCtxPF; jb < Ts; jb += 4 * kCtxPF) {
#pragma unroll
      for (int i = 0; i < kCtxPF; ++i) {
        const int j = jb + 4 * i;
        pf[i] = j < Ts ? __ldg(mem4 + (size_t)j * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < kCtxPF; ++i) {
        const int j = jb + 4 * i;
        const float a = j < Ts ? an_s[j] : 0.f;
        acc.x = fmaf(a, pf[i].x, acc.x); acc.y = fmaf(a, pf[i].y, acc.y); acc.z = fmaf(a, pf[i].z, acc.z); acc.w = fmaf(a, pf[i].w, acc.w);
      }
    }
    reinterpret_cast<float4*>(red_s)[jg * (E / 4) + d4] = acc;
  }
  __syncthreads();
  for (int d = tid; d < E; d += kThreads) {
    const float c = (red_s[d] + red_s[E + d]) + (red_s[2 * E + d] + red_s[3 * E + d]);
    p.ctx[((size_t)s * p.B + b) * E + d] = c;
    if (bf.sv.ctx) bf.sv.ctx[(((size_t)(t + 1) * p.S + s) * p.B + b) * E + d] = c;
    x_store(bf.x1 + (size_t)s * (K1 / 64) * bf.NPAD * 128, bf.NPAD, b, P + d, c);          // next frame's LSTM-1 input
    x_store(bf.x2, bf.NPAD, b, s * (H + E) + H + d, c);                                     // this frame's LSTM-2 input
  }
}

__global__ void bt_pointwise2(Params p, Bufs bf, int* t_ptr, int tf_merge) {
  const int t = tf_merge ? t_ptr[1] : t_ptr[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (tf_merge && i == 0) t_ptr[0] = t + 1;     // last kernel of a teacher-forced frame; its other blocks read t_ptr[1]
  if (i >= p.B * H) return;
  if (p.free_running && __ldcg(p.done_count) >= p.B) return;
  const int j = i / p.B, b = i - j * p.B;
  float pre[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float a = p.d_b_ih[g * H + j] + p.d_b_hh[g * H + j];
#pragma unroll
    for (int k = 0; k < SPLITS2; ++k) a += bf.g2[((size_t)k * 4 * H + g * H + j) * bf.NPAD + b];
    pre[g] = a;
  }
  const size_t idx = (size_t)b * H + j;
  const float gi = sigmoidf_(pre[0]), gf = sigmoidf_(pre[1]), gg = tanhf(pre[2]), go = sigmoidf_(pre[3]);
  float cn = gf * bf.c2[idx] + gi * gg;
  float hn = go * tanhf(cn);
  if (bf.sv.gates2) {
    float* sv = bf.sv.gates2 + ((size_t)t * 5 * H + j) * p.B + b;
    const size_t gs = (size_t)H * p.B;
    sv[0] = gi; sv[gs] = gf; sv[2 * gs] = gg; sv[3 * gs] = go; sv[4 * gs] = cn;
  }
  if (p.training) {
    const float sc = 1.0f / (1.0f - p.p_dec);
    const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 4) * p.B * H : nullptr;
    const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 5) * p.B * H : nullptr;
    hn *= keep_mult(kh, idx, p.seed, 8, t, (int)idx, p.thresh_dec, sc);
    cn *= keep_mult(kc, idx, p.seed, 9, t, (int)idx, p.thresh_dec, sc);
  }
  bf.c2[idx] = cn;
  bf.h2f[idx] = hn;
  if (bf.sv.h2) bf.sv.h2[(size_t)(t + 1) * p.B * H + idx] = hn;
  x_store(bf.x2, bf.NPAD, b, p.S * (H + E) + j, hn);     // next frame's recurrent input
}

// mel / gate projection (one warp per (row, batch)) + stop test
__global__ void __launch_bounds__(256) bt_proj(Params p, Bufs bf, const int* t_ptr) {
  const int t = *t_ptr;
  if (p.free_running && __ldcg(p.done_count) >= p.B) return;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (M + 1) * p.B) return;
  const int row = gw / p.B, b = gw - row * p.B;
  const int KD = H + p.S * E;
  const float* w = row < M ? p.proj_w + (size_t)row * KD : p.gate_w;
  float acc = 0.f;
  for (int k = lane * 4; k < KD; k += 128) {
    const float4 a = *reinterpret_cast<const float4*>(w + k);
    float4 x;
    if (k < H) x = *reinterpret_cast<const float4*>(bf.h2f + (size_t)b * H + k);
    else { const int kk = k - H, s = kk / E, d = kk - s * E; x = *reinterpret_cast<const float4*>(p.ctx + ((size_t)s * p.B + b) * E + d); }
    acc = fmaf(a.x, x.x, acc); acc = fmaf(a.y, x.y, acc); acc = fmaf(a.z, x.z, acc); acc = fmaf(a.w, x.w, acc);
  }
  acc = warp_sum(acc);
  if (lane != 0) return;
  if (row < M) {
    p.mel[((size_t)b * p.Tcap + t) * M + row] = acc + p.proj_b[row];
  } else {
    const float g = acc + p.gate_b[0];
    p.gate[(size_t)b * p.Tcap + t] = g;
    if (p.free_running && p.n_frames[b] == 0) {
      if (sigmoidf_(g) > p.gate_thr) { p.n_frames[b] = t + 1; atomicAdd(p.done_count, 1); }
      else if (t + 1 == p.max_steps) { p.n_frames[b] = t + 1; p.reached_max[b] = 1; atomicAdd(p.done_count, 1); }
    }
  }
}

// Teacher-forced mode: nothing inside the frame loop consumes mel / gate (the next prenet input is the target frame,
// model.py:417-424), so the projection of ALL frames is one pass after the loop over the stored h2 / context rows:
//   [mel | gate][(t,b)][81] = [h2 | ctx | ctx_bert][(t,b)][KD] . [Wp ; wg]^T + bias          (model.py:382-388)
// fp32 on the CUDA cores (exact like the in-loop version): 64-row x 81-column tile per block, K in chunks of 32.
constexpr int kPaRows = 64, kPaK = 32, kPaCols = 96;
__global__ void __launch_bounds__(256) bt_proj_all(Params p, Bufs bf) {
  __shared__ float y_s[kPaRows][kPaK + 1];
  __shared__ float w_s[kPaCols][kPaK + 1];
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  const int KD = H + p.S * E, R = p.T * p.B;
  const int r0 = blockIdx.x * kPaRows;
  float acc[8][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = 0.f; }
  // Blocks start at staggered K chunks: in lockstep they would all read the same 128-byte column of rows that lie
  // 2-4 KB apart, which camps on a few HBM channels (measured: 1 ms vs 39 ms for the same launch).
  const int n_chunks = KD / kPaK;
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int k0 = ((kc + blockIdx.x * 5) % n_chunks) * kPaK;
    // activations: row (t, b) of slot t+1; the chunk lies inside one of the segments h2 | ctx_0 | ctx_1
#pragma unroll
    for (int pass = 0; pass < kPaRows / 8; ++pass) {
      const int rl = pass * 8 + rg, r = r0 + rl;
      float v = 0.f;
      if (r < R) {
        const int t = r / p.B, b = r - t * p.B, k = k0 + lane;
        if (k < H) v = bf.sv.h2[((size_t)(t + 1) * p.B + b) * H + k];
        else { const int kk = k - H, s = kk / E, d = kk - s * E; v = bf.sv.ctx[(((size_t)(t + 1) * p.S + s) * p.B + b) * E + d]; }
      }
      y_s[rl][lane] = v;
    }
#pragma unroll
    for (int pass = 0; pass < kPaCols / 8; ++pass) {
      const int c = pass * 8 + rg;
      float v = 0.f;
      if (c < M) v = __ldg(p.proj_w + (size_t)c * KD + k0 + lane);
      else if (c == M) v = __ldg(p.gate_w + k0 + lane);
      w_s[c][lane] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kPaK; ++kk) {
      const float w0 = w_s[lane][kk], w1 = w_s[lane + 32][kk], w2 = w_s[lane + 64][kk];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = y_s[rg * 8 + i][kk];
        acc[i][0] = fmaf(y, w0, acc[i][0]); acc[i][1] = fmaf(y, w1, acc[i][1]); acc[i][2] = fmaf(y, w2, acc[i][2]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + rg * 8 + i;
    if (r >= R) continue;
    const int t = r / p.B, b = r - t * p.B;
    float* mel = p.mel + ((size_t)b * p.Tcap + t) * M;
    mel[lane] = acc[i][0] + p.proj_b[lane];
    mel[lane + 32] = acc[i][1] + p.proj_b[lane + 32];
    if (lane + 64 < M) mel[lane + 64] = acc[i][2] + p.proj_b[lane + 64];
    else if (lane + 64 == M) p.gate[(size_t)b * p.Tcap + t] = acc[i][2] + p.gate_b[0];
  }
}

// frame counter lives in device memory so that one captured CUDA graph serves every frame
__global__ void bt_advance(int* t_ptr) { *t_ptr += 1; }

}  // namespace bt
