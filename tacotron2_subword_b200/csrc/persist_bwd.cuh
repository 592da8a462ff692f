// persist_bwd.cuh -- persistent back-propagation through time for the teacher-forced decoder (SMA, 2 <= B <= 128): ONE
// cooperative launch runs every frame of the reverse-time loop that backward.cuh replays as a 5-kernel graph per frame.
// Included by taco2dec.cu after persist.cuh and backward.cuh (same arithmetic, same buffers, same gradient rows).
//
// Per frame t (descending) the recurrence is                                              (model.py:322-390 backwards)
//   dG2[t] = decoder-LSTM cell backward( dh2[t] = Wd_hh^T dG2[t+1] + projection gradient )
//   dX2[t] = [Wd_ih | Wd_hh]^T dG2[t]                       tcgen05, bf16 operands, 32 row tiles x 4 K-splits = 128 CTAs
//   attention backward per (utterance, stream): d ctx[t] -> d alpha' -> d energies -> dq[t], d processed_memory, dv
//   dG1[t] = attention-LSTM cell backward( dh1[t] = rows of dX2[t] + rows of dX1[t+1] + Wq^T dq[t] )
//   dX1[t] = [W_ih | W_hh]^T dG1[t]                         tcgen05, 14 row tiles x 4 K-splits per stream = 112 CTAs
// The decoder-LSTM chain (dG2 -> dX2 -> dG2 of the frame before) does not depend on the attention side at all, so every CTA runs
// it ONE FRAME AHEAD: its product and epilogue hide behind the attention chain, which is the critical path
// (attention[t] -> dG1[t] -> dX1[t] -> attention[t-1]).
//
// Roles per CTA are those of persist.cuh: warp 16 = TMA producer (lane i owns tile i of the CTA's 16 + 16 k-blocks; weight tiles
// are the transposed bf16 tiles backward.cuh packs, resident in tensor memory / shared memory first, streamed from L2 otherwise),
// warps 17/18 = one MMA-issuing thread per product, warps 0-15 = epilogues, LSTM cell updates, attention tasks.  CTAs exchange
// through L2 with one release-increment of a counter per CTA and phase; every wait has a watchdog.
#pragma once

namespace pbw {

using bt::A;
using bt::E;
using bt::H;
using bt::P;
using bt::K1;
using bw::G;

constexpr int kCtas = 128;
constexpr int kCT = 512;
constexpr int kThreads = kCT + 96;
constexpr int kKb = 16;                       // k-blocks per CTA and product (G / 64 / 4 K-splits)
constexpr int kSplits = 4;
constexpr int kRowTiles1 = K1 / 128;          // 14 per stream
enum { F_DG2 = 0, F_X2 = 1, F_DQ = 2, F_DG1 = 4, F_X1 = 6, F_COUNT = 8 };
constexpr int kFlagStride = 32;

struct PbwParams {
  const unsigned char* a1t;     // [S][14][64] transposed bf16 weight tiles (rows = [prenet | ctx | h1] features, K = gate rows)
  const unsigned char* a2t;     // [K2/128][64]
  unsigned char* dg1t;          // [S][64 kb][NPAD x 64] bf16 tiles of dG1[t]
  unsigned char* dg2t;          // [64 kb][NPAD x 64]
  float* dx1;                   // [2 parities][S][4][K1][NPAD]
  float* dx2;                   // [2 parities][4][K2][NPAD]
  float* dxc1;                  // [2 parities][S][4][NPAD][E]   context rows of dX1, utterance-major (read by the attention tasks)
  float* dxc2;                  // [2 parities][4][S][NPAD][E]   context rows of dX2
  const __half* mem16[2];       // fp16 copies of memory [B][Ts][E] and processed memory [B][Ts][A] for the attention tasks: half the
  const __half* pm16[2];        // bytes through one SM's L2 port, twice the rows in flight (made once per call)
  unsigned* flags;              // [F_COUNT][kFlagStride]
  int K2;
  int att_chunk;                // positions per attention sub-task
  long long* dbg;               // optional [64] SM-clock stamps of CTA 0 during step dbg_step (product 0): see tools/pbw_phases.py
  int dbg_step;
  int w2_stream;                // 1: the decoder-LSTM product's streamed weight tiles are read with evict_first (it has a frame of slack)
  int stages_a0, stages_a1, stages_x0, stages_x1, n_res, n_tm;   // ring depths per product: streamed weight tiles, activation tiles
};

// per-frame re-read operands of the attention tasks (memory, processed memory) must stay in L2 next to the 33 MB of streamed weight
// tiles; everything that is touched once per call (saved activations, gradient rows) is read / written with streaming hints
__device__ __forceinline__ float4 ld_keep_f4(const float* p, unsigned long long policy) {
  const uint4 r = lat::ldg_stream(reinterpret_cast<const unsigned char*>(p), policy);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}
__device__ __forceinline__ float ld_keep_f(const float* p, unsigned long long policy) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));
  return v;
}

__device__ __forceinline__ uint4 ld_keep_u4(const void* p, unsigned long long policy) { return lat::ldg_stream(reinterpret_cast<const unsigned char*>(p), policy); }
__device__ __forceinline__ uint2 ld_keep_u2(const void* p, unsigned long long policy) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ float2 h2f(unsigned u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4; i < n; i += (size_t)gridDim.x * blockDim.x * 4) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(src + i);
      __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
      uint2 o; o.x = *reinterpret_cast<unsigned*>(&a); o.y = *reinterpret_cast<unsigned*>(&b);
      *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
      for (size_t k = i; k < n; ++k) dst[k] = __float2half(src[k]);
    }
  }
}

struct Smem { size_t aring, xring, res, out, dq, wq, att, total; };
__host__ __device__ inline size_t att_floats(int max_ts) {      // attention scratch; also holds u[16][NPAD + 1] (<= 16 * 129 floats)
  const size_t a = (size_t)E + 4 * A + 4 * (size_t)(max_ts + 4) + 16;
  return a > 16 * 129 ? a : 16 * 129;
}
__host__ __device__ inline Smem smem_plan(int NPAD, int stages_a, int stages_x, int n_res, int max_ts) {
  Smem s;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 127) & ~(size_t)127; return o; };
  s.aring = take((size_t)stages_a * tc::kATileBytes);           // stages_a = slots of both products together
  s.xring = take((size_t)stages_x * (size_t)NPAD * 128);        // stages_x = slots of both products together
  s.res = take((size_t)n_res * tc::kATileBytes);
  s.out = take((size_t)4 * NPAD * 9 * 4);             // staging of 8 units x 4 gates x NPAD utterances
  s.dq = take((size_t)A * (NPAD + 1) * 4);            // dq of the CTA's stream, [a][b]
  s.wq = take((size_t)A * 16 * 4);                    // Wq columns of the CTA's 16 attention-LSTM units, [a][unit]
  s.att = take(att_floats(max_ts) * 4);
  s.total = off;
  return s;
}

template <int NPAD>
__global__ void __launch_bounds__(kThreads, 1) decoder_backward_persistent(const __grid_constant__ Params p, const __grid_constant__ bw::Grads g,
                                                                            const __grid_constant__ PbwParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_a[2][8], empty_a[2][8], full_x[2][8], empty_x[2][8], acc_full[2], acc_empty[2], res_bar;
  __shared__ int s_loc[32];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int s_exit;
  __shared__ volatile int s_ok[2];
  __shared__ volatile unsigned s_xfill0[8];            // rounds filled so far per activation-ring slot of product 0 (lane-per-tile producer)
  __shared__ long long s_ph[16];

  constexpr int kXTileBytes = NPAD * 128;
  constexpr int kTmemCols = 512;
  constexpr int LOC_STREAM = -1, LOC_SMEM = 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const int S = p.S, B = p.B, T = p.T, K2 = q.K2;
  const int n_g2 = (K2 / 128) * kSplits;                 // CTAs with a share of the decoder-LSTM product (128 for two streams)
  const int n_g1s = kRowTiles1 * kSplits;                // CTAs per stream with a share of the attention-LSTM product (56)
  const bool has_g2 = c < n_g2, has_g1 = c < S * n_g1s;
  const int m2 = c / kSplits, sig2 = c % kSplits;
  const int s1 = c / n_g1s, m1 = (c % n_g1s) / kSplits, sig1 = c % kSplits;
  const int n_tiles = (has_g1 ? kKb : 0) + (has_g2 ? kKb : 0);   // program: attention-LSTM tiles first, then decoder-LSTM tiles
  const int NSAg[2] = {q.stages_a0, q.stages_a1};        // weight ring depth per product (0: every tile of the product is resident)
  const int NSXg[2] = {q.stages_x0, q.stages_x1};        // activation ring depth per product
  unsigned* const F = q.flags;
  auto flag = [&](int id) { return F + (size_t)id * kFlagStride; };
  // pointwise ownership: decoder LSTM units [8c, 8c+8); attention LSTM of stream sp, units [16 (c % per), +16)
  const int per = kCtas / S;
  const int sp = c / per, j1 = (c % per) * (H / per);    // H / per = 16 (two streams) or 8 (one stream)
  const int nu1 = H / per;
  // decoder-LSTM cells: with two streams the CTAs of the sub-word stream (short memory: light attention tasks) take all of them
  const bool own2 = S == 2 ? c >= per : true;
  const int n_own2 = S == 2 ? per : kCtas, nu2 = H / n_own2;        // 16 or 8 units
  const int j2 = (S == 2 ? c - per : c) * nu2;

  int max_ts = 0;
  for (int s = 0; s < S; ++s) max_ts = max(max_ts, p.st[s].Ts);
  const Smem spl = smem_plan(NPAD, NSAg[0] + NSAg[1], NSXg[0] + NSXg[1], q.n_res, max_ts);
  unsigned char* aring = smem + spl.aring;
  unsigned char* xring = smem + spl.xring;
  unsigned char* res_s = smem + spl.res;
  float* out_s = (float*)(smem + spl.out);               // [4][NPAD][9]
  float* dq_s = (float*)(smem + spl.dq);                 // [A][NPAD + 1]
  float* wq_s = (float*)(smem + spl.wq);                 // [A][16]
  float* att_s = (float*)(smem + spl.att);

  if (tid == 0) {
    for (int gg = 0; gg < 2; ++gg)
      for (int i = 0; i < 8; ++i) {
        tc::mbar_init(&full_a[gg][i], 1); tc::mbar_init(&empty_a[gg][i], 1); tc::mbar_init(&full_x[gg][i], 1); tc::mbar_init(&empty_x[gg][i], 1);
      }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 1); }
    tc::mbar_init(&res_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_exit = 0; s_ok[0] = 1; s_ok[1] = 1;
    for (int i = 0; i < 16; ++i) s_ph[i] = 0;
    for (int i = 0; i < 8; ++i) s_xfill0[i] = 0;
    // placement: the attention-LSTM product sits on the critical chain -> its tiles go on-chip first
    int tm = 0, sm = 0;
    for (int i = 0; i < 32; ++i) s_loc[i] = LOC_STREAM;
    for (int i = 0; i < n_tiles; ++i) {
      if (tm < q.n_tm) s_loc[i] = tm++;
      else if (sm < q.n_res) s_loc[i] = LOC_SMEM + sm++;
    }
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_s)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < kCT) {
    for (int i = tid; i < A * 16; i += kCT) {
      const int a = i >> 4, u = i & 15;
      wq_s[i] = u < nu1 ? p.st[sp].wq[(size_t)a * H + j1 + u] : 0.f;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t acc_addr[2] = {tmem_base, tmem_base + (uint32_t)NPAD};
  const uint32_t tm_w_col = (uint32_t)(2 * NPAD);
  pb::Ctl ctl{p.abort_flag, &s_exit};
  // tile i of the program: product gg (0 = attention LSTM, 1 = decoder LSTM) and k-block index inside the CTA's K split
  auto tile_gemm = [&](int i) { return (has_g1 && i < kKb) ? 0 : 1; };
  auto tile_src = [&](int i) -> const unsigned char* {
    if (has_g1 && i < kKb) return q.a1t + (((size_t)s1 * kRowTiles1 + m1) * (G / 64) + sig1 * kKb + i) * tc::kATileBytes;
    const int j = has_g1 ? i - kKb : i;
    return q.a2t + ((size_t)m2 * (G / 64) + sig2 * kKb + j) * tc::kATileBytes;
  };

  // ---- weight tiles resident in tensor memory (A operand of tcgen05.mma), copied in once ----
  if (tid < kCT && q.n_tm > 0) {
    const int quarter = warp & 3, r = quarter * 32 + lane;
    for (int i = 0; i < n_tiles; ++i) {
      const int loc = s_loc[i];
      if (loc < 0 || loc >= LOC_SMEM || (loc & 3) != (warp >> 2)) continue;
      const unsigned char* src = tile_src(i) + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16;
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

  // =========================== TMA producers ===========================
  // All 16 activation tiles of a product depend on ONE event (the gate gradients of the frame), so each product has one thread
  // that waits for that counter and then feeds the ring in program order, and one thread that keeps the ring of streamed weight
  // tiles full.  (A lane per tile, as in the forward kernel, made every ring-slot check of the attention-LSTM product wait for
  // its sibling lanes' global-memory polls: ~1 kcyc per k-block.)  Threads that poll global memory never share a warp with a
  // thread whose loop is latency-critical at the same time: warp 16 = X producer of product 0; warp 17 = MMA issuer of product 0 +
  // the weight producers of both products (shared-memory barriers only); warp 18 = MMA issuer of product 1 + its X producer
  // (the issuer is idle until those tiles arrive).
  auto x_producer = [&](int gg) {
    const int NSX = NSXg[gg];
    unsigned char* const my_xring = xring + (size_t)gg * NSXg[0] * kXTileBytes;
    const unsigned char* const x_src = gg == 0 ? q.dg1t + ((size_t)s1 * (G / 64) + sig1 * kKb) * kXTileBytes
                                               : q.dg2t + ((size_t)sig2 * kKb) * kXTileBytes;
    const unsigned* fptr = gg == 0 ? flag(F_DG1 + s1) : flag(F_DG2);
    const unsigned mul = gg == 0 ? (unsigned)per : (unsigned)(S == 2 ? per : kCtas);
    unsigned gx = 0;
    for (int step = 0; step < T; ++step) {
      if (!pb::poll_ge(fptr, mul * (unsigned)(step + 1), ctl)) return;
      const bool dbg_on = q.dbg && c == 0 && gg == 0 && step == q.dbg_step;
      if (dbg_on) q.dbg[1] = clock64();        // counter seen
      pb::fence_proxy_async();                 // the tiles were written through the generic proxy by other SMs
      for (int i = 0; i < kKb; ++i, ++gx) {
        const int slot = (int)(gx % (unsigned)NSX);
        const unsigned round = gx / (unsigned)NSX;
        if (round > 0u && !pb::mbar_wait_ab(&empty_x[gg][slot], (round & 1u) ^ 1u, ctl)) return;
        tc::mbar_expect_tx(&full_x[gg][slot], (unsigned)kXTileBytes);
        tc::tma_load_1d(my_xring + (size_t)slot * kXTileBytes, x_src + (size_t)i * kXTileBytes, kXTileBytes, &full_x[gg][slot]);
        if (dbg_on) q.dbg[16 + i] = clock64();  // activation tile i requested
      }
    }
  };
  auto a_producer = [&](int gg) {
    const unsigned long long pol = (gg == 1 && q.w2_stream) ? lat::l2_policy_evict_first() : lat::l2_policy_evict_last();
    unsigned char* const my_aring = aring + (size_t)gg * NSAg[0] * tc::kATileBytes;
    const int NSA = NSAg[gg];
    const int g_lo = gg == 0 ? 0 : (has_g1 ? kKb : 0), g_hi = g_lo + kKb;
    unsigned ga = 0;
    if (NSA == 0) return;
    for (int step = 0; step < T; ++step)
      for (int i = g_lo; i < g_hi; ++i) {
        if (s_loc[i] != LOC_STREAM) continue;
        const int slot = (int)(ga % (unsigned)NSA);
        const unsigned round = ga / (unsigned)NSA;
        if (round > 0u && !pb::mbar_wait_ab(&empty_a[gg][slot], (round & 1u) ^ 1u, ctl)) return;
        tc::mbar_expect_tx(&full_a[gg][slot], (unsigned)tc::kATileBytes);
        pb::tma_load_1d_hint(my_aring + (size_t)slot * tc::kATileBytes, tile_src(i), tc::kATileBytes, &full_a[gg][slot], pol);
        ++ga;
      }
  };
  if (warp == 16) {
    if (lane == 0) {      // shared-memory resident tiles: one-off bulk copies
      const unsigned long long pol_once = lat::l2_policy_evict_first();
      unsigned bytes = 0;
      for (int i = 0; i < n_tiles; ++i) if (s_loc[i] >= LOC_SMEM) bytes += (unsigned)tc::kATileBytes;
      if (bytes) {
        tc::mbar_expect_tx(&res_bar, bytes);
        for (int i = 0; i < n_tiles; ++i)
          if (s_loc[i] >= LOC_SMEM)
            pb::tma_load_1d_hint(res_s + (size_t)(s_loc[i] - LOC_SMEM) * tc::kATileBytes, tile_src(i), tc::kATileBytes, &res_bar, pol_once);
      } else {
        pb::mbar_arrive(&res_bar);
      }
    }
    // activation tiles of product 0 (critical chain): lane i owns k-block i of every frame.  A single thread needs ~0.7 kcyc per
    // tile for the barrier / expect-tx / bulk-copy sequence (measured); sixteen lanes do it side by side.  They all wait for the
    // same counter, so nobody's ring-slot check is delayed by a sibling's global-memory poll.
    if (has_g1 && lane < kKb) {
      const int i = lane, NSX = NSXg[0];
      const unsigned char* const x_src = q.dg1t + ((size_t)s1 * (G / 64) + sig1 * kKb + i) * kXTileBytes;
      const unsigned* fptr = flag(F_DG1 + s1);
      unsigned seen = 0;
      bool ok = true;
      for (int step = 0; step < T && ok; ++step) {
        const unsigned gx = (unsigned)step * (unsigned)kKb + (unsigned)i, target = (unsigned)per * (unsigned)(step + 1);
        const int slot = (int)(gx % (unsigned)NSX);
        const unsigned round = gx / (unsigned)NSX;
        unsigned spins = 0;
        long long t0 = 0;
        bool ready = false;
        for (;;) {
          if (!ready) {
            ready = (int)(seen - target) >= 0;
            if (!ready) {
              seen = ld_acquire_u32(fptr);
              ready = (int)(seen - target) >= 0;
              if (ready) {
                pb::fence_proxy_async();
                if (q.dbg && c == 0 && i == 0 && step == q.dbg_step) q.dbg[1] = clock64();
              }
            }
          }
          if (ready && s_xfill0[slot] == round && (round == 0u || pb::mbar_test(&empty_x[0][slot], (round & 1u) ^ 1u))) {
            tc::mbar_expect_tx(&full_x[0][slot], (unsigned)kXTileBytes);
            tc::tma_load_1d(xring + (size_t)slot * kXTileBytes, x_src, kXTileBytes, &full_x[0][slot]);
            s_xfill0[slot] = round + 1u;
            if (q.dbg && c == 0 && step == q.dbg_step) q.dbg[16 + i] = clock64();
            break;
          }
          if ((++spins & 63u) == 0u) {
            if (s_exit) { ok = false; break; }
            if (*((volatile int*)p.abort_flag) != 0) { s_exit = 1; ok = false; break; }
            if (spins == 4096u) t0 = clock64();
            else if ((spins & 4095u) == 0u && clock64() - t0 > pb::kTimeoutClocks) { atomicExch(p.abort_flag, 1); s_exit = 1; ok = false; break; }
          }
        }
      }
    }
  } else if (warp == 17 || warp == 18) {
    // =========================== MMA issuers: warp 17 = attention-LSTM product, warp 18 = decoder-LSTM product ==========
    const int gg = warp - 17;
    if (warp == 18 && lane == 1) {
      if (has_g2) x_producer(1);
    } else if (warp == 17 && lane == 1) {
      if (has_g1) a_producer(0);
    } else if (warp == 17 && lane == 2) {
      if (has_g2) a_producer(1);
    } else if (lane == 0 && (gg == 0 ? has_g1 : has_g2)) {
      const uint32_t idesc = tc::make_idesc_f16(128, NPAD) | tc::kFmtBF16;
      constexpr uint32_t lbo_a = (128 / 8) * 128, lbo_x = (NPAD / 8) * 128, sbo = 128;
      const int i_lo = gg == 0 ? 0 : (has_g1 ? kKb : 0), i_hi = i_lo + kKb;
      unsigned char* const my_aring = aring + (size_t)gg * NSAg[0] * tc::kATileBytes;
      const int NSA = NSAg[gg] > 0 ? NSAg[gg] : 1;
      unsigned char* const my_xring = xring + (size_t)gg * NSXg[0] * kXTileBytes;
      const int NSX = NSXg[gg];
      const uint64_t desc_hi_a = tc::make_smem_desc(0u, lbo_a, sbo), desc_hi_x = tc::make_smem_desc(0u, lbo_x, sbo);
      const uint32_t a_ring_addr = tc::smem_u32(my_aring), x_ring_addr = tc::smem_u32(my_xring), res_addr = tc::smem_u32(res_s);
      const uint32_t acc = acc_addr[gg];
      bool ok = pb::mbar_wait_ab(&res_bar, 0u, ctl);
      // The issuing thread is the bottleneck of a product (measured: ~0.95 kcyc per k-block with per-tile table look-ups and ring
      // arithmetic by division): the placement of the program is three consecutive ranges -- tensor memory, shared memory,
      // streamed -- so addresses advance linearly, and the ring slots / parities are carried incrementally.
      int n_t = 0, n_s = 0;                       // tiles of this product in tensor memory / resident in shared memory
      for (int j = i_lo; j < i_hi; ++j) { const int l = s_loc[j]; n_t += (l >= 0 && l < LOC_SMEM); n_s += l >= LOC_SMEM; }
      const int loc_first = s_loc[i_lo];
      const uint32_t a_tm0 = tmem_base + tm_w_col + (uint32_t)((n_t > 0 ? loc_first : 0) * 32);
      int first_res = 0;
      for (int j = i_lo; j < i_hi; ++j) if (s_loc[j] >= LOC_SMEM) { first_res = s_loc[j] - LOC_SMEM; break; }
      const uint32_t res0 = res_addr + (uint32_t)first_res * (uint32_t)tc::kATileBytes;
      int sx = 0, sa = 0;
      uint32_t px = 0, pa = 0;
      constexpr uint64_t kx2 = (uint64_t)((2 * lbo_x) >> 4), ka2 = (uint64_t)((2 * lbo_a) >> 4);
      for (int step = 0; step < T && ok; ++step) {
        if (step > 0) {
          ok = pb::mbar_wait_ab(&acc_empty[gg], (uint32_t)((step - 1) & 1), ctl);
          if (!ok) break;
        }
        const bool dbg_on = q.dbg && c == 0 && gg == 0 && step == q.dbg_step;
        uint32_t a_tm = a_tm0, a_res = res0;
#pragma unroll 1
        for (int i = 0; i < kKb; ++i) {
          if (!pb::mbar_try(&full_x[gg][sx], px)) ok = pb::mbar_wait_ab(&full_x[gg][sx], px, ctl);
          const bool streamed = i >= n_t + n_s;
          if (ok && streamed && !pb::mbar_try(&full_a[gg][sa], pa)) ok = pb::mbar_wait_ab(&full_a[gg][sa], pa, ctl);
          if (!ok) break;
          if (dbg_on) q.dbg[32 + i] = clock64();        // operands of tile i landed
          tc::tc_fence_after();
          const uint64_t dx0 = desc_hi_x | (uint64_t)(((x_ring_addr + (uint32_t)sx * (uint32_t)kXTileBytes) >> 4) & 0x3fffu);
          const uint32_t first = i == 0 ? 0u : 1u;
          if (i < n_t) {
            pb::umma_f16_ts(acc, a_tm, dx0, idesc, first);
            pb::umma_f16_ts(acc, a_tm + 8u, dx0 + kx2, idesc, 1u);
            pb::umma_f16_ts(acc, a_tm + 16u, dx0 + 2 * kx2, idesc, 1u);
            pb::umma_f16_ts(acc, a_tm + 24u, dx0 + 3 * kx2, idesc, 1u);
            a_tm += 32u;
          } else {
            const uint32_t a_addr = streamed ? a_ring_addr + (uint32_t)sa * (uint32_t)tc::kATileBytes : a_res;
            const uint64_t da0 = desc_hi_a | (uint64_t)((a_addr >> 4) & 0x3fffu);
            tc::umma_f16(acc, da0, dx0, idesc, first);
            tc::umma_f16(acc, da0 + ka2, dx0 + kx2, idesc, 1u);
            tc::umma_f16(acc, da0 + 2 * ka2, dx0 + 2 * kx2, idesc, 1u);
            tc::umma_f16(acc, da0 + 3 * ka2, dx0 + 3 * kx2, idesc, 1u);
            if (!streamed) a_res += (uint32_t)tc::kATileBytes;
          }
          tc::umma_commit(&empty_x[gg][sx]);
          if (++sx == NSX) { sx = 0; px ^= 1u; }
          if (streamed) {
            tc::umma_commit(&empty_a[gg][sa]);
            if (++sa == NSA) { sa = 0; pa ^= 1u; }
          }
          if (i == kKb - 1) tc::umma_commit(&acc_full[gg]);
          if (dbg_on) q.dbg[48 + i] = clock64();        // tile i issued
        }
      }
    }
  } else {
    // =========================== compute warps ===========================
    int wn = 0;
    long long ph_t = clock64();
#define PBW_PH(slot)                                      \
    if (c == 0 && tid == 0) {                             \
      const long long n_ = clock64();                     \
      s_ph[slot] += n_ - ph_t;                            \
      ph_t = n_;                                          \
    }
#define PBW_WAIT_FLAG(fptr, target)                                               \
    {                                                                             \
      if (tid == 0) s_ok[wn & 1] = pb::poll_ge((fptr), (target), ctl) ? 1 : 0;    \
      pb::bar_compute();                                                          \
      const int ok_ = s_ok[wn & 1];                                               \
      ++wn;                                                                       \
      if (!ok_) goto pbw_done;                                                    \
    }
#define PBW_WAIT_MBAR(bar, parity)                                                \
    {                                                                             \
      if (tid == 0) s_ok[wn & 1] = pb::mbar_wait_ab((bar), (parity), ctl) ? 1 : 0;\
      pb::bar_compute();                                                          \
      const int ok_ = s_ok[wn & 1];                                               \
      ++wn;                                                                       \
      if (!ok_) goto pbw_done;                                                    \
    }
    const size_t dx1_par = (size_t)S * kSplits * K1 * NPAD, dx1_str = (size_t)kSplits * K1 * NPAD;
    const size_t dx2_par = (size_t)kSplits * K2 * NPAD;
    const float sc_att = 1.0f / (1.0f - p.p_att), sc_dec = 1.0f / (1.0f - p.p_dec);
    const int row_ep = (warp & 3) * 32 + lane;
    float dc2_r[4] = {0.f, 0.f, 0.f, 0.f};     // carries: d cell state of the CTA's cells (fixed cell -> thread mapping)
    float dc1_r[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t gs = (size_t)H * B;
    const unsigned long long pol_keep = lat::l2_policy_evict_last();
    const int parts0 = (p.st[0].Ts + q.att_chunk - 1) / q.att_chunk, parts1 = S == 2 ? (p.st[1].Ts + q.att_chunk - 1) / q.att_chunk : 0;
    const int n_sub0 = B * parts0, n_sub = n_sub0 + B * parts1;
    const int parts_sp = sp == 0 ? parts0 : parts1;

    // gate gradients staged in out_s[q][b][jl] -> bf16 operand tile chunk + fp32 rows (8 consecutive units per (gate, utterance))
    auto store_gates = [&](unsigned char* tiles, float* rows, int j0) {
      if (tid < 4 * B) {
        const int qg = tid / B, b = tid - qg * B;
        const float* v = out_s + ((size_t)qg * NPAD + b) * 9;
        const int k = qg * H + j0;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<unsigned*>(&h0); pk.y = *reinterpret_cast<unsigned*>(&h1);
        pk.z = *reinterpret_cast<unsigned*>(&h2); pk.w = *reinterpret_cast<unsigned*>(&h3);
        *reinterpret_cast<uint4*>(tiles + (size_t)(k >> 6) * ((size_t)NPAD * 128) + tc::tile_offset_bytes(NPAD, b, k & 63)) = pk;
        float4* dst = reinterpret_cast<float4*>(rows + (size_t)b * G + k);
        __stcs(dst, make_float4(v[0], v[1], v[2], v[3]));
        __stcs(dst + 1, make_float4(v[4], v[5], v[6], v[7]));
      }
    };

    // decoder-LSTM cell backward of frame t for units [j2, j2 + 8) x all utterances; dh2 = projection rows + (Wd_hh^T dG2[t+1])
    auto pointwise2 = [&](int t, bool have_next) {
      const float* dxn = q.dx2 + (size_t)((t + 1) & 1) * dx2_par;
      for (int half = 0; half * 8 < nu2; ++half) {
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int e = tid + ci * kCT;
          if (e >= 8 * B) break;
          const int jl = e / B, b = e - jl * B, j = j2 + half * 8 + jl;
          const size_t idx = (size_t)b * H + j;
          float dh = g.dyh[((size_t)t * H + j) * B + b];
          if (have_next) {
#pragma unroll
            for (int k = 0; k < kSplits; ++k) dh += __ldcg(dxn + ((size_t)k * K2 + S * (H + E) + j) * NPAD + b);
          }
          const float* sv = g.sv.gates2 + ((size_t)t * 5 * H + j) * B + b;
          const float cn_prev = t > 0 ? (sv - 5 * gs)[4 * gs] : 0.f;
          float mh = 1.f, mc = 1.f, mc_prev = 1.f;
          if (p.training) {
            const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 4) * B * H : nullptr;
            const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 5) * B * H : nullptr;
            mh = keep_mult(kh, idx, p.seed, 8, t, (int)idx, p.thresh_dec, sc_dec);
            mc = keep_mult(kc, idx, p.seed, 9, t, (int)idx, p.thresh_dec, sc_dec);
            if (t > 0) {
              const uint8_t* kcp = p.lstm_keep ? p.lstm_keep + ((size_t)(t - 1) * 6 + 5) * B * H : nullptr;
              mc_prev = keep_mult(kcp, idx, p.seed, 9, t - 1, (int)idx, p.thresh_dec, sc_dec);
            }
          }
          float dgate[4];
          bw::lstm_cell_backward(sv, gs, mc_prev * cn_prev, mh, mc, dh, &dc2_r[half * 2 + ci], dgate);
#pragma unroll
          for (int qg = 0; qg < 4; ++qg) out_s[((size_t)qg * NPAD + b) * 9 + jl] = dgate[qg];
        }
        pb::bar_compute();
        store_gates(q.dg2t, g.dg2 + (size_t)t * B * G, j2 + half * 8);
        pb::bar_compute();
      }
    };

    // accumulator (128 rows x NPAD utterances, lane = row) -> split-K partials.  Hidden / prenet rows are consumed with lanes along
    // the utterances: row-major [row][NPAD].  Context rows are consumed by the attention tasks (one utterance, all rows): they go
    // out utterance-major [utterance][E] -- for the writer that is the coalesced direction anyway (32 lanes = 32 consecutive rows).
    auto store_acc = [&](uint32_t acc, float* part_mine, float* ctx_mine) {
      for (int cg = warp >> 2; cg < NPAD / 16; cg += 4) {
        uint32_t v[16];
        const uint32_t taddr = acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cg * 16);
        lat::tmem_ld16(taddr, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ctx_mine) {
#pragma unroll
          for (int i = 0; i < 16; ++i) ctx_mine[(size_t)(cg * 16 + i) * E + row_ep] = __uint_as_float(v[i]);
        } else {
          float4* dst = reinterpret_cast<float4*>(part_mine + (size_t)row_ep * NPAD + cg * 16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            dst[k4] = make_float4(__uint_as_float(v[4 * k4]), __uint_as_float(v[4 * k4 + 1]), __uint_as_float(v[4 * k4 + 2]),
                                  __uint_as_float(v[4 * k4 + 3]));
        }
      }
    };

    // epilogue of the e-th decoder-LSTM product (frame T-1-e): TMEM -> split-K partials of dX2, then the counter
    auto epilogue2 = [&](int e) -> bool {
      if (!has_g2) return true;
      if (tid == 0) s_ok[wn & 1] = pb::mbar_wait_ab(&acc_full[1], (uint32_t)(e & 1), ctl) ? 1 : 0;
      pb::bar_compute();
      const int ok_ = s_ok[wn & 1];
      ++wn;
      if (!ok_) return false;
      tc::tc_fence_after();
      const int tt = T - 1 - e;
      {
        const int row0 = m2 * 128, blk = row0 / (H + E), off = row0 - blk * (H + E);
        const bool ctx_tile = row0 < S * (H + E) && off >= H;
        float* part_mine = q.dx2 + (size_t)(tt & 1) * dx2_par + ((size_t)sig2 * K2 + (size_t)row0) * NPAD;
        float* ctx_mine = q.dxc2 + ((((size_t)(tt & 1) * kSplits + sig2) * S + blk) * NPAD) * E + (off - H);
        store_acc(acc_addr[1], part_mine, ctx_tile ? ctx_mine : nullptr);
      }
      tc::tc_fence_before();
      pb::bar_compute();
      if (tid == 0) { pb::mbar_arrive(&acc_empty[1]); pb::signal(flag(F_X2)); }
      return true;
    };

    // ---- prologue: the decoder-LSTM chain (dG2 -> dX2 -> dG2 of the frame before) does not depend on the attention side; it is
    //      started two cell updates ahead and afterwards advanced inside the window in which the compute warps would otherwise
    //      wait for the attention-LSTM product (see the end of the loop body) ----
    if (own2) {
      pointwise2(T - 1, false);
      if (tid == 0) pb::signal(flag(F_DG2));
    }
    if (!epilogue2(0)) goto pbw_done;
    if (own2 && T > 1) {
      PBW_WAIT_FLAG(flag(F_X2), (unsigned)n_g2)
      pointwise2(T - 2, true);
      if (tid == 0) pb::signal(flag(F_DG2));
    }

    for (int step = 0; step < T; ++step) {
      const int t = T - 1 - step;
      const float* dx2_cur = q.dx2 + (size_t)(t & 1) * dx2_par;
      const float* dx1_nxt = q.dx1 + (size_t)((t + 1) & 1) * dx1_par;      // written by frame t + 1
      float* dx1_cur = q.dx1 + (size_t)(t & 1) * dx1_par;

      PBW_PH(0)
      PBW_WAIT_FLAG(flag(F_X2), (unsigned)n_g2 * (unsigned)(step + 1))      // dX2[t] of every CTA (normally long since there)
      PBW_PH(3)
      PBW_PH(4)
      // the attention side of frame t needs dX1[t+1] (context and hidden rows) of its stream
      if (step > 0) {
        PBW_WAIT_FLAG(flag(F_X1 + sp), (unsigned)n_g1s * (unsigned)step)
        PBW_PH(5)
        // d prenet[t+1] = prenet rows of dX1[t+1]: the CTA saves rows [4 (c % per) ...) of its stream (P / per each)
        const int npk = P / per, k0 = (c % per) * npk;
        for (int e = tid; e < npk * B; e += kCT) {
          const int kl = e / B, b = e - kl * B, k = k0 + kl;
          float acc = 0.f;
#pragma unroll
          for (int qq = 0; qq < kSplits; ++qq) acc += __ldcg(dx1_nxt + (size_t)sp * dx1_str + ((size_t)qq * K1 + k) * NPAD + b);
          __stcs(g.dpre + (((size_t)sp * T + (t + 1)) * B + b) * P + k, acc);
        }
      }

      PBW_PH(6)
      // ---------------- attention backward (attention.py:330-398) ----------------
      // A task (utterance, stream) is position-parallel except for one neighbour value, so it is cut into sub-tasks of `chunk`
      // positions: the phoneme stream's long memories would otherwise make their CTAs the slowest of every frame.  Sub-task u ->
      // (stream, utterance, part); its partial dq / dv sums are added into zero-initialised rows (at most a few addends).
      for (int u = c; u < n_sub; u += kCtas) {
        const int s = u < n_sub0 ? 0 : 1;
        const int ul = s == 0 ? u : u - n_sub0;
        const int b = ul % B, part = ul / B;
        const StreamParams& sa = p.st[s];
        const int Ts = sa.Ts;
        const int len = sa.len ? (int)sa.len[b] : Ts;
        const int jlo = part * q.att_chunk, jhi = min(jlo + q.att_chunk, Ts);     // own positions [jlo, jhi)
        const int jdn = min(jhi + 1, Ts);                                         // d alpha' is also needed at the right neighbour
        constexpr int kW = kCT / 32;
        float* dctx_s = att_s;              // E
        float* dqa_s = dctx_s + E;          // A
        float* dva_s = dqa_s + A;           // A
        const float* mem_b = sa.mem + (size_t)b * Ts * E;
        const float* pm_b = sa.pm + (size_t)b * Ts * A;
        float* dpm_b = g.dpm[s] + (size_t)b * Ts * A;
        // Every warp owns a CONTIGUOUS block of R <= 31 positions [r0, r1) of the sub-task; lane i holds the per-position scalars of
        // position r0 + i.  The only cross-position dependency -- d alpha'[j+1] -- stays inside the warp (a shuffle) except at the
        // block's right edge, where the warp computes the neighbour's value itself (one more row).  No block barrier between the
        // phases: the warps run them back to back on their own rows and hide each other's load latencies.
        const int R = (jhi - jlo + kW - 1) / kW;
        const int r0 = min(jlo + warp * R, jhi), r1 = min(r0 + R, jhi);
        const int rd = min(r1 + 1, Ts);                         // rows whose d alpha' this warp needs (own + right neighbour)
        const int jme = r0 + lane;                              // this lane's position
        const bool own = jme < r1, need = jme < rd && r1 > r0;
        float pj = 0.f, apj = 0.f, base = 0.f;
        if (need) {
          base = __ldcg(g.dalpha[s] + (size_t)b * Ts + jme);    // carry from frame t+1 (the neighbour's was written by another warp / CTA)
          if (g.d_align[s]) base += g.d_align[s][((size_t)b * T + t) * Ts + jme];
        }
        if (own) {
          pj = __ldcs(g.p_saved[s] + ((size_t)t * B + b) * Ts + jme);
          apj = t > 0 ? __ldcs(g.align[s] + ((size_t)b * T + (t - 1)) * Ts + jme) : (jme == 0 ? 1.f : 0.f);
        }
        // memory rows in fp16: lane owns elements [8 (lane + 32 h), +8), h = 0, 1; four rows in flight
        const __half* mem16_b = q.mem16[s] + (size_t)b * Ts * E;
        const __half* pm16_b = q.pm16[s] + (size_t)b * Ts * A;
        uint4 mrow[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int j = r0 + r;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
            mrow[r][hh] = (j < rd && r1 > r0) ? ld_keep_u4(mem16_b + (size_t)j * E + 8 * (lane + 32 * hh), pol_keep) : make_uint4(0u, 0u, 0u, 0u);
        }
        // attention dimensions of this lane: a = 4 lane + i
        float qv[4], vv[4];
        {
          const float4 q4 = *reinterpret_cast<const float4*>(g.sv.q + (((size_t)t * S + s) * B + b) * A + 4 * lane);
          const float4 v4 = *reinterpret_cast<const float4*>(sa.v + 4 * lane);
          qv[0] = q4.x; qv[1] = q4.y; qv[2] = q4.z; qv[3] = q4.w;
          vv[0] = v4.x; vv[1] = v4.y; vv[2] = v4.z; vv[3] = v4.w;
        }
        for (int a = tid; a < 2 * A; a += kCT) dqa_s[a] = 0.f;          // dqa_s and dva_s are adjacent
        if (step > 0 && s != sp) PBW_WAIT_FLAG(flag(F_X1 + s), (unsigned)n_g1s * (unsigned)step)
        for (int d = tid; d < E; d += kCT) {
          float acc = __ldcs(g.dyc + ((size_t)t * B + b) * (S * E) + s * E + d);
          if (step > 0) {
#pragma unroll
            for (int k = 0; k < kSplits; ++k) acc += __ldcg(q.dxc1 + ((((size_t)((t + 1) & 1) * S + s) * kSplits + k) * NPAD + b) * E + d);
          }
#pragma unroll
          for (int k = 0; k < kSplits; ++k) acc += __ldcg(q.dxc2 + ((((size_t)(t & 1) * kSplits + k) * S + s) * NPAD + b) * E + d);
          dctx_s[d] = acc;
          if (part == 0) __stcs(g.dctx + (((size_t)s * T + t) * B + b) * E + d, acc);
        }
        pb::bar_compute();
        PBW_PH(12)
        // ---- d alpha'_j = carry_j + d ctx . memory_j   (attention.py:395): two rows in flight per warp ----
        float dan = base;                                       // lane i: d alpha' of position r0 + i
        if (r1 > r0) {
          const float4* dc4 = reinterpret_cast<const float4*>(dctx_s);
          float4 dcv[2][2];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) { dcv[hh][0] = dc4[2 * (lane + 32 * hh)]; dcv[hh][1] = dc4[2 * (lane + 32 * hh) + 1]; }
          auto dot_row = [&](const uint4 (&m)[2]) {
            float acc = 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float2 m0 = h2f(m[hh].x), m1 = h2f(m[hh].y), m2 = h2f(m[hh].z), m3 = h2f(m[hh].w);
              acc = fmaf(m0.x, dcv[hh][0].x, acc); acc = fmaf(m0.y, dcv[hh][0].y, acc);
              acc = fmaf(m1.x, dcv[hh][0].z, acc); acc = fmaf(m1.y, dcv[hh][0].w, acc);
              acc = fmaf(m2.x, dcv[hh][1].x, acc); acc = fmaf(m2.y, dcv[hh][1].y, acc);
              acc = fmaf(m3.x, dcv[hh][1].z, acc); acc = fmaf(m3.y, dcv[hh][1].w, acc);
            }
            return warp_sum(acc);
          };
          for (int j = r0; j < rd; j += 4) {
            float d[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) d[r] = j + r < rd ? dot_row(mrow[r]) : 0.f;
            // next four rows
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int jn = j + 4 + r;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh)
                if (jn < rd) mrow[r][hh] = ld_keep_u4(mem16_b + (size_t)jn * E + 8 * (lane + 32 * hh), pol_keep);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (lane == j + r - r0) dan += d[r];
          }
        }
        PBW_PH(13)
        if (p.independent && jme >= len) dan = 0.f;             // alpha'_j was forced to 0 beyond the utterance's length
        // ---- alpha'_j = alpha_j p_j + alpha_{j-1} (1 - p_{j-1}),  p = sigmoid(e)   (attention.py:330-345) ----
        const float dn1 = __shfl_down_sync(0xffffffffu, dan, 1);          // lane i + 1 holds position j + 1 (zero beyond the last one)
        float de = 0.f;
        if (own) {
          const float dnn = jme + 1 < Ts ? dn1 : 0.f;
          g.dalpha[s][(size_t)b * Ts + jme] = dan * pj + dnn * (1.0f - pj);
          de = apj * (dan - dnn) * pj * (1.0f - pj);
        }
        // ---- e_j = v . tanh(q + pm_j): dq, dv, d processed_memory; rows of the block that exist, two in flight ----
        {
          float dq_acc[4] = {0.f, 0.f, 0.f, 0.f}, dv_acc[4] = {0.f, 0.f, 0.f, 0.f};
          const int re = min(r1, len);
          uint2 pmr[2];
          float4 dpr[2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int j = r0 + r;
            pmr[r] = j < re ? ld_keep_u2(pm16_b + (size_t)j * A + 4 * lane, pol_keep) : make_uint2(0u, 0u);
            dpr[r] = j < re ? *reinterpret_cast<const float4*>(dpm_b + (size_t)j * A + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          for (int j = r0; j < re; j += 2) {
            uint2 pmc[2];
            float4 dpc[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) { pmc[r] = pmr[r]; dpc[r] = dpr[r]; }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int jn = j + 2 + r;
              if (jn < re) {
                pmr[r] = ld_keep_u2(pm16_b + (size_t)jn * A + 4 * lane, pol_keep);
                dpr[r] = *reinterpret_cast<const float4*>(dpm_b + (size_t)jn * A + 4 * lane);
              }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int jj = j + r;
              const float dej = __shfl_sync(0xffffffffu, de, (jj - r0) & 31);
              if (jj < re) {
                const float2 p01 = h2f(pmc[r].x), p23 = h2f(pmc[r].y);
                const float pmv[4] = {p01.x, p01.y, p23.x, p23.y};
                const float dpv[4] = {dpc[r].x, dpc[r].y, dpc[r].z, dpc[r].w};
                float o[4];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                  const float uu = lat::fast_tanh(qv[qq] + pmv[qq]);
                  const float dz = dej * vv[qq] * (1.0f - uu * uu);
                  dq_acc[qq] += dz;
                  dv_acc[qq] = fmaf(dej, uu, dv_acc[qq]);
                  o[qq] = dpv[qq] + dz;
                }
                *reinterpret_cast<float4*>(dpm_b + (size_t)jj * A + 4 * lane) = make_float4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          PBW_PH(14)
          if (re > r0) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              atomicAdd(&dqa_s[4 * lane + qq], dq_acc[qq]);
              atomicAdd(&dva_s[4 * lane + qq], dv_acc[qq]);
            }
          }
        }
        pb::bar_compute();
        for (int a = tid; a < A; a += kCT) {
          atomicAdd(g.dq + (((size_t)s * T + t) * B + b) * A + a, dqa_s[a]);
          atomicAdd(g.dv + ((size_t)s * B + b) * A + a, dva_s[a]);
        }
        pb::bar_compute();
        if (tid == 0) pb::signal(flag(F_DQ + s));
        PBW_PH(15)
      }

      PBW_PH(7)
      // ---------------- attention-LSTM cell backward: stream sp, units [j1, j1 + nu1) x all utterances ----------------
      // Everything but Wq^T dq is known before the last attention task of the stream has finished: the split-K partials, the saved
      // gates and the dropout masks are fetched (and the cell's derivative reduced to one linear form in dh and the carry) while
      // this CTA would otherwise wait for dq;  cell e = ju * B + b, two cells per thread at most (16 units x B <= 64 utterances)
      float c_dh[2], c_mc[2], c_al[2], c_gf[2], c_b0[2], c_b1[2], c_b2[2], c_b3[2];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int e = tid + ci * kCT;
        c_dh[ci] = c_mc[ci] = c_al[ci] = c_gf[ci] = c_b0[ci] = c_b1[ci] = c_b2[ci] = c_b3[ci] = 0.f;
        if (e >= nu1 * B) continue;
        const int ju = e / B, b = e - ju * B, j = j1 + ju;
        const size_t idx = (size_t)b * H + j;
        float dh = 0.f;
        if (step > 0) {
#pragma unroll
          for (int k = 0; k < kSplits; ++k) dh += __ldcg(dx1_nxt + (size_t)sp * dx1_str + ((size_t)k * K1 + P + E + j) * NPAD + b);
        }
#pragma unroll
        for (int k = 0; k < kSplits; ++k) dh += __ldcg(dx2_cur + ((size_t)k * K2 + sp * (H + E) + j) * NPAD + b);
        const float* sv = g.sv.gates1 + (((size_t)t * S + sp) * 5 * H + j) * B + b;
        const float gi = __ldcs(sv), gf = __ldcs(sv + gs), gg = __ldcs(sv + 2 * gs), go = __ldcs(sv + 3 * gs), cn = __ldcs(sv + 4 * gs);
        const float cn_prev = t > 0 ? __ldcs(sv - (size_t)S * 5 * gs + 4 * gs) : 0.f;
        float mh = 1.f, mc = 1.f, mc_prev = 1.f;
        if (p.training) {
          const uint8_t* kh = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * sp) * B * H : nullptr;
          const uint8_t* kc = p.lstm_keep ? p.lstm_keep + ((size_t)t * 6 + 2 * sp + 1) * B * H : nullptr;
          mh = keep_mult(kh, idx, p.seed, 4 + 2 * sp, t, (int)idx, p.thresh_att, sc_att);
          mc = keep_mult(kc, idx, p.seed, 5 + 2 * sp, t, (int)idx, p.thresh_att, sc_att);
          if (t > 0) {
            const uint8_t* kcp = p.lstm_keep ? p.lstm_keep + ((size_t)(t - 1) * 6 + 2 * sp + 1) * B * H : nullptr;
            mc_prev = keep_mult(kcp, idx, p.seed, 5 + 2 * sp, t - 1, (int)idx, p.thresh_att, sc_att);
          }
        }
        // bw::lstm_cell_backward as a linear form:  dcn = carry * mc + dh * al;  carry' = dcn * gf;  dgate = {dcn b0, dcn b1, dcn b2, dh b3}
        const float tcn = tanhf(cn);
        c_dh[ci] = dh; c_mc[ci] = mc; c_gf[ci] = gf;
        c_al[ci] = mh * go * (1.0f - tcn * tcn);
        c_b0[ci] = gg * gi * (1.0f - gi);
        c_b1[ci] = (mc_prev * cn_prev) * gf * (1.0f - gf);
        c_b2[ci] = gi * (1.0f - gg * gg);
        c_b3[ci] = mh * tcn * go * (1.0f - go);
      }
      PBW_WAIT_FLAG(flag(F_DQ + sp), (unsigned)(B * parts_sp) * (unsigned)(step + 1))
      PBW_PH(8)
      for (int i = tid; i < A * B; i += kCT) {
        const int b = i / A, a = i - b * A;
        dq_s[a * (NPAD + 1) + b] = __ldcg(g.dq + (((size_t)sp * T + t) * B + b) * A + a);
      }
      pb::bar_compute();
      // u[unit][b] = sum_a Wq[a][unit] dq[b][a]   (attention.py:56 backwards): thread = (utterance, 4 units), one dq word and
      // one 16-byte weight word per 4 multiply-adds
      float* u_s = att_s;                                     // [16][NPAD + 1], the attention scratch is free in this phase
      if (tid < 4 * B) {
        const int b = tid % B, uq = tid / B;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int a = 0; a < A; ++a) {
          const float d = dq_s[a * (NPAD + 1) + b];
          const float4 w = *reinterpret_cast<const float4*>(wq_s + a * 16 + 4 * uq);
          a0 = fmaf(d, w.x, a0); a1 = fmaf(d, w.y, a1); a2 = fmaf(d, w.z, a2); a3 = fmaf(d, w.w, a3);
        }
        u_s[(4 * uq + 0) * (NPAD + 1) + b] = a0; u_s[(4 * uq + 1) * (NPAD + 1) + b] = a1;
        u_s[(4 * uq + 2) * (NPAD + 1) + b] = a2; u_s[(4 * uq + 3) * (NPAD + 1) + b] = a3;
      }
      pb::bar_compute();
      for (int half = 0; half * 8 < nu1; ++half) {
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int e = tid + ci * kCT;
          if (e >= nu1 * B) continue;
          const int ju = e / B, b = e - ju * B;
          if ((ju >> 3) != half) continue;
          const float dh = c_dh[ci] + u_s[ju * (NPAD + 1) + b];
          const float dcn = dc1_r[ci] * c_mc[ci] + dh * c_al[ci];
          dc1_r[ci] = dcn * c_gf[ci];
          float* o = out_s + (size_t)b * 9 + (ju & 7);
          o[0] = dcn * c_b0[ci]; o[(size_t)NPAD * 9] = dcn * c_b1[ci]; o[(size_t)2 * NPAD * 9] = dcn * c_b2[ci]; o[(size_t)3 * NPAD * 9] = dh * c_b3[ci];
        }
        pb::bar_compute();
        store_gates(q.dg1t + (size_t)sp * (G / 64) * NPAD * 128, g.dg1 + ((size_t)sp * T + t) * B * G, j1 + half * 8);
        pb::bar_compute();
      }
      if (tid == 0) pb::signal(flag(F_DG1 + sp));
      if (q.dbg && c == 0 && tid == 0 && step == q.dbg_step) q.dbg[0] = clock64();      // own gate gradients published
      PBW_PH(9)
      // ---------------- window: the attention-LSTM product of frame t is being issued now; advance the decoder-LSTM chain ----------------
      if (t > 0) {
        if (!epilogue2(step + 1)) goto pbw_done;                 // dX2[t-1]
        PBW_PH(1)
        if (own2 && t > 1) {
          PBW_WAIT_FLAG(flag(F_X2), (unsigned)n_g2 * (unsigned)(step + 2))
          PBW_PH(2)
          pointwise2(t - 2, true);                               // dG2[t-2]: needs the hidden rows of dX2[t-1] from every K split
          if (tid == 0) pb::signal(flag(F_DG2));
        }
      }
      PBW_PH(4)

      // ---------------- epilogue of dX1[t] ----------------
      if (has_g1) {
        PBW_WAIT_MBAR(&acc_full[0], (uint32_t)(step & 1))
        if (q.dbg && c == 0 && tid == 0 && step == q.dbg_step) q.dbg[2] = clock64();    // accumulator complete
        PBW_PH(10)
        tc::tc_fence_after();
        {
          const int row0 = m1 * 128;
          const bool ctx_tile = row0 >= P && row0 < P + E;
          float* part_mine = dx1_cur + (size_t)s1 * dx1_str + ((size_t)sig1 * K1 + (size_t)row0) * NPAD;
          float* ctx_mine = q.dxc1 + ((((size_t)(t & 1) * S + s1) * kSplits + sig1) * NPAD) * E + (row0 - P);
          store_acc(acc_addr[0], part_mine, ctx_tile ? ctx_mine : nullptr);
        }
        tc::tc_fence_before();
        pb::bar_compute();
        if (tid == 0) { pb::mbar_arrive(&acc_empty[0]); pb::signal(flag(F_X1 + s1)); }
      }
      PBW_PH(11)
    }
    // d prenet[0] = prenet rows of dX1[0]
    {
      PBW_WAIT_FLAG(flag(F_X1 + sp), (unsigned)n_g1s * (unsigned)T)
      const float* dx1_0 = q.dx1;            // parity of frame 0
      const int npk = P / per, k0 = (c % per) * npk;
      for (int e = tid; e < npk * B; e += kCT) {
        const int kl = e / B, b = e - kl * B, k = k0 + kl;
        float acc = 0.f;
#pragma unroll
        for (int qq = 0; qq < kSplits; ++qq) acc += __ldcg(dx1_0 + (size_t)sp * dx1_str + ((size_t)qq * K1 + k) * NPAD + b);
        g.dpre[(((size_t)sp * T + 0) * B + b) * P + k] = acc;
      }
    }
  pbw_done:;
    if (c == 0 && tid == 0 && p.phase_clocks)
      for (int i = 0; i < 16; ++i) p.phase_clocks[i] = s_ph[i];
#undef PBW_PH
#undef PBW_WAIT_FLAG
#undef PBW_WAIT_MBAR
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

}  // namespace pbw
