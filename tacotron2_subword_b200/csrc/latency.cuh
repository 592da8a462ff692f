// latency.cuh -- role-specialised persistent decoder for the batch-1 latency path (sm_100a).
//
// One cooperative launch, 148 co-resident CTAs, NO grid-wide barriers.  CTAs take roles:
//
//   LSTM CTAs (NL = #SM - 8*S - 8)   own a fixed slice of hidden units of the attention LSTM(s)
//       and of the decoder LSTM.  Their weights are pre-packed (fp32 or fp16) into one contiguous
//       per-CTA stream in consumption order; a producer thread streams it with TMA bulk copies
//       (cp.async.bulk + mbarrier) through a shared-memory ring, while a prefix of the stream
//       stays RESIDENT in shared memory for the whole utterance.  15 consumer warps do row dot
//       products out of shared memory with the activation slice held in registers.
//   attention CTAs (8 per stream)    keep processed_memory, a 64-feature slice of the encoder
//       memory and the alignment state resident in shared memory; fused energy -> sigmoid ->
//       stepwise-monotonic update -> context reduction with warp shuffles.
//   aux CTAs (8)                     keep the mel/gate projection and prenet weights resident and
//       run projection -> stop test -> prenet layer 0 -> layer 1 for the next frame.
//
// Vectors that cross CTAs (h, context, query partials, prenet, mel) travel through global memory
// in an "LL" protocol: each element is one 64-bit word {fp32 value, frame tag}, written with a
// single 8-byte store and polled by the consumers -- no fences, no barriers; one L2 round trip
// per dependency instead of a grid barrier.
//
// Per-frame program of an LSTM CTA (segments ordered by when their inputs become available;
// reference arithmetic: model.py:337-346 attention LSTM, :362-373 decoder LSTM):
//   a  attn-LSTM  W_hh . h1[t-1]          d  attn-LSTM  W_ih[:, :P] . prenet[t]   -> h1[t], q partials
//   b  attn-LSTM  W_ih[:, P:] . ctx[t-1]  e  dec-LSTM   W_ih[:, h cols] . h1[t]   (one step per stream)
//   c  dec-LSTM   W_hh . h2[t-1]          f  dec-LSTM   W_ih[:, ctx cols] . ctx[t] -> h2[t]
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lat {

constexpr int kThreads = 512;
constexpr int kConsumerWarps = 15;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kLLDepth = 8;            // ring depth of every LL vector (frames)
constexpr int kAttnPerStream = 8;      // attention CTAs per stream
constexpr int kAux = 8;                // aux CTAs
constexpr int kMaxU1 = 20, kMaxU2 = 12; // max hidden units per LSTM CTA (attention / decoder LSTM)
constexpr int kSteps = 7;              // a b c d e0 e1 f
constexpr int kMaxSlots = 12;
constexpr long long kWatchdogClocks = 6000000000LL;

// fixed model dims of this path (hparams defaults); other shapes use the generic kernel
constexpr int H = 1024, E = 512, P = 256, A = 128, M = 80;

struct LatStream {
  const float *b_ih, *b_hh;      // attention LSTM biases [4H]
  const float *wq;               // [A, H]
  const float *v;                // [A]
  const float *pre_w0, *pre_w1;  // [P, M], [P, P]
  const float *mem;              // [T_s, E]   (batch 1)
  const float *pm;               // [T_s, A]   processed memory (precomputed)
  const float *pre_tf;           // teacher-forced: hoisted prenet output [T+1, P]
  const float *noise;            // [T, T_s] or null
  const uint8_t *keep0, *keep1;  // prenet keep masks [rows, P] or null
  float *align;                  // [Tcap, T_s]
  const long long* len;          // [1] valid length or null
  int Ts;                        // padded width
};

struct LatParams {
  int S, NL, NL1;                // streams, LSTM CTAs, LSTM CTAs per stream (attention LSTM split)
  int free_running, training, n_steps, Tcap;
  float gate_thr, p_att, p_dec;
  unsigned thresh_pre, thresh_att, thresh_dec;
  unsigned long long seed;
  int wbytes;                    // bytes per packed weight element (4 = fp32, 2 = fp16)
  const unsigned char* packed;   // per-LSTM-CTA weight streams
  const unsigned long long* packed_off;  // [NL+1] byte offsets
  int slot_bytes, n_slots, res_budget;   // TMA ring geometry, bytes available for resident chunks
  int units_per_block;           // hidden units (4-row chunks) moved by one TMA bulk copy
  int poll_sleep_ns;             // back-off between LL polls
  int debug_direct;              // diagnostics: 1 = consumers read streamed chunks straight from global memory,
                                 // 2 = ring on, every streamed row is verified against global memory
  unsigned long long* dbg;       // [0] = mismatch count, then 8 words per record
  LatStream st[2];
  const float *d_b_ih, *d_b_hh;  // decoder LSTM biases [4H]
  const float *proj_w, *proj_b, *gate_w, *gate_b;
  const uint8_t* lstm_keep;      // [T, 6, H] or null
  float *mel, *gate;             // [Tcap, M], [Tcap]
  int *n_frames, *reached_max;
  // LL exchange buffers (64-bit {value, tag}); zeroed by the host before launch
  unsigned long long *ll_h1, *ll_h2, *ll_ctx, *ll_q, *ll_pre, *ll_mel, *ll_l0;
  unsigned* aux_done;
  int* abort_flag;
  long long* phase_clocks;       // [16] diagnostics (LSTM CTA 0)
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-defined time; the producer must not)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ long long lds_acquire_s64(const volatile long long* p) {
  long long v;
  asm volatile("ld.acquire.cta.shared.s64 %0, [%1];" : "=l"(v) : "r"(smem_u32((const void*)p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_release_s64(volatile long long* p, long long v) {
  asm volatile("st.release.cta.shared.s64 [%0], %1;" ::"r"(smem_u32((const void*)p)), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_acquire_u32(const volatile unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)p)) : "memory");
  return v;
}
__device__ __forceinline__ void reds_release_inc(volatile unsigned* p) {
  asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(smem_u32((const void*)p)) : "memory");
}

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

// LL protocol: one 64-bit word = {value bits (low), tag (high)}
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

struct Watch {
  int* abort_flag;
  long long t0;
  unsigned spins;
  unsigned sleep_ns;
  __device__ __forceinline__ void arm() { t0 = clock64(); spins = 0; }
  // returns true when the kernel must bail out
  __device__ __forceinline__ bool expired() {
    if ((++spins & 2047u) != 0u) return false;
    if (*((volatile int*)abort_flag) != 0) return true;
    if (clock64() - t0 > kWatchdogClocks) { atomicExch(abort_flag, 1); return true; }
    return false;
  }
};

// poll one LL word until it carries `tag`; false on abort
__device__ __forceinline__ bool ll_wait(const unsigned long long* p, unsigned tag, float& v, Watch& w) {
  w.arm();
  for (;;) {
    const unsigned long long x = ll_load(p);
    if ((unsigned)(x >> 32) == tag) { v = __uint_as_float((unsigned)x); return true; }
    if (w.sleep_ns) __nanosleep(w.sleep_ns);
    if (w.expired()) return false;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// fast transcendental forms (ex2.approx / rcp.approx): abs error ~1e-6, used on the attention
// energies where T*A evaluations per frame sit on the critical path.
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = exp2f(x * 2.8853900817779268f);  // e^(2x)
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// LSTM CTA
// ---------------------------------------------------------------------------------------------
struct StepPlan {
  int n_chunks;      // chunks (hidden units) in this step
  int kseg;          // columns per row in this step
  int chunk_bytes;   // 4 rows * kseg * wbytes
  int n_res;         // first n_res chunks are resident in shared memory
  int res_off;       // byte offset of the step's resident chunks inside the resident region
  long long src_off; // byte offset of the step's first chunk inside this CTA's packed stream
  int n_blocks;      // TMA blocks (groups of units_per_block streamed chunks) in this step
  int block_base;    // index of the step's first block within a frame's block sequence
};

template <int WB>  // bytes per weight element
struct RowDot;

template <>
struct RowDot<4> {
  template <int NU>  // 16-byte units per lane
  static __device__ __forceinline__ float run(const unsigned char* row, const float* x, int lane) {
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(row + (size_t)(lane + 32 * i) * 16);
      acc0 = fmaf(w.x, x[4 * i + 0], acc0);
      acc1 = fmaf(w.y, x[4 * i + 1], acc1);
      acc0 = fmaf(w.z, x[4 * i + 2], acc0);
      acc1 = fmaf(w.w, x[4 * i + 3], acc1);
    }
    return acc0 + acc1;
  }
  static constexpr int kElemsPerUnit = 4;
};

template <>
struct RowDot<2> {
  template <int NU>
  static __device__ __forceinline__ float run(const unsigned char* row, const float* x, int lane) {
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const uint4 w = *reinterpret_cast<const uint4*>(row + (size_t)(lane + 32 * i) * 16);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
      const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&w.z));
      const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&w.w));
      acc0 = fmaf(a.x, x[8 * i + 0], acc0);
      acc1 = fmaf(a.y, x[8 * i + 1], acc1);
      acc0 = fmaf(b.x, x[8 * i + 2], acc0);
      acc1 = fmaf(b.y, x[8 * i + 3], acc1);
      acc0 = fmaf(c.x, x[8 * i + 4], acc0);
      acc1 = fmaf(c.y, x[8 * i + 5], acc1);
      acc0 = fmaf(d.x, x[8 * i + 6], acc0);
      acc1 = fmaf(d.y, x[8 * i + 7], acc1);
    }
    return acc0 + acc1;
  }
  static constexpr int kElemsPerUnit = 8;
};

struct LstmShared {
  // carved from dynamic shared memory by lstm_cta()
  unsigned char* ring;
  unsigned char* resident;
  float *xh1, *xctx, *xh2, *xpre;           // activation segments [S*H], [S*E], [H], [P]
  float *wq_s;                              // [A][kMaxU1] query-weight slice of this CTA's units
  float *acc1, *acc2, *c1, *c2, *hloc;      // gate accumulators / cell state / fresh h1 of own units
  float *bias1, *bias2;                     // b_ih + b_hh of own units [u][4]
  uint64_t *full, *res_bar;                 // mbarriers (TMA completion); observed in order by the producer only
  volatile unsigned* consumed;              // [slot] rows consumed so far (4 per use) -> frees the slot
  volatile long long* landed;               // chunks whose TMA copy has landed (monotonic, published by the producer)
  StepPlan* plan;                           // [kSteps]
  volatile int* exit_flag;
  const unsigned char* gstream;             // this CTA's packed stream in global memory
  int direct;
  unsigned long long* dbg;
  int units_per_block;
};

// consume one step: every (chunk, gate-row) item is one warp-level dot product of KSEG columns
template <int WB, int KSEG>
__device__ __forceinline__ void consume_step(const LstmShared& sh, const StepPlan& sp, const float* xs, float* acc,
                                             long long frame_block_base, int n_slots, int slot_bytes, int warp,
                                             int lane, Watch& wd, bool& ok) {
  const int G = sh.units_per_block;
  unsigned long long* p_dbg = sh.dbg;
  constexpr int EPU = RowDot<WB>::kElemsPerUnit;
  constexpr int NU = KSEG / (32 * EPU);
  static_assert(NU >= 1, "segment too short");
  float x[NU * EPU];
#pragma unroll
  for (int i = 0; i < NU; ++i)
#pragma unroll
    for (int e = 0; e < EPU; e += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xs + (size_t)(lane + 32 * i) * EPU + e);
      x[i * EPU + e + 0] = v.x; x[i * EPU + e + 1] = v.y; x[i * EPU + e + 2] = v.z; x[i * EPU + e + 3] = v.w;
    }
  const int n_items = sp.n_chunks * 4;
  for (int it = warp; it < n_items; it += kConsumerWarps) {
    const int ci = it >> 2, g = it & 3;
    const unsigned char* base;
    volatile unsigned* release = nullptr;
    if (ci < sp.n_res) {
      base = sh.resident + sp.res_off + (size_t)ci * sp.chunk_bytes;
    } else if (sh.direct == 1) {
      base = sh.gstream + sp.src_off + (size_t)ci * sp.chunk_bytes;
    } else {
      const int rs = ci - sp.n_res;                       // index among the step's streamed chunks
      const long long seq = frame_block_base + sp.block_base + rs / G;
      const int slot = (int)(seq % n_slots);
      wd.arm();
      // Only the producer thread looks at the TMA mbarriers (it sees every use of every slot in order,
      // so phase parity is exact); it republishes completion as a monotonic chunk counter.
      while (lds_acquire_s64(sh.landed) <= seq) {
        if (wd.expired()) { ok = false; break; }
      }
      if (!ok) break;
      base = sh.ring + (size_t)slot * slot_bytes + (size_t)(rs % G) * sp.chunk_bytes;
      release = &sh.consumed[slot];
      if (sh.direct == 2) {
        // verify this row against the packed stream in global memory
        const unsigned char* gsrc = sh.gstream + sp.src_off + (size_t)ci * sp.chunk_bytes + (size_t)g * KSEG * WB;
        const unsigned char* ssrc = base + (size_t)g * KSEG * WB;
        int bad = 0;
        for (int i = lane; i < KSEG * WB / 16; i += 32) {
          const uint4 a = *reinterpret_cast<const uint4*>(ssrc + (size_t)i * 16);
          const uint4 b = *reinterpret_cast<const uint4*>(gsrc + (size_t)i * 16);
          if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) ++bad;
        }
        const unsigned any = __ballot_sync(0xffffffffu, bad != 0);
        if (any) {
          // re-check after a delay: does it become right (early read) or stay wrong (overwritten / wrong data)?
          const long long t0 = clock64();
          while (clock64() - t0 < 200000) {}
          int bad2 = 0, prev_match = 0, next_match = 0;
          const long long step_bytes = sp.chunk_bytes;
          for (int i = lane; i < KSEG * WB / 16; i += 32) {
            const uint4 a = *reinterpret_cast<const uint4*>(ssrc + (size_t)i * 16);
            const uint4 b = *reinterpret_cast<const uint4*>(gsrc + (size_t)i * 16);
            if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) ++bad2;
          }
          const unsigned any2 = __ballot_sync(0xffffffffu, bad2 != 0);
          if (lane == 0) {
            const unsigned long long k = atomicAdd(p_dbg, 1ull);
            if (k < 60) {
              unsigned long long* r = p_dbg + 8 + k * 8;
              r[0] = ((unsigned long long)blockIdx.x << 32) | (unsigned)(seq / 1);
              r[1] = ((unsigned long long)(unsigned)slot << 32) | (unsigned)it;
              r[2] = ((unsigned long long)any << 32) | any2;
              r[3] = (unsigned long long)(*sh.landed);
              r[4] = (unsigned long long)sp.kseg;
              r[5] = (unsigned long long)frame_block_base;
              r[6] = (unsigned long long)sp.block_base;
              r[7] = (unsigned long long)step_bytes;
            }
          }
          (void)prev_match; (void)next_match;
        }
      }
    }
    float v = RowDot<WB>::template run<NU>(base + (size_t)g * KSEG * WB, x, lane);
    v = warp_sum(v);
    if (lane == 0) {
      acc[it] += v;
      if (release) reds_release_inc(release);   // 4 rows per use free the slot
    }
  }
}

// all consumer threads poll an LL vector into shared memory
__device__ __forceinline__ bool poll_vector(const unsigned long long* src, int n, unsigned tag, float* dst, int ctid,
                                            Watch& wd) {
  bool ok = true;
  for (int i = ctid; i < n && ok; i += kConsumerThreads) {
    float v;
    ok = ll_wait(src + i, tag, v, wd);
    dst[i] = v;
  }
  return ok;
}

// philox_keep / philox_normal come from the including translation unit (taco2dec.cu)
__device__ __forceinline__ bool philox_keep_l(unsigned long long seed, int mask_id, int row, int idx, unsigned thresh) {
  return philox_keep(seed, mask_id, row, idx, thresh);
}

template <int WB>
__device__ void lstm_cta(const LatParams& p, int lc, unsigned char* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;
  // ---- unit ownership -------------------------------------------------------------------
  const int s1 = lc / p.NL1;                       // stream whose attention LSTM this CTA serves
  const int i1 = lc - s1 * p.NL1;
  const int u1_0 = (int)((long long)i1 * H / p.NL1), u1_1 = (int)((long long)(i1 + 1) * H / p.NL1);
  const int nu1 = u1_1 - u1_0;
  const int u2_0 = (int)((long long)lc * H / p.NL), u2_1 = (int)((long long)(lc + 1) * H / p.NL);
  const int nu2 = u2_1 - u2_0;

  // ---- shared memory carve-up -----------------------------------------------------------
  LstmShared sh;
  size_t off = 0;
  auto take = [&](size_t bytes) { unsigned char* q = smem + off; off += (bytes + 127) & ~(size_t)127; return q; };
  sh.ring = take((size_t)p.n_slots * p.slot_bytes);
  sh.resident = take((size_t)p.res_budget);
  sh.xh1 = (float*)take(sizeof(float) * 2 * H);
  sh.xctx = (float*)take(sizeof(float) * 2 * E);
  sh.xh2 = (float*)take(sizeof(float) * H);
  sh.xpre = (float*)take(sizeof(float) * P);
  sh.wq_s = (float*)take(sizeof(float) * A * kMaxU1);
  sh.acc1 = (float*)take(sizeof(float) * kMaxU1 * 4);
  sh.acc2 = (float*)take(sizeof(float) * kMaxU2 * 4);
  sh.c1 = (float*)take(sizeof(float) * kMaxU1);
  sh.c2 = (float*)take(sizeof(float) * kMaxU2);
  sh.hloc = (float*)take(sizeof(float) * kMaxU1);
  sh.bias1 = (float*)take(sizeof(float) * kMaxU1 * 4);
  sh.bias2 = (float*)take(sizeof(float) * kMaxU2 * 4);
  sh.full = (uint64_t*)take(sizeof(uint64_t) * kMaxSlots);
  sh.consumed = (volatile unsigned*)take(sizeof(unsigned) * kMaxSlots);
  sh.res_bar = (uint64_t*)take(sizeof(uint64_t));
  sh.plan = (StepPlan*)take(sizeof(StepPlan) * kSteps);
  sh.exit_flag = (volatile int*)take(sizeof(int));
  sh.landed = (volatile long long*)take(sizeof(long long));

  // ---- step plan (thread 0) ---------------------------------------------------------------
  if (tid == 0) {
    // steps:            a     b     c     d     e0    e1            f
    const int nch[kSteps] = {nu1, nu1, nu2, nu1, nu2, S == 2 ? nu2 : 0, nu2};
    const int ks[kSteps] = {H, E, H, P, H, H, S * E};
    long long src = 0;
    for (int s = 0; s < kSteps; ++s) {
      StepPlan& sp = sh.plan[s];
      sp.n_chunks = nch[s]; sp.kseg = ks[s]; sp.chunk_bytes = 4 * ks[s] * WB; sp.n_res = 0; sp.res_off = 0;
      sp.src_off = src;
      src += (long long)nch[s] * sp.chunk_bytes;
    }
    // residency: critical-path steps first (d, f), then e1, e0, c, b, a
    const int prio[kSteps] = {3, 6, 5, 4, 2, 1, 0};
    int left = p.res_budget, roff = 0;
    for (int k = 0; k < kSteps; ++k) {
      StepPlan& sp = sh.plan[prio[k]];
      int n = sp.chunk_bytes > 0 ? left / sp.chunk_bytes : 0;
      if (n > sp.n_chunks) n = sp.n_chunks;
      sp.n_res = n; sp.res_off = roff;
      roff += n * sp.chunk_bytes; left -= n * sp.chunk_bytes;
    }
    int bb = 0;
    for (int s = 0; s < kSteps; ++s) {
      StepPlan& sp = sh.plan[s];
      sp.n_blocks = (sp.n_chunks - sp.n_res + p.units_per_block - 1) / p.units_per_block;
      sp.block_base = bb;
      bb += sp.n_blocks;
    }
    for (int i = 0; i < p.n_slots; ++i) { mbar_init(&sh.full[i], 1); sh.consumed[i] = 0u; }
    mbar_init(sh.res_bar, 1);
    *sh.exit_flag = 0;
    *sh.landed = 0;
    fence_barrier_init();
  }
  // zero state, load biases and the query-weight slice
  for (int i = tid; i < 2 * H; i += kThreads) sh.xh1[i] = 0.f;
  for (int i = tid; i < 2 * E; i += kThreads) sh.xctx[i] = 0.f;
  for (int i = tid; i < H; i += kThreads) sh.xh2[i] = 0.f;
  for (int i = tid; i < P; i += kThreads) sh.xpre[i] = 0.f;
  for (int i = tid; i < kMaxU1 * 4; i += kThreads) {
    const int u = i >> 2, g = i & 3;
    sh.acc1[i] = 0.f;
    sh.bias1[i] = u < nu1 ? p.st[s1].b_ih[g * H + u1_0 + u] + p.st[s1].b_hh[g * H + u1_0 + u] : 0.f;
  }
  for (int i = tid; i < kMaxU2 * 4; i += kThreads) {
    const int u = i >> 2, g = i & 3;
    sh.acc2[i] = 0.f;
    sh.bias2[i] = u < nu2 ? p.d_b_ih[g * H + u2_0 + u] + p.d_b_hh[g * H + u2_0 + u] : 0.f;
  }
  for (int i = tid; i < kMaxU1; i += kThreads) { sh.c1[i] = 0.f; sh.hloc[i] = 0.f; }
  for (int i = tid; i < kMaxU2; i += kThreads) sh.c2[i] = 0.f;
  for (int i = tid; i < A * kMaxU1; i += kThreads) {
    const int a = i / kMaxU1, u = i - a * kMaxU1;
    sh.wq_s[i] = u < nu1 ? p.st[s1].wq[(size_t)a * H + u1_0 + u] : 0.f;
  }
  __syncthreads();

  const unsigned char* my_stream = p.packed + p.packed_off[lc];
  sh.gstream = my_stream;
  sh.direct = p.debug_direct;
  sh.units_per_block = p.units_per_block;
  sh.dbg = p.dbg;
  int blocks_per_frame = 0;
  for (int s = 0; s < kSteps; ++s) blocks_per_frame += sh.plan[s].n_blocks;
  const int n_steps = p.n_steps;

  // ======================= producer warp: TMA weight streaming =============================
  if (warp == kConsumerWarps) {
    if (lane == 0) {
      Watch wd{p.abort_flag, 0, 0, (unsigned)p.poll_sleep_ns};
      // resident prefix: loaded once, lives for the whole utterance
      unsigned res_total = 0;
      for (int s = 0; s < kSteps; ++s) res_total += (unsigned)(sh.plan[s].n_res * sh.plan[s].chunk_bytes);
      if (res_total) {
        mbar_expect_tx(sh.res_bar, res_total);
        for (int s = 0; s < kSteps; ++s) {
          const StepPlan& sp = sh.plan[s];
          for (int c = 0; c < sp.n_res; ++c)
            tma_load_1d(sh.resident + sp.res_off + (size_t)c * sp.chunk_bytes,
                        my_stream + sp.src_off + (size_t)c * sp.chunk_bytes, (unsigned)sp.chunk_bytes, sh.res_bar);
        }
      } else {
        mbar_arrive(sh.res_bar);
      }
      // event loop: issue the next block when its slot is free, confirm landings strictly in order
      const int G = p.units_per_block;
      const long long total = p.debug_direct == 1 ? 0 : (long long)n_steps * blocks_per_frame;
      long long issue = 0, land = 0;
      int it_s = 0, it_b = 0;            // (step, block) of the next block to issue
      unsigned exp_rows[kMaxSlots];      // rows the consumers must have retired before a slot may be refilled
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i) exp_rows[i] = 0u;
      auto advance = [&]() {             // skip steps without streamed blocks (wraps per frame)
        while (it_b >= sh.plan[it_s].n_blocks) { it_b = 0; it_s = (it_s + 1) % kSteps; }
      };
      wd.arm();
      bool live = total > 0;
      if (live) advance();
      while (live && land < total) {
        bool progressed = false;
        while (issue < total && issue - land < p.n_slots) {
          const int slot = (int)(issue % p.n_slots);
          if (lds_acquire_u32(&sh.consumed[slot]) < exp_rows[slot]) break;
          const StepPlan& sp = sh.plan[it_s];
          const int first = sp.n_res + it_b * G;
          const int units = min(G, sp.n_chunks - first);
          const unsigned bytes = (unsigned)(units * sp.chunk_bytes);
          mbar_expect_tx(&sh.full[slot], bytes);
          tma_load_1d(sh.ring + (size_t)slot * p.slot_bytes, my_stream + sp.src_off + (size_t)first * sp.chunk_bytes, bytes,
                      &sh.full[slot]);
          exp_rows[slot] += 4u * (unsigned)units;
          ++issue;
          ++it_b;
          advance();
          progressed = true;
        }
        {
          // probe up to 4 in-flight blocks with independent (overlapping) test_waits, accept the landed prefix
          bool f[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const long long q = land + k;
            f[k] = q < issue && mbar_test_wait(&sh.full[(int)(q % p.n_slots)], (unsigned)((q / p.n_slots) & 1));
          }
          int n = 0;
          if (f[0]) { n = 1; if (f[1]) { n = 2; if (f[2]) { n = 3; if (f[3]) n = 4; } } }
          if (n) { land += n; sts_release_s64(sh.landed, land); progressed = true; }
        }
        if (progressed) { wd.arm(); continue; }
        if (*sh.exit_flag) {
          // consumers are gone: stop issuing, but every issued copy must land before the CTA exits
          wd.arm();
          while (land < issue) {
            const int sl = (int)(land % p.n_slots);
            const unsigned par = (unsigned)((land / p.n_slots) & 1);
            if (mbar_try_wait(&sh.full[sl], par)) { ++land; continue; }
            if (wd.expired()) break;
          }
          break;
        }
        if (wd.expired()) live = false;
      }
      wd.arm();
      while (!mbar_try_wait(sh.res_bar, 0)) {
        if (wd.expired()) break;
      }
    }
    return;
  }

  // ======================= consumer warps ===================================================
  Watch wd{p.abort_flag, 0, 0, (unsigned)p.poll_sleep_ns};
  bool ok = true;
  const int ctid = tid;  // 0..479
  const float sc_att = 1.0f / (1.0f - p.p_att), sc_dec = 1.0f / (1.0f - p.p_dec);
  // wait for the resident prefix
  wd.arm();
  while (!mbar_try_wait(sh.res_bar, 0)) {
    if (wd.expired()) { ok = false; break; }
  }
  long long ph_t = clock64();
  long long ph[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ph[i] = 0;
#define LPH(slot)                                         \
  if (lc == 0 && tid == 0) {                              \
    const long long n_ = clock64();                       \
    ph[slot] += n_ - ph_t;                                \
    ph_t = n_;                                            \
  }

  const StepPlan* pl = sh.plan;
  for (int t = 0; t < n_steps && ok; ++t) {
    const long long fbase = (long long)t * blocks_per_frame;
    const unsigned tag_prev = (unsigned)t;        // values produced during frame t-1
    const unsigned tag_cur = (unsigned)t + 1u;    // values produced during frame t
    const int rb = t % kLLDepth, rb_prev = (t + kLLDepth - 1) % kLLDepth;

    // a: W_hh . h1[t-1]  (own stream)            b: W_ih[:, P:] . ctx[t-1]
    consume_step<WB, H>(sh, pl[0], sh.xh1 + s1 * H, sh.acc1, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    consume_step<WB, E>(sh, pl[1], sh.xctx + s1 * E, sh.acc1, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    LPH(0)
    // c: W_hh(dec) . h2[t-1]
    if (t > 0) ok = ok && poll_vector(p.ll_h2 + (size_t)rb_prev * H, H, tag_prev, sh.xh2, ctid, wd);
    consumer_sync();
    LPH(1)
    consume_step<WB, H>(sh, pl[2], sh.xh2, sh.acc2, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    LPH(2)
    // d: W_ih[:, :P] . prenet[t]
    if (p.free_running) {
      // stop word first: the aux CTAs publish it with the prenet of frame t
      int stop = 0;
      if (ctid == 0) {
        float sv = 0.f;
        ok = ll_wait(p.ll_pre + ((size_t)rb * 2 + 0) * (P + 8) + P, tag_cur, sv, wd) && ok;
        stop = sv != 0.f;
        if (stop || !ok) *sh.exit_flag = 1;
      }
      consumer_sync();
      if (*sh.exit_flag) break;
      ok = ok && poll_vector(p.ll_pre + ((size_t)rb * 2 + s1) * (P + 8), P, tag_cur, sh.xpre, ctid, wd);
    } else {
      for (int i = ctid; i < P; i += kConsumerThreads) sh.xpre[i] = __ldg(p.st[s1].pre_tf + (size_t)t * P + i);
    }
    consumer_sync();
    LPH(3)
    consume_step<WB, P>(sh, pl[3], sh.xpre, sh.acc1, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    consumer_sync();
    // attention-LSTM pointwise (gate order i,f,g,o), dropout on h and c when training
    if (ctid < nu1) {
      const int u = ctid, j = u1_0 + u;
      const float ig = sigmoid_acc(sh.acc1[u * 4 + 0] + sh.bias1[u * 4 + 0]);
      const float fg = sigmoid_acc(sh.acc1[u * 4 + 1] + sh.bias1[u * 4 + 1]);
      const float gg = tanhf(sh.acc1[u * 4 + 2] + sh.bias1[u * 4 + 2]);
      const float og = sigmoid_acc(sh.acc1[u * 4 + 3] + sh.bias1[u * 4 + 3]);
      float cn = fg * sh.c1[u] + ig * gg;
      float hn = og * tanhf(cn);
      if (p.training) {
        const bool kh = p.lstm_keep ? p.lstm_keep[((size_t)t * 6 + 2 * s1) * H + j] != 0
                                    : philox_keep_l(p.seed, 4 + 2 * s1, t, j, p.thresh_att);
        const bool kc = p.lstm_keep ? p.lstm_keep[((size_t)t * 6 + 2 * s1 + 1) * H + j] != 0
                                    : philox_keep_l(p.seed, 5 + 2 * s1, t, j, p.thresh_att);
        hn = kh ? hn * sc_att : 0.f;
        cn = kc ? cn * sc_att : 0.f;
      }
      sh.c1[u] = cn;
      sh.hloc[u] = hn;
      ll_store(p.ll_h1 + ((size_t)rb * 2 + s1) * H + j, hn, tag_cur);
    }
    consumer_sync();
    if (ctid >= A && ctid < A + kMaxU1 * 4) sh.acc1[ctid - A] = 0.f;   // next write is after two more syncs
    // query partials: q_part[a] = sum_u Wq[a, u] h1[u]  (attention.py:56, 368), one row of the reduction tree
    if (ctid < A) {
      float qv = 0.f;
      for (int u = 0; u < nu1; ++u) qv = fmaf(sh.wq_s[ctid * kMaxU1 + u], sh.hloc[u], qv);
      ll_store(p.ll_q + (((size_t)rb * 2 + s1) * p.NL1 + i1) * A + ctid, qv, tag_cur);
    }
    LPH(4)
    // e: W_ih(dec)[:, h cols] . h1[t]   (all streams)
    ok = ok && poll_vector(p.ll_h1 + (size_t)rb * 2 * H, S * H, tag_cur, sh.xh1, ctid, wd);
    consumer_sync();
    LPH(5)
    consume_step<WB, H>(sh, pl[4], sh.xh1, sh.acc2, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    if (S == 2) consume_step<WB, H>(sh, pl[5], sh.xh1 + H, sh.acc2, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    LPH(6)
    // f: W_ih(dec)[:, ctx cols] . ctx[t]
    ok = ok && poll_vector(p.ll_ctx + (size_t)rb * 2 * E, S * E, tag_cur, sh.xctx, ctid, wd);
    consumer_sync();
    LPH(7)
    if (S == 2) consume_step<WB, 2 * E>(sh, pl[6], sh.xctx, sh.acc2, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    else        consume_step<WB, E>(sh, pl[6], sh.xctx, sh.acc2, fbase, p.n_slots, p.slot_bytes, warp, lane, wd, ok);
    consumer_sync();
    // flow control (teacher-forced only): the aux CTAs are not in the dependency loop, so do not
    // overwrite LL slot t % depth before they have consumed frame t - depth
    if (!p.free_running && t >= kLLDepth && ctid == 0) {
      const unsigned need = (unsigned)kAux * (unsigned)(t - kLLDepth + 1);
      wd.arm();
      for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.aux_done) : "memory");
        if (v >= need) break;
        if (wd.expired()) { ok = false; break; }
      }
    }
    if (ctid < nu2) {
      const int u = ctid, j = u2_0 + u;
      const float ig = sigmoid_acc(sh.acc2[u * 4 + 0] + sh.bias2[u * 4 + 0]);
      const float fg = sigmoid_acc(sh.acc2[u * 4 + 1] + sh.bias2[u * 4 + 1]);
      const float gg = tanhf(sh.acc2[u * 4 + 2] + sh.bias2[u * 4 + 2]);
      const float og = sigmoid_acc(sh.acc2[u * 4 + 3] + sh.bias2[u * 4 + 3]);
      float cn = fg * sh.c2[u] + ig * gg;
      float hn = og * tanhf(cn);
      if (p.training) {
        const bool kh = p.lstm_keep ? p.lstm_keep[((size_t)t * 6 + 4) * H + j] != 0
                                    : philox_keep_l(p.seed, 8, t, j, p.thresh_dec);
        const bool kc = p.lstm_keep ? p.lstm_keep[((size_t)t * 6 + 5) * H + j] != 0
                                    : philox_keep_l(p.seed, 9, t, j, p.thresh_dec);
        hn = kh ? hn * sc_dec : 0.f;
        cn = kc ? cn * sc_dec : 0.f;
      }
      sh.c2[u] = cn;
      ll_store(p.ll_h2 + (size_t)rb * H + j, hn, tag_cur);
    }
    consumer_sync();
    if (ctid < kMaxU2 * 4) sh.acc2[ctid] = 0.f;
    LPH(8)
    // a CTA-uniform view of `ok` (a warp may have hit the watchdog)
    if (!ok) *sh.exit_flag = 1;
    consumer_sync();
    if (*sh.exit_flag) break;
  }
  if (tid == 0) *sh.exit_flag = 1;  // releases the producer if it is parked on an empty slot
  if (lc == 0 && tid == 0)
    for (int i = 0; i < 16; ++i) p.phase_clocks[i] = ph[i];
#undef LPH
}

// ---------------------------------------------------------------------------------------------
// attention CTA: stream s, context feature slice g  (attention.py:330-398)
// ---------------------------------------------------------------------------------------------
__device__ void attention_cta(const LatParams& p, int s, int g, unsigned char* smem_raw) {
  const LatStream& sp = p.st[s];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ts = sp.Ts;
  const int Teff = sp.len ? min((int)sp.len[0], Ts) : Ts;
  constexpr int FS = E / kAttnPerStream;  // 64 features per CTA
  float* sm = reinterpret_cast<float*>(smem_raw);
  float* pm_s = sm;                         // [Ts][A]
  float* mem_s = pm_s + (size_t)Ts * A;     // [Ts][FS]
  float* q_s = mem_s + (size_t)Ts * FS;     // [4][A] partial sums, then q in row 0
  float* v_s = q_s + 4 * A;                 // [A]
  float* red_s = v_s + A;                   // [8][FS]
  float* al_s = red_s + 8 * FS;             // [Ts] alignment state
  float* pr_s = al_s + Ts;                  // [Ts] probabilities
  float* an_s = pr_s + Ts;                  // [Ts] new alignment
  __shared__ int s_stop;

  for (int i = tid; i < Ts * A; i += kThreads) pm_s[i] = sp.pm[i];
  for (int i = tid; i < Ts * FS; i += kThreads) {
    const int j = i / FS, f = i - j * FS;
    mem_s[i] = sp.mem[(size_t)j * E + g * FS + f];
  }
  for (int i = tid; i < A; i += kThreads) v_s[i] = sp.v[i];
  for (int i = tid; i < Ts; i += kThreads) al_s[i] = i == 0 ? 1.0f : 0.0f;   // attention.py:324-328
  if (tid == 0) s_stop = 0;
  __syncthreads();

  Watch wd{p.abort_flag, 0, 0, (unsigned)p.poll_sleep_ns};
  const int NL1 = p.NL1;
  for (int t = 0; t < p.n_steps; ++t) {
    const unsigned tag = (unsigned)t + 1u;
    const int rb = t % kLLDepth;
    if (p.free_running) {
      if (tid == 0) {
        float sv = 0.f;
        const bool ok = ll_wait(p.ll_pre + ((size_t)rb * 2 + 0) * (P + 8) + P, tag, sv, wd);
        if (!ok || sv != 0.f) s_stop = 1;
      }
      __syncthreads();
      if (s_stop) break;
    }
    // ---- q = sum over LSTM CTAs of their partials ---------------------------------------
    {
      const int part = tid >> 7, a = tid & (A - 1);  // 4 parts x 128
      const unsigned long long* src = p.ll_q + ((size_t)rb * 2 + s) * NL1 * A;
      float acc = 0.f;
      bool ok = true;
      for (int c = part; c < NL1 && ok; c += 4) {
        float v;
        ok = ll_wait(src + (size_t)c * A + a, tag, v, wd);
        acc += v;
      }
      if (!ok) s_stop = 1;
      q_s[part * A + a] = acc;
    }
    __syncthreads();
    if (s_stop) break;
    if (tid < A) q_s[tid] = (q_s[tid] + q_s[A + tid]) + (q_s[2 * A + tid] + q_s[3 * A + tid]);
    __syncthreads();
    // ---- energies e_j = v . tanh(q + pm_j), p_j = sigmoid(e_j [+ 2 N(0,1)]) ---------------
    {
      const float q0 = q_s[lane], q1 = q_s[lane + 32], q2 = q_s[lane + 64], q3 = q_s[lane + 96];
      const float v0 = v_s[lane], v1 = v_s[lane + 32], v2 = v_s[lane + 64], v3 = v_s[lane + 96];
      for (int j = warp; j < Teff; j += kThreads / 32) {
        const float* r = pm_s + (size_t)j * A;
        float e = v0 * fast_tanh(q0 + r[lane]) + v1 * fast_tanh(q1 + r[lane + 32]) +
                  v2 * fast_tanh(q2 + r[lane + 64]) + v3 * fast_tanh(q3 + r[lane + 96]);
        e = warp_sum(e);
        if (lane == 0) {
          if (p.training) {
            const float nz = sp.noise ? sp.noise[(size_t)t * Ts + j] : philox_normal(p.seed, 10 + s, t, j);
            e += 2.0f * nz;
          }
          pr_s[j] = sigmoid_acc(e);
        }
      }
      for (int j = Teff + tid; j < Ts; j += kThreads) pr_s[j] = 0.f;   // sigmoid(-inf), attention.py:388-391
    }
    __syncthreads();
    // ---- alpha'_j = alpha_j p_j + alpha_{j-1} (1 - p_{j-1}) -------------------------------
    for (int j = tid; j < Ts; j += kThreads) {
      float a = al_s[j] * pr_s[j];
      if (j > 0) a += al_s[j - 1] * (1.0f - pr_s[j - 1]);
      if (p.free_running && j >= Teff) a = 0.f;
      an_s[j] = a;
      if (g == 0) sp.align[(size_t)t * Ts + j] = a;
    }
    __syncthreads();
    // ---- context slice: 8 position groups x 64 features ----------------------------------
    {
      const int f = tid & (FS - 1), jg = tid >> 6;
      float acc = 0.f;
      for (int j = jg; j < Ts; j += 8) acc = fmaf(an_s[j], mem_s[(size_t)j * FS + f], acc);
      red_s[jg * FS + f] = acc;
    }
    for (int j = tid; j < Ts; j += kThreads) al_s[j] = an_s[j];
    __syncthreads();
    if (tid < FS) {
      float c = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) c += red_s[k * FS + tid];
      ll_store(p.ll_ctx + ((size_t)rb * 2 + s) * E + g * FS + tid, c, tag);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// aux CTA x: mel/gate projection rows, stop test, prenet for the next frame
// (model.py:382-388, 480-485, 13-24)
// ---------------------------------------------------------------------------------------------
__device__ void aux_cta(const LatParams& p, int x, unsigned char* smem_raw) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;
  const int KD = H + S * E;
  constexpr int RPX = (M + 1 + kAux - 1) / kAux;      // projection rows per aux CTA (11)
  constexpr int PR = P / kAux;                        // prenet rows per aux CTA per stream (32)
  const int r0 = x * RPX, r1 = min(M + 1, r0 + RPX), nr = r1 - r0;
  float* sm = reinterpret_cast<float*>(smem_raw);
  float* wp_s = sm;                                   // [RPX][KD]
  float* w0_s = wp_s + (size_t)RPX * (H + 2 * E);     // [S][PR][M]
  float* w1_s = w0_s + 2 * PR * M;                    // [S][PR][P]
  float* y_s = w1_s + 2 * PR * P;                     // [KD]
  float* mel_s = y_s + (H + 2 * E);                   // [M + 16]
  float* l0_s = mel_s + M + 16;                       // [S][P]
  __shared__ int s_stop;

  for (int i = tid; i < nr * KD; i += kThreads) {
    const int r = i / KD, k = i - r * KD;
    const int row = r0 + r;
    wp_s[(size_t)r * KD + k] = row < M ? p.proj_w[(size_t)row * KD + k] : p.gate_w[k];
  }
  for (int i = tid; i < S * PR * M; i += kThreads) {
    const int s = i / (PR * M), r = (i - s * PR * M) / M, k = i - s * PR * M - r * M;
    w0_s[i] = p.st[s].pre_w0[(size_t)(x * PR + r) * M + k];
  }
  for (int i = tid; i < S * PR * P; i += kThreads) {
    const int s = i / (PR * P), r = (i - s * PR * P) / P, k = i - s * PR * P - r * P;
    w1_s[i] = p.st[s].pre_w1[(size_t)(x * PR + r) * P + k];
  }
  if (tid == 0) s_stop = 0;
  __syncthreads();

  Watch wd{p.abort_flag, 0, 0, (unsigned)p.poll_sleep_ns};
  // prenet of frame 0 = prenet(go-frame of zeros) = zeros (model.py:444-450); stop word = 0
  if (p.free_running) {
    for (int i = tid; i < S * PR; i += kThreads) {
      const int s = i / PR, r = i - s * PR;
      ll_store(p.ll_pre + ((size_t)0 * 2 + s) * (P + 8) + x * PR + r, 0.f, 1u);
    }
    if (x == 0 && tid == 0) ll_store(p.ll_pre + (size_t)P, 0.f, 1u);
  }

  for (int t = 0; t < p.n_steps; ++t) {
    const unsigned tag = (unsigned)t + 1u;
    const int rb = t % kLLDepth, rbn = (t + 1) % kLLDepth;
    // ---- y = [h2_t | ctx_t | ctx_bert_t] ---------------------------------------------------
    bool ok = true;
    for (int i = tid; i < H && ok; i += kThreads) { float v; ok = ll_wait(p.ll_h2 + (size_t)rb * H + i, tag, v, wd); y_s[i] = v; }
    for (int i = tid; i < S * E && ok; i += kThreads) { float v; ok = ll_wait(p.ll_ctx + (size_t)rb * 2 * E + i, tag, v, wd); y_s[H + i] = v; }
    if (!ok) s_stop = 1;
    __syncthreads();
    if (s_stop) break;
    // ---- projection rows (one warp per row) ------------------------------------------------
    if (warp < nr) {
      const float* w = wp_s + (size_t)warp * KD;
      float acc = 0.f;
      for (int k = lane * 4; k < KD; k += 128) {
        const float4 a = *reinterpret_cast<const float4*>(w + k);
        const float4 b = *reinterpret_cast<const float4*>(y_s + k);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const int row = r0 + warp;
        if (row < M) {
          const float v = acc + p.proj_b[row];
          p.mel[(size_t)t * M + row] = v;
          if (p.free_running) ll_store(p.ll_mel + (size_t)rb * (M + 16) + row, v, tag);
        } else {
          const float gv = acc + p.gate_b[0];
          p.gate[t] = gv;
          if (p.free_running) {
            int stop = 0;
            if (sigmoid_acc(gv) > p.gate_thr) { stop = 1; p.n_frames[0] = t + 1; }            // model.py:480-481
            else if (t + 1 == p.n_steps) { stop = 1; p.n_frames[0] = t + 1; p.reached_max[0] = 1; }  // :482-485
            ll_store(p.ll_mel + (size_t)rb * (M + 16) + M, stop ? 1.f : 0.f, tag);
          }
        }
      }
    }
    if (!p.free_running) {
      // nobody waits for the projection when teacher forcing: report progress for LL flow control
      __syncthreads();
      if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.aux_done), "r"(1u) : "memory");
      continue;
    }
    // ---- mel_t (+ stop word) from all aux CTAs ----------------------------------------------
    ok = true;
    for (int i = tid; i < M + 1 && ok; i += kThreads) { float v; ok = ll_wait(p.ll_mel + (size_t)rb * (M + 16) + i, tag, v, wd); mel_s[i] = v; }
    if (!ok) s_stop = 1;
    __syncthreads();
    if (s_stop) break;
    const bool stop = mel_s[M] != 0.f;
    if (x == 0 && tid == 0) ll_store(p.ll_pre + ((size_t)rbn * 2 + 0) * (P + 8) + P, stop ? 1.f : 0.f, tag + 1u);
    if (stop) break;
    // ---- prenet layer 0 rows of this CTA: relu(W0 mel) * keep * 2 ----------------------------
    for (int it = warp; it < S * PR; it += kThreads / 32) {
      const int s = it / PR, r = it - s * PR, row = x * PR + r;
      const float* w = w0_s + ((size_t)s * PR + r) * M;
      float acc = 0.f;
      for (int k = lane; k < M; k += 32) acc = fmaf(w[k], mel_s[k], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        const uint8_t* keep = p.st[s].keep0;
        const bool kp = keep ? keep[(size_t)(t + 1) * P + row] != 0 : philox_keep_l(p.seed, s * 2 + 0, t + 1, row, p.thresh_pre);
        ll_store(p.ll_l0 + ((size_t)rbn * 2 + s) * P + row, kp ? fmaxf(acc, 0.f) * 2.0f : 0.f, tag + 1u);
      }
    }
    // ---- prenet layer 1 -------------------------------------------------------------------
    ok = true;
    for (int i = tid; i < S * P && ok; i += kThreads) { float v; ok = ll_wait(p.ll_l0 + (size_t)rbn * 2 * P + i, tag + 1u, v, wd); l0_s[i] = v; }
    if (!ok) s_stop = 1;
    __syncthreads();
    if (s_stop) break;
    for (int it = warp; it < S * PR; it += kThreads / 32) {
      const int s = it / PR, r = it - s * PR, row = x * PR + r;
      const float* w = w1_s + ((size_t)s * PR + r) * P;
      const float* in = l0_s + (size_t)s * P;
      float acc = 0.f;
      for (int k = lane * 4; k < P; k += 128) {
        const float4 a = *reinterpret_cast<const float4*>(w + k);
        const float4 b = *reinterpret_cast<const float4*>(in + k);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const uint8_t* keep = p.st[s].keep1;
        const bool kp = keep ? keep[(size_t)(t + 1) * P + row] != 0 : philox_keep_l(p.seed, s * 2 + 1, t + 1, row, p.thresh_pre);
        ll_store(p.ll_pre + ((size_t)rbn * 2 + s) * (P + 8) + row, kp ? fmaxf(acc, 0.f) * 2.0f : 0.f, tag + 1u);
      }
    }
    __syncthreads();
  }
}

template <int WB>
__global__ void __launch_bounds__(kThreads, 1) decoder_latency(const __grid_constant__ LatParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  const int b = blockIdx.x;
  if (b < p.NL) {
    lstm_cta<WB>(p, b, dyn_smem);
  } else if (b < p.NL + p.S * kAttnPerStream) {
    const int k = b - p.NL;
    attention_cta(p, k / kAttnPerStream, k % kAttnPerStream, dyn_smem);
  } else if (b < p.NL + p.S * kAttnPerStream + kAux) {
    aux_cta(p, b - p.NL - p.S * kAttnPerStream, dyn_smem);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: per LSTM CTA, chunks in consumption order; chunk = 4 gate rows x segment columns
// ---------------------------------------------------------------------------------------------
struct PackSrc {
  const float* w_ih[2];  // attention LSTM [4H, P+E] per stream
  const float* w_hh[2];  // [4H, H]
  const float* d_w_ih;   // decoder LSTM [4H, S*(H+E)]
  const float* d_w_hh;   // [4H, H]
};

template <typename OutT>
__global__ void pack_weights_kernel(PackSrc src, int S, int NL, int NL1, const unsigned long long* packed_off,
                                    unsigned char* packed) {
  const int lc = blockIdx.x;
  const int s1 = lc / NL1, i1 = lc - s1 * NL1;
  const int u1_0 = (int)((long long)i1 * H / NL1), nu1 = (int)((long long)(i1 + 1) * H / NL1) - u1_0;
  const int u2_0 = (int)((long long)lc * H / NL), nu2 = (int)((long long)(lc + 1) * H / NL) - u2_0;
  OutT* out = reinterpret_cast<OutT*>(packed + packed_off[lc]);
  const int nch[kSteps] = {nu1, nu1, nu2, nu1, nu2, S == 2 ? nu2 : 0, nu2};
  const int ks[kSteps] = {H, E, H, P, H, H, S * E};
  const int KX1 = P + E, KX2 = S * (H + E);
  size_t o = 0;
  for (int st = 0; st < kSteps; ++st) {
    const int K = ks[st];
    const size_t n = (size_t)nch[st] * 4 * K;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
      const int k = (int)(i % K);
      const int g = (int)((i / K) & 3);
      const int u = (int)(i / ((size_t)4 * K));
      float v;
      switch (st) {
        case 0: v = src.w_hh[s1][(size_t)(g * H + u1_0 + u) * H + k]; break;
        case 1: v = src.w_ih[s1][(size_t)(g * H + u1_0 + u) * KX1 + P + k]; break;
        case 2: v = src.d_w_hh[(size_t)(g * H + u2_0 + u) * H + k]; break;
        case 3: v = src.w_ih[s1][(size_t)(g * H + u1_0 + u) * KX1 + k]; break;
        case 4: v = src.d_w_ih[(size_t)(g * H + u2_0 + u) * KX2 + k]; break;                 // h (stream 0)
        case 5: v = src.d_w_ih[(size_t)(g * H + u2_0 + u) * KX2 + (H + E) + k]; break;       // h_bert
        default: {                                                                           // ctx | ctx_bert
          const int s = k / E, kk = k - s * E;
          v = src.d_w_ih[(size_t)(g * H + u2_0 + u) * KX2 + (size_t)s * (H + E) + H + kk];
        }
      }
      out[o + i] = (OutT)v;
    }
    o += n;
  }
}

}  // namespace lat
