// latency.cuh -- role-specialised persistent decoder for the batch-1 latency path (sm_100a).
//
// One cooperative launch, 148 co-resident CTAs, NO grid-wide barriers.  CTAs take roles:
//
//   LSTM CTAs (128 on a B200)   own a fixed slice of hidden units of the attention LSTM(s) and of the
//       decoder LSTM.  Their weights are pre-packed (fp32 or fp16) into one contiguous per-CTA stream
//       in consumption order.  A prefix of the stream (critical-path segments first) is staged ONCE
//       with TMA bulk copies (cp.async.bulk + mbarrier) and stays RESIDENT in ~195 KB of shared memory
//       for the whole utterance; the rest is streamed every frame straight from L2/HBM with 16-byte
//       read-only loads issued 8 deep per lane.  One warp = one hidden unit (4 gate rows at once),
//       activation slice in registers, 6-shuffle butterfly reduction of the 4 gate sums.
//   attention CTAs (8 + 4)      keep processed_memory, a feature slice of the encoder memory and the
//       alignment state resident in shared memory; fused energy -> sigmoid -> stepwise-monotonic
//       update -> context reduction with warp shuffles.
//   aux CTAs (8)                keep the mel/gate projection and prenet weights resident and run
//       projection -> stop test -> prenet layer 0 -> layer 1 for the next frame.
//
// Vectors that cross CTAs (h, context, query partials, prenet, mel) travel through global memory in an
// "LL" protocol: each element is one 64-bit word {fp32 value, frame tag}, written with a single 8-byte
// store and polled by the consumers (all of a thread's words are requested before any is checked) --
// no fences, no barriers; one L2 round trip per dependency instead of a grid barrier.
//
// Per-frame program of an LSTM CTA (segments ordered by when their inputs become available;
// reference arithmetic: model.py:337-346 attention LSTM, :362-373 decoder LSTM):
//   a  attn-LSTM  W_hh . h1[t-1]          d  attn-LSTM  W_ih[:, :P] . prenet[t]   -> h1[t], q partials
//   b  attn-LSTM  W_ih[:, P:] . ctx[t-1]  e  dec-LSTM   W_ih[:, h cols] . h1[t]   (one step per stream)
//   c  dec-LSTM   W_hh . h2[t-1]          f  dec-LSTM   W_ih[:, ctx cols] . ctx[t] -> h2[t]
//
// Early weight requests (kernel variant PRE, fp32 storage): a weight load does not depend on the activation it will be
// multiplied with, so the L2/HBM-streamed steps request their 16-byte loads one phase early and hold them in registers --
// c before the h2 exchange is polled, b behind the h2 publish of the previous frame, the first half pass of a before the
// h1 exchange (a then runs before e).  Every such request sits BEHIND a barrier that follows the CTA's own LL stores: the
// warps that have no cells / query partials to compute would otherwise fill the load/store queue in front of them.
// With fp32 storage the 256 KB of tensor memory per SM (no MMA is issued here) hold the segments f and e1 (e0, e1 and a
// with 16-bit storage); see lat_build_plan.
//
// (An earlier revision streamed the non-resident weights through a TMA ring as well; measured on B200 the
// single producer thread + 64-96 KB ring capped a CTA at ~10-20 B/clk, below what plain deep LDG streams
// reach, so the ring was removed -- see DESIGN.md "what did not work".)
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lat {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kLLDepth = 8;             // ring depth of every LL vector (frames)
constexpr int kAux = 8;                 // aux CTAs
constexpr int kRep = 8;                 // replicas of every vector polled by all LSTM CTAs (spreads the L2 hot spot)
constexpr int kMaxU1 = 20, kMaxU2 = 12; // max hidden units per LSTM CTA (attention / decoder LSTM)
constexpr int kSteps = 7;               // a b c d e0 e1 f
constexpr long long kWatchdogClocks = 6000000000LL;

// fixed model dims of this path (hparams defaults); other shapes use the generic kernel
constexpr int H = 1024, E = 512, P = 256, A = 128, M = 80;

struct LatStream {
  const float *b_ih, *b_hh;      // attention LSTM biases [4H]
  const float *wq;               // [A, H]
  const float *v;                // [A]
  const float *loc_conv, *loc_dense;   // location-sensitive attention: conv [LF, 2, LK], dense [A, LF] (else null)
  const float *pre_w0, *pre_w1;  // [P, M], [P, P]
  const float *mem;              // [T_s, E]   (batch 1)
  const float *pm;               // [T_s, A]   processed memory (precomputed)
  const float *pre_tf;           // teacher-forced: hoisted prenet output [T+1, P]
  const float *noise;            // [T, T_s] or null
  const uint8_t *keep0, *keep1;  // prenet keep masks [rows, P] or null
  float *align;                  // [Tcap, T_s]
  const long long* len;          // [1] valid length or null
  int Ts;                        // padded width
  int na;                        // attention CTAs serving this stream (feature slice = E / na)
};

struct LatParams {
  int S, NL, NL1;                // streams, LSTM CTAs, LSTM CTAs per stream (attention LSTM split)
  int free_running, training, n_steps, Tcap;
  int lsa, LF, LK;               // location-sensitive attention (attention.py:25-85) instead of stepwise-monotonic; its conv shape
  float gate_thr, p_att, p_dec;
  unsigned thresh_pre, thresh_att, thresh_dec;
  unsigned long long seed;
  int wbytes;                    // bytes per packed weight element (4 = fp32, 2 = fp16)
  const unsigned char* packed;   // per-LSTM-CTA weight streams
  const unsigned long long* packed_off;  // [NL+1] byte offsets
  int res_budget;                // shared-memory bytes available for the resident prefix
  int l2_keep_mask;              // bit s set: step s's streamed weights use L2 evict_last, else evict_first
  int use_tmem;                  // 0: off; 1: weight segments e0, e1 (and a when it fits: 16-bit storage) live in tensor memory; 2: b, c
  int stream_prefetch;           // 1: the L2-streamed single-pass steps b and c request their weights one phase early (registers)
  LatStream st[2];
  const float *d_b_ih, *d_b_hh;  // decoder LSTM biases [4H]
  const float *proj_w, *proj_b, *gate_w, *gate_b;
  const uint8_t* lstm_keep;      // [T, 6, H] or null
  float *mel, *gate;             // [Tcap, M], [Tcap]
  int *n_frames, *reached_max;
  // LL exchange buffers (64-bit {value, tag}); zeroed by the host before launch
  unsigned long long *ll_h1, *ll_h2, *ll_ctx, *ll_q, *ll_pre, *ll_mel, *ll_l0;
  unsigned long long *ll_e;      // partial energies [depth][2 streams][8 slices][kLatTsCap]
  unsigned* aux_done;
  int* abort_flag;
  long long* phase_clocks;       // [16] diagnostics (LSTM CTA 0)
  long long* dbg;                // optional [256] cycle sums (TACO2DEC_LAT_DEBUG=1): [0,64) per-step detail of LSTM CTA 0
                                 // (step*4 + {dot, reduce, calls}), [64,80) attention CTA (0,0), [80,96) attention CTA (1,0),
                                 // [96,112) aux CTA 0
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
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// streamed weights: read-only, do not pollute L1, explicit L2 eviction policy per segment
__device__ __forceinline__ uint4 ldg_stream(const unsigned char* p, unsigned long long policy) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

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

// replica r of a replicated LL vector (layout [rep][depth][...])
__device__ __forceinline__ unsigned long long* rep_h1(const LatParams& p, int r) { return p.ll_h1 + (size_t)r * kLLDepth * 2 * H; }
__device__ __forceinline__ unsigned long long* rep_h2(const LatParams& p, int r) { return p.ll_h2 + (size_t)r * kLLDepth * H; }
__device__ __forceinline__ unsigned long long* rep_ctx(const LatParams& p, int r) { return p.ll_ctx + (size_t)r * kLLDepth * 2 * E; }
__device__ __forceinline__ unsigned long long* rep_pre(const LatParams& p, int r) { return p.ll_pre + (size_t)r * kLLDepth * 2 * (P + 8); }

struct Watch {
  int* abort_flag;
  long long t0;
  unsigned spins;
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
    if (w.expired()) return false;
  }
}

// Every thread owns the words i = tid + k * kThreads (k < KMAX) of an n-word LL vector.  All owned words
// are requested back to back before the first tag is inspected, so a thread pays ONE L2 round trip
// per poll round instead of one per word.  Values land in dst[i].
template <int KMAX>
__device__ __forceinline__ bool poll_vector(const unsigned long long* src, int n, unsigned tag, float* dst, int tid,
                                            Watch& wd) {
  unsigned pending = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (tid + k * kThreads < n) pending |= 1u << k;
  wd.arm();
  while (pending) {
    unsigned long long w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (pending & (1u << k)) w[k] = ll_load(src + tid + k * kThreads);
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if ((pending & (1u << k)) && (unsigned)(w[k] >> 32) == tag) {
        dst[tid + k * kThreads] = __uint_as_float((unsigned)w[k]);
        pending &= ~(1u << k);
      }
    if (pending && wd.expired()) return false;
  }
  return true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Reduce four per-lane partial sums over the warp with 6 shuffles (instead of 20): afterwards lanes
// 0-7 hold the total of v0, 8-15 of v1, 16-23 of v2, 24-31 of v3.
__device__ __forceinline__ float butterfly4(float v0, float v1, float v2, float v3, int lane) {
  const bool hi = (lane & 16) != 0;
  float a = hi ? v2 : v0, b = hi ? v3 : v1;
  const float sa = hi ? v0 : v2, sb = hi ? v1 : v3;
  a += __shfl_xor_sync(0xffffffffu, sa, 16);
  b += __shfl_xor_sync(0xffffffffu, sb, 16);
  const bool h8 = (lane & 8) != 0;
  float c = h8 ? b : a;
  const float sc = h8 ? a : b;
  c += __shfl_xor_sync(0xffffffffu, sc, 8);
  c += __shfl_xor_sync(0xffffffffu, c, 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}
// fast transcendental forms (ex2 / fast divide): abs error ~1e-6, used on the attention energies where
// T*A evaluations per frame sit on the critical path.
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = exp2f(x * 2.8853900817779268f);  // e^(2x)
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// philox_keep / philox_normal come from the including translation unit (taco2dec.cu)
__device__ __forceinline__ bool philox_keep_l(unsigned long long seed, int mask_id, int row, int idx, unsigned thresh) {
  return philox_keep(seed, mask_id, row, idx, thresh);
}

// ---------------------------------------------------------------------------------------------
// LSTM CTA
// ---------------------------------------------------------------------------------------------
struct StepPlan {
  int n_units;       // hidden units (4-row chunks) in this step
  int K;             // columns per row
  int ksplit;        // column split: items = n_units * ksplit, one item per warp when it fits
  int chunk_bytes;   // 4 rows * K * wbytes
  int n_res;         // first n_res units are resident in shared memory
  int res_off;       // byte offset of the step's resident chunks inside the resident region
  long long src_off; // byte offset of the step's first chunk inside this CTA's packed stream
  unsigned long long policy;  // L2 eviction policy of the step's streamed loads
  int tmem_col;      // >= 0: the step's weights live in tensor memory (one item per warp), column offset inside the
                     // warp's 128-column slice; -1: shared memory / global stream
};

// 16-byte unit -> multiply-accumulate against the activation registers
template <int WB>
struct Mac;
template <>
struct Mac<4> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void run(const uint4& w, const float* x, float& a0, float& a1) {
    a0 = fmaf(__uint_as_float(w.x), x[0], a0);
    a1 = fmaf(__uint_as_float(w.y), x[1], a1);
    a0 = fmaf(__uint_as_float(w.z), x[2], a0);
    a1 = fmaf(__uint_as_float(w.w), x[3], a1);
  }
};
template <>
struct Mac<2> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void run(const uint4& w, const float* x, float& a0, float& a1) {
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
    const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
    const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&w.z));
    const float2 p3 = __half22float2(*reinterpret_cast<const __half2*>(&w.w));
    a0 = fmaf(p0.x, x[0], a0); a1 = fmaf(p0.y, x[1], a1);
    a0 = fmaf(p1.x, x[2], a0); a1 = fmaf(p1.y, x[3], a1);
    a0 = fmaf(p2.x, x[4], a0); a1 = fmaf(p2.y, x[5], a1);
    a0 = fmaf(p3.x, x[6], a0); a1 = fmaf(p3.y, x[7], a1);
  }
};

struct LstmShared {
  unsigned char* resident;
  const unsigned char* gstream;             // this CTA's packed stream in global memory
  float *xh1, *xctx, *xh2, *xpre;           // activation segments [2H], [2E], [H], [P]
  float *wq_s;                              // [A][kMaxU1] query-weight slice of this CTA's units
  float *acc1, *acc2;                       // gate pre-activations [ksplit 2][unit][4]
  float *c1, *c2, *hloc;                    // cell state / fresh h1 of own units
  float *bias1, *bias2;                     // b_ih + b_hh of own units [u][4]
  uint64_t* res_bar;                        // mbarrier of the one-off resident TMA load
  StepPlan* plan;                           // [kSteps]
  volatile int* exit_flag;
  uint32_t tmem_base;                       // tensor-memory allocation of this CTA (all 512 columns)
  long long* dbg_s;                         // diagnostics of LSTM CTA 0 (shared memory [64]) or null
};

// ---- tensor memory as weight storage --------------------------------------------------------------------------
// The latency kernel issues no MMAs, so the SM's 256 KB of tensor memory would sit idle.  It is exactly the size of the
// two weight segments that otherwise stream from HBM every frame (b: 16 units x 8 KB, c: 8 units x 16 KB per CTA in
// fp32): they are copied in once (tcgen05.st) and read back every frame with tcgen05.ld.  A warp can address only its
// lane quarter (32 * (warp & 3)), so warp w owns columns [128 * (w >> 2), +128) of that quarter: 64 columns for its
// item of step b, 64 for its item of step c; lane l's column (chunk * 4 + gate) * 4 + e holds the e-th float of the
// 16-byte unit that lane would otherwise load (same order as dot4).
__device__ __forceinline__ uint32_t tmem_addr_of(uint32_t base, int warp, int col) {
  return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128 + col);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// one item (4 gate rows x NUC * 32 sixteen-byte units) out of tensor memory: NUC chunks of 16 columns = [gate][4 words]
template <int WB, int NUC>
__device__ __forceinline__ void dot4_tmem(uint32_t taddr, const float* xs, int lane, float (&s)[4]) {
  constexpr int EPU = Mac<WB>::kElems;
  uint32_t w[NUC][16];
#pragma unroll
  for (int c = 0; c < NUC; ++c) tmem_ld16(taddr + 16 * c, w[c]);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  float a[4][2];
#pragma unroll
  for (int g = 0; g < 4; ++g) a[g][0] = a[g][1] = 0.f;
#pragma unroll
  for (int c = 0; c < NUC; ++c) {
    float x[EPU];
#pragma unroll
    for (int e = 0; e < EPU; e += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xs + (size_t)(lane + 32 * c) * EPU + e);
      x[e + 0] = v.x; x[e + 1] = v.y; x[e + 2] = v.z; x[e + 3] = v.w;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 u = make_uint4(w[c][g * 4], w[c][g * 4 + 1], w[c][g * 4 + 2], w[c][g * 4 + 3]);
      Mac<WB>::run(u, x, a[g][0], a[g][1]);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) s[g] = a[g][0] + a[g][1];
}
// one-off: copy this warp's item of a step from the packed global stream into its tensor-memory columns
__device__ __noinline__ void tmem_fill_item(uint32_t taddr, const unsigned char* base, int row_bytes, int lane, int nuc) {
  for (int c = 0; c < nuc; ++c) {
    uint32_t w[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + (size_t)g * row_bytes + (size_t)(lane + 32 * c) * 16);
      w[g * 4] = u.x; w[g * 4 + 1] = u.y; w[g * 4 + 2] = u.z; w[g * 4 + 3] = u.w;
    }
    tmem_st16(taddr + 16 * c, w);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// One item = one hidden unit (4 gate rows) x one column range of KLEN columns, done by one warp.
// All 4 * PU sixteen-byte weight loads of a pass are issued before the first is used (memory-level
// parallelism is what hides the L2/HBM latency of the streamed part); the activation slice is read from
// shared memory right behind them.
template <int WB, int KLEN, int PU, bool kGlobal>
__device__ __forceinline__ void dot4(const unsigned char* base, int row_bytes, const float* xs, int lane, float (&s)[4],
                                     unsigned long long policy) {
  constexpr int EPU = Mac<WB>::kElems;
  constexpr int NU = KLEN / (32 * EPU);
  static_assert(NU >= 1 && NU % PU == 0, "bad pass width");
  float a[4][2];
#pragma unroll
  for (int g = 0; g < 4; ++g) a[g][0] = a[g][1] = 0.f;
#pragma unroll
  for (int i0 = 0; i0 < NU; i0 += PU) {
    uint4 w[4][PU];
#pragma unroll
    for (int i = 0; i < PU; ++i)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const unsigned char* ptr = base + (size_t)g * row_bytes + (size_t)(lane + 32 * (i0 + i)) * 16;
        w[g][i] = kGlobal ? ldg_stream(ptr, policy) : *reinterpret_cast<const uint4*>(ptr);
      }
#pragma unroll
    for (int i = 0; i < PU; ++i) {
      float x[EPU];
#pragma unroll
      for (int e = 0; e < EPU; e += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xs + (size_t)(lane + 32 * (i0 + i)) * EPU + e);
        x[e + 0] = v.x; x[e + 1] = v.y; x[e + 2] = v.z; x[e + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) Mac<WB>::run(w[g][i], x, a[g][0], a[g][1]);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) s[g] = a[g][0] + a[g][1];
}

// NOT inlined on purpose: the seven call sites per frame share three instantiations, which keeps the per-frame
// instruction footprint inside the instruction cache (inlined, the kernel was 280 KB of SASS and every step
// paid instruction-fetch misses).
template <int WB, int KLEN>
__device__ __noinline__ void consume_items(const LstmShared& sh, const StepPlan& sp, const float* xs, float* acc,
                                           int max_units, int warp, int lane) {
  constexpr int EPU = Mac<WB>::kElems;
  constexpr int NU = KLEN / (32 * EPU);
#ifndef TACO2DEC_STREAM_PU
#define TACO2DEC_STREAM_PU 4
#endif
  constexpr int PU = NU >= 4 ? 4 : NU;                                  // resident (shared memory) pass width
  constexpr int PUG = NU >= TACO2DEC_STREAM_PU ? TACO2DEC_STREAM_PU : NU;   // streamed (L2/HBM) pass width
  const int n_items = sp.n_units * sp.ksplit;
#ifdef TACO2DEC_LAT_STAMPS
  const bool dbg_on = sh.dbg_s && warp == 0 && lane == 0;
#else
  constexpr bool dbg_on = false;
#endif
  long long d_dot = 0, d_red = 0;
  for (int it = warp; it < n_items; it += kWarps) {
    const long long d0 = dbg_on ? clock64() : 0;
    const int unit = it / sp.ksplit, kh = it - unit * sp.ksplit;
    const float* xk = xs + (size_t)kh * KLEN;
    float s[4];
    const int row_bytes = sp.K * WB;
    const size_t col_off = (size_t)kh * KLEN * WB;
    if (unit < sp.n_res) {
      dot4<WB, KLEN, PU, false>(sh.resident + sp.res_off + (size_t)unit * sp.chunk_bytes + col_off, row_bytes, xk, lane, s, 0ull);
    } else {
      dot4<WB, KLEN, PUG, true>(sh.gstream + sp.src_off + (size_t)unit * sp.chunk_bytes + col_off, row_bytes, xk, lane, s,
                                 sp.policy);
    }
    long long d1 = 0;
    if (dbg_on) { volatile float sink = s[0] + s[1] + s[2] + s[3]; (void)sink; d1 = clock64(); d_dot += d1 - d0; }
    const float c = butterfly4(s[0], s[1], s[2], s[3], lane);
    if ((lane & 7) == 0) acc[((size_t)kh * max_units + unit) * 4 + (lane >> 3)] += c;
    if (dbg_on) d_red += clock64() - d1;
  }
  if (dbg_on) {
    const int si = (int)(&sp - sh.plan);
    sh.dbg_s[si * 4 + 0] += d_dot; sh.dbg_s[si * 4 + 1] += d_red; sh.dbg_s[si * 4 + 2] += 1;
  }
}

// A fully streamed step whose items map one-to-one onto the warps and fit ONE pass (16 sixteen-byte loads per lane) can
// request its weights before the exchange / pointwise phase in front of it: the weights do not depend on the activations,
// the 16 loads sit in registers while the warp polls, and the step itself shrinks to the multiply-accumulate.  Same
// summation order as dot4 (one pass, units in order), so the results are bit-identical to consume_items.
template <int WB, int KLEN>
__device__ __forceinline__ void stream_issue(const LstmShared& sh, const StepPlan& sp, int warp, int lane,
                                             uint4 (&w)[4][KLEN / (32 * Mac<WB>::kElems)]) {
  constexpr int NU = KLEN / (32 * Mac<WB>::kElems);
  const int unit = warp / sp.ksplit, kh = warp - unit * sp.ksplit;
  const unsigned char* base = sh.gstream + sp.src_off + (size_t)unit * sp.chunk_bytes + (size_t)kh * KLEN * WB;
  const int row_bytes = sp.K * WB;
#pragma unroll
  for (int i = 0; i < NU; ++i)
#pragma unroll
    for (int g = 0; g < 4; ++g) w[g][i] = ldg_stream(base + (size_t)g * row_bytes + (size_t)(lane + 32 * i) * 16, sp.policy);
}
template <int WB, int KLEN>
__device__ __forceinline__ void stream_finish(const StepPlan& sp, const float* xs, float* acc, int max_units, int warp,
                                              int lane, const uint4 (&w)[4][KLEN / (32 * Mac<WB>::kElems)]) {
  constexpr int EPU = Mac<WB>::kElems;
  constexpr int NU = KLEN / (32 * EPU);
  const int unit = warp / sp.ksplit, kh = warp - unit * sp.ksplit;
  const float* xk = xs + (size_t)kh * KLEN;
  float a[4][2];
#pragma unroll
  for (int g = 0; g < 4; ++g) a[g][0] = a[g][1] = 0.f;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    float x[EPU];
#pragma unroll
    for (int e = 0; e < EPU; e += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xk + (size_t)(lane + 32 * i) * EPU + e);
      x[e + 0] = v.x; x[e + 1] = v.y; x[e + 2] = v.z; x[e + 3] = v.w;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) Mac<WB>::run(w[g][i], x, a[g][0], a[g][1]);
  }
  const float c = butterfly4(a[0][0] + a[0][1], a[1][0] + a[1][1], a[2][0] + a[2][1], a[3][0] + a[3][1], lane);
  if ((lane & 7) == 0) acc[((size_t)kh * max_units + unit) * 4 + (lane >> 3)] += c;
}

// a step whose weights live in tensor memory: exactly one item per warp.  Kept out of consume_items so that the
// register allocation of the streamed / shared-memory paths is unaffected.
template <int WB>
__device__ __noinline__ void consume_tmem_step(const LstmShared& sh, const StepPlan& sp, const float* xs, float* acc,
                                               int max_units, int warp, int lane) {
  const int unit = warp / sp.ksplit, kh = warp - unit * sp.ksplit;
  const int klen = sp.K / sp.ksplit;
  const uint32_t taddr = tmem_addr_of(sh.tmem_base, warp, sp.tmem_col);
#ifdef TACO2DEC_LAT_STAMPS
  const bool dbg_on = sh.dbg_s && warp == 0 && lane == 0;
#else
  constexpr bool dbg_on = false;
#endif
  const long long d0 = dbg_on ? clock64() : 0;
  float s[4];
  if (klen * WB == 2048) dot4_tmem<WB, 4>(taddr, xs + (size_t)kh * klen, lane, s);     // 64 columns
  else dot4_tmem<WB, 2>(taddr, xs + (size_t)kh * klen, lane, s);                        // 32 columns
  long long d1 = 0;
  if (dbg_on) { volatile float sink = s[0] + s[1] + s[2] + s[3]; (void)sink; d1 = clock64(); }
  const float c = butterfly4(s[0], s[1], s[2], s[3], lane);
  if ((lane & 7) == 0) acc[((size_t)kh * max_units + unit) * 4 + (lane >> 3)] += c;
  if (dbg_on) {
    const int si = (int)(&sp - sh.plan);
    sh.dbg_s[si * 4 + 0] += d1 - d0; sh.dbg_s[si * 4 + 1] += clock64() - d1; sh.dbg_s[si * 4 + 2] += 1;
  }
}

// a pass of NU sixteen-byte units per lane and row, starting at unit i0 (step a is two such passes per warp)
template <int WB, int NU>
__device__ __forceinline__ void stream_issue_units(const LstmShared& sh, const StepPlan& sp, int warp, int lane, int i0,
                                                   uint4 (&w)[4][NU]) {
  const int unit = warp / sp.ksplit, kh = warp - unit * sp.ksplit;
  const int row_bytes = sp.K * WB;
  const unsigned char* base = sh.gstream + sp.src_off + (size_t)unit * sp.chunk_bytes + (size_t)kh * (row_bytes / sp.ksplit);
#pragma unroll
  for (int i = 0; i < NU; ++i)
#pragma unroll
    for (int g = 0; g < 4; ++g)
      w[g][i] = ldg_stream(base + (size_t)g * row_bytes + (size_t)(lane + 32 * (i0 + i)) * 16, sp.policy);
}
template <int WB, int NU>
__device__ __forceinline__ void stream_fma_units(const float* xk, int lane, int i0, const uint4 (&w)[4][NU], float (&a)[4][2]) {
  constexpr int EPU = Mac<WB>::kElems;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    float x[EPU];
#pragma unroll
    for (int e = 0; e < EPU; e += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xk + (size_t)(lane + 32 * (i0 + i)) * EPU + e);
      x[e + 0] = v.x; x[e + 1] = v.y; x[e + 2] = v.z; x[e + 3] = v.w;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) Mac<WB>::run(w[g][i], x, a[g][0], a[g][1]);
  }
}

// runtime column-length dispatch (K / ksplit is one of 1024, 512, 256)
template <int WB>
__device__ __forceinline__ void consume_step(const LstmShared& sh, const StepPlan& sp, const float* xs, float* acc,
                                             int max_units, int warp, int lane) {
  if (sp.n_units == 0) return;
  if (sp.tmem_col >= 0) { consume_tmem_step<WB>(sh, sp, xs, acc, max_units, warp, lane); return; }
  const int klen = sp.K / sp.ksplit;
  if (klen == 1024) consume_items<WB, 1024>(sh, sp, xs, acc, max_units, warp, lane);
  else if (klen == 512) consume_items<WB, 512>(sh, sp, xs, acc, max_units, warp, lane);
  else consume_items<WB, 256>(sh, sp, xs, acc, max_units, warp, lane);
}

// LSTM pointwise for nu hidden units: lane 4u+g evaluates gate g (i, f, g, o), the 4-lane group combines
// them with shuffles; returns true in the lanes (g == 0) that own (hn, cn) of unit u.
__device__ __forceinline__ bool lstm_pointwise(const float* acc, int max_units, const float* bias, const float* c_state,
                                               int nu, int tid, int& u_out, float& hn, float& cn) {
  const int nact = nu * 4;
  if ((tid & ~31) >= nact) return false;                 // whole warp idle
  const bool valid = tid < nact;
  const int u = valid ? (tid >> 2) : 0, g = tid & 3;
  const float pre = (acc[u * 4 + g] + acc[(max_units + u) * 4 + g]) + bias[u * 4 + g];
  const float act = (g == 2) ? tanhf(pre) : sigmoid_acc(pre);
  const int base = (tid & 31) & ~3;
  const float ig = __shfl_sync(0xffffffffu, act, base + 0);
  const float fg = __shfl_sync(0xffffffffu, act, base + 1);
  const float gg = __shfl_sync(0xffffffffu, act, base + 2);
  const float og = __shfl_sync(0xffffffffu, act, base + 3);
  if (!valid || g != 0) return false;
  cn = fg * c_state[u] + ig * gg;
  hn = og * tanhf(cn);
  u_out = u;
  return true;
}

// Per-CTA step plan: units, column split, tensor-memory / shared-memory residency.  Host and device run the same code (the
// host uses it to decide whether every LSTM CTA can take the early-request variant of the kernel).
__host__ __device__ inline void lat_build_plan(const LatParams& p, int lc, int wb, StepPlan* plan) {
  const int S = p.S;
  const int s1 = lc / p.NL1, i1 = lc - s1 * p.NL1;
  const int nu1 = (int)((long long)(i1 + 1) * H / p.NL1) - (int)((long long)i1 * H / p.NL1);
  const int nu2 = (int)((long long)(lc + 1) * H / p.NL) - (int)((long long)lc * H / p.NL);
  // steps:            a     b     c     d     e0    e1            f
  const int nun[kSteps] = {nu1, nu1, nu2, nu1, nu2, S == 2 ? nu2 : 0, nu2};
  const int ks[kSteps] = {H, E, H, P, H, H, S * E};
  long long src = 0;
  for (int s = 0; s < kSteps; ++s) {
    StepPlan& sp = plan[s];
    sp.n_units = nun[s]; sp.K = ks[s]; sp.chunk_bytes = 4 * ks[s] * wb; sp.n_res = 0; sp.res_off = 0;
    sp.ksplit = (nun[s] * 2 <= kWarps && ks[s] >= 512) ? 2 : 1;   // keep every warp busy
    sp.policy = 0ull;
    sp.src_off = src;
    src += (long long)nun[s] * sp.chunk_bytes;
    sp.tmem_col = -1;
  }
  // tensor memory (one item per warp, 128 columns per warp): the segments e0, e1 first, then a
  // (use_tmem == 2: b, c instead -- kept for the record, slower)
  if (p.use_tmem) {
    // use_tmem == 3: f (on the chain between the context and h2) and e1; e0 moves to shared memory in f's place (measured:
    // f 2.66 -> 1.47 kcyc, e 2.92 -> 3.97, frame unchanged).  use_tmem == 4: f and d; e1 half streamed (slower).
    const int cand1[3] = {4, 5, 0}, cand2[3] = {1, 2, -1}, cand3[3] = {6, 5, 0}, cand4[3] = {6, 3, -1};
    const int* cand = p.use_tmem == 2 ? cand2 : p.use_tmem == 3 ? cand3 : p.use_tmem == 4 ? cand4 : cand1;
    int col = 0;
    for (int k = 0; k < 3; ++k) {
      const int s = cand[k];
      if (s < 0) continue;
      StepPlan& sp = plan[s];
      const int item_bytes = sp.ksplit > 0 ? sp.chunk_bytes / sp.ksplit : 0;      // 4 rows x klen x WB
      const int cols = item_bytes / 128;                                          // 32 lanes x 4 bytes per column
      if (sp.n_units * sp.ksplit != kWarps || (cols != 32 && cols != 64) || col + cols > 128) continue;
      sp.tmem_col = col;
      col += cols;
    }
  }
  // residency: critical-path steps first (d, f), then e1, e0, c, b, a
  const int prio[kSteps] = {3, 6, 5, 4, 2, 1, 0};
  int left = p.res_budget, roff = 0;
  for (int k = 0; k < kSteps; ++k) {
    StepPlan& sp = plan[prio[k]];
    if (sp.tmem_col >= 0) { sp.n_res = 0; sp.res_off = roff; continue; }
    int n = sp.chunk_bytes > 0 ? left / sp.chunk_bytes : 0;
    if (n > sp.n_units) n = sp.n_units;
    sp.n_res = n; sp.res_off = roff;
    roff += n * sp.chunk_bytes; left -= n * sp.chunk_bytes;
  }
}
__host__ __device__ inline bool stream_prefetchable(const StepPlan& sp, int klen) {
  return sp.tmem_col < 0 && sp.n_res == 0 && sp.n_units * sp.ksplit == kWarps && sp.K == klen * sp.ksplit;
}
constexpr int kPreK = 512;   // column length of the early-request steps (b: own context, c: half of h2)

template <int WB, bool PRE>
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
  sh.resident = take((size_t)p.res_budget);
  sh.xh1 = (float*)take(sizeof(float) * 2 * H);
  sh.xctx = (float*)take(sizeof(float) * 2 * E);
  sh.xh2 = (float*)take(sizeof(float) * H);
  sh.xpre = (float*)take(sizeof(float) * (P + 8));
  sh.wq_s = (float*)take(sizeof(float) * A * kMaxU1);
  sh.acc1 = (float*)take(sizeof(float) * 2 * kMaxU1 * 4);
  sh.acc2 = (float*)take(sizeof(float) * 2 * kMaxU2 * 4);
  sh.c1 = (float*)take(sizeof(float) * kMaxU1);
  sh.c2 = (float*)take(sizeof(float) * kMaxU2);
  sh.hloc = (float*)take(sizeof(float) * kMaxU1);
  sh.bias1 = (float*)take(sizeof(float) * kMaxU1 * 4);
  sh.bias2 = (float*)take(sizeof(float) * kMaxU2 * 4);
  sh.res_bar = (uint64_t*)take(sizeof(uint64_t));
  sh.plan = (StepPlan*)take(sizeof(StepPlan) * kSteps);
  sh.exit_flag = (volatile int*)take(sizeof(int));
  sh.gstream = p.packed + p.packed_off[lc];

  // ---- step plan + one-off TMA load of the resident prefix (thread 0) ---------------------
  if (tid == 0) {
    lat_build_plan(p, lc, WB, sh.plan);
    // segments marked in l2_keep_mask are asked to stay in L2 across frames, the others are streamed
    // through (evict-first) so they do not push the kept ones out
    for (int s = 0; s < kSteps; ++s)
      sh.plan[s].policy = ((p.l2_keep_mask >> s) & 1) ? l2_policy_evict_last() : l2_policy_evict_first();
    mbar_init(sh.res_bar, 1);
    fence_barrier_init();
    *sh.exit_flag = 0;
    unsigned res_total = 0;
    for (int s = 0; s < kSteps; ++s) res_total += (unsigned)(sh.plan[s].n_res * sh.plan[s].chunk_bytes);
    if (res_total) {
      mbar_expect_tx(sh.res_bar, res_total);
      for (int s = 0; s < kSteps; ++s) {
        const StepPlan& sp = sh.plan[s];
        for (int c = 0; c < sp.n_res; ++c)
          tma_load_1d(sh.resident + sp.res_off + (size_t)c * sp.chunk_bytes,
                      sh.gstream + sp.src_off + (size_t)c * sp.chunk_bytes, (unsigned)sp.chunk_bytes, sh.res_bar);
      }
    } else {
      mbar_arrive(sh.res_bar);
    }
  }
  // zero state, load biases and the query-weight slice
  for (int i = tid; i < 2 * H; i += kThreads) sh.xh1[i] = 0.f;
  for (int i = tid; i < 2 * E; i += kThreads) sh.xctx[i] = 0.f;
  for (int i = tid; i < H; i += kThreads) sh.xh2[i] = 0.f;
  for (int i = tid; i < P + 8; i += kThreads) sh.xpre[i] = 0.f;
  for (int i = tid; i < 2 * kMaxU1 * 4; i += kThreads) sh.acc1[i] = 0.f;
  for (int i = tid; i < 2 * kMaxU2 * 4; i += kThreads) sh.acc2[i] = 0.f;
  for (int i = tid; i < kMaxU1 * 4; i += kThreads) {
    const int u = i >> 2, g = i & 3;
    sh.bias1[i] = u < nu1 ? p.st[s1].b_ih[g * H + u1_0 + u] + p.st[s1].b_hh[g * H + u1_0 + u] : 0.f;
  }
  for (int i = tid; i < kMaxU2 * 4; i += kThreads) {
    const int u = i >> 2, g = i & 3;
    sh.bias2[i] = u < nu2 ? p.d_b_ih[g * H + u2_0 + u] + p.d_b_hh[g * H + u2_0 + u] : 0.f;
  }
  for (int i = tid; i < kMaxU1; i += kThreads) { sh.c1[i] = 0.f; sh.hloc[i] = 0.f; }
  for (int i = tid; i < kMaxU2; i += kThreads) sh.c2[i] = 0.f;
  for (int i = tid; i < A * kMaxU1; i += kThreads) {
    const int a = i / kMaxU1, u = i - a * kMaxU1;
    sh.wq_s[i] = u < nu1 ? p.st[s1].wq[(size_t)a * H + u1_0 + u] : 0.f;
  }
  __syncthreads();

  // ---- tensor memory: allocate all 512 columns, copy this warp's items of steps b / c in once ----------------------
  __shared__ uint32_t tmem_base_s;
  bool any_tmem = false;
  for (int s = 0; s < kSteps; ++s) any_tmem = any_tmem || sh.plan[s].tmem_col >= 0;
  sh.tmem_base = 0;
  if (any_tmem) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    sh.tmem_base = tmem_base_s;
    for (int s = 0; s < kSteps; ++s) {
      const StepPlan& sp = sh.plan[s];
      if (sp.tmem_col < 0) continue;
      const int unit = warp / sp.ksplit, kh = warp - unit * sp.ksplit;
      const int klen = sp.K / sp.ksplit;
      tmem_fill_item(tmem_addr_of(sh.tmem_base, warp, sp.tmem_col),
                     sh.gstream + sp.src_off + (size_t)unit * sp.chunk_bytes + (size_t)kh * klen * WB, sp.K * WB, lane,
                     klen * WB / 512);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  Watch wd{p.abort_flag, 0, 0};
  bool ok = true;
  const int rep = lc % kRep;
  // wait for the resident prefix (async-proxy writes become visible through the mbarrier)
  wd.arm();
  while (!mbar_try_wait(sh.res_bar, 0)) {
    if (wd.expired()) { ok = false; break; }
  }
  const float sc_att = 1.0f / (1.0f - p.p_att), sc_dec = 1.0f / (1.0f - p.p_dec);
  const int n_steps = p.n_steps;
  // phase accounting of CTA 0 (diagnostics): accumulators in shared memory so that they cost no registers
  __shared__ long long ph[16];
  __shared__ long long dbg_sm[64];
  if (tid < 16) ph[tid] = 0;
  if (tid < 64) dbg_sm[tid] = 0;
  sh.dbg_s = (p.dbg && lc == 0) ? dbg_sm : nullptr;
  __syncthreads();
  long long ph_t = clock64();
#define LPH(slot)                                         \
  if (lc == 0 && tid == 0) {                              \
    const long long n_ = clock64();                       \
    ph[slot] += n_ - ph_t;                                \
    ph_t = n_;                                            \
  }

  const StepPlan* pl = sh.plan;
  // steps b and c one phase early (see stream_issue); uniform over the CTA
  // (PRE is chosen by the host: every LSTM CTA's plan has b and c fully streamed, one item per warp, one pass)
  constexpr int kPreNU = kPreK / (32 * Mac<WB>::kElems);
  uint4 wpre[4][kPreNU];
  if constexpr (PRE) stream_issue<WB, kPreK>(sh, pl[1], warp, lane, wpre);
  for (int t = 0; t < n_steps; ++t) {
    const unsigned tag_prev = (unsigned)t;        // values produced during frame t-1
    const unsigned tag_cur = (unsigned)t + 1u;    // values produced during frame t
    const int rb = t % kLLDepth, rb_prev = (t + kLLDepth - 1) % kLLDepth;

    // teacher-forced frames: request this frame's hoisted prenet row and check the LL flow control now, so that neither
    // L2 round trip sits between two dependent steps later in the frame
    float pre_tf_val = 0.f;
    if (!p.free_running) {
      if (tid < P) pre_tf_val = __ldg(p.st[s1].pre_tf + (size_t)t * P + tid);
      // flow control: the aux CTAs are not in the dependency loop, so do not overwrite LL slot t % depth (first written
      // by this frame's h1) before they have consumed frame t - depth
      if (t >= kLLDepth && tid == 0) {
        const unsigned need = (unsigned)kAux * (unsigned)(t - kLLDepth + 1);
        wd.arm();
        for (;;) {
          unsigned v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.aux_done) : "memory");
          if (v >= need) break;
          if (wd.expired()) { ok = false; break; }
        }
      }
    }
    // b: W_ih[:, P:] . ctx[t-1]      (a: W_hh . h1[t-1] already ran during the previous frame's attention wait)
    if constexpr (PRE) stream_finish<WB, kPreK>(pl[1], sh.xctx + s1 * E, sh.acc1, kMaxU1, warp, lane, wpre);
    else consume_step<WB>(sh, pl[1], sh.xctx + s1 * E, sh.acc1, kMaxU1, warp, lane);
    LPH(0)
    // c: W_hh(dec) . h2[t-1]   (weights requested before the exchange is polled)
    if constexpr (PRE) stream_issue<WB, kPreK>(sh, pl[2], warp, lane, wpre);
    if (t > 0) ok = poll_vector<2>(rep_h2(p, rep) + (size_t)rb_prev * H, H, tag_prev, sh.xh2, tid, wd) && ok;
    __syncthreads();
    LPH(1)
    if constexpr (PRE) stream_finish<WB, kPreK>(pl[2], sh.xh2, sh.acc2, kMaxU2, warp, lane, wpre);
    else consume_step<WB>(sh, pl[2], sh.xh2, sh.acc2, kMaxU2, warp, lane);
    LPH(2)
    // d: W_ih[:, :P] . prenet[t]
    if (p.free_running) {
      // prenet of frame t; word P of the block is the stop word the aux CTAs publish with it
      ok = poll_vector<1>(rep_pre(p, rep) + ((size_t)rb * 2 + s1) * (P + 8), P + 1, tag_cur, sh.xpre, tid, wd) && ok;
      if (!ok) *sh.exit_flag = 1;
    } else {
      if (tid < P) sh.xpre[tid] = pre_tf_val;
      if (tid == 0) sh.xpre[P] = 0.f;
    }
    __syncthreads();
    if (sh.xpre[P] != 0.f || *sh.exit_flag) break;
    LPH(3)
    consume_step<WB>(sh, pl[3], sh.xpre, sh.acc1, kMaxU1, warp, lane);
    LPH(13)
#ifdef TACO2DEC_LAT_EXPERIMENT_DOUBLE_D
    consume_step<WB>(sh, pl[3], sh.xpre, sh.acc1, kMaxU1, warp, lane);   // timing experiment only (results are wrong)
#endif
    __syncthreads();
    LPH(14)
    // attention-LSTM pointwise (gate order i,f,g,o), dropout on h and c when training
    {
      int u = 0; float hn = 0.f, cn = 0.f;
      if (lstm_pointwise(sh.acc1, kMaxU1, sh.bias1, sh.c1, nu1, tid, u, hn, cn)) {
        const int j = u1_0 + u;
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
#pragma unroll
        for (int r = 0; r < kRep; ++r) ll_store(rep_h1(p, r) + ((size_t)rb * 2 + s1) * H + j, hn, tag_cur);
      }
    }
    __syncthreads();
    // query partials: q_part[a] = sum_u Wq[a, u] h1[u]  (attention.py:56, 368), one row of the reduction tree
    if (tid < A) {
      float qv = 0.f;
      for (int u = 0; u < nu1; ++u) qv = fmaf(sh.wq_s[tid * kMaxU1 + u], sh.hloc[u], qv);
      ll_store(p.ll_q + (((size_t)rb * 2 + s1) * p.NL1 + i1) * A + tid, qv, tag_cur);
    } else if (tid - A < 2 * kMaxU1 * 4) {
      sh.acc1[tid - A] = 0.f;   // next accumulation is several barriers away
    }
    // a of the NEXT frame (W_hh . h1[t], two passes per warp): the first pass is requested before the h1 exchange is polled.
    // Behind a barrier: only four warps compute query partials, and the other warps' loads must not queue in front of
    // their LL stores (measured without it: +0.9 kcyc on the query partials, the context 1.5 kcyc later).
    if constexpr (PRE) __syncthreads();
    LPH(4)
    // Only HALF a pass (8 loads per lane, ~1 kcyc of queueing = the time the h1 words need to arrive anyway): the poll's own
    // loads queue behind whatever is requested here.
    constexpr int kHalfNU = kPreNU / 2;
    uint4 wha[4][kHalfNU > 0 ? kHalfNU : 1], whb[4][kHalfNU > 0 ? kHalfNU : 1];
    if constexpr (PRE) stream_issue_units<WB, kHalfNU>(sh, pl[0], warp, lane, 0, wha);
    // e: W_ih(dec)[:, h cols] . h1[t]   (all streams)
    ok = poll_vector<4>(rep_h1(p, rep) + (size_t)rb * 2 * H, S * H, tag_cur, sh.xh1, tid, wd) && ok;
    __syncthreads();
    LPH(5)
    if constexpr (PRE) {
      const float* xa = sh.xh1 + s1 * H;
      float aa[4][2];
#pragma unroll
      for (int g = 0; g < 4; ++g) aa[g][0] = aa[g][1] = 0.f;
      stream_issue_units<WB, kHalfNU>(sh, pl[0], warp, lane, kHalfNU, whb);
      stream_fma_units<WB, kHalfNU>(xa, lane, 0, wha, aa);
      stream_fma_units<WB, kHalfNU>(xa, lane, kHalfNU, whb, aa);
      stream_issue_units<WB, kPreNU>(sh, pl[0], warp, lane, kPreNU, wpre);
      stream_fma_units<WB, kPreNU>(xa, lane, kPreNU, wpre, aa);
      const float ca = butterfly4(aa[0][0] + aa[0][1], aa[1][0] + aa[1][1], aa[2][0] + aa[2][1], aa[3][0] + aa[3][1], lane);
      if ((lane & 7) == 0) sh.acc1[(size_t)warp * 4 + (lane >> 3)] += ca;     // ksplit = 1: unit = warp, column half 0
      LPH(9)
      consume_step<WB>(sh, pl[4], sh.xh1, sh.acc2, kMaxU2, warp, lane);
      consume_step<WB>(sh, pl[5], sh.xh1 + H, sh.acc2, kMaxU2, warp, lane);
      LPH(6)
    } else {
      consume_step<WB>(sh, pl[4], sh.xh1, sh.acc2, kMaxU2, warp, lane);
      consume_step<WB>(sh, pl[5], sh.xh1 + H, sh.acc2, kMaxU2, warp, lane);
      LPH(6)
      // a of the NEXT frame: W_hh . h1[t] needs only h1[t]; it fills the wait for the attention CTAs
      consume_step<WB>(sh, pl[0], sh.xh1 + s1 * H, sh.acc1, kMaxU1, warp, lane);
      LPH(9)
    }
    // f: W_ih(dec)[:, ctx cols] . ctx[t]
    ok = poll_vector<2>(rep_ctx(p, rep) + (size_t)rb * 2 * E, S * E, tag_cur, sh.xctx, tid, wd) && ok;
    __syncthreads();
    LPH(7)
    consume_step<WB>(sh, pl[6], sh.xctx, sh.acc2, kMaxU2, warp, lane);
    LPH(10)
    __syncthreads();
    LPH(11)
    {
      int u = 0; float hn = 0.f, cn = 0.f;
      if (lstm_pointwise(sh.acc2, kMaxU2, sh.bias2, sh.c2, nu2, tid, u, hn, cn)) {
        const int j = u2_0 + u;
        if (p.training) {
          const bool kh = p.lstm_keep ? p.lstm_keep[((size_t)t * 6 + 4) * H + j] != 0
                                      : philox_keep_l(p.seed, 8, t, j, p.thresh_dec);
          const bool kc = p.lstm_keep ? p.lstm_keep[((size_t)t * 6 + 5) * H + j] != 0
                                      : philox_keep_l(p.seed, 9, t, j, p.thresh_dec);
          hn = kh ? hn * sc_dec : 0.f;
          cn = kc ? cn * sc_dec : 0.f;
        }
        sh.c2[u] = cn;
#pragma unroll
        for (int r = 0; r < kRep; ++r) ll_store(rep_h2(p, r) + (size_t)rb * H + j, hn, tag_cur);
      }
    }
    LPH(12)
    if (!ok) *sh.exit_flag = 1;
    __syncthreads();
    if (tid < 2 * kMaxU2 * 4) sh.acc2[tid] = 0.f;
    LPH(8)
    if (*sh.exit_flag) break;
    // b of the NEXT frame: requested only after h2 is on its way -- behind the barrier, because only the first warp
    // evaluates the cells and the other warps' 16 loads per lane (~2 kcyc of queueing in the load/store unit) would get
    // in front of its h2 stores; the data arrives while this CTA would otherwise wait for the aux chain
    if constexpr (PRE) stream_issue<WB, kPreK>(sh, pl[1], warp, lane, wpre);
  }
  if (lc == 0 && tid == 0)
    for (int i = 0; i < 16; ++i) p.phase_clocks[i] = ph[i];
  if (sh.dbg_s && tid == 0)
    for (int i = 0; i < 64; ++i) p.dbg[i] = dbg_sm[i];
#undef LPH
  if (any_tmem) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(sh.tmem_base) : "memory");
  }
}

// partial energies of all positions over one CTA's attention dimensions: 4 lanes per position, APL dimensions per lane
// (one or two 16-byte, conflict-free loads of the processed-memory slice), partial sums published as LL words
template <int APL, bool LSA_>
__device__ __forceinline__ void partial_energies(const float* q_s, const float* v_s, const float* pm_s, const float* loc_s,
                                                 unsigned long long* dst, int AS, int Teff, int tid, unsigned tag) {
  const int el = tid & 3;
  float qv[APL], vv[APL];
#pragma unroll
  for (int h = 0; h < APL; h += 4) {
    const float4 q4 = *reinterpret_cast<const float4*>(q_s + el * APL + h);
    const float4 v4 = *reinterpret_cast<const float4*>(v_s + el * APL + h);
    qv[h] = q4.x; qv[h + 1] = q4.y; qv[h + 2] = q4.z; qv[h + 3] = q4.w;
    vv[h] = v4.x; vv[h + 1] = v4.y; vv[h + 2] = v4.z; vv[h + 3] = v4.w;
  }
  for (int j0 = 0; j0 < Teff; j0 += kThreads / 4) {
    if (j0 + ((tid & ~31) >> 2) >= Teff) break;          // this warp's 8 positions (and all later ones) are past the length
    const int j = j0 + (tid >> 2), jc = min(j, Teff - 1);
    const float* r = pm_s + (size_t)jc * AS + el * APL;
    const float* lc = loc_s + (size_t)jc * AS + el * APL;
    float e = 0.f;
#pragma unroll
    for (int h = 0; h < APL; h += 4) {
      float4 a4 = *reinterpret_cast<const float4*>(r + h);
      if (LSA_) {
        const float4 l4 = *reinterpret_cast<const float4*>(lc + h);
        a4.x += l4.x; a4.y += l4.y; a4.z += l4.z; a4.w += l4.w;
      }
      e = fmaf(vv[h], fast_tanh(qv[h] + a4.x), e);
      e = fmaf(vv[h + 1], fast_tanh(qv[h + 1] + a4.y), e);
      e = fmaf(vv[h + 2], fast_tanh(qv[h + 2] + a4.z), e);
      e = fmaf(vv[h + 3], fast_tanh(qv[h + 3] + a4.w), e);
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    if (el == 0 && j < Teff) ll_store(dst + j, e, tag);
  }
}

// ---------------------------------------------------------------------------------------------
// attention CTA: stream s, slice g of na  (attention.py:330-398)
//
// The na CTAs of a stream split the work twice:
//   * energies by ATTENTION DIMENSION: CTA g owns a in [g*AS, (g+1)*AS), AS = A / na.  It gathers only that slice of the query
//     partials (NL1 * AS words instead of NL1 * A), evaluates  e~_j = sum_{a in slice} v_a tanh(q_a + pm_ja)  for every position
//     (A * Ts / na tanh instead of A * Ts: the energies were SFU-bound) and publishes the partial energies; every CTA then sums
//     the na partials of each position in a fixed order (one more LL exchange, ~1 L2 round trip, for 4x-8x less work on the chain);
//   * the context by ENCODER FEATURE as before: CTA g owns features [g*FS, (g+1)*FS).
//
// Location-sensitive attention (attention.py:25-85, model.py:355-359) uses the same split.  Its location term
//   loc_j[a] = sum_f Wd[a][f] . conv1d([alpha[t-1]; cum[t-1]])_j[f]
// depends only on the PREVIOUS frame's weights, so every CTA computes it for its own attention dimensions right after it has
// published the context of frame t-1 -- off the critical path, while the LSTM CTAs are busy -- and the chain from the query
// to the context only gains the softmax (two block reductions) over the stepwise-monotonic variant.
// ---------------------------------------------------------------------------------------------
constexpr int kLatTsCap = 512;          // positions per stream the partial-energy exchange is sized for
__device__ void attention_cta(const LatParams& p, int s, int g, unsigned char* smem_raw) {
  const LatStream& sp = p.st[s];
  const int tid = threadIdx.x, lane = tid & 31;
  const int Ts = sp.Ts, na = sp.na;
  const int Teff = sp.len ? min((int)sp.len[0], Ts) : Ts;
  const int FS = E / na;                    // context features owned by this CTA (64 or 128)
  const int AS = A / na;                    // attention dimensions owned by this CTA (16 or 32)
  const int NJ = kThreads / FS;             // position groups in the context reduction
  const int NL1 = p.NL1;
  const int nq = NL1 * AS;                  // query-partial words this CTA gathers
  float* sm = reinterpret_cast<float*>(smem_raw);
  float* pm_s = sm;                         // [Ts][AS]   own slice of the processed memory
  float* mem_s = pm_s + (size_t)Ts * AS;    // [Ts][FS]
  float* qp_s = mem_s + (size_t)Ts * FS;    // [NL1][AS] gathered partials
  float* q8_s = qp_s + (size_t)kThreads * 4;   // [8][AS]
  float* q_s = q8_s + 8 * 32;               // [AS]
  float* v_s = q_s + 32;                    // [AS]
  float* red_s = v_s + 32;                  // [NJ][FS] = kThreads
  // location-sensitive attention only (16-byte aligned blocks before the T-dependent scalar arrays)
  const bool lsa = p.lsa != 0;
  const int LF = p.LF, LK = p.LK, pad = (LK - 1) / 2, Tp = (Ts + 2 * pad + 2 + 3) & ~3;
  float* loc_s = red_s + kThreads;          // [Ts][AS]  location term of the coming frame, own attention dimensions
  float* wf_s = loc_s + (lsa ? (size_t)Ts * AS : 0);   // [2][LK][AS]  location conv and dense folded into one map (both bias-free linear)
  float* ap_s = wf_s + (lsa ? 2 * LK * AS : 0);        // [Tp] zero-padded previous weights
  float* ac_s = ap_s + (lsa ? Tp : 0);      // [Tp] zero-padded cumulative weights
  float* al_s = ac_s + (lsa ? Tp : 0);      // [Ts] alignment state
  float* pr_s = al_s + Ts;                  // [Ts + 4] probabilities
  float* an_s = pr_s + Ts + 4;              // [Ts] new alignment
  __shared__ int s_stop;
  __shared__ float s_red[kWarps];

  for (int i = tid; i < Ts * AS; i += kThreads) {
    const int j = i / AS, a = i - j * AS;
    pm_s[i] = sp.pm[(size_t)j * A + g * AS + a];
  }
  for (int i = tid; i < Ts * FS; i += kThreads) {
    const int j = i / FS, f = i - j * FS;
    mem_s[i] = sp.mem[(size_t)j * E + g * FS + f];
  }
  for (int i = tid; i < AS; i += kThreads) v_s[i] = sp.v[g * AS + i];
  for (int i = tid; i < Ts; i += kThreads) al_s[i] = i == 0 ? 1.0f : 0.0f;   // attention.py:324-328
  if (lsa) {
    // loc_j[a] = sum_f Wd[a][f] sum_{c,k} Wc[f][c][k] x_c[j+k-pad]  =  sum_{c,k} Wf[c][k][a] x_c[j+k-pad]   (attention.py:12-23)
    for (int i = tid; i < 2 * LK * AS; i += kThreads) {
      const int a = i % AS, ck = i / AS;
      float acc = 0.f;
      for (int f = 0; f < LF; ++f) acc = fmaf(sp.loc_dense[(size_t)(g * AS + a) * LF + f], sp.loc_conv[(size_t)f * 2 * LK + ck], acc);
      wf_s[i] = acc;
    }
    for (int i = tid; i < Tp; i += kThreads) { ap_s[i] = 0.f; ac_s[i] = 0.f; }      // model.py:277-280: both start at zero
    for (int i = tid; i < Ts * AS; i += kThreads) loc_s[i] = 0.f;                   // hence so does the location term
  }
  if (tid == 0) s_stop = 0;
  __syncthreads();

  Watch wd{p.abort_flag, 0, 0};
  __shared__ long long aph[16];
  if (tid < 16) aph[tid] = 0;
  __syncthreads();
  const bool adbg = p.dbg && g == 0 && tid == 0;
  long long aph_t = clock64();
#define APH(slot) if (adbg) { const long long n_ = clock64(); aph[slot] += n_ - aph_t; aph_t = n_; }
  // energies: 4 lanes per position, APL attention dimensions per lane
  const int APL = AS / 4;
  const int el = tid & 3;
  for (int t = 0; t < p.n_steps; ++t) {
    const unsigned tag = (unsigned)t + 1u;
    const int rb = t % kLLDepth;
    if (p.free_running) {
      if (tid == 0) {
        float sv = 0.f;
        const bool got = ll_wait(rep_pre(p, (g + 3 * s) % kRep) + ((size_t)rb * 2 + s) * (P + 8) + P, tag, sv, wd);
        if (!got || sv != 0.f) s_stop = 1;
      }
      __syncthreads();
      if (s_stop) break;
    }
    APH(0)
    // ---- own slice of the query partials: word i = (LSTM CTA c, a') with a' fastest; <= 4 words per thread ----
    {
      const unsigned long long* src = p.ll_q + ((size_t)rb * 2 + s) * NL1 * A + g * AS;
      unsigned pending = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (tid + k * kThreads < nq) pending |= 1u << k;
      bool good = true;
      wd.arm();
      while (pending) {
        unsigned long long w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (pending & (1u << k)) {
            const int i = tid + k * kThreads, c = i / AS, a = i - c * AS;
            w[k] = ll_load(src + (size_t)c * A + a);
          }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((pending & (1u << k)) && (unsigned)(w[k] >> 32) == tag) {
            qp_s[tid + k * kThreads] = __uint_as_float((unsigned)w[k]);
            pending &= ~(1u << k);
          }
        if (pending && wd.expired()) { good = false; break; }
      }
      if (!good) s_stop = 1;
    }
    APH(1)
    __syncthreads();
    APH(2)
    if (s_stop) break;
    if (tid < 8 * AS) {                      // fixed-order tree: 8 groups of LSTM CTAs, then the 8 group sums
      const int part = tid / AS, a = tid - part * AS;
      float acc = 0.f;
      for (int c = part; c < NL1; c += 8) acc += qp_s[c * AS + a];
      q8_s[part * 32 + a] = acc;
    }
    __syncthreads();
    if (tid < AS) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += q8_s[k * 32 + tid];
      q_s[tid] = acc;
    }
    __syncthreads();
    APH(3)
    // ---- partial energies of every position over the own attention dimensions ----
    // (compile-time APL / lsa variants: the generic body with run-time predicates was 156 instructions per pass and the
    //  16 warps were ISSUE-bound on it -- 0.78 kcyc per 128 positions; warps whose 8 positions are all past the length skip)
    {
      unsigned long long* dst = p.ll_e + (((size_t)rb * 2 + s) * 8 + g) * kLatTsCap;
      if (APL == 4) {
        if (lsa) partial_energies<4, true>(q_s, v_s, pm_s, loc_s, dst, AS, Teff, tid, tag);
        else partial_energies<4, false>(q_s, v_s, pm_s, loc_s, dst, AS, Teff, tid, tag);
      } else if (APL == 8) {
        if (lsa) partial_energies<8, true>(q_s, v_s, pm_s, loc_s, dst, AS, Teff, tid, tag);
        else partial_energies<8, false>(q_s, v_s, pm_s, loc_s, dst, AS, Teff, tid, tag);
      } else {
        float qv[8], vv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          qv[i] = i < APL ? q_s[el * APL + i] : 0.f;
          vv[i] = i < APL ? v_s[el * APL + i] : 0.f;
        }
        for (int j0 = 0; j0 < Teff; j0 += kThreads / 4) {
          const int j = j0 + (tid >> 2), jc = min(j, Teff - 1);
          const float* r = pm_s + (size_t)jc * AS + el * APL;
          const float* lc = loc_s + (size_t)jc * AS + el * APL;
          float e = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < APL) e = fmaf(vv[i], fast_tanh(qv[i] + r[i] + (lsa ? lc[i] : 0.f)), e);
          e += __shfl_xor_sync(0xffffffffu, e, 1);
          e += __shfl_xor_sync(0xffffffffu, e, 2);
          if (el == 0 && j < Teff) ll_store(dst + j, e, tag);
        }
      }
    }
    APH(9)
    // ---- e_j = sum of the na partials (fixed order), p_j = sigmoid(e_j [+ 2 N(0,1)]) ----
    {
      const unsigned long long* src = p.ll_e + ((size_t)rb * 2 + s) * 8 * kLatTsCap;
      bool good = true;
      for (int j = tid; j < Ts; j += kThreads) {
        float pr = 0.f;                                           // sigmoid(-inf) beyond the length, attention.py:388-391
        if (j < Teff) {
          unsigned pending = (1u << na) - 1u;
          float val[8];
          wd.arm();
          while (pending) {
            unsigned long long w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (pending & (1u << k)) w[k] = ll_load(src + (size_t)k * kLatTsCap + j);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if ((pending & (1u << k)) && (unsigned)(w[k] >> 32) == tag) {
                val[k] = __uint_as_float((unsigned)w[k]);
                pending &= ~(1u << k);
              }
            if (pending && wd.expired()) { good = false; break; }
          }
          float ev = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < na) ev += val[k];
          if (lsa) {
            pr = ev;
          } else {
            if (p.training) {
              const float nz = sp.noise ? sp.noise[(size_t)t * Ts + j] : philox_normal(p.seed, 10 + s, t, j);
              ev += 2.0f * nz;
            }
            pr = sigmoid_acc(ev);
          }
        } else if (lsa) {
          pr = -INFINITY;                                         // masked_fill_(mask, -inf), attention.py:79
        }
        pr_s[j] = pr;
      }
      if (!good) s_stop = 1;
    }
    APH(10)
    __syncthreads();
    APH(4)
    if (s_stop) break;
    if (lsa) {
      // ---- alpha = softmax(e) over the positions (attention.py:81); every thread owns at most one position ----
      static_assert(kLatTsCap <= kThreads, "one position per thread in the softmax");
      const int warp = tid >> 5;
      const float ev = tid < Ts ? pr_s[tid] : -INFINITY;
      float m = ev;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) s_red[warp] = m;
      __syncthreads();
      m = s_red[0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) m = fmaxf(m, s_red[w]);
      __syncthreads();
      const float ex = tid < Ts ? expf(ev - m) : 0.f;
      float sum = warp_sum(ex);
      if (lane == 0) s_red[warp] = sum;
      __syncthreads();
      sum = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) sum += s_red[w];
      if (tid < Ts) {
        const float a = ex / sum;
        an_s[tid] = a;
        ap_s[pad + tid] = a;
        ac_s[pad + tid] += a;                                     // model.py:358-359
        if (g == 0) sp.align[(size_t)t * Ts + tid] = a;
      }
    } else {
      // ---- alpha'_j = alpha_j p_j + alpha_{j-1} (1 - p_{j-1}) -------------------------------
      for (int j = tid; j < Ts; j += kThreads) {
        float a = al_s[j] * pr_s[j];
        if (j > 0) a += al_s[j - 1] * (1.0f - pr_s[j - 1]);
        if (p.free_running && j >= Teff) a = 0.f;
        an_s[j] = a;
        if (g == 0) sp.align[(size_t)t * Ts + j] = a;
      }
    }
    __syncthreads();
    APH(5)
    // ---- context slice: NJ position groups x FS features, 4 independent chains per thread ----
    {
      const int f = tid % FS, jg = tid / FS;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int j = jg;
      for (; j + 3 * NJ < Ts; j += 4 * NJ) {
        a0 = fmaf(an_s[j], mem_s[(size_t)j * FS + f], a0);
        a1 = fmaf(an_s[j + NJ], mem_s[(size_t)(j + NJ) * FS + f], a1);
        a2 = fmaf(an_s[j + 2 * NJ], mem_s[(size_t)(j + 2 * NJ) * FS + f], a2);
        a3 = fmaf(an_s[j + 3 * NJ], mem_s[(size_t)(j + 3 * NJ) * FS + f], a3);
      }
      for (; j < Ts; j += NJ) a0 = fmaf(an_s[j], mem_s[(size_t)j * FS + f], a0);
      red_s[jg * FS + f] = (a0 + a1) + (a2 + a3);
    }
    for (int j = tid; j < Ts; j += kThreads) al_s[j] = an_s[j];
    __syncthreads();
    APH(6)
    if (tid < FS) {
      float c = 0.f;
      for (int k = 0; k < NJ; ++k) c += red_s[k * FS + tid];
#pragma unroll
      for (int r = 0; r < kRep; ++r) ll_store(rep_ctx(p, r) + ((size_t)rb * 2 + s) * E + g * FS + tid, c, tag);
    }
    APH(7)
    if (lsa) {
      // ---- off the critical path: location term of frame t+1 for the own attention dimensions from the weights just
      //      computed; a thread does 2 consecutive positions x 4 dimensions, the windows slide through registers ----
      const int A4 = AS / 4, n_items = A4 * ((Teff + 1) / 2);
      for (int it = tid; it < n_items; it += kThreads) {
        const int a4 = it % A4, j0 = 2 * (it / A4);
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float x0 = ap_s[j0], c0 = ac_s[j0];
        for (int k = 0; k < LK; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(wf_s + (size_t)k * AS + 4 * a4);
          const float4 w1 = *reinterpret_cast<const float4*>(wf_s + (size_t)(LK + k) * AS + 4 * a4);
          const float x1 = ap_s[j0 + k + 1], c1 = ac_s[j0 + k + 1];
          acc[0][0] = fmaf(w0.x, x0, acc[0][0]); acc[0][1] = fmaf(w0.y, x0, acc[0][1]);
          acc[0][2] = fmaf(w0.z, x0, acc[0][2]); acc[0][3] = fmaf(w0.w, x0, acc[0][3]);
          acc[1][0] = fmaf(w0.x, x1, acc[1][0]); acc[1][1] = fmaf(w0.y, x1, acc[1][1]);
          acc[1][2] = fmaf(w0.z, x1, acc[1][2]); acc[1][3] = fmaf(w0.w, x1, acc[1][3]);
          acc[0][0] = fmaf(w1.x, c0, acc[0][0]); acc[0][1] = fmaf(w1.y, c0, acc[0][1]);
          acc[0][2] = fmaf(w1.z, c0, acc[0][2]); acc[0][3] = fmaf(w1.w, c0, acc[0][3]);
          acc[1][0] = fmaf(w1.x, c1, acc[1][0]); acc[1][1] = fmaf(w1.y, c1, acc[1][1]);
          acc[1][2] = fmaf(w1.z, c1, acc[1][2]); acc[1][3] = fmaf(w1.w, c1, acc[1][3]);
          x0 = x1; c0 = c1;
        }
        *reinterpret_cast<float4*>(loc_s + (size_t)j0 * AS + 4 * a4) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
        if (j0 + 1 < Ts)
          *reinterpret_cast<float4*>(loc_s + (size_t)(j0 + 1) * AS + 4 * a4) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
      }
    }
    __syncthreads();
    APH(8)
  }
#undef APH
  if (adbg)
    for (int i = 0; i < 16; ++i) p.dbg[64 + 16 * s + i] = aph[i];
  (void)lane;
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
  constexpr int PR = P / kAux;                        // prenet layer-1 rows per aux CTA per stream (32)
  constexpr int LW = 2 * P + 8;                       // words per aux block of the layer-0 partial exchange
  const int r0 = x * RPX, r1 = min(M + 1, r0 + RPX), nr = r1 - r0;
  const int nmel = min(M, r1) - r0;                   // mel rows among them (the last CTA also owns the gate row)
  float* sm = reinterpret_cast<float*>(smem_raw);
  float* wp_s = sm;                                   // [RPX][KD]           projection / gate rows
  float* w0c_s = wp_s + (size_t)RPX * (H + 2 * E);    // [S*P][RPX]          prenet layer-0 COLUMNS of the own mel rows
  float* w1_s = w0c_s + 2 * P * RPX;                  // [S][PR][P]          prenet layer-1 rows
  float* y_s = w1_s + 2 * PR * P;                     // [KD]
  float* own_s = y_s + (H + 2 * E);                   // [16] own mel values
  float* l0_s = own_s + 16;                           // [S][P]
  __shared__ int s_stop, s_stopval;

  for (int i = tid; i < nr * KD; i += kThreads) {
    const int r = i / KD, k = i - r * KD;
    const int row = r0 + r;
    wp_s[(size_t)r * KD + k] = row < M ? p.proj_w[(size_t)row * KD + k] : p.gate_w[k];
  }
  for (int i = tid; i < S * P * RPX; i += kThreads) {
    const int si = i / RPX, r = i - si * RPX, s = si / P, o = si - s * P;
    w0c_s[i] = r < nmel ? p.st[s].pre_w0[(size_t)o * M + r0 + r] : 0.f;
  }
  for (int i = tid; i < S * PR * P; i += kThreads) {
    const int s = i / (PR * P), r = (i - s * PR * P) / P, k = i - s * PR * P - r * P;
    w1_s[i] = p.st[s].pre_w1[(size_t)(x * PR + r) * P + k];
  }
  if (tid == 0) { s_stop = 0; s_stopval = 0; }
  if (tid < 16) own_s[tid] = 0.f;
  const float row_bias = warp < nr ? (r0 + warp < M ? p.proj_b[r0 + warp] : p.gate_b[0]) : 0.f;
  static_assert(2 * PR <= kWarps * 4, "prenet layer 1: one pass of 4 rows per warp");
  __syncthreads();

  Watch wd{p.abort_flag, 0, 0};
  __shared__ long long xph[16];
  if (tid < 16) xph[tid] = 0;
  __syncthreads();
  const bool xdbg = p.dbg && x == 0 && tid == 0;
  long long xph_t = clock64();
#define XPH(slot) if (xdbg) { const long long n_ = clock64(); xph[slot] += n_ - xph_t; xph_t = n_; }
  // prenet of frame 0 = prenet(go-frame of zeros) = zeros (model.py:444-450); stop word = 0
  if (p.free_running) {
    for (int i = tid; i < S * PR * kRep; i += kThreads) {
      const int rr = i / (S * PR), k = i - rr * S * PR, s = k / PR, r = k - s * PR;
      ll_store(rep_pre(p, rr) + ((size_t)0 * 2 + s) * (P + 8) + x * PR + r, 0.f, 1u);
    }
    if (x == 0 && tid < S * kRep) ll_store(rep_pre(p, tid / S) + ((size_t)0 * 2 + tid % S) * (P + 8) + P, 0.f, 1u);
  }

  for (int t = 0; t < p.n_steps; ++t) {
    const unsigned tag = (unsigned)t + 1u;
    const int rb = t % kLLDepth, rbn = (t + 1) % kLLDepth;
    // ---- y = [h2_t | ctx_t | ctx_bert_t].  The context is published long before h2: its share of the projection rows and
    //      the dropout bits of the next prenet are finished while the decoder LSTM is still running, so that only the
    //      h2 half of each row (fully unrolled, every load in flight at once) sits between h2 and the prenet ----------
    XPH(0)
    bool ok = poll_vector<2>(rep_ctx(p, x % kRep) + (size_t)rb * 2 * E, S * E, tag, y_s + H, tid, wd);
    __syncthreads();
    float pc = 0.f;                     // context part of this warp's row (all lanes)
    if (warp < nr) {
      const float* w = wp_s + (size_t)warp * KD + H;
      float a0 = 0.f, a1 = 0.f;
      for (int k = lane * 4; k < S * E; k += 128) {
        const float4 a = *reinterpret_cast<const float4*>(w + k);
        const float4 b = *reinterpret_cast<const float4*>(y_s + H + k);
        a0 = fmaf(a.x, b.x, a0); a1 = fmaf(a.y, b.y, a1); a0 = fmaf(a.z, b.z, a0); a1 = fmaf(a.w, b.w, a1);
      }
      pc = warp_sum(a0 + a1) + row_bias;
    }
    // dropout bits of the layer-1 rows this lane will publish (independent of the data)
    bool kp1 = false;
    {
      const int it = warp * 4 + (lane >> 3);
      if (p.free_running && it < S * PR) {
        const int s = it / PR, r = it - s * PR, row = x * PR + r;
        const uint8_t* keep = p.st[s].keep1;
        kp1 = keep ? keep[(size_t)(t + 1) * P + row] != 0 : philox_keep_l(p.seed, s * 2 + 1, t + 1, row, p.thresh_pre);
      }
    }
    XPH(1)
    ok = poll_vector<2>(rep_h2(p, x % kRep) + (size_t)rb * H, H, tag, y_s, tid, wd) && ok;
    if (!ok) s_stop = 1;
    XPH(2)
    __syncthreads();
    XPH(3)
    if (s_stop) break;
    // ---- projection rows (one warp per row), h2 half -----------------------------------------
    if (warp < nr) {
      const float* w = wp_s + (size_t)warp * KD;
      float4 wa[H / 128], ya[H / 128];
#pragma unroll
      for (int i = 0; i < H / 128; ++i) {
        wa[i] = *reinterpret_cast<const float4*>(w + lane * 4 + 128 * i);
        ya[i] = *reinterpret_cast<const float4*>(y_s + lane * 4 + 128 * i);
      }
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < H / 128; ++i) {
        a0 = fmaf(wa[i].x, ya[i].x, a0); a1 = fmaf(wa[i].y, ya[i].y, a1);
        a2 = fmaf(wa[i].z, ya[i].z, a2); a3 = fmaf(wa[i].w, ya[i].w, a3);
      }
      const float acc = warp_sum((a0 + a1) + (a2 + a3)) + pc;
      if (lane == 0) {
        const int row = r0 + warp;
        if (row < M) {
          p.mel[(size_t)t * M + row] = acc;
          own_s[warp] = acc;
        } else {
          p.gate[t] = acc;
          if (p.free_running) {
            int stop = 0;
            if (sigmoid_acc(acc) > p.gate_thr) { stop = 1; p.n_frames[0] = t + 1; }            // model.py:480-481
            else if (t + 1 == p.n_steps) { stop = 1; p.n_frames[0] = t + 1; p.reached_max[0] = 1; }  // :482-485
            s_stopval = stop;
          }
        }
      }
    }
    XPH(4)
    __syncthreads();
    XPH(5)
    if (!p.free_running) {
      // nobody waits for the projection when teacher forcing: report progress for LL flow control
      if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.aux_done), "r"(1u) : "memory");
      continue;
    }
    // ---- prenet layer 0, PARTIAL over the own mel rows: W0[:, rows] . mel[rows]  (saves the mel exchange) ----
    {
      unsigned long long* blk = p.ll_l0 + ((size_t)rbn * kAux + x) * LW;
      if (tid < S * P) {
        const float* w = w0c_s + (size_t)tid * RPX;
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < RPX; ++r) acc = fmaf(w[r], own_s[r], acc);
        ll_store(blk + tid, acc, tag + 1u);
      }
      if (tid == 0) ll_store(blk + 2 * P, s_stopval ? 1.f : 0.f, tag + 1u);   // meaningful in the gate owner's block
    }
    XPH(6)
    // ---- gather the kAux partials (fixed order), relu + dropout(0.5) -> layer-0 activations -------------
    {
      const unsigned long long* base = p.ll_l0 + (size_t)rbn * kAux * LW;
      bool good = true;
      float sum = 0.f;
      if (tid < S * P) {
        // dropout bit first: it does not depend on the partials, so it is off the chain
        const int s = tid / P, o = tid - s * P;
        const uint8_t* keep = p.st[s].keep0;
        const bool kp = keep ? keep[(size_t)(t + 1) * P + o] != 0 : philox_keep_l(p.seed, s * 2 + 0, t + 1, o, p.thresh_pre);
        unsigned pending = (1u << kAux) - 1u;
        float val[kAux];
        wd.arm();
        while (pending) {
          unsigned long long w[kAux];
#pragma unroll
          for (int k = 0; k < kAux; ++k)
            if (pending & (1u << k)) w[k] = ll_load(base + (size_t)k * LW + tid);
#pragma unroll
          for (int k = 0; k < kAux; ++k)
            if ((pending & (1u << k)) && (unsigned)(w[k] >> 32) == tag + 1u) {
              val[k] = __uint_as_float((unsigned)w[k]);
              pending &= ~(1u << k);
            }
          if (pending && wd.expired()) { good = false; break; }
        }
#pragma unroll
        for (int k = 0; k < kAux; ++k) sum += val[k];
        l0_s[tid] = kp ? fmaxf(sum, 0.f) * 2.0f : 0.f;
      }
      if (tid == 0) {
        float sv = 0.f;
        good = ll_wait(base + (size_t)(kAux - 1) * LW + 2 * P, tag + 1u, sv, wd) && good;
        s_stopval = sv != 0.f;
      }
      if (!good) s_stop = 1;
    }
    XPH(7)
    __syncthreads();
    XPH(8)
    if (s_stop) break;
    const bool stop = s_stopval != 0;
    if (x == 0 && tid < S * kRep)   // the stop word rides in every replica / stream block of the next frame's prenet
      ll_store(rep_pre(p, tid / S) + ((size_t)rbn * 2 + tid % S) * (P + 8) + P, stop ? 1.f : 0.f, tag + 1u);
    if (stop) {
      // release everybody polling the next frame's prenet block (they read the stop word with it)
      for (int i = tid; i < S * PR * kRep; i += kThreads) {
        const int rr = i / (S * PR), k = i - rr * S * PR, s = k / PR, r = k - s * PR;
        ll_store(rep_pre(p, rr) + ((size_t)rbn * 2 + s) * (P + 8) + x * PR + r, 0.f, tag + 1u);
      }
      break;
    }
    // ---- prenet layer 1 rows of this CTA: a warp does 4 rows at once, the butterfly leaves row g's sum in lane group g
    //      (8 lanes), and those 8 lanes publish it to the 8 replicas in parallel ---------------------------------------
    static_assert(kRep == 8, "one lane of a butterfly group per replica");
    {
      const int it0 = warp * 4;
      if (it0 < S * PR) {
        const int s0 = it0 / PR;                         // the 4 rows of a warp belong to one stream (PR % 4 == 0)
        const float* in = l0_s + (size_t)s0 * P;
        const float* w = w1_s + ((size_t)s0 * PR + (it0 - s0 * PR)) * P;
        float4 xv[P / 128], wv[4][P / 128];
#pragma unroll
        for (int i = 0; i < P / 128; ++i) {
          xv[i] = *reinterpret_cast<const float4*>(in + lane * 4 + 128 * i);
#pragma unroll
          for (int q = 0; q < 4; ++q) wv[q][i] = *reinterpret_cast<const float4*>(w + (size_t)q * P + lane * 4 + 128 * i);
        }
        float acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float e0 = 0.f, e1 = 0.f;
#pragma unroll
          for (int i = 0; i < P / 128; ++i) {
            e0 = fmaf(wv[q][i].x, xv[i].x, e0); e1 = fmaf(wv[q][i].y, xv[i].y, e1);
            e0 = fmaf(wv[q][i].z, xv[i].z, e0); e1 = fmaf(wv[q][i].w, xv[i].w, e1);
          }
          acc[q] = e0 + e1;
        }
        const float v = butterfly4(acc[0], acc[1], acc[2], acc[3], lane);
        const int it = it0 + (lane >> 3), rr = lane & 7;
        const int r = it - s0 * PR, row = x * PR + r;
        ll_store(rep_pre(p, rr) + ((size_t)rbn * 2 + s0) * (P + 8) + row, kp1 ? fmaxf(v, 0.f) * 2.0f : 0.f, tag + 1u);
      }
    }
    XPH(9)
    __syncthreads();
    XPH(10)
  }
#undef XPH
  if (xdbg)
    for (int i = 0; i < 16; ++i) p.dbg[96 + i] = xph[i];
}

template <int WB, bool PRE>
__global__ void __launch_bounds__(kThreads, 1) decoder_latency(const __grid_constant__ LatParams p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  const int b = blockIdx.x;
  const int na = p.st[0].na + (p.S == 2 ? p.st[1].na : 0);
  if (b < p.NL) {
    lstm_cta<WB, PRE>(p, b, dyn_smem);
  } else if (b < p.NL + na) {
    const int k = b - p.NL;
    if (k < p.st[0].na) attention_cta(p, 0, k, dyn_smem);
    else attention_cta(p, 1, k - p.st[0].na, dyn_smem);
  } else if (b < p.NL + na + kAux) {
    aux_cta(p, b - p.NL - na, dyn_smem);
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
