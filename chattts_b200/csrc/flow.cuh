// Dataflow decode step for small batches (B <= 4):  k_flow<BT>
//
// Same arithmetic as k_step (mega.cuh) - input -> 20 x [QKV+RoPE+KV append -> attention -> O-proj -> gate/up -> down]
// -> heads - in ONE persistent cooperative kernel, but the two things that kept k_step at a quarter of the HBM
// roofline are gone:
//
//  * Weight stream decoupled from the phases.  Every warp owns a private ring of FL_SLOTS x 6 KiB shared-memory
//    slots filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  The warp's tasks for the whole step
//    (QKV pair, K/V chunk, O row, gate/up pairs, down slices, head pairs; all layers) form one static sequence;
//    after consuming task n the warp's lane 0 issues the copy of task n + FL_SLOTS into the slot it just freed.
//    Weights (and the K/V of earlier tokens) never depend on this step's activations, so 148 x 192 KiB are always
//    in flight across phase and layer boundaries and HBM streams while a CTA waits for its inputs.
//
//  * No grid barriers.  Activations cross CTAs as 8-byte {value, tag} words ("LL" protocol): the producer stores
//    value and tag with ONE 64-bit store, the consumer polls the data words themselves until the tag of this
//    (step, layer, phase) appears.  One L2 write + one L2 read per dependent edge - no release fence, no arrival
//    counter, no separate data load after an acquire.  Broadcast vectors are written to R replicas so that the
//    148 readers of a vector do not queue on the same L2 slices.
//
// Arithmetic order per row is independent of the batch; results differ from k_step only by fp32 reassociation in
// the RMSNorm sum (per-lane strided instead of per-warp-row).
#pragma once
#include "gpt_kernels.cuh"
#include "tc_common.cuh"

namespace ctb {

constexpr int FL_THREADS = 256;
constexpr int FL_WARPS = 8;
constexpr int FL_SLOTS = 4;
constexpr int FL_SLOT_BYTES = 6144;
constexpr int FL_SLOT_FLOATS = FL_SLOT_BYTES / 4;
constexpr int FL_RING_BYTES = FL_WARPS * FL_SLOTS * FL_SLOT_BYTES;  // 192 KiB
constexpr int FL_SMAX = 12;  // attention splits per (row, head)
constexpr int FL_CH = 64;    // keys per attention chunk: 8 per warp
constexpr int FL_PW = 66;    // words of one attention partial: o[64], m, l
constexpr int FL_BMAX = 4;   // batch rows the exchange arena is sized for
constexpr int FL_ROWS = 6;   // O-proj / down rows per CTA: ceil(768 / grid) for grid >= 128
constexpr int FL_GU = 3;     // gate/up pair tasks per warp: ceil(3072 / (8 * grid)) for grid >= 128
constexpr int FL_QR = 2;     // QKV pair tasks per warp
constexpr int FL_HEADS = 12; // heads the arena is sized for
constexpr int FL_RMAX = 16;  // replicas the arena is sized for

// Exchange arena, in 8-byte words.  Two parity copies (layer l uses copy l & 1); per copy FL_RMAX replicas of the
// broadcast regions followed by the point-to-point q / k_new / v_new region.
constexpr int FL_A_X = 0;                                       // [BMAX][768]  residual stream entering a layer
constexpr int FL_A_XO = FL_A_X + FL_BMAX * KC;                  // [BMAX][768]  residual stream after O-proj
constexpr int FL_A_ACT = FL_A_XO + FL_BMAX * KC;                // [BMAX][3072] silu(gate) * up
constexpr int FL_A_P = FL_A_ACT + FL_BMAX * 4 * KC;             // [BMAX][12][SMAX][66] attention partials
constexpr int FL_A_END = FL_A_P + FL_BMAX * FL_HEADS * FL_SMAX * FL_PW;
constexpr int FL_REP_STRIDE = ((FL_A_END + 1023) / 1024) * 1024;
constexpr int FL_A_Q = 0, FL_A_KN = FL_BMAX * KC, FL_A_VN = 2 * FL_BMAX * KC;
constexpr int FL_QKV_WORDS = 3 * FL_BMAX * KC;
constexpr size_t FL_PARITY_WORDS = (size_t)FL_RMAX * FL_REP_STRIDE + FL_QKV_WORDS;
constexpr size_t FL_ARENA_WORDS = 2 * FL_PARITY_WORDS;
constexpr unsigned FL_EPOCH_STEP = 256;  // tags of one launch: base + 8 * layer + kind

enum FlowTagKind { FT_X = 0, FT_QKV = 1, FT_P = 2, FT_XO = 3, FT_ACT = 4 };
enum FlowStage { FS_Q0 = 0, FS_Q1 = 1, FS_KV = 2, FS_O = 3, FS_GU0 = 4, FS_GU1 = 5, FS_GU2 = 6, FS_D0 = 7, FS_D1 = 8, FS_NLAYER = 9 };

struct FlowP {
  const float* W;  // packed fp32 blob
  int64_t layer0, layer_stride, o_wqkv, o_wo, o_wgu, o_wd, o_ln1, o_ln2, o_final_norm, o_head, o_emb_code, o_emb_text,
      o_cos, o_sin;
  int L, I, Hq, hd;
  float eps, scaling;
  float *logits, *kv;
  size_t kv_layer_floats;
  const int* block_table; int pages_per_row;
  int* seq_len;
  LoopState* st;
  int decode, col, T0, sample;
  const float* emb; const uint8_t* mask; const int32_t* ids_out;
  int max_new, num_vq, num_audio, infer_text, B;
  float* hidden_out; int hidden_stride, rows_per_item, V;
  unsigned long long* arena;  // FL_ARENA_WORDS, zeroed at create
  unsigned* epoch;            // tag base of the next launch (advanced by CTA 0 at the end of every launch)
  int R;                      // replicas in use (1..FL_RMAX)
  unsigned long long* trace;  // optional globaltimer stamps of CTA 0 (1 + 5 * L + 1)
};

// ---------------------------------------------------------------- LL words
__device__ __forceinline__ void ll_st(unsigned long long* p, float v, uint32_t tag) {
  const unsigned long long x = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory");
}
__device__ __forceinline__ unsigned long long ll_ld(const unsigned long long* p) {
  unsigned long long x;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
  return x;
}
__device__ __forceinline__ void ll_ld2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ll_tag(unsigned long long x) { return (uint32_t)(x >> 32); }
__device__ __forceinline__ float ll_val(unsigned long long x) { return __uint_as_float((uint32_t)x); }

// Watchdog: a wait that does not complete in ~2^20 polls records an error in LoopState::err and every later wait of
// the thread falls through, so a protocol bug ends the kernel in about a second instead of hanging the device.
struct FlowWd {
  int* err;
  int dead;
  int spins;
};
__device__ __forceinline__ bool fl_giveup(FlowWd& wd, int code) {
  if (wd.dead) return true;
  if (++wd.spins < (1 << 20)) {
    if ((wd.spins & 8191) == 0 && __ldcg(wd.err) != 0) { wd.dead = 1; return true; }
    return false;
  }
  atomicCAS(wd.err, 0, code);
  wd.dead = 1;
  return true;
}

__device__ __forceinline__ void fl_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void fl_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool fl_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// Geometry every thread of a CTA agrees on (registers; uniform per warp).
struct FlowGeo {
  int G, NW, gw, cta, warp, lane;
  int S;                                // attention splits per (row, head)
  int u_on, u_b, u_h, u_split, u_n, u_nchunk;  // this CTA's attention unit (u_on = 0: none / masked row)
  int nheads_tasks;                     // 2-row tasks of the heads phase (0: no heads in this launch)
};

// ---- task predicates shared by the issue side and the consume side (they MUST enumerate the same sequence)
__device__ __forceinline__ bool fl_q_valid(const FlowGeo& g, int j) { return g.gw + j * g.NW < 3 * KC / 2; }
__device__ __forceinline__ bool fl_kv_more(const FlowGeo& g, int sub, int& chunk) {
  chunk = g.u_split + sub * g.S;
  return g.u_on && chunk < g.u_nchunk;
}
__device__ __forceinline__ bool fl_kv_valid(const FlowGeo& g, int chunk) { return FL_CH * chunk + 8 * g.warp < g.u_n; }
__device__ __forceinline__ bool fl_o_valid(const FlowGeo& g) { return g.warp < FL_ROWS && g.cta + g.G * g.warp < KC; }
__device__ __forceinline__ bool fl_gu_valid(const FlowGeo& g, int j, int I) { return g.gw + j * g.NW < I; }
__device__ __forceinline__ int fl_d_rows(const FlowGeo& g, int j0, int j1) {
  int n = 0;
  for (int j = j0; j < j1; ++j) n += (g.cta + g.G * j < KC);
  return n;
}
__device__ __forceinline__ bool fl_h_valid(const FlowGeo& g, int j) { return g.gw + j * g.NW < g.nheads_tasks; }

__device__ __forceinline__ void fl_qkv_rows(const FlowP& p, int task, int& r0, int& r1) {
  const int half = p.hd / 2, nq = p.Hq * half;
  int t = task, base = 0;
  if (t >= 2 * nq) { t -= 2 * nq; base = 2 * p.Hq * p.hd; }
  else if (t >= nq) { t -= nq; base = p.Hq * p.hd; }
  r0 = base + (t / half) * p.hd + (t % half);
  r1 = r0 + half;
}

// Issue side: position in the warp's task sequence.
struct FlowIss {
  int l, st, sub, n;
};

// Find the next task at or after `it`, post its bulk copies into slot it.n % FL_SLOTS (lane 0) and advance.
__device__ __forceinline__ void fl_issue(const FlowP& p, const FlowGeo& g, FlowIss& it, uint32_t ring, uint32_t bars,
                                         uint64_t pol_w, uint64_t pol_kv) {
  const float* src[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int bytes = 0, ncopy = 0;
  bool kvtask = false;
  while (ncopy == 0) {
    if (it.l < p.L) {
      const float* Wl = p.W + p.layer0 + (int64_t)it.l * p.layer_stride;
      const int st = it.st;
      if (st == FS_KV) {
        int chunk;
        if (!fl_kv_more(g, it.sub, chunk)) { it.st = FS_O; it.sub = 0; continue; }
        it.sub++;
        if (!fl_kv_valid(g, chunk)) continue;
        const int t0 = FL_CH * chunk + 8 * g.warp;
        const int page = __ldg(p.block_table + g.u_b * p.pages_per_row + t0 / kPageTokens);
        const float* kvl = p.kv + (size_t)it.l * p.kv_layer_floats;
        src[0] = kvl + kv_off(page, 0, g.u_h, t0 % kPageTokens, p.Hq, p.hd);
        src[1] = kvl + kv_off(page, 1, g.u_h, t0 % kPageTokens, p.Hq, p.hd);
        bytes = 8 * 64 * 4; ncopy = 2; kvtask = true;
        continue;
      }
      if (st == FS_Q0 || st == FS_Q1) {
        const int j = st - FS_Q0;
        if (fl_q_valid(g, j)) {
          int r0, r1;
          fl_qkv_rows(p, g.gw + j * g.NW, r0, r1);
          src[0] = Wl + p.o_wqkv + (size_t)r0 * KC; src[1] = Wl + p.o_wqkv + (size_t)r1 * KC;
          bytes = KC * 4; ncopy = 2;
        }
      } else if (st == FS_O) {
        if (fl_o_valid(g)) { src[0] = Wl + p.o_wo + (size_t)(g.cta + g.G * g.warp) * KC; bytes = KC * 4; ncopy = 1; }
      } else if (st <= FS_GU2) {
        const int j = st - FS_GU0;
        if (fl_gu_valid(g, j, p.I)) {
          const int t = g.gw + j * g.NW;
          src[0] = Wl + p.o_wgu + (size_t)t * KC; src[1] = Wl + p.o_wgu + (size_t)(p.I + t) * KC;
          bytes = KC * 4; ncopy = 2;
        }
      } else {  // FS_D0 / FS_D1: this warp's 384-column slice of the CTA's down rows
        const int j0 = st == FS_D0 ? 0 : 4, j1 = st == FS_D0 ? 4 : FL_ROWS;
        const int nr = fl_d_rows(g, j0, j1);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nr) src[k] = Wl + p.o_wd + (size_t)(g.cta + g.G * (j0 + k)) * p.I + g.warp * (p.I / FL_WARPS);
        bytes = (p.I / FL_WARPS) * 4; ncopy = nr;
      }
      it.st = st + 1;
      if (it.st == FS_NLAYER) { it.st = 0; it.l++; }
    } else {
      if (!fl_h_valid(g, it.sub)) return;  // end of the sequence
      const int t = g.gw + it.sub * g.NW, nrows = p.rows_per_item * p.V;
      src[0] = p.W + p.o_head + (size_t)(2 * t) * KC; src[1] = p.W + p.o_head + (size_t)min(2 * t + 1, nrows - 1) * KC;
      bytes = KC * 4; ncopy = 2;
      it.sub++;
    }
  }
  if (g.lane == 0) {
    const int slot = it.n % FL_SLOTS;
    const uint32_t bar = bars + slot * 8, dst = ring + slot * FL_SLOT_BYTES;
    fl_expect(bar, (uint32_t)(bytes * ncopy));
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < ncopy) fl_bulk(dst + k * bytes, src[k], (uint32_t)bytes, bar, kvtask ? pol_kv : pol_w);
  }
  it.n++;
}

// Consume side of the ring.
struct FlowRing {
  float* base;      // this warp's slots (generic pointer)
  uint32_t ring;    // same, shared-space address
  uint32_t bars;    // this warp's FL_SLOTS mbarriers
  int n;            // tasks consumed so far
};
__device__ __forceinline__ const float* fl_ring_wait(FlowRing& r, FlowWd& wd) {
  const int slot = r.n % FL_SLOTS;
  const uint32_t parity = (uint32_t)(r.n / FL_SLOTS) & 1u;
  wd.spins = 0;
  while (!fl_try_wait(r.bars + slot * 8, parity))
    if (__any_sync(0xffffffffu, fl_giveup(wd, 0x100 + slot))) { wd.dead = 1; break; }
  return r.base + slot * FL_SLOT_FLOATS;
}
__device__ __forceinline__ void fl_ring_release(const FlowP& p, const FlowGeo& g, FlowRing& r, FlowIss& it, uint64_t pol_w,
                                                uint64_t pol_kv) {
  __syncwarp();  // every lane's reads of the slot are complete before the async proxy overwrites it
  r.n++;
  fl_issue(p, g, it, r.ring, r.bars, pol_w, pol_kv);
}

// ---------------------------------------------------------------- phase helpers
// Poll p.B x 768 LL words into xs (raw), zero rows >= B, block barrier.
template <int BT>
__device__ __forceinline__ void fl_stage768(const FlowP& p, const unsigned long long* src, uint32_t tag, float* xs, FlowWd& wd) {
  const int tid = threadIdx.x;
  unsigned long long v[BT][3];
  wd.spins = 0;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int b = 0; b < BT; ++b)
      if (b < p.B) {
#pragma unroll
        for (int k = 0; k < 3; ++k) v[b][k] = ll_ld(src + b * KC + tid + 256 * k);
      }
#pragma unroll
    for (int b = 0; b < BT; ++b)
      if (b < p.B) {
#pragma unroll
        for (int k = 0; k < 3; ++k) ok = ok && (ll_tag(v[b][k]) == tag);
      }
    if (__all_sync(0xffffffffu, ok)) break;
    if (__any_sync(0xffffffffu, fl_giveup(wd, 0x200 + (tag & 0xff)))) { wd.dead = 1; break; }
  }
#pragma unroll
  for (int b = 0; b < BT; ++b) {
#pragma unroll
    for (int k = 0; k < 3; ++k) xs[b * KC + tid + 256 * k] = b < p.B ? ll_val(v[b][k]) : 0.f;
  }
  __syncthreads();
}

template <int BT>
__device__ __forceinline__ void fl_load_x(const float* xs, float (&x)[BT][24], int lane) {
#pragma unroll
  for (int b = 0; b < BT; ++b) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 v = reinterpret_cast<const float4*>(xs)[b * (KC / 4) + i * 32 + lane];
      x[b][4 * i] = v.x; x[b][4 * i + 1] = v.y; x[b][4 * i + 2] = v.z; x[b][4 * i + 3] = v.w;
    }
  }
}

// HF LlamaRMSNorm on the register copy: w * (x * rsqrt(mean(x^2) + eps)); every warp computes the same sum.
template <int BT>
__device__ __forceinline__ void fl_norm(float (&x)[BT][24], const float4 (&nw)[6], float eps) {
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 24; ++j) ss = fmaf(x[b][j], x[b][j], ss);
    ss = warp_sum(ss);
    const float rinv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)KC), eps)));
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      x[b][4 * i] = __fmul_rn(nw[i].x, __fmul_rn(x[b][4 * i], rinv));
      x[b][4 * i + 1] = __fmul_rn(nw[i].y, __fmul_rn(x[b][4 * i + 1], rinv));
      x[b][4 * i + 2] = __fmul_rn(nw[i].z, __fmul_rn(x[b][4 * i + 2], rinv));
      x[b][4 * i + 3] = __fmul_rn(nw[i].w, __fmul_rn(x[b][4 * i + 3], rinv));
    }
  }
}
__device__ __forceinline__ void fl_load_nw(const float* normw, float4 (&nw)[6], int lane) {
#pragma unroll
  for (int i = 0; i < 6; ++i) nw[i] = ldg_stream(reinterpret_cast<const float4*>(normw) + i * 32 + lane);
}

// two weight rows (slot + 0, slot + 768 floats) against BT activation rows; k order as k_gemv
template <int BT>
__device__ __forceinline__ void fl_dot2(const float* slot, const float (&x)[BT][24], float (&a0)[BT], float (&a1)[BT], int lane) {
  const float4* w0 = reinterpret_cast<const float4*>(slot) + lane;
  const float4* w1 = w0 + KC / 4;
#pragma unroll
  for (int b = 0; b < BT; ++b) { a0[b] = 0.f; a1[b] = 0.f; }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 u = w0[i * 32], v = w1[i * 32];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      a0[b] = fmaf(u.x, x[b][4 * i], a0[b]); a0[b] = fmaf(u.y, x[b][4 * i + 1], a0[b]);
      a0[b] = fmaf(u.z, x[b][4 * i + 2], a0[b]); a0[b] = fmaf(u.w, x[b][4 * i + 3], a0[b]);
      a1[b] = fmaf(v.x, x[b][4 * i], a1[b]); a1[b] = fmaf(v.y, x[b][4 * i + 1], a1[b]);
      a1[b] = fmaf(v.z, x[b][4 * i + 2], a1[b]); a1[b] = fmaf(v.w, x[b][4 * i + 3], a1[b]);
    }
  }
}
template <int BT>
__device__ __forceinline__ void fl_dot1(const float* slot, const float (&x)[BT][24], float (&a0)[BT], int lane) {
  const float4* w0 = reinterpret_cast<const float4*>(slot) + lane;
#pragma unroll
  for (int b = 0; b < BT; ++b) a0[b] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 u = w0[i * 32];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      a0[b] = fmaf(u.x, x[b][4 * i], a0[b]); a0[b] = fmaf(u.y, x[b][4 * i + 1], a0[b]);
      a0[b] = fmaf(u.z, x[b][4 * i + 2], a0[b]); a0[b] = fmaf(u.w, x[b][4 * i + 3], a0[b]);
    }
  }
}

// Value v of batch row b lives in lane b * LPB: write it to word `idx` of every replica (LPB lanes share the work).
template <int BT>
__device__ __forceinline__ void fl_bcast_store(unsigned long long* rep0, int R, size_t idx, float v, uint32_t tag, int nb, int lane) {
  constexpr int LPB = 32 / BT;
  const float vb = __shfl_sync(0xffffffffu, v, (lane / LPB) * LPB);
  if (lane / LPB < nb)
    for (int r = lane % LPB; r < R; r += LPB) ll_st(rep0 + (size_t)r * FL_REP_STRIDE + idx, vb, tag);
}

template <int BT>
__global__ void __launch_bounds__(FL_THREADS, 1) k_flow(const FlowP p) {
  extern __shared__ __align__(128) unsigned char fl_smem[];
  float* xs = reinterpret_cast<float*>(fl_smem + FL_RING_BYTES);  // [BT][768]
  __shared__ __align__(8) uint64_t s_bar[FL_WARPS * FL_SLOTS];
  __shared__ int s_pos[BT], s_active[BT], s_page[BT];
  __shared__ float s_cos[BT * 64], s_sin[BT * 64];
  __shared__ float s_red[FL_ROWS][FL_WARPS][BT];
  __shared__ float s_ml[FL_HEADS * FL_SMAX * 2];
  __shared__ float s_am[FL_WARPS], s_al[FL_WARPS];
  __shared__ __align__(16) float s_ao[FL_WARPS][64];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int LPB = 32 / BT;
  if (tid < FL_WARPS * FL_SLOTS) mbar_init(&s_bar[tid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  if (p.decode && ldg_cg(&p.st->all_finished)) return;  // uniform over the grid; nothing has been issued yet

  FlowWd wd{&p.st->err, 0, 0};
  const uint32_t base = (uint32_t)ldg_cg(reinterpret_cast<const int*>(p.epoch));
  int tr = 0;
#define FL_TRACE() do { if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[tr++] = globaltimer_ns(); } while (0)
  FL_TRACE();

  // ---- positions / pages / RoPE rows of this step (k_input)
  if (tid < BT) {
    int act = 0, pos = 0;
    if (tid < p.B) {
      pos = ldg_cg(&p.seq_len[tid]);
      act = p.decode ? 1 : (p.mask[(size_t)tid * p.T0 + p.col] != 0);
    }
    s_pos[tid] = pos; s_active[tid] = act;
    s_page[tid] = tid < p.B ? __ldg(p.block_table + tid * p.pages_per_row + pos / kPageTokens) : 0;
  }
  __syncthreads();

  FlowGeo g;
  g.G = gridDim.x; g.NW = g.G * FL_WARPS; g.cta = blockIdx.x; g.warp = warp; g.lane = lane; g.gw = g.cta * FL_WARPS + warp;
  g.S = max(1, min(FL_SMAX, g.G / (p.Hq * p.B)));
  {
    const int u = g.cta;
    g.u_on = 0; g.u_b = 0; g.u_h = 0; g.u_split = 0; g.u_n = 0; g.u_nchunk = 0;
    if (u < p.B * p.Hq * g.S) {
      g.u_split = u % g.S; g.u_h = (u / g.S) % p.Hq; g.u_b = u / (g.S * p.Hq);
      g.u_n = s_pos[g.u_b] + 1;
      g.u_nchunk = (g.u_n + FL_CH - 1) / FL_CH;
      g.u_on = s_active[g.u_b] && g.u_split < min(g.u_nchunk, g.S);
    }
  }
  g.nheads_tasks = p.sample ? (p.rows_per_item * p.V + 1) / 2 : 0;

  uint64_t pol_w, pol_kv;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_w));
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_kv));
  FlowRing ring;
  ring.base = reinterpret_cast<float*>(fl_smem) + (size_t)warp * FL_SLOTS * FL_SLOT_FLOATS;
  ring.ring = smem_u32(ring.base);
  ring.bars = smem_u32(&s_bar[warp * FL_SLOTS]);
  ring.n = 0;
  FlowIss it{0, 0, 0, 0};
#pragma unroll 1
  for (int k = 0; k < FL_SLOTS; ++k) fl_issue(p, g, it, ring.ring, ring.bars, pol_w, pol_kv);

  for (int i = tid; i < BT * 64; i += FL_THREADS) {  // RoPE rows of this step's positions (p.hd == 64)
    const int b = i / 64;
    s_cos[i] = b < p.B ? __ldg(p.W + p.o_cos + (size_t)s_pos[b] * 64 + (i % 64)) : 0.f;
    s_sin[i] = b < p.B ? __ldg(p.W + p.o_sin + (size_t)s_pos[b] * 64 + (i % 64)) : 0.f;
  }
  const int ngen = p.decode ? ldg_cg(&p.st->n_gen) : 0;
  for (int i = tid; i < BT * KC; i += FL_THREADS) {  // step input: prompt column or sum of the code embeddings
    const int b = i / KC, k = i % KC;
    float v = 0.f;
    if (b < p.B) {
      if (!p.decode) {
        v = s_active[b] ? p.emb[((size_t)b * p.T0 + p.col) * KC + k] : 0.f;
      } else {
        const int32_t* id = p.ids_out + ((size_t)b * p.max_new + (ngen - 1)) * p.num_vq;
        if (p.infer_text) {
          v = p.W[p.o_emb_text + (size_t)ldg_cg(&id[0]) * KC + k];
        } else {
          for (int q = 0; q < p.num_vq; ++q) v += p.W[p.o_emb_code + ((size_t)q * p.num_audio + ldg_cg(&id[q])) * KC + k];
        }
      }
    }
    xs[i] = v;
  }
  __syncthreads();

  const int R = p.R, myrep = g.cta % R;
  const int o_row = g.cta + g.G * warp;  // this warp's O-proj row (valid iff fl_o_valid)

  for (int l = 0; l < p.L; ++l) {
    const float* Wl = p.W + p.layer0 + (int64_t)l * p.layer_stride;
    unsigned long long* par = p.arena + (size_t)(l & 1) * FL_PARITY_WORDS;
    unsigned long long* parn = p.arena + (size_t)((l + 1) & 1) * FL_PARITY_WORDS;
    unsigned long long* qkvw = par + (size_t)FL_RMAX * FL_REP_STRIDE;
    const unsigned long long* myr = par + (size_t)myrep * FL_REP_STRIDE;
    const uint32_t tagl = base + 8u * (uint32_t)l;

    // ============ A: QKV + RoPE + KV append ============
    {
      float4 nw[6];
      fl_load_nw(Wl + p.o_ln1, nw, lane);
      if (l > 0) fl_stage768<BT>(p, myr + FL_A_X, tagl + FT_X, xs, wd);
      float x[BT][24];
      fl_load_x<BT>(xs, x, lane);
      // residual of this warp's O-proj row (raw x), kept for phase C
      float res_o = 0.f;
      if (fl_o_valid(g)) res_o = xs[(lane / LPB) * KC + o_row];
      fl_norm<BT>(x, nw, p.eps);
      float res_keep = res_o;
#pragma unroll
      for (int j = 0; j < FL_QR; ++j) {
        if (!fl_q_valid(g, j)) continue;
        const float* slot = fl_ring_wait(ring, wd);
        float a0[BT], a1[BT];
        fl_dot2<BT>(slot, x, a0, a1, lane);
        warp_reduce_scatter<BT>(a0);
        warp_reduce_scatter<BT>(a1);
        const int b = lane / LPB;
        if ((lane % LPB) == 0 && b < p.B && s_active[b]) {
          const int half = p.hd / 2, nq = p.Hq * half;
          int t = g.gw + j * g.NW, which = 0;
          if (t >= 2 * nq) { which = 2; t -= 2 * nq; }
          else if (t >= nq) { which = 1; t -= nq; }
          const int h = t / half, jj = t % half;
          const float v0 = a0[0], v1 = a1[0];
          float o0 = v0, o1 = v1;
          if (which < 2) {
            const float* cs = s_cos + b * 64;
            const float* sn = s_sin + b * 64;
            o0 = __fadd_rn(__fmul_rn(v0, cs[jj]), __fmul_rn(-v1, sn[jj]));
            o1 = __fadd_rn(__fmul_rn(v1, cs[jj + half]), __fmul_rn(v0, sn[jj + half]));
          }
          unsigned long long* dst = qkvw + (which == 0 ? FL_A_Q : which == 1 ? FL_A_KN : FL_A_VN) + b * KC + h * 64;
          ll_st(dst + jj, o0, tagl + FT_QKV);
          ll_st(dst + jj + half, o1, tagl + FT_QKV);
          if (which > 0) {
            float* kd = p.kv + (size_t)l * p.kv_layer_floats + kv_off(s_page[b], which - 1, h, s_pos[b] % kPageTokens, p.Hq, p.hd);
            kd[jj] = o0; kd[jj + half] = o1;
          }
        }
        fl_ring_release(p, g, ring, it, pol_w, pol_kv);
      }
      FL_TRACE();

      // ============ B: attention (one unit per CTA: row, head, key split) ============
      if (g.u_on) {
        const int sub = lane & 7, grp = lane >> 3;
        const int b = g.u_b, h = g.u_h, n = g.u_n, pos = n - 1;
        // q slice of this lane: dims sub*8 .. sub*8+7
        float q[8];
        {
          const unsigned long long* qp = qkvw + FL_A_Q + b * KC + h * 64 + sub * 8;
          unsigned long long w[8];
          wd.spins = 0;
          while (true) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) ll_ld2(qp + 2 * k, w[2 * k], w[2 * k + 1]);
#pragma unroll
            for (int k = 0; k < 8; ++k) ok = ok && (ll_tag(w[k]) == tagl + FT_QKV);
            if (__all_sync(0xffffffffu, ok)) break;
            if (__any_sync(0xffffffffu, fl_giveup(wd, 0x300))) { wd.dead = 1; break; }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) q[k] = ll_val(w[k]);
        }
        float M = -INFINITY, L = 0.f, O = 0.f;
        for (int sb = 0;; ++sb) {
          int chunk;
          if (!fl_kv_more(g, sb, chunk)) break;
          const bool have = fl_kv_valid(g, chunk);
          const float* slot = have ? fl_ring_wait(ring, wd) : nullptr;
          const int tbase = chunk * FL_CH + warp * 8 + grp;
          float4 k0[2], k1[2], v0[2], v1[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int t = tbase + 4 * i;
            k0[i] = k1[i] = v0[i] = v1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have && t < pos) {
              const float* kr = slot + (grp + 4 * i) * 64 + sub * 8;
              const float* vr = kr + 8 * 64;
              k0[i] = *reinterpret_cast<const float4*>(kr); k1[i] = *reinterpret_cast<const float4*>(kr + 4);
              v0[i] = *reinterpret_cast<const float4*>(vr); v1[i] = *reinterpret_cast<const float4*>(vr + 4);
            }
          }
          // the token of THIS step: its K/V rows arrive from the QKV phase through the LL region, not the cache
          {
            const bool mine0 = (tbase == pos), mine1 = (tbase + 4 == pos);
            if (__any_sync(0xffffffffu, mine0 || mine1)) {
              const unsigned long long* kp = qkvw + FL_A_KN + b * KC + h * 64 + sub * 8;
              const unsigned long long* vp = qkvw + FL_A_VN + b * KC + h * 64 + sub * 8;
              unsigned long long kw[8], vw[8];
              wd.spins = 0;
              while (true) {
                bool ok = true;
                if (mine0 || mine1) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) { ll_ld2(kp + 2 * k, kw[2 * k], kw[2 * k + 1]); ll_ld2(vp + 2 * k, vw[2 * k], vw[2 * k + 1]); }
#pragma unroll
                  for (int k = 0; k < 8; ++k) ok = ok && (ll_tag(kw[k]) == tagl + FT_QKV) && (ll_tag(vw[k]) == tagl + FT_QKV);
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if (__any_sync(0xffffffffu, fl_giveup(wd, 0x301))) { wd.dead = 1; break; }
              }
              if (mine0 || mine1) {
                const int i = mine0 ? 0 : 1;
                const float4 a = make_float4(ll_val(kw[0]), ll_val(kw[1]), ll_val(kw[2]), ll_val(kw[3]));
                const float4 bq = make_float4(ll_val(kw[4]), ll_val(kw[5]), ll_val(kw[6]), ll_val(kw[7]));
                const float4 cq = make_float4(ll_val(vw[0]), ll_val(vw[1]), ll_val(vw[2]), ll_val(vw[3]));
                const float4 dq = make_float4(ll_val(vw[4]), ll_val(vw[5]), ll_val(vw[6]), ll_val(vw[7]));
                if (i == 0) { k0[0] = a; k1[0] = bq; v0[0] = cq; v1[0] = dq; }
                else { k0[1] = a; k1[1] = bq; v0[1] = cq; v1[1] = dq; }
              }
            }
          }
          float sc[2], m = -INFINITY;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float s = q[0] * k0[i].x + q[1] * k0[i].y + q[2] * k0[i].z + q[3] * k0[i].w + q[4] * k1[i].x + q[5] * k1[i].y +
                      q[6] * k1[i].z + q[7] * k1[i].w;
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            sc[i] = (tbase + 4 * i < n) ? s * p.scaling : -INFINITY;
            m = fmaxf(m, sc[i]);
          }
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
          float lsum = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (m > -INFINITY) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float e = expf(sc[i] - m);
              lsum += e;
              o[0] = fmaf(e, v0[i].x, o[0]); o[1] = fmaf(e, v0[i].y, o[1]); o[2] = fmaf(e, v0[i].z, o[2]);
              o[3] = fmaf(e, v0[i].w, o[3]); o[4] = fmaf(e, v1[i].x, o[4]); o[5] = fmaf(e, v1[i].y, o[5]);
              o[6] = fmaf(e, v1[i].z, o[6]); o[7] = fmaf(e, v1[i].w, o[7]);
            }
          }
          lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
          lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] += __shfl_xor_sync(0xffffffffu, o[j], 8);
            o[j] += __shfl_xor_sync(0xffffffffu, o[j], 16);
          }
          if (have) fl_ring_release(p, g, ring, it, pol_w, pol_kv);
          __syncthreads();  // previous chunk's merge no longer reads s_ao / s_am / s_al
          if (lane < 8) {
            *reinterpret_cast<float4*>(&s_ao[warp][lane * 8]) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(&s_ao[warp][lane * 8 + 4]) = make_float4(o[4], o[5], o[6], o[7]);
            if (lane == 0) { s_am[warp] = m; s_al[warp] = lsum; }
          }
          __syncthreads();
          if (tid < 64) {
            float cm = M;
#pragma unroll
            for (int w = 0; w < FL_WARPS; ++w) cm = fmaxf(cm, s_am[w]);
            const float fo = (M > -INFINITY) ? expf(M - cm) : 0.f;
            L *= fo; O *= fo;
#pragma unroll
            for (int w = 0; w < FL_WARPS; ++w) {
              const float f = (s_am[w] > -INFINITY) ? expf(s_am[w] - cm) : 0.f;
              L = fmaf(f, s_al[w], L);
              O = fmaf(f, s_ao[w][tid], O);
            }
            M = cm;
          }
        }
        if (tid < 64) {
          const size_t pidx = FL_A_P + ((size_t)(b * FL_HEADS + h) * FL_SMAX + g.u_split) * FL_PW;
          for (int r = 0; r < R; ++r) {
            unsigned long long* d = par + (size_t)r * FL_REP_STRIDE + pidx;
            ll_st(d + tid, O, tagl + FT_P);
            if (tid == 0) { ll_st(d + 64, M, tagl + FT_P); ll_st(d + 65, L, tagl + FT_P); }
          }
        }
      }
      FL_TRACE();

      // ============ C: merge the attention splits, O-proj + residual ============
      __syncthreads();  // xs (raw x) is no longer read by any warp of this CTA
      for (int b = 0; b < p.B; ++b) {
        if (!s_active[b]) {
#pragma unroll
          for (int k = 0; k < 3; ++k) xs[b * KC + tid + 256 * k] = 0.f;
          continue;
        }
        const int n = s_pos[b] + 1;
        const int ns = min((n + FL_CH - 1) / FL_CH, g.S);
        const unsigned long long* P = myr + FL_A_P + (size_t)b * FL_HEADS * FL_SMAX * FL_PW;
        float ov[3][FL_SMAX];
        wd.spins = 0;
        while (true) {
          bool ok = true;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int c = tid + 256 * k, hh = c >> 6, dd = c & 63;
#pragma unroll
            for (int s = 0; s < FL_SMAX; ++s)
              if (s < ns) {
                const unsigned long long w = ll_ld(P + (size_t)(hh * FL_SMAX + s) * FL_PW + dd);
                ok = ok && (ll_tag(w) == tagl + FT_P);
                ov[k][s] = ll_val(w);
              }
          }
          for (int i = tid; i < p.Hq * ns * 2; i += FL_THREADS) {
            const int hh = i / (2 * ns), rem = i % (2 * ns), s = rem >> 1, which = rem & 1;
            const unsigned long long w = ll_ld(P + (size_t)(hh * FL_SMAX + s) * FL_PW + 64 + which);
            ok = ok && (ll_tag(w) == tagl + FT_P);
            s_ml[(hh * FL_SMAX + s) * 2 + which] = ll_val(w);
          }
          if (__syncthreads_and(ok)) break;
          if (__syncthreads_or(fl_giveup(wd, 0x400))) { wd.dead = 1; break; }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int c = tid + 256 * k, hh = c >> 6;
          const float* ml = s_ml + hh * FL_SMAX * 2;
          float GM = -INFINITY;
#pragma unroll
          for (int s = 0; s < FL_SMAX; ++s)
            if (s < ns) GM = fmaxf(GM, ml[2 * s]);
          float GL = 0.f, GO = 0.f;
#pragma unroll
          for (int s = 0; s < FL_SMAX; ++s)
            if (s < ns) {
              const float w = expf(ml[2 * s] - GM);
              GL = fmaf(w, ml[2 * s + 1], GL);
              GO = fmaf(w, ov[k][s], GO);
            }
          xs[b * KC + c] = GO / GL;
        }
        __syncthreads();  // s_ml is rewritten by the next row
      }
      for (int b = p.B; b < BT; ++b) {
#pragma unroll
        for (int k = 0; k < 3; ++k) xs[b * KC + tid + 256 * k] = 0.f;
      }
      __syncthreads();
      if (fl_o_valid(g)) {
        fl_load_x<BT>(xs, x, lane);
        const float* slot = fl_ring_wait(ring, wd);
        float a0[BT];
        fl_dot1<BT>(slot, x, a0, lane);
        warp_reduce_scatter<BT>(a0);
        const float out = __fadd_rn(res_keep, a0[0]);
        fl_bcast_store<BT>(par, R, FL_A_XO + (size_t)(lane / LPB) * KC + o_row, out, tagl + FT_XO, p.B, lane);
        fl_ring_release(p, g, ring, it, pol_w, pol_kv);
      }
      FL_TRACE();
    }

    // ============ D: gate/up + SiLU * mul ============
    float res_d = 0.f;
    {
      float4 nw[6];
      fl_load_nw(Wl + p.o_ln2, nw, lane);
      __syncthreads();  // every warp is done with xs (attention output)
      fl_stage768<BT>(p, myr + FL_A_XO, tagl + FT_XO, xs, wd);
      if (tid < FL_ROWS * BT) {  // residual of this thread's down-phase output element (raw x')
        const int b = tid % BT, j = tid / BT, row = g.cta + g.G * j;
        if (row < KC) res_d = xs[b * KC + row];
      }
      float x[BT][24];
      fl_load_x<BT>(xs, x, lane);
      fl_norm<BT>(x, nw, p.eps);
#pragma unroll
      for (int j = 0; j < FL_GU; ++j) {
        if (!fl_gu_valid(g, j, p.I)) continue;
        const float* slot = fl_ring_wait(ring, wd);
        float a0[BT], a1[BT];
        fl_dot2<BT>(slot, x, a0, a1, lane);
        warp_reduce_scatter<BT>(a0);
        warp_reduce_scatter<BT>(a1);
        const float v0 = a0[0], v1 = a1[0];
        const float sg = __fdiv_rn(v0, __fadd_rn(1.0f, expf(-v0)));
        const float act = __fmul_rn(sg, v1);
        fl_bcast_store<BT>(par, R, FL_A_ACT + (size_t)(lane / LPB) * p.I + (g.gw + j * g.NW), act, tagl + FT_ACT, p.B, lane);
        fl_ring_release(p, g, ring, it, pol_w, pol_kv);
      }
      FL_TRACE();
    }

    // ============ E: down + residual (K = 3072 split over the 8 warps) ============
    {
      float xd[BT][12];
      {
        const int kslice = p.I / FL_WARPS;  // 384
#pragma unroll
        for (int b = 0; b < BT; ++b) {
#pragma unroll
          for (int k = 0; k < 12; ++k) xd[b][k] = 0.f;
          if (b < p.B) {  // one poll loop per batch row keeps the 64-bit words of only one row live
            const unsigned long long* ap = myr + FL_A_ACT + (size_t)b * p.I + warp * kslice;
            unsigned long long w[12];
            wd.spins = 0;
            while (true) {
              bool ok = true;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                ll_ld2(ap + (i * 32 + lane) * 4, w[4 * i], w[4 * i + 1]);
                ll_ld2(ap + (i * 32 + lane) * 4 + 2, w[4 * i + 2], w[4 * i + 3]);
              }
#pragma unroll
              for (int k = 0; k < 12; ++k) ok = ok && (ll_tag(w[k]) == tagl + FT_ACT);
              if (__all_sync(0xffffffffu, ok)) break;
              if (__any_sync(0xffffffffu, fl_giveup(wd, 0x500))) { wd.dead = 1; break; }
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) xd[b][k] = ll_val(w[k]);
          }
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j0 = half ? 4 : 0, j1 = half ? FL_ROWS : 4;
        const int nr = fl_d_rows(g, j0, j1);
        if (nr == 0) continue;
        const float* slot = fl_ring_wait(ring, wd);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= nr) break;
          const float4* wr = reinterpret_cast<const float4*>(slot + k * (3072 / FL_WARPS)) + lane;
          float acc[BT];
#pragma unroll
          for (int b = 0; b < BT; ++b) acc[b] = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float4 u = wr[i * 32];
#pragma unroll
            for (int b = 0; b < BT; ++b) {
              acc[b] = fmaf(u.x, xd[b][4 * i], acc[b]); acc[b] = fmaf(u.y, xd[b][4 * i + 1], acc[b]);
              acc[b] = fmaf(u.z, xd[b][4 * i + 2], acc[b]); acc[b] = fmaf(u.w, xd[b][4 * i + 3], acc[b]);
            }
          }
          warp_reduce_scatter<BT>(acc);
          if ((lane % LPB) == 0) s_red[j0 + k][warp][lane / LPB] = acc[0];
        }
        fl_ring_release(p, g, ring, it, pol_w, pol_kv);
      }
      __syncthreads();
      if (tid < FL_ROWS * BT) {  // K slices summed in the order 0..7 (deterministic)
        const int b = tid % BT, j = tid / BT, row = g.cta + g.G * j;
        if (row < KC && b < p.B) {
          float v = s_red[j][0][b];
#pragma unroll
          for (int w = 1; w < FL_WARPS; ++w) v = __fadd_rn(v, s_red[j][w][b]);
          const float out = __fadd_rn(res_d, v);
          for (int r = 0; r < R; ++r)
            ll_st(parn + (size_t)r * FL_REP_STRIDE + FL_A_X + (size_t)b * KC + row, out, tagl + 8u + FT_X);
        }
      }
      FL_TRACE();
    }
  }

  // ============ heads: final norm, logits, hidden state ============
  if (p.sample) {
    const unsigned long long* myr = p.arena + (size_t)(p.L & 1) * FL_PARITY_WORDS + (size_t)myrep * FL_REP_STRIDE;
    float4 nw[6];
    fl_load_nw(p.W + p.o_final_norm, nw, lane);
    __syncthreads();
    fl_stage768<BT>(p, myr + FL_A_X, base + 8u * (uint32_t)p.L + FT_X, xs, wd);
    float x[BT][24];
    fl_load_x<BT>(xs, x, lane);
    fl_norm<BT>(x, nw, p.eps);
    if (p.hidden_out != nullptr && blockIdx.x == 0 && warp == 0) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b < p.B) {
          float4* dst = reinterpret_cast<float4*>(p.hidden_out + (size_t)b * p.hidden_stride + (size_t)ngen * KC);
#pragma unroll
          for (int i = 0; i < 6; ++i) dst[i * 32 + lane] = make_float4(x[b][4 * i], x[b][4 * i + 1], x[b][4 * i + 2], x[b][4 * i + 3]);
        }
    }
    const int nrows = p.rows_per_item * p.V;
    for (int j = 0; fl_h_valid(g, j); ++j) {
      const int t = g.gw + j * g.NW;
      const float* slot = fl_ring_wait(ring, wd);
      float a0[BT], a1[BT];
      fl_dot2<BT>(slot, x, a0, a1, lane);
      warp_reduce_scatter<BT>(a0);
      warp_reduce_scatter<BT>(a1);
      const int b = lane / LPB;
      if ((lane % LPB) == 0 && b < p.B) {
        const int r0 = 2 * t, q0 = r0 / p.V, c0 = r0 % p.V;
        p.logits[((size_t)b * p.rows_per_item + q0) * p.V + c0] = a0[0];
        if (r0 + 1 < nrows) {
          const int q1 = (r0 + 1) / p.V, c1 = (r0 + 1) % p.V;
          p.logits[((size_t)b * p.rows_per_item + q1) * p.V + c1] = a1[0];
        }
      }
      fl_ring_release(p, g, ring, it, pol_w, pol_kv);
    }
  }
  FL_TRACE();
  if (blockIdx.x == 0) {
    // positions advance once per step; the tag base advances so that no word of this launch can satisfy the next
    if (tid < p.B && s_active[tid]) p.seq_len[tid] = s_pos[tid] + 1;
    if (tid == 0) *p.epoch = base + FL_EPOCH_STEP;
  }
#undef FL_TRACE
}

}  // namespace ctb
