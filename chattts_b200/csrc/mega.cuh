// One-kernel decode step for small batches (B <= 8):  k_step<BT>
//
// The 100+ dependent phases of a decode step (20 x [QKV -> attention -> O -> gate/up -> down] -> heads) are
// latency-bound when each is its own launch.  k_step runs them all in ONE persistent cooperative kernel
// (one CTA per SM): phases are separated by a grid-wide barrier in global memory instead of a kernel
// boundary, every warp requests the weights of its first task of the NEXT phase before it arrives at the
// barrier (weights never depend on activations), activations move through L2 with ld.global.cg only.
// Math, reduction order and epilogues are those of k_gemv / k_attn (gpt_kernels.cuh), so results are
// bit-identical to the multi-kernel FMA path.
#pragma once
#include "gpt_kernels.cuh"

namespace ctb {

constexpr int MG_THREADS = 256;
constexpr int MG_WARPS = MG_THREADS / 32;
constexpr int MG_DOWN_PAIRS = 3;  // row pairs of the down projection per CTA (ceil(384 / grid) <= 3 for grid >= 128)

struct MegaP {
  const float* W;           // packed fp32 blob
  int64_t layer0, layer_stride, o_wqkv, o_wo, o_wgu, o_wd, o_ln1, o_ln2, o_final_norm, o_head, o_emb_code, o_emb_text,
      o_cos, o_sin;
  int L, d, I, Hq, Hkv, hd;
  float eps, scaling;
  float *x, *qbuf, *attn, *mlp, *logits, *kv, *part;
  size_t kv_layer_floats;
  const int* block_table; int pages_per_row;
  int* seq_len; int* counter; int nsplit_max;
  LoopState* st; unsigned* bar;
  int decode, col, T0, sample;
  const float* emb; const uint8_t* mask; const int32_t* ids_out;
  int max_new, num_vq, num_audio, infer_text, B;
  float* hidden_out; int hidden_stride, rows_per_item, V;
  unsigned long long* trace;  // optional [1 + 5*L + 2] globaltimer stamps of CTA 0 (profiling aid)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid-wide barrier: monotonically increasing arrival counter (zeroed by the host before the launch)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Split in two so that the next phase's weight prefetch is issued BETWEEN arrive and wait: a release must not have
// to wait for those (long) loads, and they stream while the CTA waits for the others.
__device__ __forceinline__ void grid_arrive(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    // release: this CTA's global writes (ordered before by bar.sync) become visible with the arrival
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(unsigned* counter, unsigned& epoch) {
  if (threadIdx.x == 0) {
    while (ld_acquire_u32(counter) < epoch) {}
  }
  __syncthreads();
}

enum MgEpi { MG_QKV = 0, MG_OPROJ = 1, MG_GATEUP = 2, MG_HEADS = 3 };

template <int BT>
struct MegaCtx {
  float* xs;       // [BT][768] (or [BT][3072] for the down phase)
  float* rinv;     // [BT]
  int* pos;        // [BT]
  int* active;     // [BT]
  const float* cos_s;  // [BT][64] RoPE rows of this step's positions (shared memory)
  const float* sin_s;
  const int* page;     // [BT] KV page of this step's position
};

// rows of a 2-row warp task (same mapping as k_gemv)
template <int EPI>
__device__ __forceinline__ void mg_task_rows(const MegaP& p, int task, int nrows, int& r0, int& r1) {
  if (EPI == MG_QKV) {
    const int half = p.hd / 2;
    int t = task, base = 0;
    const int nq = p.Hq * half, nk = p.Hkv * half;
    if (t >= nq + nk) { t -= nq + nk; base = (p.Hq + p.Hkv) * p.hd; }
    else if (t >= nq) { t -= nq; base = p.Hq * p.hd; }
    r0 = base + (t / half) * p.hd + (t % half);
    r1 = r0 + half;
  } else if (EPI == MG_GATEUP) {
    r0 = task; r1 = p.I + task;
  } else {
    r0 = 2 * task; r1 = min(2 * task + 1, nrows - 1);
  }
}

template <int EPI>
__device__ __forceinline__ void mg_load_w(const MegaP& p, const float* Wm, int task, int nrows, int lane, float4 (&w0)[6],
                                          float4 (&w1)[6]) {
  int r0, r1;
  mg_task_rows<EPI>(p, task, nrows, r0, r1);
  const float4* w0p = reinterpret_cast<const float4*>(Wm + (size_t)r0 * KC) + lane;
  const float4* w1p = reinterpret_cast<const float4*>(Wm + (size_t)r1 * KC) + lane;
#pragma unroll
  for (int i = 0; i < 6; ++i) { w0[i] = ldg_stream(w0p + i * 32); w1[i] = ldg_stream(w1p + i * 32); }
}

template <int BT, int EPI>
__device__ __forceinline__ void mg_stage(const MegaP& p, const MegaCtx<BT>& c, const float* xin, const float* normw, bool staged) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = p.B;
  if (!staged) {
    for (int i = tid; i < BT * (KC / 4); i += MG_THREADS) {
      const int b = i / (KC / 4), k4 = i % (KC / 4);
      if (b < nb) cp_async16(&c.xs[i * 4], xin + (size_t)b * KC + k4 * 4);
      else reinterpret_cast<float4*>(c.xs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    __syncthreads();
  }
  if (normw != nullptr) {
    for (int b = warp; b < BT; b += MG_WARPS) {
      float ss = 0.f;
#pragma unroll
      for (int k = lane; k < KC; k += 32) { const float v = c.xs[b * KC + k]; ss = fmaf(v, v, ss); }
      ss = warp_sum(ss);
      if (lane == 0) c.rinv[b] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)KC), p.eps)));
    }
    __syncthreads();
    for (int i = tid; i < BT * KC; i += MG_THREADS) {
      const int b = i / KC, k = i % KC;
      c.xs[i] = __fmul_rn(__ldg(normw + k), __fmul_rn(c.xs[i], c.rinv[b]));
    }
    __syncthreads();
    if (EPI == MG_HEADS && p.hidden_out != nullptr && blockIdx.x == 0) {
      const int step = ldg_cg(&p.st->n_gen);
      for (int i = tid; i < nb * KC; i += MG_THREADS) {
        const int b = i / KC, k = i % KC;
        p.hidden_out[(size_t)b * p.hidden_stride + (size_t)step * KC + k] = c.xs[i];
      }
    }
  }
}

// K = 768 phase: stage activations (+ optional RMSNorm) and stream this CTA's row tasks.  `pre0/pre1` hold the
// first task's weights, requested before the preceding grid barrier.
template <int BT, int EPI>
__device__ __forceinline__ void mg_gemv_phase(const MegaP& p, const MegaCtx<BT>& c, const float* Wm, int ntasks, int nrows,
                                              const float* xin, const float* normw, float* kvl, float4 (&w0)[6],
                                              float4 (&w1)[6], bool staged, float rx0 = 0.f, float rx1 = 0.f) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = p.B;
  mg_stage<BT, EPI>(p, c, xin, normw, staged);
  constexpr int LPB = 32 / BT;
  const int tstride = gridDim.x * MG_WARPS;
  for (int task = blockIdx.x * MG_WARPS + warp; task < ntasks; task += tstride) {
    float4 n0[6], n1[6];
    const bool more = task + tstride < ntasks;
    if (more) mg_load_w<EPI>(p, Wm, task + tstride, nrows, lane, n0, n1);
    float acc0[BT], acc1[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 xv = reinterpret_cast<const float4*>(c.xs)[b * (KC / 4) + i * 32 + lane];
        acc0[b] = fmaf(w0[i].x, xv.x, acc0[b]); acc0[b] = fmaf(w0[i].y, xv.y, acc0[b]);
        acc0[b] = fmaf(w0[i].z, xv.z, acc0[b]); acc0[b] = fmaf(w0[i].w, xv.w, acc0[b]);
        acc1[b] = fmaf(w1[i].x, xv.x, acc1[b]); acc1[b] = fmaf(w1[i].y, xv.y, acc1[b]);
        acc1[b] = fmaf(w1[i].z, xv.z, acc1[b]); acc1[b] = fmaf(w1[i].w, xv.w, acc1[b]);
      }
    }
    if (more) {
#pragma unroll
      for (int i = 0; i < 6; ++i) { w0[i] = n0[i]; w1[i] = n1[i]; }
    }
    warp_reduce_scatter<BT>(acc0);
    warp_reduce_scatter<BT>(acc1);
    if ((lane % LPB) != 0 || (lane / LPB) >= nb) continue;
    const int b = lane / LPB;
    const float v0 = acc0[0], v1 = acc1[0];
    int r0, r1;
    mg_task_rows<EPI>(p, task, nrows, r0, r1);
    if (EPI == MG_QKV) {
      if (!c.active[b]) continue;
      const int half = p.hd / 2;
      const int nq = p.Hq * half, nk = p.Hkv * half;
      int t = task, which = 0;
      if (t >= nq + nk) { which = 2; t -= nq + nk; }
      else if (t >= nq) { which = 1; t -= nq; }
      const int h = t / half, j = t % half;
      const int pos = c.pos[b];
      float o0 = v0, o1 = v1;
      if (which < 2) {
        const float* cs = c.cos_s + b * 64;
        const float* sn = c.sin_s + b * 64;
        o0 = __fadd_rn(__fmul_rn(v0, cs[j]), __fmul_rn(-v1, sn[j]));
        o1 = __fadd_rn(__fmul_rn(v1, cs[j + half]), __fmul_rn(v0, sn[j + half]));
      }
      if (which == 0) {
        p.qbuf[(size_t)b * p.Hq * p.hd + h * p.hd + j] = o0;
        p.qbuf[(size_t)b * p.Hq * p.hd + h * p.hd + j + half] = o1;
      } else {
        float* dst = kvl + kv_off(c.page[b], which - 1, h, pos % kPageTokens, p.Hkv, p.hd);
        dst[j] = o0; dst[j + half] = o1;
      }
    } else if (EPI == MG_OPROJ) {
      // one task per warp (384 tasks <= grid * 8): the residual values were requested before the grid barrier
      p.x[(size_t)b * p.d + r0] = __fadd_rn(rx0, v0);
      if (2 * task + 1 < nrows) p.x[(size_t)b * p.d + r0 + 1] = __fadd_rn(rx1, v1);
    } else if (EPI == MG_GATEUP) {
      const float sg = __fdiv_rn(v0, __fadd_rn(1.0f, expf(-v0)));
      p.mlp[(size_t)b * p.I + task] = __fmul_rn(sg, v1);
    } else {
      const int q0 = r0 / p.V, c0 = r0 % p.V;
      p.logits[((size_t)b * p.rows_per_item + q0) * p.V + c0] = v0;
      if (2 * task + 1 < nrows) {
        const int q1 = (2 * task + 1) / p.V, c1 = (2 * task + 1) % p.V;
        p.logits[((size_t)b * p.rows_per_item + q1) * p.V + c1] = v1;
      }
    }
  }
}

// gate/up for BT <= 4: all (<= 3) tasks of the warp were requested before the barrier (144 registers), so the phase
// exposes no DRAM round trip at all.
constexpr int MG_GU_TASKS = 3;
__host__ __device__ constexpr int mg_chunk(int BT) { return BT >= 4 ? 128 : 64; }  // keys per attention chunk
constexpr int MG_SMAX = 12;  // max attention splits per (row, head) when the O-proj phase merges them (BT <= 2)
// Shared-memory landing zone for the warp's gate/up weights: [warp][task][row][192 float4] (18 KiB per warp).
// cp.async needs no registers, so all 36 x 16 B requests per lane are in flight while the CTA waits at the barrier.
constexpr int MG_GW_FLOATS = MG_WARPS * MG_GU_TASKS * 2 * KC;
template <int BT>
__device__ __forceinline__ void mg_gateup_load(const MegaP& p, const float* Wm, int lane, int warp, float* gws) {
  const int tstride = gridDim.x * MG_WARPS;
#pragma unroll
  for (int j = 0; j < MG_GU_TASKS; ++j) {
    const int task = blockIdx.x * MG_WARPS + warp + j * tstride;
    if (task < p.I) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float* src = Wm + (size_t)(r ? p.I + task : task) * KC;
        float* dst = gws + ((size_t)(warp * MG_GU_TASKS + j) * 2 + r) * KC;
#pragma unroll
        for (int i = 0; i < 6; ++i) cp_async16(dst + (i * 32 + lane) * 4, src + (i * 32 + lane) * 4);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int BT>
__device__ __forceinline__ void mg_gateup_small(const MegaP& p, const MegaCtx<BT>& c, const float* normw, const float* gws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mg_stage<BT, MG_GATEUP>(p, c, p.x, normw, false);  // its cp.async wait also covers the weight group
  __syncwarp();
  constexpr int LPB = 32 / BT;
  const int tstride = gridDim.x * MG_WARPS;
#pragma unroll
  for (int j = 0; j < MG_GU_TASKS; ++j) {
    const int task = blockIdx.x * MG_WARPS + warp + j * tstride;
    if (task >= p.I) break;
    const float4* g0 = reinterpret_cast<const float4*>(gws + ((size_t)(warp * MG_GU_TASKS + j) * 2) * KC);
    const float4* g1 = g0 + KC / 4;
    float acc0[BT], acc1[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 a = g0[i * 32 + lane], bq = g1[i * 32 + lane];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 xv = reinterpret_cast<const float4*>(c.xs)[b * (KC / 4) + i * 32 + lane];
        acc0[b] = fmaf(a.x, xv.x, acc0[b]); acc0[b] = fmaf(a.y, xv.y, acc0[b]);
        acc0[b] = fmaf(a.z, xv.z, acc0[b]); acc0[b] = fmaf(a.w, xv.w, acc0[b]);
        acc1[b] = fmaf(bq.x, xv.x, acc1[b]); acc1[b] = fmaf(bq.y, xv.y, acc1[b]);
        acc1[b] = fmaf(bq.z, xv.z, acc1[b]); acc1[b] = fmaf(bq.w, xv.w, acc1[b]);
      }
    }
    warp_reduce_scatter<BT>(acc0);
    warp_reduce_scatter<BT>(acc1);
    if ((lane % LPB) == 0 && (lane / LPB) < p.B) {
      const float v0 = acc0[0], v1 = acc1[0];
      const float sg = __fdiv_rn(v0, __fadd_rn(1.0f, expf(-v0)));
      p.mlp[(size_t)(lane / LPB) * p.I + task] = __fmul_rn(sg, v1);
    }
  }
}

// attention phase: units (b, h, s) strided over the grid; unit s walks key chunks s, s + S, ... with a running softmax
template <int BT, bool DEFER>
__device__ __forceinline__ void mg_attn_phase(const MegaP& p, const MegaCtx<BT>& c, const float* kvl, int S) {
  constexpr int HD = 64, NW = MG_WARPS, CH = mg_chunk(BT), PER_WARP = CH / NW, ITER = PER_WARP / 4;
  static_assert(ITER >= 1, "attention chunk too small for the warp count");
  __shared__ float s_m[NW], s_l[NW];
  __shared__ __align__(16) float s_o[NW][HD];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  const int nunits = p.B * p.Hq * S;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int split = u % S, h = (u / S) % p.Hq, b = u / (S * p.Hq);
    float* outp = p.attn + (size_t)b * p.Hq * HD + h * HD;
    __syncthreads();  // shared state of the previous unit is no longer read
    if (!c.active[b]) { if (!DEFER && split == 0 && tid < HD) outp[tid] = 0.f; continue; }
    const int n = c.pos[b] + 1;
    const int nchunk = (n + CH - 1) / CH;
    const int nsplit = min(nchunk, S);
    if (split >= nsplit) continue;
    const int hk = h / (p.Hq / p.Hkv);
    const int* bt = p.block_table + b * p.pages_per_row;
    const float* qp = p.qbuf + (size_t)b * p.Hq * HD + h * HD + sub * 8;
    const float4 q0 = ldg_cg(reinterpret_cast<const float4*>(qp));
    const float4 q1 = ldg_cg(reinterpret_cast<const float4*>(qp + 4));
    float M = -INFINITY, L = 0.f, O = 0.f;
    for (int chunk = split; chunk < nchunk; chunk += S) {
      const int tbase = chunk * CH + warp * PER_WARP + grp;
      float4 k0[ITER], k1[ITER], v0[ITER], v1[ITER];
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        const int t = tbase + 4 * i;
        k0[i] = k1[i] = v0[i] = v1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < n) {
          const int page = __ldg(bt + t / kPageTokens);
          const float* kr = kvl + kv_off(page, 0, hk, t % kPageTokens, p.Hkv, HD) + sub * 8;
          const float* vr = kvl + kv_off(page, 1, hk, t % kPageTokens, p.Hkv, HD) + sub * 8;
          k0[i] = ldg_cg(reinterpret_cast<const float4*>(kr));
          k1[i] = ldg_cg(reinterpret_cast<const float4*>(kr + 4));
          v0[i] = ldg_cg(reinterpret_cast<const float4*>(vr));
          v1[i] = ldg_cg(reinterpret_cast<const float4*>(vr + 4));
        }
      }
      float sc[ITER], m = -INFINITY;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        float s = q0.x * k0[i].x + q0.y * k0[i].y + q0.z * k0[i].z + q0.w * k0[i].w + q1.x * k1[i].x + q1.y * k1[i].y +
                  q1.z * k1[i].z + q1.w * k1[i].w;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        sc[i] = (tbase + 4 * i < n) ? s * p.scaling : -INFINITY;
        m = fmaxf(m, sc[i]);
      }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
      float l = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (m > -INFINITY) {
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
          const float e = expf(sc[i] - m);
          l += e;
          o[0] = fmaf(e, v0[i].x, o[0]); o[1] = fmaf(e, v0[i].y, o[1]); o[2] = fmaf(e, v0[i].z, o[2]);
          o[3] = fmaf(e, v0[i].w, o[3]); o[4] = fmaf(e, v1[i].x, o[4]); o[5] = fmaf(e, v1[i].y, o[5]);
          o[6] = fmaf(e, v1[i].z, o[6]); o[7] = fmaf(e, v1[i].w, o[7]);
        }
      }
      l += __shfl_xor_sync(0xffffffffu, l, 8);
      l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] += __shfl_xor_sync(0xffffffffu, o[j], 8);
        o[j] += __shfl_xor_sync(0xffffffffu, o[j], 16);
      }
      __syncthreads();
      if (lane < 8) {
        *reinterpret_cast<float4*>(&s_o[warp][lane * 8]) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(&s_o[warp][lane * 8 + 4]) = make_float4(o[4], o[5], o[6], o[7]);
        if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
      }
      __syncthreads();
      if (tid < HD) {
        float cm = M;
#pragma unroll
        for (int w = 0; w < NW; ++w) cm = fmaxf(cm, s_m[w]);
        const float fo = (M > -INFINITY) ? expf(M - cm) : 0.f;
        L *= fo; O *= fo;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const float f = (s_m[w] > -INFINITY) ? expf(s_m[w] - cm) : 0.f;
          L = fmaf(f, s_l[w], L);
          O = fmaf(f, s_o[w][tid], O);
        }
        M = cm;
      }
    }
    if (!DEFER && nsplit == 1) {
      if (tid < HD) outp[tid] = O / L;
      continue;
    }
    float* part = p.part + (((size_t)b * p.Hq + h) * p.nsplit_max + split) * (HD + 2);
    if (tid < HD) part[tid] = O;
    if (tid == 0) { part[HD] = M; part[HD + 1] = L; }
    if (DEFER) continue;  // the O-proj phase merges the splits while it stages its input (mg_attn_merge)
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&p.counter[b * p.Hq + h], 1) == nsplit - 1);
    __syncthreads();
    if (!s_last) continue;
    __threadfence();
    const float* pb = p.part + (((size_t)b * p.Hq + h) * p.nsplit_max) * (HD + 2);
    float GM = -INFINITY;
    for (int s = 0; s < nsplit; ++s) GM = fmaxf(GM, __ldcg(pb + s * (HD + 2) + HD));
    float GL = 0.f, GO = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float w = expf(__ldcg(pb + s * (HD + 2) + HD) - GM);
      GL = fmaf(w, __ldcg(pb + s * (HD + 2) + HD + 1), GL);
      if (tid < HD) GO = fmaf(w, __ldcg(pb + s * (HD + 2) + tid), GO);
    }
    if (tid < HD) outp[tid] = GO / GL;
    if (tid == 0) p.counter[b * p.Hq + h] = 0;
  }
}

// BT <= 4: merge the flash-decoding partials of every (row, head) straight into the O-proj staging buffer xs[b][h*64+d]
// (each CTA redundantly; 12 x S x 66 floats per row from L2) - no atomics / fences / last-CTA pass in the attention phase.
template <int BT>
__device__ __forceinline__ void mg_attn_merge(const MegaP& p, const MegaCtx<BT>& c, int S, float* scratch) {
  constexpr int HD = 64, PW = HD + 2;
  // 1) bulk-copy the live partials [b][h][s < S][66] into shared memory (one L2 round trip, no registers)
  const int per_row = p.Hq * S * PW;  // floats; PW even and bases 8-byte aligned -> 8-byte cp.async
  for (int i = threadIdx.x; i < p.B * per_row / 2; i += MG_THREADS) {
    const int b = i / (per_row / 2), r = i % (per_row / 2);
    const int h = r / (S * PW / 2), q = r % (S * PW / 2);
    const float* src = p.part + (((size_t)b * p.Hq + h) * p.nsplit_max) * PW + 2 * q;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(scratch + (size_t)b * per_row + 2 * r);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  }
  cp_async_wait_all();
  __syncthreads();
  // 2) merge from shared memory into the O-proj staging buffer xs[b][h*64 + d]
  for (int i = threadIdx.x; i < BT * KC; i += MG_THREADS) {
    const int b = i / KC, col = i % KC, h = col / HD, d = col % HD;
    float v = 0.f;
    if (b < p.B && c.active[b]) {
      const int n = c.pos[b] + 1;
      const int nsplit = min((n + mg_chunk(BT) - 1) / mg_chunk(BT), S);
      const float* pb = scratch + (size_t)b * per_row + (size_t)h * S * PW;
      float GM = -INFINITY;
      for (int s = 0; s < nsplit; ++s) GM = fmaxf(GM, pb[s * PW + HD]);
      float GL = 0.f, GO = 0.f;
      for (int s = 0; s < nsplit; ++s) {
        const float w = expf(pb[s * PW + HD] - GM);
        GL = fmaf(w, pb[s * PW + HD + 1], GL);
        GO = fmaf(w, pb[s * PW + d], GO);
      }
      v = GO / GL;
    }
    c.xs[i] = v;
  }
  __syncthreads();
}

// down projection (K = 3072): the CTA's 8 warps split K (384 each) for the CTA's row pairs; partials meet in
// shared memory in a fixed order.  dw[] holds this warp's weight slices, requested before the barrier.
template <int BT>
__device__ __forceinline__ void mg_down_load(const MegaP& p, const float* Wd, int lane, int warp, float4 (&dw)[MG_DOWN_PAIRS][2][3]) {
  const int npairs = p.d / 2;
#pragma unroll
  for (int j = 0; j < MG_DOWN_PAIRS; ++j) {
    const int pair = blockIdx.x + j * gridDim.x;
    if (pair < npairs) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float4* wp = reinterpret_cast<const float4*>(Wd + (size_t)(2 * pair + r) * p.I + warp * (p.I / MG_WARPS)) + lane;
#pragma unroll
        for (int i = 0; i < 3; ++i) dw[j][r][i] = ldg_stream(wp + i * 32);
      }
    }
  }
}

template <int BT>
__device__ __forceinline__ void mg_down_phase(const MegaP& p, const MegaCtx<BT>& c, float4 (&dw)[MG_DOWN_PAIRS][2][3], float rx,
                                              float* xs) {
  __shared__ float red[MG_DOWN_PAIRS][MG_WARPS][2][BT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = p.B, npairs = p.d / 2, I4 = p.I / 4;
  // stage mlp [BT][I]
  for (int i = tid; i < BT * I4; i += MG_THREADS) {
    const int b = i / I4, k4 = i % I4;
    if (b < nb) cp_async16(&xs[i * 4], p.mlp + (size_t)b * p.I + k4 * 4);
    else reinterpret_cast<float4*>(xs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();
  constexpr int LPB = 32 / BT;
  const int kq = warp * (p.I / MG_WARPS) / 4;  // float4 offset of this warp's K slice
#pragma unroll
  for (int j = 0; j < MG_DOWN_PAIRS; ++j) {
    const int pair = blockIdx.x + j * gridDim.x;
    if (pair >= npairs) break;
    float acc0[BT], acc1[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 xv = reinterpret_cast<const float4*>(xs)[b * I4 + kq + i * 32 + lane];
        acc0[b] = fmaf(dw[j][0][i].x, xv.x, acc0[b]); acc0[b] = fmaf(dw[j][0][i].y, xv.y, acc0[b]);
        acc0[b] = fmaf(dw[j][0][i].z, xv.z, acc0[b]); acc0[b] = fmaf(dw[j][0][i].w, xv.w, acc0[b]);
        acc1[b] = fmaf(dw[j][1][i].x, xv.x, acc1[b]); acc1[b] = fmaf(dw[j][1][i].y, xv.y, acc1[b]);
        acc1[b] = fmaf(dw[j][1][i].z, xv.z, acc1[b]); acc1[b] = fmaf(dw[j][1][i].w, xv.w, acc1[b]);
      }
    }
    warp_reduce_scatter<BT>(acc0);
    warp_reduce_scatter<BT>(acc1);
    if ((lane % LPB) == 0) { red[j][warp][0][lane / LPB] = acc0[0]; red[j][warp][1][lane / LPB] = acc1[0]; }
  }
  __syncthreads();
  // final: thread -> (pair j, row r, batch b); K slices summed in the order 0..7 (deterministic)
  static_assert(MG_DOWN_PAIRS * 2 * BT <= MG_THREADS, "one final-reduce element per thread");
  if (tid < MG_DOWN_PAIRS * 2 * BT) {
    const int b = tid % BT, r = (tid / BT) % 2, j = tid / (2 * BT);
    const int pair = blockIdx.x + j * gridDim.x;
    if (pair < npairs && b < nb) {
      float v = red[j][0][r][b];
#pragma unroll
      for (int w = 1; w < MG_WARPS; ++w) v = __fadd_rn(v, red[j][w][r][b]);
      p.x[(size_t)b * p.d + 2 * pair + r] = __fadd_rn(rx, v);  // rx: residual requested before the barrier
    }
  }
}

template <int BT>
__global__ void __launch_bounds__(MG_THREADS, 1) k_step(const MegaP p) {
  extern __shared__ __align__(16) float mg_smem[];
  __shared__ float s_rinv[BT];
  __shared__ int s_pos[BT], s_active[BT], s_page[BT];
  __shared__ float s_cos[BT * 64], s_sin[BT * 64];
  MegaCtx<BT> c{mg_smem, s_rinv, s_pos, s_active, s_cos, s_sin, s_page};
  float* gws = mg_smem + BT * KC;  // gate/up weight landing zone (also: attention-merge scratch, down-phase activations)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  const float* W0 = p.W + p.layer0;
  const int nqkv_rows = (p.Hq + 2 * p.Hkv) * p.hd, nqkv_tasks = nqkv_rows / 2;
  float4 w0[6], w1[6];
  {
    const int task = blockIdx.x * MG_WARPS + warp;
    if (task < nqkv_tasks) mg_load_w<MG_QKV>(p, W0 + p.o_wqkv, task, nqkv_rows, lane, w0, w1);
  }
  pdl_wait();
  if (p.decode && ldg_cg(&p.st->all_finished)) return;  // uniform over the grid: no barrier has been entered
  unsigned epoch = 0;
  int tr = 0;
#define MG_TRACE() do { if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[tr++] = globaltimer_ns(); } while (0)
  MG_TRACE();

  // ---- step input (k_input): prompt column or sum of the code embeddings; positions = tokens so far
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
  for (int i = tid; i < BT * 64; i += MG_THREADS) {  // RoPE rows of this step's positions (p.hd == 64)
    const int b = i / 64;
    s_cos[i] = b < p.B ? __ldg(p.W + p.o_cos + (size_t)s_pos[b] * 64 + (i % 64)) : 0.f;
    s_sin[i] = b < p.B ? __ldg(p.W + p.o_sin + (size_t)s_pos[b] * 64 + (i % 64)) : 0.f;
  }
  {
    const int ngen = p.decode ? ldg_cg(&p.st->n_gen) : 0;
    for (int i = tid; i < BT * KC; i += MG_THREADS) {
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
      c.xs[i] = v;
      if (blockIdx.x == 0 && b < p.B) p.x[(size_t)b * KC + k] = v;  // residual stream for the later phases
    }
    __syncthreads();
  }

  // one round of attention units: (row, head, split) <= #CTAs; each unit walks its chunks with a running softmax
  int S = max(1, min(p.nsplit_max, (int)gridDim.x / (p.Hq * p.B)));
  S = min(S, MG_SMAX);
  float4 dw[MG_DOWN_PAIRS][2][3];
  constexpr bool SMALL = true;  // O-proj merges the attention splits itself; gate/up streams through the landing zone
  for (int l = 0; l < p.L; ++l) {
    const float* Wl = W0 + (int64_t)l * p.layer_stride;
    float* kvl = p.kv + (size_t)l * p.kv_layer_floats;
    // A: QKV + RoPE + KV append   (l == 0: activations already staged from the input above)
    mg_gemv_phase<BT, MG_QKV>(p, c, Wl + p.o_wqkv, nqkv_tasks, nqkv_rows, p.x, Wl + p.o_ln1, kvl, w0, w1, l == 0);
    grid_arrive(p.bar, epoch);
    {  // next weights: O-proj (one task per warp)
      const int task = blockIdx.x * MG_WARPS + warp;
      if (task < p.d / 2) mg_load_w<MG_OPROJ>(p, Wl + p.o_wo, task, p.d, lane, w0, w1);
    }
    grid_wait(p.bar, epoch);
    MG_TRACE();
    // B: attention
    mg_attn_phase<BT, SMALL>(p, c, kvl, S);
    grid_arrive(p.bar, epoch);
    float rx0 = 0.f, rx1 = 0.f;
    {  // residual values of this warp's O-proj task (x is stable since the previous down phase / the input)
      constexpr int LPB_ = 32 / BT;
      const int task = blockIdx.x * MG_WARPS + warp, b = lane / LPB_;
      if (task < p.d / 2 && (lane % LPB_) == 0 && b < p.B) {
        rx0 = ldg_cg(p.x + (size_t)b * p.d + 2 * task);
        rx1 = ldg_cg(p.x + (size_t)b * p.d + 2 * task + 1);
      }
    }
    grid_wait(p.bar, epoch);
    MG_TRACE();
    // C: O-proj + residual
    if (blockIdx.x * MG_WARPS < p.d / 2) {  // only the CTAs that own O-proj rows stage the attention output
      if (SMALL) mg_attn_merge<BT>(p, c, S, gws);  // xs <- merged splits; the gate/up landing zone is free here
      mg_gemv_phase<BT, MG_OPROJ>(p, c, Wl + p.o_wo, p.d / 2, p.d, p.attn, nullptr, nullptr, w0, w1, SMALL, rx0, rx1);
    }
    grid_arrive(p.bar, epoch);
    if (SMALL) {
      mg_gateup_load<BT>(p, Wl + p.o_wgu, lane, warp, gws);
    } else {
      const int task = blockIdx.x * MG_WARPS + warp;
      if (task < p.I) mg_load_w<MG_GATEUP>(p, Wl + p.o_wgu, task, 2 * p.I, lane, w0, w1);
    }
    grid_wait(p.bar, epoch);
    MG_TRACE();
    // D: gate/up + SiLU*mul
    if (SMALL) mg_gateup_small<BT>(p, c, Wl + p.o_ln2, gws);
    else mg_gemv_phase<BT, MG_GATEUP>(p, c, Wl + p.o_wgu, p.I, 2 * p.I, p.x, Wl + p.o_ln2, nullptr, w0, w1, false);
    grid_arrive(p.bar, epoch);
    mg_down_load<BT>(p, Wl + p.o_wd, lane, warp, dw);
    float rxd = 0.f;
    if (tid < MG_DOWN_PAIRS * 2 * BT) {  // residual of this thread's final-reduce element (stable since the O-proj phase)
      const int b = tid % BT, r = (tid / BT) % 2, j = tid / (2 * BT);
      const int pair = blockIdx.x + j * gridDim.x;
      if (pair < p.d / 2 && b < p.B) rxd = ldg_cg(p.x + (size_t)b * p.d + 2 * pair + r);
    }
    grid_wait(p.bar, epoch);
    MG_TRACE();
    // E: down + residual
    mg_down_phase<BT>(p, c, dw, rxd, gws);  // [BT][3072] activations staged in the (now free) landing zone
    grid_arrive(p.bar, epoch);
    if (l + 1 < p.L) {
      const int task = blockIdx.x * MG_WARPS + warp;
      if (task < nqkv_tasks) mg_load_w<MG_QKV>(p, Wl + p.layer_stride + p.o_wqkv, task, nqkv_rows, lane, w0, w1);
    } else if (p.sample) {
      const int task = blockIdx.x * MG_WARPS + warp;
      if (task < (p.rows_per_item * p.V + 1) / 2)
        mg_load_w<MG_HEADS>(p, p.W + p.o_head, task, p.rows_per_item * p.V, lane, w0, w1);
    }
    grid_wait(p.bar, epoch);
    MG_TRACE();
  }
  if (p.sample) {
    const int nrows = p.rows_per_item * p.V;
    mg_gemv_phase<BT, MG_HEADS>(p, c, p.W + p.o_head, (nrows + 1) / 2, nrows, p.x, p.W + p.o_final_norm, nullptr, w0, w1, false);
  }
  MG_TRACE();
  // positions advance once per step (k_input did this in the multi-kernel path)
  if (blockIdx.x == 0 && tid < p.B && s_active[tid]) p.seq_len[tid] = s_pos[tid] + 1;
}

}  // namespace ctb
