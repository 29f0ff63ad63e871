// Decode-step GEMMs on tcgen05 tensor cores (swap-AB, 3xTF32, split-K over a thread-block cluster).
//
//   Y[b, r] = sum_k W[r, k] * x[b, k]        W: [rows, K] fp32 streamed from HBM exactly once by TMA
//
//   * swap-AB: the 128 weight rows of a tile are the UMMA M dimension, the (padded) batch is N = NPAD.
//   * fp32-equivalent accuracy for the bit-exact-ids contract: W = W_hi + W_lo is split in shared memory by
//     the worker warps (cvt.rna.tf32; weights cannot be pre-split without doubling HBM traffic), the
//     activations arrive pre-split (x_hi, x_lo written by the producing kernel's epilogue).  Two MMAs per
//     k-step:  W_hi x [x_hi ; x_lo] (N = 2*NPAD, one pass over the W_hi tile) and W_lo x x_hi (N = NPAD).
//   * split-K: the CS CTAs of a cluster own consecutive K slices of the same 128 rows, so 6..48 row tiles
//     still put ~100+ SMs on the HBM stream; partial accumulators meet through distributed shared memory
//     in a fixed order (deterministic), and the rank that owns a row slice runs the fused epilogue:
//     RMSNorm scale (norm weight folded into W at load, 1/rms applied here), RoPE + paged-KV append,
//     residual add, SiLU*up, logits.  Row pairs that must meet in one thread (RoPE j / j+32, gate_n / up_n)
//     are made adjacent by a row permutation applied once when the tensor-core weight copy is built.
//   * 4-stage TMA/mbarrier ring; weight tiles of the first stages are requested BEFORE griddepcontrol.wait
//     (PDL), activations after it.
#pragma once
#include "gpt_kernels.cuh"
#include "tc_common.cuh"

namespace ctb {

enum DecEpi { DE_QKV = 0, DE_OPROJ = 1, DE_GATEUP = 2, DE_DOWN = 3, DE_HEADS = 4 };

constexpr int TD_STAGES = 3;  // 3 x 36-40 KiB: two CTAs (this kernel + its PDL successor) fit one SM
constexpr int TD_THREADS = 192;
constexpr int TD_A_BYTES = 128 * 32 * 4;  // 16 KiB weight tile (128 rows x 32 k)

template <int NPAD>
struct TdCfg {
  static constexpr int X_BYTES = NPAD * 128;                       // one x tile (NPAD rows x 32 k)
  static constexpr int STAGE_BYTES = 2 * TD_A_BYTES + 2 * X_BYTES;  // W | W_lo | x_hi | x_lo
  static constexpr int SMEM_BYTES = TD_STAGES * STAGE_BYTES + 1024 + 512;
};

struct TcDecP {
  int K, kslice, nrows, B;
  const float* xraw;      // [Bpad][K] raw residual rows for 1/rms (nullptr: no norm)
  float eps;
  float* xres; float* x_hi; float* x_lo; int d;                      // OPROJ / DOWN
  float* qbuf; float* kv; const int* block_table; int pages_per_row;  // QKV
  const int* pos; const uint8_t* active; const float* rope_cos; const float* rope_sin;
  int Hq, Hkv, hd;
  float* h_hi; float* h_lo; int I;                                    // GATEUP
  float* logits; int rows_per_item, V;                                // HEADS
  float* hidden_out; int hidden_stride; const float* final_norm_w; const LoopState* st;
};

template <int EPI, int NPAD, int CS>
__global__ void __launch_bounds__(TD_THREADS, 1)
k_tc_dec(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_xhi,
         const __grid_constant__ CUtensorMap map_xlo, const TcDecP p) {
  using Cfg = TdCfg<NPAD>;
  pdl_trigger();
  extern __shared__ uint8_t td_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(td_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TD_STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* split = bars + TD_STAGES;
  uint64_t* empty = bars + 2 * TD_STAGES;
  uint64_t* accum_full = bars + 3 * TD_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TD_STAGES + 1);
  float* s_rinv = reinterpret_cast<float*>(bars + 3 * TD_STAGES + 2);  // [NPAD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const int tile = blockIdx.x / CS;
  const int r_tile = tile * 128;
  const int kbase = rank * p.kslice;
  const int nk = p.kslice / 32;
  constexpr uint32_t TMEM_COLS = (2 * NPAD <= 32) ? 32 : 64;

  if (threadIdx.x == 128) {  // TMA warp: hide the descriptor fetch latency
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xlo) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < TD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer
    if (lane == 0) {
      const int pre = min(nk, TD_STAGES);
      for (int t = 0; t < pre; ++t) {  // weights do not depend on earlier kernels
        uint8_t* st = smem + t * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[t], TD_A_BYTES + 2 * Cfg::X_BYTES);
        tma_load_2d(st, &map_w, &full[t], kbase + t * 32, r_tile);
      }
      pdl_wait();  // activations below were written by the previous kernel
      for (int t = 0; t < pre; ++t) {
        uint8_t* st = smem + t * Cfg::STAGE_BYTES;
        tma_load_2d(st + 2 * TD_A_BYTES, &map_xhi, &full[t], kbase + t * 32, 0);
        tma_load_2d(st + 2 * TD_A_BYTES + Cfg::X_BYTES, &map_xlo, &full[t], kbase + t * 32, 0);
      }
      for (int t = pre; t < nk; ++t) {
        const int s = t % TD_STAGES, it = t / TD_STAGES;
        mbar_wait(&empty[s], (it - 1) & 1);
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], TD_A_BYTES + 2 * Cfg::X_BYTES);
        tma_load_2d(st, &map_w, &full[s], kbase + t * 32, r_tile);
        tma_load_2d(st + 2 * TD_A_BYTES, &map_xhi, &full[s], kbase + t * 32, 0);
        tma_load_2d(st + 2 * TD_A_BYTES + Cfg::X_BYTES, &map_xlo, &full[s], kbase + t * 32, 0);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer
    const uint32_t idesc_wide = umma_idesc_tf32(128, 2 * NPAD), idesc_narrow = umma_idesc_tf32(128, NPAD);
    for (int t = 0; t < nk; ++t) {
      const int s = t % TD_STAGES, it = t / TD_STAGES;
      mbar_wait(&split[s], it & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t w_hi = st, w_lo = st + TD_A_BYTES, x_hl = st + 2 * TD_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ko = k * 32;
          // cols [0,NPAD) += W_hi x_hi ; cols [NPAD,2NPAD) += W_hi x_lo   (x_hi | x_lo tiles are contiguous)
          umma_tf32(tmem_base, umma_desc_sw128(w_hi + ko), umma_desc_sw128(x_hl + ko), idesc_wide, (t | k) ? 1u : 0u);
          umma_tf32(tmem_base, umma_desc_sw128(w_lo + ko), umma_desc_sw128(x_hl + ko), idesc_narrow, 1u);
        }
        umma_commit(&empty[s]);
        if (t == nk - 1) umma_commit(accum_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== workers
    pdl_wait();
    if (p.xraw != nullptr) {  // 1/rms of the raw residual rows (HF LlamaRMSNorm statistics)
      // rows b = warp, warp + 4, ...: all loads of all rows are issued before the first reduction
      constexpr int RPW = NPAD / 4;
      float ss[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int b = warp + 4 * i;
        ss[i] = 0.f;
        if (b < p.B) {
          const float4* xr = reinterpret_cast<const float4*>(p.xraw + (size_t)b * p.K);
#pragma unroll
          for (int k4 = 0; k4 < 6; ++k4) {  // K == 768 on every normed GEMM: 6 float4 per lane
            const float4 v = ldg_cg(xr + k4 * 32 + lane);
            ss[i] = fmaf(v.x, v.x, ss[i]); ss[i] = fmaf(v.y, v.y, ss[i]);
            ss[i] = fmaf(v.z, v.z, ss[i]); ss[i] = fmaf(v.w, v.w, ss[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const float t = warp_sum(ss[i]);
        if (lane == 0) s_rinv[warp + 4 * i] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(t, (float)p.K), p.eps)));
      }
      if (EPI == DE_HEADS && p.hidden_out != nullptr && blockIdx.x == 0) {
        // last_hidden_state (gpt.py:430-436): w * (x * rinv), written once
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int step = ldg_cg(&p.st->n_gen);
        for (int i = threadIdx.x; i < p.B * p.K; i += 128) {
          const int b = i / p.K, k = i % p.K;
          p.hidden_out[(size_t)b * p.hidden_stride + (size_t)step * p.K + k] =
              __fmul_rn(__ldg(p.final_norm_w + k), __fmul_rn(ldg_cg(p.xraw + i), s_rinv[b]));
        }
      }
    }
    for (int t = 0; t < nk; ++t) {
      const int s = t % TD_STAGES, it = t / TD_STAGES;
      mbar_wait(&full[s], it & 1);
      float4* a = reinterpret_cast<float4*>(smem + s * Cfg::STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem + s * Cfg::STAGE_BYTES + TD_A_BYTES);
#pragma unroll
      for (int j = 0; j < TD_A_BYTES / 16 / 128; ++j) {
        const int i = threadIdx.x + 128 * j;
        const float4 v = a[i];
        float4 h, l;
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
        l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
        a[i] = h; lo[i] = l;
      }
      fence_async_smem();
      mbar_arrive(&split[s]);
    }
    // ---- accumulator -> this CTA's partial tile in shared memory [128 rows][NPAD]
    mbar_wait(accum_full, 0);
    tc_fence_after();
    float* part = reinterpret_cast<float*>(smem);  // stage 0 is free: every MMA has retired
    {
      uint32_t r[2 * NPAD];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      if (NPAD == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
            "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
      } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t* q = r + 32 * half;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
              "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
                "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
                "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
                "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
              : "r"(taddr + 32 * half));
        }
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + lane;
#pragma unroll
      for (int b = 0; b < NPAD; ++b)
        part[row * NPAD + b] = __uint_as_float(r[b]) + __uint_as_float(r[NPAD + b]);  // x_hi and x_lo halves
    }
    tc_fence_before();
  }

  // ===================== split-K merge through DSMEM + fused epilogue (worker warps)
  if (CS > 1) cluster_sync_all(); else __syncthreads();
  if (warp < 4) {
    const float* part = reinterpret_cast<const float*>(smem);
    constexpr int RPR = 128 / CS;  // rows owned by this rank
    const int nitems = (RPR / 2) * p.B;
    for (int item = threadIdx.x; item < nitems; item += 128) {
      const int b = item % p.B, pi = item / p.B;
      const int lr = rank * RPR + 2 * pi;  // local row of the pair
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int rk = 0; rk < CS; ++rk) {  // fixed order: K slice 0, 1, ... (deterministic)
        if (CS > 1) {
          v0 = __fadd_rn(v0, ld_dsmem(part + lr * NPAD + b, rk));
          v1 = __fadd_rn(v1, ld_dsmem(part + (lr + 1) * NPAD + b, rk));
        } else {
          v0 = part[lr * NPAD + b]; v1 = part[(lr + 1) * NPAD + b];
        }
      }
      const int R = r_tile + lr;  // global (permuted) weight row of the pair's first member
      if (R >= p.nrows) continue;
      const float rinv = p.xraw ? s_rinv[b] : 1.0f;
      if (EPI == DE_QKV) {
        if (!ldg_cg(&p.active[b])) continue;
        const int nq = p.Hq * p.hd, nkv = p.Hkv * p.hd, half = p.hd / 2;
        const float y0 = __fmul_rn(v0, rinv), y1 = __fmul_rn(v1, rinv);
        const int pos = ldg_cg(&p.pos[b]);
        const int which = R < nq ? 0 : (R < nq + nkv ? 1 : 2);
        const int rr = R - (which == 0 ? 0 : (which == 1 ? nq : nq + nkv));
        const int h = rr / p.hd, pp = rr % p.hd;  // pp even: RoPE pair (j, j + hd/2) sits at (2j, 2j+1)
        float o0 = y0, o1 = y1;
        if (which < 2) {
          const int j = pp / 2;
          const float c0 = __ldg(p.rope_cos + (size_t)pos * p.hd + j), s0 = __ldg(p.rope_sin + (size_t)pos * p.hd + j);
          const float c1 = __ldg(p.rope_cos + (size_t)pos * p.hd + j + half), s1 = __ldg(p.rope_sin + (size_t)pos * p.hd + j + half);
          o0 = __fadd_rn(__fmul_rn(y0, c0), __fmul_rn(-y1, s0));
          o1 = __fadd_rn(__fmul_rn(y1, c1), __fmul_rn(y0, s1));
        }
        if (which == 0) {
          *reinterpret_cast<float2*>(p.qbuf + (size_t)b * nq + h * p.hd + pp) = make_float2(o0, o1);
        } else {
          const int page = __ldg(p.block_table + b * p.pages_per_row + pos / kPageTokens);
          float* dst = p.kv + kv_off(page, which - 1, h, pos % kPageTokens, p.Hkv, p.hd);
          *reinterpret_cast<float2*>(dst + pp) = make_float2(o0, o1);
        }
      } else if (EPI == DE_OPROJ || EPI == DE_DOWN) {
        float* xr = p.xres + (size_t)b * p.d + R;
        const float n0 = __fadd_rn(ldg_cg(xr), v0), n1 = __fadd_rn(ldg_cg(xr + 1), v1);
        *reinterpret_cast<float2*>(xr) = make_float2(n0, n1);
        const float h0 = to_tf32(n0), h1 = to_tf32(n1);
        *reinterpret_cast<float2*>(p.x_hi + (size_t)b * p.d + R) = make_float2(h0, h1);
        *reinterpret_cast<float2*>(p.x_lo + (size_t)b * p.d + R) = make_float2(to_tf32(n0 - h0), to_tf32(n1 - h1));
      } else if (EPI == DE_GATEUP) {
        const float g = __fmul_rn(v0, rinv), u = __fmul_rn(v1, rinv);  // rows (2n, 2n+1) = (gate_n, up_n)
        const float hv = __fmul_rn(__fdiv_rn(g, __fadd_rn(1.0f, expf(-g))), u);
        const float hh = to_tf32(hv);
        p.h_hi[(size_t)b * p.I + R / 2] = hh;
        p.h_lo[(size_t)b * p.I + R / 2] = to_tf32(hv - hh);
      } else {  // DE_HEADS
        const int q0 = R / p.V, c0 = R % p.V;
        p.logits[((size_t)b * p.rows_per_item + q0) * p.V + c0] = __fmul_rn(v0, rinv);
        if (R + 1 < p.nrows) {
          const int q1 = (R + 1) / p.V, c1 = (R + 1) % p.V;
          p.logits[((size_t)b * p.rows_per_item + q1) * p.V + c1] = __fmul_rn(v1, rinv);
        }
      }
    }
  }
  if (CS > 1) cluster_sync_all(); else __syncthreads();  // partial tiles stay alive until every rank has read them
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---- one-time construction of the tensor-core weight copy (device side, from the fp32 blob)
// out[(perm(r)) * K + k] = W[r * K + k] * (scale ? scale[k] : 1)
//   mode 0: identity rows; mode 1: q/k heads -> RoPE pairs adjacent (row j -> 2j, row j+hd/2 -> 2j+1), v rows
//   unchanged; mode 2: [gate; up] -> interleaved (gate_n -> 2n, up_n -> 2n+1)
__global__ void k_build_tc_weight(const float* __restrict__ W, const float* __restrict__ scale, float* __restrict__ out,
                                  int rows, int K, int mode, int qk_rows, int hd, int I);

}  // namespace ctb
