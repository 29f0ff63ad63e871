// Fused sampling tail (reference gpt.py:487-525 + processors.py:18-58 + HF TopP/TopK warpers).
// One CTA of 1024 threads per logits row.
#include "gpt_kernels.cuh"

namespace ctb {

namespace {

__device__ __forceinline__ double block_sum_d(double v, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll 8
  for (int w = 0; w < SAMPLE_THREADS / 32; ++w) t += s_red[w];  // fixed order => deterministic
  return t;
}

__device__ __forceinline__ int block_sum_i(int v, int* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  int t = 0;
#pragma unroll 8
  for (int w = 0; w < SAMPLE_THREADS / 32; ++w) t += s_red[w];
  return t;
}

__device__ __forceinline__ float block_max_f(float v, float* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = -INFINITY;
#pragma unroll 8
  for (int w = 0; w < SAMPLE_THREADS / 32; ++w) t = fmaxf(t, s_red[w]);
  return t;
}

}  // namespace

__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample(const SampleP p) {
  pdl_trigger();
  pdl_wait();
  if (p.check_finished && ldg_cg(&p.st->all_finished)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_x = reinterpret_cast<float*>(smem_raw);                 // [V] processed logits
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_x + p.V);        // [2][1024] sort path only
  __shared__ double s_redd[SAMPLE_THREADS / 32];
  __shared__ int s_redi[SAMPLE_THREADS / 32];
  __shared__ float s_redf[SAMPLE_THREADS / 32];
  __shared__ int s_win[32];
  __shared__ uint32_t s_thr;

  const int tid = threadIdx.x;
  const int row = blockIdx.x, V = p.V, rpi = p.rows_per_item;
  const int item = row / rpi, qi = row % rpi;
  const int n_gen = p.st ? ldg_cg(&p.st->n_gen) : p.n_gen_fixed;
  const int step = p.st ? ldg_cg(&p.st->step) : p.step_fixed;
  const ctb_sampler_config& c = p.cfg;
  const float* lg = p.logits + (size_t)row * V;

  // ---- S2 window of the last <= past_window generated ids of this (item, codebook) row
  int nwin = 0;
  const bool pen = c.penalty_on && row < c.penalty_max_ids;
  if (pen) {
    nwin = min(n_gen, c.past_window);
    if (tid < nwin) s_win[tid] = ldg_cg(&p.gen_ids[((size_t)item * p.gen_stride + (n_gen - nwin + tid)) * p.gen_inner + qi]);
  }
  __syncthreads();
  // ---- S1 temperature, S2 penalty
  const float temp = c.temperature[qi];
  for (int v = tid; v < V; v += SAMPLE_THREADS) {
    float x = __fdiv_rn(ldg_cg(&lg[v]), temp);
    if (pen) {
      int cnt = 0;
      for (int w = 0; w < nwin; ++w) cnt += (s_win[w] == v);
      const float a = c.penalty_lut[cnt];
      x = (x < 0.f) ? __fmul_rn(x, a) : __fdiv_rn(x, a);
    }
    s_x[v] = x;
  }
  __syncthreads();

  // ---- row max and softmax denominator of the unfiltered row (top-p's own softmax)
  float mx = -INFINITY;
  for (int v = tid; v < V; v += SAMPLE_THREADS) mx = fmaxf(mx, s_x[v]);
  mx = block_max_f(mx, s_redf);

  const bool use_p = c.top_p >= 0.f;
  const int kk = c.top_k > 0 ? min(max(c.top_k, c.min_tokens_to_keep), V) : 0;
  const int min_keep = min(c.min_tokens_to_keep, V);
  uint32_t thr_key = 0;  // keep x iff float_key(x) >= thr_key

  if (use_p || kk > 0) {
    double den = 0.0;
    if (use_p) {
      for (int v = tid; v < V; v += SAMPLE_THREADS) den += (double)expf(s_x[v] - mx);
      den = block_sum_d(den, s_redd);
    }
    const float denf = (float)den;
    const float pthr = c.has_removed_max ? c.top_p_removed_max : (float)(1.0 - (double)c.top_p);  // `cum <= (1 - top_p)` evaluated in fp32
    if (V <= 1024) {
      // ---------- sort path: bitonic sort of 1024 keys (pads = 0 sort first).  One key per thread in a register;
      // compare-exchange distances < 32 are warp shuffles, the 15 longer ones go through two alternating
      // shared-memory buffers (one barrier each instead of 55 barriers for an all-shared-memory network).
      uint32_t key = tid < V ? float_key(s_x[tid]) : 0u;
      int sb = 0;
      for (int k = 2; k <= 1024; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          uint32_t other;
          if (j >= 32) {
            uint32_t* buf = s_key + sb * 1024;
            buf[tid] = key;
            __syncthreads();
            other = buf[tid ^ j];
            sb ^= 1;
          } else {
            other = __shfl_xor_sync(0xffffffffu, key, j);
          }
          const bool up = (tid & k) == 0, lower = (tid & j) == 0;
          key = (lower == up) ? min(key, other) : max(key, other);
        }
      }
      __syncthreads();
      s_key[tid] = key;
      __syncthreads();
      uint32_t t_p = 0;
      if (use_p) {
        // inclusive scan (double, like ATen's CPU cumsum) of softmax(sorted) ascending
        const uint32_t key = s_key[tid];
        double pv = key ? (double)__fdiv_rn(expf(key_float(key) - mx), denf) : 0.0;
        const int lane = tid & 31, warp = tid >> 5;
        double inc = pv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double n = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += n;
        }
        if (lane == 31) s_redd[warp] = inc;
        __syncthreads();
        double base = 0.0;
        for (int w = 0; w < warp; ++w) base += s_redd[w];
        const float cum = (float)(base + inc);
        const int removed = (cum <= pthr) && (tid < 1024 - min_keep);
        const int nrem = __syncthreads_count(removed);  // removed set is a prefix of the sorted row
        t_p = s_key[nrem];
      }
      const uint32_t t_k = kk > 0 ? s_key[1024 - kk] : 0u;
      thr_key = max(t_p, t_k);
    } else {
      // ---------- search path (text head, V = 21178): bisection on the key space
      uint32_t t_p = 0;
      if (use_p) {
        // smallest key t with float(sum_{key_j <= t} p_j) > pthr
        uint32_t lo = 0u, hi = 0xffffffffu;
        while (lo < hi) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          double s = 0.0;
          for (int v = tid; v < V; v += SAMPLE_THREADS) {
            const float x = s_x[v];
            if (float_key(x) <= mid) s += (double)__fdiv_rn(expf(x - mx), denf);
          }
          s = block_sum_d(s, s_redd);
          if ((float)s > pthr) hi = mid; else lo = mid + 1;
        }
        t_p = lo;
        // always keep the min_keep largest: largest t with count(key >= t) >= min_keep
        uint32_t lo2 = 0u, hi2 = 0xffffffffu;
        while (lo2 < hi2) {
          const uint32_t mid = lo2 + (uint32_t)(((uint64_t)hi2 - lo2 + 1) >> 1);
          int cnt = 0;
          for (int v = tid; v < V; v += SAMPLE_THREADS) cnt += (float_key(s_x[v]) >= mid);
          cnt = block_sum_i(cnt, s_redi);
          if (cnt >= min_keep) lo2 = mid; else hi2 = mid - 1;
        }
        t_p = min(t_p, lo2);
      }
      uint32_t t_k = 0;
      if (kk > 0) {
        uint32_t lo2 = 0u, hi2 = 0xffffffffu;
        while (lo2 < hi2) {
          const uint32_t mid = lo2 + (uint32_t)(((uint64_t)hi2 - lo2 + 1) >> 1);
          int cnt = 0;
          for (int v = tid; v < V; v += SAMPLE_THREADS) cnt += (float_key(s_x[v]) >= mid);
          cnt = block_sum_i(cnt, s_redi);
          if (cnt >= kk) lo2 = mid; else hi2 = mid - 1;
        }
        t_k = lo2;
      }
      thr_key = max(t_p, t_k);
    }
  }
  if (c.greedy) {
    // bench config C2: keep only the row arg-max; greedy == 2 takes it over the non-EOS tokens so
    // that the EOS ban below can never empty the row
    float gm = -INFINITY;
    for (int v = tid; v < V; v += SAMPLE_THREADS)
      if (!(c.greedy == 2 && v == c.eos_token)) gm = fmaxf(gm, s_x[v]);
    gm = block_max_f(gm, s_redf);
    thr_key = float_key(gm);
  }
  if (tid == 0) s_thr = thr_key;
  __syncthreads();
  thr_key = s_thr;

  // ---- EOS ban (gpt.py:494-495), final softmax (gpt.py:497), argmax(p / q) (gpt.py:501-508)
  const bool ban = step < c.min_new_token;
  float mx2 = -INFINITY;
  for (int v = tid; v < V; v += SAMPLE_THREADS) {
    float x = s_x[v];
    if (float_key(x) < thr_key || ((ban || c.greedy == 2) && v == c.eos_token)) x = -INFINITY;
    s_x[v] = x;
    mx2 = fmaxf(mx2, x);
  }
  mx2 = block_max_f(mx2, s_redf);
  double den2 = 0.0;
  for (int v = tid; v < V; v += SAMPLE_THREADS) den2 += (double)expf(s_x[v] - mx2);
  den2 = block_sum_d(den2, s_redd);
  const float den2f = (float)den2;

  float best = -1.f;
  int besti = 0x7fffffff;
  for (int v = tid; v < V; v += SAMPLE_THREADS) {
    const float pr = __fdiv_rn(expf(s_x[v] - mx2), den2f);
    const float qn = p.q_noise ? p.q_noise[(size_t)row * V + v]
                               : philox_exp1(c.philox_seed, (uint32_t)row, (uint32_t)v, (uint32_t)step);
    const float r = __fdiv_rn(pr, qn);
    if (r > best) { best = r; besti = v; }  // ascending v within a thread: first max wins
  }
  // block arg-max, lowest index on ties (ATen argmax returns the first maximum)
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  __shared__ float s_bv[SAMPLE_THREADS / 32];
  __shared__ int s_bi[SAMPLE_THREADS / 32];
  __syncthreads();
  if ((tid & 31) == 0) { s_bv[tid >> 5] = best; s_bi[tid >> 5] = besti; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < SAMPLE_THREADS / 32; ++w)
      if (s_bv[w] > best || (s_bv[w] == best && s_bi[w] < besti)) { best = s_bv[w]; besti = s_bi[w]; }
    p.out_idx[row] = besti < V ? besti : 0;  // all-NaN row: ATen argmax returns the first index
  }
}

// finish / write-back / counters (gpt.py:512-525,572-577).  One CTA, one thread per batch row.
__global__ void k_finalize(const FinalP p) {
  pdl_trigger();
  pdl_wait();
  if (ldg_cg(&p.st->all_finished)) return;
  __shared__ int s_any, s_notall;
  if (threadIdx.x == 0) { s_any = 0; s_notall = 0; }
  __syncthreads();
  const int n = ldg_cg(&p.st->n_gen);
  const int step0 = ldg_cg(&p.st->step);
  for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
    bool eos = false;
    for (int q = 0; q < p.rows_per_item; ++q) eos |= (ldg_cg(&p.idx[b * p.rows_per_item + q]) == p.eos);
    const bool fin = ldg_cg(&p.finish[b]) || eos;
    p.finish[b] = fin ? 1 : 0;
    int32_t* dst = p.ids_out + ((size_t)b * p.max_new + n) * p.num_vq;
    for (int q = 0; q < p.num_vq; ++q) dst[q] = ldg_cg(&p.idx[b * p.rows_per_item + (p.rows_per_item == 1 ? 0 : q)]);
    if (fin) atomicOr(&s_any, 1); else { atomicOr(&s_notall, 1); }
    // gpt.py:527: at i == 0 with any finished row the reference returns before end_idx moves
    (void)0;
  }
  __syncthreads();
  const bool first_abort = (step0 == 0) && s_any;
  if (!first_abort)
    for (int b = threadIdx.x; b < p.B; b += blockDim.x)
      if (!p.finish[b]) p.end_idx[b] = ldg_cg(&p.end_idx[b]) + 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (first_abort) { p.st->any_first = 1; p.st->all_finished = 1; }
    else if (!s_notall) p.st->all_finished = 1;
    p.st->n_gen = n + 1;
    p.st->step = step0 + 1;
  }
}

}  // namespace ctb
