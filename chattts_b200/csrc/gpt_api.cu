// extern "C" entry points for hot path 1 (see include/chattts_b200.h).
#include <stdarg.h>

#include <map>
#include <mutex>
#include <vector>

#include <cudaTypedefs.h>

#define CTB_GPT_KERNELS_IMPL
#include "gpt_kernels.cuh"
#include "tc_decode.cuh"
#include "mega.cuh"
#include "flow.cuh"
#include "prefill.cuh"
#include "tc_gemm.cuh"

namespace ctb {

thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int g_num_sms = 148;


static int bt_for(int B) {
  int bt = 1;
  while (bt < B && bt < 32) bt <<= 1;
  return bt;
}

}  // namespace ctb

using namespace ctb;
using ctb::g_num_sms;

struct ctb_gpt {
  ctb_gpt_config cfg;
  ctb_gpt_layout lay;
  const float* W;
  int pages_per_row, nsplit_max, bpad_max;
  // device state
  float *x, *qbuf, *attn, *mlp, *logits, *kv, *part;
  int *block_table, *seq_len, *pos, *counter, *end_idx;
  int32_t* idx;
  uint8_t *active, *finish;
  LoopState* st;
  size_t kv_layer_floats;
  size_t kv_pages;      // pages per layer currently in the pool
  int bt_B; size_t bt_per_row;  // shape the uploaded block table was built for
  // per generate() call
  int B, T0, max_new, infer_text, started;
  ctb_sampler_config sampler;
  const float* q_noise;
  const float* emb;
  const uint8_t* mask;
  int32_t* ids_out;
  float* hiddens_out;
  cudaGraphExec_t graph_exec;
  uint64_t graph_kernels;  // kernel nodes captured in graph_exec
  cudaStream_t cap_stream;
  // ---- tensor-core decode path (tc_decode.cuh)
  bool use_tc, tc_ready;
  int tc_min_batch;
  // ---- batched prefill (prefill.cuh): lazily allocated
  bool pf_enabled;
  float *gw_hi, *gw_lo;            // tf32 hi / lo copies of the per-layer weight region of the blob
  float *pf_resid, *pf_xn, *pf_qkv, *pf_q, *pf_attn, *pf_gu, *pf_h, *pf_ones, *pf_zeros;
  int *pf_npre, *pf_nvalid;
  size_t pf_rows;                  // capacity (B * T0) of the pf_* activation buffers
  bool mega_ok;      // one-kernel decode step (mega.cuh), built for B <= 8
  int mega_max_batch; // batches that use it (default 4: measured faster up to there; CTB_MEGA_MAX_BATCH overrides)
  unsigned* bar;     // its grid-barrier counter
  unsigned long long* trace;  // CTB_MEGA_TRACE=1: per-phase timestamps of the last step
  bool flow_ok;      // dataflow decode step (flow.cuh), B <= 4: default back end for those batches
  int flow_R;        // replicas of the broadcast exchange regions (CTB_FLOW_R)
  int flow_l2_ahead;  // weight tasks prefetched one layer ahead into L2 (CTB_FLOW_L2_AHEAD)
  int flow_max_batch; // batches that use it (CTB_FLOW_MAX_BATCH, default 4)
  unsigned long long* flow_arena;
  unsigned* flow_epoch;
  int steps_enqueued;  // loop iterations enqueued since ctb_gpt_begin (host-side bound for ctb_gpt_decode)
  float *tc_wqkv, *tc_wgu, *tc_heads_code, *tc_heads_text;  // permuted / norm-folded weight copies
  float *x_hi, *x_lo, *attn_hi, *attn_lo, *h_hi, *h_lo;      // [32][K] tf32-split activations
  CUtensorMap *m_wqkv, *m_wo, *m_wgu, *m_wd;                 // [layers] host arrays
  CUtensorMap m_hcode, m_htext;
  CUtensorMap m_x[2][2], m_attn[2][2], m_h[2][2];            // [npad16|32][hi|lo]
  bool use_graph;
};

extern "C" int ctb_abi_version(void) { return CTB_ABI_VERSION; }
extern "C" const char* ctb_last_error(void) { return g_err; }
extern "C" uint64_t ctb_launch_count(void) { return g_launches.load(); }

extern "C" int ctb_gpt_layout_query(const ctb_gpt_config* c, ctb_gpt_layout* o) {
  if (!c || !o) return set_err(CTB_ERR_ARG, "null argument");
  const int64_t d = c->hidden_size, I = c->intermediate_size, hd = c->head_dim;
  const int64_t nq = (int64_t)c->num_heads * hd, nkv = (int64_t)c->num_kv_heads * hd;
  int64_t off = 0;
  o->wqkv = off; off += (nq + 2 * nkv) * d;
  o->wo = off; off += d * nq;
  o->wgate_up = off; off += 2 * I * d;
  o->wdown = off; off += d * I;
  o->ln1 = off; off += d;
  o->ln2 = off; off += d;
  o->layer_stride = off;
  o->layer0 = 0;
  off = o->layer_stride * c->num_layers;
  o->final_norm = off; off += d;
  o->head_code = off; off += (int64_t)c->num_vq * c->num_audio_tokens * d;
  o->head_text = off; off += (int64_t)c->num_text_tokens * d;
  o->emb_code = off; off += (int64_t)c->num_vq * c->num_audio_tokens * d;
  o->emb_text = off; off += (int64_t)c->num_text_tokens * d;
  o->rope_cos = off; off += (int64_t)c->max_positions * hd;
  o->rope_sin = off; off += (int64_t)c->max_positions * hd;
  o->total = off;
  return CTB_OK;
}

template <typename T>
static int dalloc(T** p, size_t n) {
  CTB_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  CTB_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
  return CTB_OK;
}

namespace ctb {
__global__ void k_build_tc_weight(const float* __restrict__ W, const float* __restrict__ scale, float* __restrict__ out,
                                  int rows, int K, int mode, int qk_rows, int hd, int I) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  int pr = r;
  if (mode == 1 && r < qk_rows) {
    const int h = r / hd, j = r % hd, half = hd / 2;
    pr = h * hd + (j < half ? 2 * j : 2 * (j - half) + 1);
  } else if (mode == 2) {
    pr = (r < I) ? 2 * r : 2 * (r - I) + 1;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    out[(size_t)pr * K + k] = scale ? __fmul_rn(W[(size_t)r * K + k], scale[k]) : W[(size_t)r * K + k];
}
}  // namespace ctb

static int encode_map_2d(CUtensorMap* m, const float* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return set_err(CTB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  }
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {K * 4};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(CTB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CTB_OK;
}

template <int EPI, int NPAD, int CS>
static int set_tc_attr() {
  return ensure_smem_attr((const void*)k_tc_dec<EPI, NPAD, CS>, TdCfg<NPAD>::SMEM_BYTES);
}

// cluster sizes of the split-K (K slices per 128-row tile)
constexpr int CS_QKV = 4, CS_O = 8, CS_GU = 2, CS_DOWN = 8, CS_HEADS = 4;

static int tc_setup(ctb_gpt* h) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  const size_t d = c.hidden_size, I = c.intermediate_size;
  const size_t nqkv = (size_t)(c.num_heads + 2 * c.num_kv_heads) * c.head_dim;
  int rc;
  if ((rc = dalloc(&h->tc_wqkv, (size_t)c.num_layers * nqkv * d))) return rc;
  if ((rc = dalloc(&h->tc_wgu, (size_t)c.num_layers * 2 * I * d))) return rc;
  if ((rc = dalloc(&h->tc_heads_code, (size_t)c.num_vq * c.num_audio_tokens * d))) return rc;
  if ((rc = dalloc(&h->tc_heads_text, (size_t)c.num_text_tokens * d))) return rc;
  if ((rc = dalloc(&h->x_hi, 32 * d))) return rc;
  if ((rc = dalloc(&h->x_lo, 32 * d))) return rc;
  if ((rc = dalloc(&h->attn_hi, 32 * d))) return rc;
  if ((rc = dalloc(&h->attn_lo, 32 * d))) return rc;
  if ((rc = dalloc(&h->h_hi, 32 * I))) return rc;
  if ((rc = dalloc(&h->h_lo, 32 * I))) return rc;
  h->m_wqkv = new CUtensorMap[c.num_layers]; h->m_wo = new CUtensorMap[c.num_layers];
  h->m_wgu = new CUtensorMap[c.num_layers]; h->m_wd = new CUtensorMap[c.num_layers];
  const int qk_rows = (c.num_heads + c.num_kv_heads) * c.head_dim;
  for (int l = 0; l < c.num_layers; ++l) {
    const float* Wl = h->W + L.layer0 + (int64_t)l * L.layer_stride;
    float* wq = h->tc_wqkv + (size_t)l * nqkv * d;
    float* wg = h->tc_wgu + (size_t)l * 2 * I * d;
    k_build_tc_weight<<<(unsigned)nqkv, 256>>>(Wl + L.wqkv, Wl + L.ln1, wq, (int)nqkv, (int)d, 1, qk_rows, c.head_dim, 0);
    k_build_tc_weight<<<(unsigned)(2 * I), 256>>>(Wl + L.wgate_up, Wl + L.ln2, wg, (int)(2 * I), (int)d, 2, 0, 0, (int)I);
    if ((rc = encode_map_2d(&h->m_wqkv[l], wq, nqkv, d, 128))) return rc;
    if ((rc = encode_map_2d(&h->m_wo[l], Wl + L.wo, d, (size_t)c.num_heads * c.head_dim, 128))) return rc;
    if ((rc = encode_map_2d(&h->m_wgu[l], wg, 2 * I, d, 128))) return rc;
    if ((rc = encode_map_2d(&h->m_wd[l], Wl + L.wdown, d, I, 128))) return rc;
  }
  const int nhc = c.num_vq * c.num_audio_tokens;
  k_build_tc_weight<<<nhc, 256>>>(h->W + L.head_code, h->W + L.final_norm, h->tc_heads_code, nhc, (int)d, 0, 0, 0, 0);
  k_build_tc_weight<<<c.num_text_tokens, 256>>>(h->W + L.head_text, h->W + L.final_norm, h->tc_heads_text,
                                                c.num_text_tokens, (int)d, 0, 0, 0, 0);
  CTB_CUDA(cudaDeviceSynchronize());
  if ((rc = encode_map_2d(&h->m_hcode, h->tc_heads_code, nhc, d, 128))) return rc;
  if ((rc = encode_map_2d(&h->m_htext, h->tc_heads_text, c.num_text_tokens, d, 128))) return rc;
  for (int n = 0; n < 2; ++n) {
    const uint32_t npad = n ? 32 : 16;
    if ((rc = encode_map_2d(&h->m_x[n][0], h->x_hi, 32, d, npad))) return rc;
    if ((rc = encode_map_2d(&h->m_x[n][1], h->x_lo, 32, d, npad))) return rc;
    if ((rc = encode_map_2d(&h->m_attn[n][0], h->attn_hi, 32, d, npad))) return rc;
    if ((rc = encode_map_2d(&h->m_attn[n][1], h->attn_lo, 32, d, npad))) return rc;
    if ((rc = encode_map_2d(&h->m_h[n][0], h->h_hi, 32, I, npad))) return rc;
    if ((rc = encode_map_2d(&h->m_h[n][1], h->h_lo, 32, I, npad))) return rc;
  }
#define TCATTR(E, C) if ((rc = set_tc_attr<E, 16, C>())) return rc; if ((rc = set_tc_attr<E, 32, C>())) return rc;
  TCATTR(DE_QKV, CS_QKV) TCATTR(DE_OPROJ, CS_O) TCATTR(DE_GATEUP, CS_GU) TCATTR(DE_DOWN, CS_DOWN) TCATTR(DE_HEADS, CS_HEADS)
#undef TCATTR
  return CTB_OK;
}

extern "C" int ctb_gpt_create(const ctb_gpt_config* c, const float* weights_dev, ctb_gpt** out) {
  if (!c || !weights_dev || !out) return set_err(CTB_ERR_ARG, "null argument");
  if (c->hidden_size != KC) return set_err(CTB_ERR_ARG, "hidden_size must be %d", KC);
  if (c->intermediate_size % KC) return set_err(CTB_ERR_ARG, "intermediate_size must be a multiple of %d", KC);
  if (c->head_dim != 64) return set_err(CTB_ERR_ARG, "head_dim must be 64");
  if (c->num_heads % c->num_kv_heads) return set_err(CTB_ERR_ARG, "num_heads %% num_kv_heads != 0");
  if (c->num_vq > 8 || c->num_vq < 1) return set_err(CTB_ERR_ARG, "num_vq out of range");
  if (c->max_batch < 1 || c->max_context < 1 || c->max_context > c->max_positions)
    return set_err(CTB_ERR_ARG, "bad max_batch/max_context");
  int dev_count = 0;
  CTB_CUDA(cudaGetDeviceCount(&dev_count));
  if (dev_count < 1) return set_err(CTB_ERR_CUDA, "no CUDA device: chattts_b200 has no CPU path");
  {
    int dev = 0, sms = 0;
    CTB_CUDA(cudaGetDevice(&dev));
    CTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_num_sms = sms;
  }
  ctb_gpt* h = new ctb_gpt();
  memset(h, 0, sizeof(*h));
  h->cfg = *c;
  ctb_gpt_layout_query(c, &h->lay);
  h->W = weights_dev;
  h->pages_per_row = (c->max_context + kPageTokens - 1) / kPageTokens;
  h->nsplit_max = (c->max_context + ATT_SPLIT_UNIT - 1) / ATT_SPLIT_UNIT;
  const int bt = bt_for(c->max_batch);
  h->bpad_max = ((c->max_batch + bt - 1) / bt) * bt;
  const size_t Bp = h->bpad_max, d = c->hidden_size;
  const size_t nq = (size_t)c->num_heads * c->head_dim;
  int rc;
#define TRY(e) if ((rc = (e)) != CTB_OK) { ctb_gpt_destroy(h); return rc; }
  TRY(dalloc(&h->x, Bp * d));
  TRY(dalloc(&h->qbuf, Bp * nq));
  TRY(dalloc(&h->attn, Bp * nq));
  TRY(dalloc(&h->mlp, Bp * c->intermediate_size));
  const size_t nlog = std::max((size_t)c->num_vq * c->num_audio_tokens, (size_t)c->num_text_tokens);
  TRY(dalloc(&h->logits, Bp * nlog));
  // The KV pool is sized by what a generate() call can actually touch (ctb_gpt_begin -> kv_reserve), not by
  // max_batch x max_context: a handle that allows 32 rows x 4096 tokens costs nothing until such a call arrives.
  h->kv = nullptr; h->kv_layer_floats = 0; h->kv_pages = 0;
  TRY(dalloc(&h->part, (size_t)c->max_batch * c->num_heads * h->nsplit_max * (c->head_dim + 2)));
  TRY(dalloc(&h->block_table, (size_t)c->max_batch * h->pages_per_row));
  TRY(dalloc(&h->seq_len, Bp));
  TRY(dalloc(&h->pos, Bp));
  TRY(dalloc(&h->counter, (size_t)c->max_batch * c->num_heads));
  TRY(dalloc(&h->end_idx, Bp));
  TRY(dalloc(&h->idx, Bp * c->num_vq));
  TRY(dalloc(&h->active, Bp));
  TRY(dalloc(&h->finish, Bp));
  TRY(dalloc(&h->st, 1));
  TRY(dalloc(&h->bar, 4));
  if (getenv("CTB_MEGA_TRACE")) TRY(dalloc(&h->trace, 4096));
  h->flow_ok = getenv("CTB_NO_FLOW") == nullptr && g_num_sms >= 128 && g_num_sms <= 191 && c->intermediate_size == 4 * KC &&
               c->num_heads == c->num_kv_heads && c->num_heads * c->head_dim == KC && c->num_heads <= FL_HEADS;
  if (h->flow_ok) {
    TRY(dalloc(&h->flow_arena, FL_ARENA_WORDS));
    TRY(dalloc(&h->flow_epoch, 4));
    const unsigned e0 = FL_EPOCH_STEP;
    cudaMemcpy(h->flow_epoch, &e0, sizeof(e0), cudaMemcpyHostToDevice);
    // measured on B200 (tools/flow_check.py): one copy of the exchange words is fastest (replicas multiply the 8-byte
    // stores; the read hot-spot they were meant to relieve is the smaller effect), and the kernel beats k_step at every batch it is built for (B <= 4)
    h->flow_R = getenv("CTB_FLOW_R") ? std::max(1, std::min(FL_RMAX, atoi(getenv("CTB_FLOW_R")))) : 1;
    h->flow_l2_ahead = getenv("CTB_FLOW_L2_AHEAD") ? atoi(getenv("CTB_FLOW_L2_AHEAD")) : 0;
    h->flow_max_batch = getenv("CTB_FLOW_MAX_BATCH") ? std::max(0, std::min(FL_BMAX, atoi(getenv("CTB_FLOW_MAX_BATCH")))) : FL_BMAX;
  }
#undef TRY
  h->use_graph = getenv("CTB_NO_GRAPH") == nullptr;
  h->pf_enabled = getenv("CTB_NO_BATCHED_PREFILL") == nullptr;
  h->mega_ok = getenv("CTB_NO_MEGA") == nullptr && g_num_sms >= 128 && c->intermediate_size == 4 * KC &&
               (c->hidden_size / 2 + g_num_sms - 1) / g_num_sms <= MG_DOWN_PAIRS;
  h->mega_max_batch = getenv("CTB_MEGA_MAX_BATCH") ? std::min(8, atoi(getenv("CTB_MEGA_MAX_BATCH"))) : 4;
  // tensor-core decode GEMMs (tc_decode.cuh): CTB_GPT_TC=1 forces them for every batch, CTB_GPT_FMA=1 disables
  // them; by default they serve batches > 16 rows, where the fp32 FMA path turns compute-bound.
  h->tc_ready = getenv("CTB_GPT_FMA") == nullptr && c->max_batch <= 32 &&
                (getenv("CTB_GPT_TC") != nullptr || c->max_batch > 16);
  h->tc_min_batch = getenv("CTB_GPT_TC") != nullptr ? 1 : 17;
  if (h->tc_ready && (rc = tc_setup(h)) != CTB_OK) { ctb_gpt_destroy(h); return rc; }
  *out = h;
  return CTB_OK;
}

extern "C" int ctb_gpt_destroy(ctb_gpt* h) {
  if (!h) return CTB_OK;
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  void* ptrs[] = {h->x, h->qbuf, h->attn, h->mlp, h->logits, h->kv, h->part, h->block_table, h->seq_len,
                  h->pos, h->counter, h->end_idx, h->idx, h->active, h->finish, h->st, h->tc_wqkv, h->tc_wgu,
                  h->tc_heads_code, h->tc_heads_text, h->x_hi, h->x_lo, h->attn_hi, h->attn_lo, h->h_hi, h->h_lo, h->bar, h->trace, h->flow_arena, h->flow_epoch, h->gw_hi, h->gw_lo, h->pf_resid, h->pf_xn, h->pf_qkv, h->pf_q,
                  h->pf_attn, h->pf_gu, h->pf_h, h->pf_ones, h->pf_zeros, h->pf_npre, h->pf_nvalid};
  delete[] h->m_wqkv; delete[] h->m_wo; delete[] h->m_wgu; delete[] h->m_wd;
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return CTB_OK;
}

// ------------------------------------------------------------------ launches
template <int BT, int EPI>
static int launch_gemv_t(const GemvP& p, int ntiles, cudaStream_t s) {
  const size_t smem = (size_t)BT * KC * sizeof(float);
  { int rc = ensure_smem_attr((const void*)k_gemv<BT, EPI>, (int)smem); if (rc) return rc; }
  // persistent: one CTA per SM strides over the warp tasks (DOWN: clusters of DOWN_SPLIT CTAs)
  int ctas = std::min(g_num_sms, (p.ntasks + GEMV_WARPS - 1) / GEMV_WARPS);
  unsigned cluster = 1;
  if (EPI == EPI_DOWN) {
    cluster = DOWN_SPLIT;
    // clusters of 4 can only occupy ~132 of the 148 SMs at once (GPC granularity): one wave of 33 clusters, not 37
    const int max_groups = (g_num_sms * 132 / 148) / DOWN_SPLIT;
    const int groups = std::max(1, std::min(max_groups, (p.ntasks + GEMV_WARPS - 1) / GEMV_WARPS));
    if ((p.ntasks + groups * GEMV_WARPS - 1) / (groups * GEMV_WARPS) > DOWN_MAX_TASKS)
      return set_err(CTB_ERR_STATE, "DOWN kernel: too few SMs (%d) for %d tasks", g_num_sms, p.ntasks);
    ctas = groups * DOWN_SPLIT;
  }
  dim3 grid(ctas, ntiles);
  CTB_CUDA(launch_pdl_cluster(k_gemv<BT, EPI>, grid, dim3(GEMV_WARPS * 32), smem, s, cluster, p));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

template <int EPI>
static int launch_gemv(int bt, const GemvP& p, int ntiles, cudaStream_t s) {
  switch (bt) {
    case 1: return launch_gemv_t<1, EPI>(p, ntiles, s);
    case 2: return launch_gemv_t<2, EPI>(p, ntiles, s);
    case 4: return launch_gemv_t<4, EPI>(p, ntiles, s);
    case 8: return launch_gemv_t<8, EPI>(p, ntiles, s);
    case 16: return launch_gemv_t<16, EPI>(p, ntiles, s);
    default: return launch_gemv_t<32, EPI>(p, ntiles, s);
  }
}

template <int BT>
static int launch_gateup_small_t(const GemvP& p, cudaStream_t s) {
  const size_t smem = ((size_t)BT * KC + GU_ZONE_FLOATS) * sizeof(float);
  { int rc = ensure_smem_attr((const void*)k_gateup_small<BT>, (int)smem); if (rc) return rc; }
  CTB_CUDA(launch_pdl(k_gateup_small<BT>, dim3(g_num_sms), dim3(GEMV_WARPS * 32), smem, s, p));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}
static int launch_gateup_small(int bt, const GemvP& p, cudaStream_t s) {
  switch (bt) {
    case 1: return launch_gateup_small_t<1>(p, s);
    case 2: return launch_gateup_small_t<2>(p, s);
    case 4: return launch_gateup_small_t<4>(p, s);
    case 8: return launch_gateup_small_t<8>(p, s);
    default: return launch_gateup_small_t<16>(p, s);
  }
}

template <int BT>
static int launch_down_small_t(const GemvP& p, cudaStream_t s) {
  const size_t smem = (size_t)BT * p.K * sizeof(float);
  { int rc = ensure_smem_attr((const void*)k_down_small<BT>, (int)smem); if (rc) return rc; }
  CTB_CUDA(launch_pdl(k_down_small<BT>, dim3(g_num_sms), dim3(GEMV_WARPS * 32), smem, s, p));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}
static int launch_down_small(int bt, const GemvP& p, cudaStream_t s) {
  switch (bt) {
    case 1: return launch_down_small_t<1>(p, s);
    case 2: return launch_down_small_t<2>(p, s);
    case 4: return launch_down_small_t<4>(p, s);
    case 8: return launch_down_small_t<8>(p, s);
    default: return launch_down_small_t<16>(p, s);
  }
}

static int launch_sample(const SampleP& sp, cudaStream_t s) {
  const size_t smem = (size_t)sp.V * sizeof(float) + 2 * 1024 * sizeof(uint32_t);
  { int rc = ensure_smem_attr((const void*)k_sample, (int)smem); if (rc) return rc; }
  CTB_CUDA(launch_pdl(k_sample, dim3(sp.rows), dim3(SAMPLE_THREADS), smem, s, sp));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

struct StepCtx {
  GemvP g;
  AttnP a;
  int bt, ntiles, decode;
};

static StepCtx make_ctx(ctb_gpt* h, int decode) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  StepCtx x{};
  x.decode = decode;
  x.bt = bt_for(h->B);
  x.ntiles = (h->B + x.bt - 1) / x.bt;
  GemvP& g = x.g;
  g.B = h->B; g.st = h->st; g.check_finished = decode; g.eps = c.rms_eps;
  g.block_table = h->block_table; g.pages_per_row = h->pages_per_row; g.pos = h->pos; g.active = h->active;
  g.rope_cos = h->W + L.rope_cos; g.rope_sin = h->W + L.rope_sin;
  g.Hq = c.num_heads; g.Hkv = c.num_kv_heads; g.hd = c.head_dim; g.I = c.intermediate_size; g.xres = h->x;
  AttnP& a = x.a;
  a.st = h->st; a.check_finished = decode; a.q = h->qbuf; a.block_table = h->block_table;
  a.pages_per_row = h->pages_per_row; a.pos = h->pos; a.active = h->active; a.out = h->attn; a.part = h->part;
  a.out_hi = h->use_tc ? h->attn_hi : nullptr; a.out_lo = h->use_tc ? h->attn_lo : nullptr;
  a.counter = h->counter; a.Hq = c.num_heads; a.Hkv = c.num_kv_heads; a.hd = c.head_dim;
  a.nsplit_max = h->nsplit_max; a.scaling = 1.0f / sqrtf((float)c.head_dim);
  return x;
}

// kind: 0 qkv(+rope+kv append), 1 attention, 2 o-proj(+residual), 3 gate/up(+silu*mul), 4 down(+residual)
static int launch_layer_kernel(ctb_gpt* h, const StepCtx& x, int l, int kind, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  const int d = c.hidden_size, I = c.intermediate_size, hd = c.head_dim;
  const float* Wl = h->W + L.layer0 + (int64_t)l * L.layer_stride;
  float* kvl = h->kv + (size_t)l * h->kv_layer_floats;
  GemvP p = x.g;
  switch (kind) {
    case 0:
      p.W = Wl + L.wqkv; p.K = d; p.nrows = (c.num_heads + 2 * c.num_kv_heads) * hd; p.ntasks = p.nrows / 2;
      p.xin = h->x; p.normw = Wl + L.ln1; p.out = h->qbuf; p.kv = kvl;
      return launch_gemv<EPI_QKV>(x.bt, p, x.ntiles, s);
    case 1: {
      AttnP a = x.a;
      a.kv = kvl;
      // context after this call <= T0 + max_new: only launch splits that can be populated
      const int max_ctx = std::min(c.max_context, h->T0 + h->max_new);
      // enough CTAs to fill the chip twice; a CTA walks chunks c, c + grid.x, ... with a running softmax,
      // so large batches need no cross-CTA merge at all
      const int want = (2 * g_num_sms + c.num_heads * h->B - 1) / (c.num_heads * h->B);
      dim3 agrid(std::max(1, std::min(want, (max_ctx + ATT_CHUNK - 1) / ATT_CHUNK)), c.num_heads, h->B);
      CTB_CUDA(launch_pdl(k_attn, agrid, dim3(ATT_THREADS), 0, s, a));
      CTB_LAUNCH_CHECK();
      return CTB_OK;
    }
    case 2:
      p.W = Wl + L.wo; p.K = c.num_heads * hd; p.nrows = d; p.ntasks = d / 2; p.xin = h->attn; p.normw = nullptr;
      return launch_gemv<EPI_OPROJ>(x.bt, p, x.ntiles, s);
    case 3:
      p.W = Wl + L.wgate_up; p.K = d; p.nrows = 2 * I; p.ntasks = I; p.xin = h->x; p.normw = Wl + L.ln2;
      p.out = h->mlp;
      if (x.bt <= 8 && x.ntiles == 1 && (I + g_num_sms * GEMV_WARPS - 1) / (g_num_sms * GEMV_WARPS) <= GU_TASKS &&
          getenv("CTB_GATEUP_GENERIC") == nullptr)
        return launch_gateup_small(x.bt, p, s);
      return launch_gemv<EPI_GATEUP>(x.bt, p, x.ntiles, s);
    case 4:
      p.W = Wl + L.wdown; p.K = I; p.nrows = d; p.ntasks = d / 2; p.xin = h->mlp; p.normw = nullptr;
      if (x.bt <= 16 && x.ntiles == 1 && I == 4 * KC && (d / 2 + g_num_sms - 1) / g_num_sms <= DS_PAIRS &&
          getenv("CTB_DOWN_CLUSTER") == nullptr)
        return launch_down_small(x.bt, p, s);
      return launch_gemv<EPI_DOWN>(x.bt, p, x.ntiles, s);
  }
  return set_err(CTB_ERR_ARG, "bad kernel kind %d", kind);
}

static int launch_heads(ctb_gpt* h, const StepCtx& x, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const int rpi = h->infer_text ? 1 : c.num_vq;
  const int V = h->infer_text ? c.num_text_tokens : c.num_audio_tokens;
  GemvP hp = x.g;
  hp.W = h->W + (h->infer_text ? h->lay.head_text : h->lay.head_code);
  hp.K = c.hidden_size; hp.nrows = rpi * V; hp.ntasks = (hp.nrows + 1) / 2; hp.xin = h->x;
  hp.normw = h->W + h->lay.final_norm; hp.out = h->logits; hp.rows_per_item = rpi; hp.V = V;
  hp.hidden_out = h->hiddens_out; hp.hidden_stride = h->max_new * c.hidden_size;
  return launch_gemv<EPI_HEADS>(x.bt, hp, x.ntiles, s);
}

static int launch_sampler(ctb_gpt* h, const StepCtx& x, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const int rpi = h->infer_text ? 1 : c.num_vq;
  SampleP sp{};
  sp.st = h->st; sp.check_finished = x.decode; sp.logits = h->logits; sp.rows = h->B * rpi;
  sp.V = h->infer_text ? c.num_text_tokens : c.num_audio_tokens;
  sp.rows_per_item = rpi; sp.cfg = h->sampler; sp.q_noise = h->q_noise; sp.gen_ids = h->ids_out;
  sp.gen_stride = h->max_new; sp.gen_inner = c.num_vq; sp.out_idx = h->idx;
  return launch_sample(sp, s);
}

template <int EPI, int CS>
static int launch_tc(int npad, const CUtensorMap& mw, const CUtensorMap& mxh, const CUtensorMap& mxl, const TcDecP& p,
                     cudaStream_t s) {
  const int tiles = (p.nrows + 127) / 128;
  dim3 grid(tiles * CS);
  if (npad == 16)
    CTB_CUDA(launch_pdl_cluster(k_tc_dec<EPI, 16, CS>, grid, dim3(TD_THREADS), (size_t)TdCfg<16>::SMEM_BYTES, s, (unsigned)CS, mw, mxh, mxl, p));
  else
    CTB_CUDA(launch_pdl_cluster(k_tc_dec<EPI, 32, CS>, grid, dim3(TD_THREADS), (size_t)TdCfg<32>::SMEM_BYTES, s, (unsigned)CS, mw, mxh, mxl, p));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

static TcDecP make_tc(ctb_gpt* h) {
  const ctb_gpt_config& c = h->cfg;
  TcDecP p{};
  p.B = h->B; p.eps = c.rms_eps; p.d = c.hidden_size; p.xres = h->x; p.x_hi = h->x_hi; p.x_lo = h->x_lo;
  p.qbuf = h->qbuf; p.block_table = h->block_table; p.pages_per_row = h->pages_per_row; p.pos = h->pos;
  p.active = h->active; p.rope_cos = h->W + h->lay.rope_cos; p.rope_sin = h->W + h->lay.rope_sin;
  p.Hq = c.num_heads; p.Hkv = c.num_kv_heads; p.hd = c.head_dim; p.h_hi = h->h_hi; p.h_lo = h->h_lo;
  p.I = c.intermediate_size; p.st = h->st;
  return p;
}

static int launch_layer_kernel_tc(ctb_gpt* h, int l, int kind, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const int n = h->B > 16 ? 1 : 0, npad = n ? 32 : 16;
  const int d = c.hidden_size, I = c.intermediate_size;
  TcDecP p = make_tc(h);
  switch (kind) {
    case 0:
      p.K = d; p.kslice = d / CS_QKV; p.nrows = (c.num_heads + 2 * c.num_kv_heads) * c.head_dim; p.xraw = h->x;
      p.kv = h->kv + (size_t)l * h->kv_layer_floats;
      return launch_tc<DE_QKV, CS_QKV>(npad, h->m_wqkv[l], h->m_x[n][0], h->m_x[n][1], p, s);
    case 2:
      p.K = c.num_heads * c.head_dim; p.kslice = p.K / CS_O; p.nrows = d; p.xraw = nullptr;
      return launch_tc<DE_OPROJ, CS_O>(npad, h->m_wo[l], h->m_attn[n][0], h->m_attn[n][1], p, s);
    case 3:
      p.K = d; p.kslice = d / CS_GU; p.nrows = 2 * I; p.xraw = h->x;
      return launch_tc<DE_GATEUP, CS_GU>(npad, h->m_wgu[l], h->m_x[n][0], h->m_x[n][1], p, s);
    case 4:
      p.K = I; p.kslice = I / CS_DOWN; p.nrows = d; p.xraw = nullptr;
      return launch_tc<DE_DOWN, CS_DOWN>(npad, h->m_wd[l], h->m_h[n][0], h->m_h[n][1], p, s);
  }
  return set_err(CTB_ERR_ARG, "bad tc kernel kind %d", kind);
}

static int launch_heads_tc(ctb_gpt* h, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const int n = h->B > 16 ? 1 : 0, npad = n ? 32 : 16;
  TcDecP p = make_tc(h);
  const int rpi = h->infer_text ? 1 : c.num_vq;
  const int V = h->infer_text ? c.num_text_tokens : c.num_audio_tokens;
  p.K = c.hidden_size; p.kslice = p.K / CS_HEADS; p.nrows = rpi * V; p.xraw = h->x;
  p.logits = h->logits; p.rows_per_item = rpi; p.V = V; p.hidden_out = h->hiddens_out;
  p.hidden_stride = h->max_new * c.hidden_size; p.final_norm_w = h->W + h->lay.final_norm;
  return launch_tc<DE_HEADS, CS_HEADS>(npad, h->infer_text ? h->m_htext : h->m_hcode, h->m_x[n][0], h->m_x[n][1], p, s);
}

template <int BT>
static int launch_step_mega_t(const MegaP& mp, cudaStream_t s) {
  // [BT][768] activations + the 144 KiB landing zone (gate/up weights; reused as merge scratch and for the down
  // phase's [BT][3072] activations)
  const size_t smem = (size_t)BT * KC * sizeof(float) + (size_t)MG_GW_FLOATS * sizeof(float);
  { int rc = ensure_smem_attr((const void*)k_step<BT>, (int)smem); if (rc) return rc; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g_num_sms); cfg.blockDim = dim3(MG_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // every CTA must be resident: the phases meet at grid barriers
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CTB_CUDA(cudaLaunchKernelEx(&cfg, k_step<BT>, mp));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

static int launch_step_mega(ctb_gpt* h, int col, bool sample, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  MegaP m{};
  m.W = h->W; m.layer0 = L.layer0; m.layer_stride = L.layer_stride; m.o_wqkv = L.wqkv; m.o_wo = L.wo; m.o_wgu = L.wgate_up;
  m.o_wd = L.wdown; m.o_ln1 = L.ln1; m.o_ln2 = L.ln2; m.o_final_norm = L.final_norm;
  m.o_head = h->infer_text ? L.head_text : L.head_code; m.o_emb_code = L.emb_code; m.o_emb_text = L.emb_text;
  m.o_cos = L.rope_cos; m.o_sin = L.rope_sin;
  m.L = c.num_layers; m.d = c.hidden_size; m.I = c.intermediate_size; m.Hq = c.num_heads; m.Hkv = c.num_kv_heads;
  m.hd = c.head_dim; m.eps = c.rms_eps; m.scaling = 1.0f / sqrtf((float)c.head_dim);
  m.x = h->x; m.qbuf = h->qbuf; m.attn = h->attn; m.mlp = h->mlp; m.logits = h->logits; m.kv = h->kv; m.part = h->part;
  m.kv_layer_floats = h->kv_layer_floats; m.block_table = h->block_table; m.pages_per_row = h->pages_per_row;
  m.seq_len = h->seq_len; m.counter = h->counter; m.nsplit_max = h->nsplit_max; m.st = h->st; m.bar = h->bar;
  m.decode = col < 0; m.col = col < 0 ? 0 : col; m.T0 = h->T0; m.sample = sample ? 1 : 0;
  m.emb = h->emb; m.mask = h->mask; m.ids_out = h->ids_out; m.max_new = h->max_new; m.num_vq = c.num_vq;
  m.num_audio = c.num_audio_tokens; m.infer_text = h->infer_text; m.B = h->B;
  m.hidden_out = h->hiddens_out; m.hidden_stride = h->max_new * c.hidden_size;
  m.rows_per_item = h->infer_text ? 1 : c.num_vq; m.V = h->infer_text ? c.num_text_tokens : c.num_audio_tokens;
  m.trace = h->trace;
  CTB_CUDA(cudaMemsetAsync(h->bar, 0, sizeof(unsigned), s));
  switch (bt_for(h->B)) {
    case 1: return launch_step_mega_t<1>(m, s);
    case 2: return launch_step_mega_t<2>(m, s);
    case 4: return launch_step_mega_t<4>(m, s);
    default: return launch_step_mega_t<8>(m, s);
  }
}

template <int BT>
static int launch_step_flow_t(const FlowP& fp, cudaStream_t s) {
  const size_t smem = (size_t)FL_RING_BYTES + (size_t)BT * KC * sizeof(float);
  { int rc = ensure_smem_attr((const void*)k_flow<BT>, (int)smem); if (rc) return rc; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g_num_sms); cfg.blockDim = dim3(FL_LAUNCH_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // every CTA must be resident: CTAs wait for each other's words
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CTB_CUDA(cudaLaunchKernelEx(&cfg, k_flow<BT>, fp));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

static bool flow_ink(const ctb_gpt* h);

// nsteps > 0: multi-step decode with the sampling tail inside the kernel (requires flow_ink(h)); 0: one step, logits out
static int launch_step_flow(ctb_gpt* h, int col, bool sample, cudaStream_t s, int nsteps = 0) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  FlowP m{};
  m.W = h->W; m.layer0 = L.layer0; m.layer_stride = L.layer_stride; m.o_wqkv = L.wqkv; m.o_wo = L.wo; m.o_wgu = L.wgate_up;
  m.o_wd = L.wdown; m.o_ln1 = L.ln1; m.o_ln2 = L.ln2; m.o_final_norm = L.final_norm;
  m.o_head = h->infer_text ? L.head_text : L.head_code; m.o_emb_code = L.emb_code; m.o_emb_text = L.emb_text;
  m.o_cos = L.rope_cos; m.o_sin = L.rope_sin;
  m.L = c.num_layers; m.I = c.intermediate_size; m.Hq = c.num_heads; m.hd = c.head_dim; m.eps = c.rms_eps;
  m.scaling = 1.0f / sqrtf((float)c.head_dim);
  m.logits = h->logits; m.kv = h->kv; m.kv_layer_floats = h->kv_layer_floats; m.block_table = h->block_table;
  m.pages_per_row = h->pages_per_row; m.seq_len = h->seq_len; m.st = h->st;
  m.decode = col < 0; m.col = col < 0 ? 0 : col; m.T0 = h->T0; m.sample = sample ? 1 : 0;
  m.emb = h->emb; m.mask = h->mask; m.ids_out = h->ids_out; m.max_new = h->max_new; m.num_vq = c.num_vq;
  m.num_audio = c.num_audio_tokens; m.infer_text = h->infer_text; m.B = h->B;
  m.hidden_out = h->hiddens_out; m.hidden_stride = h->max_new * c.hidden_size;
  m.rows_per_item = h->infer_text ? 1 : c.num_vq; m.V = h->infer_text ? c.num_text_tokens : c.num_audio_tokens;
  m.arena = h->flow_arena; m.epoch = h->flow_epoch; m.R = h->flow_R; m.trace = h->trace;
  m.l2_ahead = h->flow_l2_ahead;
  m.ink = nsteps > 0 ? 1 : 0; m.nsteps = nsteps > 0 ? nsteps : 1; m.samp = h->sampler; m.q_noise = h->q_noise;
  m.finish = h->finish; m.end_idx = h->end_idx; m.ids_w = h->ids_out;
  switch (bt_for(h->B)) {
    case 1: return launch_step_flow_t<1>(m, s);
    case 2: return launch_step_flow_t<2>(m, s);
    default: return launch_step_flow_t<4>(m, s);
  }
}

static bool use_flow(const ctb_gpt* h) { return h->flow_ok && !h->use_tc && h->B <= h->flow_max_batch; }
// decode steps of audio generation at B <= 2 sample inside k_flow and run many steps per launch
static bool flow_ink(const ctb_gpt* h) {
  static const bool off = getenv("CTB_FLOW_NO_INK") != nullptr;
  return !off && use_flow(h) && !h->infer_text && h->B <= 2 && h->cfg.num_audio_tokens <= FL_VPAD &&
         h->B * h->cfg.num_vq <= FL_SROWS && h->cfg.num_vq <= 8;
}

// One loop iteration.  col >= 0: prefill column `col` of the prompt; col < 0: decode step.
// sample: run heads + sampler + finalize (last prompt column and every decode step).
static int enqueue_step(ctb_gpt* h, int col, bool sample, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  const int decode = col < 0;
  int rc;

  if (use_flow(h) || (h->mega_ok && !h->use_tc && h->B <= h->mega_max_batch)) {
    // small batches: the whole step (input -> 20 layers -> heads) is one persistent cooperative kernel
    if ((rc = use_flow(h) ? launch_step_flow(h, col, sample, s) : launch_step_mega(h, col, sample, s))) return rc;
    if (!sample) return CTB_OK;
    const StepCtx xm = make_ctx(h, decode);
    if ((rc = launch_sampler(h, xm, s))) return rc;
    FinalP fm{};
    fm.st = h->st; fm.B = h->B; fm.rows_per_item = h->infer_text ? 1 : c.num_vq; fm.num_vq = c.num_vq;
    fm.max_new = h->max_new; fm.eos = h->sampler.eos_token; fm.idx = h->idx; fm.ids_out = h->ids_out;
    fm.finish = h->finish; fm.end_idx = h->end_idx;
    CTB_CUDA(launch_pdl(k_finalize, dim3(1), dim3(256), 0, s, fm));
    CTB_LAUNCH_CHECK();
    return CTB_OK;
  }

  InputP ip{};
  ip.st = h->st; ip.decode = decode; ip.B = h->B; ip.d = c.hidden_size; ip.col = decode ? 0 : col; ip.T0 = h->T0;
  ip.emb = h->emb; ip.mask = h->mask;
  ip.emb_code = h->W + L.emb_code; ip.emb_text = h->W + L.emb_text;
  ip.ids_out = h->ids_out; ip.max_new = h->max_new; ip.num_vq = c.num_vq; ip.num_audio = c.num_audio_tokens;
  ip.infer_text = h->infer_text;
  ip.x = h->x; ip.seq_len = h->seq_len; ip.pos = h->pos; ip.active = h->active;
  ip.x_hi = h->use_tc ? h->x_hi : nullptr; ip.x_lo = h->use_tc ? h->x_lo : nullptr;
  CTB_CUDA(launch_pdl(k_input, dim3(h->B), dim3(256), 0, s, ip));
  CTB_LAUNCH_CHECK();

  const StepCtx x = make_ctx(h, decode);
  for (int l = 0; l < c.num_layers; ++l)
    for (int kind = 0; kind < 5; ++kind)
      if ((rc = (h->use_tc && kind != 1) ? launch_layer_kernel_tc(h, l, kind, s) : launch_layer_kernel(h, x, l, kind, s)))
        return rc;
  if (!sample) return CTB_OK;
  if ((rc = h->use_tc ? launch_heads_tc(h, s) : launch_heads(h, x, s))) return rc;
  if ((rc = launch_sampler(h, x, s))) return rc;

  FinalP fp{};
  fp.st = h->st; fp.B = h->B; fp.rows_per_item = h->infer_text ? 1 : c.num_vq; fp.num_vq = c.num_vq;
  fp.max_new = h->max_new; fp.eos = h->sampler.eos_token; fp.idx = h->idx; fp.ids_out = h->ids_out;
  fp.finish = h->finish; fp.end_idx = h->end_idx;
  CTB_CUDA(launch_pdl(k_finalize, dim3(1), dim3(256), 0, s, fp));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

// Measurement hook (bench.py roofline): launch ONE kernel kind for every layer (20 launches over
// 20 different weight slabs, so nothing is L2-resident between launches) on the state left by the
// last generate call.  kind 0..4 as above, 5 = heads, 6 = sampler.  Results of a later decode are
// undefined after this call (the residual stream is overwritten); call ctb_gpt_begin again.
extern "C" int ctb_gpt_profile_kernel(ctb_gpt* h, int32_t kind, void* stream) {
  if (!h || !h->started) return set_err(CTB_ERR_STATE, "ctb_gpt_begin has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  StepCtx x = make_ctx(h, 0);
  x.g.check_finished = 0; x.a.check_finished = 0;
  int rc;
  if (kind == 5) return h->use_tc ? launch_heads_tc(h, s) : launch_heads(h, x, s);
  if (kind == 6) return launch_sampler(h, x, s);
  if (kind == 8) {  // 16 decode iterations in ONE launch of the dataflow step kernel (sampling tail inside)
    if (!flow_ink(h)) return set_err(CTB_ERR_STATE, "multi-step dataflow kernel unavailable for this handle/batch");
    if (h->steps_enqueued + 16 > h->max_new || h->T0 + h->steps_enqueued + 16 > h->cfg.max_context)
      return set_err(CTB_ERR_STATE, "no room for 16 more steps (max_new / max_context reached)");
    h->steps_enqueued += 16;
    return launch_step_flow(h, -1, true, s, 16);
  }
  if (kind == 7) {  // the one-kernel decode step alone (context grows by one token per call)
    if (h->steps_enqueued >= h->max_new || h->T0 + h->steps_enqueued >= h->cfg.max_context)
      return set_err(CTB_ERR_STATE, "no room for another step (max_new / max_context reached)");
    h->steps_enqueued++;
    if (use_flow(h)) return launch_step_flow(h, -1, true, s, flow_ink(h) ? 1 : 0);
    if (!(h->mega_ok && h->B <= 8)) return set_err(CTB_ERR_STATE, "one-kernel step unavailable for this handle/batch");
    return launch_step_mega(h, -1, true, s);
  }
  for (int l = 0; l < h->cfg.num_layers; ++l)
    if ((rc = (h->use_tc && kind != 1) ? launch_layer_kernel_tc(h, l, kind, s) : launch_layer_kernel(h, x, l, kind, s)))
      return rc;
  return CTB_OK;
}

// ------------------------------------------------------------------ batched prefill
static int prefill_reserve(ctb_gpt* h, size_t rows) {
  const ctb_gpt_config& c = h->cfg;
  const size_t d = c.hidden_size, I = c.intermediate_size, nqkv = (size_t)(c.num_heads + 2 * c.num_kv_heads) * c.head_dim;
  int rc;
  if (!h->gw_hi) {
    const size_t n = (size_t)h->lay.layer_stride * c.num_layers;
    if ((rc = dalloc(&h->gw_hi, n))) return rc;
    if ((rc = dalloc(&h->gw_lo, n))) return rc;
    k_split_tf32_t<0><<<2048, 256>>>(h->W + h->lay.layer0, h->gw_hi, h->gw_lo, (int64_t)n);
    if ((rc = dalloc(&h->pf_ones, 2 * I))) return rc;
    if ((rc = dalloc(&h->pf_zeros, 2 * I))) return rc;
    std::vector<float> ones(2 * I, 1.0f);
    CTB_CUDA(cudaMemcpy(h->pf_ones, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice));
    CTB_CUDA(cudaDeviceSynchronize());
  }
  if (rows > h->pf_rows) {
    float** bufs[] = {&h->pf_resid, &h->pf_xn, &h->pf_qkv, &h->pf_q, &h->pf_attn, &h->pf_gu, &h->pf_h};
    for (float** b : bufs) if (*b) { cudaFree(*b); *b = nullptr; }
    if (h->pf_npre) { cudaFree(h->pf_npre); h->pf_npre = nullptr; }
    if (h->pf_nvalid) { cudaFree(h->pf_nvalid); h->pf_nvalid = nullptr; }
    if ((rc = dalloc(&h->pf_resid, rows * d))) return rc;
    if ((rc = dalloc(&h->pf_xn, rows * d))) return rc;
    if ((rc = dalloc(&h->pf_qkv, rows * nqkv))) return rc;
    if ((rc = dalloc(&h->pf_q, rows * d))) return rc;
    if ((rc = dalloc(&h->pf_attn, rows * d))) return rc;
    if ((rc = dalloc(&h->pf_gu, rows * 2 * I))) return rc;
    if ((rc = dalloc(&h->pf_h, rows * I))) return rc;
    if ((rc = dalloc(&h->pf_npre, rows))) return rc;
    if ((rc = dalloc(&h->pf_nvalid, (size_t)h->cfg.max_batch))) return rc;
    h->pf_rows = rows;
  }
  return CTB_OK;
}

static int prefill_batched(ctb_gpt* h, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const ctb_gpt_layout& L = h->lay;
  const int B = h->B, T0 = h->T0, M = B * T0;
  const int d = c.hidden_size, I = c.intermediate_size, nqkv = (c.num_heads + 2 * c.num_kv_heads) * c.head_dim;
  int rc;
  if ((rc = prefill_reserve(h, (size_t)M))) return rc;
  CTB_CUDA(cudaMemcpyAsync(h->pf_resid, h->emb, (size_t)M * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
  k_prefill_positions<<<B, 32, 0, s>>>(h->mask, h->pf_npre, h->pf_nvalid, T0);
  CTB_LAUNCH_CHECK();
  PrefillP pp{};
  pp.B = B; pp.T0 = T0; pp.Hq = c.num_heads; pp.Hkv = c.num_kv_heads; pp.hd = c.head_dim; pp.d = d; pp.mask = h->mask;
  pp.npre = h->pf_npre; pp.nvalid = h->pf_nvalid; pp.qkv = h->pf_qkv; pp.q = h->pf_q; pp.block_table = h->block_table;
  pp.pages_per_row = h->pages_per_row; pp.rope_cos = h->W + L.rope_cos; pp.rope_sin = h->W + L.rope_sin;
  pp.permute_qk = h->use_tc ? 1 : 0; pp.attn = h->pf_attn; pp.scaling = 1.0f / sqrtf((float)c.head_dim);
  const int rms_blocks = (M + 7) / 8;
  for (int l = 0; l < c.num_layers; ++l) {
    const int64_t lo = (int64_t)l * L.layer_stride;
    const float* Wl = h->W + L.layer0 + lo;
    const float *Whi = h->gw_hi + lo, *Wlo = h->gw_lo + lo;
    pp.kv = h->kv + (size_t)l * h->kv_layer_floats;
    k_rms_rows<<<rms_blocks, 256, 0, s>>>(h->pf_resid, Wl + L.ln1, h->pf_xn, M, d, c.rms_eps);
    CTB_LAUNCH_CHECK();
    if ((rc = tc_gemm_launch<GE_NONE>(s, h->pf_xn, d, B, T0, nqkv, d, 1, d, 1, 0, Whi + L.wqkv, Wlo + L.wqkv, nullptr,
                                      nullptr, nullptr, 0, h->pf_qkv, nqkv))) return rc;
    k_prefill_rope_kv<<<dim3(T0, B), 256, 0, s>>>(pp);
    CTB_LAUNCH_CHECK();
    k_prefill_attn<<<dim3((T0 + PF_ATT_WARPS - 1) / PF_ATT_WARPS, c.num_heads, B), PF_ATT_WARPS * 32,
                     (size_t)PF_ATT_WARPS * T0 * sizeof(float), s>>>(pp);
    CTB_LAUNCH_CHECK();
    if ((rc = tc_gemm_launch<GE_SCALE_RES>(s, h->pf_attn, d, B, T0, d, d, 1, d, 1, 0, Whi + L.wo, Wlo + L.wo, h->pf_zeros,
                                           h->pf_ones, h->pf_resid, d, h->pf_resid, d))) return rc;
    k_rms_rows<<<rms_blocks, 256, 0, s>>>(h->pf_resid, Wl + L.ln2, h->pf_xn, M, d, c.rms_eps);
    CTB_LAUNCH_CHECK();
    if ((rc = tc_gemm_launch<GE_NONE>(s, h->pf_xn, d, B, T0, 2 * I, d, 1, d, 1, 0, Whi + L.wgate_up, Wlo + L.wgate_up,
                                      nullptr, nullptr, nullptr, 0, h->pf_gu, 2 * I))) return rc;
    k_silu_mul<<<(unsigned)(((size_t)M * I + 255) / 256), 256, 0, s>>>(h->pf_gu, h->pf_h, M, I);
    CTB_LAUNCH_CHECK();
    if ((rc = tc_gemm_launch<GE_SCALE_RES>(s, h->pf_h, I, B, T0, d, I, 1, I, 1, 0, Whi + L.wdown, Wlo + L.wdown,
                                           h->pf_zeros, h->pf_ones, h->pf_resid, d, h->pf_resid, d))) return rc;
  }
  k_prefill_finish<<<B, 256, 0, s>>>(h->pf_resid, h->x, h->use_tc ? h->x_hi : nullptr, h->use_tc ? h->x_lo : nullptr,
                                     h->pf_nvalid, h->seq_len, h->pos, h->active, T0, d);
  CTB_LAUNCH_CHECK();
  // first token: heads -> sampler -> finalize (the i == 0 iteration of gpt.py:394)
  const StepCtx x = make_ctx(h, 0);
  if ((rc = h->use_tc ? launch_heads_tc(h, s) : launch_heads(h, x, s))) return rc;
  if ((rc = launch_sampler(h, x, s))) return rc;
  FinalP fp{};
  fp.st = h->st; fp.B = B; fp.rows_per_item = h->infer_text ? 1 : c.num_vq; fp.num_vq = c.num_vq; fp.max_new = h->max_new;
  fp.eos = h->sampler.eos_token; fp.idx = h->idx; fp.ids_out = h->ids_out; fp.finish = h->finish; fp.end_idx = h->end_idx;
  CTB_CUDA(launch_pdl(k_finalize, dim3(1), dim3(256), 0, s, fp));
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

// Pages for B rows of up to `tokens` tokens each: grow the pool if needed (page = 16 tokens x K|V x heads x 64 floats per
// layer) and assign row b the pages [b * need, (b + 1) * need) - kernels only ever see the block table.
static int kv_reserve(ctb_gpt* h, int B, int tokens, cudaStream_t s) {
  const ctb_gpt_config& c = h->cfg;
  const size_t per_row = ((size_t)tokens + kPageTokens - 1) / kPageTokens;
  const size_t need = per_row * (size_t)B;
  const size_t page_floats = (size_t)2 * c.num_kv_heads * kPageTokens * c.head_dim;
  if (need > h->kv_pages) {
    CTB_CUDA(cudaStreamSynchronize(s));
    if (h->kv) { cudaFree(h->kv); h->kv = nullptr; h->kv_pages = 0; }
    const size_t pages = need + need / 8;  // a little head-room against re-allocation on slightly longer calls
    if (cudaMalloc(reinterpret_cast<void**>(&h->kv), pages * page_floats * c.num_layers * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      return set_err(CTB_ERR_NOMEM, "KV pool: %zu pages x %d layers (%.1f GB) do not fit", pages, c.num_layers,
                     (double)(pages * page_floats * c.num_layers * 4) / 1e9);
    }
    CTB_CUDA(cudaMemsetAsync(h->kv, 0, pages * page_floats * c.num_layers * sizeof(float), s));
    h->kv_pages = pages;
    h->kv_layer_floats = pages * page_floats;
    h->bt_B = 0;
  }
  if (h->bt_B == B && h->bt_per_row == per_row) return CTB_OK;  // table already describes this shape: nothing to upload
  std::vector<int> bt((size_t)B * h->pages_per_row, 0);
  for (int b = 0; b < B; ++b)
    for (size_t i = 0; i < per_row; ++i) bt[(size_t)b * h->pages_per_row + i] = (int)((size_t)b * per_row + i);
  CTB_CUDA(cudaMemcpyAsync(h->block_table, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CTB_CUDA(cudaStreamSynchronize(s));  // bt is a host temporary
  h->bt_B = B; h->bt_per_row = per_row;
  return CTB_OK;
}

extern "C" int ctb_gpt_begin(ctb_gpt* h, int32_t B, int32_t T0, const float* emb_dev, const uint8_t* mask_dev,
                             const ctb_sampler_config* sampler, const float* q_noise_dev, int32_t max_new_token,
                             int32_t infer_text, int32_t* ids_out_dev, float* hiddens_out_dev, void* stream) {
  if (!h || !emb_dev || !mask_dev || !sampler || !ids_out_dev) return set_err(CTB_ERR_ARG, "null argument");
  if (B < 1 || B > h->cfg.max_batch) return set_err(CTB_ERR_ARG, "B=%d outside [1,%d]", B, h->cfg.max_batch);
  if (T0 < 1 || max_new_token < 1 || T0 + max_new_token > h->cfg.max_context)
    return set_err(CTB_ERR_ARG, "T0=%d + max_new=%d exceeds max_context=%d", T0, max_new_token, h->cfg.max_context);
  if (sampler->past_window > 31 || sampler->past_window < 0) return set_err(CTB_ERR_ARG, "past_window out of range");
  if (sampler->min_tokens_to_keep < 1) return set_err(CTB_ERR_ARG, "min_tokens_to_keep must be >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  h->B = B; h->T0 = T0; h->max_new = max_new_token; h->infer_text = infer_text ? 1 : 0;
  h->use_tc = h->tc_ready && B >= h->tc_min_batch;
  h->sampler = *sampler; h->q_noise = q_noise_dev; h->emb = emb_dev; h->mask = mask_dev;
  h->ids_out = ids_out_dev; h->hiddens_out = hiddens_out_dev;
  if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
  { int rc0 = kv_reserve(h, B, T0 + max_new_token, s); if (rc0) return rc0; }
  CTB_CUDA(cudaMemsetAsync(h->st, 0, sizeof(LoopState), s));
  CTB_CUDA(cudaMemsetAsync(h->seq_len, 0, sizeof(int) * h->bpad_max, s));
  CTB_CUDA(cudaMemsetAsync(h->finish, 0, h->bpad_max, s));
  CTB_CUDA(cudaMemsetAsync(h->end_idx, 0, sizeof(int) * h->bpad_max, s));
  CTB_CUDA(cudaMemsetAsync(h->counter, 0, sizeof(int) * h->cfg.max_batch * h->cfg.num_heads, s));
  if (h->flow_ok) {
    // the dataflow step numbers its launches from 1 within a generate() call: tag base and arrival counters restart
    static const unsigned e0 = FL_EPOCH_STEP;
    CTB_CUDA(cudaMemcpyAsync(h->flow_epoch, &e0, sizeof(e0), cudaMemcpyHostToDevice, s));
    CTB_CUDA(cudaMemsetAsync(h->flow_arena, 0, FL_ARENA_WORDS * sizeof(unsigned long long), s));
  }
  if (h->use_tc) {
    const size_t d = h->cfg.hidden_size, I = h->cfg.intermediate_size;
    float* z768[] = {h->x_hi, h->x_lo, h->attn_hi, h->attn_lo};
    for (float* zp : z768) CTB_CUDA(cudaMemsetAsync(zp, 0, 32 * d * sizeof(float), s));
    CTB_CUDA(cudaMemsetAsync(h->h_hi, 0, 32 * I * sizeof(float), s));
    CTB_CUDA(cudaMemsetAsync(h->h_lo, 0, 32 * I * sizeof(float), s));
  }
  int rc;
  // prefill: the prompt is walked column by column through the decode kernels (left padding
  // keeps every row's last prompt token in the last column, like the reference's batches)
  if (h->pf_enabled && T0 >= 8 && T0 <= 1024) {
    // whole prompt as token-parallel tcgen05 GEMMs (prefill.cuh)
    if ((rc = prefill_batched(h, s))) return rc;
  } else {
    // short prompts: walk the columns through the decode kernels (left padding keeps every row's last prompt
    // token in the last column, like the reference's batches)
    for (int col = 0; col < T0; ++col)
      if ((rc = enqueue_step(h, col, col == T0 - 1, s))) return rc;
  }
  h->started = 1;
  h->steps_enqueued = 1;
  return CTB_OK;
}

extern "C" int ctb_gpt_decode(ctb_gpt* h, int32_t n_steps, void* stream) {
  if (!h || !h->started) return set_err(CTB_ERR_STATE, "ctb_gpt_begin has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  // ids_out / hiddens_out hold max_new steps and the KV pages max_context tokens: never enqueue past either
  n_steps = std::min(n_steps, h->max_new - h->steps_enqueued);
  if (n_steps <= 0) return CTB_OK;
  h->steps_enqueued += n_steps;
  if (flow_ink(h)) {
    // the whole loop body (step, sampling tail, finish bookkeeping) is inside k_flow: many iterations per launch
    const int per_launch = h->trace ? 1 : 64;
    for (int done = 0; done < n_steps; done += per_launch)
      if ((rc = launch_step_flow(h, -1, true, s, std::min(per_launch, n_steps - done)))) return rc;
    return CTB_OK;
  }
  if (h->use_graph && !h->graph_exec) {
    // capture on a private stream (the caller's may be the legacy default stream, which cannot
    // be captured); the instantiated graph is then launched on the caller's stream
    cudaGraph_t graph;
    if (!h->cap_stream) CTB_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    CTB_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    const uint64_t before = g_launches.load();
    rc = enqueue_step(h, -1, true, h->cap_stream);
    h->graph_kernels = g_launches.load() - before;
    g_launches.store(before);  // captured, not launched
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc) return rc;
    if (e != cudaSuccess) return set_err(CTB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    CTB_CUDA(cudaGraphInstantiate(&h->graph_exec, graph, 0));
    cudaGraphDestroy(graph);
  }
  for (int i = 0; i < n_steps; ++i) {
    if (h->graph_exec) {
      CTB_CUDA(cudaGraphLaunch(h->graph_exec, s));
      g_launches.fetch_add(h->graph_kernels, std::memory_order_relaxed);
    } else if ((rc = enqueue_step(h, -1, true, s))) {
      return rc;
    }
  }
  return CTB_OK;
}

namespace ctb {
// Embed.forward (embed.py:51-79): one CTA per prompt position
__global__ void k_embed_prompt(const int64_t* __restrict__ ids, const uint8_t* __restrict__ text_mask,
                               const float* __restrict__ emb_text, const float* __restrict__ emb_code, int num_vq,
                               int num_audio, int num_text, int d, float* __restrict__ out) {
  const size_t pos = blockIdx.x;
  const int64_t* id = ids + pos * num_vq;
  float* o = out + pos * d;
  if (text_mask[pos]) {
    const int64_t t = min(max(id[0], (int64_t)0), (int64_t)num_text - 1);
    const float* e = emb_text + (size_t)t * d;
    for (int k = threadIdx.x; k < d; k += blockDim.x) o[k] = e[k];
  } else {
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
      float s = 0.f;
      for (int q = 0; q < num_vq; ++q) {
        const int64_t c = min(max(id[q], (int64_t)0), (int64_t)num_audio - 1);
        s += emb_code[((size_t)q * num_audio + c) * d + k];
      }
      o[k] = s;
    }
  }
}
}  // namespace ctb

extern "C" int ctb_gpt_embed_prompt(ctb_gpt* h, const int64_t* ids_dev, const uint8_t* text_mask_dev, int32_t B, int32_t T,
                                    float* out_dev, void* stream) {
  if (!h || !ids_dev || !text_mask_dev || !out_dev) return set_err(CTB_ERR_ARG, "null argument");
  if (B < 1 || T < 1) return set_err(CTB_ERR_ARG, "bad shape");
  const ctb_gpt_config& c = h->cfg;
  k_embed_prompt<<<(unsigned)((size_t)B * T), 256, 0, (cudaStream_t)stream>>>(
      ids_dev, text_mask_dev, h->W + h->lay.emb_text, h->W + h->lay.emb_code, c.num_vq, c.num_audio_tokens,
      c.num_text_tokens, c.hidden_size, out_dev);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

extern "C" int ctb_gpt_debug_trace(ctb_gpt* h, unsigned long long* host_out, int n) {
  if (!h || !h->trace) return set_err(CTB_ERR_STATE, "trace disabled (set CTB_MEGA_TRACE=1 before ctb_gpt_create)");
  CTB_CUDA(cudaMemcpy(host_out, h->trace, sizeof(unsigned long long) * (size_t)std::min(n, 4096), cudaMemcpyDeviceToHost));
  return CTB_OK;
}

extern "C" int ctb_gpt_status_query(ctb_gpt* h, ctb_gpt_status* out, int32_t* end_idx_host, uint8_t* finish_host,
                                    void* stream) {
  if (!h || !out) return set_err(CTB_ERR_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  LoopState st;
  CTB_CUDA(cudaMemcpyAsync(&st, h->st, sizeof(st), cudaMemcpyDeviceToHost, s));
  if (end_idx_host) CTB_CUDA(cudaMemcpyAsync(end_idx_host, h->end_idx, sizeof(int) * h->B, cudaMemcpyDeviceToHost, s));
  if (finish_host) CTB_CUDA(cudaMemcpyAsync(finish_host, h->finish, h->B, cudaMemcpyDeviceToHost, s));
  CTB_CUDA(cudaStreamSynchronize(s));
  if (st.err != 0)
    return set_err(CTB_ERR_STATE, "decode kernel reported error 0x%x (watchdog: a cross-CTA wait never completed)", st.err);
  out->steps_done = st.step;
  out->all_finished = st.all_finished;
  out->any_finished_first_step = st.any_first;
  out->reserved = 0;
  return CTB_OK;
}

extern "C" int ctb_sample(const float* logits_dev, int32_t rows, int32_t V, int32_t rows_per_item,
                          const ctb_sampler_config* sampler, const float* q_noise_dev, const int32_t* gen_ids_dev,
                          int32_t gen_stride, int32_t n_gen, int32_t step, int32_t* out_idx_dev, void* stream) {
  if (!logits_dev || !sampler || !out_idx_dev) return set_err(CTB_ERR_ARG, "null argument");
  if (rows < 1 || V < 1 || rows_per_item < 1 || rows % rows_per_item) return set_err(CTB_ERR_ARG, "bad shape");
  if ((size_t)V * 4 + 8192 > 200 * 1024) return set_err(CTB_ERR_ARG, "V=%d too large for the sampler", V);
  if (sampler->penalty_on && n_gen > 0 && !gen_ids_dev) return set_err(CTB_ERR_ARG, "gen_ids required");
  if (sampler->past_window > 31 || sampler->past_window < 0) return set_err(CTB_ERR_ARG, "past_window out of range");
  if (sampler->min_tokens_to_keep < 1) return set_err(CTB_ERR_ARG, "min_tokens_to_keep must be >= 1");
  SampleP sp{};
  sp.st = nullptr; sp.check_finished = 0; sp.logits = logits_dev; sp.rows = rows; sp.V = V;
  sp.rows_per_item = rows_per_item; sp.cfg = *sampler; sp.q_noise = q_noise_dev; sp.gen_ids = gen_ids_dev;
  sp.gen_stride = gen_stride; sp.gen_inner = rows_per_item; sp.n_gen_fixed = n_gen; sp.step_fixed = step; sp.out_idx = out_idx_dev;
  return launch_sample(sp, (cudaStream_t)stream);
}
