// extern "C" entry points for hot path 2 (token -> waveform); see include/chattts_b200.h.
#define CTB_DECODER_KERNELS_IMPL
#include "tc_gemm.cuh"

using namespace ctb;

__global__ void k_split_tf32(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n);

namespace {

struct BlockOff {  // one ConvNeXt block inside a blob (float offsets)
  int64_t dw_w, dw_b, ln_w, ln_b, pw1_w, pw1_b, pw2_w, pw2_b, gamma;
};

struct DvaeOff {
  int64_t in0_w, in0_b, in2_w, in2_b, conv_out_w, out_conv_w, coef, vq_w, vq_b, total;
  BlockOff blk[64];
};
struct VocosOff {
  int64_t embed_w, embed_b, norm_w, norm_b, fin_w, fin_b, head_w, head_b, basis, window, total;
  BlockOff blk[64];
};

constexpr int MEL = 100, MEL_PAD = 128;

int64_t take(int64_t& off, int64_t n) { const int64_t o = off; off += (n + 3) / 4 * 4; return o; }

void block_layout(BlockOff& b, int64_t& off, int C, int inter) {
  b.dw_w = take(off, 7LL * C); b.dw_b = take(off, C); b.ln_w = take(off, C); b.ln_b = take(off, C);
  b.pw1_w = take(off, (int64_t)inter * C); b.pw1_b = take(off, inter);
  b.pw2_w = take(off, (int64_t)C * inter); b.pw2_b = take(off, C); b.gamma = take(off, C);
}

// Blob order (all fp32, every tensor padded to a multiple of 4 floats); the Python packer in
// chattts_b200/decoder.py writes the same sequence and checks the total against these functions.
DvaeOff dvae_layout(const ctb_convstack_config& c) {
  DvaeOff o{};
  int64_t off = 0;
  o.in0_w = take(off, (int64_t)c.bn_dim * 3 * c.idim);       // [bn, 3*idim]  tap-major K
  o.in0_b = take(off, c.bn_dim);
  o.in2_w = take(off, (int64_t)c.hidden * 3 * c.bn_dim);     // [hidden, 3*bn]
  o.in2_b = take(off, c.hidden);
  for (int i = 0; i < c.n_layer; ++i) block_layout(o.blk[i], off, c.hidden, 4 * c.hidden);
  o.conv_out_w = take(off, (int64_t)c.odim * c.hidden);      // [odim, hidden]
  o.out_conv_w = take(off, (int64_t)MEL_PAD * 3 * c.out_dim); // [128 (100 + zero rows), 3*out_dim]
  o.coef = take(off, MEL_PAD);
  if (c.vq_dim > 0) {
    const int per_group = c.vq_dim / c.vq_groups;
    o.vq_w = take(off, (int64_t)c.vq_groups * per_group * 4); // [G][dim/G][4]
    o.vq_b = take(off, (int64_t)c.vq_groups * per_group);
  }
  o.total = off;
  return o;
}

int spec_k(const ctb_vocos_config& c) { return ((c.n_fft + 2) + 31) / 32 * 32; }  // 1026 -> 1056

VocosOff vocos_layout(const ctb_vocos_config& c) {
  VocosOff o{};
  int64_t off = 0;
  o.embed_w = take(off, (int64_t)c.dim * 7 * MEL_PAD);       // [dim, 7*128] (mel channels padded)
  o.embed_b = take(off, c.dim);
  o.norm_w = take(off, c.dim); o.norm_b = take(off, c.dim);
  for (int i = 0; i < c.num_layers; ++i) block_layout(o.blk[i], off, c.dim, c.intermediate_dim);
  o.fin_w = take(off, c.dim); o.fin_b = take(off, c.dim);
  o.head_w = take(off, (int64_t)spec_k(c) * c.dim);          // rows interleaved (mag_k, phase_k), zero pad rows
  o.head_b = take(off, spec_k(c));
  o.basis = take(off, (int64_t)c.n_fft * spec_k(c));         // [n_fft, spec_k] windowed inverse real DFT
  o.window = take(off, c.n_fft);
  o.total = off;
  return o;
}

}  // namespace

struct ctb_decoder {
  ctb_convstack_config dc;
  ctb_vocos_config vc;
  DvaeOff dl;
  VocosOff vl;
  const float* dW;
  const float* vW;
  int max_batch, max_tokens;
  size_t max_rows;  // max_batch * 2 * max_tokens frames
  size_t cap_rows;  // frames the activation buffers currently hold
  float *bufA, *bufB, *bufH, *mel_tm, *staged_in;
  // tcgen05 path: tf32-rounded hi / lo copies of both blobs (same offsets as the fp32 blobs)
  float *dW_hi, *dW_lo, *vW_hi, *vW_lo;
  bool use_tc;
};

extern "C" int64_t ctb_dvae_blob_floats(const ctb_convstack_config* c) { return c ? dvae_layout(*c).total : -1; }
extern "C" int64_t ctb_vocos_blob_floats(const ctb_vocos_config* c) { return c ? vocos_layout(*c).total : -1; }

extern "C" int ctb_decoder_destroy(ctb_decoder* h) {
  if (!h) return CTB_OK;
  void* ptrs[] = {h->bufA, h->bufB, h->bufH, h->mel_tm, h->staged_in, h->dW_hi, h->dW_lo, h->vW_hi, h->vW_lo};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return CTB_OK;
}

// grow the activation buffers to `rows` frames (time-major [rows, C]); the contents are scratch
static int dec_reserve(ctb_decoder* h, size_t rows, cudaStream_t s) {
  if (rows <= h->cap_rows) return CTB_OK;
  CTB_CUDA(cudaStreamSynchronize(s));
  float** bufs[] = {&h->bufA, &h->bufB, &h->bufH, &h->mel_tm, &h->staged_in};
  for (float** b : bufs) if (*b) { cudaFree(*b); *b = nullptr; }
  h->cap_rows = 0;
  const ctb_convstack_config& dc = h->dc;
  const ctb_vocos_config& vc = h->vc;
  const size_t R = std::min(h->max_rows, rows + rows / 8);
  const size_t wide = std::max((size_t)std::max(4 * dc.hidden, vc.intermediate_dim), (size_t)vc.n_fft);
  const size_t narrow = std::max((size_t)std::max(std::max(dc.hidden, dc.idim), vc.dim), (size_t)dc.odim);
  cudaError_t e = cudaSuccess;
  auto A = [&](float** p, size_t n) { if (e == cudaSuccess) e = cudaMalloc((void**)p, n * sizeof(float)); };
  A(&h->bufA, R * std::max(narrow, (size_t)spec_k(vc)));
  A(&h->bufB, R * narrow);
  A(&h->bufH, R * wide);
  A(&h->mel_tm, R * MEL_PAD);
  A(&h->staged_in, R * dc.idim);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err(CTB_ERR_NOMEM, "decoder buffers for %zu frames: %s", R, cudaGetErrorString(e));
  }
  h->cap_rows = R;
  return CTB_OK;
}

extern "C" int ctb_decoder_create(const ctb_convstack_config* dc, const float* dvae_blob_dev,
                                  const ctb_vocos_config* vc, const float* vocos_blob_dev, int32_t max_batch,
                                  int32_t max_tokens, ctb_decoder** out) {
  // either blob may be NULL: the handle then serves only the other half (DVAE-only / Vocos-only)
  if (!dc || !vc || (!dvae_blob_dev && !vocos_blob_dev) || !out) return set_err(CTB_ERR_ARG, "null argument");
  if (dc->kernel != 7 || dc->n_layer > 64 || vc->num_layers > 64) return set_err(CTB_ERR_ARG, "unsupported conv stack");
  if (dc->idim % 16 || dc->bn_dim % 16 || dc->hidden % 128 || dc->hidden > 512 || dc->odim % 16 || dc->out_dim % 16)
    return set_err(CTB_ERR_ARG, "dvae channel counts must be multiples of 16 (hidden: of 128, <= 512)");
  if (vc->input_channels != MEL || vc->dim % 128 || vc->dim > 512 || vc->intermediate_dim % 16 || vc->n_fft % 16 ||
      vc->hop_length * 4 != vc->n_fft)
    return set_err(CTB_ERR_ARG, "unsupported vocos shape");
  if (dc->vq_dim > 0 && (dc->vq_levels < 2 || dc->vq_dim % dc->vq_groups || dc->vq_dim / dc->vq_groups != dc->idim))
    return set_err(CTB_ERR_ARG, "vq_dim / vq_groups must equal the stack's idim");
  int ndev = 0;
  CTB_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return set_err(CTB_ERR_CUDA, "no CUDA device: chattts_b200 has no CPU path");
  ctb_decoder* h = new ctb_decoder();
  memset(h, 0, sizeof(*h));
  h->dc = *dc; h->vc = *vc; h->dl = dvae_layout(*dc); h->vl = vocos_layout(*vc);
  h->dW = dvae_blob_dev; h->vW = vocos_blob_dev;
  h->max_batch = max_batch; h->max_tokens = max_tokens;
  h->max_rows = (size_t)max_batch * 2 * max_tokens;
  cudaError_t e = cudaSuccess;
  auto A = [&](float** p, size_t n) { if (e == cudaSuccess) e = cudaMalloc((void**)p, n * sizeof(float)); };
  // activation buffers are sized by the largest call seen so far (dec_reserve), not by max_batch x max_tokens
  h->use_tc = getenv("CTB_DECODER_FMA") == nullptr;
  if (h->use_tc) {
    if (h->dW) { A(&h->dW_hi, h->dl.total); A(&h->dW_lo, h->dl.total); }
    if (h->vW) { A(&h->vW_hi, h->vl.total); A(&h->vW_lo, h->vl.total); }
  }
  if (e != cudaSuccess) {
    ctb_decoder_destroy(h);
    return set_err(CTB_ERR_NOMEM, "decoder buffers: %s", cudaGetErrorString(e));
  }
  if (h->use_tc) {
    if (h->dW) k_split_tf32<<<1024, 256>>>(h->dW, h->dW_hi, h->dW_lo, h->dl.total);
    if (h->vW) k_split_tf32<<<1024, 256>>>(h->vW, h->vW_hi, h->vW_lo, h->vl.total);
    CTB_CUDA(cudaDeviceSynchronize());
  }
  *out = h;
  return CTB_OK;
}

// ------------------------------------------------------------------ launch helpers
__global__ void k_split_tf32(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = w[i], h = to_tf32(x);
    hi[i] = h;
    lo[i] = to_tf32(x - h);
  }
}

struct GemmCtx {  // which blob a weight pointer lives in, for the hi / lo lookup of the tensor-core path
  const float* W; const float* hi; const float* lo; bool tc;
};

template <int EPI>
static int gemm(cudaStream_t s, const GemmCtx& gc, const float* A, int lda, int M, int N, int K, int taps, int Cin,
                int dil, int pad, int F, const float* W, const float* bias, const float* gamma, const float* res,
                int ldres, float* C, int ldc) {
  if (gc.tc && K % TC_BK == 0 && Cin % TC_BK == 0 && M % F == 0)
    return tc_gemm_launch<EPI>(s, A, lda, M / F, F, N, K, taps, Cin, dil, pad, gc.hi + (W - gc.W), gc.lo + (W - gc.W), bias,
                               gamma, res, ldres, C, ldc);
  GemmP p{};
  p.A = A; p.lda = lda; p.M = M; p.N = N; p.K = K; p.taps = taps; p.Cin = Cin; p.dil = dil; p.pad = pad; p.F = F;
  p.W = W; p.bias = bias; p.gamma = gamma; p.res = res; p.ldres = ldres; p.C = C; p.ldc = ldc;
  dim3 grid((N + GBN - 1) / GBN, (M + GBM - 1) / GBM);
  k_sgemm_nt<EPI><<<grid, 256, 0, s>>>(p);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

static int dwln(cudaStream_t s, const float* x, float* out, int M, int F, int C, int taps, int dil, const float* w,
                const float* b, const float* lnw, const float* lnb) {
  DwLnP p{x, out, M, F, C, taps, dil, w, b, lnw, lnb, 1e-6f};
  const int blocks = (M + 7) / 8;
  switch (C / 128) {
    case 1: k_dwconv_ln<1><<<blocks, 256, 0, s>>>(p); break;
    case 2: k_dwconv_ln<2><<<blocks, 256, 0, s>>>(p); break;
    case 3: k_dwconv_ln<3><<<blocks, 256, 0, s>>>(p); break;
    default: k_dwconv_ln<4><<<blocks, 256, 0, s>>>(p); break;
  }
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

// x (time-major [M, C]) -> x through one ConvNeXt block; tmp = [M, C], hbuf = [M, inter]
static int convnext(cudaStream_t s, const GemmCtx& gc, const float* W, const BlockOff& b, float* x, float* tmp, float* hbuf, int M,
                    int F, int C, int inter, int dil) {
  int rc;
  if ((rc = dwln(s, x, tmp, M, F, C, 7, dil, W + b.dw_w, W + b.dw_b, W + b.ln_w, W + b.ln_b))) return rc;
  if ((rc = gemm<GE_GELU>(s, gc, tmp, C, M, inter, C, 1, C, 1, 0, F, W + b.pw1_w, W + b.pw1_b, nullptr, nullptr, 0, hbuf,
                          inter))) return rc;
  return gemm<GE_SCALE_RES>(s, gc, hbuf, inter, M, C, inter, 1, inter, 1, 0, F, W + b.pw2_w, W + b.pw2_b, W + b.gamma, x,
                            C, x, C);
}

// in_layout: 0 = channels-first [B, C, T] fp32 (DVAE.__call__ layout), 1 = token-major [B, T, C] fp32
// (the decode loop's hidden states; frame doubling is then a re-interpretation), 2 = codes [B, G*R, T] int32
static int dvae_run(ctb_decoder* h, const void* in, int layout, int B, int T, float* mel_cf, cudaStream_t s) {
  const ctb_convstack_config& c = h->dc;
  const DvaeOff& L = h->dl;
  const float* W = h->dW;
  const int F = 2 * T, M = B * F;
  const GemmCtx gc{h->dW, h->dW_hi, h->dW_lo, h->use_tc};
  int rc;
  const float* x0;
  if (layout == 1) {
    x0 = static_cast<const float*>(in);
  } else if (layout == 0) {
    dim3 g((T + 31) / 32, (2 * c.idim + 31) / 32, B);
    k_cf_to_tm_doubled<<<g, dim3(32, 8), 0, s>>>(static_cast<const float*>(in), h->staged_in, B, 2 * c.idim, T);
    CTB_LAUNCH_CHECK();
    x0 = h->staged_in;
  } else {
    if (c.vq_dim <= 0) return set_err(CTB_ERR_ARG, "this decoder has no VQ layer (use_decoder=True model)");
    if (c.vq_groups != 2) return set_err(CTB_ERR_ARG, "vq_groups must be 2 (one group per doubled frame)");
    GfsqP g{};
    g.ids = static_cast<const int32_t*>(in); g.out = h->staged_in; g.B = B; g.T = T; g.G = c.vq_groups;
    g.R = c.vq_residual; g.levels = c.vq_levels & 0xff; g.nlev = 4; g.per_group = c.vq_dim / c.vq_groups;
    g.scale_base = (float)(((c.vq_levels >> 8) & 0xff) ? ((c.vq_levels >> 8) & 0xff) : (g.levels - 1));
    g.w = W + L.vq_w; g.b = W + L.vq_b;
    k_gfsq_dequant<<<B * T * c.vq_groups, 128, 0, s>>>(g);
    CTB_LAUNCH_CHECK();
    x0 = h->staged_in;
  }
  // conv_in: Conv1d(idim -> bn, k3, p1) + GELU + Conv1d(bn -> hidden, k3, p1)   (dvae.py:144-148)
  if ((rc = gemm<GE_GELU>(s, gc, x0, c.idim, M, c.bn_dim, 3 * c.idim, 3, c.idim, 1, 1, F, W + L.in0_w, W + L.in0_b,
                          nullptr, nullptr, 0, h->bufB, c.bn_dim))) return rc;
  if ((rc = gemm<GE_BIAS>(s, gc, h->bufB, c.bn_dim, M, c.hidden, 3 * c.bn_dim, 3, c.bn_dim, 1, 1, F, W + L.in2_w,
                          W + L.in2_b, nullptr, nullptr, 0, h->bufA, c.hidden))) return rc;
  for (int i = 0; i < c.n_layer; ++i)
    if ((rc = convnext(s, gc, W, L.blk[i], h->bufA, h->bufB, h->bufH, M, F, c.hidden, 4 * c.hidden, c.dilation))) return rc;
  // conv_out 1x1 (no bias), out_conv k3 (no bias) * coef        (dvae.py:159,236,289-297)
  if ((rc = gemm<GE_NONE>(s, gc, h->bufA, c.hidden, M, c.odim, c.hidden, 1, c.hidden, 1, 0, F, W + L.conv_out_w, nullptr,
                          nullptr, nullptr, 0, h->bufB, c.odim))) return rc;
  if ((rc = gemm<GE_COEF>(s, gc, h->bufB, c.out_dim, M, MEL_PAD, 3 * c.out_dim, 3, c.out_dim, 1, 1, F, W + L.out_conv_w,
                          nullptr, W + L.coef, nullptr, 0, h->mel_tm, MEL_PAD))) return rc;
  if (mel_cf) {
    dim3 g((F + 31) / 32, (MEL + 31) / 32, B);
    k_tm_to_cf<<<g, dim3(32, 8), 0, s>>>(h->mel_tm, mel_cf, B, MEL, F, MEL_PAD);
    CTB_LAUNCH_CHECK();
  }
  return CTB_OK;
}

static int vocos_run(ctb_decoder* h, const float* mel_cf, int B, int F, float* wav, cudaStream_t s) {
  const ctb_vocos_config& c = h->vc;
  const VocosOff& L = h->vl;
  const float* W = h->vW;
  const int M = B * F, SK = spec_k(c);
  const GemmCtx gc{h->vW, h->vW_hi, h->vW_lo, h->use_tc};
  int rc;
  if (mel_cf) {
    dim3 g((F + 31) / 32, (MEL_PAD + 31) / 32, B);
    k_cf_to_tm<<<g, dim3(32, 8), 0, s>>>(mel_cf, h->mel_tm, B, MEL, F, MEL_PAD);
    CTB_LAUNCH_CHECK();
  }
  // backbone: Conv1d(100 -> dim, k7, p3) -> LN -> ConvNeXt x num_layers -> LN
  if ((rc = gemm<GE_BIAS>(s, gc, h->mel_tm, MEL_PAD, M, c.dim, 7 * MEL_PAD, 7, MEL_PAD, 1, 3, F, W + L.embed_w,
                          W + L.embed_b, nullptr, nullptr, 0, h->bufB, c.dim))) return rc;
  if ((rc = dwln(s, h->bufB, h->bufA, M, F, c.dim, 0, 1, nullptr, nullptr, W + L.norm_w, W + L.norm_b))) return rc;
  for (int i = 0; i < c.num_layers; ++i)
    if ((rc = convnext(s, gc, W, L.blk[i], h->bufA, h->bufB, h->bufH, M, F, c.dim, c.intermediate_dim, 1))) return rc;
  if ((rc = dwln(s, h->bufA, h->bufB, M, F, c.dim, 0, 1, nullptr, nullptr, W + L.fin_w, W + L.fin_b))) return rc;
  // ISTFTHead: Linear(dim -> n_fft + 2) -> (mag, phase) -> complex spectrum (interleaved re/im)
  if ((rc = gemm<GE_SPEC>(s, gc, h->bufB, c.dim, M, SK, c.dim, 1, c.dim, 1, 0, F, W + L.head_w, W + L.head_b, nullptr,
                          nullptr, 0, h->bufA, SK))) return rc;
  // inverse real DFT * window as a GEMM against the constant basis, then overlap-add / envelope
  if ((rc = gemm<GE_NONE>(s, gc, h->bufA, SK, M, c.n_fft, SK, 1, SK, 1, 0, F, W + L.basis, nullptr, nullptr, nullptr, 0,
                          h->bufH, c.n_fft))) return rc;
  const size_t total = (size_t)B * c.hop_length * (F - 1);
  k_overlap_add<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->bufH, W + L.window, wav, B, F, c.n_fft, c.hop_length);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

extern "C" int ctb_dvae_decode(ctb_decoder* h, const void* in_dev, int32_t in_layout, int32_t B, int32_t T,
                               float* mel_dev, void* stream) {
  if (!h || !in_dev) return set_err(CTB_ERR_ARG, "null argument");
  if (!h->dW) return set_err(CTB_ERR_STATE, "handle was created without DVAE weights");
  if (B < 1 || T < 1 || (size_t)B * 2 * T > h->max_rows)
    return set_err(CTB_ERR_ARG, "B=%d x T=%d exceeds this handle (max_batch=%d, max_tokens=%d)", B, T, h->max_batch,
                   h->max_tokens);
  if (in_layout < 0 || in_layout > 2) return set_err(CTB_ERR_ARG, "bad in_layout");
  { int rc = dec_reserve(h, (size_t)B * 2 * T, (cudaStream_t)stream); if (rc) return rc; }
  return dvae_run(h, in_dev, in_layout, B, T, mel_dev, (cudaStream_t)stream);
}

extern "C" int ctb_vocos_decode(ctb_decoder* h, const float* mel_dev, int32_t B, int32_t F, float* wav_dev,
                                void* stream) {
  if (!h || !wav_dev) return set_err(CTB_ERR_ARG, "null argument");
  if (!h->vW) return set_err(CTB_ERR_STATE, "handle was created without Vocos weights");
  if (B < 1 || F < 2 || (size_t)B * F > h->max_rows) return set_err(CTB_ERR_ARG, "B x F exceeds this handle");
  if (mel_dev == nullptr && (size_t)B * F > h->cap_rows)
    return set_err(CTB_ERR_STATE, "no mel of this shape was left in the handle by ctb_dvae_decode");
  { int rc = dec_reserve(h, (size_t)B * F, (cudaStream_t)stream); if (rc) return rc; }
  return vocos_run(h, mel_dev, B, F, wav_dev, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ DVAE encode branch (speaker enrolment)
// DVAE.forward(mode="encode") (dvae.py:265-274): wav -> log-mel / coef -> downsample_conv -> encoder stack -> GFSQ indices.
namespace {

constexpr int ENC_NFFT = 1024, ENC_HOP = 256, ENC_NBIN = ENC_NFFT / 2 + 1, ENC_LDMAG = 516;  // MelSpectrogramFeatures defaults (dvae.py:176-181)

struct EncOff {
  int64_t window, fb, coef, ds0_w, ds0_b, ds1_w, ds1_b, in0_w, in0_b, in2_w, in2_b, conv_out_w, vq_w, vq_b, total;
  BlockOff blk[64];
};

// Blob order of the encode branch; chattts_b200/decoder.py::pack_dvae_encoder writes the same sequence.
EncOff enc_layout(const ctb_convstack_config& c) {
  EncOff o{};
  int64_t off = 0;
  o.window = take(off, ENC_NFFT);                             // analysis window (periodic Hann, fp32 like torch.hann_window)
  o.fb = take(off, (int64_t)ENC_NBIN * MEL_PAD);              // [513][128] mel filterbank, bin-major
  o.coef = take(off, MEL_PAD);
  o.ds0_w = take(off, (int64_t)c.idim * 3 * MEL_PAD);         // Conv1d(100 -> dim, k3, p1), mel channels padded to 128
  o.ds0_b = take(off, c.idim);
  o.ds1_w = take(off, (int64_t)c.idim * 3 * 2 * c.idim);      // Conv1d(dim -> dim, k4, s2, p1) over frame PAIRS: 3 taps x 2*dim
  o.ds1_b = take(off, c.idim);
  o.in0_w = take(off, (int64_t)c.bn_dim * 3 * c.idim);
  o.in0_b = take(off, c.bn_dim);
  o.in2_w = take(off, (int64_t)c.hidden * 3 * c.bn_dim);
  o.in2_b = take(off, c.hidden);
  for (int i = 0; i < c.n_layer; ++i) block_layout(o.blk[i], off, c.hidden, 4 * c.hidden);
  o.conv_out_w = take(off, (int64_t)c.odim * c.hidden);
  const int nlev = 4;
  o.vq_w = take(off, (int64_t)c.vq_groups * nlev * (c.vq_dim / c.vq_groups));  // project_in [G][4][dim/G]
  o.vq_b = take(off, (int64_t)c.vq_groups * nlev);
  o.total = off;
  return o;
}

}  // namespace

struct ctb_encoder {
  ctb_convstack_config c;
  EncOff L;
  const float* W;
  float *W_hi, *W_lo;
  bool use_tc;
  int64_t max_samples;
  int max_frames;
  float *padded, *spec, *mel_tm, *bufX, *bufY, *bufA, *bufB, *bufH;
};

extern "C" int64_t ctb_dvae_encoder_blob_floats(const ctb_convstack_config* c) { return c ? enc_layout(*c).total : -1; }

extern "C" int ctb_dvae_encoder_destroy(ctb_encoder* h) {
  if (!h) return CTB_OK;
  void* ptrs[] = {h->W_hi, h->W_lo, h->padded, h->spec, h->mel_tm, h->bufX, h->bufY, h->bufA, h->bufB, h->bufH};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return CTB_OK;
}

extern "C" int ctb_dvae_encoder_create(const ctb_convstack_config* c, const float* blob_dev, int64_t max_samples,
                                       ctb_encoder** out) {
  if (!c || !blob_dev || !out) return set_err(CTB_ERR_ARG, "null argument");
  if (c->kernel != 7 || c->n_layer > 64 || c->idim % 32 || c->bn_dim % 16 || c->hidden % 128 || c->hidden > 512 || c->odim % 16)
    return set_err(CTB_ERR_ARG, "unsupported encoder stack");
  if (c->vq_dim != c->odim || c->vq_groups < 1 || c->vq_dim % c->vq_groups || c->vq_residual < 1 || (c->vq_levels & 0xff) < 2)
    return set_err(CTB_ERR_ARG, "the encoder's odim must equal vq_dim (GFSQ input)");
  if (max_samples <= ENC_NFFT / 2 || max_samples > (int64_t)1 << 28) return set_err(CTB_ERR_ARG, "max_samples out of range");
  int ndev = 0;
  CTB_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return set_err(CTB_ERR_CUDA, "no CUDA device: chattts_b200 has no CPU path");
  ctb_encoder* h = new ctb_encoder();
  memset(h, 0, sizeof(*h));
  h->c = *c; h->L = enc_layout(*c); h->W = blob_dev; h->max_samples = max_samples;
  h->max_frames = (int)(max_samples / ENC_HOP) + 1;
  h->use_tc = getenv("CTB_DECODER_FMA") == nullptr;
  const size_t F = h->max_frames, FP = (F + 1) / 2;
  cudaError_t e = cudaSuccess;
  auto A = [&](float** p, size_t n) { if (e == cudaSuccess) e = cudaMalloc((void**)p, n * sizeof(float)); };
  A(&h->padded, (F + 3) * ENC_HOP);
  A(&h->spec, F * ENC_LDMAG);
  A(&h->mel_tm, F * MEL_PAD);
  A(&h->bufX, 2 * FP * c->idim);
  A(&h->bufY, FP * c->idim);
  A(&h->bufA, FP * std::max(c->hidden, c->bn_dim));
  A(&h->bufB, FP * std::max(std::max(c->hidden, c->bn_dim), c->odim));
  A(&h->bufH, FP * 4 * c->hidden);
  if (h->use_tc) { A(&h->W_hi, h->L.total); A(&h->W_lo, h->L.total); }
  if (e != cudaSuccess) {
    cudaGetLastError();
    ctb_dvae_encoder_destroy(h);
    return set_err(CTB_ERR_NOMEM, "encoder buffers: %s", cudaGetErrorString(e));
  }
  if (h->use_tc) {
    k_split_tf32<<<1024, 256>>>(h->W, h->W_hi, h->W_lo, h->L.total);
    CTB_CUDA(cudaDeviceSynchronize());
  }
  *out = h;
  return CTB_OK;
}

extern "C" int ctb_dvae_encode(ctb_encoder* h, const float* wav_dev, int64_t n_samples, int32_t* ids_dev,
                               int32_t ids_capacity_tokens, int32_t* n_tokens_out, float* mel_dev, float* margin_dev,
                               void* stream) {
  if (!h || !wav_dev || !ids_dev || !n_tokens_out) return set_err(CTB_ERR_ARG, "null argument");
  if (n_samples <= ENC_NFFT / 2) return set_err(CTB_ERR_ARG, "reflect padding needs more than %d samples", ENC_NFFT / 2);
  if (n_samples > h->max_samples) return set_err(CTB_ERR_ARG, "%lld samples exceed this handle (max_samples=%lld)",
                                                 (long long)n_samples, (long long)h->max_samples);
  const ctb_convstack_config& c = h->c;
  const EncOff& L = h->L;
  const float* W = h->W;
  cudaStream_t s = (cudaStream_t)stream;
  const int F = (int)(n_samples / ENC_HOP) + 1;     // torch.stft(center=True)
  const int T = F / 2;                              // Conv1d(k4, s2, p1): floor((F + 2 - 4) / 2) + 1
  *n_tokens_out = T;
  if (T < 1) return set_err(CTB_ERR_ARG, "audio too short for one token");
  if (T > ids_capacity_tokens) return set_err(CTB_ERR_ARG, "ids buffer holds %d tokens, %d needed", ids_capacity_tokens, T);
  const GemmCtx gc{h->W, h->W_hi, h->W_lo, h->use_tc};
  int rc;
  // framing (reflect padding) -> |STFT| as a double-precision direct DFT -> mel filterbank, log, / coef
  const int64_t total = (int64_t)(F + 3) * ENC_HOP;
  k_reflect_pad<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(wav_dev, h->padded, n_samples, total, ENC_NFFT / 2);
  CTB_LAUNCH_CHECK();
  k_stft_mag<ENC_NFFT><<<dim3(F, (ENC_NBIN + 255) / 256), 256, 0, s>>>(h->padded, W + L.window, ENC_HOP, ENC_NBIN, h->spec, ENC_LDMAG);
  CTB_LAUNCH_CHECK();
  k_mel_log<MEL_PAD><<<F, MEL_PAD, ENC_NBIN * sizeof(float), s>>>(h->spec, ENC_LDMAG, ENC_NBIN, W + L.fb, W + L.coef, MEL, h->mel_tm);
  CTB_LAUNCH_CHECK();
  if (mel_dev) {
    dim3 g((F + 31) / 32, (MEL + 31) / 32, 1);
    k_tm_to_cf<<<g, dim3(32, 8), 0, s>>>(h->mel_tm, mel_dev, 1, MEL, F, MEL_PAD);
    CTB_LAUNCH_CHECK();
  }
  // downsample_conv (dvae.py:231-236): Conv1d(100 -> dim, k3, p1) + GELU; Conv1d(dim -> dim, k4, s2, p1) + GELU.
  // The stride-2 conv runs over frame pairs [F/2 rows, 2*dim]: out[t] = W0 x[2t-1] + W1 x[2t] + W2 x[2t+1] + W3 x[2t+2]
  // = 3 taps over pair rows with the weights re-packed (zeros where a pair half is not touched).
  const int FP = (F + 1) / 2;
  if (F & 1) CTB_CUDA(cudaMemsetAsync(h->bufX + (size_t)F * c.idim, 0, (size_t)c.idim * sizeof(float), s));
  if ((rc = gemm<GE_GELU>(s, gc, h->mel_tm, MEL_PAD, F, c.idim, 3 * MEL_PAD, 3, MEL_PAD, 1, 1, F, W + L.ds0_w, W + L.ds0_b,
                          nullptr, nullptr, 0, h->bufX, c.idim))) return rc;
  if ((rc = gemm<GE_GELU>(s, gc, h->bufX, 2 * c.idim, FP, c.idim, 3 * 2 * c.idim, 3, 2 * c.idim, 1, 1, FP, W + L.ds1_w,
                          W + L.ds1_b, nullptr, nullptr, 0, h->bufY, c.idim))) return rc;
  // encoder = DVAEDecoder(idim -> odim) over the first T rows (dvae.py:131-172)
  if ((rc = gemm<GE_GELU>(s, gc, h->bufY, c.idim, T, c.bn_dim, 3 * c.idim, 3, c.idim, 1, 1, T, W + L.in0_w, W + L.in0_b,
                          nullptr, nullptr, 0, h->bufB, c.bn_dim))) return rc;
  if ((rc = gemm<GE_BIAS>(s, gc, h->bufB, c.bn_dim, T, c.hidden, 3 * c.bn_dim, 3, c.bn_dim, 1, 1, T, W + L.in2_w, W + L.in2_b,
                          nullptr, nullptr, 0, h->bufA, c.hidden))) return rc;
  for (int i = 0; i < c.n_layer; ++i)
    if ((rc = convnext(s, gc, W, L.blk[i], h->bufA, h->bufB, h->bufH, T, T, c.hidden, 4 * c.hidden, c.dilation))) return rc;
  if ((rc = gemm<GE_NONE>(s, gc, h->bufA, c.hidden, T, c.odim, c.hidden, 1, c.hidden, 1, 0, T, W + L.conv_out_w, nullptr,
                          nullptr, nullptr, 0, h->bufB, c.odim))) return rc;
  FsqQuantP q{};
  q.x = h->bufB; q.ids = ids_dev; q.margin = margin_dev; q.T = T; q.G = c.vq_groups; q.R = c.vq_residual;
  q.levels = c.vq_levels & 0xff; q.nlev = 4; q.per_group = c.vq_dim / c.vq_groups;
  q.scale_base = (float)(((c.vq_levels >> 8) & 0xff) ? ((c.vq_levels >> 8) & 0xff) : (q.levels - 1));
  q.bound_input = ((c.vq_levels >> 16) & 1) ? 0 : 1;
  q.w = W + L.vq_w; q.b = W + L.vq_b;
  k_fsq_quant<<<T * c.vq_groups, 128, 0, s>>>(q);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}
