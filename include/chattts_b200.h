/*
 * chattts_b200 - C ABI of the B200-native (sm_100a) ChatTTS hot paths.
 *
 * The reference (2noise/ChatTTS) has no FFI boundary: its seams are Python objects
 * (SURVEY.md 8b).  This header is the boundary a maintainer binds from Python with
 * ctypes; every entry point names the reference interface it
 * replaces.  Conventions:
 *   - plain C types, raw *device* pointers + sizes + a cudaStream_t passed as void*;
 *     no torch types; the caller owns every buffer it passes in;
 *   - int return: 0 = ok, negative = error, message via ctb_last_error() (thread local);
 *   - one handle per device per model; a handle is not re-entrant, different handles are;
 *   - nothing here ever runs on the CPU: if no CUDA device is usable, calls fail.
 */
#ifndef CHATTTS_B200_H
#define CHATTTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTB_OK 0
#define CTB_ERR_ARG (-1)
#define CTB_ERR_CUDA (-2)
#define CTB_ERR_STATE (-3)
#define CTB_ERR_NOMEM (-4)

#define CTB_ABI_VERSION 3

/* ---- library ------------------------------------------------------------------- */
int ctb_abi_version(void);
const char* ctb_last_error(void);
/* number of kernels this library has launched in this process (bench "gpu_launches") */
uint64_t ctb_launch_count(void);

/* ---- GPT decode loop: replaces ChatTTS/model/gpt.py:315-618 (GPT.generate) ------ */

/* Model shape.  Mirrors the fields of the HF LlamaConfig the reference reads
 * (gpt.py:52,75) plus the Embed sizes (embed.py:8-35). */
typedef struct ctb_gpt_config {
  int32_t hidden_size;        /* 768  */
  int32_t intermediate_size;  /* 3072 */
  int32_t num_layers;         /* 20   */
  int32_t num_heads;          /* 12   */
  int32_t num_kv_heads;       /* 12 (GQA group = num_heads / num_kv_heads) */
  int32_t head_dim;           /* 64   */
  int32_t num_vq;             /* 4    */
  int32_t num_audio_tokens;   /* 626  */
  int32_t num_text_tokens;    /* 21178 */
  int32_t max_positions;      /* 4096: rows of the RoPE table */
  float rms_eps;              /* 1e-6 */
  int32_t max_batch;          /* rows this handle can decode at once */
  int32_t max_context;        /* prompt + generated tokens per row (KV pages are sized from it) */
} ctb_gpt_config;

/* Element offsets (in floats) of every tensor inside the packed fp32 weight blob the
 * caller uploads (one cudaMalloc / one NCCL broadcast).  Per-layer tensors are at
 * layer0 + l * layer_stride + <field>. Row-major [out_features, in_features], i.e. the
 * nn.Linear layout of the checkpoint (SURVEY.md 8b state-dict names). */
typedef struct ctb_gpt_layout {
  int64_t layer0, layer_stride;
  int64_t wqkv;      /* [(Hq + 2*Hkv)*hd, d]  q rows, then k rows, then v rows */
  int64_t wo;        /* [d, Hq*hd] */
  int64_t wgate_up;  /* [2*I, d]  gate rows then up rows */
  int64_t wdown;     /* [d, I] */
  int64_t ln1, ln2;  /* [d] each */
  int64_t final_norm;  /* [d] */
  int64_t head_code;   /* [num_vq*num_audio, d]  weight-norm already folded (embed.py:23-35) */
  int64_t head_text;   /* [num_text, d]          weight-norm already folded */
  int64_t emb_code;    /* [num_vq*num_audio, d] */
  int64_t emb_text;    /* [num_text, d] */
  int64_t rope_cos;    /* [max_positions, hd]  host-built exactly like HF LlamaRotaryEmbedding */
  int64_t rope_sin;    /* [max_positions, hd] */
  int64_t total;       /* floats in the blob */
} ctb_gpt_layout;

int ctb_gpt_layout_query(const ctb_gpt_config* cfg, ctb_gpt_layout* out);

/* Sampling tail: gpt.py:487-508 + processors.py:18-58 + HF TopP/TopK warpers.
 * Order: temperature -> repetition penalty -> top-p -> top-k -> [greedy mask] ->
 * (step < min_new: EOS ban) -> softmax -> argmax(p / q)  (== torch.multinomial). */
typedef struct ctb_sampler_config {
  float temperature[8];     /* per codebook (row r uses temperature[r % rows_per_item]) */
  float top_p;              /* <0: warper absent (gen_logits top_P=None) */
  int32_t top_k;            /* <=0: warper absent */
  int32_t min_tokens_to_keep; /* 3 (processors.py:45,47) */
  int32_t penalty_on;       /* 0: no CustomRepetitionPenaltyLogitsProcessorRepeat (penalty == 1) */
  float penalty_lut[32];    /* penalty ** count for count = 0..past_window, built by the host
                               with the same torch.pow call as processors.py:28 */
  int32_t past_window;      /* 16 */
  int32_t penalty_max_ids;  /* rows >= this get no penalty (processors.py:24-27 quirk) */
  int32_t greedy;           /* 1: keep only the row arg-max before softmax (bench config C2);
                               2: arg-max over the non-EOS tokens (EOS removed as well) */
  int32_t eos_token;
  int32_t min_new_token;
  float top_p_removed_max;  /* float32(1 - top_P) evaluated by the host exactly like HF TopPLogitsWarper (python double
                               arithmetic, then the fp32 comparison `cum <= 1 - top_p`); used when has_removed_max != 0,
                               otherwise the kernel derives it from the fp32 top_p (ABI v1 behaviour) */
  int32_t has_removed_max;
  uint64_t philox_seed;     /* used only when q_noise == NULL (manual_seed=None path) */
} ctb_sampler_config;

typedef struct ctb_gpt ctb_gpt;

/* weights_dev: packed blob laid out per ctb_gpt_layout_query, already on the device and
 * kept alive by the caller for the life of the handle. */
int ctb_gpt_create(const ctb_gpt_config* cfg, const float* weights_dev, ctb_gpt** out);
int ctb_gpt_destroy(ctb_gpt* h);

/* Start one generate() call.  Replaces gpt.py:343-381 (buffer set-up) and the i == 0
 * iteration (prefill + first sample).
 *   emb_dev      [B, T0, d] fp32   prompt embeddings (Embed.forward output, embed.py:51-79)
 *   mask_dev     [B, T0]   uint8   attention_mask; every row's valid tokens must be a contiguous SUFFIX (left padding,
 *                                  tokenizer.py:79-110) with the last column valid - the host mirror checks it
 *   q_noise_dev  [rows, V] fp32    Exp(1) noise of the seeded torch generator (gpt.py:504-508),
 *                                  rows = B*num_vq (audio) or B (text); NULL => device Philox
 *   ids_out_dev  [B, max_new, num_vq] int32   sampled ids (text: id replicated, gpt.py:521-523)
 *   hiddens_out_dev [B, max_new, d] fp32 or NULL (return_hidden, gpt.py:435-436)
 * Runs the whole prompt and the first sampling step on `stream`; does not synchronise. */
int ctb_gpt_begin(ctb_gpt* h, int32_t B, int32_t T0, const float* emb_dev, const uint8_t* mask_dev,
                  const ctb_sampler_config* sampler, const float* q_noise_dev, int32_t max_new_token,
                  int32_t infer_text, int32_t* ids_out_dev, float* hiddens_out_dev, void* stream);

/* Enqueue up to n_steps more iterations of the loop gpt.py:394-596.  Steps after every row
 * has finished (gpt.py:592) are no-ops on the device.  Does not synchronise. */
int ctb_gpt_decode(ctb_gpt* h, int32_t n_steps, void* stream);

typedef struct ctb_gpt_status {
  int32_t steps_done;     /* loop iterations executed so far (incl. the first one) */
  int32_t all_finished;   /* finish.all() (gpt.py:592) */
  int32_t any_finished_first_step; /* i == 0 and finish.any() (gpt.py:527) */
  int32_t reserved;
} ctb_gpt_status;

/* Synchronises `stream`, then reports loop state and copies per-row results:
 *   end_idx_host[B] int32 (gpt.py:345,576-577), finish_host[B] uint8; either may be NULL. */
int ctb_gpt_status_query(ctb_gpt* h, ctb_gpt_status* out, int32_t* end_idx_host, uint8_t* finish_host,
                         void* stream);

/* Measurement hook for bench.py's roofline: launches ONE kernel kind once per layer on the
 * state left by the last generate call (kind 0 qkv, 1 attention, 2 o-proj, 3 gate/up, 4 down;
 * 5 = heads, 6 = sampler, 7 = one decode step as ONE kernel launch (k_flow / k_step), 8 = 16 decode steps in one
 * k_flow launch with the sampling tail inside).  7 and 8 advance the generation.  No reference counterpart. */
int ctb_gpt_profile_kernel(ctb_gpt* h, int32_t kind, void* stream);

/* Profiling aid: with CTB_MEGA_TRACE=1 in the environment at ctb_gpt_create, the one-kernel decode step records
 * a %globaltimer stamp (ns) of CTA 0 after every grid barrier of the most recent step; copies up to n of them. */
int ctb_gpt_debug_trace(ctb_gpt* h, unsigned long long* host_out, int n);

/* Prompt embedding mix: replaces Embed.forward (ChatTTS/model/embed.py:51-79).
 *   ids_dev [B, T, num_vq] int64 (tokenizer output), text_mask_dev [B, T] uint8, tables inside the packed blob of `h`;
 *   out_dev [B, T, d] fp32: text positions get emb_text[ids[...,0]], the others sum_q emb_code[q][ids[...,q]]. */
int ctb_gpt_embed_prompt(ctb_gpt* h, const int64_t* ids_dev, const uint8_t* text_mask_dev, int32_t B, int32_t T,
                         float* out_dev, void* stream);

/* Stand-alone sampling tail over caller-provided logits (minimum slice of SURVEY.md 7.2;
 * same kernel the decode loop uses).
 *   logits_dev [rows, V] fp32 (not modified); gen_ids_dev [rows/rpi, gen_stride, rpi] int32 with
 *   n_gen tokens generated so far; out_idx_dev [rows] int32. */
int ctb_sample(const float* logits_dev, int32_t rows, int32_t V, int32_t rows_per_item,
               const ctb_sampler_config* sampler, const float* q_noise_dev, const int32_t* gen_ids_dev,
               int32_t gen_stride, int32_t n_gen, int32_t step, int32_t* out_idx_dev, void* stream);

/* ---- token -> waveform: replaces ChatTTS/core.py:512-539 (_decode_to_wavs) ------ */

typedef struct ctb_convstack_config {
  int32_t idim, odim, hidden, n_layer, bn_dim, kernel, dilation; /* dvae.py:131-172 */
  int32_t out_dim;    /* DVAE(dim=...) : out_conv input channels; 100 mel bins out (dvae.py:236) */
  int32_t vq_dim, vq_groups, vq_residual; /* GFSQ (dvae.py:69-97); vq_dim = 0: no VQ layer */
  int32_t vq_levels;  /* low byte: levels per dim (5); bits 8..15: residual scale base (0 => levels-1);
                       * bit 16 (encode only): 1 = do NOT bound() the projected input before the first residual stage */
} ctb_convstack_config;

typedef struct ctb_vocos_config {
  int32_t input_channels, dim, intermediate_dim, num_layers, n_fft, hop_length; /* config.py:74-121 */
} ctb_vocos_config;

typedef struct ctb_decoder ctb_decoder;

/* Element offsets inside the packed decoder blob (DVAE stack + out_conv + coef [+ VQ]). */
int64_t ctb_dvae_blob_floats(const ctb_convstack_config* cfg);
int64_t ctb_vocos_blob_floats(const ctb_vocos_config* cfg);

int ctb_decoder_create(const ctb_convstack_config* dvae_cfg, const float* dvae_blob_dev,
                       const ctb_vocos_config* vocos_cfg, const float* vocos_blob_dev, int32_t max_batch,
                       int32_t max_tokens, ctb_decoder** out);
int ctb_decoder_destroy(ctb_decoder* h);

/* DVAE.forward(mode="decode") (dvae.py:276-297).  in_layout selects what in_dev holds:
 *   0: hidden path, channels-first [B, C, T] fp32 (C = 2*idim) - the layout DVAE.__call__ receives
 *      from core.py:519-534;
 *   1: hidden path, token-major [B, T, C] fp32 - what ctb_gpt_* writes to hiddens_out_dev; the
 *      frame doubling of dvae.py:281-287 is then a pure re-interpretation (no copy);
 *   2: code path, ids [B, num_vq, T] int32 through GFSQ._embed (dvae.py:87-97).
 *   mel_dev [B, 100, 2T] fp32 channels-first, or NULL to keep the mel only inside the handle
 *   (time-major) for a following ctb_vocos_decode(mel_dev = NULL). */
int ctb_dvae_decode(ctb_decoder* h, const void* in_dev, int32_t in_layout, int32_t B, int32_t T, float* mel_dev,
                    void* stream);
/* Vocos.decode (core.py:505-510): mel [B,100,F] channels-first (NULL: the mel left in the handle by
 * the last ctb_dvae_decode) -> wav [B, hop*(F-1)] fp32 */
int ctb_vocos_decode(ctb_decoder* h, const float* mel_dev, int32_t B, int32_t F, float* wav_dev, void* stream);

/* ---- waveform -> codes: replaces DVAE.forward(mode="encode") (ChatTTS/model/dvae.py:265-274), i.e. ------
 * MelSpectrogramFeatures (dvae.py:175-206; n_fft 1024, hop 256, 100 mel bins, center/reflect, power 1, log(clip 1e-5))
 * -> / coef -> downsample_conv (dvae.py:231-236) -> encoder DVAEDecoder stack (dvae.py:131-172) -> GFSQ.forward indices
 * (dvae.py:102-128).  Caller: Chat.sample_audio_speaker (core.py:179-180) and the automatic speaker sample of
 * multi-sentence infer() (core.py:435-453).
 * cfg: idim = DVAE dim (512), odim = vq_dim (1024), hidden / n_layer / bn_dim / kernel / dilation of the encoder stack,
 * vq_* as for the decoder.  The blob holds the analysis window, the mel filterbank, coef, the two downsample convs,
 * the stack and the FSQ project_in matrices (order: chattts_b200/decoder.py::pack_dvae_encoder). */
typedef struct ctb_encoder ctb_encoder;
int64_t ctb_dvae_encoder_blob_floats(const ctb_convstack_config* enc_cfg);
int ctb_dvae_encoder_create(const ctb_convstack_config* enc_cfg, const float* blob_dev, int64_t max_samples,
                            ctb_encoder** out);
int ctb_dvae_encoder_destroy(ctb_encoder* h);
/* wav_dev [n_samples] fp32 (24 kHz) -> ids_dev [G*R, T] int32 with T = (n_samples / 256 + 1) / 2 written to the HOST
 * int *n_tokens_out; ids_capacity_tokens = tokens ids_dev (and margin_dev) can hold per code row.
 * mel_dev (optional) [100, n_samples / 256 + 1]: the log-mel BEFORE the division by coef is not kept; this is mel / coef.
 * margin_dev (optional) [G*R, T] fp32: distance of the closest pre-rounding value to a rounding edge (0 .. 0.5), the
 * decision margin the parity tests use to tell a real mismatch from fp32 reordering noise. */
int ctb_dvae_encode(ctb_encoder* h, const float* wav_dev, int64_t n_samples, int32_t* ids_dev,
                    int32_t ids_capacity_tokens, int32_t* n_tokens_out, float* mel_dev, float* margin_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHATTTS_B200_H */
