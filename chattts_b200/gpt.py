"""Host side of hot path 1: drop-in for the reference's ``GPT`` (ChatTTS/model/gpt.py).

Same constructor, ``generate`` signature, ``GenerationOutputs`` and ``Context`` as the
reference (gpt.py:21-57,103-111,276-337); the per-token Python loop (gpt.py:394-596) is
replaced by ``ctb_gpt_begin`` / ``ctb_gpt_decode`` (include/chattts_b200.h) which run the
whole step - embedding sum, 20 decoder layers with paged KV, the four heads, the sampling
filters, multinomial and finish bookkeeping - as CUDA kernels replayed from a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch

from . import _lib
from .config import GPTConfig
from .embed import Embed
from .processors import build_sampler_config, exp_noise


def _rope_tables(max_pos: int, head_dim: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin exactly as HF LlamaRotaryEmbedding computes them on the CPU in fp32
    ([3p]; in-tree statement examples/onnx/modeling_llama.py:119-162).  Built on the host so
    the kernel reads bit-identical values to the fp32 CPU reference."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = (inv_freq[None, :, None] @ pos[None, None, :]).transpose(1, 2)[0]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().contiguous(), emb.sin().contiguous()


def _del_all(x):
    if isinstance(x, list):
        x.clear()


class GPT:
    class Context:
        """gpt.py:103-111 - interrupt flag polled between decode chunks."""

        def __init__(self):
            self._interrupt = False

        def set(self, v: bool):
            self._interrupt = v

        def get(self) -> bool:
            return self._interrupt

    @dataclass(repr=False, eq=False)
    class GenerationOutputs:
        """gpt.py:276-285."""

        ids: List[torch.Tensor]
        attentions: List[Optional[Tuple[torch.FloatTensor, ...]]]
        hiddens: List[torch.Tensor]

        def destroy(self):
            _del_all(self.ids)
            _del_all(self.attentions)
            _del_all(self.hiddens)

    def __init__(self, gpt_config: Union[dict, GPTConfig], embed: Embed, use_flash_attn=False, use_vllm=False,
                 device=torch.device("cuda"), device_gpt=torch.device("cuda"),
                 logger=logging.getLogger(__name__), max_batch: int = 32, max_context: int = 2560):
        self.logger = logger
        self.device = torch.device(device)
        self.device_gpt = torch.device(device_gpt)
        if isinstance(gpt_config, dict):
            known = {f for f in GPTConfig.__dataclass_fields__}
            gpt_config = GPTConfig(**{k: v for k, v in gpt_config.items() if k in known})
        self.config = gpt_config
        self.num_vq = int(gpt_config.num_vq)
        self.num_audio_tokens = int(gpt_config.num_audio_tokens)
        self.num_text_tokens = int(gpt_config.num_text_tokens)
        # accepted for signature compatibility, ignored: there is one back end
        self.use_flash_attn, self.is_vllm, self.is_te_llama = use_flash_attn, False, False
        self.embed = embed
        self.max_batch, self.max_context = max_batch, max_context
        self._handle = C.c_void_p()
        self._weights: Optional[torch.Tensor] = None
        self._stream_keepalive = []

    # ------------------------------------------------------------------ loading
    def load_pretrained(self, gpt_folder: str, embed_file_path: str, experimental=False):
        """gpt.py:59-101: HF folder -> state dict (the reference's loader), then pack."""
        from transformers import LlamaModel

        model = LlamaModel.from_pretrained(gpt_folder)
        cfg = model.config  # quirk Q22: hyper-parameters come from asset/gpt/config.json
        rope = getattr(cfg, "rope_parameters", None) or {}
        self.config = GPTConfig(
            hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
            num_attention_heads=cfg.num_attention_heads,
            num_key_value_heads=getattr(cfg, "num_key_value_heads", cfg.num_attention_heads),
            head_dim=getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads,
            num_hidden_layers=cfg.num_hidden_layers, max_position_embeddings=cfg.max_position_embeddings,
            rms_norm_eps=cfg.rms_norm_eps, rope_theta=rope.get("rope_theta", getattr(cfg, "rope_theta", 10000.0)),
            num_audio_tokens=self.num_audio_tokens, num_text_tokens=self.num_text_tokens, num_vq=self.num_vq)
        state = {k: v for k, v in model.state_dict().items() if not k.startswith("embed_tokens")}
        self.load_state(state)

    def load_state(self, gpt_state: Dict[str, torch.Tensor], weights_blob: Optional[torch.Tensor] = None):
        """Pack the checkpoint into the blob layout of ``ctb_gpt_layout_query`` and create the handle.

        ``weights_blob``: an already packed device tensor (e.g. received by NCCL broadcast,
        chattts_b200/dist.py) - then ``gpt_state`` may be None."""
        _lib.require_cuda()
        lib = _lib.load()
        cc, lay = self.query_layout()
        if weights_blob is None:
            weights_blob = self.pack_weights(gpt_state, lay).to(self.device_gpt)
        assert weights_blob.numel() == lay.total and weights_blob.dtype == torch.float32
        self._weights = weights_blob.contiguous()
        if self._handle:
            lib.ctb_gpt_destroy(self._handle)
            self._handle = C.c_void_p()
        with torch.cuda.device(self.device_gpt):
            _lib.check(lib.ctb_gpt_create(C.byref(cc), C.c_void_p(self._weights.data_ptr()), C.byref(self._handle)))
        self.embed._gpt = self  # Embed.forward now runs as ctb_gpt_embed_prompt on this handle's tables

    @torch.inference_mode()
    def embed_prompt(self, input_ids: torch.Tensor, text_mask: torch.Tensor) -> torch.Tensor:
        """Embed.forward (embed.py:51-79) on the device: [B, T, num_vq] ids + text mask -> [B, T, d] fp32."""
        lib = _lib.load()
        dev = self.device_gpt
        ids = input_ids.to(dev, torch.int64).contiguous()
        tm = text_mask.to(dev).to(torch.uint8).contiguous()
        B, T = int(ids.shape[0]), int(ids.shape[1])
        out = torch.empty(B, T, self.config.hidden_size, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ctb_gpt_embed_prompt(self._handle, C.c_void_p(ids.data_ptr()), C.c_void_p(tm.data_ptr()), B, T,
                                                C.c_void_p(out.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def query_layout(self):
        """(ctb_gpt_config, ctb_gpt_layout) for this model shape - the C side owns the blob layout."""
        lib = _lib.load()
        c = self.config
        cc = _lib.GptConfig(c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.num_attention_heads,
                            c.num_key_value_heads, c.head_dim, c.num_vq, c.num_audio_tokens, c.num_text_tokens,
                            c.max_position_embeddings, c.rms_norm_eps, self.max_batch, self.max_context)
        lay = _lib.GptLayout()
        _lib.check(lib.ctb_gpt_layout_query(C.byref(cc), C.byref(lay)))
        self._cc, self._layout = cc, lay
        return cc, lay

    def pack_weights(self, s: Dict[str, torch.Tensor], lay) -> torch.Tensor:
        c = self.config
        blob = torch.empty(lay.total, dtype=torch.float32)

        def put(off, t):
            t = t.detach().to("cpu", torch.float32).contiguous().view(-1)
            blob[off: off + t.numel()] = t

        for l in range(c.num_hidden_layers):
            base, p = lay.layer0 + l * lay.layer_stride, f"layers.{l}."
            put(base + lay.wqkv, torch.cat([s[p + "self_attn.q_proj.weight"], s[p + "self_attn.k_proj.weight"],
                                            s[p + "self_attn.v_proj.weight"]], 0))
            put(base + lay.wo, s[p + "self_attn.o_proj.weight"])
            put(base + lay.wgate_up, torch.cat([s[p + "mlp.gate_proj.weight"], s[p + "mlp.up_proj.weight"]], 0))
            put(base + lay.wdown, s[p + "mlp.down_proj.weight"])
            put(base + lay.ln1, s[p + "input_layernorm.weight"])
            put(base + lay.ln2, s[p + "post_attention_layernorm.weight"])
        put(lay.final_norm, s["norm.weight"])
        e = self.embed
        put(lay.head_code, torch.cat([e.folded_head(f"head_code.{q}").cpu() for q in range(c.num_vq)], 0))
        put(lay.head_text, e.folded_head("head_text").cpu())
        put(lay.emb_code, torch.cat([e.state[f"emb_code.{q}.weight"].cpu() for q in range(c.num_vq)], 0))
        put(lay.emb_text, e.state["emb_text.weight"].cpu())
        cos, sin = _rope_tables(c.max_position_embeddings, c.head_dim, c.rope_theta)
        put(lay.rope_cos, cos)
        put(lay.rope_sin, sin)
        return blob

    def prepare(self, compile=False):
        """gpt.py:131-139 compiled the HF model with inductor; nothing to do here (quirk Q14)."""

    def eval(self):
        return self

    def __del__(self):
        try:
            if self._handle:
                _lib.load().ctb_gpt_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ outputs
    @torch.no_grad()
    def _prepare_generation_outputs(self, inputs_ids: torch.Tensor, start_idx: int, end_idx: torch.Tensor,
                                    attentions, hiddens, infer_text: bool) -> "GPT.GenerationOutputs":
        """gpt.py:287-313 (``hiddens`` here is already a [B, n, d] tensor or an empty list)."""
        ids = [inputs_ids[i].narrow(0, start_idx, int(n)) for i, n in enumerate(end_idx)]
        if infer_text:
            ids = [i.narrow(1, 0, 1).squeeze_(1) for i in ids]
        if isinstance(hiddens, list) and len(hiddens) > 0:
            hiddens = torch.stack(hiddens, 1)
        if isinstance(hiddens, torch.Tensor):
            hiddens = [hiddens[i].narrow(0, 0, int(n)) for i, n in enumerate(end_idx.int())]
        return self.GenerationOutputs(ids=ids, attentions=attentions, hiddens=hiddens)

    # ------------------------------------------------------------------ device-resident entry
    def enqueue_generate(self, emb_d: torch.Tensor, mask_d: torch.Tensor, cfg, q_d: Optional[torch.Tensor],
                         max_new_token: int, infer_text: bool, ids_out: torch.Tensor,
                         hid_out: Optional[torch.Tensor], n_steps: Optional[int] = None) -> None:
        """Enqueue prefill + ``n_steps`` loop iterations on the current stream with every buffer
        already resident on the device; never synchronises (bench.py times this with CUDA events)."""
        lib = _lib.load()
        B, T0 = int(emb_d.shape[0]), int(emb_d.shape[1])
        stream_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.ctb_gpt_begin(
            self._handle, B, T0, C.c_void_p(emb_d.data_ptr()), C.c_void_p(mask_d.data_ptr()), C.byref(cfg),
            C.c_void_p(q_d.data_ptr()) if q_d is not None else None, max_new_token, int(bool(infer_text)),
            C.c_void_p(ids_out.data_ptr()), C.c_void_p(hid_out.data_ptr()) if hid_out is not None else None,
            stream_ptr))
        n = max_new_token - 1 if n_steps is None else n_steps
        if n > 0:
            _lib.check(lib.ctb_gpt_decode(self._handle, n, stream_ptr))

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def generate(self, emb: torch.Tensor, inputs_ids: torch.Tensor, temperature: torch.Tensor,
                 eos_token: Union[int, torch.Tensor], attention_mask: Optional[torch.Tensor] = None,
                 max_new_token=2048, min_new_token=0,
                 logits_processors: Tuple[Callable[[torch.LongTensor, torch.FloatTensor], torch.FloatTensor]] = (),
                 infer_text=False, return_attn=False, return_hidden=False, stream=False, show_tqdm=True,
                 ensure_non_empty=True, stream_batch=24, manual_seed: Optional[int] = None,
                 context=Context()):
        """Generator with the reference's contract (gpt.py:315-618)."""
        if return_attn:
            raise NotImplementedError("return_attn: attention maps never leave the fused attention kernel")
        if not self._handle:
            raise _lib.CtbError("GPT weights not loaded")
        lib = _lib.load()
        dev = self.device_gpt
        B, T0 = int(inputs_ids.shape[0]), int(inputs_ids.shape[1])
        eos = int(eos_token)
        if B > self.max_batch or T0 + max_new_token > self.max_context:
            raise ValueError(f"batch {B} / context {T0}+{max_new_token} exceed this handle "
                             f"(max_batch={self.max_batch}, max_context={self.max_context})")
        rows_per_item = 1 if infer_text else self.num_vq
        V = self.num_text_tokens if infer_text else self.num_audio_tokens
        temps = [float(t) for t in torch.as_tensor(temperature).flatten().tolist()]
        seed = manual_seed
        philox = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else 0
        cfg = build_sampler_config(logits_processors, temps, eos, min_new_token, philox)
        if cfg.penalty_on and cfg.penalty_max_ids < B * rows_per_item:
            pass  # rows >= max_input_ids silently lose the penalty, like processors.py:24-27

        with torch.cuda.device(dev):
            stream_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            emb_d = emb.to(dev, torch.float32).contiguous()
            if attention_mask is None:
                attention_mask = torch.ones(B, T0, dtype=torch.bool)
            mask_d = attention_mask.to(dev).to(torch.uint8).contiguous()
            # left padding (tokenizer.py:79-110): every row's valid tokens are a contiguous suffix ending in the last column
            if not bool(mask_d[:, -1].all()) or (T0 > 1 and bool((mask_d[:, 1:] < mask_d[:, :-1]).any())):
                raise ValueError("attention_mask must be left padded: each row 0...0 1...1 with the last column valid")
            q_d = None
            if seed is not None:
                q_d = exp_noise(B * rows_per_item, V, seed).to(dev, non_blocking=True)
            ids_out = torch.zeros(B, max_new_token, self.num_vq, dtype=torch.int32, device=dev)
            hid_out = (torch.zeros(B, max_new_token, self.config.hidden_size, dtype=torch.float32, device=dev)
                       if return_hidden else None)
            end_idx = torch.zeros(B, dtype=torch.int32)
            finish = torch.zeros(B, dtype=torch.uint8)
            st = _lib.GptStatus()

            def query():
                _lib.check(lib.ctb_gpt_status_query(self._handle, C.byref(st), C.c_void_p(end_idx.data_ptr()),
                                                    C.c_void_p(finish.data_ptr()), stream_ptr))

            def outputs():
                ids64 = ids_out.to(torch.int64)
                if inputs_ids.device != ids64.device:
                    ids64 = ids64.to(inputs_ids.device)
                return self._prepare_generation_outputs(ids64, 0, end_idx.clone().long(), [],
                                                        hid_out if return_hidden else [], infer_text)

            pbar = None
            if show_tqdm:
                from tqdm import tqdm

                pbar = tqdm(total=max_new_token, desc="text" if infer_text else "code",
                            bar_format="{l_bar}{bar}| {n_fmt}/{total_fmt}(max) [{elapsed}, {rate_fmt}{postfix}]")

            _lib.check(lib.ctb_gpt_begin(
                self._handle, B, T0, C.c_void_p(emb_d.data_ptr()), C.c_void_p(mask_d.data_ptr()), C.byref(cfg),
                C.c_void_p(q_d.data_ptr()) if q_d is not None else None, max_new_token, int(bool(infer_text)),
                C.c_void_p(ids_out.data_ptr()), C.c_void_p(hid_out.data_ptr()) if hid_out is not None else None,
                stream_ptr))
            query()
            if st.any_finished_first_step:
                # gpt.py:527-570
                self.logger.warning("unexpected end at index %s", str([i for i in range(B) if finish[i]]))
                if ensure_non_empty and manual_seed is None:
                    if pbar is not None:
                        pbar.close()
                    self.logger.warning("regenerate in order to ensure non-empty")
                    yield from self.generate(emb, inputs_ids, temperature, eos_token, attention_mask, max_new_token,
                                             min_new_token, logits_processors, infer_text, return_attn,
                                             return_hidden, stream, show_tqdm, ensure_non_empty, stream_batch,
                                             manual_seed, context)
                return

            steps = 1
            chunk = int(stream_batch) if stream else int(os.environ.get("CTB_DECODE_CHUNK", "32"))
            interrupted = False
            while not st.all_finished and steps < max_new_token:
                if context.get():
                    interrupted = True
                    break
                n = min(chunk - (steps % chunk) if stream else chunk, max_new_token - steps)
                _lib.check(lib.ctb_gpt_decode(self._handle, n, stream_ptr))
                query()
                done = st.steps_done
                if done <= steps and not st.all_finished:
                    raise _lib.CtbError(f"decode made no progress (steps_done={done}); device loop state is corrupt")
                if pbar is not None:
                    pbar.update(done - steps)
                steps = done
                # gpt.py:578-589: cumulative yield every `stream_batch` unfinished steps
                if stream and not st.all_finished and steps % stream_batch == 0:
                    yield outputs()
            if pbar is not None:
                pbar.close()
            if stream and st.all_finished and steps - 1 > 0 and (steps - 1) % stream_batch == 0:
                # gpt.py:578-589 quirk: the finishing step does not advance stream_iter, so a boundary
                # reached on the previous step is yielded a second time before the final yield
                yield outputs()
            if not st.all_finished:
                if interrupted or context.get():
                    self.logger.warning("generation is interrupted")
                else:
                    self.logger.warning(f"incomplete result. hit max_new_token: {max_new_token}")
            yield outputs()
