"""``Chat`` - the public API of the reference (ChatTTS/core.py) re-hosted on the B200 hot paths.

Same surface: ``Chat.load / infer / interrupt / unload / has_loaded / sample_random_speaker``,
``Chat.RefineTextParams`` / ``Chat.InferCodeParams`` (core.py:137-273).  The two hot paths are
ours (``GPT.generate`` -> chattts_b200.gpt, ``_decode_to_wavs`` -> chattts_b200.decoder); the
out-of-scope host components (text normaliser, BERT tokenizer, speaker strings, asset download -
SURVEY.md 2 rows 7, 8, 10, 11) are *injected*: ``load()`` takes them from an installed reference
``ChatTTS`` package, ``load_states()`` accepts any objects with the same methods (tests use stubs).
"""
from __future__ import annotations

import logging
import os
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from .config import Config
from .decoder import decode_to_wavs_window, DVAE, Vocos, decode_to_wavs
from .embed import Embed
from .gpt import GPT
from .norm import Normalizer
from .processors import gen_logits


class Chat:
    def __init__(self, logger=logging.getLogger(__name__)):
        self.logger = logger
        self.config = Config()
        self.normalizer = Normalizer(logger=logger)        # no homophone map until load() finds the asset (core.py:39-42)
        self.context = GPT.Context()

    # core.py:49-64
    def has_loaded(self, use_decoder=False):
        check = ["vocos", "gpt", "tokenizer", "embed", "decoder" if use_decoder else "dvae"]
        for module in check:
            if not hasattr(self, module):
                self.logger.warning(f"{module} not initialized.")
                return False
        return True

    # ------------------------------------------------------------------ loading
    def load(self, source="local", force_redownload=False, compile: bool = False, custom_path=None,
             device: Optional[torch.device] = None, coef: Optional[torch.Tensor] = None, use_flash_attn=False,
             use_vllm=False, experimental: bool = False, spk_stat: Optional[str] = None) -> bool:
        """core.py:137-163.  ``compile`` / ``use_flash_attn`` / ``use_vllm`` / ``experimental`` are accepted and
        ignored (one back end).  The asset files are located like the reference does for
        ``source="local"`` (working directory or ``custom_path``) and ``"custom"``; downloading / sha256 checking
        (core.py:66-135) is left to the reference package, which ``source="huggingface"`` therefore still needs.
        ``spk_stat`` (config.py:132, the base16384 std|mean table random speakers are drawn from) is taken from the
        argument, else from an importable reference package; without it only ``sample_random_speaker`` is unavailable."""
        root = self.download_models(source, force_redownload, custom_path)
        if root is None:
            return False
        from dataclasses import asdict

        paths = {k: os.path.join(root, v) for k, v in asdict(self.config.path).items()}
        from safetensors.torch import load_file
        from transformers import LlamaModel

        from .speaker import Speaker
        from .tokenizer import Tokenizer

        gpt_model = LlamaModel.from_pretrained(paths["gpt_ckpt_path"])
        states = {
            "gpt": {k: v for k, v in gpt_model.state_dict().items() if not k.startswith("embed_tokens")},
            "embed": load_file(paths["embed_path"]), "decoder": load_file(paths["decoder_ckpt_path"]),
            "dvae": load_file(paths["dvae_ckpt_path"]), "vocos": load_file(paths["vocos_ckpt_path"]),
        }
        homophones = None
        if spk_stat is None or homophones is None:
            try:
                import ChatTTS as ref  # optional: only its data constants are read

                spk_stat = spk_stat or ref.config.Config().spk_stat
                cand = os.path.join(os.path.dirname(ref.__file__), "res", "homophones_map.json")   # core.py:39-42
                homophones = cand if os.path.exists(cand) else None
            except Exception:
                pass
        dev = device or torch.device("cuda")
        self.normalizer = Normalizer(homophones, self.logger)
        return self.load_states(states, tokenizer=Tokenizer(paths["tokenizer_path"]),
                                speaker=Speaker(self.config.gpt.hidden_size, spk_stat, dev), device=dev, coef=coef)

    def download_models(self, source="local", force_redownload=False, custom_path=None) -> Optional[str]:
        """core.py:66-135, without the downloader: returns the folder that holds ``asset/`` or ``None``."""
        if source == "huggingface":
            try:
                import ChatTTS as ref
            except Exception as e:  # pragma: no cover - needs the reference + network
                raise RuntimeError('source="huggingface" downloads through the reference package, which is not '
                                   'installed; fetch the assets yourself and use source="custom"') from e
            return ref.Chat(self.logger).download_models(source, force_redownload, custom_path)
        root = custom_path if custom_path is not None else os.getcwd()
        from dataclasses import asdict

        missing = [v for v in asdict(self.config.path).values() if not os.path.exists(os.path.join(root, v))]
        if missing:
            self.logger.error("assets missing under %s: %s", root, ", ".join(missing))
            return None
        return str(root)

    def load_states(self, states: Dict[str, Dict[str, torch.Tensor]], tokenizer, speaker, device=None, coef=None,
                    max_batch: int = 32, max_context: int = 4096, weights_blob: Optional[torch.Tensor] = None) -> bool:
        """Build every model from in-memory state dicts (reference names) - core.py:275-384."""
        device = torch.device(device or "cuda")
        self.device = self.device_gpt = device
        cfg = self.config
        self.vocos = Vocos(cfg.vocos, device, max_batch=max_batch, max_tokens=max_context)
        self.vocos.state = {k: v.float() for k, v in states["vocos"].items()}
        self.dvae = DVAE(cfg.dvae.decoder, cfg.dvae.encoder, cfg.dvae.vq, dim=cfg.dvae.decoder.idim, coef=coef,
                         device=device, vocos=self.vocos, max_batch=max_batch, max_tokens=max_context)
        self.dvae.load_state_dict(states["dvae"])
        self.embed = Embed(cfg.embed.hidden_size, cfg.embed.num_audio_tokens, cfg.embed.num_text_tokens,
                           cfg.embed.num_vq).load_state_dict(states["embed"]).to(device)
        self.gpt = GPT(cfg.gpt, self.embed, device=device, device_gpt=device, logger=self.logger,
                       max_batch=max_batch, max_context=max_context)
        self.gpt.load_state(states.get("gpt"), weights_blob=weights_blob)
        self.speaker = speaker
        self.decoder = DVAE(cfg.decoder, dim=cfg.decoder.idim, coef=coef, device=device, vocos=self.vocos,
                            max_batch=max_batch, max_tokens=max_context)
        self.decoder.load_state_dict(states["decoder"])
        self.tokenizer = tokenizer
        self.coef = coef
        return self.has_loaded()

    def unload(self):
        logger = self.logger
        for module in ["vocos", "gpt", "decoder", "dvae", "tokenizer", "embed", "speaker"]:
            if hasattr(self, module):
                delattr(self, module)
        self.__init__(logger)

    def sample_random_speaker(self) -> str:
        return self.speaker.sample_random()

    def sample_audio_speaker(self, wav) -> str:
        """core.py:179-180: wav (24 kHz, 1-D) -> DVAE codes -> speaker-prompt string for ``InferCodeParams.spk_smp``."""
        if self.dvae.audio_encoder is None:
            raise RuntimeError("this DVAE checkpoint carries no encoder / VQ weights: cannot sample a speaker from audio")
        return self.speaker.encode_prompt(self.dvae.sample_audio(wav))

    # ------------------------------------------------------------------ params (core.py:182-206)
    @dataclass(repr=False, eq=False)
    class RefineTextParams:
        prompt: str = ""
        top_P: float = 0.7
        top_K: int = 20
        temperature: float = 0.7
        repetition_penalty: float = 1.0
        max_new_token: int = 384
        min_new_token: int = 0
        show_tqdm: bool = True
        ensure_non_empty: bool = True
        manual_seed: Optional[int] = None

    @dataclass(repr=False, eq=False)
    class InferCodeParams(RefineTextParams):
        prompt: str = "[speed_5]"
        spk_emb: Optional[str] = None
        spk_smp: Optional[str] = None
        txt_smp: Optional[str] = None
        temperature: float = 0.3
        repetition_penalty: float = 1.05
        max_new_token: int = 2048
        stream_batch: int = 24
        stream_speed: int = 12000
        pass_first_n_batches: int = 2

    # ------------------------------------------------------------------ infer (core.py:208-270)
    def infer(self, text, stream=False, lang=None, skip_refine_text=False, refine_text_only=False, use_decoder=True,
              do_text_normalization=True, do_homophone_replacement=True, split_text=True, max_split_batch=4,
              params_refine_text=None, params_infer_code=None):
        params_refine_text = params_refine_text or Chat.RefineTextParams()
        params_infer_code = params_infer_code or Chat.InferCodeParams()
        self.context.set(False)
        if split_text and isinstance(text, str):
            if "\n" in text:
                text = text.split("\n")
            else:
                text = [t for t in re.split(r"(?<=。)|(?<=\.\s)", text) if t]
            self.logger.info("split text into %d parts", len(text))
        if len(text) == 0:
            return []
        res_gen = self._infer(text, stream, lang, skip_refine_text, refine_text_only, use_decoder,
                              do_text_normalization, do_homophone_replacement, split_text, max_split_batch,
                              params_refine_text, params_infer_code)
        if stream:
            return res_gen
        if not refine_text_only:
            stripped = []
            thr = np.float32(1e-5)
            for wavs in res_gen:
                for wav in wavs:
                    stripped.append(wav[np.abs(wav) > thr])  # quirk Q20
            if split_text:
                return [np.concatenate(stripped)]
            return stripped
        return next(res_gen)

    def interrupt(self):
        self.context.set(True)

    # core.py:386-503
    def _infer(self, text, stream, lang, skip_refine_text, refine_text_only, use_decoder, do_text_normalization,
               do_homophone_replacement, split_text, max_split_batch, params_refine_text, params_infer_code):
        assert self.has_loaded(use_decoder=use_decoder)
        if not isinstance(text, list):
            text = [text]
        text = [self.normalizer(t, do_text_normalization, do_homophone_replacement, lang) for t in text]
        if not skip_refine_text:
            tokens = []
            for lo in range(0, len(text), self.gpt.max_batch):          # one batch in the reference; chunks beyond max_batch
                refined = self._refine_text(text[lo: lo + self.gpt.max_batch], self.device, params_refine_text)
                tokens += [i[i.less(self.tokenizer.break_0_ids)] for i in refined.ids]
                refined.destroy()
            text = self.tokenizer.decode(tokens)
            if refine_text_only:
                if split_text and isinstance(text, list):
                    text = "\n".join(text)
                yield text
                return
        if split_text and len(text) > 1 and params_infer_code.spk_smp is None:
            # core.py:435-453: sentence 0 is synthesised once on its own and its audio, re-encoded by the DVAE encode
            # branch, becomes the speaker prompt (spk_smp / txt_smp) of every sentence - this keeps one voice across them
            refer_text = text[0]
            result = next(self._infer_code(refer_text, False, self.device, use_decoder, params_infer_code))
            wavs = self._decode_to_wavs(result.hiddens if use_decoder else result.ids, use_decoder)
            result.destroy()
            assert len(wavs) == 1
            params_infer_code.spk_smp = self.sample_audio_speaker(wavs[0])
            params_infer_code.txt_smp = refer_text
        if stream:
            length, pass_batch_count = 0, 0
        if split_text:
            n = (len(text) + max_split_batch - 1) // max_split_batch
        else:
            # the reference runs all texts as one batch; batches beyond the handle's max_batch run as consecutive chunks
            # (rows are independent, so the result per text is the same; noise rows restart per chunk)
            max_split_batch = min(len(text), self.gpt.max_batch)
            n = (len(text) + max_split_batch - 1) // max_split_batch
        for i in range(n):
            chunk = text[i * max_split_batch: (i + 1) * max_split_batch]
            for result in self._infer_code(chunk, stream, self.device, use_decoder, params_infer_code):
                res = result.hiddens if use_decoder else result.ids
                if stream:
                    # core.py:455-503 decodes the CUMULATIVE sequence at every yield and slices [length, length +
                    # stream_speed) out of it; the same samples are produced here from the token window they depend on
                    # (decoder.decode_to_wavs_window)
                    pass_batch_count += 1
                    last = [r.clone() for r in res]
                    total = 512 * max(int(r.size(0)) for r in res) - 256
                    result.destroy()
                    if pass_batch_count <= params_infer_code.pass_first_n_batches:
                        continue
                    a, b = length, min(length + params_infer_code.stream_speed, total)
                    length = b
                    yield self._decode_window(last, use_decoder, a, b)
                else:
                    wavs = self._decode_to_wavs(res, use_decoder)
                    result.destroy()
                    yield wavs
            if stream:
                total = 512 * max(int(r.size(0)) for r in last) - 256
                new_wavs = self._decode_window(last, use_decoder, length, total)
                keep = np.sum(np.abs(new_wavs) > 1e-5, axis=0) > 0
                yield new_wavs[:][:, keep]

    # core.py:512-539 - hot path 2
    @torch.inference_mode()
    def _decode_to_wavs(self, result_list: List[torch.Tensor], use_decoder: bool):
        return decode_to_wavs(result_list, use_decoder, self.decoder, self.dvae)

    @torch.inference_mode()
    def _decode_window(self, result_list: List[torch.Tensor], use_decoder: bool, a: int, b: int) -> np.ndarray:
        """Samples [a, b) of ``_decode_to_wavs(result_list)`` from the tokens they depend on (streaming hand-off)."""
        return decode_to_wavs_window(result_list, use_decoder, self.decoder, self.dvae, a, b)

    def _vocos_decode(self, spec: torch.Tensor) -> np.ndarray:
        return self.vocos_engine().vocos_decode(spec).cpu().numpy()

    def vocos_engine(self):
        return self.decoder.engine

    # core.py:541-662 - hot path 1 (audio codes)
    @torch.no_grad()
    def _infer_code(self, text, stream: bool, device, return_hidden: bool, params):
        if not isinstance(text, list):
            text = [text]
        assert len(text), "text should not be empty"
        temperature = params.temperature if isinstance(params.temperature, list) else [params.temperature] * self.config.gpt.num_vq
        input_ids, attention_mask, text_mask = self.tokenizer.encode(
            self.speaker.decorate_code_prompts(text, params.prompt, params.txt_smp, params.spk_emb),
            self.config.gpt.num_vq,
            prompt=(self.speaker.decode_prompt(params.spk_smp) if params.spk_smp is not None else None),
            device=self.device_gpt)
        num_code = self.config.gpt.num_audio_tokens - 1
        warpers, processors = gen_logits(num_code=num_code, top_P=params.top_P, top_K=params.top_K,
                                         repetition_penalty=params.repetition_penalty)
        emb = self.embed(input_ids, text_mask)
        if params.spk_emb is not None:
            self.speaker.apply(emb, params.spk_emb, input_ids, self.tokenizer.spk_emb_ids, self.gpt.device_gpt)
        return self.gpt.generate(
            emb, input_ids, temperature=torch.tensor(temperature), eos_token=num_code, attention_mask=attention_mask,
            max_new_token=params.max_new_token, min_new_token=params.min_new_token,
            logits_processors=(*processors, *warpers), infer_text=False, return_hidden=return_hidden, stream=stream,
            show_tqdm=params.show_tqdm, ensure_non_empty=params.ensure_non_empty, stream_batch=params.stream_batch,
            manual_seed=params.manual_seed, context=self.context)

    # core.py:664-751 - hot path 1 (text refinement)
    @torch.no_grad()
    def _refine_text(self, text, device, params):
        if not isinstance(text, list):
            text = [text]
        input_ids, attention_mask, text_mask = self.tokenizer.encode(
            self.speaker.decorate_text_prompts(text, params.prompt), self.config.gpt.num_vq, device=self.device_gpt)
        warpers, processors = gen_logits(num_code=self.tokenizer.len, top_P=params.top_P, top_K=params.top_K,
                                         repetition_penalty=params.repetition_penalty)
        emb = self.embed(input_ids, text_mask)
        return next(self.gpt.generate(
            emb, input_ids, temperature=torch.tensor([params.temperature]), eos_token=self.tokenizer.eos_token,
            attention_mask=attention_mask, max_new_token=params.max_new_token, min_new_token=params.min_new_token,
            logits_processors=(*processors, *warpers), infer_text=True, stream=False, show_tqdm=params.show_tqdm,
            ensure_non_empty=params.ensure_non_empty, manual_seed=params.manual_seed, context=self.context))
