"""Build the *reference's own* modules (imported from /root/reference) on synthetic weights.

TEST INFRASTRUCTURE, build-container only (see oracle/ref_import.py).
"""
from __future__ import annotations

import dataclasses
import logging

import torch

from .ref_import import load_reference


def build_reference_gpt(gpt_state, embed_state):
    """Mirrors core.py:336-357 / gpt.py:75-78 without on-disk assets."""
    load_reference()
    from ChatTTS.config import Config
    from ChatTTS.model import GPT, Embed
    from transformers import LlamaModel

    cfg = Config()
    embed = Embed(cfg.embed.hidden_size, cfg.embed.num_audio_tokens, cfg.embed.num_text_tokens, cfg.embed.num_vq).eval()
    embed.load_state_dict(embed_state)
    logger = logging.getLogger("ref-gpt")
    logger.setLevel(logging.ERROR)
    gpt = GPT(dataclasses.asdict(cfg.gpt), embed, logger=logger).eval()
    model = LlamaModel(gpt.llama_config).eval()
    del model.embed_tokens
    missing = model.load_state_dict(gpt_state, strict=False)
    assert not missing.unexpected_keys and all("rotary" in k for k in missing.missing_keys), missing
    gpt.gpt = model
    return gpt, embed


def reference_generate(gpt, embed, input_ids, attention_mask, text_mask, *, temperature, eos_token, max_new_token,
                       min_new_token=0, top_P=0.7, top_K=20, repetition_penalty=1.05, num_code=625,
                       infer_text=False, return_hidden=True, manual_seed=1234, extra_processors=()):
    """Mirrors core.py:582-658 (audio) / core.py:682-747 (text)."""
    from ChatTTS.model import gen_logits

    warp, proc = gen_logits(num_code=num_code, top_P=top_P, top_K=top_K, repetition_penalty=repetition_penalty)
    emb = embed(input_ids, text_mask)
    gen = gpt.generate(emb, input_ids, temperature=torch.tensor(temperature), eos_token=eos_token,
                       attention_mask=attention_mask, max_new_token=max_new_token, min_new_token=min_new_token,
                       logits_processors=(*proc, *warp, *extra_processors), infer_text=infer_text,
                       return_hidden=return_hidden, show_tqdm=False, manual_seed=manual_seed)
    return next(gen, None)


def build_reference_dvae(state, stack_cfg, dim):
    """Reference ``DVAE`` decode branch without the VQ layer (core.py:366-376 'decoder')."""
    load_reference()
    from ChatTTS.model import DVAE

    m = DVAE(decoder_config=dict(idim=stack_cfg.idim, odim=stack_cfg.odim, hidden=stack_cfg.hidden,
                                 n_layer=stack_cfg.n_layer, bn_dim=stack_cfg.bn_dim), dim=dim).eval()
    own = {k: v for k, v in state.items() if not k.startswith("vq_layer")}
    m.load_state_dict(own)
    return m


def build_reference_dvae_encoder(state, dec_cfg, enc_cfg, dim):
    """Reference ``DVAE`` with the encode-side modules (dvae.py:229-236) but no ``vq_layer`` (vector_quantize_pytorch is
    absent): exposes ``preprocessor_mel``, ``downsample_conv`` and ``encoder`` for piecewise pinning."""
    load_reference()
    from ChatTTS.model import DVAE

    d = lambda c: dict(idim=c.idim, odim=c.odim, hidden=c.hidden, n_layer=c.n_layer, bn_dim=c.bn_dim)
    m = DVAE(decoder_config=d(dec_cfg), encoder_config=d(enc_cfg), dim=dim).eval()
    missing = m.load_state_dict({k: v for k, v in state.items() if not k.startswith("vq_layer")}, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("preprocessor_mel.") for k in missing.missing_keys), missing
    return m
