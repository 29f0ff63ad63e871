"""Seeded synthetic weights in the reference's checkpoint *naming*.

There are no model assets offline, so tests, ``smoke()`` and ``bench.py``
run on random weights whose state-dict keys are exactly the ones the reference loader
reads: HF ``LlamaModel`` names for ``asset/gpt``, ``Embed.safetensors``
names (weight-norm ``original0/1``), ``Decoder.safetensors`` / ``DVAE.safetensors`` names
and the Vocos names.  Values come from a private ``torch.Generator`` so they are
independent of global RNG state and identical on every box.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import Config, ConvStackConfig, GPTConfig, VocosConfig, VQConfig

State = Dict[str, torch.Tensor]


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _normal(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def synth_gpt_state(seed: int = 0, std: float = 0.02, cfg: GPTConfig = GPTConfig()) -> State:
    """``LlamaModel.state_dict()`` minus ``embed_tokens`` (reference gpt.py:75-78)."""
    g = _gen(seed)
    d, i = cfg.hidden_size, cfg.intermediate_size
    hq = cfg.num_attention_heads * cfg.head_dim
    hkv = cfg.num_key_value_heads * cfg.head_dim
    s: State = {}
    for l in range(cfg.num_hidden_layers):
        p = f"layers.{l}."
        s[p + "self_attn.q_proj.weight"] = _normal(g, (hq, d), std)
        s[p + "self_attn.k_proj.weight"] = _normal(g, (hkv, d), std)
        s[p + "self_attn.v_proj.weight"] = _normal(g, (hkv, d), std)
        s[p + "self_attn.o_proj.weight"] = _normal(g, (d, hq), std)
        s[p + "mlp.gate_proj.weight"] = _normal(g, (i, d), std)
        s[p + "mlp.up_proj.weight"] = _normal(g, (i, d), std)
        s[p + "mlp.down_proj.weight"] = _normal(g, (d, i), std)
        s[p + "input_layernorm.weight"] = 1.0 + _normal(g, (d,), 0.05)
        s[p + "post_attention_layernorm.weight"] = 1.0 + _normal(g, (d,), 0.05)
    s["norm.weight"] = 1.0 + _normal(g, (d,), 0.05)
    return s


def synth_embed_state(seed: int = 1, cfg: GPTConfig = GPTConfig()) -> State:
    """``Embed.state_dict()`` (reference embed.py:18-35)."""
    g = _gen(seed)
    d = cfg.hidden_size
    s: State = {}
    for q in range(cfg.num_vq):
        s[f"emb_code.{q}.weight"] = _normal(g, (cfg.num_audio_tokens, d), 1.0)
    s["emb_text.weight"] = _normal(g, (cfg.num_text_tokens, d), 1.0)
    bound = d ** -0.5

    def head(prefix, rows):
        v = _uniform(g, (rows, d), bound)
        # weight_norm: original0 = g (row norm magnitude), original1 = v (direction)
        s[prefix + ".parametrizations.weight.original0"] = v.norm(dim=1, keepdim=True) * (
            1.0 + _uniform(g, (rows, 1), 0.25)
        )
        s[prefix + ".parametrizations.weight.original1"] = v

    head("head_text", cfg.num_text_tokens)
    for q in range(cfg.num_vq):
        head(f"head_code.{q}", cfg.num_audio_tokens)
    return s


def _conv(g, s, name, cout, cin_per_group, k, bias=True):
    bound = (cin_per_group * k) ** -0.5
    s[name + ".weight"] = _uniform(g, (cout, cin_per_group, k), bound)
    if bias:
        s[name + ".bias"] = _uniform(g, (cout,), bound)


def _linear(g, s, name, cout, cin, bias=True):
    bound = cin ** -0.5
    s[name + ".weight"] = _uniform(g, (cout, cin), bound)
    if bias:
        s[name + ".bias"] = _uniform(g, (cout,), bound)


def _convnext(g, s, p, dim, inter, scale_name, scale_val):
    _conv(g, s, p + "dwconv", dim, 1, 7)
    s[p + "norm.weight"] = 1.0 + _normal(g, (dim,), 0.05)
    s[p + "norm.bias"] = _normal(g, (dim,), 0.05)
    _linear(g, s, p + "pwconv1", inter, dim)
    _linear(g, s, p + "pwconv2", dim, inter)
    s[p + scale_name] = scale_val * (1.0 + _normal(g, (dim,), 0.1))


def synth_dvae_state(seed: int, stack: ConvStackConfig, dim: int, vq: VQConfig | None = None,
                     encoder: ConvStackConfig | None = None) -> State:
    """``DVAE.state_dict()`` decode-side keys (reference dvae.py:131-172,209-243).

    ``stack`` = decoder stack config, ``dim`` = ``DVAE(dim=...)`` (out_conv input channels).
    Layer-scale is drawn around 0.25 (the reference initialises 1e-6, trained values are
    O(0.1..1)); a non-trivial value keeps every block numerically visible in parity tests.
    """
    g = _gen(seed)
    s: State = {"coef": 0.5 + torch.rand((1, 100, 1), generator=g, dtype=torch.float32)}
    _conv(g, s, "decoder.conv_in.0", stack.bn_dim, stack.idim, 3)
    _conv(g, s, "decoder.conv_in.2", stack.hidden, stack.bn_dim, 3)
    for i in range(stack.n_layer):
        _convnext(g, s, f"decoder.decoder_block.{i}.", stack.hidden, stack.hidden * 4, "weight", 0.25)
    _conv(g, s, "decoder.conv_out", stack.odim, stack.hidden, 1, bias=False)
    _conv(g, s, "out_conv", 100, dim, 3, bias=False)
    if vq is not None:
        per_group = vq.dim // vq.G
        for grp in range(vq.G):
            # GroupedResidualFSQ -> rvqs.{g}.project_out : Linear(len(levels) -> dim/G)
            _linear(g, s, f"vq_layer.quantizer.rvqs.{grp}.project_in", len(vq.levels), per_group)
            _linear(g, s, f"vq_layer.quantizer.rvqs.{grp}.project_out", per_group, len(vq.levels))
    if encoder is not None:
        # encode-branch keys (dvae.py:229-236), drawn AFTER everything above so that the decode-side values (and the
        # committed fixtures made from them) do not move
        _conv(g, s, "downsample_conv.0", dim, 100, 3)
        _conv(g, s, "downsample_conv.2", dim, dim, 4)
        _conv(g, s, "encoder.conv_in.0", encoder.bn_dim, encoder.idim, 3)
        _conv(g, s, "encoder.conv_in.2", encoder.hidden, encoder.bn_dim, 3)
        for i in range(encoder.n_layer):
            _convnext(g, s, f"encoder.decoder_block.{i}.", encoder.hidden, encoder.hidden * 4, "weight", 0.25)
        _conv(g, s, "encoder.conv_out", encoder.odim, encoder.hidden, 1, bias=False)
        if vq is not None:
            for grp in range(vq.G):   # spread the projected values over several FSQ levels (bound() saturates beyond ~2)
                s[f"vq_layer.quantizer.rvqs.{grp}.project_in.weight"] *= 32.0
    return s


def synth_vocos_state(seed: int = 5, cfg: VocosConfig = VocosConfig(), mag_shift: float = -4.0) -> State:
    """Vocos backbone + ISTFT head keys ([3p] vocos; SURVEY.md §8b seam 2)."""
    g = _gen(seed)
    s: State = {}
    _conv(g, s, "backbone.embed", cfg.dim, cfg.input_channels, 7)
    s["backbone.norm.weight"] = 1.0 + _normal(g, (cfg.dim,), 0.05)
    s["backbone.norm.bias"] = _normal(g, (cfg.dim,), 0.05)
    for i in range(cfg.num_layers):
        _convnext(g, s, f"backbone.convnext.{i}.", cfg.dim, cfg.intermediate_dim, "gamma", 1.0 / cfg.num_layers)
    s["backbone.final_layer_norm.weight"] = 1.0 + _normal(g, (cfg.dim,), 0.05)
    s["backbone.final_layer_norm.bias"] = _normal(g, (cfg.dim,), 0.05)
    _linear(g, s, "head.out", cfg.n_fft + 2, cfg.dim)
    # keep exp(mag) well below the clip at 100: shift the magnitude half
    s["head.out.bias"][: cfg.n_fft // 2 + 1] += mag_shift
    s["head.istft.window"] = torch.hann_window(cfg.n_fft)
    return s


def synth_all(seed: int = 0, std: float = 0.02, cfg: Config = Config()) -> Dict[str, State]:
    return {
        "gpt": synth_gpt_state(seed, std, cfg.gpt),
        "embed": synth_embed_state(seed + 1, cfg.gpt),
        "decoder": synth_dvae_state(seed + 2, cfg.decoder, cfg.decoder.idim),
        "dvae": synth_dvae_state(seed + 3, cfg.dvae.decoder, cfg.dvae.decoder.idim, cfg.dvae.vq, encoder=cfg.dvae.encoder),
        "vocos": synth_vocos_state(seed + 4, cfg.vocos),
    }


def synth_speech_like(seconds: float = 2.0, seed: int = 0, sample_rate: int = 24000) -> torch.Tensor:
    """A waveform with speech-like variety for the encode-branch tests: 80 ms segments that are silent, voiced (harmonic
    stack with a moving pitch) or noisy, under a slow amplitude envelope, so that the log-mel moves over its whole range."""
    g = _gen(seed)
    n = int(seconds * sample_rate)
    seg = int(0.08 * sample_rate)
    t = torch.arange(n, dtype=torch.float32) / sample_rate
    out = torch.zeros(n)
    kinds = torch.randint(0, 4, ((n + seg - 1) // seg,), generator=g)
    pitch = 90.0 + 160.0 * torch.rand(kinds.numel(), generator=g)
    for i, k in enumerate(kinds.tolist()):
        sl = slice(i * seg, min(n, (i + 1) * seg))
        if k == 0:
            continue
        if k in (1, 2):
            f0 = float(pitch[i])
            for h in range(1, 9 if k == 1 else 4):
                out[sl] += torch.sin(2 * math.pi * f0 * h * t[sl]) / h
        else:
            out[sl] += torch.randn(t[sl].numel(), generator=g) * 0.5
    env = 0.15 + 0.35 * (1 + torch.sin(2 * math.pi * 1.7 * t)) / 2
    return (out * env * 0.5).clamp_(-1.0, 1.0)
