"""Shape/hyper-parameter tree for the two hot paths.

Mirrors the *fields and defaults* of the reference's dataclass tree so that
``Chat.config.gpt.num_vq`` & friends keep working for callers
(reference: ChatTTS/config/config.py:5-11 paths, :15-20 decoder, :24-28 VQ, :32-47 DVAE,
:51-63 GPT, :67-71 embed, :74-121 vocos).  Only fields the hot paths read are kept; the
base16384 speaker-statistics blob (config.py:134) belongs to the out-of-scope ``Speaker``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


@dataclass(frozen=True)
class PathConfig:
    vocos_ckpt_path: str = "asset/Vocos.safetensors"
    dvae_ckpt_path: str = "asset/DVAE.safetensors"
    gpt_ckpt_path: str = "asset/gpt"
    decoder_ckpt_path: str = "asset/Decoder.safetensors"
    tokenizer_path: str = "asset/tokenizer"
    embed_path: str = "asset/Embed.safetensors"


@dataclass(frozen=True)
class ConvStackConfig:
    """conv_in -> n_layer ConvNeXt blocks -> 1x1 conv_out (reference DVAEDecoder)."""

    idim: int
    odim: int
    hidden: int
    n_layer: int = 12
    bn_dim: int = 128
    kernel: int = 7
    dilation: int = 2


@dataclass(frozen=True)
class VQConfig:
    dim: int = 1024
    levels: Tuple[int, ...] = (5, 5, 5, 5)
    G: int = 2
    R: int = 2


@dataclass(frozen=True)
class DVAEConfig:
    encoder: ConvStackConfig = ConvStackConfig(idim=512, odim=1024, hidden=256)
    decoder: ConvStackConfig = ConvStackConfig(idim=512, odim=512, hidden=256)
    vq: VQConfig = VQConfig()


@dataclass(frozen=True)
class GPTConfig:
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_attention_heads: int = 12
    num_key_value_heads: int = 12
    head_dim: int = 64
    num_hidden_layers: int = 20
    max_position_embeddings: int = 4096
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    num_audio_tokens: int = 626
    num_text_tokens: int = 21178
    num_vq: int = 4
    use_cache: bool = False
    spk_emb_dim: int = 192
    spk_KL: bool = False


@dataclass(frozen=True)
class EmbedConfig:
    hidden_size: int = 768
    num_audio_tokens: int = 626
    num_text_tokens: int = 21178
    num_vq: int = 4


@dataclass(frozen=True)
class VocosConfig:
    input_channels: int = 100
    dim: int = 512
    intermediate_dim: int = 1536
    num_layers: int = 8
    n_fft: int = 1024
    hop_length: int = 256
    sample_rate: int = 24000
    padding: str = "center"


@dataclass(frozen=True)
class Config:
    path: PathConfig = PathConfig()
    decoder: ConvStackConfig = ConvStackConfig(idim=384, odim=384, hidden=512)
    dvae: DVAEConfig = DVAEConfig()
    gpt: GPTConfig = GPTConfig()
    embed: EmbedConfig = EmbedConfig()
    vocos: VocosConfig = VocosConfig()


#: one speech token = 2 mel frames = 512 samples @ 24 kHz  
SAMPLES_PER_TOKEN = 512
SAMPLE_RATE = 24000
