"""Embedding tables + weight-normed heads (reference ChatTTS/model/embed.py).

Holds the checkpoint tensors (same state-dict names as the reference, embed.py:18-35) on the
device.  The heads are folded once at load (``W = g * v / ||v||``) and, together with the
tables, become part of the packed GPT weight blob the decode kernels stream.
"""
from __future__ import annotations

from typing import Dict

import torch


class Embed:
    def __init__(self, hidden_size: int, num_audio_tokens: int, num_text_tokens: int, num_vq: int = 4):
        self.model_dim = hidden_size
        self.num_audio_tokens = num_audio_tokens
        self.num_text_tokens = num_text_tokens
        self.num_vq = num_vq
        self.state: Dict[str, torch.Tensor] = {}
        self.device = torch.device("cpu")
        self._gpt = None  # set by GPT.load_state: the tables then live in the packed blob of that handle

    # embed.py:37-41
    def load_pretrained(self, filename: str, device: torch.device):
        from safetensors.torch import load_file

        self.load_state_dict(load_file(filename))
        self.to(device)

    def load_state_dict(self, state: Dict[str, torch.Tensor]):
        need = [f"emb_code.{q}.weight" for q in range(self.num_vq)] + ["emb_text.weight"]
        for prefix in ["head_text"] + [f"head_code.{q}" for q in range(self.num_vq)]:
            need += [prefix + ".parametrizations.weight.original0", prefix + ".parametrizations.weight.original1"]
        missing = [k for k in need if k not in state]
        if missing:
            raise KeyError(f"Embed state dict misses {missing}")
        self.state = {k: state[k].detach().to(torch.float32).contiguous() for k in need}
        return self

    def to(self, device):
        self.device = torch.device(device)
        self.state = {k: v.to(self.device) for k, v in self.state.items()}
        return self

    def eval(self):
        return self

    def folded_head(self, prefix: str) -> torch.Tensor:
        """weight_norm(dim=0): W = g * v / ||v||_row, evaluated like torch's parametrization."""
        g = self.state[prefix + ".parametrizations.weight.original0"]
        v = self.state[prefix + ".parametrizations.weight.original1"]
        return torch._weight_norm(v, g, 0)

    @torch.inference_mode()
    def __call__(self, input_ids: torch.Tensor, text_mask: torch.Tensor) -> torch.Tensor:
        """Prompt embedding mix (embed.py:51-79): text rows use ``emb_text(ids[...,0])``, the
        others the sum of the four code embeddings."""
        if self._gpt is not None and self._gpt._handle:
            return self._gpt.embed_prompt(input_ids, text_mask)
        dev = self.device
        ids = input_ids.to(dev)
        tm = text_mask.to(dev).bool()
        text = torch.nn.functional.embedding(ids[..., 0], self.state["emb_text.weight"])
        code_ids = ids.clamp(max=self.num_audio_tokens - 1)
        code = torch.stack(
            [torch.nn.functional.embedding(code_ids[..., q], self.state[f"emb_code.{q}.weight"])
             for q in range(self.num_vq)], -1).sum(-1)
        return torch.where(tm[..., None], text, code)
