"""Synthetic prompt batches in the reference tokenizer's output format.

``Tokenizer.encode`` (reference tokenizer.py:35-126) returns left-padded
``input_ids[B,T,num_vq]`` (text id replicated over the vq axis), ``attention_mask[B,T]``
and ``text_mask[B,T]`` (= attention_mask.bool(), tokenizer.py:112).  No tokenizer assets
exist offline, so tests and bench draw ids directly.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def synth_prompt_batch(lengths: Sequence[int], seed: int = 1, num_vq: int = 4, lo: int = 1, hi: int = 1000
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    B, T = len(lengths), max(lengths)
    ids = torch.zeros(B, T, num_vq, dtype=torch.long)
    mask = torch.zeros(B, T, dtype=torch.bool)
    for b, n in enumerate(lengths):
        row = torch.randint(lo, hi, (n,), generator=g)
        ids[b, T - n:] = row[:, None]
        mask[b, T - n:] = True
    return ids, mask, mask.clone()
