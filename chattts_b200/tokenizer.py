"""Prompt tokenizer with the reference's interface (``ChatTTS/model/tokenizer.py:16-160``): a HF ``BertTokenizerFast``
loaded from the asset folder plus the batch layout ``GPT.generate`` expects.

``encode(text, num_vq, prompt=None, device)`` -> ``(input_ids [B, T, num_vq] , attention_mask [B, T], text_mask [B, T])``:

* every text is tokenised on its own (no special tokens added) and LEFT-padded to the longest one (tokenizer.py:79-103) -
  which is what the device-side position logic (``pos = number of valid tokens so far``) relies on;
* the text id is repeated over the ``num_vq`` code slots (tokenizer.py:118-119);
* an audio prompt ``[num_vq, P]`` (DVAE codes of a speaker sample, ``Speaker.decode_prompt``) is appended AFTER the text of
  every row with ``attention_mask = 1`` and ``text_mask = 0`` - those columns are embedded as audio codes (tokenizer.py:120-133).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")


class Tokenizer:
    def __init__(self, tokenizer_path):
        from transformers import BertTokenizerFast

        tok = BertTokenizerFast.from_pretrained(tokenizer_path)
        self._tokenizer = tok
        self.len = len(tok)
        self.spk_emb_ids = tok.convert_tokens_to_ids("[spk_emb]")
        self.break_0_ids = tok.convert_tokens_to_ids("[break_0]")
        self.eos_token = tok.convert_tokens_to_ids("[Ebreak]")

    @torch.inference_mode()
    def encode(self, text: List[str], num_vq: int, prompt: Optional[torch.Tensor] = None, device="cpu"
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        P = 0
        if prompt is not None:
            assert prompt.size(0) == num_vq, "prompt dim 0 must equal to num_vq"
            P = int(prompt.size(1))
        # tokenizer.py:59-62 calls encode_plus, which transformers >= 5 dropped; __call__ is the same operation in both
        rows = [self._tokenizer(t, return_tensors="pt", add_special_tokens=False, padding=True) for t in text]
        ids = [r["input_ids"].squeeze(0) for r in rows]
        att = [r["attention_mask"].squeeze(0) for r in rows]
        T = max(int(i.size(0)) for i in ids) + P
        input_ids = torch.zeros(len(ids), T, device=device, dtype=ids[0].dtype)
        attention_mask = torch.zeros(len(ids), T, device=device, dtype=att[0].dtype)
        for b, (i, a) in enumerate(zip(ids, att)):
            n = int(i.size(0))
            input_ids[b, T - P - n: T - P] = i
            attention_mask[b, T - P - n: T - P] = a
        if P:
            attention_mask[:, T - P:] = 1
        text_mask = attention_mask.bool()
        input_ids = input_ids.unsqueeze(-1).expand(-1, -1, num_vq).clone()
        if P:
            text_mask[:, T - P:] = False
            input_ids[:, T - P:] = prompt.t().unsqueeze(0).to(device=device, dtype=input_ids.dtype)
        return input_ids, attention_mask, text_mask

    @torch.inference_mode()
    def decode(self, sequences, skip_special_tokens: bool = False, clean_up_tokenization_spaces: bool = None, **kwargs):
        return self._tokenizer.batch_decode(sequences, skip_special_tokens, clean_up_tokenization_spaces, **kwargs)
