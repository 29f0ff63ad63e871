"""Speaker strings and prompt decoration (host side): the reference's ``Speaker`` interface
(``ChatTTS/model/speaker.py:10-160``) on top of :mod:`chattts_b200.b14`.

* a speaker embedding travels as ``b14(lzma2_raw(fp16[768]))`` (speaker.py:139-160) - such strings start with "蘁淰";
* an audio prompt (DVAE codes ``[num_vq, T]``) travels as ``b14(<u2 shape> + lzma2_raw(<u2 codes>))`` (speaker.py:88-121);
* ``spk_stat`` (config.py:132) is ``b14(fp16[std(768) | mean(768)])``; a random speaker is ``N(mean, std)`` (speaker.py:123-130);
* ``apply`` overwrites the embedding of the ``[spk_emb]`` position with the L2-normalised speaker vector (speaker.py:22-50).
"""
from __future__ import annotations

import lzma
from typing import List, Optional, Union

import numpy as np
import torch

from . import b14

_LZMA = dict(format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "preset": 9 | lzma.PRESET_EXTREME}])


class Speaker:
    def __init__(self, dim: int, spk_cfg: Optional[str], device=torch.device("cpu")) -> None:
        self.dim = dim
        self.std = self.mean = None
        if spk_cfg is None:          # no statistics: everything but random sampling works
            return
        stat = np.frombuffer(b14.decode_from_string(spk_cfg), dtype=np.float16).copy()
        if stat.size != 2 * dim:
            raise ValueError(f"spk_stat holds {stat.size} values, expected std|mean of width {dim}")
        spk_stat = torch.from_numpy(stat).to(device=device)
        self.std, self.mean = spk_stat.chunk(2)

    # ------------------------------------------------------------------ sampling / strings
    def sample_random(self) -> str:
        return self._encode(self._sample_random())

    @torch.no_grad()
    def _sample_random(self) -> torch.Tensor:
        if self.std is None:
            raise RuntimeError("this Speaker was built without spk_stat (config.py:132): pass it to Chat.load(spk_stat=...)")
        return torch.randn(self.dim, device=self.std.device, dtype=self.std.dtype).mul_(self.std).add_(self.mean)

    @staticmethod
    def _encode(spk_emb: torch.Tensor) -> str:
        arr = spk_emb.detach().to(dtype=torch.float16, device="cpu").numpy()
        return b14.encode_to_string(lzma.compress(arr.tobytes(), **_LZMA))

    @staticmethod
    def _decode(spk_emb: str) -> np.ndarray:
        return np.frombuffer(lzma.decompress(b14.decode_from_string(spk_emb), **_LZMA), dtype=np.float16).copy()

    @staticmethod
    def encode_prompt(prompt: torch.Tensor) -> str:
        arr = prompt.detach().cpu().numpy().astype(np.uint16)
        if arr.ndim != 2:
            raise AssertionError("prompt must be a 2D tensor")
        head = np.array(arr.shape, dtype="<u2").tobytes()
        return b14.encode_to_string(head + lzma.compress(arr.astype("<u2").tobytes(), **_LZMA))

    @staticmethod
    def decode_prompt(prompt: str) -> torch.Tensor:
        dec = b14.decode_from_string(prompt)
        shp = np.frombuffer(dec[:4], dtype="<u2")
        codes = np.frombuffer(lzma.decompress(dec[4:], **_LZMA), dtype="<u2")
        return torch.from_numpy(codes.astype(np.int32)).view(int(shp[0]), int(shp[1]))

    # ------------------------------------------------------------------ embedding injection (speaker.py:22-50)
    @torch.inference_mode()
    def apply(self, emb: torch.Tensor, spk_emb: Union[str, torch.Tensor], input_ids: torch.Tensor, spk_emb_ids: int,
              device: torch.device, inplace: bool = True) -> torch.Tensor:
        vec = torch.from_numpy(self._decode(spk_emb)) if isinstance(spk_emb, str) else spk_emb
        n = torch.nn.functional.normalize(vec, p=2.0, dim=0, eps=1e-12).to(device).view(1, 1, -1).expand(emb.shape)
        cond = input_ids.narrow(-1, 0, 1).eq(spk_emb_ids).expand(emb.shape)
        return torch.where(cond, n, emb, out=emb if inplace else None)

    # ------------------------------------------------------------------ prompt decoration (speaker.py:52-87)
    @staticmethod
    def decorate_code_prompts(text: List[str], prompt: str, txt_smp: Optional[str], spk_emb: Optional[str]) -> List[str]:
        for i, t in enumerate(text):     # in place, like the reference (ChatTTS issue 459: user text must not carry these)
            text[i] = t.replace("[Stts]", "").replace("[spk_emb]", "").replace("[empty_spk]", "").strip()
        if prompt:
            text = [prompt + t for t in text]
        smp = "" if txt_smp is None else txt_smp
        slot = "[spk_emb]" if spk_emb is not None else "[empty_spk]"
        return [f"[Stts]{slot}{smp}{t}[Ptts]" for t in text]

    @staticmethod
    def decorate_text_prompts(text: List[str], prompt: str) -> List[str]:
        return [f"[Sbreak]{t}[Pbreak]{prompt}" for t in text]
