"""Stand-alone entry to the fused sampling kernel (``ctb_sample``) - the minimum slice of
SURVEY.md §7.2; the decode loop uses the same kernel."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def sample_rows(logits: torch.Tensor, cfg: "_lib.SamplerConfig", rows_per_item: int,
                q_noise: Optional[torch.Tensor], gen_ids: Optional[torch.Tensor], step: int = 0) -> torch.Tensor:
    """logits [rows, V] fp32 cuda; gen_ids [rows/rpi, n_gen, rpi] int32 cuda (tokens so far)."""
    _lib.require_cuda()
    lib = _lib.load()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    rows, V = logits.shape
    out = torch.empty(rows, dtype=torch.int32, device=logits.device)
    n_gen = 0 if gen_ids is None else int(gen_ids.shape[1])
    if gen_ids is not None:
        gen_ids = gen_ids.to(torch.int32).contiguous()
    with torch.cuda.device(logits.device):
        _lib.check(lib.ctb_sample(
            C.c_void_p(logits.data_ptr()), rows, V, rows_per_item, C.byref(cfg),
            C.c_void_p(q_noise.data_ptr()) if q_noise is not None else None,
            C.c_void_p(gen_ids.data_ptr()) if gen_ids is not None else None, max(n_gen, 1), n_gen, step,
            C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out
