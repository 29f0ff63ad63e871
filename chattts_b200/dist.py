"""Multi-GPU data parallelism for the hot paths.

Utterances are independent (attention is per row, the sampler per (row, codebook), the decoder
convolves within a row), so the path shards with NO data-path collective: one process per GPU,
contiguous blocks of the global batch per rank, results gathered on the host.  The only
collective is one broadcast of the packed fp32 weight blob at load (NCCL over NVLink/NVSwitch
on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of the global batch owned by ``rank`` (global row = start + local row)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_list(items: Sequence, rank: int, world: int) -> list:
    a, b = shard_range(len(items), rank, world)
    return list(items[a:b])


def broadcast_blob(blob: Optional[torch.Tensor], numel: int, device, src: int = 0) -> torch.Tensor:
    """One broadcast of a flat fp32 buffer from ``src``; other ranks pass ``blob=None``."""
    if dist.get_rank() == src:
        assert blob is not None and blob.numel() == numel
        buf = blob.to(device, torch.float32).contiguous()
    else:
        buf = torch.empty(numel, dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)
    return buf


def broadcast_weights(gpt, gpt_state: Optional[Dict[str, torch.Tensor]], src: int = 0) -> torch.Tensor:
    """Rank ``src`` packs the GPT/Embed checkpoint into the kernels' blob layout; everyone receives
    it with a single collective and can then call ``gpt.load_state(None, weights_blob=blob)``."""
    _, lay = gpt.query_layout()
    blob = gpt.pack_weights(gpt_state, lay) if dist.get_rank() == src else None
    return broadcast_blob(blob, int(lay.total), gpt.device_gpt, src)


def gather_object_lists(local: list, dst: int = 0) -> Optional[List]:
    """Host-side gather of per-rank result lists (token id tensors / waveforms are KB-MB)."""
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    if out is None:
        return None
    return [x for part in out for x in part]
