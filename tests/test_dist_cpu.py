"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: utterance sharding and the single
weight-blob broadcast.  No collective exists in the step loop, so this is all of it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chattts_b200.dist import shard_list, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_list(list("abcdefg"), 1, 3) == ["d", "e"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chattts_b200.config import GPTConfig
        from chattts_b200.dist import broadcast_weights, gather_object_lists
        from chattts_b200.embed import Embed
        from chattts_b200.gpt import GPT
        from chattts_b200.synth import synth_embed_state, synth_gpt_state

        cfg = GPTConfig(num_hidden_layers=1, num_text_tokens=512, max_position_embeddings=64)
        embed = Embed(768, 626, 512, 4).load_state_dict(synth_embed_state(1, cfg))
        gpt = GPT(cfg, embed, device="cpu", device_gpt="cpu", max_batch=2, max_context=32)
        state = synth_gpt_state(0, cfg=cfg) if rank == 0 else None
        blob = broadcast_weights(gpt, state, src=0)
        _, lay = gpt.query_layout()
        assert blob.numel() == lay.total
        ref = gpt.pack_weights(synth_gpt_state(0, cfg=cfg), lay)
        ok = torch.equal(blob, ref)
        items = shard_list(list(range(5)), rank, world)
        gathered = gather_object_lists([i * 10 for i in items], dst=0)
        if rank == 0:
            q.put((ok, gathered))
        else:
            q.put((ok, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_weight_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(ok for ok, _ in res)
    assert [g for _, g in res if g is not None][0] == [0, 10, 20, 30, 40]
