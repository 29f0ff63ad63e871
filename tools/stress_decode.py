"""Stress / timing helper: repeated device-resident generate() passes at a given batch (no oracle)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_inputs
from chattts_b200.config import Config
from chattts_b200.embed import Embed
from chattts_b200.gpt import GPT
from chattts_b200.synth import synth_embed_state, synth_gpt_state

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
embed = Embed(768, 626, 21178, 4).load_state_dict(synth_embed_state(1)).to(dev)
gpt = GPT(Config().gpt, embed, device=dev, device_gpt=dev, max_batch=max(B, 1), max_context=16 + tokens + 16)
gpt.load_state(synth_gpt_state(0))
ids, mask, tmask, procs, scfg, q = build_inputs(B, tokens, seed=1)
emb_d, mask_d, q_d = embed(ids, tmask).to(dev), mask.to(dev).to(torch.uint8), q.to(dev)
ids_out = torch.zeros(B, tokens, 4, dtype=torch.int32, device=dev)
for r in range(reps):
    t = time.time()
    gpt.enqueue_generate(emb_d, mask_d, scfg, q_d, tokens, False, ids_out, None)
    torch.cuda.synchronize()
    dt = time.time() - t
    print(f"B={B} rep {r}: {dt*1e3:.1f} ms  {dt*1e6/(tokens+15):.1f} us/step  ids[0,-1]={ids_out[0,-1].tolist()}", flush=True)

if os.environ.get("CTB_KERNEL_US"):
    import ctypes as C

    from chattts_b200 import _lib

    lib = _lib.load()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for kind, name in ((0, "qkv"), (1, "attn"), (2, "oproj"), (3, "gateup"), (4, "down"), (5, "heads"), (6, "sample")):
        for _ in range(2):
            _lib.check(lib.ctb_gpt_profile_kernel(gpt._handle, kind, sp))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _lib.check(lib.ctb_gpt_profile_kernel(gpt._handle, kind, sp))
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) * 1e3 / (10 * (20 if kind < 5 else 1)), 2)
    print(f"B={B} kernel_us {out}  layer_sum={sum(v for k, v in out.items() if k not in ('heads', 'sample')):.1f}", flush=True)

if os.environ.get("CTB_MEGA_TRACE"):
    import ctypes as C
    import numpy as np

    from chattts_b200 import _lib

    buf = np.zeros(128, dtype=np.uint64)
    _lib.check(_lib.load().ctb_gpt_debug_trace(gpt._handle, buf.ctypes.data_as(C.c_void_p), 128))
    t = buf[:103].astype(np.int64)
    d = np.diff(t) / 1e3
    names = ["qkv", "attn", "oproj", "gateup", "down"]
    per = {n: [] for n in names}
    for l in range(20):
        for j, n in enumerate(names):
            per[n].append(d[l * 5 + j])
    print("phase us (median over layers):", {n: round(float(np.median(v)), 2) for n, v in per.items()},
          "heads", round(float(d[100]), 2), "total", round(float(t[101] - t[0]) / 1e3, 1))
    print("layer 10:", [round(float(x), 2) for x in d[50:55]])
