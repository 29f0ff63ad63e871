timeout 150 python -m pytest tests/test_gpu_gpt.py -x -q --timeout 100 -k "batched_prefill" 2>&1 | tail -4
python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, ".")
from bench import build_inputs
from chattts_b200.config import Config
from chattts_b200.embed import Embed
from chattts_b200.gpt import GPT
from chattts_b200.synth import synth_embed_state, synth_gpt_state
from chattts_b200.prompts import synth_prompt_batch
from chattts_b200.processors import build_sampler_config, gen_logits, exp_noise
dev = torch.device("cuda")
embed = Embed(768, 626, 21178, 4).load_state_dict(synth_embed_state(1)).to(dev)
gs = synth_gpt_state(0)
for env in ({}, {"CTB_NO_BATCHED_PREFILL": "1"}):
    os.environ.update(env)
    gpt = GPT(Config().gpt, embed, device=dev, device_gpt=dev, max_batch=32, max_context=192)
    gpt.load_state(gs)
    ids, mask, tmask = synth_prompt_batch([128] * 32, seed=1)
    warp, proc = gen_logits(num_code=625)
    cfg = build_sampler_config((*proc, *warp), [0.3] * 4, 625, 4)
    q = exp_noise(128, 626, 1).to(dev)
    emb_d, mask_d = embed(ids, tmask).to(dev), mask.to(dev).to(torch.uint8)
    out = torch.zeros(32, 4, 4, dtype=torch.int32, device=dev)
    for r in range(3):
        torch.cuda.synchronize(); t = time.time()
        gpt.enqueue_generate(emb_d, mask_d, cfg, q, 4, False, out, None, n_steps=0)
        torch.cuda.synchronize(); dt = time.time() - t
    print("prefill B=32 T0=128", "columns" if env else "batched", f"{dt*1e3:.1f} ms", out[0, 0].tolist(), flush=True)
    for k in env: os.environ.pop(k)
    del gpt
PY
