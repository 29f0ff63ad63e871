#!/usr/bin/env python
"""bench.py - speech-tokens/s of the GPT decode hot path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A "step" is one whole ``generate`` pass of the hot path over one batch: a 16-token prompt and
``--tokens`` (512) forced speech tokens per row, greedy + EOS excluded (BASELINE.json configs[1];
SURVEY.md 8d C2).  ``value`` = speech tokens/s with inputs resident in HBM (prompt embeddings,
mask, Exp(1) noise already on the device; CUDA events on the launching stream).  ``e2e`` = the same
metric through the public ``GPT.generate`` call with HOST buffers (pinned prompt embeddings +
noise H2D, sampled ids D2H inside the timed region).  N > 1: one process per GPU (torchrun),
utterances sharded, one NCCL broadcast of the packed weights at load, no step-loop collective.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PROMPT_LEN = 16
# streamed weight elements per audio step (20 layers + 41 norms + 4 heads)
W_ELEMS = 190_698_240
KV_BYTES_PER_TOKEN_ROW = 20 * 2 * 768 * 4  # 122,880 B per row per context token (read), same per step (write)


def algorithmic_bytes_per_step(B: int, T: float) -> float:
    return W_ELEMS * 4 + B * T * KV_BYTES_PER_TOKEN_ROW + B * KV_BYTES_PER_TOKEN_ROW


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


def build_inputs(B: int, tokens: int, seed: int):
    from chattts_b200.processors import ArgmaxOnly, build_sampler_config, exp_noise, gen_logits
    from chattts_b200.prompts import synth_prompt_batch

    ids, mask, tmask = synth_prompt_batch([PROMPT_LEN] * B, seed=seed)
    warp, proc = gen_logits(num_code=625, top_P=0.7, top_K=20, repetition_penalty=1.05)
    procs = (*proc, *warp, ArgmaxOnly(exclude_eos=True))
    cfg = build_sampler_config(procs, [0.3] * 4, 625, tokens)
    q = exp_noise(B * 4, 626, 1234)
    return ids, mask, tmask, procs, cfg, q


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist

    from chattts_b200 import _lib
    from chattts_b200.config import Config
    from chattts_b200.dist import broadcast_weights
    from chattts_b200.embed import Embed
    from chattts_b200.gpt import GPT
    from chattts_b200.synth import synth_embed_state, synth_gpt_state

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    cfg = Config()
    B, tokens = args.batch, args.tokens
    es = synth_embed_state(1)
    embed = Embed(768, 626, 21178, 4).load_state_dict(es).to(dev)
    gpt = GPT(cfg.gpt, embed, device=dev, device_gpt=dev, max_batch=max(B, 32), max_context=PROMPT_LEN + tokens + 16)
    if world > 1:
        # one NCCL broadcast of the packed blob at load; only rank 0 builds it
        blob = broadcast_weights(gpt, synth_gpt_state(0) if rank == 0 else None, src=0)
        gpt.load_state(None, weights_blob=blob)
    else:
        gpt.load_state(synth_gpt_state(0))

    ids, mask, tmask, procs, scfg, q = build_inputs(B, tokens, seed=1 + rank)
    emb_host = embed(ids, tmask).cpu().pin_memory()
    # ---- resident buffers for `value`
    emb_d, mask_d, q_d = emb_host.to(dev), mask.to(dev).to(torch.uint8), q.to(dev)
    ids_out = torch.zeros(B, tokens, 4, dtype=torch.int32, device=dev)

    def step_resident():
        gpt.enqueue_generate(emb_d, mask_d, scfg, q_d, tokens, False, ids_out, None)

    def step_e2e():
        out = list(gpt.generate(emb_host, ids, temperature=torch.tensor([0.3] * 4), eos_token=625,
                                attention_mask=mask, max_new_token=tokens, min_new_token=tokens,
                                logits_processors=procs, return_hidden=False, show_tqdm=False, manual_seed=1234))[-1]
        return [t.cpu() for t in out.ids]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    l0 = lib.ctb_launch_count()
    with ClockSampler(local_rank) as clk:
        ms = timed(step_resident, args.steps)
    launches = int(lib.ctb_launch_count() - l0)
    ms_per_step = ms / args.steps
    value = world * B * tokens / (ms_per_step / 1e3)

    # ---- e2e through the public API with host buffers
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
    e2e_value = world * B * tokens / (ms_e2e / 1e3)
    h2d = emb_host.numel() * 4 + mask.numel() + q.numel() * 4
    d2h = B * tokens * 4 * 4 + (16 + 5 * B) * ((tokens + 31) // 32 + 1)

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream
    peak, peak_src = measured_peaks()
    stream_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    step_resident()
    torch.cuda.synchronize()

    def time_kind(kind, reps, per_call):
        for _ in range(3):
            _lib.check(lib.ctb_gpt_profile_kernel(gpt._handle, kind, stream_ptr))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _lib.check(lib.ctb_gpt_profile_kernel(gpt._handle, kind, stream_ptr))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * per_call)  # us per launch

    t_avg = PROMPT_LEN + tokens / 2
    step_bytes = algorithmic_bytes_per_step(B, t_avg)
    step_us = ms_per_step * 1e3 / (tokens + PROMPT_LEN - 1)
    one_kernel = B == 1 and not os.environ.get("CTB_NO_MEGA")
    kern = {}
    if one_kernel:
        # B = 1 runs the whole step as ONE persistent cooperative kernel (k_step: 101 grid-barrier phases)
        step_resident()  # fresh state: the hook advances the context by one token per call
        torch.cuda.synchronize()
        us = time_kind(7, 16, 1)
        kbytes = algorithmic_bytes_per_step(B, PROMPT_LEN + tokens + 10)
        achieved = kbytes / (us * 1e-6) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_step<1> (one launch = the whole decode step: 20 layers + heads, 101 phases)",
                    "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": 771580416, "peak_source": peak_src, "bytes_per_launch": int(kbytes),
                    "us_per_launch": round(us, 2),
                    "traffic_note": "dram__bytes_read+write of one k_step<1> launch at context 64 from ncu --set full "
                                    "(profiles/r01_k_step_b1_full_raw.csv); algorithmic bytes at that context: 770.8 MB"}
    else:
        for kind, name in ((3, "gateup"), (4, "down"), (0, "qkv"), (2, "oproj"), (1, "k_attn"), (5, "heads"), (6, "k_sample")):
            kern[name] = time_kind(kind, 20, 20 if kind < 5 else 1)
        gu_bytes = 2 * 3072 * 768 * 4 + B * 768 * 4 + B * 3072 * 4  # weights + x in + mlp out
        achieved = gu_bytes / (kern["gateup"] * 1e-6) / 1e9
        kname = "k_tc_dec<DE_GATEUP> (tcgen05 3xTF32)" if B > 16 else "k_gemv<BT,EPI_GATEUP>"
        roofline = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                    "bytes_per_launch": gu_bytes, "us_per_launch": round(kern["gateup"], 3),
                    "kernel_us": {k: round(v, 3) for k, v in kern.items()}}
    roofline["whole_step"] = {"algorithmic_bytes": int(step_bytes), "us": round(step_us, 2),
                              "achieved_gbs": round(step_bytes / (step_us * 1e-6) / 1e9, 1),
                              "frac": round(step_bytes / (step_us * 1e-6) / 1e9 / peak, 4)}

    # ---- the metric at the other batch sizes it is quoted on, and hot path 2 (BASELINE configs[3]); short runs
    sweep = {}
    if world == 1 and not args.no_sweep:
        del gpt
        torch.cuda.empty_cache()
        for bb in (8, 32):
            g2 = GPT(cfg.gpt, embed, device=dev, device_gpt=dev, max_batch=bb, max_context=PROMPT_LEN + tokens + 16)
            g2.load_state(synth_gpt_state(0))
            i2, m2, tm2, _, sc2, q2 = build_inputs(bb, tokens, seed=1)
            e2, mk2, qd2 = embed(i2, tm2).to(dev), m2.to(dev).to(torch.uint8), q2.to(dev)
            o2 = torch.zeros(bb, tokens, 4, dtype=torch.int32, device=dev)
            g2.enqueue_generate(e2, mk2, sc2, qd2, tokens, False, o2, None)  # warm-up (graph capture, clocks)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                g2.enqueue_generate(e2, mk2, sc2, qd2, tokens, False, o2, None)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / 2
            sb = algorithmic_bytes_per_step(bb, t_avg)
            us2 = ms2 * 1e3 / (tokens + PROMPT_LEN - 1)
            sweep[f"batch_{bb}"] = {"value": round(bb * tokens / (ms2 / 1e3), 1), "unit": "speech-tokens/s",
                                    "ms_per_step": round(ms2, 2), "rtf": round((ms2 / 1e3) / (bb * tokens * 512 / 24000.0), 6),
                                    "step_us": round(us2, 1), "hbm_frac": round(sb / (us2 * 1e-6) / 1e9 / peak, 4)}
            del g2
            torch.cuda.empty_cache()
        sweep["decoder_c4"] = bench_decoder(dev)

    # the CPU baseline is timed on rank 0 at N = 1 only (it would otherwise compete with the other ranks' host threads)
    cpu = cpu_baseline_sample(B) if world == 1 else None
    audio_s = B * tokens * 512 / 24000.0
    line = {
        "metric": "speech-tokens/sec (GPT decode loop, 4-codebook tokens; RTF = wall / audio seconds @ 24 kHz)",
        "value": round(value, 2), "unit": "speech-tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random-init weights of the ChatTTS GPT shape)",
        "config": {"workload": f"GPT decode: batch {B}/GPU x (16-token prompt + {tokens} forced speech tokens), greedy "
                               "(BASELINE configs[1]); one step = one whole generate pass",
                   "batch_per_gpu": B, "tokens": tokens, "prompt_len": PROMPT_LEN, "parallelism": f"dp{world}",
                   "l2_policy": "inputs larger than L2: every decode iteration streams 763 MB of fp32 weights (> 126 MB L2)"},
        "rtf": round((ms_per_step / 1e3) / audio_s, 6),
        "e2e": {"value": round(e2e_value, 2), "unit": "speech-tokens/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e, 3),
                "rtf": round((ms_e2e / 1e3) / audio_s, 6)},
        "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
    }
    if sweep:
        line["other_configs"] = sweep
    return line


def bench_decoder(dev, B: int = 64, T: int = 469):
    """Hot path 2 at BASELINE configs[3]: DVAE decoder + Vocos + iSTFT of 10 s of hidden states, batch 64."""
    from chattts_b200.config import Config
    from chattts_b200.decoder import DVAE, Vocos
    from chattts_b200.synth import synth_dvae_state, synth_vocos_state

    cfg = Config()
    voc = Vocos(cfg.vocos, dev, max_batch=B, max_tokens=T)
    voc.state = synth_vocos_state(5)
    dec = DVAE(cfg.decoder, dim=cfg.decoder.idim, device=dev, vocos=voc, max_batch=B, max_tokens=T)
    dec.load_state_dict(synth_dvae_state(2, cfg.decoder, cfg.decoder.idim))
    x = torch.randn(B, T, 768, generator=torch.Generator().manual_seed(1)).to(dev)
    for _ in range(2):
        wav = dec.engine.tokens_to_wav(x, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        wav = dec.engine.tokens_to_wav(x, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    frames = B * 2 * T
    flops = frames * 78.7e6  # 78.7 MFLOP per mel frame on the hidden path
    return {"workload": f"DVAE decoder + Vocos + iSTFT, batch {B} x {T} tokens (10 s each), hidden path, tcgen05 3xTF32 GEMMs",
            "ms": round(ms, 2), "audio_samples_per_s": round(wav.numel() / (ms / 1e3), 1),
            "rtf": round((ms / 1e3) / (wav.numel() / 24000.0), 7), "tflops_fp32_equiv": round(flops / (ms / 1e3) / 1e12, 1)}


def best_cpu_threads(run4, candidates=(16, 32, 64, 128)):
    """Pick the torch thread count that makes the oracle port fastest on this host (the reference arm must use the
    host as well as it can; more threads are not always faster for a batch-1 GEMV chain)."""
    cores = os.cpu_count() or 1
    best, best_t = None, None
    for n in candidates:
        if n > cores:
            break
        torch.set_num_threads(n)
        run4()
        t = run4()
        if best_t is None or t < best_t:
            best, best_t = n, t
    if best is None:
        best = cores
    torch.set_num_threads(best)
    return best


def cpu_baseline_sample(B: int, budget_s: float = 8.0):
    """The oracle port (torch fp32 CPU, same ops as the reference's HF path) on this host's cores,
    on a bounded sample of the same workload."""
    from chattts_b200.prompts import synth_prompt_batch
    from chattts_b200.synth import synth_embed_state, synth_gpt_state
    from oracle.gpt_oracle import GPTOracle, SamplerParams

    cores = os.cpu_count() or 1
    orc = GPTOracle(synth_gpt_state(0), synth_embed_state(1))
    ids, mask, tmask = synth_prompt_batch([PROMPT_LEN] * B, seed=1)
    sp = SamplerParams(greedy=True, greedy_exclude_eos=True)

    def run(n):
        t = time.perf_counter()
        orc.generate(orc.embed_prompt(ids, tmask), ids, torch.tensor([0.3] * 4), 625, attention_mask=mask,
                     max_new_token=n, min_new_token=n, sampler=sp, manual_seed=1234)
        return time.perf_counter() - t

    threads = best_cpu_threads(lambda: run(4))
    t4 = run(4)
    n = int(max(8, min(256, budget_s / max(t4 / 4, 1e-3))))
    t = run(n)
    return {"value": round(B * n / t, 2), "unit": "speech-tokens/s", "cores": threads, "kind": "port",
            "sample": f"oracle/gpt_oracle.py generate(): batch {B}, 16-token prompt + {n} tokens, {threads} torch threads "
                      f"on a {cores}-core host ({t:.1f} s)", "ms_per_token_step": round(1e3 * t / n, 2)}


def run_reference(args, rank: int):
    """Reference arm: the reference's CPU implementation of the path (oracle port: torch fp32 CPU),
    each step a bounded sample of the same workload."""
    if rank != 0:
        return None
    from chattts_b200.prompts import synth_prompt_batch
    from chattts_b200.synth import synth_embed_state, synth_gpt_state
    from oracle.gpt_oracle import GPTOracle, SamplerParams

    cores = os.cpu_count() or 1
    B, n = args.batch, args.ref_tokens
    orc = GPTOracle(synth_gpt_state(0), synth_embed_state(1))
    ids, mask, tmask = synth_prompt_batch([PROMPT_LEN] * B, seed=1)
    sp = SamplerParams(greedy=True, greedy_exclude_eos=True)

    def gen(k):
        t = time.perf_counter()
        orc.generate(orc.embed_prompt(ids, tmask), ids, torch.tensor([0.3] * 4), 625, attention_mask=mask,
                     max_new_token=k, min_new_token=k, sampler=sp, manual_seed=1234)
        return time.perf_counter() - t

    threads = best_cpu_threads(lambda: gen(4))

    def step():
        gen(n)

    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t) / args.steps
    value = B * n / dt
    sample = (f"oracle port of GPT.generate on CPU: batch {B}, 16-token prompt + {n} tokens per step "
              f"(bounded sample of the {args.tokens}-token workload), {threads} torch threads of {cores} cores")
    return {
        "impl": "reference", "metric": "speech-tokens/sec (GPT decode loop, 4-codebook tokens; RTF = wall / audio seconds @ 24 kHz)",
        "value": round(value, 2), "unit": "speech-tokens/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random-init weights of the ChatTTS GPT shape)",
        "config": {"workload": f"GPT decode: batch {B} x (16-token prompt + {args.tokens} forced speech tokens), greedy "
                               "(BASELINE configs[1]); reference arm times a bounded sample",
                   "batch_per_gpu": B, "tokens": args.tokens, "prompt_len": PROMPT_LEN},
        "rtf": round(dt / (B * n * 512 / 24000.0), 4),
        "cpu_baseline": {"value": round(value, 2), "unit": "speech-tokens/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "speech-tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("CTB_BENCH_BATCH", "1")))
    ap.add_argument("--tokens", type=int, default=512)
    ap.add_argument("--ref-tokens", type=int, default=48)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the short batch-8/32 and decoder side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = run_reference(args, rank)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        line = run_ours(args, rank, world, local_rank)
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
