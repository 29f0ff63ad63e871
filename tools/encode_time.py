"""Timing of the DVAE encode branch (speaker enrolment): ctb_dvae_encode on the GPU (CUDA events, inputs
resident / end to end from host memory) beside the CPU oracle on the host cores.  One JSON line per audio length."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from chattts_b200.config import Config
from chattts_b200.decoder import AudioEncoder, pack_dvae_encoder
from chattts_b200.synth import synth_dvae_state, synth_speech_like


def main():
    cfg = Config()
    st = synth_dvae_state(3, cfg.dvae.decoder, cfg.dvae.decoder.idim, cfg.dvae.vq, encoder=cfg.dvae.encoder)
    enc = AudioEncoder(cfg.dvae.encoder, cfg.dvae.decoder.idim, cfg.dvae.vq,
                       pack_dvae_encoder(st, cfg.dvae.encoder, cfg.dvae.decoder.idim, cfg.dvae.vq), "cuda", max_samples=24000 * 31)
    from oracle.dvae_oracle import dvae_encode   # cpu_baseline leg only

    for seconds in (3.0, 10.0, 30.0):
        wav_host = synth_speech_like(seconds, 1).pin_memory()
        wav = wav_host.cuda()
        for _ in range(3):
            enc.encode(wav)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ids = enc.encode(wav)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t0 = time.perf_counter()
        for _ in range(10):
            out = enc.encode(wav_host.cuda(non_blocking=True)).cpu()
        ms_e2e = (time.perf_counter() - t0) * 100
        torch.set_num_threads(16)
        t0 = time.perf_counter()
        ref = dvae_encode(wav_host.clone(), st, precise_stft=True)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        same = float((ref[0].int() == out).float().mean())
        print(json.dumps({"audio_s": seconds, "tokens": int(ids.shape[1]), "gpu_ms": round(ms, 3), "gpu_e2e_ms": round(ms_e2e, 3),
                          "samples_per_s": round(wav.numel() / (ms / 1e3), 1), "cpu_oracle_ms_16_threads": round(cpu_ms, 1),
                          "ids_equal_frac": same}))


if __name__ == "__main__":
    main()
