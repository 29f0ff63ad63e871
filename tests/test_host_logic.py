"""CPU tests of host-side logic that the GPU parity design relies on."""
import numpy as np
import pytest
import torch

from chattts_b200.processors import (ArgmaxOnly, CustomRepetitionPenaltyLogitsProcessorRepeat, TopKLogitsWarper,
                                     TopPLogitsWarper, build_sampler_config, exp_noise, gen_logits)


@pytest.mark.parametrize("rows,V,seed", [(4, 626, 1234), (128, 626, 42), (3, 21178, 7)])
def test_multinomial_is_argmax_of_p_over_exponential_noise(rows, V, seed):
    """The identity the sampler kernel is built on: with the generator re-seeded,
    torch.multinomial(p, 1) == argmax(p / q), q = Exp(1) noise of that generator, prefix-stable in the rows."""
    g = torch.Generator().manual_seed(99)
    p = torch.softmax(torch.randn(rows, V, generator=g) * 2, -1)
    want = torch.multinomial(p, 1, generator=torch.Generator().manual_seed(seed))[:, 0]
    q = exp_noise(rows, V, seed)
    assert torch.equal(torch.argmax(p / q, -1), want)
    assert torch.equal(exp_noise(rows + 5, V, seed)[:rows], q)  # prefix stability in the row dimension


def test_gen_logits_contract_matches_reference_factory():
    warp, proc = gen_logits(num_code=625, top_P=0.7, top_K=20, repetition_penalty=1.05)
    assert [type(w).__name__ for w in warp] == ["TopPLogitsWarper", "TopKLogitsWarper"]
    assert warp[0].min_tokens_to_keep == 3 and warp[1].top_k == 20
    assert len(proc) == 1 and proc[0].past_window == 16 and proc[0].max_input_ids == 625
    warp, proc = gen_logits(num_code=625, top_P=None, top_K=1, repetition_penalty=1.0)
    assert len(warp) == 1 and warp[0].top_k == 3 and proc == []  # min_tokens_to_keep clamps k (quirk Q3); penalty 1 => none


def test_sampler_config_translation_and_order_enforcement():
    warp, proc = gen_logits(num_code=625, top_P=0.7, top_K=20, repetition_penalty=1.05)
    cfg = build_sampler_config((*proc, *warp, ArgmaxOnly(exclude_eos=True)), [0.3] * 4, 625, 7)
    assert cfg.penalty_on == 1 and cfg.past_window == 16 and cfg.penalty_max_ids == 625 and cfg.greedy == 2
    assert abs(cfg.top_p - 0.7) < 1e-7 and cfg.top_k == 20 and cfg.min_tokens_to_keep == 3 and cfg.min_new_token == 7
    lut = torch.pow(1.05, torch.arange(32))  # the reference's own alpha computation (processors.py:28)
    assert all(cfg.penalty_lut[i] == float(lut[i]) for i in range(17))
    with pytest.raises(ValueError):  # top-k before top-p is not the reference's order (core.py:649)
        build_sampler_config((TopKLogitsWarper(5), TopPLogitsWarper(0.5)), [0.3], 625, 0)
    with pytest.raises(ValueError):
        CustomRepetitionPenaltyLogitsProcessorRepeat(0.0, 625, 16)
    with pytest.raises(TypeError):
        build_sampler_config((object(),), [0.3], 625, 0)


def test_hf_warpers_are_accepted_by_attribute_names():
    from transformers.generation import TopKLogitsWarper as HFK, TopPLogitsWarper as HFP

    cfg = build_sampler_config((HFP(0.9, min_tokens_to_keep=3), HFK(5, min_tokens_to_keep=3)), [1.0], 10, 0)
    assert abs(cfg.top_p - 0.9) < 1e-7 and cfg.top_k == 5 and cfg.min_tokens_to_keep == 3


def test_idft_basis_is_windowed_irfft():
    from chattts_b200.decoder import idft_basis

    w = torch.hann_window(1024)
    B = idft_basis(1024, w, 1056)
    S = torch.randn(513, dtype=torch.complex64, generator=torch.Generator().manual_seed(1))
    v = torch.zeros(1056)
    v[0:1026:2], v[1:1026:2] = S.real, S.imag
    assert (B @ v - torch.fft.irfft(S, 1024) * w).abs().max() < 1e-6
    assert float(B[:, 1026:].abs().max()) == 0.0  # padded spectrum columns contribute nothing


def test_chat_api_surface_matches_reference_names():
    from chattts_b200 import Chat

    c = Chat()
    for name in ("load", "infer", "interrupt", "unload", "has_loaded", "sample_random_speaker", "sample_audio_speaker",
                 "_infer", "_infer_code", "_refine_text", "_decode_to_wavs", "_vocos_decode"):
        assert hasattr(c, name), name
    p = Chat.InferCodeParams()
    assert (p.prompt, p.temperature, p.repetition_penalty, p.max_new_token, p.stream_batch, p.stream_speed,
            p.pass_first_n_batches, p.top_P, p.top_K) == ("[speed_5]", 0.3, 1.05, 2048, 24, 12000, 2, 0.7, 20)
    r = Chat.RefineTextParams()
    assert (r.temperature, r.repetition_penalty, r.max_new_token, r.top_P, r.top_K) == (0.7, 1.0, 384, 0.7, 20)
    assert c.config.gpt.num_vq == 4 and c.config.gpt.num_audio_tokens == 626 and not c.has_loaded()
