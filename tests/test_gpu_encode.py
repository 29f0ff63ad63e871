"""DVAE encode branch on the GPU: ``ctb_dvae_encode`` through the C ABI against the oracle
(``oracle/dvae_oracle.dvae_encode``) and against the fixture made from the reference's own modules
(``tests/golden/dvae_encode.npz``, ``oracle/make_golden.py::gen_dvae_encode``).

Codes are integers: the bar is bit-exact.  Each index is a rounding of a float, so a mismatch is only tolerated where the
oracle's own decision margin (distance of the pre-rounding value to the rounding edge) is inside fp32 reordering noise
(< 1e-3); the fixture was chosen with every margin >= 9e-3, so there the comparison is exact with no exceptions.
The GPU evaluates |STFT| as a float64 DFT (correctly rounded magnitudes); the oracle is run with ``precise_stft=True``
for the live comparison, while the fixture carries the reference's fp32-FFT mel - whose rounding error the log amplifies
in the bins far below a frame's peak, hence the linear-domain mel comparison and the looser margin tolerance there."""
import os

import numpy as np
import pytest
import torch

from chattts_b200.config import Config
from chattts_b200.synth import synth_all, synth_dvae_state, synth_speech_like

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
_e = {}


def assert_mel_close(logmel_over_coef: torch.Tensor, ref: torch.Tensor, coef: torch.Tensor):
    """Compare in the LINEAR mel domain: an fp32 FFT and an fp32 DFT-as-GEMM both carry an absolute error of ~1e-6 of the
    frame's peak, which the log turns into a large difference wherever a mel bin is far below that peak."""
    a = torch.exp(logmel_over_coef.double().cpu() * coef.double().view(-1, 1))
    b = torch.exp(ref.double().cpu() * coef.double().view(-1, 1))
    tol = 2e-5 * b.amax(dim=0, keepdim=True) + 2e-6
    assert bool(((a - b).abs() <= tol).all()), float(((a - b).abs() / tol).max())
    big = b > 1e-2 * b.amax(dim=0, keepdim=True)            # bins that matter: log-domain agreement
    assert float((logmel_over_coef.cpu() - ref.cpu()).abs()[big].max()) < 2e-3



def _state():
    cfg = Config()
    return synth_dvae_state(3, cfg.dvae.decoder, cfg.dvae.decoder.idim, cfg.dvae.vq, encoder=cfg.dvae.encoder)


def encoder():
    if not _e:
        from chattts_b200.decoder import AudioEncoder, pack_dvae_encoder

        cfg = Config()
        st = _state()
        _e["st"] = st
        _e["enc"] = AudioEncoder(cfg.dvae.encoder, cfg.dvae.decoder.idim, cfg.dvae.vq,
                                 pack_dvae_encoder(st, cfg.dvae.encoder, cfg.dvae.decoder.idim, cfg.dvae.vq), "cuda",
                                 max_samples=24000 * 8)
    return _e["enc"], _e["st"]


def test_encode_matches_the_reference_generated_fixture_exactly():
    enc, _ = encoder()
    z = np.load(os.path.join(GOLDEN, "dvae_encode.npz"))
    wav = synth_speech_like(float(z["seconds"]), int(z["seed"]))
    ids, mel, margin = enc.encode(wav, want_mel=True, want_margin=True)
    assert tuple(ids.shape) == tuple(z["ids"].shape[1:]) and ids.dtype == torch.int32
    assert float(z["margin"].min()) > 5e-3
    assert np.array_equal(ids.cpu().numpy(), z["ids"][0])
    assert np.abs(margin.cpu().numpy() - z["margin"][0]).max() < 2e-2      # fp32-FFT noise of the reference mel
    assert_mel_close(mel, torch.from_numpy(z["mel_over_coef"]), _state()["coef"].reshape(-1))


@pytest.mark.parametrize("seconds,seed", [(0.55, 2), (2.0, 1), (3.013, 5)])
def test_encode_matches_oracle_on_other_lengths(seconds, seed):
    from oracle.dvae_oracle import dvae_encode

    enc, st = encoder()
    wav = synth_speech_like(seconds, seed)
    ids, mel, margin = enc.encode(wav, want_mel=True, want_margin=True)
    ref_ids, ref_margin, ref_mel, _ = dvae_encode(wav, st, return_parts=True, precise_stft=True)
    F = wav.numel() // 256 + 1
    assert tuple(ids.shape) == (4, F // 2) == tuple(ref_ids.shape[1:])
    same = ids.cpu() == ref_ids[0].int()
    assert bool(same[ref_margin[0] > 1e-3].all()), "an index with a clear decision margin differs from the oracle"
    assert float(same.float().mean()) > 0.99
    assert float((margin.cpu() - ref_margin[0]).abs().max()) < 2e-3
    assert_mel_close(mel, ref_mel, st["coef"].reshape(-1))
    assert float((mel.cpu() - ref_mel).abs().max()) < 2e-3        # against the precise oracle the LOG mel agrees everywhere


def test_fma_twin_gives_the_same_codes(monkeypatch):
    from chattts_b200.decoder import AudioEncoder, pack_dvae_encoder

    enc, st = encoder()
    cfg = Config()
    monkeypatch.setenv("CTB_DECODER_FMA", "1")
    fma = AudioEncoder(cfg.dvae.encoder, cfg.dvae.decoder.idim, cfg.dvae.vq,
                       pack_dvae_encoder(st, cfg.dvae.encoder, cfg.dvae.decoder.idim, cfg.dvae.vq), "cuda", max_samples=24000 * 4)
    monkeypatch.delenv("CTB_DECODER_FMA")
    wav = synth_speech_like(1.37, 7)
    a, ma = enc.encode(wav, want_margin=True)
    b = fma.encode(wav)
    assert bool((a == b)[ma > 1e-3].all()) and float((a == b).float().mean()) > 0.99


def test_encode_rejects_bad_sizes():
    from chattts_b200._lib import CtbError

    enc, _ = encoder()
    with pytest.raises(CtbError):
        enc.encode(torch.zeros(512))                 # reflect padding needs more than n_fft / 2 samples
    with pytest.raises(CtbError):
        enc.encode(torch.zeros(24000 * 8 + 1))       # beyond max_samples


def test_chat_samples_a_speaker_from_audio_and_uses_it_for_every_sentence():
    """core.py:179-180,435-453: sample_audio_speaker round trip, and the automatic speaker sample of multi-sentence infer()."""
    from chattts_b200 import Chat
    from chattts_b200.speaker import Speaker
    from stubs import StubSpeaker, StubTokenizer

    c = Chat()
    assert c.load_states(synth_all(0), tokenizer=StubTokenizer(), speaker=StubSpeaker(), device="cuda", max_batch=4,
                         max_context=512)
    wav = synth_speech_like(1.0, 3).numpy()
    s = c.sample_audio_speaker(wav)
    codes = Speaker.decode_prompt(s)
    assert tuple(codes.shape) == (4, (len(wav) // 256 + 1) // 2) and int(codes.min()) >= 0 and int(codes.max()) < 625
    assert torch.equal(codes, c.dvae.sample_audio(torch.from_numpy(wav)).cpu())
    # the codes are a valid input of the decode branch (dvae.py:276-297)
    assert tuple(c.dvae(codes[None].cuda()).shape) == (1, 100, 2 * codes.shape[1])

    p = c.InferCodeParams(manual_seed=3, max_new_token=24, min_new_token=24, show_tqdm=False)
    out = c.infer("first sentence here. second one. and a third.", skip_refine_text=True, params_infer_code=p)
    assert len(out) == 1 and out[0].ndim == 1 and out[0].size > 0 and np.isfinite(out[0]).all()
    assert p.spk_smp is not None and p.txt_smp == "first sentence here. "     # the sampled prompt is kept on the params
    prompt = Speaker.decode_prompt(p.spk_smp)
    assert prompt.shape[0] == 4 and prompt.shape[1] >= 1
    # an explicit speaker sample switches the automatic one off and conditions the generation
    q = c.InferCodeParams(manual_seed=3, max_new_token=24, min_new_token=24, show_tqdm=False, spk_smp=s, txt_smp="abc")
    out2 = c.infer(["hello there", "hi"], skip_refine_text=True, params_infer_code=q)
    assert q.spk_smp == s and len(out2) == 1 and out2[0].size > 0
