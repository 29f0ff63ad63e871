"""chattts_b200.tokenizer.Tokenizer against the reference's Tokenizer on a small BERT vocabulary written on the fly."""
import os

import pytest
import torch

SPECIAL = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "[Stts]", "[Ptts]", "[spk_emb]", "[empty_spk]", "[Sbreak]",
           "[Pbreak]", "[Ebreak]", "[break_0]", "[uv_break]", "[speed_5]", "[oral_2]"]
WORDS = ["hello", "there", "world", "hi", "a", "b", "test", "##ing", "speech", ".", ","]


def _write_vocab(tmp_path):
    from transformers import BertTokenizerFast

    tok = BertTokenizerFast(vocab={w: i for i, w in enumerate(SPECIAL + WORDS)}, do_lower_case=True)
    tok.add_special_tokens({"additional_special_tokens": SPECIAL[5:]})
    out = tmp_path / "tok"
    tok.save_pretrained(str(out))
    return str(out)


def test_layout_left_padding_and_audio_prompt(tmp_path):
    from chattts_b200.tokenizer import Tokenizer

    t = Tokenizer(_write_vocab(tmp_path))
    assert t.len == len(SPECIAL) + len(WORDS)
    assert (t.spk_emb_ids, t.break_0_ids, t.eos_token) == (7, 12, 11)
    ids, att, tm = t.encode(["[Stts][spk_emb]hello there world[Ptts]", "[Stts][empty_spk]hi[Ptts]"], 4)
    assert ids.shape == (2, 6, 4) and att.shape == tm.shape == (2, 6)
    assert att.tolist() == [[1] * 6, [0, 0, 1, 1, 1, 1]] and torch.equal(tm, att.bool())
    assert ids[0, :, 0].tolist() == [5, 7, 16, 17, 18, 6] and bool((ids == ids[:, :, :1]).all())
    prompt = torch.arange(12).view(4, 3)
    ids2, att2, tm2 = t.encode(["hello", "hi there"], 4, prompt=prompt)
    assert ids2.shape == (2, 5, 4)
    assert att2.tolist() == [[0, 1, 1, 1, 1], [1, 1, 1, 1, 1]]
    assert tm2.tolist() == [[False, True, False, False, False], [True, True, False, False, False]]
    assert torch.equal(ids2[0, 2:], prompt.t()) and torch.equal(ids2[1, 2:], prompt.t())
    assert t.decode(ids[:, :, 0])[1].replace(" ", "").endswith("[Stts][empty_spk]hi[Ptts]")


@pytest.mark.reference
def test_matches_reference_tokenizer(tmp_path):
    from oracle.ref_import import load_reference, reference_available

    if not reference_available():
        pytest.skip("/root/reference not present on this box")
    load_reference()
    from ChatTTS.model.tokenizer import Tokenizer as RefTokenizer

    from chattts_b200.tokenizer import Tokenizer

    path = _write_vocab(tmp_path)
    ours, ref = Tokenizer(path), RefTokenizer(path)
    if not hasattr(ref._tokenizer, "encode_plus"):       # API drift: transformers >= 5 removed encode_plus (same as __call__)
        ref._tokenizer.encode_plus = ref._tokenizer.__call__
    assert (ours.len, ours.spk_emb_ids, ours.break_0_ids, ours.eos_token) == (ref.len, ref.spk_emb_ids, ref.break_0_ids, ref.eos_token)
    texts = ["[Stts][spk_emb]hello there world[Ptts]", "[Stts][empty_spk]hi[Ptts]", "testing speech, a b."]
    for prompt in (None, torch.randint(0, 626, (4, 7))):
        a = ours.encode(list(texts), 4, prompt=prompt)
        b = ref.encode(list(texts), 4, prompt=None if prompt is None else prompt.clone())
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)
    seq = [[16, 17, 11], [19, 12]]
    assert ours.decode(seq) == ref.decode(seq)
