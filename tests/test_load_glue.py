"""``Chat.load(source="custom")`` glue on CPU: asset discovery, safetensors / LlamaModel loading, key filtering, this package's
own Tokenizer / Speaker / Normalizer - everything up to ``load_states`` (which needs the GPU and is replaced by a recorder)."""
import numpy as np
import torch


def _make_assets(root):
    from safetensors.torch import save_file
    from transformers import BertTokenizerFast, LlamaConfig, LlamaModel

    asset = root / "asset"
    asset.mkdir()
    cfg = LlamaConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=16, max_position_embeddings=64)
    LlamaModel(cfg).save_pretrained(str(asset / "gpt"))
    for name in ("Vocos", "DVAE", "Decoder", "Embed"):
        save_file({"w": torch.full((2,), float(len(name)))}, str(asset / f"{name}.safetensors"))
    tok = BertTokenizerFast(vocab={w: i for i, w in enumerate(
        ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "[spk_emb]", "[break_0]", "[Ebreak]", "hi"])})
    tok.add_special_tokens({"additional_special_tokens": ["[spk_emb]", "[break_0]", "[Ebreak]"]})
    tok.save_pretrained(str(asset / "tokenizer"))


def test_load_custom_assets_reaches_load_states_with_own_host_components(tmp_path, monkeypatch):
    from chattts_b200 import Chat, b14
    from chattts_b200.norm import Normalizer
    from chattts_b200.speaker import Speaker
    from chattts_b200.tokenizer import Tokenizer

    c = Chat()
    assert c.load(source="custom", custom_path=str(tmp_path)) is False          # nothing there yet
    _make_assets(tmp_path)
    seen = {}

    def fake_load_states(states, tokenizer, speaker, device=None, coef=None, **kw):
        seen.update(states=states, tokenizer=tokenizer, speaker=speaker, device=device, coef=coef)
        return True

    monkeypatch.setattr(c, "load_states", fake_load_states)
    stat = b14.encode_to_string(np.concatenate([np.ones(768, np.float16), np.zeros(768, np.float16)]).tobytes())
    assert c.load(source="custom", custom_path=str(tmp_path), device=torch.device("cpu"), spk_stat=stat) is True
    st = seen["states"]
    assert set(st) == {"gpt", "embed", "decoder", "dvae", "vocos"}
    assert "layers.0.self_attn.q_proj.weight" in st["gpt"] and "norm.weight" in st["gpt"]
    assert not any(k.startswith("embed_tokens") for k in st["gpt"])
    assert float(st["vocos"]["w"][0]) == 5.0 and float(st["decoder"]["w"][0]) == 7.0
    assert isinstance(seen["tokenizer"], Tokenizer) and seen["tokenizer"].spk_emb_ids == 5
    assert isinstance(seen["speaker"], Speaker) and seen["speaker"].sample_random().startswith("蘁淰")
    assert isinstance(c.normalizer, Normalizer)
