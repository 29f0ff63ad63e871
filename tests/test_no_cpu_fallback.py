"""The product path must fail loudly without a CUDA device - there is no CPU fallback (north_star)."""
import pytest
import torch

from chattts_b200 import _lib
from chattts_b200.config import Config
from chattts_b200.embed import Embed
from chattts_b200.gpt import GPT


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_loading_weights_without_a_gpu_raises():
    cfg = Config()
    embed = Embed(768, 626, 21178, 4)
    gpt = GPT(cfg.gpt, embed, device="cpu", device_gpt="cpu")
    with pytest.raises(_lib.CtbError):
        gpt.load_state({})
    from chattts_b200.decoder import TokenDecoder

    with pytest.raises(_lib.CtbError):
        TokenDecoder(cfg.decoder, 384, None, cfg.vocos, None, torch.zeros(4), "cpu")


def test_product_never_imports_the_oracle():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chattts_b200")
    for name in os.listdir(root):
        if name.endswith(".py"):
            src = open(os.path.join(root, name)).read()
            hits = [l for l in src.splitlines() if re.match(r"\s*(from|import)\s+oracle", l)]
            assert not hits, (name, hits)
            assert "oracle" not in src or name in ("__init__.py",) or "oracle" not in [w for l in src.splitlines()
                                                                                       if "import" in l for w in l.split()], name
