"""Stand-ins for the out-of-scope host components (tokenizer / speaker) used by the API tests:
same method names and tensor formats as the reference (tokenizer.py:35-138, speaker.py:54-87)."""
import torch


class StubTokenizer:
    len = 21178
    break_0_ids = 21150
    eos_token = 21001
    spk_emb_ids = 21143

    def encode(self, text, num_vq, prompt=None, device="cpu"):
        rows = [[(ord(c) * 37 + 11) % 20000 + 1 for c in t] or [1] for t in text]
        T = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), T, num_vq, dtype=torch.long)
        mask = torch.zeros(len(rows), T, dtype=torch.bool)
        for b, r in enumerate(rows):  # left padding (tokenizer.py:79-103)
            ids[b, T - len(r):] = torch.tensor(r)[:, None]
            mask[b, T - len(r):] = True
        return ids, mask, mask.clone()

    def decode(self, tokens):
        return ["".join(chr(97 + int(t) % 26) for t in row) for row in tokens]


class StubSpeaker:
    def decorate_code_prompts(self, text, prompt, txt_smp, spk_emb):
        return [f"{prompt}{t}" for t in text]

    def decorate_text_prompts(self, text, prompt):
        return [f"{t}{prompt}" for t in text]

    def decode_prompt(self, s):
        raise NotImplementedError

    def sample_random(self):
        return "stub"
