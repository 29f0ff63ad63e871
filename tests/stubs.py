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
        P = 0 if prompt is None else int(prompt.size(1))
        T = max(len(r) for r in rows) + P
        ids = torch.zeros(len(rows), T, num_vq, dtype=torch.long)
        mask = torch.zeros(len(rows), T, dtype=torch.bool)
        for b, r in enumerate(rows):  # left padding (tokenizer.py:79-103); the audio prompt follows the text (:120-133)
            ids[b, T - P - len(r): T - P] = torch.tensor(r)[:, None]
            mask[b, T - P - len(r):] = True
        text_mask = mask.clone()
        if P:
            assert prompt.size(0) == num_vq
            ids[:, T - P:] = prompt.t().long()[None]
            text_mask[:, T - P:] = False
        return ids, mask, text_mask

    def decode(self, tokens):
        return ["".join(chr(97 + int(t) % 26) for t in row) for row in tokens]


class StubSpeaker:
    def decorate_code_prompts(self, text, prompt, txt_smp, spk_emb):
        return [f"{prompt}{t}" for t in text]

    def decorate_text_prompts(self, text, prompt):
        return [f"{t}{prompt}" for t in text]

    def decode_prompt(self, s):
        from chattts_b200.speaker import Speaker

        return Speaker.decode_prompt(s)

    def encode_prompt(self, codes):
        from chattts_b200.speaker import Speaker

        return Speaker.encode_prompt(codes)

    def sample_random(self):
        return "stub"
