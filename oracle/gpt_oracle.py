"""CPU (torch fp32) restatement of hot path 1: the ChatTTS GPT decode loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines
it restates; ``[3p]`` marks HF ``transformers`` code reached from the reference call site
``ChatTTS/model/gpt.py:419-427`` whose in-tree statement is
``examples/onnx/modeling_llama.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- sampler
def apply_temperature(logits: torch.Tensor, temperature: torch.Tensor) -> torch.Tensor:
    """gpt.py:350-355,487 - rows are (b, q) pairs; row r uses temperature[r % len]."""
    rows = logits.shape[0]
    t = temperature.reshape(1, -1).expand(rows // temperature.numel(), -1).reshape(-1, 1)
    return logits / t


def repetition_penalty(
    ids: torch.Tensor, scores: torch.Tensor, penalty: float, max_input_ids: int, past_window: int = 16
) -> torch.Tensor:
    """processors.py:18-35 (window, one-hot count, the row>=max_input_ids drop-out quirk,
    ``penalty**count``; negative scores multiplied, others divided)."""
    if ids.size(1) > past_window:
        ids = ids[:, -past_window:]
    freq = F.one_hot(ids, scores.size(1)).sum(1)
    if freq.size(0) > max_input_ids:
        freq[max_input_ids:] = 0
    alpha = torch.pow(penalty, freq)
    return torch.where(scores < 0, scores * alpha, scores / alpha)


def top_p_filter(scores: torch.Tensor, top_p: float, min_keep: int = 3) -> torch.Tensor:
    """[3p] HF TopPLogitsWarper built at processors.py:44-45: ascending sort, softmax of the
    sorted row, cumsum, remove ``cum <= 1 - top_p`` except the last ``min_keep``."""
    srt, idx = torch.sort(scores, descending=False)
    cum = srt.softmax(dim=-1).cumsum(dim=-1)
    remove = cum <= (1 - top_p)
    remove[..., -min_keep:] = False
    remove = remove.scatter(1, idx, remove)
    return scores.masked_fill(remove, -float("inf"))


def top_k_filter(scores: torch.Tensor, top_k: int, min_keep: int = 3) -> torch.Tensor:
    """[3p] HF TopKLogitsWarper built at processors.py:46-48: k=max(top_k,min_keep); remove
    ``score < k-th largest``."""
    k = min(max(top_k, min_keep), scores.size(-1))
    kth = torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(scores < kth, -float("inf"))


def exp_noise(rows: int, cols: int, seed: int) -> torch.Tensor:
    """The Exp(1) tensor ``torch.multinomial`` draws (ATen multinomial fast path:
    ``q = empty_like(p).exponential_(1, gen)``).  gpt.py:504-508 re-seeds the generator on
    every step, so this tensor is the same at every step of one ``generate`` call."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.empty(rows, cols, dtype=torch.float32).exponential_(1, generator=g)


def decision_margins(logits: torch.Tensor, generated: torch.Tensor, temperature: torch.Tensor, sp: "SamplerParams",
                     q: torch.Tensor, eos: int, ban_eos: bool):
    """How close one sampling pass is to flipping under fp32 reorder noise (the 'margin monitor' of SURVEY.md 7):
    returns (relative gap between the two largest p/q values, distance of the top-p cumulative sum from its
    threshold at the cut), both minimised over the rows.  Used by the tests to show the committed fixtures sit far
    (>= 1e-4) from any tie, so ~1e-6 differences in logits cannot change an id."""
    x = apply_temperature(logits, temperature)
    if sp.repetition_penalty is not None and sp.repetition_penalty != 1:
        x = repetition_penalty(generated, x, sp.repetition_penalty, sp.penalty_max_ids, sp.penalty_window)
    p_margin = float("inf")
    if sp.top_p is not None:
        srt, _ = torch.sort(x, descending=False)
        cum = srt.softmax(dim=-1).cumsum(dim=-1)
        thr = 1 - sp.top_p
        p_margin = float((cum[..., :-sp.min_keep] - thr).abs().min())
        x = top_p_filter(x, sp.top_p, sp.min_keep)
    if sp.top_k is not None:
        x = top_k_filter(x, sp.top_k, sp.min_keep)
    if sp.greedy:
        if sp.greedy_exclude_eos:
            x = x.clone()
            x[:, eos] = -float("inf")
        x = x.masked_fill(x < x.max(dim=-1, keepdim=True)[0], -float("inf"))
    if ban_eos:
        x = x.clone()
        x[:, eos] = -float("inf")
    r = F.softmax(x, dim=-1) / q
    top2 = torch.topk(r, 2, dim=-1)[0]
    a_margin = float(((top2[:, 0] - top2[:, 1]) / top2[:, 0]).min())
    return a_margin, p_margin


def sample_from_scores(scores: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """``torch.multinomial(scores, 1, generator)`` == ``argmax(scores / q)`` (ATen fast path)."""
    return torch.argmax(scores / q, dim=-1)


@dataclass
class SamplerParams:
    top_p: Optional[float] = 0.7
    top_k: Optional[int] = 20
    repetition_penalty: float = 1.05
    penalty_max_ids: int = 625  # gen_logits(num_code=...) -> max_input_ids (processors.py:53-55)
    penalty_window: int = 16
    min_keep: int = 3
    greedy: bool = False  # bench config C2: extra arg-max mask processor
    greedy_exclude_eos: bool = False


def sample_step(
    logits: torch.Tensor,
    generated: torch.Tensor,
    temperature: torch.Tensor,
    sp: SamplerParams,
    q: torch.Tensor,
    eos: int,
    ban_eos: bool,
) -> torch.Tensor:
    """One pass of gpt.py:487-508 over rows ``[(B*num_vq) or B, V]``.

    Order (core.py:649): temperature -> repetition penalty -> top-p -> top-k ->
    (min_new_token EOS ban, gpt.py:494-495) -> softmax -> multinomial."""
    logits = apply_temperature(logits, temperature)
    if sp.repetition_penalty is not None and sp.repetition_penalty != 1:
        logits = repetition_penalty(generated, logits, sp.repetition_penalty, sp.penalty_max_ids, sp.penalty_window)
    if sp.top_p is not None:
        logits = top_p_filter(logits, sp.top_p, sp.min_keep)
    if sp.top_k is not None:
        logits = top_k_filter(logits, sp.top_k, sp.min_keep)
    if sp.greedy:
        if sp.greedy_exclude_eos:
            logits = logits.clone()
            logits[:, eos] = -float("inf")
        logits = logits.masked_fill(logits < logits.max(dim=-1, keepdim=True)[0], -float("inf"))
    if ban_eos:
        logits = logits.clone()
        logits[:, eos] = -float("inf")
    scores = F.softmax(logits, dim=-1)
    return sample_from_scores(scores, q)


# ----------------------------------------------------------------------------- model
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """[3p] LlamaRMSNorm; in-tree: examples/onnx/modeling_llama.py:102-116."""
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float):
    """[3p] LlamaRotaryEmbedding; in-tree: examples/onnx/modeling_llama.py:119-162.
    ``positions`` float [B,t] (the reference casts them to float: gpt.py:151-159,417)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = (inv_freq[None, :, None].expand(positions.shape[0], -1, 1) @ positions[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """examples/onnx/modeling_llama.py:239-244."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """embed.py:23-35 weight_norm(dim=0): W = g * v / ||v||_row."""
    return g * v / v.norm(dim=1, keepdim=True)


@dataclass
class GenerationOutputs:
    """gpt.py:276-285."""

    ids: List[torch.Tensor]
    hiddens: List[torch.Tensor]
    attentions: list = field(default_factory=list)
    steps: int = 0
    trace: Optional[dict] = None


class GPTOracle:
    def __init__(self, gpt_state: State, embed_state: State, *, num_heads=12, head_dim=64, eps=1e-6,
                 theta=10000.0, num_vq=4):
        self.s = gpt_state
        self.e = embed_state
        self.L = 1 + max(int(k.split(".")[1]) for k in gpt_state if k.startswith("layers."))
        self.H, self.hd, self.eps, self.theta, self.num_vq = num_heads, head_dim, eps, theta, num_vq
        self.head_code = [
            fold_weight_norm(embed_state[f"head_code.{q}.parametrizations.weight.original0"],
                             embed_state[f"head_code.{q}.parametrizations.weight.original1"])
            for q in range(num_vq)
        ]
        self.head_text = fold_weight_norm(embed_state["head_text.parametrizations.weight.original0"],
                                          embed_state["head_text.parametrizations.weight.original1"])

    # embed.py:51-79
    def embed_prompt(self, input_ids: torch.Tensor, text_mask: torch.Tensor) -> torch.Tensor:
        B, T, _ = input_ids.shape
        emb = torch.zeros(B, T, self.e["emb_text.weight"].shape[1])
        emb[text_mask] = F.embedding(input_ids[text_mask][:, 0], self.e["emb_text.weight"])
        code_ids = input_ids[~text_mask]
        emb[~text_mask] = sum(F.embedding(code_ids[:, q], self.e[f"emb_code.{q}.weight"]) for q in range(self.num_vq))
        return emb

    # gpt.py:403-415
    def embed_step(self, ids: torch.Tensor, infer_text: bool) -> torch.Tensor:
        if infer_text:
            return F.embedding(ids[:, :, 0], self.e["emb_text.weight"])
        return torch.stack([F.embedding(ids[:, :, q], self.e[f"emb_code.{q}.weight"]) for q in range(self.num_vq)], 3).sum(3)

    # [3p] LlamaModel.forward; in-tree: examples/onnx/modeling_llama.py:375-505,519-579
    def forward(self, x, positions, key_mask, past):
        """x [B,t,d]; positions float [B,t]; key_mask bool [B,Ttot] (True = attend);
        past: list of (k,v) or None.  Returns (normed last hidden [B,t,d], new past)."""
        B, t, d = x.shape
        cos, sin = rope_cos_sin(positions, self.hd, self.theta)
        cos, sin = cos[:, None], sin[:, None]
        Ttot = key_mask.shape[1]
        causal = torch.ones(t, Ttot, dtype=torch.bool).tril(diagonal=Ttot - t)
        allow = causal[None, None] & key_mask[:, None, None, :]
        add = torch.zeros(B, 1, t, Ttot).masked_fill(~allow, -float("inf"))
        new_past = []
        s = self.s
        for l in range(self.L):
            p = f"layers.{l}."
            h = rms_norm(x, s[p + "input_layernorm.weight"], self.eps)
            q = F.linear(h, s[p + "self_attn.q_proj.weight"]).view(B, t, self.H, self.hd).transpose(1, 2)
            k = F.linear(h, s[p + "self_attn.k_proj.weight"]).view(B, t, self.H, self.hd).transpose(1, 2)
            v = F.linear(h, s[p + "self_attn.v_proj.weight"]).view(B, t, self.H, self.hd).transpose(1, 2)
            q = q * cos + rotate_half(q) * sin
            k = k * cos + rotate_half(k) * sin
            if past is not None:
                k = torch.cat([past[l][0], k], dim=2)
                v = torch.cat([past[l][1], v], dim=2)
            new_past.append((k, v))
            w = torch.matmul(q, k.transpose(2, 3)) * (self.hd ** -0.5) + add
            w = torch.softmax(w, dim=-1, dtype=torch.float32)
            w = torch.nan_to_num(w)  # fully-masked (pad) query rows; their outputs are never read
            a = torch.matmul(w, v).transpose(1, 2).reshape(B, t, self.H * self.hd)
            x = x + F.linear(a, s[p + "self_attn.o_proj.weight"])
            h = rms_norm(x, s[p + "post_attention_layernorm.weight"], self.eps)
            m = F.silu(F.linear(h, s[p + "mlp.gate_proj.weight"])) * F.linear(h, s[p + "mlp.up_proj.weight"])
            x = x + F.linear(m, s[p + "mlp.down_proj.weight"])
        return rms_norm(x, s["norm.weight"], self.eps), new_past

    # gpt.py:438-464
    def logits_rows(self, hidden_last: torch.Tensor, infer_text: bool) -> torch.Tensor:
        if infer_text:
            return F.linear(hidden_last, self.head_text)
        lg = torch.stack([F.linear(hidden_last, w) for w in self.head_code], dim=2)  # [B,V,4]
        return lg.permute(0, 2, 1).reshape(-1, lg.size(1))  # rows (b,q)

    @torch.no_grad()
    def generate(self, emb, inputs_ids, temperature, eos_token, attention_mask=None, max_new_token=2048,
                 min_new_token=0, sampler: SamplerParams = SamplerParams(), infer_text=False,
                 return_hidden=False, manual_seed: Optional[int] = None, trace=False,
                 forced_ids: Optional[torch.Tensor] = None) -> GenerationOutputs:
        """gpt.py:315-618 (non-stream, seeded).  ``forced_ids`` [B,n,4] teacher-forces the
        appended tokens (used by margin tests); sampled ids are still recorded in the trace."""
        B, T0, nvq = inputs_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones(B, T0, dtype=torch.bool)
        mask = torch.ones(B, T0 + max_new_token, dtype=torch.bool)
        mask[:, :T0] = attention_mask.bool()
        ids_buf = torch.zeros(B, T0 + max_new_token, nvq, dtype=torch.long)
        ids_buf[:, :T0] = inputs_ids
        finish = torch.zeros(B, dtype=torch.bool)
        end_idx = torch.zeros(B, dtype=torch.long)
        rows = B if infer_text else B * nvq
        V = self.head_text.shape[0] if infer_text else self.head_code[0].shape[0]
        assert manual_seed is not None, "oracle covers the seeded path (unseeded has no parity target)"
        q = exp_noise(rows, V, manual_seed)
        past, hiddens = None, []
        tr = {"logits": [], "sampled": [], "argmax_margin": [], "top_p_margin": []} if trace else None
        progress = T0
        steps = 0
        for i in range(max_new_token):
            m = mask[:, :progress]
            pos = (m.long().cumsum(-1) - 1).masked_fill(~m, 1)  # gpt.py:234-241
            if i == 0:
                x, positions = emb, pos
            else:
                x, positions = self.embed_step(ids_buf[:, progress - 1: progress], infer_text), pos[:, -1:]
            hidden, past = self.forward(x.float(), positions.float(), m, past)
            last = hidden[:, -1]
            if return_hidden:
                hiddens.append(last)
            logits = self.logits_rows(last, infer_text)
            if trace:
                tr["logits"].append(logits.clone())
            gen = ids_buf[:, T0:progress]
            gen_rows = gen[:, :, 0] if infer_text else gen.permute(0, 2, 1).reshape(rows, -1)
            idx = sample_step(logits, gen_rows, temperature, sampler, q, eos_token, i < min_new_token)
            idx = idx.view(B, -1)
            if trace:
                tr["sampled"].append(idx.clone())
                am, pm = decision_margins(logits, gen_rows, temperature, sampler, q, eos_token, i < min_new_token)
                tr["argmax_margin"].append(am)
                tr["top_p_margin"].append(pm)
            finish |= (idx == eos_token).any(1)
            app = idx if forced_ids is None else forced_ids[:, i]
            ids_buf[:, progress] = app if not infer_text else app[:, :1].expand(-1, nvq)
            steps += 1
            if i == 0 and finish.any():
                # gpt.py:527-570: seeded => warn and stop without yielding anything
                return GenerationOutputs(ids=[], hiddens=[], steps=steps, trace=tr)
            progress += 1
            end_idx += (~finish).long()
            if finish.all():
                break
        ids = [ids_buf[b, T0: T0 + int(end_idx[b])] for b in range(B)]
        if infer_text:
            ids = [t[:, 0] for t in ids]
        hs = []
        if return_hidden:
            hst = torch.stack(hiddens, 1)
            hs = [hst[b, : int(end_idx[b])] for b in range(B)]
        return GenerationOutputs(ids=ids, hiddens=hs, steps=steps, trace=tr)
