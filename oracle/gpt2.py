"""oracle.gpt2 — the reference's CPU-runnable configuration (BASELINE.json configs[0]).

TEST INFRASTRUCTURE (see oracle/__init__.py).  `examples/cpu_offline_inference.py`
runs gpt2-124M on CPU; the engine forces fp32 there (src/engine/llm_engine.cpp:28-31)
and uses RefHandler + F::linear + LayerNorm + gelu_new (no custom kernels).  The wheel
cannot be built in this image, so this is a torch-CPU restatement of
src/models/openai/gpt2.h:33-305 with random-init weights (no HF weights offline),
greedy decoding through a per-layer KV cache.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List

import torch
import torch.nn.functional as F


def gelu_new(x: torch.Tensor) -> torch.Tensor:  # src/layers/activation.cpp gelu_new
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def init_gpt2(seed: int = 0, n_layers=12, hidden=768, n_heads=12, vocab=50257, n_pos=1024) -> Dict:
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g) * 0.02
    layers = []
    for _ in range(n_layers):
        layers.append(dict(ln1_w=torch.ones(hidden), ln1_b=torch.zeros(hidden),
                           attn_w=rn(3 * hidden, hidden), attn_b=torch.zeros(3 * hidden),
                           proj_w=rn(hidden, hidden), proj_b=torch.zeros(hidden),
                           ln2_w=torch.ones(hidden), ln2_b=torch.zeros(hidden),
                           fc_w=rn(4 * hidden, hidden), fc_b=torch.zeros(4 * hidden),
                           fc2_w=rn(hidden, 4 * hidden), fc2_b=torch.zeros(hidden)))
    return dict(wte=rn(vocab, hidden), wpe=rn(n_pos, hidden), layers=layers,
                lnf_w=torch.ones(hidden), lnf_b=torch.zeros(hidden), n_heads=n_heads,
                hidden=hidden)


def _forward(m: Dict, tokens: torch.Tensor, pos0: int, caches: List) -> torch.Tensor:
    hidden, H = m["hidden"], m["n_heads"]
    D = hidden // H
    T = tokens.numel()
    h = m["wte"][tokens] + m["wpe"][pos0: pos0 + T]          # gpt2.h:232
    for li, L in enumerate(m["layers"]):
        x = F.layer_norm(h, (hidden,), L["ln1_w"], L["ln1_b"], 1e-5)
        qkv = F.linear(x, L["attn_w"], L["attn_b"])
        q, k, v = qkv.chunk(3, dim=-1)                       # gpt2.h:122
        q, k, v = (t.view(T, H, D) for t in (q, k, v))
        kc, vc = caches[li]
        kc = torch.cat([kc, k], 0)
        vc = torch.cat([vc, v], 0)
        caches[li] = (kc, vc)
        S = kc.shape[0]
        scores = torch.einsum("qhd,khd->hqk", q, kc) * (D ** -0.5)
        mask = torch.tril(torch.ones(T, S, dtype=torch.bool), diagonal=S - T)
        scores = scores.masked_fill(~mask, float("-inf"))
        attn = torch.einsum("hqk,khd->qhd", torch.softmax(scores, -1), vc).reshape(T, hidden)
        h = h + F.linear(attn, L["proj_w"], L["proj_b"])
        x = F.layer_norm(h, (hidden,), L["ln2_w"], L["ln2_b"], 1e-5)
        h = h + F.linear(gelu_new(F.linear(x, L["fc_w"], L["fc_b"])), L["fc2_w"], L["fc2_b"])
    h = F.layer_norm(h, (hidden,), m["lnf_w"], m["lnf_b"], 1e-5)
    return F.linear(h[-1:], m["wte"])                        # tied lm_head, gpt2.h:272


@torch.no_grad()
def generate(m: Dict, prompt: torch.Tensor, new_tokens: int = 32) -> Dict:
    """Greedy offline generate, batch 1.  Returns tokens + wall-clock timings."""
    D = m["hidden"] // m["n_heads"]
    caches = [(torch.zeros(0, m["n_heads"], D), torch.zeros(0, m["n_heads"], D))
              for _ in m["layers"]]
    out = []
    t0 = time.perf_counter()
    logits = _forward(m, prompt, 0, caches)
    t_first = time.perf_counter() - t0
    nxt = logits.argmax(-1)
    out.append(int(nxt))
    pos = prompt.numel()
    t1 = time.perf_counter()
    for _ in range(new_tokens - 1):
        logits = _forward(m, nxt, pos, caches)
        nxt = logits.argmax(-1)
        out.append(int(nxt))
        pos += 1
    t_decode = time.perf_counter() - t1
    return dict(tokens=out, ttft_s=t_first, decode_s=t_decode,
                decode_tok_s=(new_tokens - 1) / max(t_decode, 1e-9))
