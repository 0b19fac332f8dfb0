"""Scaled-down Llama (random init) built twice from the same checkpoint tensors: the CPU oracle
(dequantised dense weights, oracle.llama) and the GPU plugin layers (scalellm_b200).  Shared by
tests/test_gpu_decode_step.py and __graft_entry__.smoke()."""
import torch

from oracle import llama as ollama, ops, quant
from scalellm_b200.decode_step import LlamaArgs, LlamaDecoder
from scalellm_b200.layers import QuantArgs
from scalellm_b200.model_parallel import ParallelArgs

DEV = torch.device("cuda")


def small_args():
    return LlamaArgs(hidden_size=512, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64,
                     intermediate_size=1024, vocab_size=2048, max_position_embeddings=512)


def build_pair(method, seed=0, args=None):
    """Same random checkpoint on both sides: CPU oracle (dequantised dense) and GPU plugin layers."""
    a = args or small_args()
    qa = QuantArgs(quant_method=method, bits=4, group_size=128, is_sym=(method == "gptq"))
    model = LlamaDecoder(a, qa, ParallelArgs(), DEV)
    cfg = ollama.LlamaConfig(hidden=a.hidden_size, n_layers=a.n_layers, n_heads=a.n_heads,
                             n_kv_heads=a.n_kv_heads, head_dim=a.head_dim, inter=a.intermediate_size,
                             vocab=a.vocab_size, max_pos=a.max_position_embeddings)
    g = torch.Generator().manual_seed(seed)
    h, D = a.hidden_size, a.head_dim
    shapes = dict(qkv=(h, (a.n_heads + 2 * a.n_kv_heads) * D), o=(a.n_heads * D, h),
                  gate_up=(h, 2 * a.intermediate_size), down=(a.intermediate_size, h))
    olayers = []
    for i in range(a.n_layers):
        sd, ol = {}, {}
        for j, (name, (K, N)) in enumerate(shapes.items()):
            if method == "none":
                w = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
                sd[name] = dict(weight=w)
                ol[name] = ollama.Linear(w.t().contiguous())
            else:
                mk = quant.random_awq_checkpoint if method == "awq" else quant.random_gptq_checkpoint
                ck = mk(K, N, 128, seed=seed * 100 + i * 10 + j)
                ck["scales"] = (ck["scales"].float() * 0.6).bfloat16()
                sd[name] = dict(qweight=ck["qweight"], scales=ck["scales"],
                                qzeros=ck["qzeros"] if method == "awq" else None)
                ol[name] = ollama.Linear(quant.dequant(ck["q"], ck["z"], ck["scales"], 128))
        for nm in ("input_norm", "post_norm"):
            wn = (1 + 0.1 * torch.randn(h, generator=g)).bfloat16()
            sd[nm] = wn
            ol[nm] = wn
        model.load_layer(i, sd)
        olayers.append(ol)
    embed = (torch.randn(a.vocab_size, h, generator=g) * 0.5).bfloat16()
    head = (torch.randn(a.vocab_size, h, generator=g) * 0.05).bfloat16()
    fn = (1 + 0.1 * torch.randn(h, generator=g)).bfloat16()
    model.embed.copy_(embed)
    model.lm_head.weight.copy_(head)
    model.final_norm.weight.copy_(fn)
    cos_sin = ops.build_cos_sin_cache(D, a.max_position_embeddings, ollama.inv_freq_for(cfg),
                                      torch.bfloat16)
    omodel = dict(embed=embed, layers=olayers, final_norm=fn, lm_head=head.t().contiguous(),
                  cos_sin=cos_sin)
    return a, cfg, model, omodel
