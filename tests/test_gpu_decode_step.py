"""GPU parity: one full Llama decode step (the §8a op sequence through the plugin layer) vs the
CPU oracle restating models/meta/llama.h, on a scaled-down Llama with random-init weights.
North-star bar: logits within 1e-3 rtol (bf16) of the reference path; KV-cache contents and
block-table indexing bit exact."""
import numpy as np
import pytest
import torch

from oracle import llama as ollama, ops, quant
from scalellm_b200.decode_step import (BlockPool, GraphedStep, LlamaArgs, LlamaDecoder,
                                       StepBuffers, build_decode_batch)
from scalellm_b200.layers import QuantArgs
from scalellm_b200.model_parallel import ParallelArgs

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


from tests.smallcase import build_pair, small_args  # noqa: E402,F401


@pytest.mark.parametrize("method", ["awq", "gptq", "none"])
@pytest.mark.parametrize("bs", [8, 16])
def test_decode_step_logits_and_cache(method, bs):
    a, cfg, model, omodel = build_pair(method)
    kv_lens, q_lens = [100, 17, 255, 64, 1], [1, 1, 1, 1, 1]
    n_blocks = 80 * (16 // bs) + 7
    pool = BlockPool(n_blocks, bs, seed=2)
    for kv in kv_lens:
        pool.add_sequence(kv + 8)
    hb = build_decode_batch(pool, kv_lens, q_lens, a.vocab_size)
    model.alloc_kv(n_blocks, bs, randomize=True, seed=1)
    ok = [c.key_cache.cpu().clone() for c in model.kv_caches]
    ov = [c.value_cache.cpu().clone() for c in model.kv_caches]
    snap_k = [c.clone() for c in ok]
    snap_v = [c.clone() for c in ov]

    bufs = StepBuffers(DEV, 64, 16, 1024)
    tokens, positions, params = bufs.upload(hb)
    logits = model(tokens, positions, params)
    torch.cuda.synchronize()

    meta = dict(q_cu_lens=hb.q_cu_lens, kv_cu_lens=hb.kv_cu_lens,
                block_table=torch.from_numpy(hb.block_tables), block_cu_lens=hb.cu_block_lens,
                block_size=bs)
    ref = ollama.decode_step(torch.from_numpy(hb.tokens), torch.from_numpy(hb.positions), omodel,
                             cfg, ok, ov, torch.from_numpy(hb.new_cache_slots), meta)
    # fp32 "truth": the same op sequence and the same (bf16-valued) weights with NO intermediate
    # rounding.  The B200 path must be as close to it as the per-op-rounded bf16 oracle is:
    # that is what "within bf16 rounding of the reference" means for a 3-layer pipeline.
    f32 = lambda t: t.float() if torch.is_tensor(t) and t.is_floating_point() else t
    omodel32 = dict(embed=f32(omodel["embed"]), final_norm=f32(omodel["final_norm"]),
                    lm_head=f32(omodel["lm_head"]), cos_sin=omodel["cos_sin"],
                    layers=[{k: (ollama.Linear(v.w.float()) if isinstance(v, ollama.Linear) else f32(v))
                             for k, v in L.items()} for L in omodel["layers"]])
    ok32 = [c.float() for c in snap_k]
    ov32 = [c.float() for c in snap_v]
    truth = ollama.decode_step(torch.from_numpy(hb.tokens), torch.from_numpy(hb.positions),
                               omodel32, cfg, ok32, ov32, torch.from_numpy(hb.new_cache_slots), meta)
    slots = torch.from_numpy(hb.new_cache_slots).long()
    for i, c in enumerate(model.kv_caches):
        kc, vc = c.key_cache.cpu(), c.value_cache.cpu()
        mask = torch.ones(kc.shape[0], dtype=torch.bool)
        mask[slots] = False
        # rows no new token maps to are untouched, bit for bit
        assert torch.equal(kc[mask], snap_k[i][mask]) and torch.equal(vc[mask], snap_v[i][mask])
        # written rows: as close to the fp32 truth as the bf16 oracle's rows
        for got, orc, tru in ((kc[slots], ok[i][slots], ok32[i][slots]),
                              (vc[slots], ov[i][slots], ov32[i][slots])):
            e_got = (got.float() - tru).abs().mean()
            e_orc = (orc.float() - tru).abs().mean()
            assert e_got <= 1.5 * e_orc + 1e-4 * tru.abs().mean(), (i, float(e_got), float(e_orc))
    lo, lr = logits.float().cpu(), ref.float()
    scale = truth.abs().mean()
    e_ours = (lo - truth).abs().mean()
    e_oracle = (lr - truth).abs().mean()
    print(f"logits: |ours-truth|={float(e_ours/scale):.3e} |oracle-truth|={float(e_oracle/scale):.3e} "
          f"|ours-oracle|={float((lo-lr).abs().mean()/scale):.3e} (relative to mean |logit|)")
    assert e_ours <= 1.5 * e_oracle + 1e-4 * scale
    assert (lo - lr).abs().mean() / scale < 2.5 * e_oracle / scale + 1e-3   # two bf16-noise realisations
    assert (lo.argmax(-1) == truth.argmax(-1)).float().mean() >= 0.8


def test_graph_replay_matches_eager_and_multi_step():
    a, cfg, model, omodel = build_pair("awq", seed=1)
    bs = 8
    kv_lens = [33, 64, 7, 120]
    pool = BlockPool(128, bs, seed=2)
    for kv in kv_lens:
        pool.add_sequence(kv + 16)
    model.alloc_kv(128, bs, randomize=True, seed=1)
    snap = [(c.key_cache.clone(), c.value_cache.clone()) for c in model.kv_caches]
    bufs = StepBuffers(DEV, 16, 8, 256)
    hb = build_decode_batch(pool, kv_lens, [1] * 4, a.vocab_size)
    tokens, positions, params = bufs.upload(hb)
    eager = model(tokens, positions, params).clone()
    for c, (k, v) in zip(model.kv_caches, snap):
        c.key_cache.copy_(k)
        c.value_cache.copy_(v)
    step = GraphedStep(model, bufs, hb, greedy=False)
    for c, (k, v) in zip(model.kv_caches, snap):
        c.key_cache.copy_(k)
        c.value_cache.copy_(v)
    out = step.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager), "CUDA-graph replay must be bit identical to eager"
    # a second replay with refreshed metadata (next decode position) stays finite and changes
    hb2 = build_decode_batch(pool, [k + 1 for k in kv_lens], [1] * 4, a.vocab_size)
    bufs.upload(hb2)
    out2 = step.replay().clone()
    torch.cuda.synchronize()
    assert torch.isfinite(out2.float()).all() and not torch.equal(out2, eager)
