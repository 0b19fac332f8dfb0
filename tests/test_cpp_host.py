"""C++ host side above the C ABI (shim/b200_layers.{h,cpp}): the reference's plugin-level interfaces
(AttentionHandler, ParallelLinearImpl, RMSNormImpl, KVCache, InputParameters) and the Llama decode
step, exposed through the _b200_shim pybind module.

CPU: the module builds, the classes construct, load a state dict and enforce the reference's
argument checks (qlinear_awq_marlin_impl.cpp:28-31,150-151).
GPU: the C++ decode step must produce bit-identical logits (and KV-cache contents) to the Python
mirror (scalellm_b200/decode_step.py) on the same weights, with and without the fused
GEMM-partials consumers."""

import numpy as np
import pytest
import torch


def _shim():
    import __graft_entry__ as g
    from tests.test_shim import load_shim     # one import name for the extension module
    g._build_shim()
    return load_shim()


CFG = dict(hidden=256, n_layers=2, n_heads=4, n_kv_heads=2, head_dim=64, inter=512, vocab=1024,
           max_pos=512, eps=1e-5)


def _state_dict(seed=0, method="awq", g=128):
    rng = np.random.default_rng(seed)
    h, H, Hkv, D, I, V = (CFG[k] for k in ("hidden", "n_heads", "n_kv_heads", "head_dim", "inter", "vocab"))
    ri = lambda *s: torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=s, dtype=np.int64).astype(np.int32))
    sd = {}
    shapes = dict(qkv=(h, (H + 2 * Hkv) * D), o=(H * D, h), gate_up=(h, 2 * I), down=(I, h))
    for i in range(CFG["n_layers"]):
        for name, (K, N) in shapes.items():
            p = f"layers.{i}.{name}."
            sc = (torch.from_numpy(rng.random((K // g, N), dtype=np.float32)) * 0.02 + 1e-3).bfloat16()
            if method == "awq":
                sd[p + "qweight"], sd[p + "qzeros"], sd[p + "scales"] = ri(K, N // 8), ri(K // g, N // 8), sc
            else:
                sd[p + "qweight"], sd[p + "scales"] = ri(K // 8, N), sc
        for n in ("input_norm", "post_norm"):
            sd[f"layers.{i}.{n}.weight"] = (1 + 0.1 * torch.from_numpy(rng.standard_normal(h).astype(np.float32))).bfloat16()
    sd["final_norm.weight"] = torch.ones(h).bfloat16()
    sd["embed.weight"] = (torch.from_numpy(rng.standard_normal((V, h)).astype(np.float32)) * 0.5).bfloat16()
    sd["lm_head.weight"] = (torch.from_numpy(rng.standard_normal((V, h)).astype(np.float32)) * 0.05).bfloat16()
    return sd


def _inv_freq():
    D = CFG["head_dim"]
    return 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))


def _make(shim, like, method="awq", group=128, is_sym=False):
    c = CFG
    return shim.LlamaDecoderStep(c["hidden"], c["n_layers"], c["n_heads"], c["n_kv_heads"], c["head_dim"],
                                 c["inter"], c["vocab"], c["max_pos"], c["eps"], method, group, is_sym,
                                 _inv_freq(), like)


def test_cpp_host_constructs_loads_and_checks_arguments():
    shim = _shim()
    like = torch.empty(0, dtype=torch.bfloat16)
    m = _make(shim, like)
    m.load_state_dict(_state_dict())
    assert m.fuse_partials in (True, False)
    with pytest.raises(RuntimeError, match="group_size"):        # qlinear_awq_marlin_impl.cpp:28-31
        _make(shim, like, group=48)
    with pytest.raises(RuntimeError, match="quant_method"):
        _make(shim, like, method="int8")
    bad = _state_dict()
    del bad["layers.1.down.scales"]
    with pytest.raises(RuntimeError, match="scales"):
        _make(shim, like).load_state_dict(bad)
    with pytest.raises(RuntimeError, match="set_kv_caches"):    # forward without caches
        z = torch.zeros(1, dtype=torch.int32)
        m.forward(z, z, z, z, 1, 1, z, z, z)


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [True, False])
def test_cpp_decode_step_matches_python_mirror(fuse):
    from scalellm_b200.decode_step import (BlockPool, LlamaArgs, LlamaDecoder, StepBuffers,
                                           build_decode_batch)
    from scalellm_b200.layers import QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs
    shim = _shim()
    dev = torch.device("cuda")
    sd = _state_dict()
    c = CFG
    args = LlamaArgs(hidden_size=c["hidden"], n_layers=c["n_layers"], n_heads=c["n_heads"],
                     n_kv_heads=c["n_kv_heads"], head_dim=c["head_dim"], intermediate_size=c["inter"],
                     vocab_size=c["vocab"], rope_theta=10000.0, rms_norm_eps=c["eps"],
                     max_position_embeddings=c["max_pos"], rope_scaling=None)
    py = LlamaDecoder(args, QuantArgs("awq", 4, 128), ParallelArgs(0, 1, None), dev)
    py.fuse_splitk = fuse
    for i in range(c["n_layers"]):
        layer = {n: {k: sd[f"layers.{i}.{n}.{k}"] for k in ("qweight", "qzeros", "scales")}
                 for n in ("qkv", "o", "gate_up", "down")}
        layer["input_norm"] = sd[f"layers.{i}.input_norm.weight"]
        layer["post_norm"] = sd[f"layers.{i}.post_norm.weight"]
        py.load_layer(i, layer)
    py.final_norm.weight.copy_(sd["final_norm.weight"])
    py.embed.copy_(sd["embed.weight"])
    py.lm_head.weight.copy_(sd["lm_head.weight"])
    bs, B = 16, 5
    pool = BlockPool(64, bs, seed=1)
    kv = [37, 64, 5, 100, 17]
    for k in kv:
        pool.add_sequence(k + 8)
    py.alloc_kv(64, bs, randomize=True, seed=3)
    hb = build_decode_batch(pool, kv, [1] * B, c["vocab"], seed=9)
    bufs = StepBuffers(dev, 16, 8, 256)
    tokens, positions, params = bufs.upload(hb)

    cpp = _make(shim, torch.empty(0, dtype=torch.bfloat16, device=dev))
    cpp.load_state_dict(sd)
    cpp.fuse_partials = fuse
    # separate caches with the same contents: both steps write the new token's K/V
    kc = [cc.key_cache.clone() for cc in py.kv_caches]
    vc = [cc.value_cache.clone() for cc in py.kv_caches]
    cpp.set_kv_caches(kc, vc, bs)

    want = py(tokens, positions, params)
    got = cpp.forward(tokens, positions, params.q_cu_seq_lens, params.kv_cu_seq_lens,
                      params.kv_max_seq_len, params.q_max_seq_len, params.new_cache_slots,
                      params.block_tables, params.cu_block_lens)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    for a, b in zip(kc, py.kv_caches):
        assert torch.equal(a, b.key_cache)
    for a, b in zip(vc, py.kv_caches):
        assert torch.equal(a, b.value_cache)


@pytest.mark.gpu
def test_cpp_cuda_graph_step_replays_like_eager():
    """CudaGraphStep (ModelRunner::CudaGraph, model_runner.cpp:141-210): capture one step, replay
    it with other metadata (different kv lengths, a shorter block table) — logits bit-identical to
    the eager C++ step on the same inputs, and the KV writes of the replay land in the caches."""
    from scalellm_b200.decode_step import BlockPool, StepBuffers, build_decode_batch
    shim = _shim()
    dev = torch.device("cuda")
    c = CFG
    sd = _state_dict(seed=2)
    bs, B, n_blocks = 16, 5, 64

    def fresh():
        m = _make(shim, torch.empty(0, dtype=torch.bfloat16, device=dev))
        m.load_state_dict(sd)
        g = torch.Generator(device=dev).manual_seed(3)
        kc = [torch.randn(n_blocks * bs, c["n_kv_heads"], c["head_dim"], generator=g, device=dev).bfloat16()
              for _ in range(c["n_layers"])]
        vc = [torch.randn(n_blocks * bs, c["n_kv_heads"], c["head_dim"], generator=g, device=dev).bfloat16()
              for _ in range(c["n_layers"])]
        m.set_kv_caches(kc, vc, bs)
        return m, kc, vc

    pool = BlockPool(n_blocks, bs, seed=1)
    for _ in range(B):
        pool.add_sequence(120)
    bufs = StepBuffers(dev, 16, 8, 256)

    def args_of(kv, seed):
        hb = build_decode_batch(pool, kv, [1] * B, c["vocab"], seed=seed)
        tokens, positions, p = bufs.upload(hb)
        # the graph owns its copies; clone so the next upload cannot alias what we compare against
        return [t.clone() for t in (tokens, positions, p.q_cu_seq_lens, p.kv_cu_seq_lens)] + \
               [120, 1] + [t.clone() for t in (p.new_cache_slots, p.block_tables, p.cu_block_lens)]

    graphed, g_kc, g_vc = fresh()
    eager, e_kc, e_vc = fresh()
    step = shim.CudaGraphStep()
    cap = args_of([100, 64, 5, 111, 17], seed=9)
    step.capture(graphed, *cap, B * (120 // bs + 2), False)
    eager.forward(*cap)                                        # the capture's warm-up + capture also wrote K/V
    for kv, seed in (([37, 64, 5, 100, 17], 4), ([1, 2, 3, 4, 5], 5), ([100, 64, 5, 111, 17], 9)):
        a = args_of(kv, seed)
        got = step.replay(*a).clone()
        want = eager.forward(*a)
        torch.cuda.synchronize()
        assert torch.equal(got, want), kv
        for x, y in zip(g_kc + g_vc, e_kc + e_vc):
            assert torch.equal(x, y)
    bad = args_of([5, 6, 7, 8, 9], 1)
    bad[2], bad[3] = bad[2][:-1], bad[3][:-1]                  # one sequence fewer than captured
    with pytest.raises(RuntimeError, match="batch size"):
        step.replay(*bad)


@pytest.mark.parametrize("method", ["awq", "gptq"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_cpp_tensor_parallel_shards_are_slices_of_the_full_weight(method, world):
    """shard_llama_layer (C++, views of the checkpoint tensors): dequantising a rank's shard gives
    exactly the q | k | v / gate | up columns (kv heads replicated when n_kv_heads < world) or the
    input rows that the rank owns in the dequantised unsharded weight."""
    from oracle import quant
    shim = _shim()
    H, Hkv, D, I, h, g = 8, 2, 128, 1024, 512, 128
    shapes = dict(qkv=(h, (H + 2 * Hkv) * D), o=(H * D, h), gate_up=(h, 2 * I), down=(I, h))
    full, dense = {}, {}
    for i, (name, (K, N)) in enumerate(shapes.items()):
        ck = (quant.random_awq_checkpoint if method == "awq" else quant.random_gptq_checkpoint)(K, N, g, seed=i)
        full[name] = {k: ck[k] for k in ("qweight", "qzeros", "scales") if ck.get(k) is not None}
        dense[name] = _dequant(quant, method, full[name], g)

    def kv_head(rank):
        return rank * (Hkv // world) if Hkv >= world else rank // (world // Hkv)

    for rank in range(world):
        qkv, o, gu, down = shim.shard_llama_layer(full["qkv"], full["o"], full["gate_up"], full["down"],
                                                  H, Hkv, D, I, method, g, rank, world)
        Hl, Hkvl, Il = H // world, max(1, Hkv // world), I // world
        q_cols = slice(rank * Hl * D, (rank + 1) * Hl * D)
        k_cols = slice(H * D + kv_head(rank) * D, H * D + (kv_head(rank) + Hkvl) * D)
        v_cols = slice((H + Hkv) * D + kv_head(rank) * D, (H + Hkv) * D + (kv_head(rank) + Hkvl) * D)
        want_qkv = torch.cat([dense["qkv"][:, c] for c in (q_cols, k_cols, v_cols)], dim=1)
        assert torch.equal(_dequant(quant, method, qkv, g), want_qkv)
        want_gu = torch.cat([dense["gate_up"][:, rank * Il:(rank + 1) * Il],
                             dense["gate_up"][:, I + rank * Il: I + (rank + 1) * Il]], dim=1)
        assert torch.equal(_dequant(quant, method, gu, g), want_gu)
        assert torch.equal(_dequant(quant, method, o, g), dense["o"][rank * Hl * D:(rank + 1) * Hl * D])
        assert torch.equal(_dequant(quant, method, down, g), dense["down"][rank * Il:(rank + 1) * Il])
    with pytest.raises(RuntimeError, match="quant groups"):     # 1024 / 16 = 64-row shards < group 128
        shim.shard_llama_layer(full["qkv"], full["o"], full["gate_up"], full["down"], 16, Hkv, 32, I, method,
                               g, 0, 16)


def _dequant(quant, method, t, g):
    """checkpoint tensors -> dense bf16 [K, N] with the oracle's unpackers."""
    if method == "awq":
        q, z = quant.unpack_awq(t["qweight"].contiguous()), quant.unpack_awq(t["qzeros"].contiguous())
    else:
        q = quant.unpack_gptq(t["qweight"].contiguous())
        z = (quant.unpack_gptq_zeros(t["qzeros"].contiguous(), plus_one=True) if "qzeros" in t
             else np.full((q.shape[0] // g, q.shape[1]), 8, dtype=q.dtype))
    return quant.dequant(q, z, t["scales"].contiguous(), g)


@pytest.mark.gpu
def test_cpp_model_runner_replays_captured_batch_sizes_and_falls_back():
    """ModelRunner (model_runner.cpp:25-139): a decode step whose batch size was captured is a
    graph replay, anything else (other batch size, multi-token queries, context beyond the
    captured maximum) runs eagerly — same logits either way."""
    from scalellm_b200.decode_step import BlockPool, StepBuffers, build_decode_batch
    shim = _shim()
    dev = torch.device("cuda")
    c = CFG
    sd = _state_dict(seed=6)
    bs, n_blocks, max_seq = 16, 64, 120

    def fresh():
        m = _make(shim, torch.empty(0, dtype=torch.bfloat16, device=dev))
        m.load_state_dict(sd)
        g = torch.Generator(device=dev).manual_seed(8)
        kc = [torch.randn(n_blocks * bs, c["n_kv_heads"], c["head_dim"], generator=g, device=dev).bfloat16()
              for _ in range(c["n_layers"])]
        vc = [torch.randn(n_blocks * bs, c["n_kv_heads"], c["head_dim"], generator=g, device=dev).bfloat16()
              for _ in range(c["n_layers"])]
        m.set_kv_caches(kc, vc, bs)
        return m, kc + vc

    graphed, g_caches = fresh()
    eager, e_caches = fresh()
    runner = shim.ModelRunner(graphed, 0, [5], 1, max_seq, bs, False)
    runner.capture_cuda_graphs(5)
    for x, y in zip(g_caches, e_caches):          # the capture's warm-up step wrote slot 0
        y.copy_(x)

    pool = BlockPool(n_blocks, bs, seed=1)
    for _ in range(6):
        pool.add_sequence(max_seq)
    bufs = StepBuffers(dev, 16, 8, 256)

    def args_of(kv, q, seed):
        hb = build_decode_batch(pool, kv, q, c["vocab"], seed=seed)
        hb.kv_max = max_seq       # the host scalar the graph was captured with: same work partition
        tokens, positions, p = bufs.upload(hb)
        return [t.clone() for t in (tokens, positions, p.q_cu_seq_lens, p.kv_cu_seq_lens)] + \
               [p.kv_max_seq_len, p.q_max_seq_len] + \
               [t.clone() for t in (p.new_cache_slots, p.block_tables, p.cu_block_lens)]

    cases = [([37, 64, 5, 100, 17], [1] * 5, True),        # captured batch size: replay
             ([1, 2, 3, 4, 5], [1] * 5, True),
             ([37, 64, 5, 100], [1] * 4, False),           # batch size not captured: eager
             ([37, 64, 5, 100, 17], [2, 1, 1, 1, 1], False)]  # multi-token query: eager
    replayed = 0
    for i, (kv, q, graph) in enumerate(cases):
        a = args_of(kv, q, seed=i)
        got = runner.forward(*a).clone()
        want = eager.forward(*a)
        torch.cuda.synchronize()
        assert torch.equal(got, want), (kv, q)
        replayed += int(graph)
        assert runner.num_cuda_graph_replayed() == replayed and runner.num_eager_execution() == i + 1 - replayed
        for x, y in zip(g_caches, e_caches):
            assert torch.equal(x, y)
