"""CPU: host-side logic of the decode step — batch metadata (Batch::prepare_model_input
contract), block pool, head / shard arithmetic."""
import numpy as np
import pytest
import torch

from oracle import ops
from scalellm_b200.decode_step import BlockPool, build_decode_batch
from scalellm_b200.model_parallel import kv_head_for_rank, local_heads, shard_range


def test_decode_batch_metadata_contract():
    bs = 8
    pool = BlockPool(n_blocks=64, block_size=bs, seed=2)
    kv_lens, q_lens = [17, 8, 33], [1, 1, 1]
    for kv in kv_lens:
        pool.add_sequence(kv + 4)
    hb = build_decode_batch(pool, kv_lens, q_lens, vocab=1000)
    assert hb.q_cu_lens.tolist() == [0, 1, 2, 3]
    assert hb.kv_cu_lens.tolist() == [0, 17, 25, 58]
    assert hb.cu_block_lens.tolist() == [0, 3, 4, 9]
    assert hb.positions.tolist() == [16, 7, 32]
    assert hb.q_max == 1 and hb.kv_max == 33
    assert all(hb.block_tables % bs == 0)                       # first-slot ids
    assert len(set(hb.block_tables.tolist())) == len(hb.block_tables)  # blocks never shared
    # the new token's slot is where the reference's lookup finds position kv-1
    bt = torch.from_numpy(hb.block_tables)
    for b, kv in enumerate(kv_lens):
        slots = ops.slot_ids_for_sequence(bt, torch.from_numpy(hb.cu_block_lens), b, kv, bs)
        assert int(slots[-1]) == int(hb.new_cache_slots[b])
        assert len(set(slots.tolist())) == kv


def test_multi_token_queries_and_pool_exhaustion():
    pool = BlockPool(n_blocks=4, block_size=16)
    pool.add_sequence(40)
    hb = build_decode_batch(pool, [37], [3], vocab=10)
    assert hb.positions.tolist() == [34, 35, 36] and hb.q_max == 3
    with pytest.raises(RuntimeError):
        pool.add_sequence(40)


def test_shard_arithmetic():
    assert shard_range(4096, 3, 8) == slice(1536, 2048)
    with pytest.raises(AssertionError):
        shard_range(10, 0, 4)
    assert local_heads(64, 8, 8) == (8, 1)
    assert local_heads(32, 8, 2) == (16, 4)
    assert local_heads(64, 8, 16) == (4, 1)          # kv heads replicated (qkv_parallel_linear.cpp:28-38)
    assert kv_head_for_rank(8, 5, 8) == slice(5, 6)
    assert kv_head_for_rank(8, 5, 16) == slice(2, 3)  # ranks 4,5 share kv head 2
    assert kv_head_for_rank(8, 1, 2) == slice(4, 8)


def test_build_decode_batch_matches_per_token_loop():
    """The vectorised step-metadata builder (batch.cpp:77-270 semantics) against a per-token loop:
    ragged kv lengths, multi-token queries, block-table rows of different lengths."""
    import numpy as np
    import torch
    from scalellm_b200.decode_step import BlockPool, StepBuffers, build_decode_batch
    rng = np.random.default_rng(0)
    pool = BlockPool(4000, 16, seed=3)
    caps = [int(x) for x in rng.integers(20, 600, size=9)]
    for c in caps:
        pool.add_sequence(c)
    kv = [int(rng.integers(5, c + 1)) for c in caps]
    ql = [int(rng.integers(1, min(k, 5) + 1)) for k in kv]
    hb = build_decode_batch(pool, kv, ql, 1000)
    bs = pool.block_size
    positions, slots, tables, q_cu, kv_cu, blk_cu = [], [], [], [0], [0], [0]
    for b, (k, q) in enumerate(zip(kv, ql)):
        blocks = pool.seq_blocks[b]
        nb = (k + bs - 1) // bs
        for p in range(k - q, k):
            positions.append(p)
            slots.append(blocks[p // bs] * bs + p % bs)
        tables.extend(blk * bs for blk in blocks[:nb])       # first-slot ids (batch.cpp:206-209)
        q_cu.append(q_cu[-1] + q)
        kv_cu.append(kv_cu[-1] + k)
        blk_cu.append(blk_cu[-1] + nb)
    assert hb.positions.tolist() == positions and hb.new_cache_slots.tolist() == slots
    assert hb.block_tables.tolist() == tables and hb.q_cu_lens.tolist() == q_cu
    assert hb.kv_cu_lens.tolist() == kv_cu and hb.cu_block_lens.tolist() == blk_cu
    assert hb.q_max == max(ql) and hb.kv_max == max(kv) and len(hb.tokens) == sum(ql)
    # one staging buffer, device views at fixed offsets
    bufs = StepBuffers(torch.device("cpu"), 64, 16, 4000)
    tok, pos, params = bufs.upload(hb)
    assert torch.equal(params.block_tables, torch.from_numpy(hb.block_tables))
    assert torch.equal(params.new_cache_slots, torch.from_numpy(hb.new_cache_slots))
    assert torch.equal(pos, torch.from_numpy(hb.positions)) and torch.equal(tok, torch.from_numpy(hb.tokens))
    p0 = params.block_tables.data_ptr()
    _, _, params2 = bufs.upload(hb)
    assert params2.block_tables.data_ptr() == p0


@pytest.mark.parametrize("K,N,ctas,nsub", [(4096, 4096, 148, 1), (4096, 6144, 148, 1), (4096, 28672, 148, 1),
                                           (14336, 4096, 148, 1), (4096, 28672, 148, 2), (128, 128, 148, 1),
                                           (512, 256, 148, 1), (512, 4096, 148, 1), (4096, 768, 148, 1),
                                           (1792, 4096, 132, 1), (4096, 3584, 7, 1),
                                           # Llama-3-70B: per-rank shapes at TP=8, then TP=1
                                           (8192, 1280, 148, 1), (1024, 8192, 148, 1), (8192, 7168, 148, 1),
                                           (3584, 8192, 148, 1), (8192, 10240, 148, 1), (8192, 57344, 148, 1),
                                           (28672, 8192, 148, 1)])
def test_w4a16_stream_k_partition_invariants(K, N, ctas, nsub):
    """The stream-K partition shared by the GEMM and the kernels that sum its partials (host
    arithmetic in libb200decode, no GPU needed): equal contiguous shares, every tile's contributors
    are consecutive CTAs, slot ranks are 0..contrib-1, and no tile needs more than 8 slots."""
    import ctypes as C
    import numpy as np
    from scalellm_b200 import _lib
    lib = _lib.load()
    NT = N // 128 // nsub
    plan = (C.c_int32 * 5)()
    first = (C.c_int32 * NT)()
    contrib = (C.c_int32 * NT)()
    rc = lib.b200_debug_w4a16_plan(N, K, ctas, nsub, plan, first, contrib)
    assert rc == 0, lib.b200_last_error()
    units, P, KT, NTp, slots = list(plan)
    assert KT == K // 128 and NTp == NT and units == KT * NT
    assert 1 <= P <= min(ctas, units) and 1 <= slots <= 8
    begin = [(p * units) // P for p in range(P + 1)]
    shares = np.diff(begin)
    assert shares.min() >= 1 and shares.max() - shares.min() <= 1      # equal, non-empty shares
    owner = np.repeat(np.arange(P), shares)                              # unit -> CTA
    for nt in range(NT):
        o = owner[nt * KT:(nt + 1) * KT]
        assert first[nt] == o[0] and contrib[nt] == o[-1] - o[0] + 1
        assert sorted(set(o.tolist())) == list(range(o[0], o[-1] + 1))   # consecutive CTAs, no holes
        assert contrib[nt] <= slots
    assert max(contrib) == slots


def test_prefill_chunk_schedule():
    from scalellm_b200.decode_step import prefill_chunks
    assert prefill_chunks(1, 128) == [(1, 1)]
    assert prefill_chunks(128, 128) == [(128, 128)]
    assert prefill_chunks(300, 128) == [(128, 128), (128, 256), (44, 300)]
    for P in (5, 129, 2048):
        s = prefill_chunks(P, 128)
        assert sum(q for q, _ in s) == P and s[-1][1] == P and all(0 < q <= 128 for q, _ in s)
        assert all(b[1] - a[1] == b[0] for a, b in zip(s, s[1:]))


@pytest.mark.parametrize("model", ["llama3_8b", "llama3_70b"])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_llama_tp_shard_shapes_are_valid_w4a16_shapes(world, model):
    """Every per-rank projection of Llama-3-8B / 70B under TP=1/2/4/8 must be a legal W4A16 shape
    (multiples of 128, quant groups aligned on the row-parallel K split:
    qlinear_awq_marlin_impl.cpp:150-151,287) — the decoder constructs without a GPU."""
    from scalellm_b200.decode_step import LlamaArgs, LlamaDecoder
    from scalellm_b200.layers import QuantArgs
    from scalellm_b200.model_parallel import ParallelArgs
    args = getattr(LlamaArgs, model)()
    args.n_layers = 1
    quant = QuantArgs("awq", 4, 128) if model == "llama3_8b" else QuantArgs("gptq", 4, 128, is_sym=True)
    m = LlamaDecoder(args, quant, ParallelArgs(world - 1, world, None), "cpu")
    L = m.layers[0]
    H, Hkv = local_heads(args.n_heads, args.n_kv_heads, world)
    assert (m.H, m.Hkv) == (H, Hkv) and H % Hkv == 0 and Hkv >= 1
    D, h, I = args.head_dim, args.hidden_size, args.intermediate_size
    assert (L["qkv"].K, L["qkv"].N) == (h, (H + 2 * Hkv) * D)
    assert (L["o"].K, L["o"].N) == (H * D, h)
    assert (L["gate_up"].K, L["gate_up"].N) == (h, 2 * I // world)
    assert (L["down"].K, L["down"].N) == (I // world, h)
    for name in ("qkv", "o", "gate_up", "down"):
        assert L[name].K % 128 == 0 and L[name].N % 128 == 0, name
    assert m.embed.shape == (args.vocab_size, h // world)


@pytest.mark.parametrize("heads", [(32, 8), (4, 1), (8, 8), (64, 8)])
@pytest.mark.parametrize("batch,max_q", [(1, 1), (7, 3), (64, 1), (256, 1)])
def test_attention_stream_partition_invariants(heads, batch, max_q):
    """Host arithmetic of the paged-attention work partition (libb200decode, no GPU needed; the
    SM count falls back to 148): every (sequence, row block, kv head) is cut into at most n_splits
    pieces, a piece's block-table window fits the shared-memory table for every block size, and
    the workspace query (which assumes block_size 1) covers every block size."""
    import ctypes as C
    from scalellm_b200 import _lib
    lib = _lib.load()
    H, Hkv = heads
    D = 128
    for max_kv in (1, 17, 2048, 8192, 100000):
        plans = {}
        for bs in (1, 8, 16, 128):
            out = (C.c_int64 * 8)()
            assert lib.b200_debug_attn_plan(batch, max_q, max_kv, H, Hkv, D, bs, out) == 0
            impl, n_splits, tpw, ntm, n_seq, n_rb, total, window = list(out)
            plans[bs] = n_splits
            if impl != 2:       # absurdly long context with tiny blocks falls back to the mma kernel
                continue
            assert ntm == (max_kv + 15) // 16 and n_rb == (max_q * (H // Hkv) + 15) // 16
            assert n_seq == batch * n_rb * Hkv and total == n_seq * ntm and tpw >= 1
            assert tpw * 16 // bs + 8 <= window or tpw * 16 <= bs, (tpw, bs, window)
            # pieces of the worst-placed sequence: its ntm tiles start anywhere in the stream
            worst = max((((s * ntm) % tpw) + ntm - 1) // tpw + 1 for s in range(min(n_seq, 4096)))
            assert worst <= n_splits, (worst, n_splits)
        ws = lib.b200_paged_attn_workspace_bytes(batch, max_q, max_kv, H, Hkv, D)
        need = max(plans.values())
        if need > 1:
            assert ws >= batch * max_q * H * need * (D + 1) * 4


def test_step_timeline_summary_arithmetic():
    """tools/step_timeline.py: coverage / gap / overlap arithmetic over kernel intervals."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "step_timeline", os.path.join(os.path.dirname(__file__), "..", "tools", "step_timeline.py"))
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    ev = [("void b200::k1<(int)3>(int)", 0.0, 10.0), ("k2", 9.0, 5.0), ("void b200::k1<(int)3>(int)", 20.0, 4.0)]
    rows, s = st.summarise(ev[::-1], 1)
    assert rows[0] == ("k1<3>", 2, 14.0) and rows[1] == ("k2", 1, 5.0)
    assert s["window_us"] == 24.0 and s["covered_us"] == 18.0 and s["overlap_us"] == 1.0
    assert s["gap_us"] == 6.0 and s["n_gaps"] == 1 and s["gap_max_us"] == 6.0
    assert "k1<3>" in st.render(rows, s, "t")
    assert st.summarise([], 1) == ([], {})


def test_attention_transposed_tile_fragment_algebra():
    """tools/attn_tr_model.py: the transposed attention tile (S^T = K Q^T, O^T = V^T P^T with
    ldmatrix / movmatrix fragments) restated lane by lane reproduces softmax(QK^T)V, for 1..8
    packed rows, a ragged causal end (stale K, NaN V past the end) and head_dim 64."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "attn_tr_model", os.path.join(os.path.dirname(__file__), "..", "tools", "attn_tr_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for n_rows in (1, 4, 7, 8):
        assert mod.check(seed=n_rows, n_rows=n_rows) < 5e-3
    assert mod.check(seed=9, n_rows=4, n_tiles=4, kv_end=53) < 5e-3
    assert mod.check(seed=3, D=64, n_rows=2, n_tiles=3, kv_end=40) < 5e-3


def test_bench_defaults_name_the_configuration_the_metric_is_quoted_on(monkeypatch):
    """bench.py with no flags = Llama-3-8B AWQ-int4, batch 64, kv_len 2048, block_size 8, 1 GPU
    (BASELINE.json configs[2]); the other models / quantisations are explicit opt-ins."""
    import importlib.util
    import os
    import sys
    spec = importlib.util.spec_from_file_location(
        "bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.gpus, a.batch, a.seqlen, a.block_size, a.quant, a.model, a.impl) == \
           (1, 64, 2048, 8, "awq", "llama3-8b", "b200")
    assert a.warmup >= 3 and a.steps >= 1 and not a.ttft
    assert bench.workload_name(a, 1) == \
        "Llama-3-8B AWQ-int4 g128 decode step, batch 64, kv_len 2048, block_size 8, TP=1"
    monkeypatch.setattr(sys, "argv", ["bench.py", "--model", "llama3-70b", "--quant", "gptq", "--gpus", "8"])
    b = bench.parse()
    assert bench.workload_name(b, 8).startswith("Llama-3-70B GPTQ-int4 g128") and bench.METRIC == "decode_tokens_per_s"


def test_every_environment_switch_is_documented():
    """Each B200_* switch read anywhere in the library, the host layers, bench.py or the tests is
    listed in DESIGN.md's table of environment switches."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(__file__), "..")
    files = (glob.glob(os.path.join(root, "scalellm_b200", "csrc", "*.cu*")) +
             glob.glob(os.path.join(root, "scalellm_b200", "*.py")) + glob.glob(os.path.join(root, "shim", "*.cpp")) +
             [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] +
             glob.glob(os.path.join(root, "tests", "*.py")))
    used = set()
    for f in files:
        src = open(f).read()
        used |= set(re.findall(r'getenv\("(B200_[A-Z0-9_]+)"\)', src))
        used |= set(re.findall(r'environ(?:\.get)?[\[(]"(B200_[A-Z0-9_]+)"', src))
        used |= set(re.findall(r'setenv\("(B200_[A-Z0-9_]+)"', src))
    design = open(os.path.join(root, "DESIGN.md")).read()
    missing = sorted(n for n in used if n not in design)
    assert used and not missing, missing


def test_attention_transposed_tile_source_runs_on_the_host():
    """tools/attn_tr_emu.py: the TR = 1 blocks of paged_attn.cu (Q fragments, the tile loop body,
    the output scatter), cut out of the .cu file and run by 32 host threads with emulated ldmatrix /
    mma.sync / movmatrix / shuffles over TMA-swizzled shared memory, reproduce softmax(QK^T)V —
    ragged causal ends, split-KV partials and multi-token queries included.  The default
    (GPU-validated) blocks go through the same harness first, which validates the emulation."""
    import os
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(__file__), "..", "tools", "attn_tr_emu.py")
    r = subprocess.run([sys.executable, tool], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count("\nok") == 2, r.stdout + r.stderr


def test_attention_stream_kernel_runs_on_the_host_in_every_instantiation():
    """tools/attn_emu.py: the whole paged-attention stream kernel cut out of paged_attn.cu and run
    on the host, every CTA by 32 threads over emulated mbarriers / tensor-map loads / ldmatrix /
    mma.sync / movmatrix, with the library's own work partition, on paged caches with shuffled
    block ids (GQA / MHA / MQA, block sizes 1 / 8 / 16, multi-token queries, many pieces, garbage
    past the sequence ends): the default instantiation (GPU-validated, so it validates the
    harness) and the opt-in ones (11 CTAs/SM ring, transposed tile, both); then the CUDA-core kernel
    (4 warps per CTA) at head_dim 128 (GPU-validated) and at the padded head dims 96 and 32, each
    followed by the combine kernel."""
    import os
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(__file__), "..", "tools", "attn_emu.py")
    r = subprocess.run([sys.executable, tool, "quick"], capture_output=True, text=True, timeout=2400)
    assert r.returncode == 0 and r.stdout.count("\nok") == 6, r.stdout + r.stderr
