"""GPU (>= 2 devices): the NVLink peer-memory all-reduce (csrc/allreduce.cu) vs NCCL and vs a host
sum, one process per GPU.  Mirrors src/model_parallel/process_group_test.cpp:48-171 (all-reduce vs
host sum for fp32/fp16/bf16); bit-identical results on every rank are checked as well."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, algo=""):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), B200_AR_ALGO=algo)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from scalellm_b200.model_parallel import ProcessGroup
    pg = ProcessGroup(rank, world, dev)
    try:
        assert pg._comm is not None, "NVLink communicator was not created"
        assert pg._twoshot == (algo != "oneshot")
        for dtype in (torch.bfloat16, torch.float16, torch.float32):
            # (33, 1000): last row of the two-shot partition is short; (3, 4096): fewer rows than ranks
            for shape in ((64, 4096), (32, 8192), (1, 8), (7, 1024), (128, 4096), (33, 1000), (3, 4096)):
                g = torch.Generator(device=dev).manual_seed(100 * rank + shape[0])
                x = torch.randn(shape, generator=g, device=dev).to(dtype)
                ref = x.clone()
                dist.all_reduce(ref)                      # NCCL
                gathered = [torch.empty_like(x) for _ in range(world)]
                dist.all_gather(gathered, x)
                host_sum = sum(t.float() for t in gathered)  # fp32 sum in rank order
                y = x.clone()
                for it in range(3):                       # repeated calls: epochs / double buffering
                    y.copy_(x)
                    if (it + rank) % world == 0:          # rank skew: somebody always arrives late
                        torch.cuda._sleep(2_000_000)
                    pg.allreduce(y)
                torch.cuda.synchronize()
                # ours == fp32 rank-order sum rounded once.  A message above the peer-memory limit (fp32
                # [128, 4096] = 2 MiB) goes through NCCL, whose summation order is its own: close, not equal
                if y.numel() * y.element_size() <= pg._nvlink_max_bytes:
                    assert torch.equal(y, host_sum.to(dtype)), (dtype, shape)
                else:
                    assert torch.allclose(y.float(), host_sum, rtol=1e-5, atol=1e-5 * world), (dtype, shape)
                # NCCL rounds its partial sums to the element type in its own order: only a sanity
                # bound against it (the exact statement is the host sum above)
                tol = 1e-2 * world if dtype != torch.float32 else 1e-5 * world
                assert torch.allclose(y.float(), ref.float(), rtol=tol, atol=tol)
                # every rank holds the same bits
                ys = [torch.empty_like(y) for _ in range(world)]
                dist.all_gather(ys, y)
                assert all(torch.equal(ys[0], t) for t in ys)
        # the row-parallel GEMM's stream-K partials as the local operand: the all-reduce's copy-in
        # sums each tile's slots -> bit-identical to reduce_partials followed by the plain all-reduce
        from scalellm_b200 import kernels
        import numpy as np
        K, N, M = 2048, 4096, 64
        rng = np.random.default_rng(11 + rank)
        qw = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(K, N // 8), dtype=np.int64).astype(np.int32)).to(dev)
        qz = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(K // 128, N // 8), dtype=np.int64).astype(np.int32)).to(dev)
        sc = (torch.rand(K // 128, N, generator=torch.Generator().manual_seed(rank)) * 0.01 + 1e-3).bfloat16().to(dev)
        packed = kernels.w4a16_prepack_awq(qw, qz, sc, 128)
        a = torch.randn(M, K, generator=torch.Generator().manual_seed(5 + rank)).bfloat16().to(dev)
        parts = kernels.w4a16_gemm_splitk(a, packed, N, 128, poison=True)
        got = pg.allreduce_partials(parts, torch.bfloat16)
        want = kernels.w4a16_reduce_partials(parts)
        pg.allreduce(want)
        torch.cuda.synchronize()
        assert torch.isfinite(got.float()).all()
        assert torch.equal(got, want)
        # ... and with the residual add + RMSNorm fused in as well (one launch instead of three)
        gen = torch.Generator().manual_seed(99)               # residual stream: replicated on all ranks
        res = torch.randn(M, N, generator=gen).bfloat16().to(dev)
        wn = (1 + 0.1 * torch.randn(N, generator=gen)).bfloat16().to(dev)
        assert pg.supports_partials_norm(M, N, torch.bfloat16)
        r_ref, out_ref = res.clone(), torch.empty_like(res)
        kernels.rms_norm_residual(out_ref, r_ref, want, wn, 1e-5)
        for _ in range(2):                                     # twice: epochs / double buffering
            r_fused = res.clone()
            out_fused = pg.allreduce_partials_norm(parts, r_fused, wn, 1e-5)
        torch.cuda.synchronize()
        assert torch.equal(r_fused, r_ref) and torch.equal(out_fused, out_ref)
        # hidden 8192 (two vectors per thread in the two-shot kernel; the one-shot form stops at 4096)
        if pg._twoshot:
            K2, N2, M2 = 1024, 8192, 32
            qw2 = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(K2, N2 // 8), dtype=np.int64).astype(np.int32)).to(dev)
            qz2 = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(K2 // 128, N2 // 8), dtype=np.int64).astype(np.int32)).to(dev)
            sc2 = (torch.rand(K2 // 128, N2, generator=torch.Generator().manual_seed(rank)) * 0.01 + 1e-3).bfloat16().to(dev)
            packed2 = kernels.w4a16_prepack_awq(qw2, qz2, sc2, 128)
            a2 = torch.randn(M2, K2, generator=torch.Generator().manual_seed(7 + rank)).bfloat16().to(dev)
            parts2 = kernels.w4a16_gemm_splitk(a2, packed2, N2, 128, poison=True)
            want2 = kernels.w4a16_reduce_partials(parts2)
            pg.allreduce(want2)
            res2 = torch.randn(M2, N2, generator=gen).bfloat16().to(dev)
            wn2 = (1 + 0.1 * torch.randn(N2, generator=gen)).bfloat16().to(dev)
            assert pg.supports_partials_norm(M2, N2, torch.bfloat16)
            r_ref2, out_ref2 = res2.clone(), torch.empty_like(res2)
            kernels.rms_norm_residual(out_ref2, r_ref2, want2, wn2, 1e-5)
            r_f2 = res2.clone()
            out_f2 = pg.allreduce_partials_norm(parts2, r_f2, wn2, 1e-5)
            torch.cuda.synchronize()
            assert torch.equal(r_f2, r_ref2) and torch.equal(out_f2, out_ref2)
        # greedy sampling over a vocabulary-sharded lm_head == argmax of the gathered logits
        for dtype in (torch.bfloat16, torch.float32):
            for rows, nl in ((64, 16032), (5, 1000), (128, 264), (1, 7)):
                gl = torch.Generator(device=dev).manual_seed(31 * rank + rows)
                lg = torch.randn(rows, nl, generator=gl, device=dev).to(dtype)
                lg[0, :] = 1.0                                          # ties everywhere: first index wins
                if rows > 2:
                    lg[1, nl - 1] = 100.0 if rank == world - 1 else lg[1, nl - 1]   # winner in the last shard
                    lg[2, 3] = float("nan") if rank == world // 2 else lg[2, 3]    # NaN counts as the maximum
                outs = [torch.empty_like(lg) for _ in range(world)]
                dist.all_gather(outs, lg)
                want_ids = torch.argmax(torch.cat(outs, dim=-1).float(), dim=-1)
                for _ in range(2):
                    ids = pg.argmax_sharded(lg)
                torch.cuda.synchronize()
                assert ids is not None and torch.equal(ids, want_ids), (dtype, rows, nl)
        # larger than the symmetric buffer -> NCCL path, still correct
        big = torch.ones(2 << 20, device=dev)
        pg.allreduce(big)
        assert torch.equal(big, torch.full_like(big, world))
        # CUDA-graph capture + replay of the kernel (device-side epoch)
        x = torch.full((64, 4096), float(rank + 1), device=dev, dtype=torch.bfloat16)
        buf = x.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            pg.allreduce(buf)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        buf.copy_(x)
        with torch.cuda.graph(graph):
            pg.allreduce(buf)
        expect = float(sum(range(1, world + 1)))
        for _ in range(4):
            buf.copy_(x)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(buf, torch.full_like(buf, expect))
        # all-reduces directly after one another inside one graph (a sliced large message, a micro-
        # benchmark): each launch may start while its predecessor is still exchanging — the epoch must be
        # the one after the predecessor's (this sequence hung before the epoch was read after the wait)
        small = torch.full((64, 4096), 1.0, device=dev, dtype=torch.float32)
        chain = torch.cuda.CUDAGraph()
        with torch.cuda.graph(chain):
            for _ in range(6):
                pg.allreduce(small)
        for rep in range(3):
            small.fill_(1.0)
            chain.replay()
            torch.cuda.synchronize()
            assert torch.equal(small, torch.full_like(small, float(world) ** 6)), rep
    finally:
        pg.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,algo", [(2, "oneshot"), (2, "twoshot"), (4, ""), (8, ""), (8, "oneshot")])
def test_nvlink_allreduce_matches_nccl_and_host_sum(world, algo):
    """process_group_test.cpp:48-171 loops world_size = 1,2,4,...,device_count: so do we, with
    both algorithms (two-shot over LL lines = default, one-shot pull) at the ends of the range."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    mp.spawn(_worker, args=(world, _free_port(), algo), nprocs=world, join=True)




def _gather_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), B200_AR_GATHER="1")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from scalellm_b200.model_parallel import (ParallelArgs, ProcessGroup,
                                              gather_from_model_parallel_region)
    pg = ProcessGroup(rank, world, dev)
    try:
        pa = ParallelArgs(rank, world, pg)
        for dtype in (torch.bfloat16, torch.float32):
            for shape in ((64, 512), (1, 8), (7, 1024), (64, 16032), (3, 5, 64)):
                g = torch.Generator(device=dev).manual_seed(7 * rank + shape[-1])
                x = torch.randn(shape, generator=g, device=dev).to(dtype)
                x[..., 0] = -0.0                                   # a byte copy keeps the sign of zero
                outs = [torch.empty_like(x) for _ in range(world)]
                dist.all_gather(outs, x)
                want = torch.cat(outs, dim=-1)
                for _ in range(3):                                 # epochs / double buffering
                    got = pg.allgather_lastdim(x)
                    pg.allreduce(torch.ones(8, device=dev))        # collectives of both kinds interleave
                assert got is not None, (dtype, shape)
                torch.cuda.synchronize()
                assert got.shape == want.shape and torch.equal(got.view(torch.uint8), want.view(torch.uint8))
                assert torch.equal(gather_from_model_parallel_region(x, pa).view(torch.uint8),
                                   want.view(torch.uint8))
        odd = torch.randn(4, 6, device=dev)                        # 24-byte rows: not the fast path
        assert pg.allgather_lastdim(odd) is None
        assert gather_from_model_parallel_region(odd, pa).shape == (4, 6 * world)
    finally:
        pg.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_nvlink_allgather_lastdim_is_a_bit_exact_cat(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    mp.spawn(_gather_worker, args=(world, _free_port()), nprocs=world, join=True)


def test_same_process_group_like_ncclCommInitAll():
    """b200_ar_create_all: every rank in ONE process (the reference engine's model, one thread per
    GPU): peers mapped by peer access; launches issued from one thread, one device after the other."""
    import ctypes as C
    from scalellm_b200 import _lib
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    lib = _lib.load()
    comms = (C.c_void_p * world)()
    devs = (C.c_int * world)(*range(world))
    _lib.check(lib.b200_ar_create_all(comms, devs, world, 1 << 20))
    try:
        xs = [torch.randn(64, 4096, generator=torch.Generator().manual_seed(r)).bfloat16().to(f"cuda:{r}")
              for r in range(world)]
        want = sum(x.float().cpu() for x in xs).bfloat16()         # fp32 rank-order sum, one rounding
        for it in range(3):
            ys = [x.clone() for x in xs]
            for r in range(world):                                 # async launches: rank r spins until all arrived
                with torch.cuda.device(r):
                    _lib.check(lib.b200_ar_allreduce(comms[r], ys[r].data_ptr(), ys[r].numel(), 0,
                                                     torch.cuda.current_stream().cuda_stream))
            for r in range(world):
                torch.cuda.synchronize(r)
                assert torch.equal(ys[r].cpu(), want), (it, r)
        outs = [torch.empty(64, 4096 * world, dtype=torch.bfloat16, device=f"cuda:{r}") for r in range(world)]
        for r in range(world):
            with torch.cuda.device(r):
                _lib.check(lib.b200_ar_allgather(comms[r], outs[r].data_ptr(), xs[r].data_ptr(), 64,
                                                 4096 * 2, torch.cuda.current_stream().cuda_stream))
        cat = torch.cat([x.cpu() for x in xs], dim=-1)
        for r in range(world):
            torch.cuda.synchronize(r)
            assert torch.equal(outs[r].cpu(), cat)
    finally:
        for r in range(world):
            lib.b200_ar_destroy(comms[r])


def test_cpp_process_groups_in_one_process():
    """shim/b200_process_group: ProcessGroup::create_process_groups + the model-parallel region
    helpers with the reference's signatures, all ranks in this process."""
    from tests.test_shim import load_shim
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    shim = load_shim()
    pgs = shim.create_process_groups(list(range(world)))
    assert [pg.rank() for pg in pgs] == list(range(world)) and pgs[0].world_size() == world
    xs = [torch.randn(64, 4096, generator=torch.Generator().manual_seed(r)).bfloat16().to(f"cuda:{r}")
          for r in range(world)]
    want = sum(x.float().cpu() for x in xs).bfloat16()
    ys = [x.clone() for x in xs]
    outs = [pg.reduce_from_model_parallel_region(y) for pg, y in zip(pgs, ys)]   # async per device
    for r in range(world):
        torch.cuda.synchronize(r)
        assert outs[r].data_ptr() == ys[r].data_ptr() and torch.equal(ys[r].cpu(), want)
    cols = [pg.gather_from_model_parallel_region(x[:, :512].contiguous()) for pg, x in zip(pgs, xs)]
    cat = torch.cat([x[:, :512].cpu() for x in xs], dim=-1)
    for r in range(world):
        torch.cuda.synchronize(r)
        assert torch.equal(cols[r].cpu(), cat)
    lists = [[torch.empty(7, 8, device=f"cuda:{r}") for _ in range(world)] for r in range(world)]
    ins = [torch.full((7, 8), float(r + 1), device=f"cuda:{r}") for r in range(world)]
    for pg, i, o in zip(pgs, ins, lists):
        pg.allgather(i, o)
    for r in range(world):
        torch.cuda.synchronize(r)
        assert all(torch.equal(lists[r][q].cpu(), torch.full((7, 8), float(q + 1))) for q in range(world))
    sc = pgs[1].scatter_to_model_parallel_region(xs[1])
    assert torch.equal(sc, xs[1][:, 4096 // world: 2 * 4096 // world])
    del pgs


@pytest.mark.parametrize("fuse", [True, False])
def test_cpp_tensor_parallel_decode_step(fuse):
    """The C++ LlamaDecoderStep under TP=2, both ranks in this process (one Python thread each, the
    GIL is released inside forward): every rank ends with the same logits bit for bit, and they
    agree with the single-GPU step up to the summation order of the row-parallel reductions."""
    import threading
    from tests.test_cpp_host import CFG, _inv_freq, _state_dict, _shim
    from scalellm_b200.decode_step import BlockPool, StepBuffers, build_decode_batch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    shim = _shim()
    c, world, bs, B, n_blocks = CFG, 2, 16, 5, 64
    sd = _state_dict(seed=4)
    pgs = shim.create_process_groups(list(range(world)))

    def make(dev, pg):
        m = shim.LlamaDecoderStep(c["hidden"], c["n_layers"], c["n_heads"], c["n_kv_heads"], c["head_dim"],
                                  c["inter"], c["vocab"], c["max_pos"], c["eps"], "awq", 128, False,
                                  _inv_freq(), torch.empty(0, dtype=torch.bfloat16, device=dev), pg)
        m.load_state_dict(sd)
        m.fuse_partials = fuse
        return m

    g = torch.Generator().manual_seed(3)
    full_k = [torch.randn(n_blocks * bs, c["n_kv_heads"], c["head_dim"], generator=g).bfloat16()
              for _ in range(c["n_layers"])]
    full_v = [torch.randn(n_blocks * bs, c["n_kv_heads"], c["head_dim"], generator=g).bfloat16()
              for _ in range(c["n_layers"])]
    single = make("cuda:0", None)
    single.set_kv_caches([k.to("cuda:0") for k in full_k], [v.to("cuda:0") for v in full_v], bs)
    ranks = []
    for r in range(world):                                    # n_kv_heads == world: rank r owns kv head r
        m = make(f"cuda:{r}", pgs[r])
        m.set_kv_caches([k[:, r:r + 1].contiguous().to(f"cuda:{r}") for k in full_k],
                        [v[:, r:r + 1].contiguous().to(f"cuda:{r}") for v in full_v], bs)
        ranks.append(m)

    pool = BlockPool(n_blocks, bs, seed=1)
    kv = [37, 64, 5, 100, 17]
    for k in kv:
        pool.add_sequence(k + 8)
    hb = build_decode_batch(pool, kv, [1] * B, c["vocab"], seed=9)

    def args_on(dev):
        bufs = StepBuffers(torch.device(dev), 16, 8, 256)
        tokens, positions, p = bufs.upload(hb)
        return (tokens, positions, p.q_cu_seq_lens, p.kv_cu_seq_lens, p.kv_max_seq_len, p.q_max_seq_len,
                p.new_cache_slots, p.block_tables, p.cu_block_lens)

    want = single.forward(*args_on("cuda:0")).float().cpu()
    outs, errs = [None] * world, []
    rank_args = [args_on(f"cuda:{r}") for r in range(world)]   # pinned staging etc. before any rank spins
    for r in range(world):
        torch.cuda.synchronize(r)

    def run(r):
        try:
            with torch.cuda.device(r):
                outs[r] = ranks[r].forward(*rank_args[r])
                torch.cuda.synchronize(r)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs and all(not t.is_alive() for t in ts), errs
    got = [o.float().cpu() for o in outs]
    assert torch.equal(got[0], got[1])
    assert got[0].shape == want.shape
    assert torch.allclose(got[0], want, rtol=2e-2, atol=2e-2 * want.abs().max().item())
    del ranks, pgs
