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


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from scalellm_b200.model_parallel import ProcessGroup
    pg = ProcessGroup(rank, world, dev)
    try:
        assert pg._comm is not None, "NVLink communicator was not created"
        for dtype in (torch.bfloat16, torch.float16, torch.float32):
            for shape in ((64, 4096), (32, 8192), (1, 8), (7, 1024), (128, 4096)):
                g = torch.Generator(device=dev).manual_seed(100 * rank + shape[0])
                x = torch.randn(shape, generator=g, device=dev).to(dtype)
                ref = x.clone()
                dist.all_reduce(ref)                      # NCCL
                gathered = [torch.empty_like(x) for _ in range(world)]
                dist.all_gather(gathered, x)
                host_sum = sum(t.float() for t in gathered)  # fp32 sum in rank order
                y = x.clone()
                for _ in range(3):                        # repeated calls: epochs / double buffering
                    y.copy_(x)
                    pg.allreduce(y)
                torch.cuda.synchronize()
                # ours == fp32 rank-order sum rounded once
                assert torch.equal(y, host_sum.to(dtype)), (dtype, shape)
                tol = 1e-2 if dtype != torch.float32 else 1e-5
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
    finally:
        pg.close()
        dist.destroy_process_group()


def test_nvlink_allreduce_matches_nccl_and_host_sum():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=True)
