"""CPU, world_size 2, gloo: the tensor-parallel split of the int4 linears reproduces the
unsharded result — column split concatenates, row split sums through an all-reduce
(qlinear_awq_marlin_impl.cpp:129-365, model_parallel.cpp:13-65)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import quant
from scalellm_b200.layers import QuantArgs, _shard_qtensors
from scalellm_b200.model_parallel import (ParallelArgs, ProcessGroup,
                                          gather_from_model_parallel_region,
                                          reduce_from_model_parallel_region,
                                          scatter_to_model_parallel_region)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dequant_shard(method, qw, qz, sc, g):
    if method == "awq":
        return quant.dequant(quant.unpack_awq(qw), quant.unpack_awq(qz), sc, g)
    return quant.dequant(quant.unpack_gptq(qw), quant.unpack_gptq_zeros(qz, True), sc, g)


def _worker(rank, world, port, method):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pg = ProcessGroup(rank, world, torch.device("cpu"))
        pa = ParallelArgs(rank, world, pg)
        K, N, g = 512, 256, 128
        ck = (quant.random_awq_checkpoint if method == "awq" else quant.random_gptq_checkpoint)(
            K, N, g, seed=7)
        sd = dict(qweight=ck["qweight"], qzeros=ck["qzeros"], scales=ck["scales"])
        qa = QuantArgs(quant_method=method, bits=4, group_size=g)
        w_full = quant.dequant(ck["q"], ck["z"], ck["scales"], g)
        torch.manual_seed(3)
        x = torch.randn(4, K).bfloat16()
        y_full = x.float() @ w_full.float()

        # column parallel: each rank owns N/world columns; gather == full
        qw, qz, sc = _shard_qtensors(sd, qa, 1, rank, world, K, N)
        w_col = _dequant_shard(method, qw, qz, sc, g)
        assert torch.equal(w_col, w_full[:, rank * N // world:(rank + 1) * N // world])
        y_col = gather_from_model_parallel_region((x.float() @ w_col.float()), pa)
        assert torch.allclose(y_col, y_full, rtol=1e-5, atol=1e-5)

        # row parallel: each rank owns K/world rows; all-reduce of partial products == full
        qw, qz, sc = _shard_qtensors(sd, qa, 0, rank, world, K, N)
        w_row = _dequant_shard(method, qw, qz, sc, g)
        assert torch.equal(w_row, w_full[rank * K // world:(rank + 1) * K // world])
        x_loc = scatter_to_model_parallel_region(x, pa)
        y_row = reduce_from_model_parallel_region(x_loc.float() @ w_row.float(), pa)
        assert torch.allclose(y_row, y_full, rtol=1e-4, atol=1e-4)
    finally:
        dist.destroy_process_group()


def test_awq_column_and_row_split_world2():
    mp.spawn(_worker, args=(2, _free_port(), "awq"), nprocs=2, join=True)


def test_gptq_column_and_row_split_world2():
    mp.spawn(_worker, args=(2, _free_port(), "gptq"), nprocs=2, join=True)
