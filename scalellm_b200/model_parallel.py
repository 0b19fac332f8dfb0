"""Tensor-parallel plumbing: mirror of src/model_parallel/{process_group,model_parallel,parallel_args}.

One process per GPU (torchrun) instead of the reference's one thread per GPU
(src/engine/worker.h:90).  `torch.distributed` is only the plumbing (rendezvous,
IPC-handle exchange, all-gather); the latency-critical row-parallel all-reduce
(src/model_parallel/process_group.cpp:135-153, 2 per layer) runs through the
NVLink peer-memory kernel of libb200decode (csrc/allreduce.cu) when the tensor is
on CUDA and fits the symmetric buffer, else through NCCL.

The sharding helpers are pure index arithmetic and are unit-tested on CPU with
gloo (tests/test_tp_gloo.py).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class ProcessGroup:
    """ProcessGroup interface (src/model_parallel/process_group.h:10-60) over torch.distributed."""

    def __init__(self, rank: int, world_size: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, nvlink_max_bytes: int = 1 << 20,
                 gather_max_bytes: int = 8 << 20):
        self._rank, self._world, self._device, self._group = rank, world_size, device, group
        self._comm = None
        self._nvlink_max_bytes = nvlink_max_bytes      # peer-memory all-reduce: latency-bound sizes only
        # all-gathers go through the peer-memory kernel too (B200_AR_GATHER=0: NCCL + cat); their
        # messages may be larger (a one-shot gather moves no more bytes than a ring), so the
        # symmetric buffers are sized for them
        self._gather_max_bytes = gather_max_bytes if os.environ.get("B200_AR_GATHER", "1") != "0" else 0
        self._buffer_bytes = max(nvlink_max_bytes, self._gather_max_bytes)
        algo = os.environ.get("B200_AR_ALGO", "")
        self._twoshot = algo != "oneshot"      # two-shot over LL lines is the default at every world size
        if device.type == "cuda" and world_size > 1:
            self._init_nvlink()

    # -- reference accessors -------------------------------------------------
    def rank(self) -> int:
        return self._rank

    def world_size(self) -> int:
        return self._world

    def device(self) -> torch.device:
        return self._device

    # -- NVLink communicator -------------------------------------------------
    def _init_nvlink(self) -> None:
        lib = _lib.load()
        comm = C.c_void_p()
        handle = (C.c_uint8 * _lib.AR_HANDLE_BYTES)()
        with torch.cuda.device(self._device):
            check(lib.b200_ar_create(C.byref(comm), self._rank, self._world,
                                     self._buffer_bytes, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8)
        if dist.get_backend(self._group) == "nccl":
            mine = mine.to(self._device)
        gathered = [torch.empty_like(mine) for _ in range(self._world)]
        dist.all_gather(gathered, mine, group=self._group)
        blob = bytes(torch.cat([g.cpu() for g in gathered]).tolist())
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        with torch.cuda.device(self._device):
            check(lib.b200_ar_open_peers(comm, buf))
        dist.barrier(group=self._group)
        self._comm = comm

    # -- collectives -----------------------------------------------------------
    def allreduce(self, input: torch.Tensor) -> None:
        """In-place sum (process_group.cpp:135-153)."""
        if self._world == 1:
            return
        nbytes = input.numel() * input.element_size()
        # the peer-memory / NCCL choice uses rank-invariant properties only (size, dtype): ranks
        # that disagreed would dead-lock (one side spinning on flags, the other inside NCCL).  A
        # view that is not contiguous or not 16-byte aligned is staged through a temporary.
        if (self._comm is not None and input.is_cuda
                and 0 < nbytes <= self._nvlink_max_bytes and nbytes % 16 == 0
                and input.dtype in (torch.bfloat16, torch.float16, torch.float32)):
            dt = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}[input.dtype]
            direct = input.is_contiguous() and input.data_ptr() % 16 == 0
            x = input if direct else input.contiguous().clone()
            check(_lib.load().b200_ar_allreduce(self._comm, x.data_ptr(), x.numel(), dt,
                                                torch.cuda.current_stream().cuda_stream))
            if not direct:
                input.copy_(x)
            return
        dist.all_reduce(input, op=dist.ReduceOp.SUM, group=self._group)

    def allreduce_partials(self, partials, dtype: torch.dtype) -> torch.Tensor:
        """Sum over ranks of this rank's GEMM result delivered as stream-K partials
        (kernels.W4Partials: data [slots, rows, n] fp32 + the GEMM's K): the all-reduce's copy-in
        stage sums each tile's slots and rounds once to `dtype` (what a GEMM epilogue would store)."""
        data = partials.data
        S, rows, n = data.shape
        out = torch.empty((rows, n), dtype=dtype, device=data.device)
        nbytes = out.numel() * out.element_size()
        if (self._comm is not None and nbytes <= self._nvlink_max_bytes and self._fits(rows, n * out.element_size())
                and nbytes % 16 == 0 and dtype in (torch.bfloat16, torch.float16)):
            dt = 0 if dtype == torch.bfloat16 else 1
            check(_lib.load().b200_ar_allreduce_splitk(self._comm, out.data_ptr(), data.data_ptr(),
                                                       S, partials.K, n, out.numel(), dt,
                                                       torch.cuda.current_stream().cuda_stream))
            return out
        from . import kernels
        kernels.w4a16_reduce_partials(partials, out)
        self.allreduce(out)
        return out

    def supports_partials_norm(self, rows: int, n: int, dtype: torch.dtype) -> bool:
        max_rows, max_n = (128, 8192) if self._twoshot else (64, 4096)
        return (self._comm is not None and 0 < rows <= max_rows and n % 128 == 0 and n <= max_n
                and self._fits(rows, n * 2) and dtype in (torch.bfloat16, torch.float16))

    def _fits(self, rows: int, row_bytes: int) -> bool:
        """A message of `rows` rows fits the symmetric buffers (the two-shot form ships LL lines:
        twice the payload, one extra vector per rank and row of slack)."""
        if self._twoshot:   # column chunks are rounded up to whole 16-byte vectors per owner
            return rows * (row_bytes + 16 * self._world) * 2 <= self._buffer_bytes
        return rows * row_bytes <= self._buffer_bytes

    def allreduce_partials_norm(self, partials, residual: torch.Tensor, weight: torch.Tensor,
                                eps: float) -> torch.Tensor:
        """residual += allreduce(partials); returns rms_norm(residual) * weight — the row-parallel
        GEMM's reduction, the all-reduce, the residual add and the RMSNorm in one launch."""
        data = partials.data
        S, rows, n = data.shape
        out = torch.empty_like(residual)
        dt = 0 if residual.dtype == torch.bfloat16 else 1
        check(_lib.load().b200_ar_allreduce_splitk_norm(
            self._comm, out.data_ptr(), residual.data_ptr(), data.data_ptr(), S, partials.K,
            weight.data_ptr(), rows, n, eps, dt, torch.cuda.current_stream().cuda_stream))
        return out

    def allgather(self, input: torch.Tensor, outputs: List[torch.Tensor]) -> None:
        if self._world == 1:
            outputs[0].copy_(input)
            return
        dist.all_gather(outputs, input.contiguous(), group=self._group)

    def allgather_lastdim(self, input: torch.Tensor) -> Optional[torch.Tensor]:
        """cat(all-gather(input), dim=-1) in one launch over peer memory (bit exact), or None when
        the fast path does not apply (the caller then uses allgather + cat)."""
        if self._comm is None or self._gather_max_bytes == 0 or not input.is_cuda:
            return None
        row_bytes = input.size(-1) * input.element_size()
        nbytes = input.numel() * input.element_size()
        if nbytes == 0 or row_bytes % 16 or nbytes > self._gather_max_bytes:   # rank-invariant
            return None
        x = input.contiguous()
        if x.data_ptr() % 16:
            x = x.clone()
        rows = x.numel() // x.size(-1)
        out = torch.empty((*x.shape[:-1], x.size(-1) * self._world), dtype=x.dtype, device=x.device)
        check(_lib.load().b200_ar_allgather(self._comm, out.data_ptr(), x.data_ptr(), rows, row_bytes,
                                            torch.cuda.current_stream().cuda_stream))
        return out

    def argmax_sharded(self, logits: torch.Tensor) -> Optional[torch.Tensor]:
        """Greedy sampling over a vocabulary-sharded lm_head: argmax over ALL ranks' columns of
        each row == torch.argmax(gather_from_model_parallel_region(logits), -1), without moving
        the logits (one launch, 8 bytes per rank and row over NVLink).  None if not applicable."""
        if (self._comm is None or not logits.is_cuda or logits.dim() != 2 or logits.size(0) > 128
                or logits.dtype not in (torch.bfloat16, torch.float16, torch.float32)):
            return None
        x = logits if logits.stride(1) == 1 else logits.contiguous()
        out = torch.empty(x.size(0), dtype=torch.int64, device=x.device)
        dt = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}[x.dtype]
        check(_lib.load().b200_ar_argmax(self._comm, out.data_ptr(), x.data_ptr(), x.size(0), x.size(1),
                                         x.stride(0), dt, torch.cuda.current_stream().cuda_stream))
        return out

    def close(self) -> None:
        if self._comm is not None:
            _lib.load().b200_ar_destroy(self._comm)
            self._comm = None


@dataclass
class ParallelArgs:
    """src/model_parallel/parallel_args.h:10-22"""
    rank: int = 0
    world_size: int = 1
    process_group: Optional[ProcessGroup] = None


def gather_from_model_parallel_region(input: torch.Tensor, pa: ParallelArgs) -> torch.Tensor:
    """model_parallel.cpp:13-31: all-gather then cat on the last dim."""
    if pa.world_size == 1:
        return input
    fast = getattr(pa.process_group, "allgather_lastdim", None)
    if fast is not None:
        out = fast(input)
        if out is not None:
            return out
    outs = [torch.empty_like(input) for _ in range(pa.world_size)]
    pa.process_group.allgather(input, outs)
    return torch.cat(outs, dim=-1).contiguous()


def reduce_from_model_parallel_region(input: torch.Tensor, pa: ParallelArgs) -> torch.Tensor:
    """model_parallel.cpp:33-44"""
    if pa.world_size == 1:
        return input
    pa.process_group.allreduce(input)
    return input


def scatter_to_model_parallel_region(input: torch.Tensor, pa: ParallelArgs) -> torch.Tensor:
    """model_parallel.cpp:46-65"""
    if pa.world_size == 1:
        return input
    last = input.size(-1)
    assert last % pa.world_size == 0, f"last dim {last} not divisible by world {pa.world_size}"
    return input.split(last // pa.world_size, dim=-1)[pa.rank]


# ---------------------------------------------------------------------------
# shard index arithmetic (host logic; CPU-testable)
# ---------------------------------------------------------------------------
def shard_range(total: int, rank: int, world: int) -> slice:
    assert total % world == 0, f"{total} not divisible by world size {world}"
    per = total // world
    return slice(rank * per, (rank + 1) * per)


def local_heads(n_heads: int, n_kv_heads: int, world: int):
    """models/meta/llama.h:83-90: local q heads, local kv heads (>= 1 with replication)."""
    assert n_heads % world == 0
    return n_heads // world, max(1, n_kv_heads // world)


def kv_head_for_rank(n_kv_heads: int, rank: int, world: int) -> slice:
    """kv-head rows owned by `rank`; with n_kv_heads < world the heads are replicated
    world/n_kv_heads times (qkv_parallel_linear.cpp:28-70)."""
    if n_kv_heads >= world:
        return shard_range(n_kv_heads, rank, world)
    assert world % n_kv_heads == 0
    rep = world // n_kv_heads
    h = rank // rep
    return slice(h, h + 1)
