"""The decode step: the reference's Llama operator sequence driven through the B200 plugins.

This is the caller-side glue the benchmark and the parity tests need, not a model zoo:
  * `LlamaDecoder`  — models/meta/llama.h:61-64,123-133,170-177,220-232,281-289 restated over
    scalellm_b200.layers (RMSNorm -> qkv -> rope+kv-write -> paged attention -> o_proj
    -> [all-reduce] -> RMSNorm -> gate_up -> silu*up -> down -> [all-reduce]).  The two residual
    adds per layer are folded into the following RMSNorm (kernel::rms_norm_residual semantics,
    bit-identical to the separate torch adds of llama.h:174-176).
  * `build_decode_batch` — the tensor contract of Batch::prepare_model_input
    (src/engine/batch.cpp:77-270) for a synthetic decode batch: flat tokens, positions,
    q/kv cu_seq_lens, new_cache_slots, block_tables holding FIRST-SLOT ids (:206-209),
    cu_block_lens; block ids are a random permutation of the pool (attention_test.cpp:63-68).
  * `GraphedStep` — CUDA-graph capture / replay of one step with static input buffers
    (src/engine/model_runner.cpp:141-210).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import kernels
from .layers import (Attention, B200AttnHandler, ColumnParallelLinear, ColumnParallelQLinear,
                     InputParameters, KVCache, QuantArgs, RMSNorm, RowParallelLinear,
                     RowParallelQLinear, apply_llama3_rope_scaling, compute_default_inv_freq)
from .model_parallel import (ParallelArgs, gather_from_model_parallel_region, local_heads)


@dataclass
class LlamaArgs:
    """The ModelArgs fields the path reads (models/meta/llama.h:341-406)."""
    hidden_size: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 14336
    vocab_size: int = 128256
    rope_theta: float = 500000.0
    rms_norm_eps: float = 1e-5
    max_position_embeddings: int = 8192
    rope_scaling: Optional[Dict[str, float]] = field(
        default_factory=lambda: dict(factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                     original_max_position_embeddings=8192))

    @staticmethod
    def llama3_8b() -> "LlamaArgs":
        return LlamaArgs()

    @staticmethod
    def llama3_70b() -> "LlamaArgs":
        return LlamaArgs(hidden_size=8192, n_layers=80, n_heads=64, n_kv_heads=8,
                         intermediate_size=28672)


def _inv_freq(args: LlamaArgs) -> torch.Tensor:
    f = compute_default_inv_freq(args.head_dim, args.rope_theta)
    if args.rope_scaling:
        rs = args.rope_scaling
        f = apply_llama3_rope_scaling(f, rs["factor"], rs["low_freq_factor"],
                                      rs["high_freq_factor"],
                                      int(rs["original_max_position_embeddings"]))
    return f


class LlamaDecoder:
    def __init__(self, args: LlamaArgs, qa: QuantArgs, pa: ParallelArgs, device,
                 dtype: torch.dtype = torch.bfloat16):
        self.args, self.qa, self.pa, self.device, self.dtype = args, qa, pa, device, dtype
        w = pa.world_size
        self.H, self.Hkv = local_heads(args.n_heads, args.n_kv_heads, w)
        D, h, I = args.head_dim, args.hidden_size, args.intermediate_size
        self.q_size, self.kv_size = self.H * D, self.Hkv * D
        self.I_local = I // w
        quant = qa.quant_method in ("awq", "gptq")
        self.handler = B200AttnHandler.create_handler_with_rope(
            D, D, args.max_position_embeddings, _inv_freq(args), False, dtype, device)
        self.layers: List[Dict] = []
        local = ParallelArgs(0, 1, None)  # shards are materialised per rank already
        for _ in range(args.n_layers):
            if quant:
                qkv = ColumnParallelQLinear(h, self.q_size + 2 * self.kv_size, False, False, qa,
                                            local, device)
                o = RowParallelQLinear(self.q_size * w, h, False, True, qa, pa, device)
                gate_up = ColumnParallelQLinear(h, 2 * self.I_local, False, False, qa, local, device)
                down = RowParallelQLinear(self.I_local * w, h, False, True, qa, pa, device)
            else:
                qkv = ColumnParallelLinear(h, self.q_size + 2 * self.kv_size, False, local, dtype,
                                           device)
                o = RowParallelLinear(self.q_size * w, h, True, pa, dtype, device)
                gate_up = ColumnParallelLinear(h, 2 * self.I_local, False, local, dtype, device)
                down = RowParallelLinear(self.I_local * w, h, True, pa, dtype, device)
            self.layers.append(dict(
                input_norm=RMSNorm(h, args.rms_norm_eps, dtype, device), qkv=qkv, o=o,
                post_norm=RMSNorm(h, args.rms_norm_eps, dtype, device), gate_up=gate_up, down=down,
                attn=Attention(self.H, self.Hkv, D, self.handler)))
        self.final_norm = RMSNorm(h, args.rms_norm_eps, dtype, device)
        assert h % w == 0 and args.vocab_size % w == 0
        self.embed = torch.empty((args.vocab_size, h // w), dtype=dtype, device=device)
        self.lm_head = ColumnParallelLinear(h, args.vocab_size, True, pa, dtype, device)
        self.kv_caches: List[KVCache] = []
        # fuse the split-K reduction of o_proj / down_proj into the following RMSNorm
        self.fuse_splitk = os.environ.get("B200_FUSE_SPLITK", "1") != "0"

    # -- weights -----------------------------------------------------------------
    def load_layer(self, i: int, sd: Dict[str, Dict[str, torch.Tensor]]) -> None:
        """sd: {"qkv"|"o"|"gate_up"|"down": checkpoint tensors of THIS RANK's shard,
        "input_norm"/"post_norm": weight}.  Row-parallel entries are given unsharded-by-K
        local shards too (the modules were built with world=1 sharding for local tensors)."""
        L = self.layers[i]
        for name in ("qkv", "o", "gate_up", "down"):
            m = L[name]
            if isinstance(m, (ColumnParallelQLinear, RowParallelQLinear)):
                m._set_shard(sd[name]["qweight"], sd[name].get("qzeros"), sd[name]["scales"])
            else:
                m.weight.copy_(sd[name]["weight"])
        L["input_norm"].weight.copy_(sd["input_norm"])
        L["post_norm"].weight.copy_(sd["post_norm"])

    def init_random(self, seed: int = 0) -> None:
        """Random-init weights of the architecture on the device:
        dense N(0, 0.02); int4: q,z ~ U{0..15} (uniform random words), s = |randn| * 0.01."""
        g = torch.Generator(device=self.device).manual_seed(seed * 1000 + self.pa.rank)
        quant = self.qa.quant_method in ("awq", "gptq")
        gsz = self.qa.group_size

        def rand_q(K: int, N: int) -> Dict[str, torch.Tensor]:
            ng = 1 if gsz <= 0 else K // gsz
            ri = lambda *s: torch.randint(-2 ** 31, 2 ** 31 - 1, s, generator=g, device=self.device,
                                          dtype=torch.int64).to(torch.int32)
            sc = (torch.randn((ng, N), generator=g, device=self.device).abs() * 0.01 + 1e-4).to(
                torch.bfloat16)
            if self.qa.quant_method == "awq":
                return dict(qweight=ri(K, N // 8), qzeros=ri(ng, N // 8), scales=sc)
            return dict(qweight=ri(K // 8, N), qzeros=None, scales=sc)

        def rand_w(N: int, K: int) -> Dict[str, torch.Tensor]:
            return dict(weight=(torch.randn((N, K), generator=g, device=self.device) * 0.02).to(
                self.dtype))

        h = self.args.hidden_size
        for i, L in enumerate(self.layers):
            shapes = dict(qkv=(h, self.q_size + 2 * self.kv_size), o=(self.q_size, h),
                          gate_up=(h, 2 * self.I_local), down=(self.I_local, h))
            sd = {}
            for name, (K, N) in shapes.items():
                sd[name] = rand_q(K, N) if quant else rand_w(N, K)
            sd["input_norm"] = torch.ones(h, dtype=self.dtype, device=self.device)
            sd["post_norm"] = torch.ones(h, dtype=self.dtype, device=self.device)
            self.load_layer(i, sd)
            if quant:  # pack now and free the checkpoint-format tensors
                for name in ("qkv", "o", "gate_up", "down"):
                    L[name]._ensure_packed()
        self.embed.copy_((torch.randn(self.embed.shape, generator=g, device=self.device) * 0.02).to(
            self.dtype))
        self.lm_head.weight.copy_((torch.randn(self.lm_head.weight.shape, generator=g,
                                               device=self.device) * 0.02).to(self.dtype))

    def alloc_kv(self, n_blocks: int, block_size: int, randomize: bool = True, seed: int = 1) -> None:
        self.kv_caches = []
        g = torch.Generator(device=self.device).manual_seed(seed)
        for _ in range(self.args.n_layers):
            c = KVCache(n_blocks, block_size, self.Hkv, self.args.head_dim, self.dtype, self.device)
            if randomize:
                c.key_cache.normal_(generator=g)
                c.value_cache.normal_(generator=g)
            else:
                c.key_cache.zero_()
                c.value_cache.zero_()
            self.kv_caches.append(c)

    # -- forward -------------------------------------------------------------------
    def forward(self, tokens: torch.Tensor, positions: torch.Tensor, params: InputParameters,
                last_token_idxes: Optional[torch.Tensor] = None, greedy: bool = False,
                return_hidden: bool = False) -> torch.Tensor:
        """Returns logits [n_tokens, vocab] (llama.h:220-232 then :281-289); with
        last_token_idxes only those rows go through the final norm's output -> lm_head
        (the reference's `h.index_select(0, last_token_idxes)` for prefill chunks).
        greedy=True returns the sampled token ids [n_tokens] int64 instead (argmax of the logits);
        under tensor parallelism the vocabulary-sharded logits are never gathered — local argmax
        plus an 8-byte exchange per rank and row (ProcessGroup.argmax_sharded), same ids.
        return_hidden=True returns (logits, residual stream after the last layer) — for layer-level
        parity tests against the reference's kernels."""
        h = self.embed.index_select(0, tokens)
        if self.pa.world_size > 1:  # ParallelEmbedding: split on hidden + all-gather (embedding.h:74-79)
            h = gather_from_model_parallel_region(h, self.pa)
        I = self.I_local
        T = h.shape[0]
        # `pending` = output of the previous block, not yet added to the residual stream; either a
        # bf16 tensor or the producing GEMM's split-K partials (reduction fused into the norm).
        pending, pending_is_partials = None, False

        def norm_residual(norm, pend, is_partials):
            if is_partials:
                return norm.forward_residual_partials(pend, h, self.pa)
            return norm.forward_residual(pend, h)

        for L, cache in zip(self.layers, self.kv_caches):
            n1 = L["input_norm"](h) if pending is None else norm_residual(L["input_norm"], pending,
                                                                           pending_is_partials)
            fuse_qkv = (self.fuse_splitk and hasattr(L["qkv"], "supports_partials")
                        and L["qkv"].supports_partials(T) and L["attn"].supports_partials()
                        and not cache.empty())
            if fuse_qkv:   # GEMM partials -> (sum, RoPE, KV write) in one launch
                attn = L["attn"].forward_partials(L["qkv"].forward_partials(n1), positions, cache,
                                                  params, n1.dtype)
            else:
                qkv = L["qkv"](n1)
                q = qkv[:, : self.q_size]
                k = qkv[:, self.q_size: self.q_size + self.kv_size]
                v = qkv[:, self.q_size + self.kv_size:]
                attn = L["attn"](q, k, v, positions, cache, params)
            fuse_o = self.fuse_splitk and hasattr(L["o"], "supports_partials") and L["o"].supports_partials(T)
            o = L["o"].forward_partials(attn) if fuse_o else L["o"](attn)
            n2 = norm_residual(L["post_norm"], o, fuse_o)            # h = h + o ; n2 = norm(h)
            if (self.fuse_splitk and hasattr(L["gate_up"], "supports_partials")
                    and L["gate_up"].supports_partials(T)):
                act = kernels.silu_mul_splitk(L["gate_up"].forward_partials(n2), n2.dtype)
            else:
                gu = L["gate_up"](n2)
                act = kernels.silu_mul(gu[:, :I], gu[:, I:])
            fuse_d = self.fuse_splitk and hasattr(L["down"], "supports_partials") and L["down"].supports_partials(T)
            pending = L["down"].forward_partials(act) if fuse_d else L["down"](act)
            pending_is_partials = fuse_d
        if pending is None:
            hn = self.final_norm(h)
        else:
            hn = norm_residual(self.final_norm, pending, pending_is_partials)
        if return_hidden:   # h now holds the residual stream after the last block's add
            return self.lm_head(hn), h
        if last_token_idxes is not None:
            hn = hn.index_select(0, last_token_idxes)
        if greedy:
            pg = self.pa.process_group
            if self.pa.world_size > 1 and hasattr(pg, "argmax_sharded"):
                ids = pg.argmax_sharded(self.lm_head.forward_local(hn))
                if ids is not None:
                    return ids
            return kernels.argmax(self.lm_head(hn))
        return self.lm_head(hn)

    __call__ = forward


def prefill_chunks(prompt_len: int, chunk: int) -> List[Tuple[int, int]]:
    """Chunked prefill schedule of one request (scheduler token budget, continuous_scheduler.cpp):
    [(q_len, kv_len_after_chunk)] with q_len <= chunk, covering the prompt in order."""
    assert prompt_len > 0 and chunk > 0
    out, done = [], 0
    while done < prompt_len:
        q = min(chunk, prompt_len - done)
        done += q
        out.append((q, done))
    return out


# ---------------------------------------------------------------------------
# synthetic decode batch (host side; numpy)
# ---------------------------------------------------------------------------
@dataclass
class HostBatch:
    tokens: np.ndarray          # int32 [T]
    positions: np.ndarray       # int32 [T]
    q_cu_lens: np.ndarray       # int32 [B+1]
    kv_cu_lens: np.ndarray      # int32 [B+1]
    new_cache_slots: np.ndarray  # int32 [T]
    block_tables: np.ndarray    # int32 [sum blocks]  (first-slot ids)
    cu_block_lens: np.ndarray   # int32 [B+1]
    q_max: int
    kv_max: int


class BlockPool:
    """Minimal stand-in for the reference's BlockAllocator output: block ids are handed out from
    a random permutation (src/memory/block_allocator.cpp semantics are out of scope; only the ids
    it would produce matter to the kernels).  Per-sequence block lists are kept as rows of one
    int32 matrix of FIRST-SLOT ids (block_id * block_size, batch.cpp:206-209) so a step's metadata
    is built with a handful of vectorised numpy calls."""

    def __init__(self, n_blocks: int, block_size: int, seed: int = 2):
        self.block_size = block_size
        self.free = list(np.random.default_rng(seed).permutation(n_blocks).astype(np.int64))
        self.seq_blocks: List[List[int]] = []
        self._table = np.zeros((0, 0), dtype=np.int32)   # [n_seqs, max blocks per seq] first-slot ids

    def add_sequence(self, n_tokens_capacity: int) -> int:
        nb = (n_tokens_capacity + self.block_size - 1) // self.block_size
        if nb > len(self.free):
            raise RuntimeError("block pool exhausted")
        blocks = [self.free.pop() for _ in range(nb)]
        self.seq_blocks.append(blocks)
        n, w = self._table.shape
        t = np.zeros((n + 1, max(w, nb)), dtype=np.int32)
        t[:n, :w] = self._table
        t[n, :nb] = np.asarray(blocks, dtype=np.int64) * self.block_size
        self._table = t
        return n

    def n_blocks_of(self, seq: int) -> int:
        return len(self.seq_blocks[seq])


def build_decode_batch(pool: BlockPool, kv_lens: List[int], q_lens: List[int], vocab: int,
                       seed: int = 5) -> HostBatch:
    """One step's InputParameters on the host (Batch::prepare_model_input, batch.cpp:77-270) for
    sequences 0..B-1 of the pool.  kv_lens INCLUDE the new tokens (kv_cu_seq_lens semantics,
    parameters.h:35-37)."""
    bs = pool.block_size
    B = len(kv_lens)
    kv = np.asarray(kv_lens, dtype=np.int64)
    ql = np.asarray(q_lens, dtype=np.int64)
    nb = (kv + bs - 1) // bs
    cap = np.asarray([pool.n_blocks_of(b) for b in range(B)], dtype=np.int64)
    assert bool((nb <= cap).all()), "sequence outgrew its blocks"
    q_cu = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(ql, out=q_cu[1:])
    kv_cu = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(kv, out=kv_cu[1:])
    blk_cu = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(nb, out=blk_cu[1:])
    T = int(q_cu[-1])
    seq_of_tok = np.repeat(np.arange(B), ql)
    positions = np.arange(T) - q_cu[seq_of_tok] + (kv - ql)[seq_of_tok]      # kv-ql .. kv-1 per seq
    table = pool._table[:B]
    slots = table[seq_of_tok, positions // bs].astype(np.int64) + positions % bs
    tables = table[np.arange(table.shape[1])[None, :] < nb[:, None]]         # row-major: per-seq order
    tokens = np.random.default_rng(seed).integers(0, vocab, size=T)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    return HostBatch(i32(tokens), i32(positions), i32(q_cu), i32(kv_cu), i32(slots), i32(tables),
                     i32(blk_cu), int(ql.max()), int(kv.max()))


class StepBuffers:
    """One pinned host staging buffer + its device twin for a step's inputs (worker.cpp:132-135
    H2D copies): the seven int32 arrays live at fixed offsets, so a step is ONE host->device copy
    and the device views keep their addresses (CUDA-graph safe)."""

    FIELDS = ("tokens", "positions", "slots", "q_cu", "kv_cu", "blk_cu", "tables")

    def __init__(self, device, max_tokens: int, max_seqs: int, max_blocks: int):
        caps = {"tokens": max_tokens, "positions": max_tokens, "slots": max_tokens,
                "q_cu": max_seqs + 1, "kv_cu": max_seqs + 1, "blk_cu": max_seqs + 1,
                "tables": max_blocks}
        self.device = device
        self.off, total = {}, 0
        for f in self.FIELDS:
            self.off[f] = (total, caps[f])
            total += (caps[f] + 3) // 4 * 4          # keep every field 16-byte aligned
        self.total = total
        self.host = torch.zeros(total, dtype=torch.int32, pin_memory=torch.cuda.is_available())
        self.host_np = self.host.numpy()
        self.dev = torch.zeros(total, dtype=torch.int32, device=device)

    def h2d_bytes(self, hb: HostBatch) -> int:
        """Bytes of the single staging copy a step performs (up to the end of the block table)."""
        return 4 * (self.off["tables"][0] + len(hb.block_tables))

    def upload(self, hb: HostBatch) -> Tuple[torch.Tensor, torch.Tensor, InputParameters]:
        """Async H2D of one step's metadata from pinned memory on the current stream."""
        arrs = {"tokens": hb.tokens, "positions": hb.positions, "slots": hb.new_cache_slots,
                "q_cu": hb.q_cu_lens, "kv_cu": hb.kv_cu_lens, "blk_cu": hb.cu_block_lens,
                "tables": hb.block_tables}
        view = {}
        for f in self.FIELDS:
            o, cap = self.off[f]
            n = len(arrs[f])
            assert n <= cap, f"{f}: {n} > capacity {cap}"
            self.host_np[o:o + n] = arrs[f]
            view[f] = self.dev[o:o + n]
        used = self.off["tables"][0] + len(hb.block_tables)
        self.dev[:used].copy_(self.host[:used], non_blocking=True)
        B = len(hb.q_cu_lens) - 1
        params = InputParameters(
            num_sequences=B, q_cu_seq_lens=view["q_cu"], kv_cu_seq_lens=view["kv_cu"],
            kv_max_seq_len=hb.kv_max, q_max_seq_len=hb.q_max, new_cache_slots=view["slots"],
            block_tables=view["tables"], cu_block_lens=view["blk_cu"])
        return view["tokens"], view["positions"], params


class GraphedStep:
    """Capture one decode step (fixed batch size / max kv len) and replay it.  Inputs are
    refreshed by copying into the static buffers (model_runner.cpp:180-210)."""

    def __init__(self, model: LlamaDecoder, bufs: StepBuffers, hb: HostBatch, greedy: bool = True,
                 last_token_idxes: Optional[torch.Tensor] = None):
        """last_token_idxes (a static device tensor): only those rows reach lm_head — a captured
        prefill chunk needs the logits of its last token only (llama.h:281-289)."""
        self.model, self.bufs = model, bufs
        self.tokens, self.positions, self.params = bufs.upload(hb)
        self.greedy, self.last = greedy, last_token_idxes
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):      # warm-up outside capture: workspaces, tensor maps, cuBLAS
            for _ in range(2):
                self._run()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self) -> torch.Tensor:
        return self.model(self.tokens, self.positions, self.params, last_token_idxes=self.last,
                          greedy=self.greedy)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out
