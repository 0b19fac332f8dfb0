"""oracle/ — CPU restatement of ScaleLLM's decode hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product (``scalellm_b200/``) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU baseline.

Every function cites the reference file:line it restates (paths relative to the
ScaleLLM tree @ffee4ffd, v0.2.6).  Arithmetic is plain torch-CPU / numpy in fp32
with the reference's per-op rounding made explicit.

Pinning: the oracle is checked in ``tests/test_oracle_golden.py``
against (a) golden vectors generated HERE by importing the reference's own
Python (``tests/kernels/quant_utils.py``, ``tests/kernels/attention/ref_attention.py``;
generator committed as ``tests/golden/make_golden.py``), (b) the llama3
rope-scaling known answers of ``src/layers/pos_embedding_test.cpp:98-138`` and
(c) the real GPTQ tensors of ``src/layers/quantization/data/gptq_small.safetensors``.
The AWQ zero-point Marlin path has no reference test (``tests/kernels/marlin_gemm_test.py:97``
"TODO"), so for ``has_zp=True`` the GEMM parity is pinned only through
quant_utils' ``w_ref = (q - zp) * s`` formula: parity unpinned by a reference KAT.
``oracle/_ref`` (git-ignored, built by ``oracle/ref/Makefile`` where ``/root/reference`` exists,
shipped to the GPU box as a built file): the reference's OWN elementwise kernels
(``src/kernels/{layernorm,pos_embedding,kv_cache,activation}_kernels.cu``) and its paged attention
(``src/kernels/attention``: attn_api.cpp + explicit bf16/fp16 head_dim-128 instantiations) and its
Marlin GEMM (``quantization/marlin/gptq_gemm.cu`` + the instantiations selected for 17 <= M <= 64)
compiled for sm_100a from the sources where they lie, behind a pybind module written here
(``oracle/ref/bindings.cpp``), plus a known-answer table of the Marlin int4->bf16 weight arithmetic
computed by the reference's own device functions (``oracle/ref/marlin_dequant_kat.cu``).
``tests/test_gpu_vs_reference_kernels.py`` runs our kernels against them on the same inputs.  The whole engine is not buildable here; see DESIGN.md.
"""

from . import ops, quant, llama, gpt2  # noqa: F401
