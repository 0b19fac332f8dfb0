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
The reference's C++/CUDA kernels cannot be built in this image (no glog/gflags/
folly/libtorch-2.9/vcpkg, no GPU in the authoring container), so there is no
``oracle/_ref`` binary; see DESIGN.md.
"""

from . import ops, quant, llama, gpt2  # noqa: F401
