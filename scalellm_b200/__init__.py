"""scalellm_b200 — B200-native (sm_100a) decode hot path behind ScaleLLM's operator API.

Scope: paged-KV decode attention, AWQ/GPTQ int4 x bf16 matmul, RMSNorm,
RoPE, KV-slot write, SiLU*mul and the tensor-parallel all-reduce — hand-written CUDA in
`csrc/` behind the C ABI of `include/b200_decode.h`; `kernels.py` mirrors the reference's
`src/kernels` signatures, `layers.py` / `model_parallel.py` its `src/layers` /
`src/model_parallel` plugin interfaces, `decode_step.py` drives one Llama decode step.
There is no CPU fallback: every op needs the built extension and a CUDA device.
"""
__version__ = "0.1.0"
