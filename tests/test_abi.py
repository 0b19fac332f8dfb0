"""CPU: the C-ABI library loads and exports every symbol include/b200_decode.h declares
(no compute calls — there is no GPU here)."""
import os
import re
import subprocess

import pytest

from scalellm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200_decode.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ("b200_rms_norm", "b200_rms_norm_residual", "b200_rope_inplace", "b200_kv_write",
                 "b200_rope_kv_write", "b200_silu", "b200_silu_mul", "b200_paged_attn_decode",
                 "b200_paged_attn_workspace_bytes", "b200_w4a16_prepack_awq",
                 "b200_w4a16_prepack_gptq", "b200_w4a16_gemm", "b200_ar_create",
                 "b200_ar_allreduce", "b200_ar_destroy"):
        assert must in syms, must


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                        text=True, check=True).stdout
    exported = set(re.findall(r"\bT\s+(b200_\w+)", nm))
    declared = set(header_symbols())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    # nothing undeclared leaks out of the C ABI
    assert exported <= declared, f"exported but not declared: {sorted(exported - declared)}"


def test_ctypes_signatures_cover_the_header():
    assert set(_lib.SIGNATURES) == set(header_symbols())
    lib = _lib.load()
    assert lib.b200_abi_version() == 1
    assert lib.b200_last_error() is not None


def test_argument_validation_without_gpu():
    """Error paths return negative codes and a message instead of aborting."""
    lib = _lib.load()
    rc = lib.b200_rms_norm(None, None, None, 4, 128, 1e-5, 0, None)
    assert rc == -1 and b"null" in lib.b200_last_error()
    assert lib.b200_w4a16_packed_bytes(4096, 4096, 128) == (4096 // 128) ** 2 * (8192 + 256 + 128)
    assert lib.b200_w4a16_packed_bytes(4096, 4100, 128) == -1
    rc = lib.b200_w4a16_gemm(1, 1, 1, None, 4, 100, 128, 128, 100, 128, None, 0, None)
    assert rc == -1
    rc = lib.b200_paged_attn_decode(1, 1, 1, 1, 1, 1, 1, 1, None, 1, 32, 8, 128, 64, 4096, 128,
                                    4096, 128, 1024, 128, 7, 1, 16, 0.1, 0.0, -1, None, 0, 0, None)
    assert rc == -1 and b"power of two" in lib.b200_last_error()
    assert lib.b200_paged_attn_workspace_bytes(64, 1, 2048, 32, 8, 128) > 0
    # same-process communicator group (ncclCommInitAll counterpart): checked before any CUDA call
    import ctypes as C
    comms = (C.c_void_p * 2)()
    devs = (C.c_int * 2)(0, 0)
    assert lib.b200_ar_create_all(comms, devs, 2, 1 << 20) == -1 and b"twice" in lib.b200_last_error()
    assert lib.b200_ar_create_all(comms, devs, 9, 1 << 20) == -1
    assert lib.b200_ar_create_all(comms, devs, 1, 24) == -1
    assert lib.b200_ar_create_all(None, devs, 1, 1 << 20) == -1
    assert lib.b200_ar_allgather(None, 16, 16, 1, 16, None) == -1 and b"null" in lib.b200_last_error()
    assert lib.b200_ar_allgather(1, 16, 32, 1, 24, None) == -1 and b"16 bytes" in lib.b200_last_error()


def test_kernels_reject_cpu_tensors():
    import torch
    from scalellm_b200 import kernels
    x = torch.zeros(2, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kernels.rms_norm(torch.empty_like(x), x, torch.ones(128, dtype=torch.bfloat16), 1e-5)
