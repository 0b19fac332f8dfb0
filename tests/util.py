"""Shared helpers for the parity tests."""
import numpy as np
import torch


def bf16_from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.astype(np.int16)).view(torch.bfloat16)


def f16_from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.astype(np.int16)).view(torch.float16)


def ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Distance in units-in-the-last-place between two bf16/fp16 tensors (same dtype)."""
    assert a.dtype == b.dtype and a.dtype in (torch.bfloat16, torch.float16)
    ai = a.contiguous().view(torch.int16).to(torch.int32)
    bi = b.contiguous().view(torch.int16).to(torch.int32)
    # map sign-magnitude to a monotonic integer line
    ai = torch.where(ai < 0, -(ai & 0x7FFF), ai)
    bi = torch.where(bi < 0, -(bi & 0x7FFF), bi)
    return (ai - bi).abs()


def assert_ulp(a: torch.Tensor, b: torch.Tensor, max_ulp: int = 1, max_frac: float = 0.0,
               what: str = ""):
    """All elements within max_ulp; at most max_frac of them different at all."""
    d = ulp_diff(a.cpu(), b.cpu())
    worst = int(d.max()) if d.numel() else 0
    frac = float((d > 0).float().mean()) if d.numel() else 0.0
    assert worst <= max_ulp, f"{what}: max ulp diff {worst} > {max_ulp} (mismatch frac {frac:.2e})"
    assert frac <= max_frac or max_frac >= 1.0, f"{what}: mismatch fraction {frac:.3e} > {max_frac}"


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.float().cpu(), ref.float().cpu()
    return float((a - ref).abs().mean() / ref.abs().mean().clamp_min(1e-12))


def assert_ulp_or_abs(a: torch.Tensor, b: torch.Tensor, max_ulp: int, abs_frac: float,
                      what: str = ""):
    """Every element within `max_ulp` ulps of the reference OR within abs_frac * max|ref|
    (ulp counts explode for results that cancel to ~0, where only the absolute error matters)."""
    a, b = a.cpu(), b.cpu()
    d = ulp_diff(a, b)
    err = (a.float() - b.float()).abs()
    atol = abs_frac * float(b.float().abs().max())
    bad = (d > max_ulp) & (err > atol)
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())} elements off by more than {max_ulp} ulp "
                                 f"and {atol:.3e} abs (worst abs err {float(err[bad].max()):.3e})")
