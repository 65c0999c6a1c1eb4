"""Parity metrics shared by the tests (max-abs, rel-L2 against an fp32 reference)."""
import torch


def rel_l2(a: torch.Tensor, ref: torch.Tensor) -> float:
    a = a.double().flatten()
    ref = ref.double().flatten()
    return float((a - ref).norm() / ref.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, ref: torch.Tensor) -> float:
    return float((a.double() - ref.double()).abs().max())


def report(name: str, a: torch.Tensor, ref: torch.Tensor) -> tuple[float, float]:
    r, m = rel_l2(a, ref), max_abs(a, ref)
    print(f"[parity] {name}: rel_l2={r:.3e} max_abs={m:.3e} ref_absmax={float(ref.abs().max()):.3e}")
    return r, m


# one bf16 rounding of an exact result has RMS relative error 2^-9/sqrt(3) ~ 1.1e-3
BF16_ONE_ROUNDING_REL_L2 = 2.0e-3
