"""Helpers shared by the kernel wrappers."""
import torch

WS_UNIT = 256  # workspace sizes cross the 32-bit `int` JIT ABI in units of 256 bytes


def require_cuda() -> None:
    """The product path has no CPU fallback: fail loudly when there is no device."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "voltrix (B200 build): CUDA device required -- this package has no CPU fallback. "
            "torch.cuda.is_available() is False.")


def current_stream() -> torch.cuda.Stream:
    return torch.cuda.current_stream()


def check(rc: int, what: str) -> None:
    if rc != 0:
        names = {1: "invalid argument", 2: "CUDA error", 3: "workspace too small", 4: "unsupported configuration",
                 5: "overflow"}
        raise RuntimeError(f"voltrix {what} failed with code {rc} ({names.get(rc, 'unknown')})")


def ws_units(nbytes: int) -> int:
    return (int(nbytes) + WS_UNIT - 1) // WS_UNIT


def alloc_workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(ws_units(nbytes) * WS_UNIT, dtype=torch.uint8, device=device)
