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


def expect_cuda(**tensors) -> None:
    """``expect_cuda(name=(tensor, dtype), ...)``: every operand of a kernel wrapper lives on the GPU with the dtype the
    native entry point casts its pointer to (the reference asserts the same, one line per operand)."""
    for name, (t, dtype) in tensors.items():
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype):
            got = f"{t.dtype} on {t.device}" if isinstance(t, torch.Tensor) else type(t).__name__
            raise AssertionError(f"`{name}` must be a CUDA tensor of dtype {dtype}, got {got}")


def launch_untuned(name: str, includes: tuple, template: str, operands) -> None:
    """Build (or fetch) the single variant of a kernel that has no tuning space and run it on the current stream.
    ``operands``: ordered ``(arg name, declared type, value)`` triples -- the ``launch`` signature and the call in one."""
    from .tuner import jit_tuner
    operands = tuple(operands) + (("stream", torch.cuda.Stream, current_stream()),)
    values = tuple(v for _, _, v in operands)
    runtime = jit_tuner.compile_and_tune(name=name, keys={}, space=tuple(), includes=includes,
                                         arg_defs=tuple((n, t) for n, t, _ in operands), template=template, args=values)
    check(runtime(*values), name)
