"""Shared helpers of the force-layer mirrors: a process-wide Philox key/counter for the per-episode
re-draws (the reference uses torch's global RNG [ref: OIGE/envs/USV/Hydrodynamics.py:146-151]) and the
CUDA-only guard.  [ref: OIGE/envs/USV/Utils.py] (the quaternion helpers there are fused into the kernels)."""
from __future__ import annotations

import ctypes
import itertools

import torch

from ... import _lib

_SEED = [1234]
_COUNTER = itertools.count(1)


def manual_seed(seed: int) -> None:
    """Seeds the Philox key used by reset_coefficients / reset_thruster_randomization."""
    _SEED[0] = int(seed)


def next_counter() -> int:
    return next(_COUNTER)


def seed() -> int:
    return _SEED[0]


def require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.UsvLibraryError(f"the B200 force layer runs on CUDA only (no CPU fallback); got device={device!r}")
    return dev


def f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


def randomize_rows(dst: torch.Tensor, env_ids: torch.Tensor, base, lo, hi, stream_id: int, log_space: bool = False) -> None:
    """dst[env_ids, c] = base[c] + U(lo[c], hi[c]) with Philox uniforms keyed (seed, env_id, counter)."""
    L = _lib.lib()
    dev = dst.device
    ids = env_ids.to(device=dev, dtype=torch.long).contiguous()
    ncols = dst.shape[1]
    mk = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev).expand(ncols).contiguous()
    b, l, h = mk(base), mk(lo), mk(hi)
    _lib.check(L.usv_randomize_rows_f32(_lib.ptr(dst), ctypes.c_int64(dst.stride(0)), _lib.ptr(ids), ctypes.c_int64(ids.numel()),
                                        ctypes.c_int32(ncols), _lib.ptr(b), _lib.ptr(l), _lib.ptr(h), ctypes.c_int32(int(log_space)),
                                        ctypes.c_uint64(seed()), ctypes.c_uint64(next_counter()), ctypes.c_uint32(stream_id),
                                        _lib.stream()), "usv_randomize_rows_f32")
