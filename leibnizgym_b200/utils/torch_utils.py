"""CUDA versions of the reference's jit math primitives, same names and argument meaning
(ref leibnizgym/utils/torch_utils.py:18-180, leibnizgym/envs/trifinger/rewards.py:20-34).
Each call is one launch of the batched kernel behind the C ABI; inputs must be CUDA fp32."""
from __future__ import annotations

import torch

from .. import _native as nat


def _prep(*tensors):
    out = []
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("leibnizgym_b200.utils.torch_utils: CUDA tensors only (no CPU fallback)")
        out.append(t.to(torch.float32).contiguous())
    return out


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    assert a.shape == b.shape
    shape = a.shape
    a, b = _prep(a.reshape(-1, 4), b.reshape(-1, 4))
    out = torch.empty_like(a)
    nat.check(nat.load().lg_quat_mul(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], _stream(a)), "lg_quat_mul")
    return out.view(shape)


def quat_conjugate(a: torch.Tensor) -> torch.Tensor:
    shape = a.shape
    a = a.reshape(-1, 4)
    return torch.cat((-a[:, :3], a[:, -1:]), dim=-1).view(shape)


def quat_diff_rad(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = _prep(a.reshape(-1, 4), b.reshape(-1, 4))
    out = torch.empty(a.shape[0], device=a.device, dtype=torch.float32)
    nat.check(nat.load().lg_quat_diff_rad(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], _stream(a)),
              "lg_quat_diff_rad")
    return out


def _rowwise(fn_name, x, lower, upper):
    x2, lo, hi = _prep(x.reshape(-1, x.shape[-1]), lower, upper)
    out = torch.empty_like(x2)
    fn = getattr(nat.load(), fn_name)
    nat.check(fn(x2.data_ptr(), lo.data_ptr(), hi.data_ptr(), out.data_ptr(), x2.shape[0], x2.shape[1], _stream(x2)), fn_name)
    return out.view(x.shape)


def scale_transform(x, lower, upper):
    return _rowwise("lg_scale_transform", x, lower, upper)


def unscale_transform(x, lower, upper):
    return _rowwise("lg_unscale_transform", x, lower, upper)


def saturate(x, lower, upper):
    return _rowwise("lg_saturate", x, lower, upper)


def lgsk_kernel(x: torch.Tensor, scale: float = 50.0) -> torch.Tensor:
    (x1,) = _prep(x.reshape(-1))
    out = torch.empty_like(x1)
    nat.check(nat.load().lg_lgsk_kernel(x1.data_ptr(), float(scale), out.data_ptr(), x1.numel(), _stream(x1)), "lg_lgsk_kernel")
    return out.view(x.shape)


def cube_keypoints(pose: torch.Tensor, cube_size: float = 0.065) -> torch.Tensor:
    """Extension (no reference code): [n,7] poses -> [n,8,3] world-frame cube corners."""
    (p,) = _prep(pose.reshape(-1, 7))
    out = torch.empty((p.shape[0], 8, 3), device=p.device, dtype=torch.float32)
    nat.check(nat.load().lg_cube_keypoints(p.data_ptr(), float(cube_size), out.data_ptr(), p.shape[0], _stream(p)),
              "lg_cube_keypoints")
    return out


def compact_mask(mask: torch.Tensor) -> torch.Tensor:
    """torch.nonzero(mask).view(-1) through the ordered look-back compaction (ref env_base.py:374)."""
    if not mask.is_cuda:
        raise RuntimeError("CUDA tensors only")
    m = mask.to(torch.bool).contiguous()
    n = m.numel()
    lib = nat.load()
    ids = torch.empty(max(n, 1), device=m.device, dtype=torch.long)
    counts = torch.zeros(2, device=m.device, dtype=torch.int32)
    status = torch.zeros(max(int(lib.lg_scan_tiles(n)), 1), device=m.device, dtype=torch.int64)
    control = torch.zeros(4, device=m.device, dtype=torch.int64)
    nat.check(lib.lg_compact(m.data_ptr(), n, ids.data_ptr(), counts.data_ptr(), status.data_ptr(),
                             control.data_ptr(), _stream(m)), "lg_compact")
    return ids[: int(counts[0])]
