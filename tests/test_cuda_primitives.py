"""GPU parity of the batched math primitives (C ABI) against the oracle's restatement of
leibnizgym/utils/torch_utils.py and rewards.py:20-34, plus the exhaustive division self-test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _quats(n, seed, unit=True):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True) if unit else q


def test_quat_mul_bit_exact():
    from leibnizgym_b200.utils import torch_utils as tu
    from oracle import trifinger_oracle as orc
    a, b = _quats(100_000, 1, unit=False), _quats(100_000, 2, unit=False)
    got = tu.quat_mul(a.cuda(), b.cuda()).cpu()
    assert torch.equal(got, orc.quat_mul(a, b))  # pure +,-,* in the reference's order: no rounding freedom


def test_quat_diff_rad():
    from leibnizgym_b200.utils import torch_utils as tu
    from oracle import trifinger_oracle as orc
    a, b = _quats(200_000, 3), _quats(200_000, 4)
    # edge cases of SURVEY.md A.7: identical, antipodal, non-unit (|v| > 1 -> clamp -> pi)
    a[0] = b[0]
    a[1] = -b[1]
    a[2] = torch.tensor([1.5, -0.7, 0.9, 0.1])
    got = tu.quat_diff_rad(a.cuda(), b.cuda()).cpu().numpy().astype(np.float64)
    exp = orc.quat_diff_rad(a, b).numpy().astype(np.float64)
    # asin differs by <= 2 ulp between libdevice and SLEEF; its argument is bit-identical
    np.testing.assert_allclose(got, exp, rtol=1e-6, atol=1e-7)
    assert abs(got[2] - np.pi) < 1e-6 and got[1] < 1e-3


def test_scale_unscale_saturate_bit_exact():
    from leibnizgym_b200.params import state_scale
    from leibnizgym_b200.config import difficulty_config, resolve_config
    from leibnizgym_b200.utils import torch_utils as tu
    from oracle import trifinger_oracle as orc
    lo, hi = state_scale(resolve_config(difficulty_config(4, 8)))
    lo, hi = torch.tensor(lo), torch.tensor(hi)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4096, lo.numel(), generator=g) * 2.0
    for name, fn in (("scale", orc.scale_transform), ("unscale", orc.unscale_transform), ("saturate", orc.saturate)):
        got = getattr(tu, {"scale": "scale_transform", "unscale": "unscale_transform", "saturate": "saturate"}[name])(
            x.cuda(), lo.cuda(), hi.cuda()).cpu()
        assert torch.equal(got, fn(x, lo, hi)), name


def test_lgsk_kernel():
    from leibnizgym_b200.utils import torch_utils as tu
    from oracle import trifinger_oracle as orc
    x = torch.cat([torch.linspace(0, 0.6, 10_000), torch.tensor([0.0, 1.7, 1.78, 1.8, 2.5, 10.0])])
    got = tu.lgsk_kernel(x.cuda()).cpu().numpy().astype(np.float64)
    exp = orc.lgsk(x).numpy().astype(np.float64)
    np.testing.assert_allclose(got, exp, rtol=2e-6, atol=1e-44)   # exp: <= 2 ulp on both sides
    assert got[-1] == 0.0 and got[-2] == 0.0                      # overflow -> 0 is intended (rewards.py:34)


def test_compaction_matches_nonzero():
    from leibnizgym_b200.utils import torch_utils as tu
    g = torch.Generator().manual_seed(6)
    for n, p in ((1, 1.0), (1, 0.0), (127, 0.5), (128, 1.0), (129, 0.3), (16384, 0.3), (300_000, 0.01), (300_000, 0.0)):
        m = torch.rand(n, generator=g) < p
        got = tu.compact_mask(m.cuda()).cpu()
        assert torch.equal(got, torch.nonzero(m).view(-1)), (n, p)


def test_fast_division_contract_exhaustive():
    """The fused kernel divides by per-column constants with a 3-instruction sequence.  For EVERY one
    of the 2^32 numerators and every distinct span the env uses: bit-identical to IEEE x / span when
    2^-100 <= |x| < 2^100 or x = 0; within 1 ulp below that; flagged for the exact path above it."""
    from leibnizgym_b200 import _native as nat
    from leibnizgym_b200.params import state_scale
    from leibnizgym_b200.config import difficulty_config, resolve_config
    lib = nat.load()
    spans = set()
    for mode in ("torque", "position", "position_impedance"):
        for norm_a in (True, False):
            lo, hi = state_scale(resolve_config(difficulty_config(4, 8, command_mode=mode, normalize_action=norm_a)))
            spans |= set(((hi - lo).astype(np.float32) * np.float32(0.5)).tolist())  # the kernel divides by span / 2
    spans.add(1.0)                               # normalize_obs = False
    spans.add(float(np.float32(0.02)))           # sim dt: finger_move_penalty divides by it (rewards.py:261)
    for s in sorted(spans):
        out = torch.zeros(3, dtype=torch.int64, device="cuda")
        s32 = np.float32(s)
        rcp = np.float32(1.0) / s32
        nat.check(lib.lg_selftest_division(float(s32), float(rcp), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "selftest")
        bad, worst_ulp, unflagged = (int(x) for x in out.cpu())
        assert bad == 0 and worst_ulp <= 1 and unflagged == 0, (s, bad, worst_ulp, unflagged)


def test_cube_keypoints_extension():
    from leibnizgym_b200.utils import torch_utils as tu
    q = _quats(1000, 8)
    p = torch.randn(1000, 3)
    pose = torch.cat([p, q], dim=-1)
    got = tu.cube_keypoints(pose.cuda(), 0.065).cpu().double()
    # float64 restatement: R(q) v + p for the 8 corners
    x, y, z, w = (q[:, i].double() for i in range(4))
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    h = 0.0325
    corners = torch.tensor([[(h if k & 1 else -h), (h if k & 2 else -h), (h if k & 4 else -h)] for k in range(8)],
                           dtype=torch.float64)
    exp = torch.einsum("nij,kj->nki", R, corners) + p.double()[:, None, :]
    assert (got - exp).abs().max() < 1e-6
