"""Synthetic simulator state for the TriFinger MDP hot path.

PhysX is out of scope (BASELINE.json north_star): the per-step inputs the path
reads from the simulator are replaced by seeded random tensors of the real
shapes and layouts:

    dof_state   [N, 9, 2]    (pos, vel) interleaved   ref trifinger_env.py:611-613
    root_state  [4N, 13]     robot 4e, stage 4e+1, object 4e+2, goal 4e+3   ref :617, :811-825
    rigid_body  [N, 20, 13]  fingertips at bodies 6 / 11 / 16   ref :615, :881-883
    dof_force   [N, 9]       ref :594-596
    ft_sensors  [N, 18]      3 x (fx,fy,fz,tx,ty,tz)   ref :598-600
    action      [N, 9]       policy output in [-1, 1]

The same generator feeds the reference harness (tests/golden/make_golden.py),
the oracle, the CUDA path and bench.py, so all of them see identical bits when
the sequence is generated on the CPU and uploaded.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

NUM_DOF = 9
NUM_BODIES = 20
NUM_ACTORS = 4
ROW = 13
FINGERTIP_BODIES = (6, 11, 16)
ROBOT_SLOT, STAGE_SLOT, OBJECT_SLOT, GOAL_SLOT = 0, 1, 2, 3

# joint limits of the real robot, ref trifinger_env.py:156-157
_JOINT_LOW = (-0.33, 0.0, -2.7)
_JOINT_HIGH = (1.0, 1.57, 0.0)


@dataclass
class StateSequence:
    """T consecutive synthetic simulator states for N envs (all float32)."""

    dof_state: torch.Tensor    # [T, N, 9, 2]
    root_state: torch.Tensor   # [T, 4N, 13]
    rigid_body: torch.Tensor   # [T, N, 20, 13]
    dof_force: torch.Tensor    # [T, N, 9]
    ft_sensors: torch.Tensor   # [T, N, 18]
    action: torch.Tensor       # [T, N, 9]

    @property
    def num_steps(self) -> int:
        return self.dof_state.shape[0]

    @property
    def num_envs(self) -> int:
        return self.dof_state.shape[1]

    def to(self, device, pin: bool = False) -> "StateSequence":
        def mv(t):
            t = t.to(device)
            return t.pin_memory() if pin else t
        return StateSequence(*(mv(getattr(self, f)) for f in self.__dataclass_fields__))

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * 4 for f in self.__dataclass_fields__)


def _unit_quat(g, shape, device):
    q = torch.randn(*shape, 4, generator=g, device=device, dtype=torch.float32)
    return q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def make_sequence(seed: int, num_steps: int, num_envs: int, device: str = "cpu",
                  first_env: int = 0, action_dim: int = NUM_DOF) -> StateSequence:
    """Seeded synthetic state sequence (distributions of SURVEY.md §8(d)).

    `first_env` offsets the seed so that a shard [first_env, first_env+num_envs)
    generated on its own rank does not repeat rank 0's numbers.
    """
    T, N = num_steps, num_envs
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 1_000_003 + int(first_env))
    kw = dict(generator=g, device=device, dtype=torch.float32)

    lo = torch.tensor(_JOINT_LOW * 3, device=device)
    hi = torch.tensor(_JOINT_HIGH * 3, device=device)
    dof_pos = lo + (hi - lo) * torch.rand(T, N, NUM_DOF, **kw)
    dof_vel = torch.randn(T, N, NUM_DOF, **kw).clamp_(-10.0, 10.0)
    dof_state = torch.stack((dof_pos, dof_vel), dim=-1).contiguous()

    root = torch.zeros(T, N, NUM_ACTORS, ROW, device=device, dtype=torch.float32)
    root[..., 6] = 1.0  # identity quaternion (xyzw) for every actor
    obj = root[:, :, OBJECT_SLOT]
    obj[..., 0:2] = -0.15 + 0.30 * torch.rand(T, N, 2, **kw)
    obj[..., 2] = 0.0325 + (0.2 - 0.0325) * torch.rand(T, N, **kw)
    obj[..., 3:7] = _unit_quat(g, (T, N), device)
    obj[..., 7:13] = 0.1 * torch.randn(T, N, 6, **kw)
    root_state = root.reshape(T, NUM_ACTORS * N, ROW).contiguous()

    rigid_body = torch.randn(T, N, NUM_BODIES, ROW, **kw)
    tip_lo = torch.tensor((-0.2, -0.2, 0.0), device=device)
    tip_hi = torch.tensor((0.2, 0.2, 0.3), device=device)
    for b in FINGERTIP_BODIES:
        rigid_body[:, :, b, 0:3] = tip_lo + (tip_hi - tip_lo) * torch.rand(T, N, 3, **kw)
        rigid_body[:, :, b, 3:7] = _unit_quat(g, (T, N), device)
        rigid_body[:, :, b, 7:13] = 0.05 * torch.randn(T, N, 6, **kw)

    dof_force = -0.36 + 0.72 * torch.rand(T, N, NUM_DOF, **kw)
    ft_sensors = 0.3 * torch.randn(T, N, 18, **kw)
    action = -1.0 + 2.0 * torch.rand(T, N, action_dim, **kw)   # 18 for the position_impedance command mode
    return StateSequence(dof_state, root_state, rigid_body.contiguous(), dof_force, ft_sensors, action)


def plant_goal_rows(seq: StateSequence, seed: int) -> None:
    """Fill the goal actor's root rows of every step with a seeded random pose (position inside the arena, unit
    quaternion) and angular velocity: what a simulator that integrates a moving goal body would leave there."""
    T, N = seq.num_steps, seq.num_envs
    dev = seq.root_state.device
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) * 7919 + 17)
    kw = dict(generator=g, device=dev, dtype=torch.float32)
    goal = seq.root_state.view(T, N, NUM_ACTORS, ROW)[:, :, GOAL_SLOT]
    goal[..., 0:2] = -0.12 + 0.24 * torch.rand(T, N, 2, **kw)
    goal[..., 2] = 0.0325 + 0.1 * torch.rand(T, N, **kw)
    goal[..., 3:7] = _unit_quat(g, (T, N), dev)
    goal[..., 10:13] = 0.5 * torch.randn(T, N, 3, **kw)


def plant_edge_cases(seq: StateSequence, goal_pose: torch.Tensor, first_step: int = 1) -> None:
    """Overwrite the first envs of every step >= first_step with the edge cases of SURVEY.md §A.7.

    `goal_pose` [N,7] is the goal the envs will hold (the caller forces it into
    the goal buffer).  env 0: object quat == goal quat; env 1: antipodal quat;
    env 2: non-unit object quat (|v| > 1 -> clamp -> pi); env 3: object > 1.78 m
    from the goal (lgsk denormal / zero); env 4: object exactly on the goal
    (success tolerance, d = 0); env 5: object at exactly the position tolerance
    along x (tie on <=); env 6: fingertips coincide with the object centre.
    """
    N = seq.num_envs
    assert N >= 8
    root = seq.root_state.view(seq.num_steps, N, NUM_ACTORS, ROW)
    obj = root[first_step:, :, OBJECT_SLOT]
    obj[:, 0, 3:7] = goal_pose[0, 3:7]
    obj[:, 1, 3:7] = -goal_pose[1, 3:7]
    obj[:, 2, 3:7] = torch.tensor([1.5, -0.7, 0.9, 0.1])
    obj[:, 3, 0:3] = goal_pose[3, 0:3] + torch.tensor([1.5, 1.2, 0.4])
    obj[:, 4, 0:7] = goal_pose[4, 0:7]
    obj[:, 5, 0:3] = goal_pose[5, 0:3]
    obj[:, 5, 0] += 0.01
    for b in FINGERTIP_BODIES:
        seq.rigid_body[first_step:, 6, b, 0:3] = obj[:, 6, 0:3]


def bernoulli_masks(seed: int, num_steps: int, num_envs: int, p: float,
                    device: str = "cpu") -> Optional[torch.Tensor]:
    """[T, N] bool masks used to inject resets (BASELINE.json config 5: p = 0.3)."""
    if p <= 0.0:
        return None
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) * 7919 + 17)
    return torch.rand(num_steps, num_envs, generator=g, device=device) < p


TWO_PI = 2.0 * math.pi
