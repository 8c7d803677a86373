"""Synthetic stand-in for the IsaacGym tensor API.

PhysX is out of scope; what the hot path needs from the simulator is (a) the five
state tensors it reads zero-copy (ref envs/trifinger/trifinger_env.py:592-617),
(b) a frame counter (ref envs/env_base.py:286-289) and (c) somewhere for the
reset rows to go.  `SyntheticSim` owns those tensors on the GPU and "steps" by
copying the next state of a `StateSequence` into them in place, exactly what the
reference harness' FakeGym does on the CPU (tests/golden/ref_harness.py).

The sequence may live on the device (device-to-device copy) or in pinned host
memory (host-to-device copy: the situation of the reference's default
`use_gpu_pipeline: False`, where the simulator state is host-resident).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .synthetic import StateSequence


class LazyCount:
    """Number of valid entries of a device-side index list, known only on the device until somebody asks.

    The fused step compacts the reset flags on the GPU and never synchronises the host, so the count the reference
    passes to `set_*_tensor_indexed` as a Python int (`len(indices)`, ref trifinger_env.py:419-423, :437-440) exists
    as a device scalar.  The simulator adapter decides how to consume it: `int(count)` reads it back (one stream
    synchronisation, what the reference pays anyway), `count.device_scalar` / `count.scale` feed an API that takes
    device-side counts."""

    __slots__ = ("device_scalar", "scale")

    def __init__(self, device_scalar: torch.Tensor, scale: int = 1):
        self.device_scalar, self.scale = device_scalar, int(scale)

    def __int__(self) -> int:
        return int(self.device_scalar.item()) * self.scale

    __index__ = __int__

    def __repr__(self) -> str:
        return f"LazyCount({self.scale} x device scalar)"


class SyntheticSim:
    def __init__(self, seq: StateSequence, device: str = "cuda:0"):
        self.seq = seq
        self.device = torch.device(device)
        self.num_envs = seq.num_envs
        N = self.num_envs
        f32 = dict(device=self.device, dtype=torch.float32)
        self.dof_state = torch.zeros(N, 9, 2, **f32)
        self.root_state = torch.zeros(4 * N, 13, **f32)
        self.rigid_body = torch.zeros(N, 20, 13, **f32)
        self.dof_force = torch.zeros(N, 9, **f32)
        self.ft_sensors = torch.zeros(N, 18, **f32)
        self.frame_count = 0
        self.cursor = 0
        self.fingertip_bodies = (6, 11, 16)
        self.bodies_per_env, self.actors_per_env = 20, 4
        self.slots = (0, 2, 3)  # robot, object, goal actor slots inside one env
        self.before_simulate: Optional[Callable[["SyntheticSim"], None]] = None  # test hook
        self.applied_torque = None
        self._uploader = None
        self._integrate = None
        self._load(0)  # what refresh_* shows before the first simulate

    def _load(self, t: int) -> None:
        s = self.seq
        if self._uploader is not None:   # pinned host sequence: stage only the rows the path reads
            self._uploader(t)
            return
        nb = s.dof_state.device.type == "cpu"
        self.dof_state.copy_(s.dof_state[t], non_blocking=nb)
        self.root_state.copy_(s.root_state[t], non_blocking=nb)
        self.rigid_body.copy_(s.rigid_body[t], non_blocking=nb)
        self.dof_force.copy_(s.dof_force[t], non_blocking=nb)
        self.ft_sensors.copy_(s.ft_sensors[t], non_blocking=nb)

    def use_sparse_upload(self, params) -> None:
        """Host-resident sequences: stage each state with lg_upload_sim_state (whole small tensors, and of the
        rigid-body tensor only the run of bodies holding the fingertips) instead of five full copies."""
        from . import _native as nat
        if self.seq.dof_state.device.type != "cpu" or not self.seq.dof_state.is_pinned():
            return
        lib = nat.load()
        S = nat.LgSimState(*(x.data_ptr() for x in (self.dof_state, self.root_state, self.rigid_body,
                                                     self.dof_force, self.ft_sensors)))

        def upload(t):
            s = self.seq
            H = nat.LgHostStep()
            H.dof_state_host, H.root_state_host = s.dof_state[t].data_ptr(), s.root_state[t].data_ptr()
            H.rigid_body_host = s.rigid_body[t].data_ptr()
            H.dof_force_host, H.ft_sensors_host = s.dof_force[t].data_ptr(), s.ft_sensors[t].data_ptr()
            nat.check(lib.lg_upload_sim_state(params, S, H, torch.cuda.current_stream(self.device).cuda_stream),
                      "lg_upload_sim_state")
        self._uploader = upload

    def enable_goal_integration(self, params, dt: float) -> None:
        """Moving-goal task (ref trifinger_env.py:947, :1267-1284): instead of taking the goal actor's root rows from
        the sequence, integrate them like a simulator would (lg_integrate_goal: p += v dt, q <- exp(w dt/2) q), so
        the goal really rotates with the angular velocity the env imposes each step."""
        from . import _native as nat
        lib = nat.load()
        S = nat.LgSimState(*(x.data_ptr() for x in (self.dof_state, self.root_state, self.rigid_body,
                                                     self.dof_force, self.ft_sensors)))
        goal_rows = self.root_state.view(self.num_envs, self.actors_per_env, 13)[:, self.slots[2]]

        def integrate(load):
            keep = goal_rows.clone()
            load()
            goal_rows.copy_(keep)
            nat.check(lib.lg_integrate_goal(params, S, float(dt), torch.cuda.current_stream(self.device).cuda_stream),
                      "lg_integrate_goal")
        self._integrate = integrate

    # -- the slice of the gym API the path uses ---------------------------------------
    def simulate(self) -> None:
        if self.before_simulate is not None:
            self.before_simulate(self)
        t = self.cursor % self.seq.num_steps
        if self._integrate is not None:
            self._integrate(lambda: self._load(t))
        else:
            self._load(t)
        self.cursor += 1
        self.frame_count += 1

    def begin_step(self) -> int:
        """Advance the simulator clock WITHOUT staging the new state; returns the sequence index of that
        state.  For callers that upload it themselves in pieces (host_pipeline.HostPipeline)."""
        if self.before_simulate is not None:
            self.before_simulate(self)
        t = self.cursor % self.seq.num_steps
        self.cursor += 1
        self.frame_count += 1
        return t

    def get_frame_count(self) -> int:
        return self.frame_count

    def set_dof_actuation_force_tensor(self, torque: torch.Tensor) -> None:
        self.applied_torque = torque

    def set_dof_state_tensor_indexed(self, indices: torch.Tensor, count) -> None:
        """Reset rows are already in `dof_state` (simulator memory); a real backend is notified here."""

    def set_actor_root_state_tensor_indexed(self, indices: torch.Tensor, count) -> None:
        """As above for `root_state`."""


class HostZeroCopySim:
    """Simulator whose state lives in pinned HOST memory and is read by the kernels in place.

    The situation of the reference's default CPU pipeline (`use_gpu_pipeline: False`, ref
    envs/env_base.py:60): PhysX keeps its tensors on the host.  Instead of staging them through a
    device copy every step, the env kernels read the rows they need (and write reset rows) directly
    over PCIe — pinned memory is device-addressable under UVA.  `simulate()` makes the next state of
    the sequence current by swapping tensor references (`rebinds_tensors`), which is what a host
    simulator that double-buffers its output would do; nothing is copied on the host either.
    """

    rebinds_tensors = True

    def __init__(self, seq: StateSequence):
        assert seq.dof_state.device.type == "cpu" and seq.dof_state.is_pinned(), "needs a pinned host sequence"
        self.seq = seq
        self.num_envs = seq.num_envs
        self.frame_count = 0
        self.cursor = 0
        self.fingertip_bodies = (6, 11, 16)
        self.bodies_per_env, self.actors_per_env = 20, 4
        self.slots = (0, 2, 3)
        self.applied_torque = None
        self._select(0)

    def _select(self, t: int) -> None:
        s = self.seq
        self.dof_state, self.root_state, self.rigid_body = s.dof_state[t], s.root_state[t], s.rigid_body[t]
        self.dof_force, self.ft_sensors = s.dof_force[t], s.ft_sensors[t]

    def simulate(self) -> None:
        self._select(self.cursor % self.seq.num_steps)
        self.cursor += 1
        self.frame_count += 1

    def get_frame_count(self) -> int:
        return self.frame_count

    def set_dof_actuation_force_tensor(self, torque: torch.Tensor) -> None:
        self.applied_torque = torque

    def set_dof_state_tensor_indexed(self, indices, count) -> None:
        pass

    def set_actor_root_state_tensor_indexed(self, indices, count) -> None:
        pass
