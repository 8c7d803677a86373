"""RL-facing vectorised wrapper (ref leibnizgym/wrappers/vec_task.py:26-170).

Same surface as the reference's `VecTask` / `VecTaskPython`.  The three clamps the
reference runs as separate ATen passes per step (actions +-clip_actions, obs and
states +-clip_obs) are folded into the env's fused kernels: the wrapper switches
the clipped outputs on and hands those buffers out.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from ..env import IsaacEnvBase

try:  # gym is optional here; the reference needs it only for the Box descriptors
    from gym import spaces as _spaces
    Box = _spaces.Box
except Exception:  # pragma: no cover - gym is not part of this image
    class Box:
        """Minimal stand-in for gym.spaces.Box (bounds + shape)."""

        def __init__(self, low, high):
            self.low, self.high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
            self.shape, self.dtype = self.low.shape, np.float32

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, float32)"


class VecTask:
    def __init__(self, task: IsaacEnvBase, rl_device: str, clip_obs: float = 5.0, clip_actions: float = 1.0):
        assert isinstance(task, IsaacEnvBase)
        self._task = task
        self._clip_obs = float(clip_obs)
        self._clip_actions = float(clip_actions)
        self._rl_device = rl_device
        self._obs_space = Box(np.full(self.num_obs, -self._clip_obs), np.full(self.num_obs, self._clip_obs))
        self._state_space = Box(np.full(self.num_states, -self._clip_obs), np.full(self.num_states, self._clip_obs))
        self._act_space = Box(np.full(self.num_actions, -self._clip_actions),
                              np.full(self.num_actions, self._clip_actions))

    def __str__(self) -> str:
        return (f"Vectorized Environment around task: {type(self._task).__name__} \n"
                f"\t Number of instances   : {self.num_envs} \n"
                f"\t Number of observations: {self.num_obs} \n"
                f"\t Number of states      : {self.num_states} \n"
                f"\t Number of actions     : {self.num_actions} \n"
                f"\t Observation clipping  : {self._clip_obs} \n"
                f"\t Actions clipping      : {self._clip_actions} \n")

    def get_number_of_agents(self) -> int:
        if hasattr(self._task, "get_number_of_agents"):
            return self._task.get_number_of_agents()
        return 1

    @property
    def num_envs(self) -> int:
        return self._task.get_num_instances()

    @property
    def num_states(self) -> int:
        return self._task.get_state_dim()

    @property
    def num_obs(self) -> int:
        return self._task.get_obs_dim()

    @property
    def num_actions(self) -> int:
        return self._task.get_action_dim()

    @property
    def observation_space(self):
        return self._obs_space

    @property
    def state_space(self):
        return self._state_space

    @property
    def action_space(self):
        return self._act_space

    def dump_config(self, filename: str):
        self._task.dump_config(filename)

    def reset(self) -> torch.Tensor:
        raise NotImplementedError

    def step(self, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, dict]:
        raise NotImplementedError


class VecTaskPython(VecTask):
    def __init__(self, task: IsaacEnvBase, rl_device: str, clip_obs: float = 5.0, clip_actions: float = 1.0,
                 host_pipeline_chunks: int = 0, obs_dtype: torch.dtype = torch.float32, blocking: bool = True):
        """`host_pipeline_chunks` > 0 (host-resident simulator and learner, rl_device 'cpu'): step through
        host_pipeline.HostPipeline, which overlaps the state upload, the kernels and the result download.
        `obs_dtype=torch.bfloat16`: hand the learner bf16 observations / states, emitted by the fused pass itself.
        `blocking` (host-resident results only): `step()` returns once the results ARE in the returned pinned host
        tensors and the kernels have consumed the caller's action buffer — what the reference's blocking
        `.to(rl_device)` copies guarantee (ref wrappers/vec_task.py:164-170).  `blocking=False` returns as soon as
        the work is enqueued; the caller then owns the stream synchronisation before it reads the results or
        overwrites a pinned action buffer it passed in."""
        super().__init__(task, rl_device, clip_obs, clip_actions)
        self._pipeline = None
        self._blocking = bool(blocking)
        if obs_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("obs_dtype must be torch.float32 or torch.bfloat16")
        self._bf16 = obs_dtype == torch.bfloat16
        if self._bf16 and (host_pipeline_chunks > 0 or not hasattr(task, "enable_bf16_outputs")):
            raise ValueError("bf16 outputs need the fused device path")
        self._fused = hasattr(task, "enable_clipped_outputs")
        if self._fused:
            if host_pipeline_chunks > 0:
                from ..host_pipeline import HostPipeline
                task.enable_clipped_outputs(self._clip_obs, self._clip_actions)
                self._pipeline = HostPipeline(task, chunks=host_pipeline_chunks)
            elif torch.device(rl_device).type == "cpu" and hasattr(task, "enable_host_outputs"):
                # host-side learner: the clipped results are stored straight into pinned host memory
                task.enable_host_outputs(self._clip_obs, self._clip_actions)
            else:
                task.enable_clipped_outputs(self._clip_obs, self._clip_actions)
            if self._bf16:
                task.enable_bf16_outputs()

    def _host_results_ready(self) -> None:
        """Host-resident results: wait until the step's kernels / copies have completed (see `blocking`)."""
        if self._blocking:
            torch.cuda.current_stream(self._task._torch_device).synchronize()

    def get_state(self) -> torch.Tensor:
        if self._pipeline is not None:
            return self._pipeline.h_states
        if self._bf16:
            return self._task._states_bf16.to(self._rl_device)
        if self._fused and self._task._states_clipped is not None:
            return self._task._states_clipped.to(self._rl_device)
        return torch.clamp(self._task.states_buf, -self._clip_obs, self._clip_obs).to(self._rl_device)

    def reset(self) -> torch.Tensor:
        obs = torch.clamp(self._task.reset(), -self._clip_obs, self._clip_obs)
        return (obs.to(torch.bfloat16) if self._bf16 else obs).to(self._rl_device)

    def step(self, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, dict]:
        if self._task.visualize:
            self._task.render()
        if self._pipeline is not None:
            out = self._pipeline.step(actions)    # pinned host tensors
            self._host_results_ready()
            return out
        if self._fused:
            # action clamp, obs clamp and states clamp all happen inside the two fused launches
            _, rew, is_done, info = self._task.step(actions)
            obs = self._task._obs_bf16 if self._bf16 else self._task._obs_clipped
            if getattr(self._task, "_host_io", False):
                self._host_results_ready()        # the kernels wrote into pinned host memory: .to('cpu') below copies nothing
        else:
            obs, rew, is_done, info = self._task.step(torch.clamp(actions, -self._clip_actions, self._clip_actions))
            obs = torch.clamp(obs, -self._clip_obs, self._clip_obs)
        return obs.to(self._rl_device), rew.to(self._rl_device), is_done.to(self._rl_device), info
