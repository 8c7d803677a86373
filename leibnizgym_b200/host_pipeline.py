"""Chunked step for a HOST-resident simulator: overlap upload, compute and download.

When the simulator state lives in host memory (the reference's default `use_gpu_pipeline:
False`, ref envs/env_base.py:60) a step is PCIe-bound: ~16 MB up, ~10 MB down, ~13 us of
kernels.  PCIe is full duplex, so the shard is cut into a few env ranges and pipelined on
three streams:

    main   action H2D | lg_pre_physics (all envs) | post(chunk 0) | post(chunk 1) | ...
    up                                     upload(chunk 0) | upload(chunk 1) | ...
    down                                                    download(chunk 0) | ...

`lg_post_physics` runs per chunk on offset pointers (`LgParams.num_envs/env_offset` describe
the chunk, `stats_num_envs` keeps the statistics' denominator at the shard size); the
pre-physics pass stays whole because its ordered compaction spans the shard.  Results are
bit-identical to the un-chunked step (tests/test_cuda_vs_oracle.py).
"""
from __future__ import annotations

import torch

from . import _native as nat
from .env import TrifingerEnv
from .sim import SyntheticSim


class HostPipeline:
    def __init__(self, env: TrifingerEnv, chunks: int = 2):
        sim = env._sim
        if not isinstance(sim, SyntheticSim) or sim.seq.dof_state.device.type != "cpu" or not sim.seq.dof_state.is_pinned():
            raise ValueError("HostPipeline needs a SyntheticSim over a pinned host sequence")
        if env._obs_clipped is None:
            raise ValueError("enable_clipped_outputs() first: the pipeline downloads the clipped results")
        self.env, self.sim, self.lib = env, sim, env._lib
        N, dev = env.num_instances, env._torch_device
        A, od, sd = env.get_action_dim(), env.get_obs_dim(), env.get_state_dim()
        self.asym = sd > 0
        self.chunks = int(max(1, min(16, chunks)))
        pin = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt).pin_memory()  # noqa: E731
        self.h_obs, self.h_reward, self.h_dones = pin(N, od), pin(N), pin(N, dt=torch.bool)
        self.h_states = pin(N, sd) if self.asym else torch.zeros(N, 0)
        self.act_dev = torch.zeros((N, A), device=dev)
        self._h_action = pin(N, A)
        H = nat.LgHostStep()
        H.obs_host, H.reward_host, H.dones_host = self.h_obs.data_ptr(), self.h_reward.data_ptr(), self.h_dones.data_ptr()
        H.states_host = self.h_states.data_ptr() if self.asym else None
        H.action_staging = self.act_dev.data_ptr()
        self._H = H
        self.up, self.down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def step(self, action_host: torch.Tensor):
        """One env step; all copies and launches are enqueued by a single native call
        (lg_step_host_pipelined) — issuing ~12 operations per chunk from Python would cost more host
        time than the PCIe transfers take."""
        env, lib = self.env, self.lib
        main = torch.cuda.current_stream(env._torch_device)
        shape = (env.num_instances, env.get_action_dim())
        if tuple(action_host.size()) != shape:
            raise ValueError(f"Invalid shape for tensor `action`. Input: {tuple(action_host.size())} != {shape}.")
        if action_host.device.type != "cpu" or not action_host.is_pinned() or action_host.dtype != torch.float:
            self._h_action.copy_(action_host)           # slow path: stage through our own pinned buffer
            action_host = self._h_action
        self.sim.set_dof_actuation_force_tensor(env._applied_torque)
        t = self.sim.begin_step()                       # the simulator produced state t (in host memory)
        s = self.sim.seq
        H = self._H
        H.dof_state_host, H.root_state_host = s.dof_state[t].data_ptr(), s.root_state[t].data_ptr()
        H.rigid_body_host = s.rigid_body[t].data_ptr()
        H.dof_force_host, H.ft_sensors_host = s.dof_force[t].data_ptr(), s.ft_sensors[t].data_ptr()
        H.action_host = action_host.data_ptr()
        env._P.fuse_bookkeeping = 1
        nat.check(lib.lg_step_host_pipelined(env._P, env._S, env._B, H, float(env.env_steps_count), self.chunks,
                                             main.cuda_stream, self.up.cuda_stream, self.down.cuda_stream),
                  "lg_step_host_pipelined")
        env._clear_injection()
        env._step_info = env._make_info()
        return self.h_obs, self.h_reward, self.h_dones, env._step_info
