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
    def __init__(self, env: TrifingerEnv, chunks: int = 1, obs_from_states: bool = True):
        """`obs_from_states`: with asymmetric observations and no observation noise the actor observation IS the
        first `obs_dim` columns of the critic state (ref trifinger_env.py:995-1051: the state bounds start with
        the observation bounds, the state tensor with the observation's groups), so only the states cross PCIe
        and `h_obs` is the strided view `h_states[:, :obs_dim]` — same values, 26 % fewer bytes down."""
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
        self.shared_obs = bool(obs_from_states and self.asym and not env._P.dr_activate)
        pin = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt).pin_memory()  # noqa: E731
        self.h_reward, self.h_dones = pin(N), pin(N, dt=torch.bool)
        self.h_states = pin(N, sd) if self.asym else torch.zeros(N, 0)
        self.h_obs = self.h_states[:, :od] if self.shared_obs else pin(N, od)
        self.act_dev = torch.zeros((N, A), device=dev)
        self._h_action = pin(N, A)
        H = nat.LgHostStep()
        H.reward_host, H.dones_host = self.h_reward.data_ptr(), self.h_dones.data_ptr()
        H.obs_host = None if self.shared_obs else self.h_obs.data_ptr()
        H.states_host = self.h_states.data_ptr() if self.asym else None
        H.action_staging = self.act_dev.data_ptr()
        self._H = H
        self._shape = (N, A)
        s = sim.seq   # host pointers of every state of the sequence, resolved once
        self._ptrs = [tuple(x[t].data_ptr() for x in (s.dof_state, s.root_state, s.rigid_body, s.dof_force, s.ft_sensors))
                      for t in range(s.num_steps)]
        self.up, self.down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._streams = (self.up.cuda_stream, self.down.cuda_stream)

    def step(self, action_host: torch.Tensor):
        """One env step; all copies and launches are enqueued by a single native call
        (lg_step_host_pipelined) — issuing ~12 operations per chunk from Python would cost more host
        time than the PCIe transfers take."""
        env = self.env
        if tuple(action_host.shape) != self._shape:
            raise ValueError(f"Invalid shape for tensor `action`. Input: {tuple(action_host.size())} != {self._shape}.")
        if action_host.device.type != "cpu" or not action_host.is_pinned() or action_host.dtype != torch.float:
            self._h_action.copy_(action_host)           # slow path: stage through our own pinned buffer
            action_host = self._h_action
        self.sim.set_dof_actuation_force_tensor(env._applied_torque)
        t = self.sim.begin_step()                       # the simulator produced state t (in host memory)
        H = self._H
        (H.dof_state_host, H.root_state_host, H.rigid_body_host, H.dof_force_host, H.ft_sensors_host) = self._ptrs[t]
        H.action_host = action_host.data_ptr()
        env._P.fuse_bookkeeping = 1
        main = torch.cuda.current_stream(env._torch_device).cuda_stream
        nat.check(self.lib.lg_step_host_pipelined(env._P, env._S, env._B, H, float(env.env_steps_count), self.chunks,
                                                  main, *self._streams), "lg_step_host_pipelined")
        env._clear_injection()
        env._notify_resets()
        env._step_info = env._make_info()
        return self.h_obs, self.h_reward, self.h_dones, env._step_info
