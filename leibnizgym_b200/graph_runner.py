"""CUDA-graph replay of consecutive env steps over a ring of simulator states.

The per-step sequence (lg_pre_physics -> [simulator] -> lg_post_physics) is two
launches of a few microseconds each, so host launch cost would dominate at 16k envs.
`GraphRunner` captures C consecutive steps into one CUDA graph.  Step t reads the
simulator tensors of ring slot t % R and writes obs/states into output slot t % R
(the layout a rollout buffer [horizon, N, D] has anyway), so successive steps touch
different HBM lines and nothing is served from L2 by accident of the benchmark.

Schedule step and RNG epoch come from the device-side LgControl block
(`use_device_clock`), so a replayed graph still advances `env_steps_count` and draws
fresh random numbers.  Resets write their rows into the ring slot the simulator would
consume next, so a ring that wraps around replays slightly different states on later
passes (harmless for timing; tests compare like with like).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _native as nat
from .env import TrifingerEnv
from .synthetic import StateSequence


class GraphRunner:
    def __init__(self, env: TrifingerEnv, ring: StateSequence, rotate_outputs: bool = True,
                 inject_reset_masks: Optional[torch.Tensor] = None, device_clock: bool = True,
                 inject_goal_masks: Optional[torch.Tensor] = None, chain_pre: bool = True):
        assert ring.dof_state.is_cuda, "the ring must be device resident"
        self.env, self.ring = env, ring
        self.R = ring.num_steps
        self.lib = env._lib
        N, dev = env.num_instances, env._torch_device
        self.obs_slots = torch.zeros((self.R if rotate_outputs else 1, N, env.get_obs_dim()), device=dev)
        sd = env.get_state_dim()
        self.state_slots = torch.zeros((self.R if rotate_outputs else 1, N, sd), device=dev) if sd else None
        # [R, N] bool or None: OR-ed into _reset_buf / _goal_reset_buf at the start of each step, by lg_pre_physics
        # itself (LgBuffers.force_reset / force_goal_reset) — no separate pass over the flags
        self.reset_masks = inject_reset_masks
        self.goal_masks = inject_goal_masks
        # The runner's actions and joint states sit in the ring long before they are used: the contract of
        # lg_pre_physics_chained holds, so the pre-physics pass of step t+1 is chained to the post-physics pass of step
        # t.  With injected reset masks (reset-heavy workloads) each of its CTAs gets an SM of its own.
        self.chain_pre = chain_pre
        self.exclusive_sm = int(inject_reset_masks is not None)
        self.P = nat.LgParams.from_buffer_copy(env._P)
        self.P.use_device_clock = int(device_clock)
        self.P.fuse_bookkeeping = 1
        self.device_clock = device_clock
        self.frame0 = env._sim.get_frame_count()   # host clock: frames before the runner's first step
        if device_clock:
            env._control[1] = self.frame0           # LgControl.frame_count
        self._S: List[nat.LgSimState] = []
        self._B: List[nat.LgBuffers] = []
        for t in range(self.R):
            self._S.append(nat.LgSimState(ring.dof_state[t].data_ptr(), ring.root_state[t].data_ptr(),
                                          ring.rigid_body[t].data_ptr(), ring.dof_force[t].data_ptr(),
                                          ring.ft_sensors[t].data_ptr()))
            b = nat.LgBuffers()
            C.memmove(C.byref(b), C.byref(env._B), C.sizeof(b))
            o = t if rotate_outputs else 0
            b.obs = self.obs_slots[o].data_ptr()
            b.states = self.state_slots[o].data_ptr() if self.state_slots is not None else None
            b.obs_clipped = b.states_clipped = None
            b.term_rewards = None
            # the pre-physics pass of step t runs on the buffers of slot t-1 (see _launch_step): masks of step t
            nxt = (t + 1) % self.R
            b.force_reset = self.reset_masks[nxt].data_ptr() if self.reset_masks is not None else None
            b.force_goal_reset = self.goal_masks[nxt].data_ptr() if self.goal_masks is not None else None
            self._B.append(b)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.steps_per_graph = 0
        self.t = 0   # ring cursor of the next eager step
        self.t0 = 0  # ring cursor of the runner's first step (host clock only)

    # -- one step, eager (also what gets captured) -----------------------------------------
    def _launch_step(self, t: int, stream: int, post_only: bool = False, pre_only: bool = False) -> None:
        s = t % self.R
        prev = (t - 1) % self.R
        if not post_only:
            # resets write into the tensors the simulator consumes next (slot of the previous state)
            if self.chain_pre:
                nat.check(self.lib.lg_pre_physics_chained(self.P, self._S[prev], self._B[prev], self.ring.action[s].data_ptr(),
                                                          self.exclusive_sm, stream), "lg_pre_physics_chained")
            else:
                nat.check(self.lib.lg_pre_physics(self.P, self._S[prev], self._B[prev],
                                                  self.ring.action[s].data_ptr(), stream), "lg_pre_physics")
        if pre_only:
            return
        sched = 0.0 if self.device_clock else float((self.frame0 + (t - self.t0) + 1) * self.env._global_N)
        nat.check(self.lib.lg_post_physics(self.P, self._S[s], self._B[s], sched, stream), "lg_post_physics")

    def step_eager(self, n: int = 1, post_only: bool = False, pre_only: bool = False) -> None:
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(n):
            self._launch_step(self.t, stream, post_only, pre_only)
            self.t += 1

    # -- graph ----------------------------------------------------------------------------------
    def capture(self, steps: int, post_only: bool = False, pre_only: bool = False) -> torch.cuda.CUDAGraph:
        """Captures `steps` consecutive steps starting at ring slot 0 (steps % R == 0 keeps replays
        aligned with the ring).  `post_only` / `pre_only`: one of the two kernels alone, back to back
        (per-kernel timing).  Returns the graph (also kept as `self.graph`)."""
        self.step_eager(2, post_only, pre_only)  # warm: module load, first-touch
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            stream = torch.cuda.current_stream().cuda_stream
            for t in range(steps):
                self._launch_step(t, stream, post_only, pre_only)
        self.graph, self.steps_per_graph = g, steps
        return g

    def replay(self, times: int = 1) -> None:
        for _ in range(times):
            self.graph.replay()

    @property
    def launches_per_step(self) -> int:
        return 2
