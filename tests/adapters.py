"""Uniform test-side view of an env (oracle or CUDA) for the golden replay."""
from __future__ import annotations

import numpy as np
import torch

from oracle.trifinger_oracle import OracleEnv, OracleSim


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class _CapturingSim(OracleSim):
    def simulate(self):
        self.pre_sim_dof = self.dof.clone()
        self.pre_sim_root = self.root.clone()
        super().simulate()


class OracleAdapter:
    """OracleEnv behind the reference's attribute names."""

    def __init__(self, config, seq):
        from leibnizgym_b200.config import resolve_config
        self.cfg = resolve_config(config)
        self.N = self.cfg["num_instances"]
        self.sim = _CapturingSim(seq, self.N)
        self.env = OracleEnv(self.cfg, self.sim)

    # reference-style buffer names used by the replay driver
    @property
    def _reset_buf(self):
        return self.env.reset_buf

    @_reset_buf.setter
    def _reset_buf(self, v):
        self.env.reset_buf = v

    @property
    def _goal_reset_buf(self):
        return self.env.goal_reset_buf

    @_goal_reset_buf.setter
    def _goal_reset_buf(self, v):
        self.env.goal_reset_buf = v

    def inject_draws(self, reset=None, goal=None):
        self.env.inject_draws(reset=reset, goal=goal)

    def reset(self):
        self.env.index_lists = {}
        self.env.last_ids = (torch.arange(self.N), torch.zeros(0, dtype=torch.long))
        return self.env.reset()

    def step(self, action):
        self.env.index_lists = {}
        return self.env.step(action)

    def observe(self, key):
        e, s, N = self.env, self.sim, self.N
        if key == "reset_in":
            return _np(e.reset_buf)
        if key == "goal_reset_in":
            return _np(e.goal_reset_buf)
        if key == "reset_ids":
            return _np(e.last_ids[0])
        if key == "goal_reset_ids":
            return _np(e.last_ids[1])
        if key == "pre_sim_dof":
            return _np(s.pre_sim_dof)
        if key == "pre_sim_obj_root":
            return _np(s.pre_sim_root.view(N, 4, 13)[:, 2])
        if key == "pre_sim_goal_root":
            return _np(s.pre_sim_root.view(N, 4, 13)[:, 3])
        if key == "dof_index_list":
            return _np(e.index_lists["dof"])
        if key == "root_index_list0":
            return _np(e.index_lists["root_reset"] if "root_reset" in e.index_lists else e.index_lists["root_goal"])
        if key == "root_index_list1":
            return _np(e.index_lists["root_goal"])
        if key == "root_index_list_move":
            return _np(e.index_lists["root_move"])
        if key == "applied_torque":
            return _np(e.applied_torque)
        if key == "goal_pose":
            return _np(e.goal_poses)
        if key == "goal_movement":
            return _np(e.goal_movement)
        if key == "action_buf":
            return _np(e.action_buf)
        if key == "obs":
            return _np(e.obs_buf)
        if key == "states":
            return _np(e.states_buf)
        if key == "reset_buf":
            return _np(e.reset_buf)
        if key == "goal_reset_buf":
            return _np(e.goal_reset_buf)
        if key == "steps_count":
            return _np(e.steps_count_buf)
        if key == "successes":
            return _np(e.successes)
        if key == "terms":
            return _np(e.last_terms)
        if key == "reward":
            return _np(e.reward_buf)
        if key == "sched_step":
            return np.asarray(e.env_steps_count)
        if key == "info":
            return {k: float(v) for k, v in e.step_info.items()}
        raise KeyError(key)


class CudaAdapter:
    """The CUDA TrifingerEnv behind the same test-side view."""

    def __init__(self, config, seq, device="cuda:0", fused=True):
        from leibnizgym_b200.env import TrifingerEnv
        from leibnizgym_b200.sim import SyntheticSim
        self.N = config["num_instances"]
        self.sim = SyntheticSim(seq.to(device), device=device)
        self.sim.before_simulate = self._capture
        self.sim.set_actor_root_state_tensor_indexed = self._root_indexed   # record what the simulator is told
        self.env = TrifingerEnv(config=config, device=device, verbose=False, sim=self.sim)
        self.env.enable_term_rewards(True)
        self.fused = fused
        self.k_reset = self.N
        self.k_goal = 0
        self._ids = (torch.arange(self.N), torch.zeros(0, dtype=torch.long))

    def _root_indexed(self, indices, count):
        self.sim.last_root_indexed = indices[:int(count)].clone()

    def _capture(self, sim):
        self.pre_sim_dof = sim.dof_state.clone()
        self.pre_sim_root = sim.root_state.clone()

    @property
    def _reset_buf(self):
        return self.env._reset_buf

    @_reset_buf.setter
    def _reset_buf(self, v):
        if v is not self.env._reset_buf:
            self.env._reset_buf.copy_(v)

    @property
    def _goal_reset_buf(self):
        return self.env._goal_reset_buf

    @_goal_reset_buf.setter
    def _goal_reset_buf(self, v):
        if v is not self.env._goal_reset_buf:
            self.env._goal_reset_buf.copy_(v)

    def inject_draws(self, reset=None, goal=None):
        self.env.inject_draws(reset=reset, goal=goal)

    def reset(self):
        out = self.env.reset()
        self._ids = (torch.arange(self.N), torch.zeros(0, dtype=torch.long))
        return out

    def step(self, action):
        if self.fused:
            out = self.env.step(action.to(self.env.device))
            self._ids = (self.env.reset_env_ids.cpu(), self.env.goal_reset_env_ids.cpu())
        else:  # the reference's hook-by-hook sequencing through the individual C-ABI entry points
            from leibnizgym_b200.env import IsaacEnvBase
            r = torch.nonzero(self.env._reset_buf).view(-1).cpu()
            g = torch.nonzero(self.env._goal_reset_buf).view(-1).cpu()
            out = IsaacEnvBase.step(self.env, action.to(self.env.device))
            self._ids = (r, g)
        return out

    def observe(self, key):
        e, N = self.env, self.N
        kr, kg = len(self._ids[0]), len(self._ids[1])
        if key == "reset_in":
            return _np(e._reset_buf)
        if key == "goal_reset_in":
            return _np(e._goal_reset_buf)
        if key == "reset_ids":
            return _np(self._ids[0])
        if key == "goal_reset_ids":
            return _np(self._ids[1])
        if key == "pre_sim_dof":
            return _np(self.pre_sim_dof)
        if key == "pre_sim_obj_root":
            return _np(self.pre_sim_root.view(N, 4, 13)[:, 2])
        if key == "pre_sim_goal_root":
            return _np(self.pre_sim_root.view(N, 4, 13)[:, 3])
        if key == "dof_index_list":
            return _np(e._robot_indices[:kr])
        if key == "root_index_list0":
            return _np(e._reset_root_indices[:3 * kr] if kr else e._goal_root_indices[:kg])
        if key == "root_index_list1":
            return _np(e._goal_root_indices[:kg])
        if key == "root_index_list_move":
            return _np(self.sim.last_root_indexed)
        if key == "applied_torque":
            return _np(e._applied_torque)
        if key == "goal_pose":
            return _np(e._object_goal_poses_buf)
        if key == "goal_movement":
            return _np(e._object_goal_movement_buf)
        if key == "action_buf":
            return _np(e._action_buf)
        if key == "obs":
            return _np(e._obs_buf)
        if key == "states":
            return _np(e._states_buf)
        if key == "reset_buf":
            return _np(e._reset_buf)
        if key == "goal_reset_buf":
            return _np(e._goal_reset_buf)
        if key == "steps_count":
            return _np(e._steps_count_buf)
        if key == "successes":
            return _np(e._successes)
        if key == "terms":
            return _np(e._term_rewards[:6])
        if key == "reward":
            return _np(e._reward_buf)
        if key == "sched_step":
            return np.asarray(e.env_steps_count)
        if key == "info":
            return {k: float(v) for k, v in e._step_info.items()}
        raise KeyError(key)
