"""Runs the UNMODIFIED reference (`/root/reference`) on synthetic simulator state.

Only usable in the build container (the reference does not travel to the GPU
box).  It is the generator of the fixtures in this directory — nothing in the
test-suite, smoke() or bench.py imports it at run time.

Recipe: stub `termcolor`, `isaacgym` and `gym`, serve the
simulator API from `FakeGym`, then construct the real `TrifingerEnv`.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("LEIBNIZ_REFERENCE_ROOT", "/root/reference")

FINGERTIP_LINKS = {"finger_tip_link_0": 6, "finger_tip_link_120": 11, "finger_tip_link_240": 16}


class FakeGym:
    """Minimal stand-in for `gymapi.acquire_gym()` that plays back a StateSequence."""

    def __init__(self, seq, num_envs):
        self.seq = seq
        self.N = num_envs
        self.cursor = 0            # next sequence step `simulate` will load
        self.frame = 0
        self.actor_counter = 0
        self.dof_names = []
        self.calls = {}
        # simulator-owned tensors (the env wraps views of these, zero copy)
        self.root = torch.zeros(4 * num_envs, 13)
        self.dof = torch.zeros(9 * num_envs, 2)
        self.rb = torch.zeros(20 * num_envs, 13)
        self.dof_force = torch.zeros(9 * num_envs)
        self.ft = torch.zeros(3 * num_envs, 6)
        # captured side effects
        self.applied_torque = None
        self.dof_indexed = None
        self.root_indexed = []
        self.pre_sim_dof = None
        self.pre_sim_root = None
        self._load(0)  # what refresh_* would show before the first simulate

    # -- state playback -------------------------------------------------
    def _load(self, t):
        s = self.seq
        self.root.copy_(s.root_state[t])
        self.dof.copy_(s.dof_state[t].reshape(-1, 2))
        self.rb.copy_(s.rigid_body[t].reshape(-1, 13))
        self.dof_force.copy_(s.dof_force[t].reshape(-1))
        self.ft.copy_(s.ft_sensors[t].reshape(-1, 6))

    def simulate(self, sim):
        self.pre_sim_dof = self.dof.clone()
        self.pre_sim_root = self.root.clone()
        self._load(self.cursor)
        self.cursor += 1
        self.frame += 1

    def get_frame_count(self, sim):
        return self.frame

    # -- assets -----------------------------------------------------------
    def load_asset(self, sim, root, file, options):
        if "trifingerpro" in file:
            return "robot"
        if "high_table" in file:
            return "stage"
        return "cube"

    def get_asset_rigid_body_count(self, asset):
        return 17 if asset == "robot" else 1

    def get_asset_rigid_shape_count(self, asset):
        return 17 if asset == "robot" else 1

    def get_asset_dof_count(self, asset):
        return 9 if asset == "robot" else 0

    def get_asset_dof_properties(self, asset):
        keys = ("driveMode", "stiffness", "damping", "effort", "velocity", "lower", "upper")
        return {k: np.zeros(9, dtype=np.float32) for k in keys}

    def get_asset_rigid_shape_properties(self, asset):
        return [SimpleNamespace(friction=0.0, torsion_friction=0.0, restitution=0.0)]

    def find_asset_rigid_body_index(self, asset, name):
        return FINGERTIP_LINKS[name]

    def find_asset_dof_index(self, asset, name):
        if name not in self.dof_names:
            self.dof_names.append(name)
        return self.dof_names.index(name)

    # -- actors -----------------------------------------------------------
    def create_env(self, *a):
        return object()

    def create_actor(self, *a):
        handle = self.actor_counter
        self.actor_counter += 1
        return handle

    def get_actor_index(self, env, handle, domain):
        return handle

    # -- tensor API -------------------------------------------------------
    def acquire_actor_root_state_tensor(self, sim):
        return self.root

    def acquire_dof_state_tensor(self, sim):
        return self.dof

    def acquire_rigid_body_state_tensor(self, sim):
        return self.rb

    def acquire_dof_force_tensor(self, sim):
        return self.dof_force

    def acquire_force_sensor_tensor(self, sim):
        return self.ft

    def set_dof_actuation_force_tensor(self, sim, torque):
        self.applied_torque = torque.clone()
        return True

    def set_dof_state_tensor_indexed(self, sim, state, indices, n):
        self.dof_indexed = indices.clone()
        return True

    def set_actor_root_state_tensor_indexed(self, sim, state, indices, n):
        self.root_indexed.append(indices.clone())
        return True

    def __getattr__(self, name):
        def _noop(*a, **k):
            self.calls[name] = self.calls.get(name, 0) + 1
            return MagicMock()
        return _noop


def install_stubs(fake: FakeGym):
    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    sys.modules["termcolor"] = tc
    gymapi = MagicMock()
    gymapi.INVALID_HANDLE = -1
    gymapi.SIM_PHYSX = 1
    gymapi.SIM_FLEX = 0
    gymapi.acquire_gym = lambda: fake
    gymtorch = MagicMock()
    gymtorch.wrap_tensor = lambda t: t
    gymtorch.unwrap_tensor = lambda t: t
    isaac = MagicMock()
    isaac.gymapi = gymapi
    isaac.gymtorch = gymtorch
    sys.modules["isaacgym"] = isaac
    sys.modules["isaacgym.gymapi"] = gymapi
    sys.modules["isaacgym.gymtorch"] = gymtorch
    gym = MagicMock()
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = gym.spaces
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class DrawRecorder:
    """Records every torch.rand / torch.randn the reference issues (PYTORCH_JIT=0 only)."""

    def __init__(self):
        self.log = []
        self._rand, self._randn = torch.rand, torch.randn

    def __enter__(self):
        def rand(*a, **k):
            out = self._rand(*a, **k)
            self.log.append(("u", out.clone()))
            return out

        def randn(*a, **k):
            out = self._randn(*a, **k)
            self.log.append(("n", out.clone()))
            return out
        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn

    def take(self):
        out, self.log = self.log, []
        return out


def build_reference_env(config: dict, seq, quiet: bool = True):
    """Construct the reference's TrifingerEnv over a FakeGym that plays `seq`."""
    fake = FakeGym(seq, config["num_instances"])
    install_stubs(fake)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink if quiet else sys.stdout):
        from leibnizgym.envs.trifinger.trifinger_env import TrifingerEnv
        env = TrifingerEnv(config=config, device="cpu", verbose=False, visualize=False)
    return env, fake


def stash_goal_before_movement(env):
    """Moving-goal scenarios: keep a copy of the goal pose the reward terms saw, by wrapping the (name-mangled)
    post-movement hook on the instance.  The reference's behaviour is unchanged."""
    inner = env._TrifingerEnv__update_goal_movement_post

    def wrapped():
        env._goal_pose_used_by_rewards = env._object_goal_poses_buf.clone()
        inner()
    env._TrifingerEnv__update_goal_movement_post = wrapped


def reward_terms_of(env):
    """Re-evaluates the six terms on the histories the last `_post_step` used -> [6, N]."""
    t = env._reward_terms
    dt = env.config["sim"]["dt"]
    T = env.env_steps_count
    ft0, ft1 = env._fingertips_frames_state_history[0], env._fingertips_frames_state_history[1]
    ob0, ob1 = env._object_state_history[0], env._object_state_history[1]
    # moving goal: the terms were evaluated BEFORE __update_goal_movement_post replaced the goal pose (ref :559)
    goal = getattr(env, "_goal_pose_used_by_rewards", env._object_goal_poses_buf)
    out = [
        t["finger_reach_object_rate"].compute(T, ft0, ft1, ob0, ob1),
        t["finger_move_penalty"].compute(dt, ft0, ft1),
        t["object_dist"].compute(dt, T, ob0, goal),
        t["object_rot"].compute(dt, T, ob0, goal),
        t["object_rot_delta"].compute(dt, T, ob0, ob1, goal),
        t["object_move"].compute(ob0, ob1, goal),
    ]
    return torch.stack(out, dim=0)
