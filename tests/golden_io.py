"""Loading and replaying the fixtures of tests/golden/*.npz.

`replay(env, golden, check)` drives any env that offers the reference's surface
(`reset`, `step`, `_reset_buf`, `_goal_reset_buf`, `inject_draws`) through the
scenario the reference ran, injecting the same reset masks and random draws,
and hands every stored array to `check(step, key, expected, env)`.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from leibnizgym_b200.synthetic import FINGERTIP_BODIES, StateSequence

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name: str):
        self.name = name
        self.z = np.load(name if os.path.isabs(name) else os.path.join(GOLDEN_DIR, f"{name}.npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.N, self.T = self.meta["N"], self.meta["T"]
        self.config = self.meta["config"]

    def has(self, t, key):
        return f"s{t}/{key}" in self.z.files

    def get(self, t, key):
        return self.z[f"s{t}/{key}"]

    def info(self, t):
        return json.loads(str(self.z[f"s{t}/info"]))

    def masks(self, kind):
        k = f"in/{kind}"
        return torch.from_numpy(self.z[k]) if k in self.z.files else None

    def sequence(self) -> StateSequence:
        """Rebuilds full-layout simulator tensors from the compact stored inputs
        (bodies/actors the path never reads are zero)."""
        z, T, N = self.z, self.T, self.N
        root = torch.zeros(T, N, 4, 13)
        root[..., 6] = 1.0
        root[:, :, 2] = torch.from_numpy(z["in/object_root"])
        if "in/goal_root" in z.files:   # moving-goal scenarios: the goal body's rows as the simulator left them
            root[:, :, 3] = torch.from_numpy(z["in/goal_root"])
        rb = torch.zeros(T, N, 20, 13)
        rb[:, :, list(FINGERTIP_BODIES)] = torch.from_numpy(z["in/fingertips"])
        return StateSequence(
            dof_state=torch.from_numpy(z["in/dof_state"]).clone(),
            root_state=root.reshape(T, 4 * N, 13).contiguous(),
            rigid_body=rb,
            dof_force=torch.from_numpy(z["in/dof_force"]).clone(),
            ft_sensors=torch.from_numpy(z["in/ft_sensors"]).clone(),
            action=torch.from_numpy(z["in/action"]).clone(),
        )


def _draws(g: Golden, t: int, kind: str):
    if not g.has(t, f"{kind}_u"):
        return None
    return g.get(t, f"{kind}_u"), g.get(t, f"{kind}_n")


def replay(env, g: Golden, check, inject: bool = True):
    """Runs the scenario; `check(t, key, expected_numpy)` is called for every stored output."""
    seq_action = torch.from_numpy(g.z["in/action"])
    rmask, gmask = g.masks("reset_masks"), g.masks("goal_masks")
    for t in range(g.T):
        if inject:
            env.inject_draws(reset=_draws(g, t, "reset"), goal=_draws(g, t, "goal"))
        if t == 0:
            env.reset()
        else:
            if rmask is not None:
                env._reset_buf |= rmask[t].to(env._reset_buf.device)
            if gmask is not None:
                env._goal_reset_buf |= gmask[t].to(env._goal_reset_buf.device)
            check(t, "reset_in", g.get(t, "reset_in"))
            check(t, "goal_reset_in", g.get(t, "goal_reset_in"))
            env.step(seq_action[t].clone())
        for key in ("reset_ids", "goal_reset_ids", "pre_sim_dof", "pre_sim_obj_root", "pre_sim_goal_root",
                    "dof_index_list", "root_index_list0", "root_index_list1", "root_index_list_move", "applied_torque",
                    "goal_pose", "goal_movement", "action_buf", "obs", "states", "reset_buf",
                    "goal_reset_buf", "steps_count", "successes", "terms", "reward", "sched_step"):
            if g.has(t, key):
                check(t, key, g.get(t, key))
        if g.has(t, "info"):
            check(t, "info", g.info(t))
