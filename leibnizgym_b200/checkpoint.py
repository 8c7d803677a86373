"""Env-state checkpoint: everything the MDP carries from one step to the next, on disk.

The reference has no resume path for the environment (`dump_config`, ref
envs/env_base.py:295-309, only writes the config).  A run of this build can stop after any step
and continue bit-identically: the file holds the goal poses and movement, step counters, reset /
goal-reset / success flags, the one-deep history, the last action / outputs, the per-step
statistics, and the device control block (RNG epoch of the kernel's Philox stream, frame counter,
compaction epoch) — plus, for `SyntheticSim`, the simulator tensors and its clock, because the
pre-physics pass of the next step reads (and the resets wrote) simulator memory.

Format: one uncompressed `.npz`; `meta` is a JSON string (format version, shard geometry, config),
`env/<name>` and `sim/<name>` are the arrays.  Shards checkpoint independently (one file per rank).
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
import torch

FORMAT_VERSION = 1

# name -> attribute of TrifingerEnv (device tensors, restored in place: the pointers the C ABI holds stay valid)
_ENV_TENSORS = {
    "obs": "_obs_buf", "states": "_states_buf", "obs_clipped": "_obs_clipped", "states_clipped": "_states_clipped",
    "action": "_action_buf", "reward": "_reward_buf", "reset": "_reset_buf", "goal_reset": "_goal_reset_buf",
    "successes": "_successes", "dones": "_dones", "steps_count": "_steps_count_buf",
    "goal_pose": "_object_goal_poses_buf", "goal_movement": "_object_goal_movement_buf", "history": "_history",
    "applied_torque": "_applied_torque", "step_stats": "_step_stats", "control": "_control",
    "scan_status": "_scan_status", "counts": "_counts", "reset_ids": "_reset_ids", "goal_reset_ids": "_goal_reset_ids",
    "reward_coef": "_reward_coef",
}
_SIM_TENSORS = ("dof_state", "root_state", "rigid_body", "dof_force", "ft_sensors")


def env_state_dict(env) -> Dict[str, object]:
    """Snapshot as host numpy arrays (synchronises the device)."""
    out: Dict[str, object] = {}
    for name, attr in _ENV_TENSORS.items():
        t = getattr(env, attr, None)
        if t is not None and t.numel():
            out[f"env/{name}"] = t.detach().cpu().numpy()
    sim = env._sim
    sim_meta = {"frame_count": int(sim.get_frame_count()), "cursor": int(getattr(sim, "cursor", 0)),
                "kind": type(sim).__name__}
    if not getattr(sim, "rebinds_tensors", False):   # simulators that own device tensors we can read back
        for name in _SIM_TENSORS:
            t = getattr(sim, name, None)
            if t is not None:
                out[f"sim/{name}"] = t.detach().cpu().numpy()
    out["meta"] = {
        "format": FORMAT_VERSION, "num_instances": int(env.num_instances), "global_num_instances": int(env._global_N),
        "env_offset": int(env._env_offset), "obs_dim": int(env.get_obs_dim()), "state_dim": int(env.get_state_dim()),
        "action_dim": int(env.get_action_dim()), "sim": sim_meta, "config": json.loads(json.dumps(env.config, default=str)),
    }
    return out


def load_env_state_dict(env, state: Dict[str, object]) -> None:
    meta = state["meta"]
    if isinstance(meta, (str, bytes, np.ndarray)):
        meta = json.loads(str(meta))
    if meta.get("format") != FORMAT_VERSION:
        raise ValueError(f"checkpoint format {meta.get('format')} != {FORMAT_VERSION}")
    for key, have in (("num_instances", env.num_instances), ("global_num_instances", env._global_N),
                      ("env_offset", env._env_offset), ("obs_dim", env.get_obs_dim()),
                      ("state_dim", env.get_state_dim()), ("action_dim", env.get_action_dim())):
        if int(meta[key]) != int(have):
            raise ValueError(f"checkpoint was written for {key}={meta[key]}, this env has {have}")
    for name, attr in _ENV_TENSORS.items():
        t = getattr(env, attr, None)
        arr = state.get(f"env/{name}")
        if t is None or not t.numel():
            continue
        if arr is None:
            if name in ("obs_clipped", "states_clipped"):
                continue          # the writer had no wrapper attached; the next step refills them
            raise ValueError(f"checkpoint lacks env/{name}")
        src = torch.from_numpy(np.ascontiguousarray(arr))
        if tuple(src.shape) != tuple(t.shape) or src.dtype != t.dtype:
            raise ValueError(f"env/{name}: checkpoint {tuple(src.shape)} {src.dtype} != env {tuple(t.shape)} {t.dtype}")
        t.copy_(src)
    sim = env._sim
    if any(k.startswith("sim/") for k in state):
        for name in _SIM_TENSORS:
            arr, t = state.get(f"sim/{name}"), getattr(sim, name, None)
            if arr is not None and t is not None:
                t.copy_(torch.from_numpy(np.ascontiguousarray(arr)).view(t.shape))
    if hasattr(sim, "frame_count"):
        sim.frame_count = int(meta["sim"]["frame_count"])
    if hasattr(sim, "cursor"):
        sim.cursor = int(meta["sim"]["cursor"])
    env._clear_injection()
    env._step_info = env._make_info()
    torch.cuda.synchronize(env._torch_device)


def save_env_state(env, path: str) -> str:
    if not path.endswith(".npz"):
        path += ".npz"
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    state = env_state_dict(env)
    state["meta"] = np.array(json.dumps(state["meta"]))
    tmp = path + ".tmp.npz"
    np.savez(tmp, **state)
    os.replace(tmp, path)      # a crash mid-write never leaves a truncated checkpoint under the final name
    return path


def load_env_state(env, path: str) -> None:
    if not path.endswith(".npz"):
        path += ".npz"
    with np.load(path) as z:
        load_env_state_dict(env, {k: z[k] for k in z.files})
