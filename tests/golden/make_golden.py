#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference.

    python tests/golden/make_golden.py            # all scenarios
    python tests/golden/make_golden.py d4_asym    # one scenario

Every scenario is run twice, each in its own process (the reference keeps
class-level mutable state): once with TorchScript on (the
reference's real execution path — these outputs are what is stored) and once
with PYTORCH_JIT=0 and torch.rand/randn wrapped, which yields the random draws
the samplers consumed.  Both runs must agree bit for bit, otherwise the
generator aborts.  The stored draws let the CUDA path and the oracle replay the
reference's resets exactly ("injected draws", SURVEY.md §7 RNG).

Only this script needs /root/reference; the fixtures travel with the repo.
"""
from __future__ import annotations

import json
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from leibnizgym_b200.synthetic import (bernoulli_masks, make_sequence, plant_edge_cases,  # noqa: E402
                                       plant_goal_rows, FINGERTIP_BODIES)
from scenarios import SCENARIOS  # noqa: E402

U_COLS, N_COLS = 24, 8


def canonical_draws(log, cfg, k, goal_only):
    """Reorders the recorded rand/randn calls of one `_reset_impl` (or
    `_goal_reset_impl`) into the injected-draw layout of include/leibniz_b200.h:
    uniforms [k,24] = robot noise 0:18 | object r,theta,yaw 18:21 | goal u0,u1,u2 21:24;
    normals  [k,8]  = goal quaternion 0:4 | goal angular-velocity axis 4:7 | magnitude 7.
    Draw order: ref trifinger_env.py:394-411 / SURVEY.md §A.5."""
    u = np.full((k, U_COLS), np.nan, np.float32)
    n = np.full((k, N_COLS), np.nan, np.float32)
    it = iter(log)

    def nxt(kind, shape):
        got_kind, val = next(it)
        assert got_kind == kind and tuple(val.shape) == shape, (got_kind, tuple(val.shape), kind, shape)
        return val.numpy()

    rd = cfg["reset_distribution"]
    if not goal_only:
        if rd["robot_initial_state"]["type"] == "random":
            u[:, 0:18] = nxt("u", (k, 18))
        if rd["object_initial_state"]["type"] == "random":
            for c in (18, 19, 20):
                u[:, c] = nxt("u", (k,))
    d = cfg["task_difficulty"]
    if d == -1:
        for c in (21, 22, 23):
            u[:, c] = nxt("u", (k,))
    elif d == 1:
        for c in (21, 22):
            u[:, c] = nxt("u", (k,))
    elif d in (3, 4, 5):
        for c in (21, 22, 23):
            u[:, c] = nxt("u", (k,))
    if d in (4, 5, 6):
        n[:, 0:4] = nxt("n", (k, 4))
    if cfg["goal_movement"]["rotation"]["activate"]:
        n[:, 4:7] = nxt("n", (k, 3))
        n[:, 7:8] = nxt("n", (k, 1))
    rest = list(it)
    assert not rest, f"unconsumed draws: {[(a, tuple(b.shape)) for a, b in rest]}"
    return u, n


def run_scenario(name: str, record_draws: bool) -> dict:
    from ref_harness import DrawRecorder, build_reference_env, reward_terms_of, stash_goal_before_movement
    sc = SCENARIOS[name]
    N, T, seed = sc["N"], sc["T"], sc["seed"]
    seq = make_sequence(seed, T, N, action_dim=sc.get("action_dim", 9))
    reset_masks = bernoulli_masks(seed, T, N, sc.get("reset_p", 0.0))
    goal_masks = bernoulli_masks(seed + 1, T, N, sc.get("goal_reset_p", 0.0))
    if sc.get("plant_goal_rows"):   # what a simulator integrating the goal body would leave in its root rows
        plant_goal_rows(seq, seed)
    env, fake = build_reference_env(sc["config"], seq)
    cfg = env.config
    if cfg["goal_movement"]["rotation"]["activate"]:
        stash_goal_before_movement(env)
    out = {"meta": dict(name=name, N=N, T=T, seed=seed, torch=torch.__version__,
                        config=json.loads(json.dumps(sc["config"])))}
    steps = []

    rec = DrawRecorder()
    with rec:
        for t in range(T):
            s = {}
            if t == 0:
                ids_r = torch.arange(N)
                ids_g = torch.zeros(0, dtype=torch.long)
                fake.root_indexed, fake.dof_indexed = [], None
                env.reset()
                log = rec.take()
                if record_draws:
                    s["reset_u"], s["reset_n"] = canonical_draws(log, cfg, N, goal_only=False)
                goal = env._object_goal_poses_buf.clone()
                if sc.get("plant_edges"):
                    plant_edge_cases(seq, goal)
                if sc.get("plant_success"):
                    root = seq.root_state.view(T, N, 4, 13)
                    for e in (1, 5, 9, 13):
                        root[2:, e, 2, 0:7] = goal[e]
                        root[2:, e, 2, 0] += 0.004  # inside the 0.01 m tolerance
            else:
                if reset_masks is not None:
                    env._reset_buf |= reset_masks[t]
                if goal_masks is not None:
                    env._goal_reset_buf |= goal_masks[t]
                s["reset_in"] = env._reset_buf.numpy().copy()
                s["goal_reset_in"] = env._goal_reset_buf.numpy().copy()
                ids_r = torch.nonzero(env._reset_buf).view(-1)
                ids_g = torch.nonzero(env._goal_reset_buf).view(-1)
                fake.root_indexed, fake.dof_indexed = [], None
                env.step(seq.action[t].clone())
                log = rec.take()
                if record_draws:
                    # reset draws come first, goal-reset draws second (env_base.py:374-379)
                    n_goal_calls = _num_goal_calls(cfg)
                    n_reset_calls = _num_reset_calls(cfg) if len(ids_r) else 0
                    s["reset_u"], s["reset_n"] = canonical_draws(
                        log[:n_reset_calls], cfg, len(ids_r), goal_only=False) if len(ids_r) else _empty()
                    s["goal_u"], s["goal_n"] = canonical_draws(
                        log[n_reset_calls:], cfg, len(ids_g), goal_only=True) if len(ids_g) else _empty()
                    assert len(log) == n_reset_calls + (n_goal_calls if len(ids_g) else 0)
                s["terms"] = reward_terms_of(env).numpy()
                s["reward"] = env._reward_buf.numpy().copy()
                s["info"] = {k: float(v) for k, v in env._step_info.items()}
            s["reset_ids"] = ids_r.numpy()
            s["goal_reset_ids"] = ids_g.numpy()
            # simulator-facing side effects of the reset (captured just before `simulate`)
            s["pre_sim_dof"] = fake.pre_sim_dof.numpy().reshape(N, 9, 2).copy()
            s["pre_sim_obj_root"] = fake.pre_sim_root.view(N, 4, 13)[:, 2].numpy().copy()
            s["pre_sim_goal_root"] = fake.pre_sim_root.view(N, 4, 13)[:, 3].numpy().copy()
            s["dof_index_list"] = None if fake.dof_indexed is None else fake.dof_indexed.numpy()
            lists = [x.numpy() for x in fake.root_indexed]
            if cfg["goal_movement"]["rotation"]["activate"]:   # the last call of the step is __update_goal_movement_pre's
                s["root_index_list_move"], lists = lists[-1], lists[:-1]
            s["root_index_lists"] = lists
            s["applied_torque"] = fake.applied_torque.numpy().copy()
            s["goal_pose"] = env._object_goal_poses_buf.numpy().copy()
            s["goal_movement"] = env._object_goal_movement_buf.numpy().copy()
            s["action_buf"] = env._action_buf.numpy().copy()
            s["obs"] = env._obs_buf.numpy().copy()
            s["states"] = env._states_buf.numpy().copy()
            s["reset_buf"] = env._reset_buf.numpy().copy()
            s["goal_reset_buf"] = env._goal_reset_buf.numpy().copy()
            s["steps_count"] = env._steps_count_buf.numpy().copy()
            s["successes"] = env._successes.numpy().copy()
            s["sched_step"] = int(env.env_steps_count)
            steps.append(s)
    out["steps"] = steps
    # the inputs actually played (after planting), compact: only what the path reads
    out["inputs"] = dict(
        dof_state=seq.dof_state.numpy(),
        object_root=seq.root_state.view(T, N, 4, 13)[:, :, 2].numpy().copy(),
        fingertips=seq.rigid_body[:, :, list(FINGERTIP_BODIES)].numpy().copy(),
        dof_force=seq.dof_force.numpy(),
        ft_sensors=seq.ft_sensors.numpy(),
        action=seq.action.numpy(),
        reset_masks=None if reset_masks is None else reset_masks.numpy(),
        goal_masks=None if goal_masks is None else goal_masks.numpy(),
        goal_root=seq.root_state.view(T, N, 4, 13)[:, :, 3].numpy().copy() if sc.get("plant_goal_rows") else None,
    )
    return out


def _empty():
    return np.zeros((0, U_COLS), np.float32), np.zeros((0, N_COLS), np.float32)


def _num_goal_calls(cfg):
    d = cfg["task_difficulty"]
    n = {-1: 3, 1: 2, 2: 0, 3: 3, 4: 4, 5: 4, 6: 1}[d]
    if cfg["goal_movement"]["rotation"]["activate"]:
        n += 2
    return n


def _num_reset_calls(cfg):
    rd = cfg["reset_distribution"]
    n = _num_goal_calls(cfg)
    if rd["robot_initial_state"]["type"] == "random":
        n += 1
    if rd["object_initial_state"]["type"] == "random":
        n += 3
    return n


def _child(name, jit_on):
    env = dict(os.environ)
    env["PYTORCH_JIT"] = "1" if jit_on else "0"
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        path = f.name
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, path,
                    "0" if jit_on else "1"], check=True, env=env)
    with open(path, "rb") as f:
        data = pickle.load(f)
    os.unlink(path)
    return data


def _flatten(out):
    flat = {"meta": np.array(json.dumps(out["meta"]))}
    for k, v in out["inputs"].items():
        if v is not None:
            flat[f"in/{k}"] = v
    for t, s in enumerate(out["steps"]):
        for k, v in s.items():
            if v is None:
                continue
            if k == "info":
                flat[f"s{t}/info"] = np.array(json.dumps(v))
            elif k == "root_index_lists":
                for i, a in enumerate(v):
                    flat[f"s{t}/root_index_list{i}"] = a
            else:
                flat[f"s{t}/{k}"] = np.asarray(v)
    return flat


def generate(name):
    jit = _child(name, True)
    eager = _child(name, False)
    fj, fe = _flatten(jit), _flatten(eager)
    for k, v in fj.items():
        if k == "meta":
            continue
        a, b = np.asarray(v), np.asarray(fe[k])
        same = (a == b) | ((a != a) & (b != b)) if a.dtype.kind == "f" else (a == b)
        if not np.all(same):
            raise SystemExit(f"{name}: scripted and eager runs disagree on {k}")
    for k, v in fe.items():  # draws exist only in the eager run
        fj.setdefault(k, v)
    path = os.path.join(os.environ.get("LEIBNIZ_GOLDEN_OUT", HERE), f"{name}.npz")
    np.savez_compressed(path, **fj)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(fj)} arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        _, _, nm, out_path, rec = sys.argv
        torch.set_num_threads(1)
        result = run_scenario(nm, record_draws=(rec == "1"))
        with open(out_path, "wb") as f:
            pickle.dump(result, f)
    else:
        for nm in (sys.argv[1:] or list(SCENARIOS)):
            generate(nm)
