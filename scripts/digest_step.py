#!/usr/bin/env python
"""Bit-level fingerprint of the CUDA path on a fixed scenario (development tool, needs a GPU).

    LG_LIB_PATH=ab/new.so python scripts/digest_step.py        # prints one sha256 per buffer group

Two builds of the library whose arithmetic is meant to be unchanged (a data-path or scheduling refactor) must print
identical digests: 20 000 envs (ragged last tile), difficulty 4, asymmetric, 30 % forced resets + 5 % goal resets drawn
from the kernels' own Philox stream, 24 fused steps replayed from a CUDA graph plus 3 eager steps through env.step().
"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leibnizgym_b200.config import difficulty_config  # noqa: E402
from leibnizgym_b200.env import TrifingerEnv  # noqa: E402
from leibnizgym_b200.graph_runner import GraphRunner  # noqa: E402
from leibnizgym_b200.sim import SyntheticSim  # noqa: E402
from leibnizgym_b200.synthetic import bernoulli_masks, make_sequence  # noqa: E402


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()[:16]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    dev, R = "cuda:0", 8
    for asym, diff in ((True, 4), (False, 1)):
        cfg = difficulty_config(diff, N, asymmetric_obs=asym, seed=5)
        cfg["termination_conditions"] = {"success": {"activate": True, "bonus": 5000.0, "position_tolerance": 0.05,
                                                     "orientation_tolerance": 1.0}}
        ring = make_sequence(5, R, N, device=dev)
        masks = bernoulli_masks(5, R, N, 0.3, device=dev)
        gmasks = bernoulli_masks(12, R, N, 0.05, device=dev)
        env = TrifingerEnv(cfg, device=dev, verbose=False, sim=SyntheticSim(ring, dev))
        env.enable_term_rewards(True)
        env.reset()
        runner = GraphRunner(env, ring, rotate_outputs=True, inject_reset_masks=masks, inject_goal_masks=gmasks)
        runner.capture(R)
        for _ in range(3):
            runner.graph.replay()
        torch.cuda.synchronize()
        print(f"asym={asym} graph : obs/states {digest(runner.obs_slots, *( [runner.state_slots] if asym else []))}"
              f" reward {digest(env._reward_buf)} flags {digest(env._reset_buf, env._goal_reset_buf, env._successes, env._dones)}"
              f" steps {digest(env._steps_count_buf)} goal {digest(env._object_goal_poses_buf, env._history)}"
              f" ids {digest(env._reset_ids[: int(env._counts[0])], env._goal_reset_ids[: int(env._counts[1])], env._counts)}"
              f" sim {digest(ring.dof_state, ring.root_state)} torque {digest(env._applied_torque, env._action_buf)}")
        env.set_forced_resets(masks[0], gmasks[0])
        for t in range(3):
            env.step(ring.action[t])
        torch.cuda.synchronize()
        print(f"asym={asym} eager : obs {digest(env._obs_buf)} states {digest(env._states_buf)} reward {digest(env._reward_buf)}"
              f" terms {digest(env._term_rewards)} flags {digest(env._reset_buf, env._goal_reset_buf, env._successes, env._dones)}"
              f" stats {[round(float(x), 9) for x in env._step_stats.cpu()[:13]]}")
        try:   # the wrapper's clamped copies (not part of -DLG_FAST_BUILD libraries)
            from leibnizgym_b200.wrappers import VecTaskPython
            vec = VecTaskPython(env, rl_device=dev, clip_obs=0.7, clip_actions=0.5)
            for t in range(3, 5):
                obs, rew, done, _ = vec.step(2.0 * ring.action[t])
            torch.cuda.synchronize()
            print(f"asym={asym} clip  : obs {digest(obs, env._obs_buf)} states {digest(vec.get_state(), env._states_buf)}"
                  f" reward {digest(rew, done)} action {digest(env._action_buf, env._applied_torque)}")
        except Exception as ex:  # noqa: BLE001
            print(f"asym={asym} clip  : n/a ({type(ex).__name__})")


if __name__ == "__main__":
    main()
