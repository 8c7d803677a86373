"""Random whole-env configurations (the generator of tests/golden/fuzz_reference.py): CUDA path against the oracle —
ids, flags and counters bit-exact, obs / states / rewards within tests/tolerances.py."""
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from tolerances import compare  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idx", range(12))
def test_random_configuration_matches_oracle(idx):
    from fuzz_reference import random_scenario
    from leibnizgym_b200.config import resolve_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import bernoulli_masks, make_sequence, plant_goal_rows
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    from tolerances import angle_slack
    sc = random_scenario(random.Random(1000 + idx), idx)
    N, T, cfg = sc["N"] * 37, sc["T"], dict(sc["config"])     # a few hundred envs: several tiles, ragged
    cfg["num_instances"] = N
    seq = make_sequence(sc["seed"], T, N, action_dim=sc.get("action_dim", 9))
    if sc.get("plant_goal_rows"):
        plant_goal_rows(seq, sc["seed"])
    rm = bernoulli_masks(sc["seed"], T, N, sc["reset_p"])
    gm = bernoulli_masks(sc["seed"] + 1, T, N, sc["goal_reset_p"])
    env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    env.enable_term_rewards(True)
    ora = OracleEnv(resolve_config(cfg), OracleSim(seq, N))
    g = torch.Generator().manual_seed(idx)
    draws = lambda k: (torch.rand(k, 24, generator=g).numpy(), torch.randn(k, 8, generator=g).numpy())  # noqa: E731
    d = draws(N)
    env.inject_draws(reset=d)
    ora.inject_draws(reset=d)
    env.reset()
    ora.reset()
    asym = bool(cfg["asymmetric_obs"])
    for t in range(1, T):
        if rm is not None:
            ora.reset_buf |= rm[t]
            env._reset_buf |= rm[t].cuda()
        if gm is not None:
            ora.goal_reset_buf |= gm[t]
            env._goal_reset_buf |= gm[t].cuda()
        kr, kg = int(ora.reset_buf.sum()), int(ora.goal_reset_buf.sum())
        dr, dg = (draws(kr) if kr else None), (draws(kg) if kg else None)
        env.inject_draws(reset=dr, goal=dg)
        ora.inject_draws(reset=dr, goal=dg)
        env.step(seq.action[t].cuda())
        ora.step(seq.action[t].clone())
        where = (idx, t, sc["config"]["task_difficulty"], sc["config"]["command_mode"])

        def chk(key, got, exp, extra=None):
            ok, detail = compare(key, got.cpu().numpy() if torch.is_tensor(got) else got,
                                 exp.numpy() if torch.is_tensor(exp) else exp, extra_atol=extra)
            assert ok, (where, key, detail)

        chk("reset_ids", env.reset_env_ids, ora.last_ids[0])
        chk("goal_reset_ids", env.goal_reset_env_ids, ora.last_ids[1])
        chk("obs", env.obs_buf, ora.obs_buf)
        if asym:
            chk("states", env.states_buf, ora.states_buf)
        cur = angle_slack(ora.obj_hist[0][:, 3:7], ora.goal_poses_used if hasattr(ora, "goal_poses_used") else ora.goal_poses[:, 3:7])
        prev = angle_slack(ora.obj_hist[1][:, 3:7], ora.goal_poses[:, 3:7])
        tp = ora.terms
        slack = (abs(tp["object_rot"]["weight"]) * cfg["sim"]["dt"] / tp["object_rot"]["scale"] * cur * tp["object_rot"]["activate"]
                 + abs(tp["object_rot_delta"]["weight"]) * (cur + prev) * tp["object_rot_delta"]["activate"])
        if not cfg["goal_movement"]["rotation"]["activate"]:   # with a moving goal the slack's goal is one step stale
            chk("reward", env.reward_buf, ora.reward_buf, slack)
        chk("reset_buf", env._reset_buf, ora.reset_buf)
        chk("goal_reset_buf", env._goal_reset_buf, ora.goal_reset_buf)
        chk("steps_count", env._steps_count_buf, ora.steps_count_buf)
        chk("successes", env._successes, ora.successes)
        chk("goal_pose", env._object_goal_poses_buf, ora.goal_poses)
        chk("applied_torque", env._applied_torque, ora.applied_torque)
