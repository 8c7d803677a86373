"""Scenario table shared by the fixture generator (make_golden.py) and the tests.

Each scenario is one short episode of the reference env over synthetic state:
`reset()` followed by T-1 `step()` calls.  Sizes are kept small so the
fixtures stay small; the full-size configs of BASELINE.json are exercised by
the GPU tests through the oracle and through size-independent properties.
"""
from __future__ import annotations

import copy

from leibnizgym_b200.config import difficulty_config


def _cfg(d, n, asym, seed, **kw):
    return difficulty_config(d, n, asymmetric_obs=asym, seed=seed, **kw)


_ALL_TERMS = {  # the reference's module defaults (trifinger_env.py:76-113), spelled out
    "finger_reach_object_rate": {"activate": True, "weight": -750, "norm_p": 2},
    "finger_move_penalty": {"activate": True, "weight": -0.1},
    "object_dist": {"activate": True, "weight": 2000},
    "object_rot": {"activate": True, "weight": 300},
    "object_rot_delta": {"activate": True, "weight": -250},
    "object_move": {"activate": True, "weight": -750},
}
_SUCCESS_ON = {"success": {"activate": True, "bonus": 5000.0,
                           "position_tolerance": 0.01, "orientation_tolerance": 0.2}}

SCENARIOS = {
    # BASELINE.json config 1 (the reference's own CPU-runnable case), both obs modes
    "d1_sym": dict(N=24, T=6, seed=1001, config=_cfg(1, 24, False, 1001)),
    "d1_asym": dict(N=24, T=6, seed=1001, config=_cfg(1, 24, True, 1001)),
    # config 2
    "d2_asym": dict(N=24, T=6, seed=1002, config=_cfg(2, 24, True, 1002)),
    # config 3: goal resampling forced on a fraction of envs, random robot reset
    "d3_goal_resample": dict(
        N=28, T=7, seed=1003, goal_reset_p=0.2, reset_p=0.1,
        config=_cfg(3, 28, True, 1003, reset_distribution={
            "robot_initial_state": {"type": "random", "dof_pos_stddev": 0.4, "dof_vel_stddev": 0.2}})),
    # config 4: Hydra difficulty-4 constants; schedule gates straddled by env_steps_count
    "d4_asym": dict(N=24, T=6, seed=1004, config=_cfg(4, 24, True, 1004)),
    # config 5: reset-heavy stress, ~30 % of envs reset every step
    "d4_reset30": dict(N=64, T=8, seed=1005, reset_p=0.3, config=_cfg(4, 64, True, 1005)),
    # module defaults: all six terms, success bonus/goal reset ON, position control,
    # short episodes so the timeout path fires naturally
    "defaults_all_terms": dict(
        N=32, T=9, seed=1006, plant_success=True,
        config=dict(num_instances=32, seed=1006, command_mode="position", episode_length=3,
                    task_difficulty=4, asymmetric_obs=True, reward_terms=copy.deepcopy(_ALL_TERMS),
                    termination_conditions=copy.deepcopy(_SUCCESS_ON))),
    # edge cases of SURVEY.md §A.7 planted in envs 0..6; all terms on, gates with boundaries
    "edge_cases": dict(
        N=16, T=5, seed=1007, plant_edges=True,
        config=dict(num_instances=16, seed=1007, command_mode="torque", episode_length=None,
                    task_difficulty=5, asymmetric_obs=True,
                    reward_terms={
                        "finger_reach_object_rate": {"activate": True, "weight": -750, "norm_p": 2,
                                                     "thresh_sched_start": 0, "thresh_sched_end": 32},
                        "finger_move_penalty": {"activate": True, "weight": -0.1},
                        "object_dist": {"activate": True, "weight": 2000,
                                        "thresh_sched_start": 32, "thresh_sched_end": 48},
                        "object_rot": {"activate": True, "weight": 300, "scale": 3.0},
                        "object_rot_delta": {"activate": True, "weight": -250,
                                             "linear_schedule_start": 16, "linear_schedule_end": 80},
                        "object_move": {"activate": True, "weight": -750},
                    },
                    termination_conditions=copy.deepcopy(_SUCCESS_ON))),
    # remaining goal samplers: difficulty -1 (yaw goal) and 6 (orientation-only goal),
    # default object pose, un-normalised obs, no safety damping, "none" robot reset
    "dm1_raw_obs": dict(
        N=20, T=5, seed=1008, reset_p=0.25,
        config=_cfg(-1, 20, True, 1008, normalize_obs=False, apply_safety_damping=False,
                    reset_distribution={"object_initial_state": {"type": "default"},
                                        "robot_initial_state": {"type": "none"}})),
    # moving goal (trifinger_env.py:1248-1253, :1267-1284): angular velocity sampled at goal resets, re-imposed on the
    # goal body every step, goal pose read back from the simulator's goal rows (planted: random poses per step)
    "moving_goal": dict(
        N=24, T=7, seed=1010, reset_p=0.1, goal_reset_p=0.2, plant_goal_rows=True,
        config=_cfg(4, 24, True, 1010, goal_movement={"rotation": {"activate": True, "rate_magnitude": 0.5}})),
    # variable-impedance command mode: 18-dim action (position + stiffness), 50-dim obs, 122-dim states
    "d3_impedance": dict(N=20, T=5, seed=1011, reset_p=0.2, action_dim=18,
                         config=_cfg(3, 20, True, 1011, command_mode="position_impedance")),
    "d6_sym": dict(N=20, T=5, seed=1009, reset_p=0.25, goal_reset_p=0.25,
                   config=_cfg(6, 20, False, 1009, normalize_action=False,
                               reset_distribution={"object_initial_state": {"type": "none"}})),
}

# Extra scenarios for the fuzzer (tests/golden/fuzz_reference.py): a JSON file of {name: scenario} named by the
# environment, so that the generator's child processes see them too.  Never set by the test-suite.
import json as _json   # noqa: E402
import os as _os       # noqa: E402
_extra = _os.environ.get("LEIBNIZ_EXTRA_SCENARIOS")
if _extra and _os.path.exists(_extra):
    with open(_extra) as _f:
        SCENARIOS.update(_json.load(_f))

