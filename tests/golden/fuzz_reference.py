#!/usr/bin/env python
"""Fuzzes the oracle against the UNMODIFIED reference on random configurations (build container only).

    python tests/golden/fuzz_reference.py [count] [seed]

Each random scenario (difficulty, obs mode, command mode, normalisation flags, reset distributions, every reward term
with random activation / weight / schedule, success termination, moving goal, episode length, forced reset and
goal-reset rates) is run through the real reference exactly like the committed fixtures (make_golden.py: scripted +
eager child processes that must agree bitwise), then the oracle replays it twice — with the recorded draws injected and
drawing from torch's seeded generator — and every stored array must match bit for bit.  Nothing is written into the
repository; failures print the offending scenario as JSON.
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)
sys.path.insert(0, HERE)


def random_scenario(rng: random.Random, idx: int) -> dict:
    d = rng.choice([-1, 1, 2, 3, 4, 5, 6])
    N, T = rng.randint(8, 36), rng.randint(4, 7)
    seed = 5000 + idx
    mode = rng.choice(["torque", "position", "position_impedance"])

    def sched():
        if rng.random() < 0.5:
            return {}
        a = rng.randint(0, 3 * N)
        return {"thresh_sched_start": a, "thresh_sched_end": a + rng.randint(1, 4 * N)}

    terms = {
        "finger_reach_object_rate": {"activate": rng.random() < 0.8, "weight": -rng.randint(1, 900), "norm_p": 2, **sched()},
        "finger_move_penalty": {"activate": rng.random() < 0.8, "weight": -rng.random()},
        "object_dist": {"activate": rng.random() < 0.8, "weight": rng.randint(1, 3000), **sched()},
        "object_rot": {"activate": rng.random() < 0.8, "weight": rng.randint(1, 3000), "epsilon": 0.01,
                       "scale": rng.choice([1.0, 3.0, 0.5 + rng.random() * 4]), **sched()},
        "object_rot_delta": {"activate": rng.random() < 0.8, "weight": -rng.randint(1, 500)},
        "object_move": {"activate": rng.random() < 0.8, "weight": -rng.randint(1, 900)},
    }
    if rng.random() < 0.5:
        a = rng.randint(0, 3 * N)
        terms["object_rot_delta"].update(linear_schedule_start=a, linear_schedule_end=a + rng.randint(1, 6 * N))
    keys = list(terms)
    rng.shuffle(keys)                     # the reference accumulates in the user's dict order
    cfg = {
        "num_instances": N, "seed": seed, "task_difficulty": d, "command_mode": mode,
        "episode_length": rng.choice([None, 2, 3, 5, 750]),
        "asymmetric_obs": rng.random() < 0.6, "enable_ft_sensors": rng.random() < 0.5,
        "normalize_obs": rng.random() < 0.8, "normalize_action": rng.random() < 0.8,
        "apply_safety_damping": rng.random() < 0.7,
        "reset_distribution": {
            "robot_initial_state": {"type": rng.choice(["default", "random", "none"]),
                                    "dof_pos_stddev": round(rng.random() * 0.5, 3), "dof_vel_stddev": round(rng.random() * 0.3, 3)},
            "object_initial_state": {"type": rng.choice(["default", "random", "none"])},
        },
        "goal_movement": {"rotation": {"activate": rng.random() < 0.3, "rate_magnitude": round(0.1 + rng.random(), 3)}},
        "reward_terms": {k: terms[k] for k in keys},
        "termination_conditions": {"success": {"activate": rng.random() < 0.5, "bonus": float(rng.randint(1, 5000)),
                                               "position_tolerance": rng.choice([0.01, 0.02, 0.3]),
                                               "orientation_tolerance": rng.choice([0.1, 0.25, 3.2])}},
        "sim": {"dt": rng.choice([0.02, 0.01, 0.005])},
    }
    sc = dict(N=N, T=T, seed=seed, reset_p=rng.choice([0.0, 0.1, 0.4]), goal_reset_p=rng.choice([0.0, 0.2]), config=cfg)
    if mode == "position_impedance":
        sc["action_dim"] = 18
    if cfg["goal_movement"]["rotation"]["activate"]:
        sc["plant_goal_rows"] = True
    if rng.random() < 0.3:
        sc["plant_success"] = N >= 14
    return sc


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    from adapters import OracleAdapter
    from golden_io import Golden, replay
    from test_oracle_golden import _exact
    failures = 0
    with tempfile.TemporaryDirectory() as tmp:
        scen = {f"fuzz{i:03d}": random_scenario(rng, i) for i in range(count)}
        extra = os.path.join(tmp, "scenarios.json")
        with open(extra, "w") as f:
            json.dump(scen, f)
        env = dict(os.environ, LEIBNIZ_EXTRA_SCENARIOS=extra, LEIBNIZ_GOLDEN_OUT=tmp)
        for name, sc in scen.items():
            res = subprocess.run([sys.executable, os.path.join(HERE, "make_golden.py"), name], env=env,
                                 capture_output=True, text=True)
            if res.returncode != 0:
                tail = (res.stdout + res.stderr).strip().splitlines()[-1:]
                print(f"{name}: reference run failed ({tail}) — scenario skipped: {json.dumps(sc['config'])[:200]}")
                continue
            g = Golden(os.path.join(tmp, f"{name}.npz"))
            for inject in (True, False):
                oracle = OracleAdapter(g.config, g.sequence())
                bad = []

                def check(t, key, expected):
                    got = oracle.observe(key)
                    ok = (all(got[k] == expected[k] for k in expected) and set(got) == set(expected)) if key == "info" \
                        else _exact(got, expected)
                    if not ok:
                        bad.append((t, key))
                try:
                    replay(oracle, g, check, inject=inject)
                except Exception as ex:  # noqa: BLE001
                    bad.append(("exception", repr(ex)))
                if bad:
                    failures += 1
                    print(f"{name} inject={inject}: MISMATCH {bad[:6]}\n  {json.dumps(sc)}")
            print(f"{name}: ok (d={sc['config']['task_difficulty']}, {sc['config']['command_mode']}, "
                  f"asym={sc['config']['asymmetric_obs']}, N={sc['N']}, T={sc['T']})", flush=True)
    print(f"{count} scenarios, {failures} mismatching replays")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
