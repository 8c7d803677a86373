"""Host-side logic that needs no GPU: config handling, parameter flattening, error behaviour."""
import copy

import numpy as np
import pytest
import torch

from leibnizgym_b200 import _native as nat
from leibnizgym_b200.config import (default_trifinger_config, difficulty_config, merge_config, resolve_config)
from leibnizgym_b200.params import build_params, observation_scale, state_scale


def test_merge_config_semantics_of_update_dict():
    base = {"a": 1, "n": {"x": 1, "y": 2}}
    out = merge_config(base, {"n": {"y": 3, "z": 4}, "b": 5})
    assert out is base and base == {"a": 1, "b": 5, "n": {"x": 1, "y": 3, "z": 4}}


def test_defaults_are_not_shared_between_envs():
    a, b = resolve_config({"command_mode": "torque"}), resolve_config({"command_mode": "position"})
    a["reward_terms"]["object_dist"]["weight"] = -1
    assert b["reward_terms"]["object_dist"]["weight"] == 2000            # the reference leaks this
    assert default_trifinger_config()["reward_terms"]["object_dist"]["weight"] == 2000


def test_asymmetric_forces_ft_sensors():
    assert resolve_config(difficulty_config(1, 8, asymmetric_obs=True))["enable_ft_sensors"] is True
    assert resolve_config(difficulty_config(1, 8, asymmetric_obs=False))["enable_ft_sensors"] is False


def test_scale_tables_match_the_oracle():
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    from leibnizgym_b200.synthetic import make_sequence
    for mode in ("torque", "position", "position_impedance"):
        for norm_a in (True, False):
            cfg = resolve_config(difficulty_config(4, 4, command_mode=mode, normalize_action=norm_a))
            ora = OracleEnv(cfg, OracleSim(make_sequence(0, 1, 4), 4))
            lo, hi = state_scale(cfg)
            assert np.array_equal(lo, ora.st_lo.numpy()) and np.array_equal(hi, ora.st_hi.numpy())
            lo, hi = observation_scale(cfg)
            assert np.array_equal(lo, ora.obs_lo.numpy()) and np.array_equal(hi, ora.obs_hi.numpy())
            p = build_params(cfg, 4)
            n = len(ora.st_lo)
            centre = ((ora.st_lo + ora.st_hi) * 0.5).numpy()
            assert np.array_equal(np.array(p.scale_centre[:n], np.float32), centre)
            assert np.array_equal(np.array(p.scale_span[:n], np.float32), (ora.st_hi - ora.st_lo).numpy())


def test_params_flattening():
    cfg = resolve_config(difficulty_config(4, 64, seed=9))
    p = build_params(cfg, 32, env_offset=32, global_num_envs=64)
    assert (p.num_envs, p.env_offset, p.global_num_envs, p.episode_length) == (32, 32, 64, 750)
    assert p.task_difficulty == 4 and p.asymmetric_obs == 1 and p.command_mode == nat.CMD_MODES["torque"]
    assert p.terms[3].activate == 1 and p.terms[3].scale == 3.0 and p.terms[3].sched_start == 1e7
    assert p.terms[0].sched_end == 1e7 and p.terms[4].activate == 0
    assert p.term_active_mask == 0b001111
    assert p.position_tolerance == 0.02 and p.orientation_tolerance == 0.25 and p.success_activate == 0
    assert abs(p.max_com_distance - (0.195 - 0.065 * np.sqrt(3) / 2)) < 1e-15
    cfg["episode_length"] = None
    assert build_params(cfg, 8).episode_length == -1


@pytest.mark.parametrize("patch, exc", [
    ({"command_mode": "velocity"}, ValueError),
    ({"task_difficulty": 7}, ValueError),
    ({"task_difficulty": 0}, ValueError),
    ({"reset_distribution": {"robot_initial_state": {"type": "gaussian"}}}, ValueError),
    ({"reset_distribution": {"object_initial_state": {"type": "grid"}}}, ValueError),
])
def test_invalid_configs_raise_like_the_reference(patch, exc):
    cfg = resolve_config(difficulty_config(1, 8, **copy.deepcopy(patch)))
    with pytest.raises(exc):
        build_params(cfg, 8)


def test_missing_reward_term_raises_keyerror():
    cfg = resolve_config(difficulty_config(1, 8))
    del cfg["reward_terms"]["object_move"]
    with pytest.raises(KeyError):
        build_params(cfg, 8)


def test_no_cpu_fallback():
    from leibnizgym_b200.env import TrifingerEnv
    with pytest.raises(RuntimeError, match="CUDA-only"):
        TrifingerEnv(difficulty_config(1, 8), device="cpu", verbose=False)
    with pytest.raises(KeyError):
        TrifingerEnv({"num_instances": 8}, device="cuda:0")   # command_mode is mandatory (trifinger_env.py:277)
    from leibnizgym_b200.utils import torch_utils as tu
    with pytest.raises(RuntimeError, match="CUDA"):
        tu.quat_mul(torch.zeros(2, 4), torch.zeros(2, 4))


def test_product_code_never_imports_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "leibnizgym_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_rl_games_adapter_shapes_without_a_gpu():
    """Adapter contract on a stub env (ref utils/rlg_train.py:89-154): dict vs tensor, persistent dict, [[], info]."""
    import torch
    from leibnizgym_b200.wrappers.rlg_adapter import RlGamesGpuEnvAdapter

    class Stub:
        def __init__(self, ns):
            self.num_states, self.num_envs, self.calls = ns, 4, []
            self.action_space, self.observation_space, self.state_space = "A", "O", "S"

        def get_number_of_agents(self):
            return 1

        def reset(self):
            self.calls.append("reset")
            return torch.zeros(4, 3)

        def get_state(self):
            return torch.ones(4, self.num_states)

        def step(self, a):
            return torch.full((4, 3), 2.0), torch.zeros(4), torch.zeros(4, dtype=torch.bool), {"k": 1.0}

    sym = RlGamesGpuEnvAdapter(env=Stub(0))
    assert sym.env.calls == ["reset"] and not sym.use_global_obs and "state_space" not in sym.get_env_info()
    out, r, d, extra = sym.step(torch.zeros(4, 9))
    assert torch.is_tensor(out) and extra == [[], {"k": 1.0}] and torch.is_tensor(sym.reset())
    asym = RlGamesGpuEnvAdapter(env=Stub(5))
    out, r, d, extra = asym.step(torch.zeros(4, 9))
    assert out is asym.full_state and set(out) == {"obs", "states"} and out["states"].shape == (4, 5)
    assert asym.reset() is asym.full_state and asym.get_env_info()["state_space"] == "S"


def test_every_bench_workload_resolves_to_kernel_parameters():
    """bench.py's workloads (SURVEY.md 8d C2-C5 and their variants) produce valid env configs and parameter blocks,
    with and without the extension features."""
    import importlib.util
    import os
    from leibnizgym_b200.config import resolve_config
    from leibnizgym_b200.params import build_params
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert {"c2", "c3", "c4", "c5"} <= set(bench.WORKLOADS)
    for name, wl in bench.WORKLOADS.items():
        for ext in (True, False):
            cfg = resolve_config(bench.workload_config(wl, 256, extensions=ext))
            P = build_params(cfg, 256)
            assert P.num_envs == 256 and P.asymmetric_obs == int(wl["asym"]), name
            assert bool(P.dr_activate) == bool(ext and wl.get("dr")), name
            assert bool((P.term_active_mask >> 6) & 1) == bool(ext and wl.get("keypoint")), name
