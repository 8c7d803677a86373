"""GPU parity against the UNMODIFIED reference: every fixture of tests/golden is replayed
through the CUDA TrifingerEnv (C ABI) with the reference's random draws injected."""
import pytest

from golden_io import Golden, golden_names, replay
from tolerances import compare

pytestmark = pytest.mark.gpu


def _run(name, fused):
    import numpy as np
    from adapters import CudaAdapter
    from leibnizgym_b200.config import resolve_config
    from tolerances import angle_slack
    g = Golden(name)
    env = CudaAdapter(g.config, g.sequence(), fused=fused)
    failures, worst = [], {}
    terms_cfg = resolve_config(g.config)["reward_terms"]
    dt = resolve_config(g.config)["sim"]["dt"]
    seen = {}   # expected arrays of the current step the angle allowance is derived from

    def angle_allowance(key):
        """Conditioning-aware slack of the two angle-derived terms (tolerances.angle_slack), from the reference's own
        arrays: current object / goal quaternions are observation columns 21:25 / 28:32 (identity scaling), the
        previous object pose is the object root row the simulator started the step from."""
        if "obs" not in seen or "pre_sim_obj_root" not in seen:
            return None
        obs, prev_row = seen["obs"], seen["pre_sim_obj_root"]
        cur = angle_slack(obs[:, 21:25], obs[:, 28:32])
        prev = angle_slack(prev_row[:, 3:7], obs[:, 28:32])
        rot, delta = terms_cfg["object_rot"], terms_cfg["object_rot_delta"]
        a_rot = abs(rot["weight"]) * dt / rot.get("scale", 3.0) * cur
        a_delta = abs(delta["weight"]) * (cur + prev)
        if key == "terms":
            out = np.zeros((6, len(cur)))
            out[3], out[4] = a_rot, a_delta
            return out
        return a_rot * bool(rot["activate"]) + a_delta * bool(delta["activate"])

    def check(t, key, expected):
        got = env.observe(key)
        if key in ("obs", "pre_sim_obj_root"):
            seen[key] = np.asarray(expected, dtype=np.float64)
        if key == "info":
            assert set(got) == set(expected), (t, sorted(got), sorted(expected))
            for k, v in expected.items():
                tol = 1e-5 * abs(v) + (1e-4 if "rewards" in k else 1e-6)
                if abs(got[k] - v) > tol:
                    failures.append((t, k, got[k], v))
            return
        extra = angle_allowance(key) if key in ("terms", "reward") else None
        ok, detail = compare(key, got, expected, extra_atol=extra)
        if not ok:
            failures.append((t, key, detail))
        elif detail:
            worst[key] = max(worst.get(key, ""), detail)

    replay(env, g, check, inject=True)
    assert not failures, f"{name}: {failures[:8]}"
    return worst


@pytest.mark.parametrize("name", golden_names())
def test_fused_step_matches_reference(name):
    _run(name, fused=True)


@pytest.mark.parametrize("name", ["d3_goal_resample", "defaults_all_terms", "d4_reset30"])
def test_hook_by_hook_step_matches_reference(name):
    """Same fixtures through the individual hooks (_reset_impl, _goal_reset_impl, _pre_step,
    _post_step) in the reference's own sequencing (IsaacEnvBase.step)."""
    _run(name, fused=False)
