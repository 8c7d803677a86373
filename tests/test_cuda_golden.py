"""GPU parity against the UNMODIFIED reference: every fixture of tests/golden is replayed
through the CUDA TrifingerEnv (C ABI) with the reference's random draws injected."""
import pytest

from golden_io import Golden, golden_names, replay
from tolerances import compare

pytestmark = pytest.mark.gpu


def _run(name, fused):
    from adapters import CudaAdapter
    g = Golden(name)
    env = CudaAdapter(g.config, g.sequence(), fused=fused)
    failures, worst = [], {}

    def check(t, key, expected):
        got = env.observe(key)
        if key == "info":
            assert set(got) == set(expected), (t, sorted(got), sorted(expected))
            for k, v in expected.items():
                tol = 1e-5 * abs(v) + (1e-4 if "rewards" in k else 1e-6)
                if abs(got[k] - v) > tol:
                    failures.append((t, k, got[k], v))
            return
        ok, detail = compare(key, got, expected)
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
