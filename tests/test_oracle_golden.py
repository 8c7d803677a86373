"""The oracle must reproduce the unmodified reference (tests/golden/*.npz) exactly.

Two modes per scenario: with the reference's recorded random draws injected, and with
the oracle drawing from torch's seeded CPU generator in the reference's order (which
yields the same numbers, proving the draw order of SURVEY.md §A.5)."""
import numpy as np
import pytest

from adapters import OracleAdapter
from golden_io import Golden, golden_names, replay


def _exact(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a.astype(np.int64) == b.astype(np.int64)
    return bool(np.all(same))


@pytest.mark.parametrize("inject", [True, False], ids=["injected-draws", "seeded-generator"])
@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference(name, inject):
    g = Golden(name)
    env = OracleAdapter(g.config, g.sequence())
    seen = []

    def check(t, key, expected):
        got = env.observe(key)
        if key == "info":
            assert set(got) == set(expected), (t, got.keys(), expected.keys())
            for k in expected:
                assert got[k] == expected[k], (name, t, k, got[k], expected[k])
        else:
            assert _exact(got, expected), (name, t, key)
        seen.append(key)

    replay(env, g, check, inject=inject)
    assert {"obs", "reward", "terms", "reset_ids", "steps_count"} <= set(seen)
