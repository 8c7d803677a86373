"""Parity tolerances (BASELINE.json north_star): masks / ids / counters bit-exact;
rewards, observations and states within 1e-5 relative for fp32.

A small absolute floor is needed next to the relative bound because several quantities are
differences of nearly equal numbers (norm - previous norm, |theta| - |previous theta|) whose
value can be arbitrarily close to zero while each operand carries a 1-ulp sampling error
(sin/cos/exp/asin differ by <= 2 ulp between CUDA libdevice and torch's SLEEF).  The floor
is stated per key, relative to the natural magnitude of the quantity."""
import numpy as np

RTOL = 1e-5

EXACT = {"reset_in", "goal_reset_in", "reset_ids", "goal_reset_ids", "dof_index_list", "root_index_list0",
         "root_index_list1", "reset_buf", "goal_reset_buf", "steps_count", "successes", "sched_step"}

# absolute floors
ATOL = {
    "obs": 1e-6, "states": 1e-6, "action_buf": 0.0, "applied_torque": 1e-7,
    "goal_pose": 1e-7, "goal_movement": 1e-7,
    "pre_sim_dof": 1e-7, "pre_sim_obj_root": 1e-7, "pre_sim_goal_root": 1e-7,
    "terms": 2e-4, "reward": 5e-4,
}


def compare(key, got, expected):
    """Returns (ok, detail)."""
    got, expected = np.asarray(got), np.asarray(expected)
    if got.shape != expected.shape:
        return False, f"shape {got.shape} != {expected.shape}"
    if key in EXACT:
        ok = np.array_equal(got.astype(np.int64), expected.astype(np.int64))
        return ok, "" if ok else f"{int((got.astype(np.int64) != expected.astype(np.int64)).sum())} entries differ"
    g, x = got.astype(np.float64), expected.astype(np.float64)
    err = np.abs(g - x)
    bound = ATOL[key] + RTOL * np.abs(x)
    bad = ~((err <= bound) | (np.isnan(g) & np.isnan(x)))
    if bad.any():
        i = np.unravel_index(np.argmax(np.where(bad, err, 0)), err.shape)
        return False, f"{int(bad.sum())} entries off; worst at {i}: got {g[i]!r} expected {x[i]!r} err {err[i]:.3e}"
    return True, f"max abs err {err.max() if err.size else 0:.3e}"
