"""The two extension features the north star names but the reference does not implement:
the keypoint pose reward (against oracle/extensions.py) and the DR observation / action noise (stated
distributional check).  Both default off; switched off they leave the path bit-identical."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cfg, seq, N):
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    env.enable_term_rewards(True)
    return env


def test_keypoint_reward_term():
    from leibnizgym_b200.config import difficulty_config, resolve_config
    from leibnizgym_b200.synthetic import make_sequence
    from oracle.extensions import keypoint_reward
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    N, T = 4096, 3
    base = difficulty_config(4, N, seed=3)
    cfg = copy.deepcopy(base)
    cfg["reward_terms"]["keypoint"] = {"activate": True, "weight": 1500, "scale": 30.0, "eps": 2.0}
    seq = make_sequence(41, T, N)
    # bring a third of the cubes close to their goals so the kernel term is not all ~0
    env = _make(cfg, seq, N)
    ora = OracleEnv(resolve_config(base), OracleSim(seq, N))
    g = torch.Generator().manual_seed(1)
    d = (torch.rand(N, 24, generator=g).numpy(), torch.randn(N, 8, generator=g).numpy())
    env.inject_draws(reset=d)
    ora.inject_draws(reset=d)
    env.reset()
    ora.reset()
    goal = ora.goal_poses.clone()
    near = torch.arange(0, N, 3)
    root = seq.root_state.view(T, N, 4, 13)
    root[1:, near, 2, 0:3] = goal[near, 0:3] + 0.01 * torch.randn(len(near), 3, generator=g)
    root[1:, near, 2, 3:7] = torch.nn.functional.normalize(goal[near, 3:7] + 0.05 * torch.randn(len(near), 4, generator=g), dim=-1)
    env._sim.seq = seq.to("cuda:0")
    for t in range(1, T):
        env.step(seq.action[t].cuda())
        ora.step(seq.action[t].clone())
        exp_kp = keypoint_reward(1500, 0.02, ora.obj_hist[0][:, 0:7], ora.goal_poses)
        got_kp = env._term_rewards[6].cpu()
        assert float(exp_kp.max()) > 1.0                                  # the term is exercised
        np.testing.assert_allclose(got_kp.numpy(), exp_kp.numpy(), rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(env.reward_buf.cpu().numpy(), (ora.reward_buf + exp_kp).numpy(), rtol=2e-5, atol=5e-4)
        info = env._step_info
        assert abs(float(info["env/rewards/keypoint"]) - float(exp_kp.double().mean())) < 1e-4
        assert abs(float(info["env/rewards/object_dist"]) - float(ora.step_info["env/rewards/object_dist"])) < 1e-4


def test_dr_noise_off_is_bit_identical_and_on_is_gaussian():
    from scipy import stats
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.params import observation_scale
    from leibnizgym_b200.config import resolve_config
    from leibnizgym_b200.synthetic import make_sequence
    N, T = 120_000, 3
    seq = make_sequence(51, T, N)
    seq.action.mul_(0.5)                                              # keep the noisy action inside the clamp
    base = difficulty_config(3, N, seed=9)
    sig = {"robot_q": 0.01, "robot_u": 0.05, "object_q": 0.002, "object_q_des": 0.0, "command": 0.0}
    dr_off = dict(copy.deepcopy(base), domain_randomization={"activate": True, "action_noise_std": 0.0,
                                                             "obs_noise_std": {k: 0.0 for k in sig}})
    dr_on = dict(copy.deepcopy(base), domain_randomization={"activate": True, "action_noise_std": 0.03, "obs_noise_std": sig})

    def run(cfg, steps=2):
        env = _make(cfg, seq, N)
        env.enable_clipped_outputs(5.0, 1.0)
        env.reset()
        outs = []
        for t in range(1, 1 + steps):
            env.step(seq.action[t].cuda())
            outs.append((env.obs_buf.clone(), env.states_buf.clone(), env.action_buf.clone(), env.reward_buf.clone()))
        return outs

    clean, off, on, on2 = run(base), run(dr_off), run(dr_on), run(dr_on)
    for a, b in zip(clean, off):
        assert all(torch.equal(x, y) for x, y in zip(a, b))          # sigma = 0 -> the reference path, bit for bit
    for a, b in zip(on, on2):
        assert all(torch.equal(x, y) for x, y in zip(a, b))          # deterministic in (seed, step, env)
    lo, hi = observation_scale(resolve_config(base))
    half = torch.tensor((hi - lo) / 2).cuda()
    resid = [((on[t][0] - clean[t][0]) * half).cpu().numpy().astype(np.float64) for t in range(2)]
    assert torch.equal(on[0][1][:, 41:], clean[0][1][:, 41:])        # privileged part of the critic state stays clean
    assert torch.equal(on[0][1][:, :32], clean[0][1][:, :32])        # ... and so does its copy of the observation
    assert not torch.equal(on[0][1][:, 32:41], clean[0][1][:, 32:41])  # (the stored action itself is the noisy one)
    for col, s in ((0, 0.01), (5, 0.01), (9, 0.05), (17, 0.05), (18, 0.002), (24, 0.002)):
        r = resid[0][:, col]
        assert abs(r.mean()) < 4 * s / np.sqrt(N) and abs(r.std() / s - 1) < 0.02, (col, r.mean(), r.std())
        assert stats.kstest(r / s, "norm").pvalue > 1e-3, col
    assert np.abs(resid[0][:, 25:32]).max() == 0.0                   # sigma 0 group (goal pose) untouched
    # independence: across steps, across envs, across columns
    assert abs(np.corrcoef(resid[0][:, 9], resid[1][:, 9])[0, 1]) < 0.01
    assert abs(np.corrcoef(resid[0][:-1, 9], resid[0][1:, 9])[0, 1]) < 0.01
    assert abs(np.corrcoef(resid[0][:, 9], resid[0][:, 10])[0, 1]) < 0.01
    # action noise: additive N(0, 0.03^2) before the clamp (|action| <= 0.5 here, so the clamp is inactive)
    ra = (on[0][2] - clean[0][2]).cpu().numpy().astype(np.float64)
    assert abs(ra.std() / 0.03 - 1) < 0.02 and stats.kstest(ra[:, 3] / 0.03, "norm").pvalue > 1e-3


def test_goal_integrator_matches_its_oracle_and_rotates_at_the_imposed_rate():
    """lg_integrate_goal (stand-in for PhysX on the moving goal body) against oracle/extensions.integrate_goal_rows;
    properties: unit quaternions, rotation angle after K steps == |w| K dt, zero angular velocity is the identity."""
    from leibnizgym_b200 import _native as nat
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.params import build_params
    from oracle.extensions import integrate_goal_rows
    N, dt, K = 4099, 0.02, 25
    lib = nat.load()
    from leibnizgym_b200.config import resolve_config
    P = build_params(resolve_config(difficulty_config(4, N)), N)
    g = torch.Generator().manual_seed(9)
    root = torch.zeros(N, 4, 13)
    root[..., 6] = 1.0
    q = torch.randn(N, 4, generator=g)
    root[:, 3, 3:7] = q / q.norm(dim=1, keepdim=True)
    root[:, 3, 0:3] = torch.rand(N, 3, generator=g)
    root[:, 3, 7:10] = 0.1 * torch.randn(N, 3, generator=g)
    root[:, 3, 10:13] = 0.5 * torch.randn(N, 3, generator=g)
    root[:7, 3, 10:13] = 0.0                               # w = 0: orientation must not move
    root[7, 3, 10:13] = torch.tensor([0.0, 0.0, 3.0])
    dev = root.cuda().contiguous()
    S = nat.LgSimState(None, dev.data_ptr(), None, None, None)
    q0 = dev[:, 3, 3:7].clone()
    exp = root[:, 3].numpy().astype(np.float64)
    for _ in range(K):
        nat.check(lib.lg_integrate_goal(P, S, dt, torch.cuda.current_stream().cuda_stream), "lg_integrate_goal")
        exp = integrate_goal_rows(exp, dt)
    got = dev[:, 3].cpu().numpy().astype(np.float64)
    assert np.abs(got - exp).max() < 5e-6, np.abs(got - exp).max()
    assert torch.equal(dev[:, :3], root[:, :3].cuda())     # other actors untouched
    assert np.abs(np.linalg.norm(got[:, 3:7], axis=1) - 1.0).max() < 1e-6
    assert torch.equal(dev[:7, 3, 3:7], q0[:7])
    ang = torch.zeros(N, device="cuda")
    nat.check(lib.lg_quat_diff_rad(dev[:, 3, 3:7].contiguous().data_ptr(), q0.contiguous().data_ptr(), ang.data_ptr(), N,
                                   torch.cuda.current_stream().cuda_stream), "lg_quat_diff_rad")
    want = (root[:, 3, 10:13].norm(dim=1) * K * dt).numpy()
    sel = want < 3.0                                        # below pi the geodesic angle is the integrated one
    assert np.abs(ang.cpu().numpy()[sel] - want[sel]).max() < 2e-4


def test_moving_goal_end_to_end_with_the_integrating_simulator():
    """goal_movement.rotation on top of an integrating simulator: the goal pose buffer follows the goal body, which
    turns at the sampled angular velocity; CUDA env == oracle env over the same (integrated) simulator state."""
    from leibnizgym_b200.config import difficulty_config, resolve_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    from oracle.extensions import integrate_goal_rows
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    from tolerances import compare
    N, T, dt = 1500, 6, 0.02
    cfg = difficulty_config(4, N, seed=21, goal_movement={"rotation": {"activate": True, "rate_magnitude": 0.5}})
    seq = make_sequence(77, T, N)

    class IntegratingOracleSim(OracleSim):
        def simulate(self):
            keep = self.root.view(N, 4, 13)[:, 3].clone()
            super().simulate()
            rows = integrate_goal_rows(keep.numpy(), dt)
            self.root.view(N, 4, 13)[:, 3] = torch.from_numpy(rows.astype(np.float32))

    sim = SyntheticSim(seq.to("cuda:0"), "cuda:0")
    env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=sim)
    sim.enable_goal_integration(env._P, dt)
    ora = OracleEnv(resolve_config(cfg), IntegratingOracleSim(seq, N))
    g = torch.Generator().manual_seed(4)
    d = (torch.rand(N, 24, generator=g).numpy(), torch.randn(N, 8, generator=g).numpy())
    env.inject_draws(reset=d)
    ora.inject_draws(reset=d)
    env.reset()
    ora.reset()
    w = env._object_goal_movement_buf[:, 3:6].clone()
    assert float(w.norm(dim=1).min()) > 0.0
    q_start = env._object_goal_poses_buf[:, 3:7].clone()
    for t in range(1, T):
        env.step(seq.action[t].cuda())
        ora.step(seq.action[t].clone())
        for key, a, b in (("obs", env.obs_buf, ora.obs_buf), ("states", env.states_buf, ora.states_buf),
                          ("goal_pose", env._object_goal_poses_buf, ora.goal_poses)):
            ok, detail = compare(key, a.cpu().numpy(), b.numpy(), extra_atol=2e-6)
            assert ok, (t, key, detail)
        ok, detail = compare("reward", env.reward_buf.cpu().numpy(), ora.reward_buf.numpy(),
                             extra_atol=1e-4 * np.abs(ora.reward_buf.numpy()).max())
        assert ok, (t, detail)
        # the goal pose the env holds IS the goal body's row, and the imposed angular velocity survives the step
        assert torch.equal(env._object_goal_poses_buf, env._actors_root_state.view(N, 4, 13)[:, 3, 0:7])
        assert torch.equal(env._actors_root_state.view(N, 4, 13)[:, 3, 10:13], w)
    # reset() integrated the sampled pose once, each of the T-1 steps once more: total rotation |w| T dt
    ang = torch.zeros(N, device="cuda")
    from leibnizgym_b200 import _native as nat
    nat.check(nat.load().lg_quat_diff_rad(env._object_goal_poses_buf[:, 3:7].contiguous().data_ptr(), q_start.data_ptr(),
                                          ang.data_ptr(), N, torch.cuda.current_stream().cuda_stream), "lg_quat_diff_rad")
    want = (w.norm(dim=1) * T * dt)
    assert float((ang - want).abs().max()) < 2e-4


@pytest.mark.parametrize("asym", [True, False])
def test_bf16_outputs_are_the_rounded_fp32_outputs(asym):
    """Optional bf16 emission: bit-identical to `.to(torch.bfloat16)` of the clipped fp32 outputs,
    through the wrapper; the fp32 outputs themselves do not change when it is switched on."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    from leibnizgym_b200.wrappers import VecTaskPython
    N, T = 2077, 4
    seq = make_sequence(3, T, N)
    seq.dof_state[:, :, :, 1] *= 20.0   # some values beyond the clamp
    cfg = difficulty_config(3, N, asymmetric_obs=asym, seed=8)

    def build(dtype):
        env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
        return env, VecTaskPython(env, rl_device="cuda:0", obs_dtype=dtype)

    env_f, vec_f = build(torch.float32)
    env_b, vec_b = build(torch.bfloat16)
    o_f, o_b = vec_f.reset(), vec_b.reset()
    assert o_b.dtype == torch.bfloat16 and torch.equal(o_b, o_f.to(torch.bfloat16))
    for t in range(1, T):
        of, rf, df, _ = vec_f.step(seq.action[t].cuda())
        ob, rb, db, _ = vec_b.step(seq.action[t].cuda())
        assert ob.dtype == torch.bfloat16 and ob.shape == of.shape
        assert torch.equal(ob, of.to(torch.bfloat16)) and torch.equal(rf, rb) and torch.equal(df, db)
        assert torch.equal(env_f.obs_buf, env_b.obs_buf) and torch.equal(env_f._obs_clipped, env_b._obs_clipped)
        if asym:
            sb = vec_b.get_state()
            assert sb.dtype == torch.bfloat16 and torch.equal(sb, vec_f.get_state().to(torch.bfloat16))
            assert torch.equal(env_f.states_buf, env_b.states_buf)
    assert float(of.abs().max()) == 5.0
    with pytest.raises(ValueError):
        VecTaskPython(env_f, rl_device="cuda:0", obs_dtype=torch.float16)
