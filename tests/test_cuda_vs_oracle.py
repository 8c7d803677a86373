"""GPU parity against the CPU oracle at BASELINE.json sizes, plus size-independent properties:
sharding invariance, graph replay == eager, the wrapper's fused clipping, and the distribution of
the kernel's own Philox stream (SURVEY.md §D option B: stated distributional check)."""
import numpy as np
import pytest
import torch

from tolerances import compare

pytestmark = pytest.mark.gpu


def _pair(cfg, seq, N):
    from leibnizgym_b200.config import resolve_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    env.enable_term_rewards(True)
    ora = OracleEnv(resolve_config(cfg), OracleSim(seq, N))
    return env, ora


def _check(key, got, exp, where, extra=None):
    ok, detail = compare(key, got.cpu().numpy() if torch.is_tensor(got) else got,
                         exp.numpy() if torch.is_tensor(exp) else exp, extra_atol=extra)
    assert ok, (where, key, detail)


def _angle_allowance(ora):
    """Conditioning-aware slack of the two angle-derived terms and of the reward (tolerances.angle_slack)."""
    from tolerances import angle_slack
    cur = angle_slack(ora.obj_hist[0][:, 3:7], ora.goal_poses[:, 3:7])
    prev = angle_slack(ora.obj_hist[1][:, 3:7], ora.goal_poses[:, 3:7])
    t = ora.terms
    rot = abs(t["object_rot"]["weight"]) * 0.02 / t["object_rot"]["scale"] * cur      # |d term / d theta| <= w dt / scale
    delta = abs(t["object_rot_delta"]["weight"]) * (cur + prev)
    terms = np.zeros((6, len(cur)))
    terms[3], terms[4] = rot, delta
    reward = rot * t["object_rot"]["activate"] + delta * t["object_rot_delta"]["activate"]
    frac_ill = float(np.mean((cur > 1e-4) | (prev > 1e-4)))
    assert frac_ill < 0.03, frac_ill   # a material slack only for the ~1 % of pairs within ~0.01 rad of pi
    return terms, reward


@pytest.mark.parametrize("difficulty, N, reset_p", [(2, 16384, 0.0), (4, 16384, 0.3), (3, 65536, 0.05)])
def test_full_size_steps_match_oracle(difficulty, N, reset_p):
    """BASELINE.json configs 2, 5 and 3 at their real env counts, random draws injected on both sides."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.synthetic import bernoulli_masks, make_sequence
    T = 4
    cfg = difficulty_config(difficulty, N, asymmetric_obs=True, seed=5, episode_length=3)
    seq = make_sequence(100 + difficulty, T, N)
    masks = bernoulli_masks(7, T, N, reset_p)
    env, ora = _pair(cfg, seq, N)
    g = torch.Generator().manual_seed(3)
    draws = lambda k: (torch.rand(k, 24, generator=g).numpy(), torch.randn(k, 8, generator=g).numpy())  # noqa: E731
    d = draws(N)
    env.inject_draws(reset=d)
    ora.inject_draws(reset=d)
    env.reset()
    ora.reset()
    for t in range(1, T):
        if masks is not None:
            ora.reset_buf |= masks[t]
            env._reset_buf |= masks[t].cuda()
        k = int(ora.reset_buf.sum())
        d = draws(k) if k else None
        env.inject_draws(reset=d)
        ora.inject_draws(reset=d)
        env.step(seq.action[t].cuda())
        ora.step(seq.action[t].clone())
        w = (difficulty, t)
        _check("reset_ids", env.reset_env_ids, ora.last_ids[0], w)
        _check("obs", env.obs_buf, ora.obs_buf, w)
        _check("states", env.states_buf, ora.states_buf, w)
        slack_terms, slack_reward = _angle_allowance(ora)
        _check("terms", env._term_rewards[:6], ora.last_terms, w, slack_terms)
        _check("reward", env.reward_buf, ora.reward_buf, w, slack_reward)
        _check("reset_buf", env._reset_buf, ora.reset_buf, w)
        _check("steps_count", env._steps_count_buf, ora.steps_count_buf, w)
        _check("goal_pose", env._object_goal_poses_buf, ora.goal_poses, w)
        _check("pre_sim_dof", env._dof_state, ora.sim.dof, w)
        info = {k_: float(v) for k_, v in env._step_info.items()}
        for k_, v in ora.step_info.items():
            assert abs(info[k_] - float(v)) <= 1e-5 * abs(float(v)) + 1.35e-4, (w, k_, info[k_], float(v))   # tolerances.ATOL['terms']


def test_sharding_is_invisible():
    """Two shards (contiguous halves of the global env range) reproduce the un-sharded run bit for bit,
    including the kernel's own random stream (keyed by GLOBAL env id) and the schedule step."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import StateSequence, make_sequence
    N, T = 4096, 5
    cfg = difficulty_config(4, N, asymmetric_obs=True, seed=11, episode_length=2,
                            reset_distribution={"robot_initial_state": {"type": "random", "dof_pos_stddev": 0.4,
                                                                        "dof_vel_stddev": 0.2}})
    cfg["reward_terms"]["finger_reach_object_rate"].update(thresh_sched_start=0, thresh_sched_end=3 * N)
    seq = make_sequence(12, T, N)

    def shard(first, last):
        return StateSequence(seq.dof_state[:, first:last].contiguous(), seq.root_state[:, 4 * first:4 * last].contiguous(),
                             seq.rigid_body[:, first:last].contiguous(), seq.dof_force[:, first:last].contiguous(),
                             seq.ft_sensors[:, first:last].contiguous(), seq.action[:, first:last].contiguous())

    def run(rank, world):
        n = N // world
        sq = shard(rank * n, (rank + 1) * n).to("cuda:0")
        env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(sq, "cuda:0"), rank=rank, world_size=world)
        env.reset()
        outs = []
        for t in range(1, T):
            env.step(sq.action[t])
            outs.append([x.clone() for x in (env.obs_buf, env.states_buf, env.reward_buf, env._reset_buf,
                                             env._object_goal_poses_buf, env._dof_state)])
        return outs

    whole = run(0, 1)
    halves = [run(0, 2), run(1, 2)]
    for t in range(T - 1):
        for i in range(6):
            joined = torch.cat([halves[0][t][i], halves[1][t][i]], dim=0)
            assert torch.equal(joined, whole[t][i]), (t, i)


def test_graph_replay_equals_eager_steps():
    """CUDA-graph replay with the device-side clock (frame counter, reward coefficients, RNG epoch in
    device memory) gives the same buffers as eager launches driven by the host clock."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.graph_runner import GraphRunner
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    N, R = 2048, 4
    cfg = difficulty_config(4, N, asymmetric_obs=True, seed=21, episode_length=3)
    cfg["reward_terms"]["object_dist"].update(thresh_sched_start=0, thresh_sched_end=5 * N)  # gate flips inside the run

    def fresh(device_clock):
        ring = make_sequence(22, R, N, device="cuda:0")   # own copy: resets write into the ring
        env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(ring, "cuda:0"))
        env.reset()
        runner = GraphRunner(env, ring, rotate_outputs=False, device_clock=device_clock)
        runner.t = runner.t0 = env._sim.cursor
        return env, runner

    eager_env, eager = fresh(False)
    eager.step_eager(2 * R)
    graph_env, runner = fresh(True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            for t in range(R):
                runner._launch_step(runner.t + t, st)
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(runner.obs_slots[0], eager.obs_slots[0])
    assert torch.equal(runner.state_slots[0], eager.state_slots[0])
    assert torch.equal(graph_env.reward_buf, eager_env.reward_buf)
    assert torch.equal(graph_env._steps_count_buf, eager_env._steps_count_buf)
    assert torch.equal(graph_env._reset_buf, eager_env._reset_buf)
    assert torch.equal(graph_env._history, eager_env._history)
    assert torch.equal(graph_env._object_goal_poses_buf, eager_env._object_goal_poses_buf)  # same RNG epochs
    assert float(eager_env.reward_buf.abs().sum()) > 0


def test_vec_task_fused_clipping():
    """VecTaskPython's three clamps (vec_task.py:146-170) come out of the fused kernels."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    from leibnizgym_b200.wrappers import VecTaskPython
    N = 1000  # ragged last tile
    seq = make_sequence(31, 3, N)
    seq.dof_state[:, :, :, 1] *= 20.0           # joint velocities far outside +-10 -> |obs| > 5 after scaling
    env = TrifingerEnv(difficulty_config(4, N, seed=1), device="cuda:0", verbose=False,
                       sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    vec = VecTaskPython(env, rl_device="cuda:0", clip_obs=5.0, clip_actions=1.0)
    assert (vec.num_envs, vec.num_obs, vec.num_states, vec.num_actions) == (N, 41, 113, 9)
    vec.reset()
    act = 3.0 * seq.action[1].cuda()
    obs, rew, done, info = vec.step(act)
    assert torch.equal(obs, torch.clamp(env.obs_buf, -5.0, 5.0)) and float(obs.abs().max()) == 5.0
    assert torch.equal(vec.get_state(), torch.clamp(env.states_buf, -5.0, 5.0))
    assert torch.equal(env.action_buf, torch.clamp(act, -1.0, 1.0))
    assert obs.data_ptr() != env.obs_buf.data_ptr() and rew.shape == (N,) and done.dtype == torch.bool
    with pytest.raises(ValueError):
        vec.step(act[:, :8])


def test_rl_games_adapter_contract():
    """RlGamesGpuEnvAdapter (ref utils/rlg_train.py:89-154): persistent `full_state` dict with obs (+ states when
    asymmetric), `[[], info]` in slot 3, spaces from the wrapper; values equal the oracle's clipped outputs."""
    from leibnizgym_b200.config import difficulty_config, resolve_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    from leibnizgym_b200.wrappers import RlGamesGpuEnvAdapter, VecTaskPython
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    N, T = 777, 4
    seq = make_sequence(5, T, N)
    seq.dof_state[:, :, :, 1] *= 20.0            # |obs| > 5 after scaling: the clamp is exercised
    g = torch.Generator().manual_seed(11)
    d = (torch.rand(N, 24, generator=g).numpy(), torch.randn(N, 8, generator=g).numpy())
    clip = lambda x: torch.clamp(x, -5.0, 5.0)   # noqa: E731
    for asym in (True, False):
        cfg = difficulty_config(2, N, asymmetric_obs=asym, seed=3)
        made = []

        def creator(**kw):
            env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
            env.inject_draws(reset=d)
            made.append(env)
            return VecTaskPython(env, rl_device="cuda:0", clip_obs=5.0, clip_actions=1.0)

        ad = RlGamesGpuEnvAdapter("rlgpu", N, env_creator=creator)     # resets once (ref :100)
        env = made[-1]
        ora = OracleEnv(resolve_config(cfg), OracleSim(seq, N))
        ora.inject_draws(reset=d)
        ora.reset()
        info = ad.get_env_info()
        assert info["num_envs"] == N and info["action_space"].shape == (9,) and info["observation_space"].shape == (41,)
        assert ("state_space" in info) == asym and ad.get_number_of_agents() == 1
        env.inject_draws(reset=d)
        ora.inject_draws(reset=d)
        first = ad.reset()
        ora.reset()
        assert (first is ad.full_state) if asym else torch.is_tensor(first)
        _check("obs", first["obs"] if asym else first, clip(ora.obs_buf), ("reset", asym))
        for t in range(1, T):
            out, rew, done, extra = ad.step(seq.action[t].cuda())
            ora.step(seq.action[t].clone())
            w = ("adapter", asym, t)
            assert isinstance(extra, list) and len(extra) == 2 and extra[0] == [] and isinstance(extra[1], dict)
            assert (out is ad.full_state) if asym else torch.is_tensor(out)
            _check("obs", out["obs"] if asym else out, clip(ora.obs_buf), w)
            _check("reward", rew, ora.reward_buf, w)
            assert torch.equal(done.cpu(), ora.reset_buf & ora.goal_reset_buf)
            if asym:
                _check("states", out["states"], clip(ora.states_buf), w)
                assert out["states"].data_ptr() == env._states_clipped.data_ptr()   # no copy on the way to the learner
            assert "env/rewards/object_dist" in extra[1]
    with pytest.raises(ValueError):
        RlGamesGpuEnvAdapter("rlgpu", N)


def test_own_random_stream_distributions():
    """The samplers on the kernel's Philox stream: r^2/R^2, theta, z uniform; goal quaternion uniform on S^3
    (components^2 ~ Beta(1/2, 3/2)); joint noise uniform; deterministic in (seed, epoch, env)."""
    from scipy import stats
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    N = 200_000
    cfg = difficulty_config(4, N, asymmetric_obs=False, seed=77,
                            reset_distribution={"robot_initial_state": {"type": "random", "dof_pos_stddev": 0.4,
                                                                        "dof_vel_stddev": 0.2}})
    seq = make_sequence(1, 1, N, device="cuda:0")

    def sample(seed):
        c = dict(cfg, seed=seed)
        env = TrifingerEnv(c, device="cuda:0", verbose=False, sim=SyntheticSim(seq, "cuda:0"))
        env._reset_impl(torch.arange(N, device="cuda:0"))
        torch.cuda.synchronize()
        return (env._object_goal_poses_buf.cpu().numpy().astype(np.float64), env._dof_state.cpu().numpy().astype(np.float64),
                env._actors_root_state.view(N, 4, 13)[:, 2].cpu().numpy().astype(np.float64), env)

    goal, dof, obj, env = sample(77)
    R = 0.195 - 0.065 * np.sqrt(3) / 2
    ks = lambda x, cdf: stats.kstest(x, cdf).pvalue  # noqa: E731
    r2 = (goal[:, 0] ** 2 + goal[:, 1] ** 2) / R ** 2
    assert r2.max() <= 1.0 + 1e-6 and ks(r2, "uniform") > 1e-3
    theta = np.mod(np.arctan2(goal[:, 1], goal[:, 0]), 2 * np.pi) / (2 * np.pi)
    assert ks(theta, "uniform") > 1e-3
    z = (goal[:, 2] - 0.065 * np.sqrt(3) / 2) / (0.1 - 0.065 * np.sqrt(3) / 2)
    assert z.min() >= -1e-6 and z.max() <= 1 + 1e-6 and ks(z, "uniform") > 1e-3
    q = goal[:, 3:7]
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-6
    for c in range(4):
        assert ks(q[:, c] ** 2, stats.beta(0.5, 1.5).cdf) > 1e-3, c
    assert abs(np.corrcoef(q[:, 0], q[:, 1])[0, 1]) < 0.01 and abs(np.corrcoef(goal[:-1, 0], goal[1:, 0])[0, 1]) < 0.01
    # object: disc + yaw; robot: default + std * U(-1, 1)
    r2o = (obj[:, 0] ** 2 + obj[:, 1] ** 2) / R ** 2
    assert ks(r2o, "uniform") > 1e-3 and np.allclose(obj[:, 2], 0.0325)
    yaw = np.mod(2 * np.arctan2(obj[:, 5], obj[:, 6]), 2 * np.pi) / (2 * np.pi)
    assert ks(yaw, "uniform") > 1e-3
    default = np.array([0.0, 0.9, -1.7] * 3)
    u = (dof[:, :, 0] - default) / 0.4
    assert np.abs(u).max() <= 1 + 1e-6 and ks((u[:, 4] + 1) / 2, "uniform") > 1e-3
    assert ks((dof[:, 7, 1] / 0.2 + 1) / 2, "uniform") > 1e-3
    assert abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.01
    # determinism and seed sensitivity
    goal2, _, _, _ = sample(77)
    goal3, _, _, _ = sample(78)
    assert np.array_equal(goal, goal2) and not np.array_equal(goal, goal3)
    # a second reset call draws fresh numbers (the epoch advanced)
    env._reset_impl(torch.arange(N, device="cuda:0"))
    assert not np.array_equal(goal, env._object_goal_poses_buf.cpu().numpy().astype(np.float64))


@pytest.mark.parametrize("chunks,asym", [(1, True), (3, True), (2, False)])
def test_host_pipeline_equals_plain_step(chunks, asym):
    """The (optionally chunked) upload / compute / download step for host-resident simulators returns exactly
    what the plain step returns (chunks on ragged boundaries, resets and statistics included; with asymmetric
    observations the host observation is a view of the downloaded states)."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    from leibnizgym_b200.wrappers import VecTaskPython
    N, T = 10_000, 6
    cfg = difficulty_config(4, N, asymmetric_obs=asym, seed=17, episode_length=2)
    host = make_sequence(61, T, N).to("cpu", pin=True)

    def build(chunks):
        sim = SyntheticSim(host, "cuda:0")
        env = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=sim)
        sim.use_sparse_upload(env._P)
        vec = VecTaskPython(env, rl_device="cuda:0" if not chunks else "cpu", host_pipeline_chunks=chunks)
        vec.reset()
        return env, vec

    env_a, plain = build(0)
    env_b, piped = build(chunks)
    assert piped._pipeline.shared_obs == asym
    for t in range(1, T):
        o1, r1, d1, i1 = plain.step(host.action[t])
        s1 = plain.get_state() if asym else None
        o2, r2, d2, i2 = piped.step(host.action[t])
        s2 = piped.get_state() if asym else None
        torch.cuda.synchronize()
        assert torch.equal(o1.cpu(), o2) and torch.equal(r1.cpu(), r2) and torch.equal(d1.cpu(), d2)
        assert not asym or torch.equal(s1.cpu(), s2)
        assert torch.equal(env_a._reset_buf, env_b._reset_buf) and torch.equal(env_a._history, env_b._history)
        assert torch.equal(env_a._object_goal_poses_buf, env_b._object_goal_poses_buf)
        for k in i1:
            assert abs(float(i1[k]) - float(i2[k])) <= 1e-9 * max(1.0, abs(float(i1[k]))), k
    assert int(env_a._steps_count_buf.max()) <= 2 and o2.device.type == "cpu" and o2.is_pinned()


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """save_checkpoint after step k, load into a fresh env: steps k+1.. equal the uninterrupted run bit for bit —
    goal poses, counters, history, the kernel's own Philox epoch, the schedule clock."""
    import copy
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    from leibnizgym_b200.wrappers import VecTaskPython
    N, T = 3000, 12
    cfg = difficulty_config(4, N, seed=9, episode_length=3,
                            reset_distribution={"robot_initial_state": {"type": "random", "dof_pos_stddev": 0.4,
                                                                        "dof_vel_stddev": 0.2}},
                            termination_conditions={"success": {"activate": True, "bonus": 5000.0,
                                                                "position_tolerance": 0.2, "orientation_tolerance": 3.0}},
                            goal_movement={"rotation": {"activate": True, "rate_magnitude": 0.5}})
    seq = make_sequence(13, T, N)

    def build():
        env = TrifingerEnv(copy.deepcopy(cfg), device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
        return env, VecTaskPython(env, rl_device="cuda:0")

    def snap(env, vec, out):
        return [x.clone() for x in (out[0], out[1], out[2], vec.get_state(), env._object_goal_poses_buf, env._steps_count_buf,
                                    env._reset_buf, env._goal_reset_buf, env._successes, env.reset_env_ids,
                                    env._dof_state, env._actors_root_state, env._step_stats)]

    env_a, vec_a = build()
    vec_a.reset()
    for t in range(1, 6):
        vec_a.step(seq.action[t].cuda())
    path = env_a.save_checkpoint(str(tmp_path / "ckpt" / "env_state"))
    assert path.endswith(".npz")
    want = [snap(env_a, vec_a, vec_a.step(seq.action[t].cuda())) for t in range(6, T)]
    assert int(sum(w[9].numel() for w in want)) > 0            # resets (own random stream) happened after the checkpoint

    env_b, vec_b = build()                                      # fresh process state: never reset, clock at zero
    env_b.load_checkpoint(path)
    assert env_b.env_steps_count == 6 * N                       # reset() + 5 steps
    got = [snap(env_b, vec_b, vec_b.step(seq.action[t].cuda())) for t in range(6, T)]
    for t, (g, w) in enumerate(zip(got, want)):
        for i, (a, b) in enumerate(zip(g, w)):
            if i == len(g) - 1:   # the fp64 statistics are atomic sums over CTAs: order-dependent in the last bits
                assert torch.allclose(a, b, rtol=1e-12, atol=0.0), (t, i)
            else:
                assert torch.equal(a, b), (t, i)

    other = TrifingerEnv(difficulty_config(4, N + 1, seed=9), device="cuda:0", verbose=False,
                         sim=SyntheticSim(make_sequence(13, 2, N + 1).to("cuda:0"), "cuda:0"))
    with pytest.raises(ValueError):
        other.load_checkpoint(path)


def test_forced_reset_masks_equal_or_into_the_flag_buffers():
    """LgBuffers.force_reset / force_goal_reset (masks folded into lg_pre_physics) == `env._reset_buf |= mask` between steps."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import bernoulli_masks, make_sequence
    N, T = 5000, 6
    cfg = difficulty_config(4, N, seed=31)
    seq = make_sequence(9, T, N)
    rm = bernoulli_masks(3, T, N, 0.3, device="cuda:0")
    gm = bernoulli_masks(4, T, N, 0.1, device="cuda:0")
    a = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    b = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    a.reset()
    b.reset()
    for t in range(1, T):
        a._reset_buf |= rm[t]
        a._goal_reset_buf |= gm[t]
        b.set_forced_resets(rm[t].contiguous(), gm[t].contiguous())
        oa = a.step(seq.action[t].cuda())
        ob = b.step(seq.action[t].cuda())
        assert torch.equal(a.reset_env_ids, b.reset_env_ids) and torch.equal(a.goal_reset_env_ids, b.goal_reset_env_ids)
        assert len(a.reset_env_ids) > 0.2 * N
        for x, y in zip(oa[:3], ob[:3]):
            assert torch.equal(x, y)
        assert torch.equal(a.states_buf, b.states_buf) and torch.equal(a._object_goal_poses_buf, b._object_goal_poses_buf)
        assert torch.equal(a._dof_state, b._dof_state) and torch.equal(a._reset_buf, b._reset_buf)
    b.set_forced_resets(None, None)
    b.step(seq.action[1].cuda())
    assert len(b.reset_env_ids) == 0
    with pytest.raises(ValueError):
        b.set_forced_resets(torch.zeros(N, dtype=torch.uint8, device="cuda:0"))


@pytest.mark.parametrize("N", [1, 2, 3, 15, 17, 31, 33, 127, 129, 255, 18943, 18945])
@pytest.mark.parametrize("asym", [True, False])
def test_ragged_and_tiny_shards_match_oracle(N, asym):
    """Env counts around every tile boundary of both kernels (1 env, one short of / one past a 16-, 32-, 128-env tile,
    one wave +/- 1 env): obs, states, rewards, flags, counters and reset ids against the oracle, with 40 % resets."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.synthetic import bernoulli_masks, make_sequence
    T = 4
    cfg = difficulty_config(4, N, asymmetric_obs=asym, seed=5, episode_length=2)
    seq = make_sequence(200 + N, T, N)
    masks = bernoulli_masks(11, T, N, 0.4)
    env, ora = _pair(cfg, seq, N)
    g = torch.Generator().manual_seed(N)
    draws = lambda k: (torch.rand(k, 24, generator=g).numpy(), torch.randn(k, 8, generator=g).numpy())  # noqa: E731
    d = draws(N)
    env.inject_draws(reset=d)
    ora.inject_draws(reset=d)
    env.reset()
    ora.reset()
    for t in range(1, T):
        ora.reset_buf |= masks[t]
        env._reset_buf |= masks[t].cuda()
        k = int(ora.reset_buf.sum())
        d = draws(k) if k else None
        env.inject_draws(reset=d)
        ora.inject_draws(reset=d)
        env.step(seq.action[t].cuda())
        ora.step(seq.action[t].clone())
        w = (N, asym, t)
        _check("reset_ids", env.reset_env_ids, ora.last_ids[0], w)
        _check("obs", env.obs_buf, ora.obs_buf, w)
        if asym:
            _check("states", env.states_buf, ora.states_buf, w)
        slack_terms, slack_reward = _angle_allowance_small(ora)
        _check("reward", env.reward_buf, ora.reward_buf, w, slack_reward)
        _check("reset_buf", env._reset_buf, ora.reset_buf, w)
        _check("steps_count", env._steps_count_buf, ora.steps_count_buf, w)
        _check("pre_sim_dof", env._dof_state, ora.sim.dof.view(N, 9, 2), w)


def _angle_allowance_small(ora):
    """_angle_allowance without the population-level assertion (a handful of envs may all be ill-conditioned)."""
    from tolerances import angle_slack
    cur = angle_slack(ora.obj_hist[0][:, 3:7], ora.goal_poses[:, 3:7])
    prev = angle_slack(ora.obj_hist[1][:, 3:7], ora.goal_poses[:, 3:7])
    t = ora.terms
    rot = abs(t["object_rot"]["weight"]) * 0.02 / t["object_rot"]["scale"] * cur
    delta = abs(t["object_rot_delta"]["weight"]) * (cur + prev)
    terms = np.zeros((6, len(cur)))
    terms[3], terms[4] = rot, delta
    return terms, rot * t["object_rot"]["activate"] + delta * t["object_rot_delta"]["activate"]


def test_misaligned_action_views_are_accepted():
    """An action that is an offset view of a bigger buffer (not 16-byte aligned) steps like its aligned copy."""
    from leibnizgym_b200.config import difficulty_config
    from leibnizgym_b200.env import TrifingerEnv
    from leibnizgym_b200.sim import SyntheticSim
    from leibnizgym_b200.synthetic import make_sequence
    N = 640
    seq = make_sequence(2, 3, N)
    cfg = difficulty_config(2, N, seed=1)
    a = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    b = TrifingerEnv(cfg, device="cuda:0", verbose=False, sim=SyntheticSim(seq.to("cuda:0"), "cuda:0"))
    a.reset()
    b.reset()
    big = torch.zeros(N * 9 + 1, device="cuda:0")
    view = big[1:].view(N, 9)
    view.copy_(seq.action[1])
    assert view.data_ptr() % 16 != 0
    oa = a.step(seq.action[1].cuda())
    ob = b.step(view)
    for x, y in zip(oa[:3], ob[:3]):
        assert torch.equal(x, y)
    assert torch.equal(a.action_buf, b.action_buf)
