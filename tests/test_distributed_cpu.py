"""World-size-2 gloo test of the multi-GPU host logic: env sharding and the episode-statistics
all-reduce reproduce the single-process statistics (run on CPU; the CUDA path uses NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leibnizgym_b200 import _native as nat
from leibnizgym_b200.config import difficulty_config, resolve_config
from leibnizgym_b200.distributed import all_reduce_stats, shard_range, stats_to_info
from leibnizgym_b200.synthetic import StateSequence, make_sequence

N, T, SEED = 64, 4, 321


def _shard_of(seq: StateSequence, first: int, last: int) -> StateSequence:
    return StateSequence(seq.dof_state[:, first:last].contiguous(), seq.root_state[:, 4 * first:4 * last].contiguous(),
                         seq.rigid_body[:, first:last].contiguous(), seq.dof_force[:, first:last].contiguous(),
                         seq.ft_sensors[:, first:last].contiguous(), seq.action[:, first:last].contiguous())


def _stats_vector(env) -> torch.Tensor:
    """An oracle env's per-step statistics in the layout of LgBuffers.step_stats."""
    v = torch.zeros(nat.LG_NUM_STATS, dtype=torch.float64)
    for i, name in enumerate(nat.TERM_NAMES[:6]):
        if env.terms[name]["activate"]:
            v[i] = env.last_terms[i].double().mean()
    v[nat.STAT_POSITION_GOAL] = float(env.step_info["env/current_position_goal/count"])
    v[nat.STAT_ORIENTATION_GOAL] = float(env.step_info["env/current_orientation_goal/count"])
    v[nat.STAT_SUCCESSES] = env.successes.double().mean()
    v[nat.STAT_REWARD] = env.reward_buf.double().mean()
    return v


def _run_oracle(cfg, seq, n, draws):
    from oracle.trifinger_oracle import OracleEnv, OracleSim
    env = OracleEnv(cfg, OracleSim(seq, n))
    env.inject_draws(reset=draws)
    env.reset()
    out = []
    for t in range(1, T):
        env.step(seq.action[t].clone())
        out.append(_stats_vector(env))
    return out, env


def _config(n):
    # difficulty 2: the goal is fixed, so goal resets after a success draw no random numbers and the
    # shards stay comparable with the single-process run
    cfg = resolve_config(difficulty_config(2, n, seed=SEED))
    cfg["reward_terms"]["object_rot"]["activate"] = True
    cfg["termination_conditions"]["success"].update(activate=True, position_tolerance=0.12, orientation_tolerance=1.5)
    return cfg


def _draws():
    g = torch.Generator().manual_seed(SEED)
    return torch.rand(N, 24, generator=g).numpy(), torch.randn(N, 8, generator=g).numpy()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    first, last = shard_range(N, rank, world)
    seq = _shard_of(make_sequence(SEED, T, N), first, last)
    u, n = _draws()
    cfg = _config(last - first)
    local, env = _run_oracle(cfg, seq, last - first, (u[first:last], n[first:last]))
    merged = [all_reduce_stats(v, last - first, N) for v in local]
    if rank == 0:
        active = [k for k, v in cfg["reward_terms"].items() if v["activate"]]
        q.put(([m.numpy() for m in merged], stats_to_info(merged[-1], active)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_equal_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, info = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole, env = _run_oracle(_config(N), make_sequence(SEED, T, N), N, _draws())
    for got, exp in zip(merged, whole):
        np.testing.assert_allclose(got, exp.numpy(), rtol=1e-12, atol=1e-12)
    assert whole[-1][nat.STAT_POSITION_GOAL] > 0          # the scenario does exercise the counters
    assert info["env/current_position_goal/count"] == float(whole[-1][nat.STAT_POSITION_GOAL])
    assert set(info) >= {"env/rewards/object_dist", "env/average_consecutive_success"}


def test_shard_range():
    assert shard_range(262144, 3, 8) == (98304, 131072)
    assert [shard_range(16, r, 4) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    import pytest
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
