"""CPU ORACLE for the TriFinger MDP hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A torch (CPU, fp32) restatement of the reference algorithm for the path named by
BASELINE.json: reward + observation/state + reset of pairlab/leibnizgym.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module; the product path
(`leibnizgym_b200/`) never does and fails loudly without its CUDA library.

Pinning: the reference ships no golden vectors (its tests hold no assertion,
SURVEY.md §4), so this restatement is pinned against outputs of the UNMODIFIED
reference run in the build container: `tests/golden/make_golden.py` drives the
real `TrifingerEnv` over synthetic simulator state and stores inputs, outputs
and the random draws in `tests/golden/*.npz`; `tests/test_oracle_golden.py`
requires this oracle to reproduce every stored array exactly (same torch ops in
the same order, so the comparison is bit-for-bit on the build box).  The two
extension features without a reference implementation (keypoint reward, DR
noise) are "parity unpinned" and are restated in oracle/extensions.py instead.

Op granularity deliberately follows the reference (one ATen op per reference
op) so that timing this module on host cores is a fair stand-in for the
reference's CPU torch path (bench.py `cpu_baseline.kind == "port"`).

Every function cites the reference lines it restates (paths relative to
/root/reference/leibnizgym/).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch

F32 = torch.float32

# ----------------------------------------------------------------------------
# constants (ref envs/trifinger/trifinger_env.py:143-224, envs/trifinger/utils.py:54-131)
# ----------------------------------------------------------------------------
CUBE_SIZE = 0.065
ARENA_RADIUS = 0.195
CUBE_RADIUS_3D = CUBE_SIZE * np.sqrt(3) / 2
MAX_COM_DISTANCE = ARENA_RADIUS - CUBE_RADIUS_3D
CUBE_MIN_HEIGHT = CUBE_SIZE / 2
CUBE_MAX_HEIGHT = 0.1
MAX_TORQUE = 0.36
MAX_VELOCITY = 10

JOINT_LOW = [-0.33, 0.0, -2.7] * 3
JOINT_HIGH = [1.0, 1.57, 0.0] * 3
JOINT_DEFAULT = [0.0, 0.9, -1.7] * 3
KP = [10.0, 10.0, 10.0] * 3
KD = [0.1, 0.3, 0.001] * 3
SAFETY_KD = [0.08, 0.08, 0.04] * 3
STIFFNESS_LOW, STIFFNESS_HIGH = [1.0] * 9, [50.0] * 9

TERM_ORDER = ("finger_reach_object_rate", "finger_move_penalty", "object_dist",
              "object_rot", "object_rot_delta", "object_move")


def _t(x):
    return torch.tensor(np.asarray(x, dtype=np.float32))


# ----------------------------------------------------------------------------
# math primitives (ref utils/torch_utils.py)
# ----------------------------------------------------------------------------
def scale_transform(x, lower, upper):
    """ref utils/torch_utils.py:18-36"""
    centre = (lower + upper) * 0.5
    return 2 * (x - centre) / (upper - lower)


def unscale_transform(x, lower, upper):
    """ref utils/torch_utils.py:39-57"""
    centre = (lower + upper) * 0.5
    return x * (upper - lower) * 0.5 + centre


def saturate(x, lower, upper):
    """ref utils/torch_utils.py:60-75"""
    return torch.max(torch.min(x, upper), lower)


def quat_mul(a, b):
    """Hamilton product, xyzw, 8-multiplication form (ref utils/torch_utils.py:83-113)."""
    shape = a.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 4)
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bx, by, bz, bw = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    t_ww = (az + ax) * (bx + by)
    t_yy = (aw - ay) * (bw + bz)
    t_zz = (aw + ay) * (bw - bz)
    t_xx = t_ww + t_yy + t_zz
    half = 0.5 * (t_xx + (az - ax) * (bx - by))
    w = half - t_ww + (az - ay) * (by - bz)
    x = half - t_xx + (ax + aw) * (bx + bw)
    y = half - t_yy + (aw - ax) * (by + bz)
    z = half - t_zz + (az + ay) * (bw - bx)
    return torch.stack([x, y, z, w], dim=-1).view(shape)


def quat_conjugate(a):
    """ref utils/torch_utils.py:116-128"""
    shape = a.shape
    a = a.reshape(-1, 4)
    return torch.cat((-a[:, :3], a[:, -1:]), dim=-1).view(shape)


def quat_diff_rad(a, b):
    """theta = 2 asin(min(|(a (x) conj b)_xyz|, 1))  (ref utils/torch_utils.py:131-150)"""
    prod = quat_mul(a, quat_conjugate(b))
    return 2.0 * torch.asin(torch.clamp(torch.norm(prod[:, 0:3], p=2, dim=-1), max=1.0))


def quat_from_euler_xyz(roll, pitch, yaw):
    """ref utils/torch_utils.py:153-180"""
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


# ----------------------------------------------------------------------------
# samplers (ref envs/trifinger/sample.py); `draw` supplies the random numbers so a
# test can inject the ones the reference consumed
# ----------------------------------------------------------------------------
class Draws:
    """Source of random numbers.  Default: torch's global CPU generator, called with the
    reference's shapes in the reference's order — after the same `torch.manual_seed` it
    yields the very numbers the reference draws.  With `inject(u, n)` the next reset reads
    columns of the canonical injected-draw arrays instead (layout: include/leibniz_b200.h)."""

    def __init__(self):
        self.u = self.n = None

    def inject(self, u, n):
        self.u = None if u is None else torch.as_tensor(u, dtype=F32)
        self.n = None if n is None else torch.as_tensor(n, dtype=F32)

    def clear(self):
        self.u = self.n = None

    def uniform(self, k, cols):
        if self.u is None:
            return torch.rand(k, dtype=F32) if isinstance(cols, int) else torch.rand((k, len(cols)), dtype=F32)
        return self.u[:, cols].clone() if isinstance(cols, int) else self.u[:, list(cols)].clone()

    def normal(self, k, cols):
        if self.n is None:
            return torch.randn((k, len(cols)), dtype=F32)
        return self.n[:, list(cols)].clone()


def random_xy(draw: Draws, k: int, max_radius: float, cols: Tuple[int, int]):
    """Uniform on a disc (ref envs/trifinger/sample.py:22-34)."""
    radius = torch.sqrt(draw.uniform(k, cols[0]))
    radius *= max_radius
    theta = 2 * np.pi * draw.uniform(k, cols[1])
    return radius * torch.cos(theta), radius * torch.sin(theta)


def random_z(draw: Draws, k: int, lo: float, hi: float, col: int):
    """ref envs/trifinger/sample.py:37-43"""
    z = draw.uniform(k, col)
    return (hi - lo) * z + lo


def default_orientation(k: int):
    """ref envs/trifinger/sample.py:46-52"""
    q = torch.zeros((k, 4), dtype=F32)
    q[..., -1] = 1.0
    return q


def random_orientation(draw: Draws, k: int):
    """normalize(randn(k,4)) (ref envs/trifinger/sample.py:55-65)"""
    q = draw.normal(k, (0, 1, 2, 3))
    return torch.nn.functional.normalize(q, p=2.0, dim=-1, eps=1e-12)


def random_angular_vel(draw: Draws, k: int, magnitude_std: float):
    """ref envs/trifinger/sample.py:67-75"""
    axis = draw.normal(k, (4, 5, 6))
    axis /= torch.norm(axis, p=2, dim=-1).view(-1, 1)
    magnitude = draw.normal(k, (7,))
    magnitude *= magnitude_std
    return magnitude * axis


def random_yaw_orientation(draw: Draws, k: int, col: int):
    """ref envs/trifinger/sample.py:77-84"""
    zero = torch.zeros(k, dtype=F32)
    yaw = 2 * np.pi * draw.uniform(k, col)
    return quat_from_euler_xyz(zero, torch.zeros(k, dtype=F32), yaw)


# ----------------------------------------------------------------------------
# reward terms (ref envs/trifinger/rewards.py)
# ----------------------------------------------------------------------------
def lgsk(x, scale: float = 50.0):
    """1 / (e^{sx} + 2 + e^{-sx}) (ref envs/trifinger/rewards.py:20-34)."""
    sx = x * scale
    return 1.0 / (sx.exp() + 2 + (-sx).exp())


def _gate(start: float, end: float, T: float) -> float:
    """Threshold schedule (ref rewards.py:56-60, :123-127, :229-233)."""
    if start != end:
        return 1.0 if start <= T <= end else 0.0
    return 1.0


def _ramp(start: float, end: float, T: float) -> float:
    """Linear schedule (ref rewards.py:14-17, :169-172)."""
    if start != end:
        return max(0.0, min(1.0, (T - start) / (end - start)))
    return 1.0


def finger_reach_object_rate(p, T, tips, tips_prev, obj, obj_prev):
    """ref rewards.py:204-235"""
    cur = torch.stack([torch.norm(tips[:, i, 0:3] - obj[:, 0:3], p=p["norm_p"], dim=-1)
                       for i in range(3)], dim=-1)
    prev = torch.stack([torch.norm(tips_prev[:, i, 0:3] - obj_prev[:, 0:3], p=p["norm_p"], dim=-1)
                        for i in range(3)], dim=-1)
    return p["weight"] * _gate(p["start"], p["end"], T) * (cur - prev).sum(dim=-1)


def finger_move_penalty(p, dt, tips, tips_prev):
    """ref rewards.py:248-263"""
    vel = (tips[:, :, 0:3] - tips_prev[:, :, 0:3]) / dt
    return p["weight"] * vel.pow(2).view(-1, 9).sum(dim=-1)


def object_dist(p, dt, T, obj, goal):
    """ref rewards.py:53-63"""
    dist = torch.norm(obj[:, 0:3] - goal[:, 0:3], p=2, dim=-1)
    return p["weight"] * dt * _gate(p["start"], p["end"], T) * lgsk(dist)


def object_rot(p, dt, T, obj, goal):
    """ref rewards.py:119-139"""
    angles = quat_diff_rad(obj[:, 3:7], goal[:, 3:7])
    rew = _gate(p["start"], p["end"], T) * dt / (p["scale"] * torch.abs(angles) + p["scale"])
    return p["weight"] * rew


def object_rot_delta(p, dt, T, obj, obj_prev, goal):
    """ref rewards.py:165-184"""
    last = torch.abs(quat_diff_rad(obj_prev[:, 3:7], goal[:, 3:7]))
    cur = torch.abs(quat_diff_rad(obj[:, 3:7], goal[:, 3:7]))
    rew = _ramp(p["start"], p["end"], T) * (cur - last)
    return p["weight"] * rew


def object_move(p, obj, obj_prev, goal):
    """ref rewards.py:75-91"""
    cur = torch.norm(obj[:, 0:3] - goal[:, 0:3], dim=-1)
    prev = torch.norm(obj_prev[:, 0:3] - goal[:, 0:3], dim=-1)
    return p["weight"] * (cur - prev)


def reward_term_params(cfg_terms: dict) -> "OrderedDict[str, dict]":
    """Constructor-time parameters of the six terms (ref rewards.py:40-51, :68-73, :102-117,
    :150-163, :190-202, :241-246).  Iteration order = the user's dict order, as in the
    reference (trifinger_env.py:344-350); accumulation order is fixed (TERM_ORDER)."""
    out = OrderedDict()
    for name, c in cfg_terms.items():
        if name not in TERM_ORDER:
            continue
        p = {"activate": bool(c["activate"]), "weight": c.get("weight")}
        if name == "object_rot_delta":
            p["start"] = float(c.get("linear_schedule_start", 0))
            p["end"] = float(c.get("linear_schedule_end", 0))
        else:
            p["start"] = float(c.get("thresh_sched_start", 0))
            p["end"] = float(c.get("thresh_sched_end", 0))
        defaults = {"finger_reach_object_rate": -250, "finger_move_penalty": -1.0e-4, "object_dist": 2000,
                    "object_rot": 100, "object_rot_delta": 100, "object_move": -750}
        if p["weight"] is None:
            p["weight"] = defaults[name]
        p["scale"] = c.get("scale", 1.0)
        p["norm_p"] = c.get("norm_p", 2)
        out[name] = p
    return out


# ----------------------------------------------------------------------------
# the environment-level restatement
# ----------------------------------------------------------------------------
class OracleSim:
    """Holds the simulator-owned tensors and plays a StateSequence into them in place
    (the role FakeGym plays for the reference)."""

    def __init__(self, seq, num_envs: int):
        self.seq, self.N = seq, num_envs
        self.cursor = 0
        self.frame = 0
        self.root = torch.zeros(4 * num_envs, 13)
        self.dof = torch.zeros(num_envs, 9, 2)
        self.rb = torch.zeros(num_envs, 20, 13)
        self.dof_force = torch.zeros(num_envs, 9)
        self.ft = torch.zeros(num_envs, 18)
        self._load(0)

    def _load(self, t):
        s = self.seq
        self.root.copy_(s.root_state[t])
        self.dof.copy_(s.dof_state[t])
        self.rb.copy_(s.rigid_body[t])
        self.dof_force.copy_(s.dof_force[t])
        self.ft.copy_(s.ft_sensors[t])

    def simulate(self):
        self._load(self.cursor % self.seq.num_steps)
        self.cursor += 1
        self.frame += 1


class OracleEnv:
    """Restates IsaacEnvBase.step/reset (ref envs/env_base.py:322-401) and the TrifingerEnv
    hooks (ref envs/trifinger/trifinger_env.py:373-559, :959-1265) over an OracleSim."""

    fingertip_bodies = (6, 11, 16)

    def __init__(self, config: dict, sim: OracleSim, seed_torch: bool = True):
        self.cfg = config
        self.sim = sim
        N = self.N = config["num_instances"]
        self.episode_length = config["episode_length"]
        self.decimation = config["control_decimation"]
        self.asym = bool(config["asymmetric_obs"])
        mode = config["command_mode"]
        if mode not in ("position", "torque", "position_impedance"):
            raise ValueError(f"Invalid command mode. Input: {mode} not in ['torque', 'position'].")
        self.A = 18 if mode == "position_impedance" else 9
        self.obs_dim = 9 + 9 + 7 + 7 + self.A
        self.state_dim = self.obs_dim + 6 + 39 + 9 + 18 if self.asym else 0
        # buffers (ref env_base.py:560-572, trifinger_env.py:336, :588-590)
        self.states_buf = torch.zeros((N, self.state_dim), dtype=F32)
        self.obs_buf = torch.zeros((N, self.obs_dim), dtype=F32)
        self.action_buf = torch.zeros((N, self.A), dtype=F32)
        self.reset_buf = torch.zeros(N, dtype=torch.bool)
        self.goal_reset_buf = torch.zeros(N, dtype=torch.bool)
        self.reward_buf = torch.zeros(N, dtype=F32)
        self.steps_count_buf = torch.zeros(N, dtype=torch.long)
        self.successes = torch.zeros(N, dtype=torch.bool)
        self.goal_poses = torch.zeros((N, 7), dtype=F32)
        self.goal_movement = torch.zeros((N, 6), dtype=F32)
        # actor indices: robot, stage, object, goal per env (ref trifinger_env.py:811-849)
        e = torch.arange(N, dtype=torch.long)
        self.idx_robot, self.idx_object, self.idx_goal = 4 * e, 4 * e + 2, 4 * e + 3
        # views of simulator memory (ref :611-617)
        self.dof_pos = sim.dof[..., 0]
        self.dof_vel = sim.dof[..., 1]
        # 2-deep histories seeded with the initial state (ref :619-628)
        tips = list(self.fingertip_bodies)
        self.tip_hist = [sim.rb[:, tips], sim.rb[:, tips]]
        self.obj_hist = [sim.root[self.idx_object], sim.root[self.idx_object]]
        self._build_scales(mode)
        self.terms = reward_term_params(config["reward_terms"])
        self.draw = Draws()
        self.step_info: Dict[str, object] = {}
        self.last_terms: Optional[torch.Tensor] = None
        self.applied_torque: Optional[torch.Tensor] = None
        self.index_lists: Dict[str, torch.Tensor] = {}
        self._pending_draws = {"reset": None, "goal": None}
        if seed_torch:
            torch.manual_seed(config["seed"])  # ref env_base.py:169, :312-320

    # -- scales (ref trifinger_env.py:630-710) -----------------------------------
    def _build_scales(self, mode):
        jl, jh = _t(JOINT_LOW), _t(JOINT_HIGH)
        tq_l, tq_h = torch.full((9,), -MAX_TORQUE, dtype=F32), torch.full((9,), MAX_TORQUE, dtype=F32)
        if mode == "position":
            self.act_lo, self.act_hi = jl, jh
        elif mode == "torque":
            self.act_lo, self.act_hi = tq_l, tq_h
        else:
            self.act_lo = torch.cat([jl, _t(STIFFNESS_LOW)])
            self.act_hi = torch.cat([jh, _t(STIFFNESS_HIGH)])
        if self.cfg["normalize_action"]:
            oa_lo, oa_hi = torch.full((self.A,), -1, dtype=F32), torch.full((self.A,), 1, dtype=F32)
        else:
            oa_lo, oa_hi = self.act_lo, self.act_hi
        vel_l, vel_h = torch.full((9,), -MAX_VELOCITY, dtype=F32), torch.full((9,), MAX_VELOCITY, dtype=F32)
        pos_l, pos_h = _t([-0.3, -0.3, 0]), _t([0.3, 0.3, 0.3])
        q_l, q_h = -torch.ones(4), torch.ones(4)
        self.obs_lo = torch.cat([jl, vel_l, pos_l, q_l, pos_l, q_l, oa_lo])
        self.obs_hi = torch.cat([jh, vel_h, pos_h, q_h, pos_h, q_h, oa_hi])
        if self.asym:
            tip_l = torch.cat([_t([-0.4, -0.4, 0]), q_l, torch.full((6,), -0.2, dtype=F32)])
            tip_h = torch.cat([_t([0.4, 0.4, 0.5]), q_h, torch.full((6,), 0.2, dtype=F32)])
            ov_l, ov_h = torch.full((6,), -0.5, dtype=F32), torch.full((6,), 0.5, dtype=F32)
            w_l, w_h = torch.full((6,), -1.0, dtype=F32), torch.full((6,), 1.0, dtype=F32)
            self.st_lo = torch.cat([self.obs_lo, ov_l, tip_l.repeat(3), tq_l, w_l.repeat(3)])
            self.st_hi = torch.cat([self.obs_hi, ov_h, tip_h.repeat(3), tq_h, w_h.repeat(3)])
        self.tq_lo, self.tq_hi = tq_l, tq_h

    @property
    def env_steps_count(self) -> int:
        """ref env_base.py:286-289"""
        return self.sim.frame * self.N

    # -- IsaacEnvBase.reset / step (ref env_base.py:322-401) ------------------------
    def reset(self):
        self.reset_impl(torch.arange(0, self.N))
        self.pre_step()
        self.sim.simulate()
        self.fill_observations_and_states()
        return self.obs_buf.clone().detach()

    def step(self, action):
        self.step_info = {}
        if isinstance(action, np.ndarray):
            action = torch.tensor(action, dtype=F32)
        if tuple(action.size()) != (self.N, self.A):
            raise ValueError(f"Invalid shape for tensor `action`. Input: {tuple(action.size())} != {(self.N, self.A)}.")
        self.action_buf = action.clone()
        env_ids = torch.nonzero(self.reset_buf).view(-1)
        if len(env_ids) > 0:
            self.reset_impl(env_ids)
        goal_env_ids = torch.nonzero(self.goal_reset_buf).view(-1)
        if len(goal_env_ids) > 0:
            self.goal_reset_impl(goal_env_ids)
        self.last_ids = (env_ids, goal_env_ids)
        self.pre_step()
        for _ in range(self.decimation):
            self.sim.simulate()
        self.post_step()
        self.steps_count_buf += 1
        if self.episode_length is not None:
            timeout = torch.greater_equal(self.steps_count_buf, self.episode_length)
            self.reset_buf = torch.logical_or(self.reset_buf, timeout)
        dones = torch.logical_and(self.reset_buf, self.goal_reset_buf)
        return self.obs_buf, self.reward_buf, dones, self.step_info

    # -- reset hooks (ref trifinger_env.py:373-440) ---------------------------------
    def inject_draws(self, reset=None, goal=None):
        """Queue (uniforms, normals) for the next `reset_impl` / `goal_reset_impl`."""
        self._pending_draws = {"reset": reset, "goal": goal}

    def reset_impl(self, ids, draws=None):
        draws = draws if draws is not None else self._pending_draws["reset"]
        self._pending_draws["reset"] = None
        if draws is not None:
            self.draw.inject(*draws)
        self.reset_buf[ids] = 0
        self.steps_count_buf[ids] = 0
        self.successes[ids] = 0
        self.action_buf[ids] = 0
        rd = self.cfg["reset_distribution"]
        r = rd["robot_initial_state"]
        self._sample_robot_state(ids, r["type"], r["dof_pos_stddev"], r["dof_vel_stddev"])
        self._sample_object_poses(ids, rd["object_initial_state"]["type"])
        self._sample_goal_poses(ids, self.cfg["task_difficulty"])
        robot = self.idx_robot[ids].to(torch.int32)
        obj = self.idx_object[ids].to(torch.int32)
        goal = self.idx_goal[ids].to(torch.int32)
        self.index_lists = {"dof": robot, "root_reset": torch.unique(torch.cat([robot, obj, goal]))}
        self.draw.clear()

    def goal_reset_impl(self, ids, draws=None):
        draws = draws if draws is not None else self._pending_draws["goal"]
        self._pending_draws["goal"] = None
        if draws is not None:
            self.draw.inject(*draws)
        self.goal_reset_buf[ids] = 0
        self._sample_goal_poses(ids, self.cfg["task_difficulty"])
        goal = self.idx_goal[ids].to(torch.int32)
        self.index_lists["root_goal"] = torch.unique(torch.cat([goal]))
        self.draw.clear()

    def _sample_robot_state(self, ids, kind, pos_std, vel_std):
        """ref trifinger_env.py:1101-1147"""
        k = ids.size()[0]
        if kind == "none":
            return
        if kind == "default":
            self.dof_pos[ids] = _t(JOINT_DEFAULT)
            self.dof_vel[ids] = torch.zeros(9)
        elif kind == "random":
            noise = 2 * self.draw.uniform(k, tuple(range(18))) - 1
            self.dof_pos[ids] = _t(JOINT_DEFAULT)
            self.dof_vel[ids] = torch.zeros(9)
            self.dof_pos[ids] += pos_std * noise[:, 0:9]
            self.dof_vel[ids] += vel_std * noise[:, 9:18]
        else:
            raise ValueError(f"Invalid robot initial state distribution. Input: {kind} not in [`default`, `random`].")
        self.tip_hist[1][ids] = 0.0

    def _sample_object_poses(self, ids, kind):
        """ref trifinger_env.py:1149-1192"""
        k = ids.size()[0]
        if kind == "none":
            return
        if kind == "default":
            px, py, pz = _t([0, 0, CUBE_MIN_HEIGHT])
            quat = _t([0.0, 0.0, 0.0, 1.0])
        elif kind == "random":
            px, py = random_xy(self.draw, k, MAX_COM_DISTANCE, (18, 19))
            pz = CUBE_SIZE / 2
            quat = random_yaw_orientation(self.draw, k, 20)
        else:
            raise ValueError(f"Invalid object initial state distribution. Input: {kind} "
                             "not in [`default`, `random`, `none`].")
        h0 = self.obj_hist[0]
        h0[ids, 0] = px
        h0[ids, 1] = py
        h0[ids, 2] = pz
        h0[ids, 3:7] = quat
        h0[ids, 7:13] = 0
        self.obj_hist[1][ids] = 0.0
        self.sim.root[self.idx_object[ids]] = h0[ids]

    def _sample_goal_poses(self, ids, difficulty):
        """ref trifinger_env.py:1194-1265; draw order SURVEY.md §A.5"""
        k = ids.size()[0]
        d = self.draw
        if difficulty == -1:
            px, py = random_xy(d, k, MAX_COM_DISTANCE, (21, 22))
            pz = CUBE_SIZE / 2
            quat = random_yaw_orientation(d, k, 23)
        elif difficulty == 1:
            px, py = random_xy(d, k, MAX_COM_DISTANCE, (21, 22))
            pz = CUBE_SIZE / 2
            quat = default_orientation(k)
        elif difficulty == 2:
            px, py = 0.0, 0.0
            pz = CUBE_MIN_HEIGHT + 0.05
            quat = default_orientation(k)
        elif difficulty == 3:
            px, py = random_xy(d, k, MAX_COM_DISTANCE, (21, 22))
            pz = random_z(d, k, CUBE_MIN_HEIGHT, CUBE_MAX_HEIGHT, 23)
            quat = default_orientation(k)
        elif difficulty in (4, 5):
            px, py = random_xy(d, k, MAX_COM_DISTANCE, (21, 22))
            pz = random_z(d, k, CUBE_RADIUS_3D, CUBE_MAX_HEIGHT, 23)
            quat = random_orientation(d, k)
        elif difficulty == 6:
            px, py = 0.0, 0.0
            pz = CUBE_MIN_HEIGHT + 0.05
            quat = random_orientation(d, k)
        else:
            raise ValueError(f"Invalid difficulty index for task: {difficulty}.")
        rot = self.cfg["goal_movement"]["rotation"]
        if rot["activate"]:
            self.goal_movement[ids, 3:6] = random_angular_vel(d, k, rot["rate_magnitude"])
        else:
            self.goal_movement[ids, 3:6] = 0.0
        self.goal_poses[ids, 0] = px
        self.goal_poses[ids, 1] = py
        self.goal_poses[ids, 2] = pz
        self.goal_poses[ids, 3:7] = quat
        rows = self.idx_goal[ids]
        self.sim.root[rows, 0:7] = self.goal_poses[ids]
        self.sim.root[rows, 7:13] = self.goal_movement[ids]

    # -- action -> torque (ref trifinger_env.py:442-498) -----------------------------
    def pre_step(self):
        cfg = self.cfg
        if cfg["normalize_action"]:
            act = unscale_transform(self.action_buf, self.act_lo, self.act_hi)
        else:
            act = self.action_buf
        mode = cfg["command_mode"]
        if mode == "torque":
            torque = act
        elif mode == "position":
            torque = _t(KP) * (act - self.dof_pos)
            torque -= _t(KD) * self.dof_vel
        elif mode == "position_impedance":
            torque = act[:, 9:18] * (act[:, 0:9] - self.dof_pos)
            torque -= _t(KD) * self.dof_vel
        else:
            raise ValueError(f"Invalid command mode. Input: {mode} not in ['torque', 'position'].")
        applied = saturate(torque, self.tq_lo, self.tq_hi)
        if cfg["apply_safety_damping"]:
            applied -= _t(SAFETY_KD) * self.dof_vel
            applied = saturate(applied, self.tq_lo, self.tq_hi)
        self.applied_torque = applied
        if cfg["goal_movement"]["rotation"]["activate"]:  # ref :1267-1277
            self.sim.root[self.idx_goal, 10:13] = self.goal_movement[:, 3:6]
            self.index_lists["root_move"] = self.idx_goal.to(torch.int32)

    # -- observations / states (ref trifinger_env.py:959-1051) -----------------------
    def fill_observations_and_states(self):
        tips = list(self.fingertip_bodies)
        self.tip_hist = [self.sim.rb[:, tips], self.tip_hist[0]]
        self.obj_hist = [self.sim.root[self.idx_object], self.obj_hist[0]]
        o = self.obs_buf
        o[:, 0:9] = self.dof_pos
        o[:, 9:18] = self.dof_vel
        o[:, 18:25] = self.obj_hist[0][:, 0:7]
        o[:, 25:32] = self.goal_poses
        o[:, 32:32 + self.A] = self.action_buf
        if self.asym:
            s = self.states_buf
            n0 = self.obs_dim
            s[:, 0:n0] = o
            s[:, n0:n0 + 6] = self.obj_hist[0][:, 7:13]
            s[:, n0 + 6:n0 + 45] = self.tip_hist[0].reshape(self.N, 39)
            s[:, n0 + 45:n0 + 54] = self.sim.dof_force
            s[:, n0 + 54:n0 + 72] = self.sim.ft
        if self.cfg["normalize_obs"]:
            self.obs_buf = scale_transform(self.obs_buf, self.obs_lo, self.obs_hi)
            if self.asym:
                self.states_buf = scale_transform(self.states_buf, self.st_lo, self.st_hi)

    # -- reward + termination (ref trifinger_env.py:500-559, :1053-1099) -------------
    def post_step(self):
        self.fill_observations_and_states()
        self.reward_buf[:] = 0
        dt = self.cfg["sim"]["dt"]
        T = self.env_steps_count
        tp = self.terms
        t0, t1, o0, o1, g = self.tip_hist[0], self.tip_hist[1], self.obj_hist[0], self.obj_hist[1], self.goal_poses
        vals = OrderedDict([
            ("finger_reach_object_rate", finger_reach_object_rate(tp["finger_reach_object_rate"], T, t0, t1, o0, o1)),
            ("finger_move_penalty", finger_move_penalty(tp["finger_move_penalty"], dt, t0, t1)),
            ("object_dist", object_dist(tp["object_dist"], dt, T, o0, g)),
            ("object_rot", object_rot(tp["object_rot"], dt, T, o0, g)),
            ("object_rot_delta", object_rot_delta(tp["object_rot_delta"], dt, T, o0, o1, g)),
            ("object_move", object_move(tp["object_move"], o0, o1, g)),
        ])
        for name, v in vals.items():
            if tp[name]["activate"]:
                self.reward_buf += v
                self.step_info[f"env/rewards/{name}"] = v.mean()
        self.last_terms = torch.stack(list(vals.values()), dim=0)
        self._check_termination()
        if self.cfg["goal_movement"]["rotation"]["activate"]:  # ref :1279-1284
            self.goal_poses[:] = self.sim.root[self.idx_goal, 0:7]

    def _check_termination(self):
        succ = self.cfg["termination_conditions"]["success"]
        o0, g = self.obj_hist[0], self.goal_poses
        dist = torch.norm(g[:, 0:3] - o0[:, 0:3], p=2, dim=-1)
        pos_ok = torch.le(dist, succ["position_tolerance"])
        self.step_info["env/current_position_goal/count"] = torch.sum(pos_ok)
        ang = quat_diff_rad(o0[:, 3:7], g[:, 3:7])
        rot_ok = torch.le(ang, succ["orientation_tolerance"])
        self.step_info["env/current_orientation_goal/count"] = torch.sum(rot_ok)
        d = self.cfg["task_difficulty"]
        if d < 4:
            done = pos_ok
        elif d == 4:
            done = torch.logical_and(pos_ok, rot_ok)
        else:
            done = rot_ok
        if succ["activate"]:
            ids = torch.nonzero(done).squeeze()
            self.reward_buf[ids] += succ["bonus"]
            self.goal_reset_buf = done
            self.successes = self.successes + self.goal_reset_buf
        else:
            self.successes = torch.logical_and(self.goal_reset_buf, self.successes)
        self.step_info["env/average_consecutive_success"] = np.mean(self.successes.cpu().numpy())


def clip_for_learner(x, limit: float):
    """VecTaskPython's clamp (ref wrappers/vec_task.py:146-170)."""
    return torch.clamp(x, -limit, limit)
