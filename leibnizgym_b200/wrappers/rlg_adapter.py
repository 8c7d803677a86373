"""rl_games-facing adapter (ref leibnizgym/utils/rlg_train.py:89-154).

The reference subclasses `rl_games.common.vecenv.IVecEnv`; rl_games is not a dependency of this
package, so the adapter is a plain class with the same methods, return shapes and quirks — an
rl_games user registers it exactly as the reference does (see `register`).  What a learner gets:

  * `reset()` -> `{"obs": ..., "states": ...}` when the env has critic states, else the obs tensor
    (ref :127-135);
  * `step(a)` -> `(full_state | obs, reward, is_done, [[], info])` (ref :137-148) — `info` rides in
    slot 1 of a two-element list so that `LeibnizAlgoObserver.process_infos` logs it directly
    (ref :176-194);
  * the SAME `full_state` dict object is returned every call and its entries are the env's
    persistent clipped output buffers: nothing is cloned or re-clamped between the fused kernels and
    the learner.
"""
from __future__ import annotations

from typing import Callable, Optional

from .vec_task import VecTaskPython


class RlGamesGpuEnvAdapter:
    def __init__(self, config_name: str = "rlgpu", num_actors: int = 0, *, env: Optional[VecTaskPython] = None,
                 env_creator: Optional[Callable[..., VecTaskPython]] = None, **kwargs):
        """Either pass the wrapped env, or an `env_creator(**kwargs)` (what rl_games' `env_configurations`
        registry holds under `config_name`, ref :96)."""
        if env is None:
            if env_creator is None:
                raise ValueError("RlGamesGpuEnvAdapter needs `env` or `env_creator`")
            env = env_creator(**kwargs)
        self.env = env
        self.config_name, self.num_actors = config_name, num_actors
        self.use_global_obs = self.env.num_states > 0          # asymmetric PPO (ref :98)
        self.full_state = {"obs": self.env.reset()}            # ref :100-102
        if self.use_global_obs:
            self.full_state["states"] = self.env.get_state()   # ref :104-105

    # -- properties (ref :111-129) -------------------------------------------------------------
    def get_number_of_agents(self) -> int:
        return self.env.get_number_of_agents()

    def get_env_info(self) -> dict:
        info = {"num_envs": self.env.num_envs, "action_space": self.env.action_space,
                "observation_space": self.env.observation_space}
        if self.use_global_obs:
            info["state_space"] = self.env.state_space
        return info

    # -- operations (ref :135-158) --------------------------------------------------------------
    def reset(self):
        self.full_state["obs"] = self.env.reset()
        if self.use_global_obs:
            self.full_state["states"] = self.env.get_state()
            return self.full_state
        return self.full_state["obs"]

    def step(self, action):
        next_obs, reward, is_done, info = self.env.step(action)
        self.full_state["obs"] = next_obs
        if self.use_global_obs:
            self.full_state["states"] = self.env.get_state()
            return self.full_state, reward, is_done, [[], info]
        return self.full_state["obs"], reward, is_done, [[], info]


def register(env_creator: Callable[..., VecTaskPython], vecenv_type: str = "RLGPU", config_name: str = "rlgpu") -> None:
    """Register with rl_games the way the reference does (ref :151-156).  Imports rl_games lazily and
    raises ImportError when it is absent (it is not part of this image)."""
    from rl_games.common import env_configurations, vecenv  # type: ignore

    class _Adapter(RlGamesGpuEnvAdapter, vecenv.IVecEnv):
        def __init__(self, cfg_name, num_actors, **kwargs):
            creator = env_configurations.configurations[cfg_name]["env_creator"]
            RlGamesGpuEnvAdapter.__init__(self, cfg_name, num_actors, env_creator=creator, **kwargs)

    vecenv.register(vecenv_type, lambda cfg_name, num_actors, **kw: _Adapter(cfg_name, num_actors, **kw))
    env_configurations.register(config_name, {"vecenv_type": vecenv_type, "env_creator": env_creator})
