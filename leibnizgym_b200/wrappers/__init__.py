from .rlg_adapter import RlGamesGpuEnvAdapter
from .vec_task import VecTask, VecTaskPython

__all__ = ["VecTask", "VecTaskPython", "RlGamesGpuEnvAdapter"]
