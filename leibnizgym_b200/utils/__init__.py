from . import torch_utils

__all__ = ["torch_utils"]
