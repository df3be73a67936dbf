"""TEST INFRASTRUCTURE ONLY -- `ignite.utils.convert_tensor` (used by reference modified_ignite_engine.py:9-10)."""
import torch


def convert_tensor(x, device=None, non_blocking=False):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, non_blocking=non_blocking) if device is not None else x
    if isinstance(x, (list, tuple)):
        return type(x)(convert_tensor(v, device, non_blocking) for v in x)
    if isinstance(x, dict):
        return {k: convert_tensor(v, device, non_blocking) for k, v in x.items()}
    return x
