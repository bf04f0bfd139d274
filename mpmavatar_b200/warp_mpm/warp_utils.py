"""warp_mpm/warp_utils.py mirror: from_torch_safe wrapped a tensor as a wp.array without a
copy (warp_utils.py:12-89).  The B200 solver holds torch tensors directly, so it only checks
dtype and returns the (contiguous, detached) tensor."""
import torch


def from_torch_safe(t: torch.Tensor, dtype=None, requires_grad=None, grad=None) -> torch.Tensor:
    if t.dtype not in (torch.float32, torch.int32):
        raise RuntimeError(f"Incompatible data types: {t.dtype}")
    return t.contiguous().detach()
