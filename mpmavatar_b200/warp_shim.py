"""Minimal stand-in for the `warp` module, used ONLY when NVIDIA Warp is not installed.

The callers of the hot path touch three Warp entry points outside warp_mpm itself:
wp.init() (train_material_params.py:399), wp.to_torch(state.particle_x) (:628, :811,
run_demo.py:532) and wp.synchronize().  With the B200 solver the state arrays ARE torch
tensors, so these become trivial."""
from __future__ import annotations

import torch


def init():
    return None


def to_torch(a, requires_grad=None):
    if isinstance(a, torch.Tensor):
        return a
    raise TypeError("warp shim: to_torch expects a torch.Tensor-backed array")


def from_torch(t, dtype=None, requires_grad=None, grad=None):
    return t


def synchronize():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def synchronize_device(device=None):
    synchronize()


class config:  # wp.config.mode / verify_cuda are only ever assigned (train_material_params.py:400-401)
    mode = "release"
    verify_cuda = False
