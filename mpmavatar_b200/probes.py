"""Concurrent finite-difference probes on ONE GPU.

train_material_params.py estimates the gradient of its loss by forward differences: every optimiser step
runs four rollouts of the same garment on the same body motion, differing only in the material parameters
(base, +dD, +dE, +dH; train_material_params.py:583-660).  The reference runs them one after the other.  A
single 500 k-particle rollout keeps a B200 about one third busy (its substep is a chain of dependent,
latency-bound kernels), so the probes are independent work that fits in the gaps: each probe
gets its own solver handle and its own CUDA stream, the captured substep graphs of the K handles are
launched back to back from the one host thread and execute concurrently.  No reference counterpart."""
from __future__ import annotations

import torch


class ProbeBatch:
    """K independent (solver, model, state) triples stepped concurrently on K streams of one device."""

    def __init__(self, triples, device="cuda:0"):
        self.triples = list(triples)
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.triples]

    def __len__(self):
        return len(self.triples)

    def step(self, dt, nsub, mesh_x=None, mesh_v=None, joint_traditional_v=None, joint_verts_v=None, joint_faces_v=None):
        """nsub substeps of every probe (same body motion, as the caller's probes have).  Returns at once; the
        probes run concurrently, `sync()` or a read of a state tensor on the current stream joins them."""
        cur = torch.cuda.current_stream(self.device)
        for (solver, model, state), st in zip(self.triples, self.streams):
            st.wait_stream(cur)  # inputs produced on the caller's stream
            with torch.cuda.stream(st):
                solver.step(model, state, dt, nsub, mesh_x, mesh_v, joint_traditional_v, joint_verts_v, joint_faces_v)
        for st in self.streams:
            cur.wait_stream(st)

    def positions(self):
        return [state.particle_x for _, _, state in self.triples]

    def sync(self):
        for st in self.streams:
            st.synchronize()
