"""Concurrent finite-difference probes (mpmavatar_b200/probes.py): K rollouts stepped on K streams give what the
same rollouts give one after the other."""
import dataclasses

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_concurrent_probes_equal_sequential_rollouts():
    from mpmavatar_b200 import synthetic as S
    from mpmavatar_b200.probes import ProbeBatch
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    base = S.scene_small_cloth_body()
    # the caller's probes: base, stiffer, denser (train_material_params.py:652-660 perturbs D, E, H)
    scenes = [base, dataclasses.replace(base, E=base.E * 1.05), dataclasses.replace(base, density=base.density * 1.05)]
    ft = frame_tensors(base, 0)
    args = (ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
    seq = []
    for sc in scenes:
        solver, model, state = build_from_scene(sc)
        solver.step(model, state, sc.dt, 48, *args)
        seq.append(state.particle_x.cpu().numpy())
    batch = ProbeBatch([build_from_scene(sc) for sc in scenes])
    batch.step(base.dt, 48, *args)
    batch.sync()
    for x_seq, x_con in zip(seq, batch.positions()):
        x_con = x_con.cpu().numpy()
        assert np.isfinite(x_con).all()
        assert np.abs(x_con - x_seq).max() / np.abs(x_seq).max() < 1e-5  # float atomics order only
    assert np.abs(seq[0] - seq[1]).max() > 0  # the probes do differ
