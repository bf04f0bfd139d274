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


def test_each_concurrent_probe_matches_its_own_oracle_rollout():
    """The probes differ in their material parameters (train_material_params.py:583-610: D -> density, E -> stiffness,
    H -> rest directions): stepped concurrently, each must reproduce the CPU oracle run with ITS parameters
    (1e-4 relative on x, v after N = 10 substeps, BASELINE.json's tolerance)."""
    from mpmavatar_b200 import synthetic as S
    from mpmavatar_b200.probes import ProbeBatch
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    from oracle.oracle import OracleSim
    base = S.scene_small_cloth_body()
    scenes = [base, dataclasses.replace(base, E=base.E * 1.3), dataclasses.replace(base, density=base.density * 0.7),
              dataclasses.replace(base, R_inv=(base.R_inv * np.float32(1.02)).astype(np.float32))]
    nsub = 10
    ft = frame_tensors(base, 0)
    batch = ProbeBatch([build_from_scene(sc) for sc in scenes])
    batch.step(base.dt, nsub, ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
    batch.sync()
    fi = base.frame_inputs(0)
    finals = []
    for sc, (solver, model, state) in zip(scenes, batch.triples):
        o = OracleSim.from_scene(sc, "f32", threads=4)
        for k in range(nsub):
            o.p2g2p(sc.dt, fi["mesh_x"] + np.float32(sc.dt * k) * fi["mesh_v"], fi["mesh_v"], None, fi["joint_verts_v"], fi["joint_faces_v"])
        x, v = state.particle_x.cpu().numpy(), state.particle_v.cpu().numpy()
        assert np.abs(x - o.x).max() / np.abs(o.x).max() < 1e-4
        assert np.abs(v - o.v).max() / np.abs(o.v).max() < 1e-4
        finals.append(v)
    for other in finals[1:]:
        assert np.abs(other - finals[0]).max() > 1e-6  # the parameter changes are visible in the result
