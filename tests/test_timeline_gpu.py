"""The timeline probe (mpm_measure_timeline): stamps are ordered along the substep chain, and the probed substeps
are ordinary substeps -- the state after them equals the state after the same number of p2g2p calls."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_timeline_stamps_and_state():
    from mpmavatar_b200 import synthetic as S
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    from mpmavatar_b200.timeline import measure, summarise
    sc = S.scene_small_cloth_body()
    ft = frame_tensors(sc, 0)
    args = dict(mesh_x=ft["mesh_x"], mesh_v=ft["mesh_v"], joint_verts_v=ft["joint_verts_v"], joint_faces_v=ft["joint_faces_v"])
    n = 8
    a_solver, a_model, a_state = build_from_scene(sc)
    a_solver.p2g2p(a_model, a_state, sc.dt, **args)
    tl = measure(a_solver, sc.dt, ft, n)  # the captured graph of n substeps is replayed twice: 2n substeps, body at rest
    b_solver, b_model, b_state = build_from_scene(sc)
    for _ in range(1 + 2 * n):
        b_solver.p2g2p(b_model, b_state, sc.dt, **args)
    xa, xb = a_state.particle_x.cpu().numpy(), b_state.particle_x.cpu().numpy()
    va, vb = a_state.particle_v.cpu().numpy(), b_state.particle_v.cpu().numpy()
    assert np.abs(xa - xb).max() / np.abs(xb).max() < 1e-5
    assert np.abs(va - vb).max() / max(np.abs(vb).max(), 1e-6) < 1e-3  # float atomics order
    # kernels that ran: element P2G, vertex P2G, scatter, grid update, both G2P; no traditional particles here
    ran = tl[0, :, 0] >= 0
    assert list(ran) == [True, False, True, True, True, True, False, True]
    assert (tl[:, ran, 1] >= tl[:, ran, 0]).all()
    # along the chain: element P2G starts first, the grid update ends before the vertex G2P ends, element G2P ends last
    for i in range(n):
        assert tl[i, 0, 0] == tl[i, ran, 0].min()
        assert tl[i, 4, 1] <= tl[i, 5, 1] <= tl[i, 7, 1] == tl[i, ran, 1].max()
        if i:
            assert tl[i, 0, 0] >= tl[i - 1, 5, 1]  # the next substep's P2G cannot start before the vertex G2P is done
    s = summarise(measure(a_solver, sc.dt, ft, 16))
    assert 0 < s["p2g_union_us"] + s["g2p_union_us"] <= 1.05 * s["substep_us"] + 1.0
    assert torch.isfinite(a_state.particle_x).all()
