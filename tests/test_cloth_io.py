"""Caller-side glue on the device (mpmavatar_b200/cloth_io.py, csrc/mpm_mesh.cuh) against
golden vectors produced by executing the reference's own source text (tests/golden/make_mesh_golden.py ->
tests/golden/cloth_particles.npz): compute_dir_vol / compute_rest_dir_inv(_from_vf) / wld2sim, the export
(un-permute + sim2wld + scatter + MSE), the OBJ writer, the split_idx.npz schema and MeshGaussianModel.set_mesh_by_verts.
fp32 work: 2e-6 relative (a few ulps: the reference is torch fp32 with a different summation order inside norm())."""
import os

import numpy as np
import pytest
import torch

from oracle import mesh_oracle as MO

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cloth_particles.npz"))
TOL = 2e-6


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ------------------------------------------------------------------ CPU: the oracle restatement is pinned to the reference
def test_oracle_matches_reference_source():
    scale, shift = MO.cloth_normalisation(G["verts_wld"])
    assert rel(scale, G["scale"]) < TOL and rel(shift, G["shift"]) < TOL
    init_dir, rest_dir, evol, vvol = MO.compute_dir_vol(G["verts_sim"], G["faces"], float(G["thickness"]))
    assert rel(init_dir, G["init_dir"]) < TOL and rel(rest_dir, G["rest_dir"]) < TOL
    assert rel(evol, G["element_vol"]) < TOL and rel(vvol, G["vertex_vol"]) < TOL
    assert rel(MO.compute_rest_dir_inv(rest_dir), G["rest_dir_inv"]) < TOL
    assert rel(G["rest_dir_inv_from_vf"], G["rest_dir_inv"]) < TOL
    c, o, s = MO.face_frames(G["cloth_wld"], G["faces"])
    assert rel(c, G["face_center"]) < TOL and rel(o, G["face_orien"]) < 1e-5 and rel(s, G["face_scale"]) < 1e-5


def test_write_obj_round_trips_every_float32(tmp_path):
    from mpmavatar_b200 import cloth_io
    rng = np.random.default_rng(0)
    v = rng.normal(size=(500, 3)).astype(np.float32) * np.float32(3.0)
    v[0] = [1.0, -2.0, 0.0]
    v[1] = [1e-5, 123456.7, -3.4028235e38]
    v[2] = [np.float32(0.1), np.float32(1) / np.float32(3), 5e-39]  # incl. a subnormal
    tail = ["vt 0.5 0.25\n", "f 1/1 2/1 3/1\n"]
    p = tmp_path / "a.obj"
    cloth_io.write_obj(p, v, tail)
    lines = open(p).read().split("\n")
    assert lines[-3:] == ["vt 0.5 0.25", "f 1/1 2/1 3/1", ""]
    assert lines[0] == "v 1.0 -2.0 0.0"
    back = np.array([[np.float32(t) for t in ln.split()[1:]] for ln in lines[:500]], np.float32)
    assert back.tobytes() == v.tobytes()  # binary-safe: every value reads back bit-identically
    # the reference's own line format for the same numbers (f-string of numpy float32) parses to the same values
    ref_lines = [f"v {x[0]} {x[1]} {x[2]}" for x in v]
    for a, b in zip(lines[:500], ref_lines):
        assert [np.float32(t) for t in a.split()[1:]] == [np.float32(t) for t in b.split()[1:]]
    cloth_io.write_obj(tmp_path / "b.obj", torch.from_numpy(v))  # tensors, no tail
    assert open(tmp_path / "b.obj").read() == "\n".join(lines[:500]) + "\n"


def test_split_idx_schema(tmp_path):
    from mpmavatar_b200 import cloth_io
    d = dict(num_joint_v=3, num_joint_f=2, reordered_cloth_v_idx=np.arange(5), reordered_cloth_f_idx=np.arange(4),
             reordered_human_v_idx=np.arange(7), reordered_human_f_idx=np.arange(6), new_cloth_faces=np.zeros((4, 3), np.int32),
             new_human_faces=np.zeros((6, 3), np.int32))
    np.savez(tmp_path / "split_idx.npz", **d)  # the file preprocess/split_garments.py:84-94 writes
    s = cloth_io.load_split_idx(tmp_path / "split_idx.npz")
    assert s["num_joint_v"] == 3 and s["num_joint_f"] == 2 and s["new_cloth_faces"].shape == (4, 3)
    del d["new_human_faces"]
    np.savez(tmp_path / "bad.npz", **d)
    with pytest.raises(KeyError):
        cloth_io.load_split_idx(tmp_path / "bad.npz")


def test_no_cpu_path():
    from mpmavatar_b200 import cloth_io
    with pytest.raises(RuntimeError):
        cloth_io.compute_dir_vol(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int64), 1e-5)


# ------------------------------------------------------------------ GPU: the CUDA kernels against the same golden vectors
@pytest.mark.gpu
def test_build_cloth_particles_matches_reference_source():
    from mpmavatar_b200 import cloth_io
    T = lambda a: torch.as_tensor(a, device="cuda")
    b = cloth_io.build_cloth_particles(T(G["verts_wld"]), T(G["faces"]), float(G["thickness"]))
    Ne = G["faces"].shape[0]
    assert b["n_elements"] == Ne and b["n_vertices"] == G["verts_wld"].shape[0]
    assert rel(b["scale"], G["scale"]) < TOL and rel(b["shift"].cpu().numpy().reshape(-1), G["shift"].reshape(-1)) < TOL
    x = b["x"].cpu().numpy()
    assert rel(x[Ne:], G["verts_sim"]) < TOL and rel(x[:Ne], G["elts"]) < TOL
    # everything below is built from DIFFERENCES of the sim-space vertices (edges of ~0.08 between positions of ~1): one
    # ulp of a position -- the scale computed here vs. by torch -- is 1.5e-6 of an edge, twice that in an area
    vol = b["vol"].cpu().numpy()
    assert rel(vol[:Ne], G["element_vol"]) < 1e-5 and rel(vol[Ne:], G["vertex_vol"]) < 1e-5
    assert rel(b["init_dir"].cpu().numpy(), G["init_dir"]) < 1e-5
    assert rel(b["rest_dir"].cpu().numpy(), G["rest_dir"]) < 1e-5
    assert rel(b["rest_dir_inv"].cpu().numpy(), G["rest_dir_inv"]) < 1e-5


@pytest.mark.gpu
def test_reference_named_methods_on_device():
    from mpmavatar_b200 import cloth_io
    T = lambda a: torch.as_tensor(a, device="cuda")
    init_dir, rest_dir, evol, vvol = cloth_io.compute_dir_vol(T(G["verts_sim"]), T(G["faces"]), float(G["thickness"]))
    assert rel(init_dir.cpu().numpy(), G["init_dir"]) < TOL and rel(rest_dir.cpu().numpy(), G["rest_dir"]) < TOL
    assert rel(evol.cpu().numpy(), G["element_vol"]) < TOL and rel(vvol.cpu().numpy(), G["vertex_vol"]) < TOL
    assert rel(cloth_io.compute_rest_dir_inv(rest_dir).cpu().numpy(), G["rest_dir_inv"]) < TOL
    assert rel(cloth_io.compute_rest_dir_inv_from_vf(T(G["verts_sim"]), T(G["faces"])).cpu().numpy(), G["rest_dir_inv_from_vf"]) < TOL


def _solver_with(x_all, b):
    """A solver holding the golden garment with its vertices at x_all[Ne:] (the caller's set-up sequence)."""
    from mpmavatar_b200.warp_mpm.mpm_data_structure import MPMModelStruct, MPMStateStruct
    from mpmavatar_b200.warp_mpm.mpm_solver import MPMWARP
    import contextlib, io
    Ne, Nv = b["n_elements"], b["n_vertices"]
    N = Ne + Nv
    state, model = MPMStateStruct(), MPMModelStruct()
    state.init(N, Ne, Nv, device="cuda:0")
    elem = np.zeros(N, np.int32); elem[:Ne] = 1
    vert = np.zeros(N, np.int32); vert[Ne:] = 1
    with contextlib.redirect_stdout(io.StringIO()):
        state.from_torch(x_all, b["vol"], torch.linalg.inv(b["init_dir"]), b["rest_dir_inv"],
                         torch.as_tensor(G["faces"], device="cuda").float(), np.zeros(N, np.int32), vert, elem,
                         torch.zeros(Ne, 6), n_grid=48, grid_lim=2.0, device="cuda:0")
    model.init(N, device="cuda:0")
    model.init_other_params(n_grid=48, grid_lim=2.0, device="cuda:0")
    solver = MPMWARP(N, Ne, Nv, n_grid=48, grid_lim=2.0, device="cuda:0")
    solver.set_parameters_dict(model, state, {"material": "cloth", "g": [0.0, -9.8, 0.0], "density": 1.0, "friction_angle": 40.0})
    state.reset_state(Nv, x_all.clone(), b["init_dir"], None, torch.zeros_like(x_all), tensor_R_inv=b["rest_dir_inv"], device="cuda:0")
    state.reset_density(torch.ones(N, device="cuda"), None, "cuda:0", update_mass=True)
    one = torch.ones(N, device="cuda")
    solver.set_E_nu_from_torch(model, one * 100.0, one * 0.3, one * 500.0, one * 500.0, "cuda:0")
    solver.prepare_mu_lam(model, state, "cuda:0")
    solver.set_particles(model, state)
    return solver, model, state


@pytest.mark.gpu
def test_export_cloth_verts_unpermute_sim2wld_scatter_mse():
    from mpmavatar_b200 import cloth_io
    T = lambda a: torch.as_tensor(a, device="cuda")
    b = cloth_io.build_cloth_particles(T(G["verts_wld"]), T(G["faces"]), float(G["thickness"]))
    Ne, Nv = b["n_elements"], b["n_vertices"]
    x_all = b["x"].clone()
    x_all[Ne:] = T(G["moved_sim"])
    solver, model, state = _solver_with(x_all, b)
    rng = np.random.default_rng(5)
    scatter = rng.permutation(Nv + 40)[:Nv]  # reordered_cloth_v_idx into a larger full-body array
    full = torch.full((Nv + 40, 3), -7.0, device="cuda")
    out, mse = cloth_io.export_cloth_verts(solver, b["scale"], b["shift"], scatter_idx=scatter, full_verts=full, target=T(G["target"]))
    assert rel(out.cpu().numpy(), G["cloth_wld"]) < TOL  # un-permuted + sim2wld
    assert rel(float(mse), float(G["mse"])) < 1e-5
    fl = full.cpu().numpy()
    assert rel(fl[scatter], G["cloth_wld"]) < TOL
    rest = np.setdiff1d(np.arange(Nv + 40), scatter)
    assert (fl[rest] == -7.0).all()
    # after real substeps: identical to the caller's own route, wp.to_torch(state.particle_x)[Ne:] through sim2wld
    solver.step(model, state, 1e-4, 70)  # crosses a re-sort
    out2, _ = cloth_io.export_cloth_verts(solver, b["scale"], b["shift"])
    ref = (state.particle_x[Ne:] - b["shift"]) / b["scale"]
    assert rel(out2.cpu().numpy(), ref.cpu().numpy()) < TOL
    assert float((out2 - out).abs().max()) > 0


@pytest.mark.gpu
def test_face_frames_match_set_mesh_by_verts():
    from mpmavatar_b200 import cloth_io
    T = lambda a: torch.as_tensor(a, device="cuda")
    c, o, q, s = cloth_io.face_frames(T(G["cloth_wld"]), T(G["faces"]))
    assert rel(c.cpu().numpy(), G["face_center"]) < TOL
    assert rel(o.cpu().numpy(), G["face_orien"]) < 1e-5
    assert s.shape == (G["faces"].shape[0], 1) and rel(s.cpu().numpy(), G["face_scale"]) < 1e-5
    qn = q.cpu().numpy().astype(np.float64)
    assert np.abs(np.linalg.norm(qn, axis=1) - 1.0).max() < 1e-5
    assert rel(MO.quat_wxyz_to_rotmat(qn), G["face_orien"]) < 2e-5  # the unit quaternion (w,x,y,z) of that rotation, up to sign
