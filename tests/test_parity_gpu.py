"""GPU parity: the CUDA path (through the reference-facing API and the C-ABI) against the CPU
oracle on identical seeded inputs.  Tolerance (BASELINE.json north_star): 1e-4 relative on
positions / velocities after N substeps; 1e-3 of max|.| on C, d, F_trial, stress."""
import numpy as np
import pytest
import torch

from mpmavatar_b200 import synthetic as S

pytestmark = pytest.mark.gpu

TOL_XV = 1e-4
TOL_AUX = 1e-3


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_oracle(sc, nsub, precision="f32", threads=8, joint_t=None):
    from oracle.oracle import OracleSim
    o = OracleSim.from_scene(sc, precision, threads=threads)
    fi = sc.frame_inputs(0)
    for k in range(nsub):
        mx = None if fi["mesh_x"] is None else fi["mesh_x"] + np.float32(sc.dt * k) * fi["mesh_v"]
        o.p2g2p(sc.dt, mx, fi["mesh_v"], joint_t, fi["joint_verts_v"], fi["joint_faces_v"])
    return o


def run_cuda(sc, nsub, per_call=True, debug=False, joint_t=None, resort_interval=0):
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    solver, model, state = build_from_scene(sc, resort_interval=resort_interval)
    if debug:
        solver.set_debug(True)
    ft = frame_tensors(sc, 0)
    jt = None if joint_t is None else torch.as_tensor(joint_t, dtype=torch.float32, device="cuda")
    if per_call:  # exactly the caller's loop (train_material_params.py:622-626)
        for k in range(nsub):
            mx = None if ft["mesh_x"] is None else ft["mesh_x"] + sc.dt * k * ft["mesh_v"]
            solver.p2g2p(model, state, sc.dt, mesh_x=mx, mesh_v=ft["mesh_v"], joint_traditional_v=jt,
                         joint_verts_v=ft["joint_verts_v"], joint_faces_v=ft["joint_faces_v"])
    else:
        solver.step(model, state, sc.dt, nsub, ft["mesh_x"], ft["mesh_v"], jt, ft["joint_verts_v"], ft["joint_faces_v"])
    st = solver.stats()
    assert st["overflow"] == 0, st
    return solver, model, state


def compare(o, state, sc, tol_xv=TOL_XV, tol_aux=TOL_AUX):
    x = state.particle_x.cpu().numpy()
    v = state.particle_v.cpu().numpy()
    assert np.isfinite(x).all() and np.isfinite(v).all()
    assert rel(x, o.x) < tol_xv, ("x", rel(x, o.x))
    assert rel(v, o.v) < tol_xv, ("v", rel(v, o.v))
    # C = 4/dx * sum w v (x_i - x_p): an allowed velocity error eps_v maps to 4*inv_dx*eps_v in C, which is
    # the floor when the true affine field is ~0 (cloth at rest: C is pure round-off)
    Cc = state.particle_C.cpu().numpy()
    inv_dx = sc.n_grid / sc.grid_lim
    c_tol = tol_aux * np.abs(o.C).max() + tol_xv * np.abs(o.v).max() * 4.0 * inv_dx
    assert np.abs(Cc - o.C).max() < c_tol, ("C", np.abs(Cc - o.C).max(), c_tol)
    if sc.n_elements:
        assert rel(state.particle_d.cpu().numpy(), o.d) < tol_aux
        # cloth at rest has stress ~ round-off of mu*vol: compare against that scale, not against noise
        se = state.particle_stress.cpu().numpy()[: sc.n_elements]
        floor = float((o.mu[: sc.n_elements] * o.vol[: sc.n_elements]).max())  # stress at 100% strain
        assert np.abs(se - o.stress[: sc.n_elements]).max() < tol_aux * max(np.abs(o.stress[: sc.n_elements]).max(), floor)
    if sc.n_traditional:
        sl = slice(sc.n_elements, sc.n_elements + sc.n_traditional)
        assert rel(state.particle_F_trial.cpu().numpy()[sl], o.F_trial[sl]) < tol_aux
        assert rel(state.particle_F.cpu().numpy()[sl], o.F[sl]) < tol_aux
        assert rel(state.particle_stress.cpu().numpy()[sl], o.stress[sl]) < tol_aux


def test_c1_jelly_one_substep_with_grid():
    sc = S.scene_c1()
    o = run_oracle(sc, 1)
    solver, model, state = run_cuda(sc, 1, debug=True)
    compare(o, state, sc)
    gm, gvi, gvo = state.export_grid()
    gm, gvi, gvo = gm.cpu().numpy().reshape(-1), gvi.cpu().numpy().reshape(-1, 3), gvo.cpu().numpy().reshape(-1, 3)
    assert abs(gm.sum() - o.mass.sum()) < 1e-5 * o.mass.sum()
    assert rel(gm, o.grid_m) < 1e-5
    assert rel(gvi, o.grid_v_in) < 1e-4
    has = o.grid_m > 1e-15
    assert rel(gvo[has], o.grid_v_out[has]) < 1e-4


@pytest.mark.parametrize("material", ["jelly", "metal", "sand", "foam", "plasticine", "snow"])
def test_traditional_materials(material):
    sc = S.scene_c1(n=3000, n_grid=32, seed=11, material=material)
    o = run_oracle(sc, 5)
    _, _, state = run_cuda(sc, 5)
    compare(o, state, sc)


def reference_envelope(sc, nsub, joint_t=None):
    """Variability of the REFERENCE algorithm itself after nsub substeps: its float atomics make
    the accumulation order nondeterministic (mpm_utils.py:554-557) and fp32 round-off is
    amplified by contact and by the return mapping's branch at R22 = 1 (mpm_utils.py:196-204).
    Measured with the oracle as the max rel. deviation between (a) sequential and 8-thread
    accumulation, (b) fp32 and fp64, (c) inputs moved by one ulp (1e-7)."""
    from oracle.oracle import OracleSim
    a = run_oracle(sc, nsub, "f32", 1, joint_t)
    b = run_oracle(sc, nsub, "f32", 8, joint_t)
    c = run_oracle(sc, nsub, "f64", 1, joint_t)
    import copy
    sp = copy.copy(sc)
    sp.x = (sc.x + np.random.default_rng(1).uniform(-1e-7, 1e-7, sc.x.shape)).astype(np.float32)
    d = run_oracle(sp, nsub, "f32", 1, joint_t)
    ex = max(rel(b.x, a.x), rel(c.x, a.x), rel(d.x, a.x))
    ev = max(rel(b.v, a.v), rel(c.v, a.v), rel(d.v, a.v))
    return a, ex, ev


def test_small_cloth_body_joints_one_substep():
    sc = S.scene_small_cloth_body()
    o = run_oracle(sc, 1)
    _, _, state = run_cuda(sc, 1)
    compare(o, state, sc)


@pytest.mark.parametrize("nsub", [10, 100])
def test_small_cloth_body_joints_within_reference_envelope(nsub):
    """1e-4 relative on x and v after N substeps (BASELINE.json); where the reference's own
    variability (reference_envelope) is larger than that, 3x the envelope."""
    sc = S.scene_small_cloth_body()
    o, ex, ev = reference_envelope(sc, nsub)
    _, _, state = run_cuda(sc, nsub)
    x = state.particle_x.cpu().numpy()
    v = state.particle_v.cpu().numpy()
    assert np.isfinite(x).all() and np.isfinite(v).all()
    assert rel(x, o.x) < max(TOL_XV, 3 * ex), (rel(x, o.x), ex)
    assert rel(v, o.v) < max(TOL_XV, 3 * ev), (rel(v, o.v), ev)


def _load_oracle_state(o, sc, solver, model, state):
    """Put the oracle's current particle state into the CUDA solver (continue_from_torch path)."""
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    state.continue_from_torch(T(o.x), T(o.v), T(o.d) if sc.n_elements else None, T(o.C),
                              T(o.R_inv) if sc.n_elements else None)
    if sc.n_traditional:
        state.particle_F = T(o.F)
        state.particle_F_trial = T(o.F_trial)
    solver.time = o.time


@pytest.mark.parametrize("scene,k0", [("small_cloth_body", 10), ("small_cloth_body", 60), ("demo_like", 25)])
def test_one_substep_from_reference_midstate(scene, k0):
    """Pins the substep map away from the rest state: the oracle runs k0 substeps, its state is
    loaded into the CUDA solver, both advance ONE substep on the same inputs.  Strict 1e-4 on
    x and v (99.9th percentile at 1e-4, max at 3e-4: an element within one ulp of the R22 = 1
    branch of the return mapping may take the other branch)."""
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    sc = getattr(S, "scene_" + scene)()
    jt = np.zeros((sc.num_joint_t, 3), np.float32) if sc.num_joint_t else None
    o = run_oracle(sc, k0, "f32", 1, jt)
    solver, model, state = build_from_scene(sc)
    _load_oracle_state(o, sc, solver, model, state)
    fi = sc.frame_inputs(0)
    ft = frame_tensors(sc, 0)
    mx = fi["mesh_x"] + np.float32(sc.dt * k0) * fi["mesh_v"]
    o.p2g2p(sc.dt, mx, fi["mesh_v"], jt, fi["joint_verts_v"], fi["joint_faces_v"])
    jtt = None if jt is None else torch.as_tensor(jt, device="cuda")
    solver.p2g2p(model, state, sc.dt, mesh_x=torch.as_tensor(mx, device="cuda"), mesh_v=ft["mesh_v"],
                 joint_traditional_v=jtt, joint_verts_v=ft["joint_verts_v"], joint_faces_v=ft["joint_faces_v"])
    assert solver.stats()["overflow"] == 0
    x = state.particle_x.cpu().numpy()
    v = state.particle_v.cpu().numpy()
    assert rel(x, o.x) < TOL_XV
    ev = np.abs(v - o.v).max(1) / np.abs(o.v).max()
    assert np.quantile(ev, 0.999) < TOL_XV, np.quantile(ev, 0.999)
    assert ev.max() < 3 * TOL_XV, ev.max()
    Cc = state.particle_C.cpu().numpy()
    inv_dx = sc.n_grid / sc.grid_lim
    ec = np.abs(Cc - o.C).reshape(len(Cc), -1).max(1)
    assert np.quantile(ec, 0.999) < TOL_AUX * np.abs(o.C).max() + TOL_XV * np.abs(o.v).max() * 4 * inv_dx
    if sc.n_elements:
        assert rel(state.particle_d.cpu().numpy(), o.d) < TOL_AUX


def test_vertex_force_and_first_substep_details():
    sc = S.scene_small_cloth_body()
    # pre-stretch the cloth so that stress and vertex forces are non-trivial
    sc.x = sc.x.copy()
    Ne = sc.n_elements
    verts = sc.x[Ne:] * np.array([1.0, 1.03, 1.0], np.float32) + np.array([0, -0.03, 0], np.float32)
    sc.x[Ne:] = verts
    sc.x[:Ne] = verts[sc.faces].mean(1)
    d1 = verts[sc.faces[:, 1]] - verts[sc.faces[:, 0]]
    d2 = verts[sc.faces[:, 2]] - verts[sc.faces[:, 0]]
    d3 = np.cross(d1, d2)
    d3 /= np.linalg.norm(d3, axis=1, keepdims=True)  # keep d3 the unit normal: R22 = 1, no shear
    sc.d = np.stack([d1, d2, d3], -1).astype(np.float32)
    o = run_oracle(sc, 1)
    solver, model, state = run_cuda(sc, 1, debug=True)
    compare(o, state, sc)
    vf = state.vertex_force.cpu().numpy()
    assert np.abs(o.vertex_force).max() > 0
    assert rel(vf, o.vertex_force) < 1e-3


def test_multi_substep_call_equals_caller_loop():
    sc = S.scene_small_cloth_body()
    _, _, a = run_cuda(sc, 20, per_call=True)
    _, _, b = run_cuda(sc, 20, per_call=False)
    assert rel(b.particle_x.cpu().numpy(), a.particle_x.cpu().numpy()) < 1e-5
    assert rel(b.particle_v.cpu().numpy(), a.particle_v.cpu().numpy()) < 1e-3  # different mesh-advance arithmetic


def test_resort_interval_does_not_change_results():
    sc = S.scene_small_cloth_body()
    _, _, a = run_cuda(sc, 30, per_call=False, resort_interval=1000)
    s2, _, b = run_cuda(sc, 30, per_call=False, resort_interval=4)
    assert s2.stats()["n_resorts"] >= 7
    assert rel(b.particle_x.cpu().numpy(), a.particle_x.cpu().numpy()) < 1e-5
    assert rel(b.particle_v.cpu().numpy(), a.particle_v.cpu().numpy()) < 1e-3  # re-sorting changes the atomics order


def test_demo_like_sand_plane_pinned_tail():
    sc = S.scene_demo_like()
    jt = np.zeros((sc.num_joint_t, 3), np.float32)
    o1 = run_oracle(sc, 1, joint_t=jt)
    _, _, st1 = run_cuda(sc, 1, joint_t=jt)
    compare(o1, st1, sc)
    o, ex, ev = reference_envelope(sc, 10, jt)
    _, _, state = run_cuda(sc, 10, joint_t=jt)
    assert rel(state.particle_x.cpu().numpy(), o.x) < max(TOL_XV, 3 * ex)
    assert rel(state.particle_v.cpu().numpy(), o.v) < max(TOL_XV, 3 * ev)


def _xv_within(state, o, sc, nsub, label, joint_t=None):
    """1e-4 relative on x, v (BASELINE.json north_star); only if that is exceeded, the reference's own variability
    (reference_envelope) is measured and 3x of it allowed.  Prints both."""
    x, v = state.particle_x.cpu().numpy(), state.particle_v.cpu().numpy()
    assert np.isfinite(x).all() and np.isfinite(v).all()
    ex, ev = rel(x, o.x), rel(v, o.v)
    print(f"{label} N={nsub}: rel err x={ex:.2e} v={ev:.2e} (tolerance {TOL_XV:.0e})")
    if ex < TOL_XV and ev < TOL_XV:
        return
    _, env_x, env_v = reference_envelope(sc, nsub, joint_t)
    print(f"{label} N={nsub}: reference envelope x={env_x:.2e} v={env_v:.2e}")
    assert ex < max(TOL_XV, 3 * env_x), (ex, env_x)
    assert ev < max(TOL_XV, 3 * env_v), (ev, env_v)


@pytest.mark.parametrize("nsub", [1, 10, 100])
def test_c2_cloth_100k(nsub):
    """SURVEY 8c fixes N in {1, 10, 100} for C2."""
    sc = S.scene_c2()
    o = run_oracle(sc, nsub, threads=16)
    _, _, state = run_cuda(sc, nsub, per_call=(nsub <= 10))
    if nsub <= 10:
        compare(o, state, sc)
    _xv_within(state, o, sc, nsub, "C2")


@pytest.mark.parametrize("nsub", [1, 10])
def test_c3_oracle_parity(nsub):
    """BASELINE.json's headline config (499 968 particles / 256^3 / body collider + joints) against the fp32 oracle."""
    sc = S.scene_c3()
    o = run_oracle(sc, nsub, threads=16)
    _, _, state = run_cuda(sc, nsub, per_call=(nsub == 1))
    _xv_within(state, o, sc, nsub, "C3")
    # d, C: 1e-3 of max|.| for 99.9 % of the elements; an element within an ulp of the return mapping's R22 = 1 branch
    # (mpm_utils.py:196-204) may take the other branch -- shear kept instead of projected onto the cone -- which moves
    # its d3 by the size of that shear (the reference's own atomics order does the same to it); those stay below 1e-2
    ed = np.abs(state.particle_d.cpu().numpy() - o.d).reshape(sc.n_elements, -1).max(1) / np.abs(o.d).max()
    print(f"C3 N={nsub}: d err q99.9={np.quantile(ed, 0.999):.2e} max={ed.max():.2e}")
    assert np.quantile(ed, 0.999) < TOL_AUX, np.quantile(ed, 0.999)
    assert ed.max() < 10 * TOL_AUX, ed.max()
    Cc = state.particle_C.cpu().numpy()
    inv_dx = sc.n_grid / sc.grid_lim
    c_tol = TOL_AUX * np.abs(o.C).max() + TOL_XV * np.abs(o.v).max() * 4.0 * inv_dx
    ec = np.abs(Cc - o.C).reshape(len(Cc), -1).max(1)
    print(f"C3 N={nsub}: C err q99.9={np.quantile(ec, 0.999):.2e} max={ec.max():.2e} tol={c_tol:.2e}")
    assert np.quantile(ec, 0.999) < c_tol


def test_c3_full_size_properties():
    """BASELINE.json's headline size: size-independent properties instead of the dense oracle."""
    sc = S.scene_c3()
    solver, model, state = run_cuda(sc, 1, per_call=False, debug=True)
    gm, gvi, gvo = state.export_grid()
    total_mass = float(state.particle_mass.double().sum())
    assert abs(float(gm.double().sum()) - total_mass) < 1e-4 * total_mass  # mass conservation
    x0 = torch.as_tensor(sc.x, device="cuda")
    v = state.particle_v
    # free cloth starts at rest: after one substep every unconstrained particle has v ~ dt*g
    free = torch.ones(sc.n_particles, dtype=torch.bool, device="cuda")
    assert torch.isfinite(state.particle_x).all()
    assert float((state.particle_x - x0).abs().max()) < 1e-3
    assert float(v[free][:, 1].median()) == pytest.approx(-9.8 * sc.dt, rel=1e-2)
    st = solver.stats()
    assert st["overflow"] == 0 and st["n_active_nodes"] > 0
    solver.set_debug(False)
    ft = __import__("mpmavatar_b200.scene_setup", fromlist=["frame_tensors"]).frame_tensors(sc, 0)
    solver.step(model, state, sc.dt, 200, ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
    assert torch.isfinite(state.particle_x).all() and torch.isfinite(state.particle_v).all()
    assert solver.stats()["overflow"] == 0


def test_c5_full_size_properties():
    """BASELINE.json config 5 (2 002 176 particles / 512^3 / body collider + joints): the dense oracle grid would be 11 GB,
    so size-independent properties instead: mass conservation on the grid, free fall of the unconstrained cloth in the
    first substep, the joint ring driven at the body velocity, finiteness and no overflow over 100 substeps."""
    from mpmavatar_b200.scene_setup import frame_tensors
    sc = S.scene_c5()
    solver, model, state = run_cuda(sc, 1, per_call=False, debug=True)
    gm, gvi, gvo = state.export_grid()
    total_mass = float(state.particle_mass.double().sum())
    assert abs(float(gm.double().sum()) - total_mass) < 1e-4 * total_mass
    del gm, gvi, gvo
    v = state.particle_v
    assert float(v[:, 1].median()) == pytest.approx(-9.8 * sc.dt, rel=1e-2)
    ft = frame_tensors(sc, 0)
    Nnv = sc.n_no_vertices
    jv = v[Nnv:Nnv + sc.num_joint_v // 2]  # the top ring is surrounded by prescribed nodes only
    assert float((jv - ft["joint_verts_v"][: sc.num_joint_v // 2]).abs().max()) < 0.05 * float(ft["joint_verts_v"].abs().max())
    st = solver.stats()
    assert st["overflow"] == 0 and st["n_active_nodes"] > 300_000
    solver.set_debug(False)
    solver.step(model, state, sc.dt, 100, ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
    assert torch.isfinite(state.particle_x).all() and torch.isfinite(state.particle_v).all()
    assert solver.stats()["overflow"] == 0


def test_kats_on_gpu_free_fall_and_clamp():
    sc = S.scene_c1(n=500, n_grid=16, seed=5, material="snow")  # no stress branch -> zero stress
    sc.v[:] = np.array([0.3, -0.2, 0.1], np.float32)
    sc.F_trial = None
    _, _, state = run_cuda(sc, 1)
    v = state.particle_v.cpu().numpy()
    assert np.abs(v - (sc.v + sc.dt * np.array(sc.g, np.float32))).max() < 1e-5
    assert np.abs(state.particle_C.cpu().numpy()).max() < 1e-3
    dx = 2.0 / 16
    sc2 = S.scene_c1(n=2, n_grid=16, seed=5, material="snow")
    sc2.F_trial = None
    sc2.x = np.array([[2 * dx + 1e-6, 1, 1], [1, 2 - 2 * dx - 1e-6, 1]], np.float32)
    sc2.v = np.array([[-5, 0, 0], [0, 5, 0]], np.float32)
    sc2.dt = 1e-2
    _, _, st2 = run_cuda(sc2, 1)
    x = st2.particle_x.cpu().numpy()
    assert x[0, 0] == np.float32(2 * dx) and x[1, 1] == np.float32(2.0) - np.float32(2 * dx)


def test_model_change_invalidates_captured_graphs():
    """step(32) replays captured 16-substep graphs whose kernels take the model scalars by value; after
    set_parameters_dict(g=...) the next step(32) must use the new gravity (stale graphs)."""
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    sc = S.scene_small_cloth_body()
    ft = frame_tensors(sc, 0)
    args = (None, ft["joint_verts_v"], ft["joint_faces_v"])
    g2 = [3.0, -2.0, 1.0]

    def run(chunk):
        solver, model, state = build_from_scene(sc)
        for part in range(2):
            if part == 1:
                solver.set_parameters_dict(model, state, {"g": g2, "rpic_damping": 0.1})
            for k in range(0, 32, chunk):
                kk = part * 32 + k
                mx = ft["mesh_x"] + float(np.float32(sc.dt * kk)) * ft["mesh_v"]
                solver.step(model, state, sc.dt, chunk, mx, ft["mesh_v"], *args)
        return state.particle_x.cpu().numpy(), state.particle_v.cpu().numpy()
    xa, va = run(32)  # graphs
    xb, vb = run(1)   # direct launches pick the new values up by construction
    assert rel(xa, xb) < 1e-5
    assert rel(va, vb) < 1e-3
    solver, model, state = build_from_scene(sc)  # and the change does matter
    for part in range(2):
        mx = ft["mesh_x"] + float(np.float32(sc.dt * part * 32)) * ft["mesh_v"]
        solver.step(model, state, sc.dt, 32, mx, ft["mesh_v"], *args)
    assert rel(state.particle_v.cpu().numpy(), vb) > 1e-2


def test_new_stiffness_after_a_step_is_not_overwritten_by_the_lazy_export():
    """step(); set_E_nu_from_torch(); prepare_mu_lam(); step() -- the caller's rollout reset without reset_state
    (train_material_params.py:607-610): the second step must run with the NEW mu / lam (_bind exported the
    solver's old values over them)."""
    from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
    sc = S.scene_small_cloth_body()
    ft = frame_tensors(sc, 0)
    args = (ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
    T = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda")

    def run(E2):
        solver, model, state = build_from_scene(sc)
        solver.step(model, state, sc.dt, 10, *args)
        solver.set_E_nu_from_torch(model, T(sc.E * E2), T(sc.nu), T(sc.gamma), T(sc.kappa), "cuda:0")
        solver.prepare_mu_lam(model, state, "cuda:0")
        mu_set = model.mu.clone()
        solver.step(model, state, sc.dt, 10, *args)
        assert torch.equal(model.mu, mu_set)  # cloth: no return map mutates mu
        return state.particle_v.cpu().numpy(), float(mu_set[0])
    v1, mu1 = run(1.0)
    v50, mu50 = run(50.0)
    assert mu50 == pytest.approx(50.0 * mu1, rel=1e-5)
    assert rel(v50, v1) > 1e-3  # a 50x stiffer cloth moves differently: the new values did reach the kernels
    # velocity assigned through the lazy property after a step survives the pending export
    solver, model, state = build_from_scene(sc)
    solver.step(model, state, sc.dt, 4, *args)
    vnew = torch.full((sc.n_particles, 3), 0.25, device="cuda")
    state.particle_v = vnew
    assert torch.equal(state.particle_v, vnew)
    solver.step(model, state, sc.dt, 1, *args)
    assert abs(float(state.particle_v[:, 0].median()) - 0.25) < 0.05


def test_vertex_force_is_the_last_substeps_without_debug_mode():
    """state.vertex_force after p2g2p holds the last substep's cloth forces (the reference zeroes it at the START of
    p2g2p, mpm_solver.py:251-256) -- also without set_debug."""
    sc = S.scene_small_cloth_body()
    Ne = sc.n_elements
    sc.x = sc.x.copy()
    verts = sc.x[Ne:] * np.array([1.0, 1.03, 1.0], np.float32) + np.array([0, -0.03, 0], np.float32)
    sc.x[Ne:] = verts
    sc.x[:Ne] = verts[sc.faces].mean(1)
    d1 = verts[sc.faces[:, 1]] - verts[sc.faces[:, 0]]
    d2 = verts[sc.faces[:, 2]] - verts[sc.faces[:, 0]]
    d3 = np.cross(d1, d2)
    # d3 = 0.98 x the unit normal: R22 stays clear of the return mapping's R22 = 1 branch, whose outcome at exactly 1 is
    # decided by the last ulp (this test is about WHICH substep's forces are exported, not about that branch)
    d3 *= 0.98 / np.linalg.norm(d3, axis=1, keepdims=True)
    sc.d = np.stack([d1, d2, d3], -1).astype(np.float32)
    for nsub in (1, 3):
        o = run_oracle(sc, nsub)
        _, _, state = run_cuda(sc, nsub, debug=False)
        vf = state.vertex_force.cpu().numpy()
        assert np.abs(o.vertex_force).max() > 0
        assert rel(vf, o.vertex_force) < 1e-3, (nsub, rel(vf, o.vertex_force))
