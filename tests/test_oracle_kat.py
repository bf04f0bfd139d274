"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors, so each KAT below is an
analytic consequence of the cited reference lines; it must hold for the
reference and therefore for the oracle (and, in test_parity_gpu.py, for the
CUDA path)."""
import math

import numpy as np
import pytest

from oracle.oracle import OracleSim
from mpmavatar_b200 import synthetic as S


def _trad_sim(n, n_grid=16, precision="f32", material="jelly"):
    o = OracleSim(n, 0, 0, n_grid, 2.0, precision)
    o.set_parameters(material=material)
    return o


def test_struct_layout_and_sizes():
    for p in ("f32", "f64"):
        OracleSim(4, 0, 0, 8, 2.0, p).sim()


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_qr3_matches_gram_schmidt_and_rest_state(precision):
    rng = np.random.default_rng(0)
    o = _trad_sim(1, precision=precision)
    for _ in range(20):
        A = rng.normal(size=(3, 3))
        Q, R = o.qr3_signed(A)
        tol = 1e-5 if precision == "f32" else 1e-12
        assert np.allclose(Q @ R, A, atol=tol)
        assert np.allclose(Q.T @ Q, np.eye(3), atol=tol)
        assert np.linalg.det(Q) > 0 and R[0, 0] > 0 and R[1, 1] > 0
        assert abs(R[1, 0]) + abs(R[2, 0]) + abs(R[2, 1]) == 0


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_svd3_convention(precision):
    rng = np.random.default_rng(1)
    o = _trad_sim(1, precision=precision)
    tol = 2e-5 if precision == "f32" else 1e-10
    for k in range(30):
        A = rng.normal(size=(3, 3))
        if k % 3 == 0:
            A = np.eye(3) + 0.05 * rng.normal(size=(3, 3))
        U, s, V = o.svd3(A)
        assert np.allclose(U @ np.diag(s) @ V.T, A, atol=tol)
        assert np.linalg.det(U) > 0 and np.linalg.det(V) > 0
        assert s[0] >= s[1] >= abs(s[2]) - tol
        assert np.sign(s[2]) == np.sign(np.linalg.det(A))
    # cloth call site: upper-triangular 2x2 embedded in 3x3 -> polar rotation closed form
    F11, F12, F22 = 1.3, 0.4, 0.8
    U, s, V = o.svd3(np.array([[F11, F12, 0], [0, F22, 0], [0, 0, 0]]))
    Rot = (U @ V.T)[:2, :2]
    nrm = math.hypot(F11 + F22, F12)
    assert np.allclose(Rot, np.array([[F11 + F22, F12], [-F12, F11 + F22]]) / nrm, atol=tol)


def _rest_triangle():
    v = np.array([[1.0, 1.0, 1.0], [1.02, 1.001, 1.0], [1.003, 1.0, 1.015]])
    f = np.array([[0, 1, 2]])
    init_dir, rest_dir, evol, _ = S.compute_dir_vol(v, f)
    return init_dir[0].astype(np.float64), S.compute_rest_dir_inv(rest_dir)[0].astype(np.float64), float(evol[0])


def test_cloth_rest_state_is_stress_free():
    d, R_inv, vol = _rest_triangle()
    o = _trad_sim(1, precision="f64")
    mu, lam = 38.0, 57.0
    stress, f = o.aniso_stress(R_inv, d, vol, mu, lam, 500.0, 500.0)
    assert np.abs(stress).max() < 1e-5 * mu * vol
    for fi in f:
        assert np.abs(fi).max() < 1e-5 * mu * vol / 0.02
    nd = o.aniso_return_map(d, 500.0, 500.0, math.tan(math.radians(40)))
    assert np.allclose(nd, d, atol=1e-7)  # float32 rest data, float64 arithmetic


def test_cloth_inplane_stretch_closed_form():
    # orthonormal frame, stretch s along d1: F11=s, F22=1, F12=0
    o = _trad_sim(1, precision="f64")
    mu, lam, s, vol = 10.0, 20.0, 1.1, 2.0
    d = np.diag([s, 1.0, 1.0])
    R_inv = np.array([1.0, 0.0, 1.0])
    stress, (f1, f2, f3) = o.aniso_stress(R_inv, d, vol, mu, lam, 5.0, 7.0)
    K11 = 2 * mu * (s - 1) + lam * (s * 1.0 - 1) * 1.0
    # P = Q K3sym RiDT^-1 with Q=I, RiDT=diag(s,1,1): P11 = K11*s/s = K11
    assert np.allclose(f2, [-vol * K11, 0, 0], atol=1e-9)
    assert np.allclose(f1, -(f2 + f3), atol=1e-12)
    assert np.abs(stress).max() < 1e-12  # no normal / shear part


def test_cloth_normal_compression_and_return_map():
    o = _trad_sim(1, precision="f64")
    kappa, gamma, cf = 7.0, 5.0, math.tan(math.radians(40))
    r = 0.9
    d = np.diag([1.0, 1.0, r])
    stress, _ = o.aniso_stress(np.array([1.0, 0, 1.0]), d, 1.0, 10.0, 20.0, gamma, kappa)
    # dr33 = -kappa (1-r)^2 ; P33 = dr33*r / r ; stress = vol * P3 (x) d3 -> [2,2] = dr33 * r
    assert np.isclose(stress[2, 2], -kappa * (1 - r) ** 2 * r, atol=1e-12)
    d2 = np.diag([1.0, 1.0, 1.2])
    stress2, _ = o.aniso_stress(np.array([1.0, 0, 1.0]), d2, 1.0, 10.0, 20.0, gamma, kappa)
    assert np.abs(stress2).max() < 1e-12
    nd = o.aniso_return_map(d2, kappa, gamma, cf)
    assert np.isclose(nd[2, 2], 1.0)  # R22 > 1 is clamped to 1 (mpm_utils.py:196-197)
    # shear friction cone (mpm_utils.py:199-204)
    d3 = np.array([[1.0, 0, 0.3], [0, 1.0, 0.4], [0, 0, 0.9]])
    nd3 = o.aniso_return_map(d3, kappa, gamma, cf)
    fn = kappa * (1 - 0.9) ** 2
    ff = gamma * 0.5
    assert ff > cf * fn
    assert np.allclose(nd3[:, 2], [0.3 * cf * fn / ff, 0.4 * cf * fn / ff, 0.9], atol=1e-12)
    # R22>1 keeps the shear (quirk)
    d4 = np.array([[1.0, 0, 0.3], [0, 1.0, 0.4], [0, 0, 1.5]])
    assert np.allclose(o.aniso_return_map(d4, kappa, gamma, cf)[:, 2], [0.3, 0.4, 1.0], atol=1e-12)


def _fcr_stress(F, mu, lam, precision="f64"):
    o = _trad_sim(1, precision=precision)
    o.x[:] = 1.0
    o.vol[:] = 1.0
    o.mass[:] = 1.0
    o.mu[:] = mu
    o.lam[:] = lam
    o.F_trial[0] = F
    o.call("orc_compute_stress_from_F_trial", o.real(1e-4))
    return o.stress[0].copy(), o.F[0].copy()


def test_fcr_closed_forms():
    mu, lam = 3.0, 5.0
    st, F = _fcr_stress(np.eye(3), mu, lam)
    assert np.abs(st).max() < 1e-12 and np.allclose(F, np.eye(3))
    c, s = math.cos(0.7), math.sin(0.7)
    Q = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    st, _ = _fcr_stress(Q, mu, lam)
    assert np.abs(st).max() < 1e-10
    sc = 1.2
    st, _ = _fcr_stress(sc * np.eye(3), mu, lam)
    expect = (2 * mu * (sc - 1) * sc + lam * sc ** 3 * (sc ** 3 - 1)) * np.eye(3)
    assert np.allclose(st, expect, atol=1e-9)


def test_materials_without_a_stress_branch_give_zero():
    # material ids 4, 6, 7 have no branch for traditional particles (mpm_utils.py:1079-1095)
    for mat in ("snow", "neo-hookean", "cloth"):
        o = _trad_sim(1, precision="f64", material=mat)
        o.mu[:] = 1.0
        o.lam[:] = 1.0
        o.F_trial[0] = np.eye(3) * 1.3
        o.call("orc_compute_stress_from_F_trial", o.real(1e-4))
        assert np.abs(o.stress).max() == 0.0


def _cloud(n=200, n_grid=16, precision="f64", seed=3):
    rng = np.random.default_rng(seed)
    o = _trad_sim(n, n_grid, precision)
    o.x[:] = rng.uniform(0.6, 1.4, (n, 3))
    o.v[:] = rng.normal(0, 0.3, (n, 3))
    o.vol[:] = 1e-3
    o.mass[:] = rng.uniform(0.5, 1.5, n) * 1e-3
    return o, rng


def test_bspline_partition_mass_momentum_conservation():
    o, _ = _cloud()
    o.call("orc_zero_grid")
    o.call("orc_p2g_apic_with_stress", o.real(1e-4))
    assert np.isclose(o.grid_m.sum(), o.mass.sum(), rtol=1e-12)
    assert np.allclose(o.grid_v_in.sum(0), (o.mass[:, None] * o.v).sum(0), rtol=1e-10)


def test_free_fall_and_affine_roundtrip():
    o, _ = _cloud()
    o.set_parameters(material="snow", g=(0.0, -9.8, 0.0))  # no stress branch -> zero stress
    o.v[:] = np.array([0.3, -0.2, 0.1])
    v0 = o.v.copy()
    x0 = o.x.copy()
    dt = 1e-3
    o.p2g2p(dt)
    assert np.allclose(o.v, v0 + dt * np.array([0, -9.8, 0]), atol=1e-12)
    assert np.abs(o.C).max() < 1e-10
    assert np.allclose(o.x, x0 + dt * o.v, atol=1e-12)
    assert np.allclose(o.F_trial, np.eye(3)[None], atol=1e-10)  # grad v = 0


def test_position_clamp():
    o = _trad_sim(2, 16, "f64", "snow")
    dx = 2.0 / 16
    o.x[0] = [2 * dx + 1e-6, 1.0, 1.0]
    o.x[1] = [1.0, 2.0 - 2 * dx - 1e-6, 1.0]
    o.v[0] = [-5.0, 0, 0]
    o.v[1] = [0, 5.0, 0]
    o.vol[:] = 1.0
    o.mass[:] = 1.0
    o.p2g2p(1e-2)
    assert o.x[0, 0] == 2 * dx and o.x[1, 1] == 2.0 - 2 * dx


def test_collider_projection_closed_form():
    n_grid = 16
    o = _trad_sim(1, n_grid, "f64", "snow")
    # one big triangle near the centre with normal +y, moving with velocity vb
    verts = np.array([[0.9, 1.0, 0.9], [0.9, 1.0, 1.2], [1.2, 1.0, 0.9]])
    o.set_body_mesh(verts, np.array([[0, 1, 2]]), friction=0.5)
    vb = np.array([0.1, 0.2, 0.0])
    o.mesh_velocities[:] = vb
    s = o.sim()
    o.grid_v_out[:] = np.array([1.0, -2.0, 0.5])  # inward (n=+y): v_rel_n = -2.2
    before = o.grid_v_out.copy()
    o.call("orc_mesh_collider")
    hit = o.col_weight > 1e-15
    assert hit.sum() == 27
    assert np.array_equal(o.grid_v_out[~hit], before[~hit])
    vrel = before[0] - vb
    vproj = vrel - min(vrel[1], 0) * np.array([0, 1.0, 0])
    expect = max(0.0, np.linalg.norm(vproj) + vrel[1] * 0.5) * vproj / np.linalg.norm(vproj) + vb
    assert np.allclose(o.grid_v_out[hit], expect[None], atol=1e-12)


def test_mover_overrides_and_sticky_floor():
    sc = S.scene_demo_like()
    o = OracleSim.from_scene(sc, "f64")
    fi = sc.frame_inputs(0)
    jt = np.zeros((sc.num_joint_t, 3))
    o.p2g2p(sc.dt, fi["mesh_x"], fi["mesh_v"], jt, fi["joint_verts_v"], fi["joint_faces_v"])
    n = sc.n_grid
    gy = (np.arange(n ** 3) // n) % n
    below = gy * (2.0 / n) - 0.1 < 0
    assert np.abs(o.grid_v_out[below]).max() == 0.0
    hit = (o.mov_weight > 1e-15) & ~below
    want = o.mov_velocity[hit] / o.mov_weight[hit][:, None]
    assert np.allclose(o.grid_v_out[hit], want, atol=1e-12)
    assert np.isfinite(o.x).all() and np.isfinite(o.d).all()


def test_f32_tracks_f64_on_cloth_scene():
    sc = S.scene_small_cloth_body()
    a = OracleSim.from_scene(sc, "f32")
    b = OracleSim.from_scene(sc, "f64")
    fi = sc.frame_inputs(0)
    for k in range(10):
        mx = fi["mesh_x"] + np.float32(sc.dt * k) * fi["mesh_v"]
        for o in (a, b):
            o.p2g2p(sc.dt, mx, fi["mesh_v"], None, fi["joint_verts_v"], fi["joint_faces_v"])
    assert np.abs(a.x - b.x).max() < 1e-6
    assert np.abs(a.v - b.v).max() < 1e-4 * max(1.0, np.abs(b.v).max())


def test_openmp_matches_sequential():
    sc = S.scene_small_cloth_body()
    a = OracleSim.from_scene(sc, "f32", threads=1)
    b = OracleSim.from_scene(sc, "f32", threads=4)
    fi = sc.frame_inputs(0)
    for o in (a, b):
        for k in range(3):
            o.p2g2p(sc.dt, fi["mesh_x"], fi["mesh_v"], None, fi["joint_verts_v"], fi["joint_faces_v"])
    assert np.abs(a.x - b.x).max() < 1e-6


def test_reference_variability_envelope_and_no_input_aliasing():
    """Size of the reference algorithm's own variability (inputs moved by one ulp, fp32 vs fp64) after
    10 substeps with contact: a few 1e-5 relative in velocity, i.e. the 1e-4 parity tolerance of
    BASELINE.json is meaningful at N=10.  Also guards the harness: the oracle must not write into
    the Scene's arrays (p2g2p overwrites the body-mesh points in place)."""
    import copy
    sc = S.scene_small_cloth_body()
    body0 = sc.body_verts.copy()
    sp = copy.copy(sc)
    sp.x = (sc.x + np.random.default_rng(1).uniform(-1e-7, 1e-7, sc.x.shape)).astype(np.float32)
    sims = [OracleSim.from_scene(sc, "f32", threads=1), OracleSim.from_scene(sp, "f32", threads=1),
            OracleSim.from_scene(sc, "f64", threads=1)]
    for k in range(10):
        fi = sc.frame_inputs(0)
        mx = fi["mesh_x"] + np.float32(sc.dt * k) * fi["mesh_v"]
        for o in sims:
            o.p2g2p(sc.dt, mx, fi["mesh_v"], None, fi["joint_verts_v"], fi["joint_faces_v"])
    assert np.array_equal(sc.body_verts, body0)
    a, b, c = sims
    vm = np.abs(c.v).max()
    assert np.abs(a.v - b.v).max() / vm < 1e-4
    assert np.abs(a.v - c.v).max() / vm < 1e-4
    assert np.abs(a.x - c.x).max() < 1e-6
