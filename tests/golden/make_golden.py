#!/usr/bin/env python
"""Generates tests/golden/*.npz by executing the REFERENCE's own solver source -- unmodified, imported from
/root/reference/warp_mpm -- under oracle/warp_emu.py (a sequential pure-Python stand-in for the Warp API;
warp-lang 0.10.1 is not installable offline).  Only runs in the build container (the GPU box has no
/root/reference); the .npz files it writes are committed and are what tests/ compare against.

    python tests/golden/make_golden.py            # regenerate every fixture (about a minute)

Each fixture stores the canonical inputs (so the tests do not depend on synthetic.py staying unchanged), the
per-substep solver inputs, and the reference's state after N substeps in fp64 emulation ("ref64_*": pins the
algorithm) and in fp32 emulation ("ref32_*": Warp's storage / arithmetic width, sequential atomics).
The call sequence is the caller's: train_material_params.py:403-506 (setup) and :589-626 (rollout).
Assumptions about wp.qr3 / wp.svd3 are stated in oracle/warp_emu.py.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/warp_mpm"

from oracle import warp_emu  # noqa: E402
from mpmavatar_b200 import synthetic as S  # noqa: E402


def import_reference():
    wp = warp_emu.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import mpm_data_structure as ds  # the reference's files
    import mpm_solver as sv
    return wp, ds, sv


def tiny_scenes():
    out = {}
    for mat, n in (("jelly", 120), ("metal", 60), ("sand", 60), ("foam", 60), ("plasticine", 60), ("snow", 40)):
        sc = S.scene_c1(n=n, n_grid=12, seed=21, material=mat)
        out[f"trad_{mat}"] = (sc, 2, None)
    sc = S._cloth_scene("golden_cloth_body", 5, 14, 8, 20, with_body=True)
    sc.body_verts, sc.body_faces = S.capsule_mesh(segs=14, rings_cyl=6, rings_cap=3, radius=0.22, cyl_len=0.7,
                                                  center=(1.0, 1.0, 1.0))
    # pre-stretch + a velocity field so that stress, vertex forces and the collider all act
    Ne = sc.n_elements
    rng = np.random.default_rng(3)
    verts = sc.x[Ne:] * np.array([1.0, 1.04, 1.0], np.float32) + np.array([0, -0.04, 0], np.float32)
    sc.x = sc.x.copy()
    sc.x[Ne:] = verts
    sc.x[:Ne] = verts[sc.faces].mean(1)
    sc.v = (0.3 * rng.normal(size=sc.x.shape)).astype(np.float32)
    d1 = verts[sc.faces[:, 1]] - verts[sc.faces[:, 0]]
    d2 = verts[sc.faces[:, 2]] - verts[sc.faces[:, 0]]
    d3 = np.cross(d1, d2)
    d3 /= np.linalg.norm(d3, axis=1, keepdims=True)
    d3 = d3 * (1.0 + 0.05 * rng.normal(size=(Ne, 1))) + 0.05 * rng.normal(size=(Ne, 3))  # both return-map branches
    sc.d = np.stack([d1, d2, d3], -1).astype(np.float32)
    out["cloth_body_joints"] = (sc, 3, None)
    sd = S.scene_demo_like(Nu=10, Nr=6, n_sand=80, n_grid=20, seed=7)
    sd.body_verts, sd.body_faces = S.capsule_mesh(segs=10, rings_cyl=4, rings_cap=2, radius=0.22, cyl_len=0.7,
                                                  center=(1.0, 1.0, 1.0))
    sd.v = (0.2 * np.random.default_rng(4).normal(size=sd.x.shape)).astype(np.float32)
    out["demo_sand_plane_pinned"] = (sd, 2, np.zeros((sd.num_joint_t, 3), np.float32))
    # pre-P2G particle operations (mpm_solver.py:1058-1328, 1360-1417): every modifier the API offers, with time
    # windows that open / close inside the 4 recorded substeps (dt = 1e-4)
    so = S.scene_c1(n=150, n_grid=12, seed=22, material="jelly")
    out["trad_jelly_particle_ops"] = (so, 4, None)
    # grid boundary conditions beyond the sticky plane (mpm_solver.py:564-658, 929-1053, 1330-1355), grid damping
    # (mpm_utils.py:1162-1174) and RPIC damping (:528-542): a jelly block next to the x = 0 wall
    sg = S.scene_c1(n=150, n_grid=12, seed=23, material="jelly")
    sg.x = (sg.x + np.array([-0.4, 0.0, 0.0], np.float32)).astype(np.float32)
    sg.v = (0.3 * np.random.default_rng(5).normal(size=sg.x.shape)).astype(np.float32)
    sg.grid_v_damping_scale = 0.98
    sg.rpic_damping = 0.2
    out["trad_jelly_grid_bcs"] = (sg, 4, None)
    # non-default plasticity parameters (set_parameters_dict keys, mpm_solver.py:86-124): von Mises hardening
    # (mpm_utils.py:250-252), viscoplastic return map (:315-359), damage softening down to mu = lam = 0 (:287-292)
    for name, (mat, _) in MATERIAL_VARIANTS.items():
        out[name] = (S.scene_c1(n=60, n_grid=12, seed=21, material=mat), 4, None)
    # grid sizes that are NOT multiples of the solver's 4^3 blocks -- the reference's real operating points are 200, 150
    # (scripts/physics/actorshq_a2_lower.sh:36) and 250 (run_demo.py:142): 150 = 250 = 2 (mod 4).  Particles are driven into
    # the upper walls so that the position clamp [2dx, lim - 2dx] (mpm_utils.py:767-778) acts in the last, partial block.
    for n_grid, mat, seed in ((14, "jelly", 31), (15, "sand", 32)):
        sw = S.scene_c1(n=90, n_grid=n_grid, seed=seed, material=mat)
        rng = np.random.default_rng(seed)
        dxw = 2.0 / n_grid
        lo, hi = 2.0 - 4.5 * dxw, 2.0 - 2.0 * dxw  # a block of particles within 2.5 cells of the x, y upper walls
        sw.x = np.stack([rng.uniform(lo, hi, 90), rng.uniform(lo, hi, 90), rng.uniform(0.9, 1.3, 90)], 1).astype(np.float32)
        sw.x[:6] = np.float32(hi)  # some start exactly on the clamp value
        sw.v = (np.array([6.0, 9.0, -1.0], np.float32) + 2.0 * rng.normal(size=(90, 3))).astype(np.float32)
        sw.dt = 1e-3  # the fastest particles cross the clamp in the first substep, the rest within three
        out[f"trad_{mat}_grid{n_grid}_walls"] = (sw, 3, None)
    sc18 = S._cloth_scene("golden_cloth_body_grid18", 6, 12, 7, 18, with_body=True)
    sc18.body_verts, sc18.body_faces = S.capsule_mesh(segs=12, rings_cyl=5, rings_cap=3, radius=0.22, cyl_len=0.7,
                                                      center=(1.0, 1.0, 1.0))
    sc18.v = (0.3 * np.random.default_rng(6).normal(size=sc18.x.shape)).astype(np.float32)
    out["cloth_body_joints_grid18"] = (sc18, 3, None)
    s50 = S.scene_demo_like(Nu=16, Nr=8, n_sand=120, n_grid=50, seed=9)
    s50.body_verts, s50.body_faces = S.capsule_mesh(segs=16, rings_cyl=8, rings_cap=3, radius=0.22, cyl_len=0.7,
                                                    center=(1.0, 1.0, 1.0))
    s50.v = (0.2 * np.random.default_rng(10).normal(size=s50.x.shape)).astype(np.float32)
    out["demo_sand_plane_pinned_grid50"] = (s50, 2, np.zeros((s50.num_joint_t, 3), np.float32))
    return out


MATERIAL_VARIANTS = {
    "trad_metal_hardening": ("metal", dict(hardening=1, xi=5.0)),
    "trad_foam_viscous": ("foam", dict(plastic_viscosity=0.05)),
    "trad_plasticine_softening": ("plasticine", dict(softening=40.0)),
}


def apply_material_variant(name, set_params):
    """set_params(dict) forwards to the solver's parameter setter (set_parameters_dict / OracleSim.set_parameters)."""
    set_params(dict(MATERIAL_VARIANTS[name][1]))


# (method, kwargs) issued after the solver is set up; the order is the order the reference applies them in.
# The first cuboid is long expired but has reset = 1: the reference then zeroes the WHOLE grid while
# time < end_time + 15 dt (:966-970), here exactly for the first substep.  The second one switches on inside the run
# and moves with its velocity (host-side modify_bc, :975-981).
GRID_BCS = [
    ("add_surface_collider", dict(point=[1.0, 0.86, 1.0], normal=[0.0, 1.0, 0.0], surface="slip", friction=0.3)),
    ("set_velocity_on_cuboid", dict(point=[0.6, 1.0, 1.0], size=[0.1, 0.1, 0.1], velocity=[0.0, 0.0, 0.0], start_time=-1.0,
                                    end_time=-14.5e-4, reset=1)),
    ("set_velocity_on_cuboid", dict(point=[0.6, 1.0, 1.0], size=[0.12, 0.25, 0.25], velocity=[0.3, 0.0, 0.1], start_time=1.5e-4,
                                    end_time=10.0, reset=0)),
    ("add_bounding_box", dict(start_time=0.5e-4, end_time=999.0)),
    ("enforce_grid_velocity_by_mask", dict(selection_mask="upper_y")),
]


def grid_mask(name, n):
    m = np.zeros((n, n, n), np.int32)
    if name == "upper_y":
        m[:, 7:, :] = 1
    else:
        raise KeyError(name)
    return m


def apply_grid_bcs(solver, n_grid, to_mask=None):
    """Issues GRID_BCS on `solver` (the reference's MPMWARP, this repo's mirror, or the oracle: same names and arguments);
    to_mask converts the [n,n,n] int32 numpy mask into what the solver takes."""
    import torch
    for meth, kw in GRID_BCS:
        kw = dict(kw)
        if "selection_mask" in kw:
            m = grid_mask(kw.pop("selection_mask"), n_grid)
            getattr(solver, meth)(to_mask(m) if to_mask else torch.from_numpy(m))
        else:
            getattr(solver, meth)(**kw)


# (method, kwargs) in the order the caller issues them; replayed on the oracle / CUDA mirror by tests/test_golden.py.
# The reference applies all impulses first, then all velocity modifiers, whatever the order of the calls
# (mpm_solver.py:260-279) -- the translation below is therefore issued BEFORE the first impulse on purpose.
PARTICLE_OPS = [
    ("enforce_particle_velocity_translation", dict(point=[0.9, 1.0, 1.0], size=[0.06, 0.3, 0.3], velocity=[0.0, 0.2, -0.1],
                                                   start_time=1.5e-4, end_time=10.0)),
    ("add_impulse_on_particles", dict(force=[3e-4, 0.0, -1e-4], dt=1e-4, point=[1.1, 1.0, 1.0], size=[0.08, 0.3, 0.3], num_dt=2,
                                      start_time=0.5e-4)),
    ("add_impulse_on_particles_with_mask", dict(force=[0.0, 5.0, 0.0], dt=1e-4, particle_mask="first_third",
                                                point=[1.0, 1.1, 1.0], size=[0.3, 0.07, 0.3], end_time=2.5e-4, start_time=0.0)),
    ("enforce_particle_velocity_by_mask", dict(selection_mask="every_seventh", velocity=[0.05, 0.0, 0.0], start_time=0.0,
                                               end_time=1.0)),
    ("enforce_particle_velocity_rotation", dict(point=[1.0, 0.95, 1.0], normal=[0.0, 2.0, 0.0], half_height_and_radius=[0.08, 0.12],
                                                rotation_scale=3.0, translation_scale=0.4, start_time=0.0, end_time=3.5e-4)),
]


def named_mask(name, n):
    m = np.zeros(n, np.int32)
    if name == "first_third":
        m[: n // 3] = 1
    elif name == "every_seventh":
        m[::7] = 1
    else:
        raise KeyError(name)
    return m


def apply_particle_ops(solver, state, n, device=None, to_tensor=None):
    """Issues PARTICLE_OPS on `solver` (the reference's MPMWARP or this repo's mirror: same method names and arguments)."""
    import torch
    for meth, kw in PARTICLE_OPS:
        kw = dict(kw)
        for k in ("particle_mask", "selection_mask"):
            if k in kw:
                t = torch.from_numpy(named_mask(kw[k], n))
                kw[k] = to_tensor(t) if to_tensor else t
        if device is not None and meth not in ("enforce_particle_velocity_by_mask",):
            kw["device"] = device
        getattr(solver, meth)(state, **kw)


def run_reference(sc, nsub, joint_t, precision, with_ops=False, with_bcs=False, variant=None):
    """setup_simulation + rollout exactly as the reference's caller does, on the emulated Warp."""
    warp_emu.set_precision(precision)
    wp, ds, sv = import_reference()
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
    N, Ne, Nv, Nt = sc.n_particles, sc.n_elements, sc.n_vertices, sc.n_traditional
    dev = "cpu"
    with contextlib.redirect_stdout(io.StringIO()):
        state = ds.MPMStateStruct()
        state.init(N, Ne, Nv, device=dev, requires_grad=True)
        trad = np.zeros(N, np.int32); trad[Ne:Ne + Nt] = 1
        vert = np.zeros(N, np.int32); vert[Ne + Nt:] = 1
        elem = np.zeros(N, np.int32); elem[:Ne] = 1
        d = T(sc.d) if Ne else torch.zeros(0, 3, 3)
        R_inv = T(sc.R_inv) if Ne else torch.zeros(0, 3)
        faces = T(sc.faces.astype(np.float32)) if Ne else torch.zeros(0, 3)
        D_inv = torch.linalg.inv(d) if Ne else d
        state.from_torch(T(sc.x), T(sc.vol), D_inv, R_inv, faces, trad, vert, elem, torch.zeros(N - Nv, 6),
                         device=dev, requires_grad=True, n_grid=sc.n_grid, grid_lim=sc.grid_lim)
        model = ds.MPMModelStruct()
        model.init(N, device=dev, requires_grad=True)
        model.init_other_params(n_grid=sc.n_grid, grid_lim=sc.grid_lim, device=dev)
        solver = sv.MPMWARP(N, Ne, Nv, n_grid=sc.n_grid, grid_lim=sc.grid_lim, mesh_vertices=sc.body_verts,
                            mesh_faces=sc.body_faces, num_joint_t=sc.num_joint_t, num_joint_v=sc.num_joint_v,
                            num_joint_f=sc.num_joint_f, device=dev)
        solver.set_parameters_dict(model, state, {"material": sc.material, "g": list(sc.g), "density": 1.0,
                                                  "grid_v_damping_scale": sc.grid_v_damping_scale,
                                                  "friction_angle": sc.friction_angle,
                                                  "rpic_damping": sc.rpic_damping}, device=dev)
        for b in sc.surface_colliders:
            solver.add_surface_collider(**b)
        if sc.body_verts is not None:
            solver.add_mesh_collider(solver.mesh.id, n_grid=sc.n_grid, friction=sc.mesh_friction)
        if sc.num_joint_v or sc.num_joint_f:
            solver.add_particle_mover(n_grid=sc.n_grid)
        # rollout reset: train_material_params.py:584-610
        state.reset_state(Nv, T(sc.x).clone(), d, None, T(sc.v).clone(), tensor_R_inv=R_inv, device=dev,
                          requires_grad=True)
        if sc.F_trial is not None:
            state.particle_F_trial = wp.from_numpy(sc.F_trial, dtype=wp.mat33)
        state.reset_density(T(sc.density), None, dev, update_mass=True)
        solver.set_E_nu_from_torch(model, T(sc.E), T(sc.nu), T(sc.gamma), T(sc.kappa), dev)
        if sc.yield_stress is not None:
            model.yield_stress = wp.from_numpy(sc.yield_stress, dtype=float)
        solver.prepare_mu_lam(model, state, dev)
        if variant:
            apply_material_variant(variant, lambda kw: solver.set_parameters_dict(model, state, kw, device=dev))
        if with_ops:
            apply_particle_ops(solver, state, N, device=dev)
        if with_bcs:
            apply_grid_bcs(solver, sc.n_grid)
        fi = sc.frame_inputs(0)
        t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32)
        for k in range(nsub):
            mx = None if fi["mesh_x"] is None else fi["mesh_x"] + np.float32(sc.dt * k) * fi["mesh_v"]
            solver.p2g2p(model, state, sc.dt, mesh_x=t(mx), mesh_v=t(fi["mesh_v"]), joint_traditional_v=t(joint_t),
                         joint_verts_v=t(fi["joint_verts_v"]), joint_faces_v=t(fi["joint_faces_v"]), device=dev)
    g = lambda a: np.array(a.numpy(), dtype=np.float64)
    out = dict(x=g(state.particle_x), v=g(state.particle_v), C=g(state.particle_C), F=g(state.particle_F),
               F_trial=g(state.particle_F_trial), stress=g(state.particle_stress), d=g(state.particle_d),
               vertex_force=g(state.vertex_force), grid_m=g(state.grid_m), grid_v_in=g(state.grid_v_in),
               grid_v_out=g(state.grid_v_out), time=np.float64(solver.time),
               yield_stress=g(model.yield_stress), mu=g(model.mu), lam=g(model.lam))  # damage / hardening mutate these
    return out


SCENE_FIELDS = ("x", "v", "vol", "density", "E", "nu", "gamma", "kappa", "faces", "d", "R_inv", "F_trial",
                "yield_stress", "body_verts", "body_faces")
SCENE_SCALARS = ("name", "n_grid", "grid_lim", "dt", "material", "n_elements", "n_traditional", "n_vertices",
                 "friction_angle", "grid_v_damping_scale", "rpic_damping", "substeps_per_frame", "mesh_friction",
                 "num_joint_v", "num_joint_f", "num_joint_t")


def make_cov_fixture():
    """export_particle_cov_to_torch / compute_cov_from_F (mpm_solver.py:543-561, mpm_utils.py:1108-1132) run from the
    reference source on random F_trial and covariances; stored as tests/golden/cov_export.npz."""
    warp_emu.set_precision("f64")
    wp, ds, sv = import_reference()
    rng = np.random.default_rng(41)
    N, Ne, Nv = 40, 12, 10
    Nnv = N - Nv
    Ft = np.eye(3)[None] + 0.3 * rng.normal(size=(Nnv, 3, 3))
    a = rng.normal(size=(Nnv, 3, 3))
    cov33 = a @ a.transpose(0, 2, 1)
    cov6 = np.stack([cov33[:, 0, 0], cov33[:, 0, 1], cov33[:, 0, 2], cov33[:, 1, 1], cov33[:, 1, 2], cov33[:, 2, 2]], 1)
    with contextlib.redirect_stdout(io.StringIO()):
        state = ds.MPMStateStruct()
        state.init(N, Ne, Nv, device="cpu", requires_grad=False)
        state.particle_F_trial = wp.from_numpy(Ft, dtype=wp.mat33)
        state.particle_cov = wp.from_numpy(cov6.reshape(-1), dtype=float)
        solver = sv.MPMWARP.__new__(sv.MPMWARP)
        solver.n_no_vertices = Nnv
        solver.time_profile = {}
        out = solver.export_particle_cov_to_torch(state, device="cpu")
    path = os.path.join(HERE, "cov_export.npz")
    np.savez_compressed(path, F_trial=Ft.astype(np.float32), cov=cov6.astype(np.float32).reshape(-1),
                        ref64_new_cov=np.asarray(out.numpy(), np.float64), counts=np.asarray([N, Ne, Nv]))
    print(f"cov_export: Nnv={Nnv} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    only = set(sys.argv[1:])
    if not only or "cov_export" in only:
        make_cov_fixture()
    for name, (sc, nsub, joint_t) in tiny_scenes().items():
        if only and name not in only:
            continue
        rec = {"nsub": np.int64(nsub), "g": np.asarray(sc.g, np.float64)}
        for f in SCENE_FIELDS:
            v = getattr(sc, f)
            if v is not None:
                rec["in_" + f] = np.asarray(v)
        for f in SCENE_SCALARS:
            rec["sc_" + f] = np.asarray(getattr(sc, f))
        fi = sc.frame_inputs(0)
        for k, v in fi.items():
            if v is not None:
                rec["fi_" + k] = v
        if joint_t is not None:
            rec["fi_joint_traditional_v"] = joint_t
        if name in MATERIAL_VARIANTS:
            rec["material_variant"] = np.int64(1)  # replay MATERIAL_VARIANTS[name]
        if name.endswith("grid_bcs"):
            rec["grid_bcs"] = np.int64(1)  # replay tests/golden/make_golden.py GRID_BCS
        if name.endswith("particle_ops"):
            rec["particle_ops"] = np.int64(1)  # replay tests/golden/make_golden.py PARTICLE_OPS
        if sc.surface_colliders:
            rec["plane_point"] = np.asarray(sc.surface_colliders[0]["point"], np.float64)
            rec["plane_normal"] = np.asarray(sc.surface_colliders[0]["normal"], np.float64)
        for prec, tag in (("f64", "ref64_"), ("f32", "ref32_")):
            r = run_reference(sc, nsub, joint_t, prec, with_ops=name.endswith("particle_ops"), with_bcs=name.endswith("grid_bcs"),
                              variant=name if name in MATERIAL_VARIANTS else None)
            for k, v in r.items():
                if k.startswith("grid_") and tag == "ref32_":
                    continue
                rec[tag + k] = v.astype(np.float64 if tag == "ref64_" else np.float32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: N={sc.n_particles} grid={sc.n_grid} nsub={nsub} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
