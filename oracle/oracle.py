"""ctypes binding of the CPU oracle (oracle/mpm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mpm_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The reference has no tests or golden vectors and warp-lang is
not installable offline; the oracle is pinned (a) by analytic KATs
(tests/test_oracle_kat.py) and (b) against the reference's OWN kernel source executed
under oracle/warp_emu.py (tests/golden/*.npz, tests/test_golden.py: 1e-9 in fp64).
What stays assumed: the conventions of wp.qr3 / wp.svd3 (see warp_emu.py).

The Python surface mirrors the reference's call order (mpm_solver.py:229-536):
build an OracleSim from the canonical particle arrays, then call p2g2p().
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

MATERIALS = {"jelly": 0, "metal": 1, "sand": 2, "foam": 3, "snow": 4,
             "plasticine": 5, "neo-hookean": 6, "cloth": 7}  # mpm_solver.py:58-76


def build(force: bool = False) -> None:
    """Compile liboracle_f32.so / liboracle_f64.so with the committed Makefile."""
    need = force or not all(os.path.exists(os.path.join(_HERE, f"liboracle_{p}.so")) for p in ("f32", "f64"))
    src = os.path.join(_HERE, "mpm_oracle.c")
    if not need:
        for p in ("f32", "f64"):
            if os.path.getmtime(os.path.join(_HERE, f"liboracle_{p}.so")) < os.path.getmtime(src):
                need = True
    if need:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, env=env, stdout=subprocess.DEVNULL)


def _lib(precision: str):
    if precision not in _LIBS:
        build()
        lib = C.CDLL(os.path.join(_HERE, f"liboracle_{precision}.so"))
        lib.orc_sizeof_real.restype = C.c_int
        lib.orc_sizeof_sim.restype = C.c_int
        lib.orc_sizeof_bc.restype = C.c_int
        lib.orc_max_threads.restype = C.c_int
        _LIBS[precision] = lib
    return _LIBS[precision]


def _structs(real):
    rp = C.POINTER(real)
    ip = C.POINTER(C.c_int)

    class OrcBC(C.Structure):
        _fields_ = [("kind", C.c_int), ("surface_type", C.c_int), ("reset", C.c_int), ("pad_", C.c_int),
                    ("point", real * 3), ("normal", real * 3), ("size", real * 3), ("velocity", real * 3),
                    ("friction", real), ("start_time", real), ("end_time", real), ("mask", ip)]

    class OrcSim(C.Structure):
        _fields_ = [
            ("n_particles", C.c_int), ("n_elements", C.c_int), ("n_vertices", C.c_int), ("n_grid", C.c_int),
            ("grid_lim", real), ("dx", real), ("inv_dx", real),
            ("material", C.c_int), ("hardening_i", C.c_int),
            ("friction_coeff", real), ("alpha", real), ("g", real * 3),
            ("rpic_damping", real), ("grid_v_damping_scale", real),
            ("xi", real), ("plastic_viscosity", real), ("softening", real),
            ("x", rp), ("v", rp), ("C", rp), ("F", rp), ("F_trial", rp), ("stress", rp),
            ("d", rp), ("R_inv", rp), ("faces", rp), ("vertex_force", rp), ("vol", rp), ("mass", rp),
            ("traditional", ip), ("vertices", ip), ("elements", ip), ("selection", ip),
            ("mu", rp), ("lam", rp), ("gamma", rp), ("kappa", rp), ("yield_stress", rp),
            ("grid_m", rp), ("grid_v_in", rp), ("grid_v_out", rp),
            ("has_collider", C.c_int), ("n_mesh_v", C.c_int), ("n_mesh_f", C.c_int),
            ("collider_friction", real), ("mesh_faces", ip), ("mesh_points", rp), ("mesh_velocities", rp),
            ("col_weight", rp), ("col_v_in", rp), ("col_v_out", rp), ("col_normal", rp),
            ("has_mover", C.c_int), ("num_joint_v", C.c_int), ("num_joint_f", C.c_int),
            ("mov_weight", rp), ("mov_velocity", rp),
            ("n_bc", C.c_int), ("bc", C.POINTER(OrcBC)),
        ]

    return OrcSim, OrcBC


class OracleSim:
    """Dense-grid CPU solver in the reference's data layout.

    Particle order is [elements | traditional | vertices]
    (train_material_params.py:387); F/F_trial/stress have n_no_vertices rows,
    d/R_inv/faces have n_elements rows (mpm_data_structure.py:61-110)."""

    def __init__(self, n_particles, n_elements, n_vertices, n_grid, grid_lim, precision="f32", threads=1):
        self.precision = precision
        self.lib = _lib(precision)
        self.real = C.c_float if precision == "f32" else C.c_double
        self.np_real = np.float32 if precision == "f32" else np.float64
        assert self.lib.orc_sizeof_real() == C.sizeof(self.real)
        self._Sim, self._BC = _structs(self.real)
        assert self.lib.orc_sizeof_sim() == C.sizeof(self._Sim), "OrcSim layout mismatch"
        assert self.lib.orc_sizeof_bc() == C.sizeof(self._BC), "OrcBC layout mismatch"
        self.threads = threads
        self._impulses, self._modifiers = [], []  # pre-P2G particle operations
        N, Ne, Nv = n_particles, n_elements, n_vertices
        Nnv = N - Nv
        self.N, self.Ne, self.Nv, self.Nnv, self.Nt = N, Ne, Nv, Nnv, Nnv - Ne
        self.n_grid, self.grid_lim = n_grid, grid_lim
        r = self.np_real
        z = lambda *s: np.zeros(s, dtype=r)
        self.x, self.v, self.C = z(N, 3), z(N, 3), z(N, 3, 3)
        self.F, self.F_trial, self.stress = z(Nnv, 3, 3), z(Nnv, 3, 3), z(Nnv, 3, 3)
        self.F[:] = np.eye(3)
        self.F_trial[:] = np.eye(3)
        self.d, self.R_inv, self.faces = z(Ne, 3, 3), z(Ne, 3), z(Ne, 3)
        self.vertex_force = z(max(Nv, 1), 3)
        self.vol, self.mass = z(N), z(N)
        zi = lambda n: np.zeros(n, dtype=np.int32)
        self.traditional, self.vertices, self.elements, self.selection = zi(N), zi(N), zi(N), zi(N)
        self.elements[:Ne] = 1
        self.traditional[Ne:Nnv] = 1
        self.vertices[Nnv:] = 1
        self.mu, self.lam, self.gamma, self.kappa, self.yield_stress = z(N), z(N), z(N), z(N), z(N)
        n3 = n_grid ** 3
        self.grid_m, self.grid_v_in, self.grid_v_out = z(n3), z(n3, 3), z(n3, 3)
        # model defaults (mpm_data_structure.py:686-715)
        self.material = 0
        self.friction_coeff = 0.0
        self.alpha = 0.0
        self.g = (0.0, 0.0, 0.0)
        self.rpic_damping = 0.0
        self.grid_v_damping_scale = 1.1
        self.hardening = 0
        self.xi = 0.0
        self.plastic_viscosity = 0.0
        self.softening = 0.1
        self.has_collider = False
        self.has_mover = False
        self.num_joint_v = self.num_joint_f = 0
        self.collider_friction = 0.0
        self.bcs = []
        self._bc_keep = []
        self.time = 0.0
        self._sim = None

    # ---- parameter setters mirroring MPMWARP.set_parameters_dict (mpm_solver.py:57-126)
    def set_parameters(self, material=None, g=None, friction_angle=None, density=None, rpic_damping=None,
                       grid_v_damping_scale=None, yield_stress=None, hardening=None, xi=None,
                       plastic_viscosity=None, softening=None):
        if material is not None:
            if material not in MATERIALS:
                raise TypeError("Undefined material type")
            self.material = MATERIALS[material]
        if g is not None:
            self.g = tuple(float(a) for a in g)
        if friction_angle is not None:
            # float32 arithmetic like wp.sin/wp.tan on a python float (mpm_solver.py:90-94)
            ang = friction_angle / 180.0 * 3.14159265
            sin_phi = math.sin(ang)
            self.friction_coeff = math.tan(ang)
            self.alpha = math.sqrt(2.0 / 3.0) * 2.0 * sin_phi / (3.0 - sin_phi)
        if density is not None:
            self.mass[:] = (np.asarray(density, dtype=self.np_real) * self.vol).astype(self.np_real)
        if rpic_damping is not None:
            self.rpic_damping = rpic_damping
        if grid_v_damping_scale is not None:
            self.grid_v_damping_scale = grid_v_damping_scale
        if yield_stress is not None:
            self.yield_stress[:] = yield_stress
        if hardening is not None:
            self.hardening = hardening
        if xi is not None:
            self.xi = xi
        if plastic_viscosity is not None:
            self.plastic_viscosity = plastic_viscosity
        if softening is not None:
            self.softening = softening
        self._sim = None

    def set_E_nu(self, E, nu, gamma, kappa):
        """set_E_nu + prepare_mu_lam (mpm_solver.py:128-227, mpm_utils.py:402-408)."""
        r = self.np_real
        E = np.broadcast_to(np.asarray(E, dtype=r), (self.N,)).astype(r)
        nu = np.broadcast_to(np.asarray(nu, dtype=r), (self.N,)).astype(r)
        self.gamma[:] = gamma
        self.kappa[:] = kappa
        self.mu[:] = E / (r(2.0) * (r(1.0) + nu))
        self.lam[:] = E * nu / ((r(1.0) + nu) * (r(1.0) - r(2.0) * nu))

    def set_body_mesh(self, verts, faces, friction=0.0):
        r = self.np_real
        self.mesh_points = np.array(verts, dtype=r, order="C", copy=True)  # p2g2p overwrites these in place
        self.mesh_velocities = np.zeros_like(self.mesh_points)
        self.mesh_faces = np.array(faces, dtype=np.int32, order="C", copy=True)
        n3 = self.n_grid ** 3
        self.col_weight = np.zeros(n3, r)
        self.col_v_in = np.zeros((n3, 3), r)
        self.col_v_out = np.zeros((n3, 3), r)
        self.col_normal = np.zeros((n3, 3), r)
        self.collider_friction = friction
        self.has_collider = True
        self._sim = None

    def add_particle_mover(self, num_joint_v, num_joint_f):
        r = self.np_real
        n3 = self.n_grid ** 3
        self.mov_weight = np.zeros(n3, r)
        self.mov_velocity = np.zeros((n3, 3), r)
        self.num_joint_v, self.num_joint_f = num_joint_v, num_joint_f
        self.has_mover = True
        self._sim = None

    def add_surface_collider(self, point, normal, surface="sticky", friction=0.0, start_time=0.0, end_time=999.0):
        nrm = np.asarray(normal, dtype=np.float64)
        nrm = nrm / math.sqrt(float((nrm ** 2).sum()))
        if surface == "sticky" and friction != 0:
            raise ValueError("friction must be 0 on sticky surfaces.")
        st = {"sticky": 0, "slip": 1, "cut": 11}.get(surface, 2)
        self.bcs.append(dict(kind=0, surface_type=st, point=point, normal=nrm, friction=friction,
                             start_time=start_time, end_time=end_time))
        self._sim = None

    def set_velocity_on_cuboid(self, point, size, velocity, start_time=0.0, end_time=999.0, reset=0):
        self.bcs.append(dict(kind=1, point=point, size=size, velocity=velocity, start_time=start_time,
                             end_time=end_time, reset=reset))
        self._sim = None

    def add_bounding_box(self, start_time=0.0, end_time=999.0):
        self.bcs.append(dict(kind=2, start_time=start_time, end_time=end_time))
        self._sim = None

    def enforce_grid_velocity_by_mask(self, mask):
        m = np.ascontiguousarray(mask, dtype=np.int32).reshape(-1)
        self._bc_keep.append(m)
        self.bcs.append(dict(kind=3, mask=m))
        self._sim = None

    # ---- struct assembly
    def _ptr(self, a, t=None):
        t = t or self.real
        return a.ctypes.data_as(C.POINTER(t))

    def _build(self):
        s = self._Sim()
        s.n_particles, s.n_elements, s.n_vertices, s.n_grid = self.N, self.Ne, self.Nv, self.n_grid
        s.grid_lim = self.grid_lim
        # dx, inv_dx as in init_other_params (mpm_data_structure.py:692-697)
        s.dx = self.grid_lim / self.n_grid
        s.inv_dx = float(self.n_grid / self.grid_lim)
        s.material = self.material
        s.hardening_i = 1 if self.hardening == 1 else 0
        s.friction_coeff, s.alpha = self.friction_coeff, self.alpha
        s.g = (self.real * 3)(*self.g)
        s.rpic_damping, s.grid_v_damping_scale = self.rpic_damping, self.grid_v_damping_scale
        s.xi, s.plastic_viscosity, s.softening = self.xi, self.plastic_viscosity, self.softening
        for name in ("x", "v", "C", "F", "F_trial", "stress", "d", "R_inv", "faces", "vertex_force", "vol",
                     "mass", "mu", "lam", "gamma", "kappa", "yield_stress", "grid_m", "grid_v_in", "grid_v_out"):
            a = getattr(self, name)
            assert a.flags["C_CONTIGUOUS"] and a.dtype == self.np_real, name
            setattr(s, name, self._ptr(a))
        for name in ("traditional", "vertices", "elements", "selection"):
            setattr(s, name, self._ptr(getattr(self, name), C.c_int))
        s.has_collider = int(self.has_collider)
        if self.has_collider:
            s.n_mesh_v, s.n_mesh_f = self.mesh_points.shape[0], self.mesh_faces.shape[0]
            s.collider_friction = self.collider_friction
            s.mesh_faces = self._ptr(self.mesh_faces, C.c_int)
            for name in ("mesh_points", "mesh_velocities", "col_weight", "col_v_in", "col_v_out", "col_normal"):
                setattr(s, name, self._ptr(getattr(self, name)))
        s.has_mover = int(self.has_mover)
        if self.has_mover:
            s.num_joint_v, s.num_joint_f = self.num_joint_v, self.num_joint_f
            s.mov_weight, s.mov_velocity = self._ptr(self.mov_weight), self._ptr(self.mov_velocity)
        self._bc_arr = (self._BC * max(len(self.bcs), 1))()
        for i, b in enumerate(self.bcs):
            e = self._bc_arr[i]
            e.kind = b["kind"]
            e.surface_type = b.get("surface_type", 0)
            e.reset = b.get("reset", 0)
            for key in ("point", "normal", "size", "velocity"):
                if key in b:
                    setattr(e, key, (self.real * 3)(*[float(q) for q in b[key]]))
            e.friction = b.get("friction", 0.0)
            e.start_time = b.get("start_time", 0.0)
            e.end_time = b.get("end_time", 999.0)
            if "mask" in b:
                e.mask = self._ptr(b["mask"], C.c_int)
        s.n_bc = len(self.bcs)
        s.bc = C.cast(self._bc_arr, C.POINTER(self._BC))
        self._sim = s

    def sim(self):
        if self._sim is None:
            self._build()
        return self._sim

    # ---- phases
    def _opt(self, a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=self.np_real)
        self._keep = getattr(self, "_keep", [])
        self._keep.append(a)
        return self._ptr(a)

    # ---- pre-P2G particle operations (mpm_solver.py:1058-1328, 1360-1417; selection kernels mpm_utils.py:1198-1248).
    # They only touch particle_v, so they are restated here in numpy at the working precision instead of in C.
    # Same method names and arguments as MPMWARP (the first, mpm_state, argument is ignored).
    def _box_mask(self, point, size):
        r = self.np_real
        off = self.x.astype(r) - np.asarray(point, r)
        return (np.abs(off) < np.asarray(size, r)).all(1).astype(np.int32)

    def add_impulse_on_particles(self, mpm_state, force, dt, point=(1, 1, 1), size=(1, 1, 1), num_dt=1, start_time=0.0,
                                 device=None):
        r = self.np_real
        self._impulses.append(dict(kind="per_mass", force=np.asarray(force, r), mask=self._box_mask(point, size),
                                   t0=r(start_time), t1=r(r(start_time) + r(dt) * r(num_dt))))

    def add_impulse_on_particles_with_mask(self, mpm_state, force, dt, particle_mask, point=(1, 1, 1), size=(1, 1, 1),
                                           end_time=1, start_time=0.0, device=None):
        # the reference aliases the caller's mask and then OVERWRITES it with the box selection (:1381-1398): the
        # mask that acts is the box, and the caller's tensor holds it afterwards
        r = self.np_real
        assert len(particle_mask) == self.x.shape[0], "mask should have n_particles elements"
        box = self._box_mask(point, size)
        try:
            particle_mask[...] = particle_mask.new_tensor(box) if hasattr(particle_mask, "new_tensor") else box
        except Exception:  # noqa: BLE001 -- read-only input: the side effect is not observable
            pass
        self._impulses.append(dict(kind="plain", force=np.asarray(force, r), mask=box, t0=r(start_time), t1=r(end_time)))

    def enforce_particle_velocity_translation(self, mpm_state, point, size, velocity, start_time, end_time, device=None):
        r = self.np_real
        self._modifiers.append(dict(kind="set", velocity=np.asarray(velocity, r), mask=self._box_mask(point, size),
                                    t0=r(start_time), t1=r(end_time)))

    def enforce_particle_velocity_by_mask(self, mpm_state, selection_mask, velocity, start_time, end_time):
        r = self.np_real
        self._modifiers.append(dict(kind="set", velocity=np.asarray(velocity, r),
                                    mask=np.asarray(selection_mask).astype(np.int32), t0=r(start_time), t1=r(end_time)))

    def enforce_particle_velocity_rotation(self, mpm_state, point, normal, half_height_and_radius, rotation_scale,
                                           translation_scale, start_time, end_time, device=None):
        r = self.np_real
        n = np.asarray(normal, r)
        n = n * (r(1.0) / np.sqrt(r(n[0] ** 2 + n[1] ** 2 + n[2] ** 2)))
        h1 = np.ones(3, r)
        if abs(np.dot(n, h1)) < 0.01:
            h1 = np.asarray([0.72, 0.37, -0.67], r)
        h1 = h1 - np.dot(h1, n) * n
        h1 = h1 * (r(1.0) / np.sqrt(np.dot(h1, h1)))
        h2 = np.cross(h1, n)
        off = self.x.astype(r) - np.asarray(point, r)
        vert = np.abs(off @ n)
        hor = np.linalg.norm(off - (off @ n)[:, None] * n, axis=1)
        mask = ((vert < r(half_height_and_radius[0])) & (hor < r(half_height_and_radius[1]))).astype(np.int32)
        self._modifiers.append(dict(kind="rotate", point=np.asarray(point, r), n=n, h1=h1, h2=h2, rot=r(rotation_scale),
                                    trans=r(translation_scale), mask=mask, t0=r(start_time), t1=r(end_time)))

    def _apply_particle_ops(self, dt):
        """mpm_solver.py:260-279: all impulses in the order added, then all velocity modifiers in the order added."""
        r = self.np_real
        t, dt = r(self.time), r(dt)
        for op in self._impulses:
            if t >= op["t0"] and t < op["t1"]:
                if op["kind"] == "per_mass":
                    sel = op["mask"] == 1
                    self.v[sel] = (self.v[sel].astype(r) + op["force"][None, :] / self.mass[sel].astype(r)[:, None] * dt).astype(self.v.dtype)
                else:
                    sel = op["mask"] >= 1
                    self.v[sel] = (self.v[sel].astype(r) + op["force"][None, :] * dt).astype(self.v.dtype)
        for op in self._modifiers:
            if not (t >= op["t0"] and t < op["t1"]):
                continue
            sel = op["mask"] == 1
            if op["kind"] == "set":
                self.v[sel] = op["velocity"].astype(self.v.dtype)
            else:
                off = self.x[sel].astype(r) - op["point"]
                hd = np.linalg.norm(off - (off @ op["n"])[:, None] * op["n"], axis=1)
                theta = np.arccos((off @ op["h1"]) / hd)
                theta = np.where(off @ op["h2"] > 0, theta, -theta)
                a1 = -hd * np.sin(theta) * op["rot"]
                a2 = hd * np.cos(theta) * op["rot"]
                self.v[sel] = (a1[:, None] * op["h1"] + a2[:, None] * op["h2"] + op["trans"] * op["n"]).astype(self.v.dtype)

    def p2g2p(self, dt, mesh_x=None, mesh_v=None, joint_traditional_v=None, joint_verts_v=None, joint_faces_v=None):
        """One substep, mpm_solver.py:229-536."""
        if self._impulses or self._modifiers:
            self._apply_particle_ops(dt)
        self.lib.orc_set_threads(C.c_int(self.threads))
        self._keep = []
        njt = 0 if joint_traditional_v is None else int(np.asarray(joint_traditional_v).shape[0])
        self.lib.orc_p2g2p(C.byref(self.sim()), self.real(dt), self.real(self.time), self._opt(mesh_x),
                           self._opt(mesh_v), self._opt(joint_traditional_v), C.c_int(njt),
                           self._opt(joint_verts_v), self._opt(joint_faces_v))
        self.time = self.time + dt

    def call(self, fn, *args):
        self.lib.orc_set_threads(C.c_int(self.threads))
        getattr(self.lib, fn)(C.byref(self.sim()), *args)

    # ---- small function-level entry points for KATs
    def qr3_signed(self, A):
        A = np.ascontiguousarray(A, dtype=self.np_real)
        Q = np.zeros((3, 3), self.np_real)
        Rm = np.zeros((3, 3), self.np_real)
        self.lib.orc_qr3_signed(self._ptr(A), self._ptr(Q), self._ptr(Rm))
        return Q, Rm

    def svd3(self, A):
        A = np.ascontiguousarray(A, dtype=self.np_real)
        U = np.zeros((3, 3), self.np_real)
        V = np.zeros((3, 3), self.np_real)
        S = np.zeros(3, self.np_real)
        self.lib.orc_svd3(self._ptr(A), self._ptr(U), self._ptr(S), self._ptr(V))
        return U, S, V

    def aniso_stress(self, R_inv, d, vol, mu, lam, gamma, kappa):
        r = self.np_real
        R_inv = np.ascontiguousarray(R_inv, dtype=r)
        d = np.ascontiguousarray(d, dtype=r)
        out = np.zeros((3, 3), r)
        f = [np.zeros(3, r) for _ in range(3)]
        R = self.real
        self.lib.orc_kirchoff_stress_anisotropy(self._ptr(R_inv), self._ptr(d), R(vol), R(mu), R(lam), R(gamma),
                                                R(kappa), self._ptr(out), *[self._ptr(a) for a in f])
        return out, f

    def aniso_return_map(self, d, kappa, gamma, friction_coeff):
        r = self.np_real
        d = np.ascontiguousarray(d, dtype=r)
        nd = np.zeros((3, 3), r)
        R = self.real
        self.lib.orc_anisotropy_return_mapping(self._ptr(d), R(kappa), R(gamma), R(friction_coeff), self._ptr(nd))
        return nd

    @classmethod
    def from_scene(cls, sc, precision="f32", threads=1):
        """Build from a mpmavatar_b200.synthetic.Scene (plain numpy container)."""
        o = cls(sc.n_particles, sc.n_elements, sc.n_vertices, sc.n_grid, sc.grid_lim, precision, threads)
        r = o.np_real
        o.x[:] = sc.x
        o.v[:] = sc.v
        o.vol[:] = sc.vol
        if sc.n_elements:
            o.d[:] = sc.d
            o.R_inv[:] = sc.R_inv
            o.faces[:] = sc.faces.astype(r)
        if sc.F_trial is not None:
            o.F_trial[:] = sc.F_trial
        o.set_parameters(material=sc.material, g=sc.g, friction_angle=sc.friction_angle,
                         grid_v_damping_scale=sc.grid_v_damping_scale, rpic_damping=sc.rpic_damping)
        o.set_parameters(density=sc.density)
        o.set_E_nu(sc.E, sc.nu, sc.gamma, sc.kappa)
        if sc.yield_stress is not None:
            o.yield_stress[:] = sc.yield_stress
        if sc.body_verts is not None:
            o.set_body_mesh(sc.body_verts, sc.body_faces, sc.mesh_friction)
        if sc.num_joint_v or sc.num_joint_f:
            o.add_particle_mover(sc.num_joint_v, sc.num_joint_f)
        for b in sc.surface_colliders:
            o.add_surface_collider(**b)
        return o


def max_threads() -> int:
    """Host threads the oracle may use: the cores this process is allowed on.  omp_get_max_threads() alone is not
    enough: torchrun exports OMP_NUM_THREADS=1, which would make the CPU arm of bench.py a one-thread run."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(int(_lib("f32").orc_max_threads()), n)
