"""MPMWARP with the reference's method signatures (/root/reference/warp_mpm/mpm_solver.py)
driving libmpm_b200.so through its C-ABI (include/mpm_b200.h).

Host code only: every per-particle / per-node operation of p2g2p runs in the hand-written
sm_100a kernels.  There is no CPU or PyTorch fallback -- a missing library or CUDA device
raises."""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace

import numpy as np
import torch

from .. import _lib
from .mpm_data_structure import MPMModelStruct, MPMStateStruct  # noqa: F401

_MATERIALS = {"jelly": 0, "metal": 1, "sand": 2, "foam": 3, "snow": 4, "plasticine": 5, "neo-hookean": 6,
              "cloth": 7}  # mpm_solver.py:58-76


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, dev):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous() or (dev is not None and t.device != dev and t.is_cuda):
        t = t.to(dtype=torch.float32).contiguous()
    return t


class MPMWARP(object):
    def __init__(self, n_particles, n_elements, n_vertices, n_grid=100, grid_lim=1.0, mesh_vertices=None,
                 mesh_faces=None, num_joint_t=0, num_joint_v=0, num_joint_f=0, device="cuda:0", resort_interval=0):
        self._h = None
        self.initialize(n_particles, n_elements, n_vertices, n_grid, grid_lim, mesh_vertices, mesh_faces, num_joint_t,
                        num_joint_v, num_joint_f, device=device, resort_interval=resort_interval)
        self.time_profile = {}

    # ---- mpm_solver.py:18-51
    def initialize(self, n_particles, n_elements, n_vertices, n_grid=100, grid_lim=1.0, mesh_vertices=None,
                   mesh_faces=None, num_joint_t=0, num_joint_v=0, num_joint_f=0, device="cuda:0", resort_interval=0):
        if not torch.cuda.is_available():
            raise RuntimeError("mpmavatar_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self._libh = _lib.load()
        self.device = torch.device(device)
        self.n_particles, self.n_elements, self.n_vertices = n_particles, n_elements, n_vertices
        self.n_no_vertices = n_particles - n_vertices
        self.num_joint_t, self.num_joint_v, self.num_joint_f = num_joint_t, num_joint_v, num_joint_f
        self.n_grid, self.grid_lim = n_grid, grid_lim
        self._time = 0.0
        self.grid_postprocess, self.collider_params, self.modify_bc = [], [], []
        self.mesh_colliders, self.mesh_collider_params = [], []
        self.particle_movers, self.particle_mover_params = [], []
        self.pre_p2g_operations, self.impulse_params = [], []
        self.particle_velocity_modifiers, self.particle_velocity_modifier_params = [], []
        self.num_mesh_v = self.num_mesh_f = 0
        if mesh_vertices is not None and mesh_faces is not None:
            self.num_mesh_v, self.num_mesh_f = int(mesh_vertices.shape[0]), int(mesh_faces.shape[0])
        cfg = _lib.MpmConfig(n_particles, n_elements, n_vertices, n_grid, float(grid_lim), self.num_mesh_v,
                             self.num_mesh_f, num_joint_v, num_joint_f, self.device.index or 0, int(resort_interval))
        if self._h is not None:
            self._libh.mpm_destroy(self._h)
        h = C.c_void_p()
        if self._libh.mpm_create(C.byref(cfg), C.byref(h)) != 0:
            raise RuntimeError("mpm_create: " + self._libh.mpm_last_error(None).decode())
        self._h = h
        self._bound_state = None
        self._bound_model = None
        self._model_sig = None
        if self.num_mesh_f:
            pts = torch.as_tensor(np.ascontiguousarray(mesh_vertices, dtype=np.float32), device=self.device)
            fcs = torch.as_tensor(np.ascontiguousarray(mesh_faces).astype(np.int32), device=self.device).contiguous()
            self._ck(self._libh.mpm_set_body_mesh(self._h, _ptr(fcs), _ptr(pts), self._stream()))
            # stands in for wp.Mesh: callers only read .id (train_material_params.py:505)
            self.mesh = SimpleNamespace(id=1, points=pts, velocities=torch.zeros_like(pts), indices=fcs.reshape(-1))
            torch.cuda.synchronize(self.device)

    def __del__(self):
        try:
            if self._h is not None:
                self._libh.mpm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("libmpm_b200: " + self._libh.mpm_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def time(self):
        return self._time

    @time.setter
    def time(self, t):
        self._time = float(t)
        self._ck(self._libh.mpm_set_time(self._h, C.c_double(self._time)))

    # ---- mpm_solver.py:54-126
    def set_parameters(self, device="cuda:0", **kwargs):
        self.set_parameters_dict(device, kwargs)

    def set_parameters_dict(self, mpm_model, mpm_state, kwargs={}, device="cuda:0"):
        if "material" in kwargs:
            if kwargs["material"] not in _MATERIALS:
                raise TypeError("Undefined material type")
            mpm_model.material = _MATERIALS[kwargs["material"]]
        if "yield_stress" in kwargs:
            mpm_model.yield_stress = torch.full_like(mpm_model.yield_stress, float(kwargs["yield_stress"]))
        if "hardening" in kwargs:
            mpm_model.hardening = kwargs["hardening"]
        if "xi" in kwargs:
            mpm_model.xi = kwargs["xi"]
        if "friction_angle" in kwargs:
            mpm_model.friction_angle = kwargs["friction_angle"]
            sin_phi = math.sin(mpm_model.friction_angle / 180.0 * 3.14159265)
            mpm_model.friction_coeff = math.tan(mpm_model.friction_angle / 180.0 * 3.14159265)
            mpm_model.alpha = math.sqrt(2.0 / 3.0) * 2.0 * sin_phi / (3.0 - sin_phi)
        if "g" in kwargs:
            mpm_model.gravitational_accelaration = (kwargs["g"][0], kwargs["g"][1], kwargs["g"][2])
        if "density" in kwargs:
            mpm_state.particle_density = torch.full_like(mpm_state.particle_vol, float(kwargs["density"]))
            mpm_state.particle_mass = mpm_state.particle_density * mpm_state.particle_vol
            mpm_state._dirty = True
        if "rpic_damping" in kwargs:
            mpm_model.rpic_damping = kwargs["rpic_damping"]
        if "plastic_viscosity" in kwargs:
            mpm_model.plastic_viscosity = kwargs["plastic_viscosity"]
        if "softening" in kwargs:
            mpm_model.softening = kwargs["softening"]
        if "grid_v_damping_scale" in kwargs:
            mpm_model.grid_v_damping_scale = kwargs["grid_v_damping_scale"]
        mpm_model._dirty = True

    # ---- mpm_solver.py:128-218
    def set_E_nu(self, mpm_model, E, nu, gamma, kappa, device="cuda:0"):
        def fill(cur, val):
            if isinstance(val, (float, int)):
                return torch.full_like(cur, float(val))
            return val.detach().to(cur.device, torch.float32).contiguous().clone()
        mpm_model.E = fill(mpm_model.E, E)
        mpm_model.nu = fill(mpm_model.nu, nu)
        mpm_model.gamma = fill(mpm_model.gamma, gamma)
        mpm_model.kappa = fill(mpm_model.kappa, kappa)
        mpm_model._dirty = True

    def set_E_nu_from_torch(self, mpm_model, E, nu, gamma, kappa, device="cuda:0"):
        conv = lambda t: t.item() if t.ndim == 0 else t
        self.set_E_nu(mpm_model, conv(E), conv(nu), conv(gamma), conv(kappa), device=device)

    # ---- mpm_solver.py:220-227, mpm_utils.py:402-408
    def prepare_mu_lam(self, mpm_model, mpm_state, device="cuda:0"):
        mpm_model.mu = mpm_model.E / (2.0 * (1.0 + mpm_model.nu))
        mpm_model.lam = mpm_model.E * mpm_model.nu / ((1.0 + mpm_model.nu) * (1.0 - 2.0 * mpm_model.nu))
        mpm_model._dirty = True

    # ---- binding canonical tensors to the C solver
    def _push_model_scalars(self, m):
        g = getattr(m, "gravitational_accelaration", (0.0, 0.0, 0.0))
        sig = (m.material, getattr(m, "hardening", 0), m.friction_coeff, m.alpha, tuple(g), m.rpic_damping,
               m.grid_v_damping_scale, getattr(m, "xi", 0.0), m.plastic_viscosity, m.softening)
        if sig == self._model_sig:
            return
        p = _lib.MpmModelParams(int(m.material), 1 if getattr(m, "hardening", 0) == 1 else 0, float(m.friction_coeff),
                                float(m.alpha), _lib.f3(g), float(m.rpic_damping), float(m.grid_v_damping_scale),
                                float(getattr(m, "xi", 0.0)), float(m.plastic_viscosity), float(m.softening))
        self._ck(self._libh.mpm_set_model(self._h, C.byref(p)))
        self._model_sig = sig

    def _bind(self, model, state):
        self._push_model_scalars(model)
        rebound = state is not self._bound_state or model is not self._bound_model
        if not (rebound or state._dirty or model._dirty):
            return
        if state._stale and state._solver is self:
            self._export_into(state)  # keep the dynamic fields the caller did not touch
        if int(state.particle_selection.abs().sum().item()) != 0:
            raise NotImplementedError("particle_selection != 0 is not supported by the B200 solver")
        dev = self.device
        keep = []

        def t(x):
            x = _f32(x, dev)
            keep.append(x)
            return _ptr(x)
        a = _lib.MpmParticleArrays()
        a.x, a.v, a.C = t(state._particle_x), t(state._particle_v), t(state._particle_C)
        a.F, a.F_trial = t(state._particle_F), t(state._particle_F_trial)
        a.d, a.R_inv, a.faces = t(state._particle_d), t(state.particle_R_inv), t(state.faces)
        a.vol, a.mass = t(state.particle_vol), t(state.particle_mass)
        a.mu, a.lam, a.gamma, a.kappa = t(model.mu), t(model.lam), t(model.gamma), t(model.kappa)
        a.yield_stress = t(model.yield_stress)
        self._ck(self._libh.mpm_import_state(self._h, C.byref(a), self._stream()))
        torch.cuda.current_stream(dev).synchronize()  # staging tensors in `keep` may be temporaries
        state._dirty = model._dirty = False
        state._stale = False
        state._solver = self
        model._solver = self
        self._bound_state, self._bound_model = state, model

    def _export_into(self, state):
        a = _lib.MpmParticleArrays()
        a.x, a.v, a.C = _ptr(state._particle_x), _ptr(state._particle_v), _ptr(state._particle_C)
        a.F, a.F_trial, a.stress = _ptr(state._particle_F), _ptr(state._particle_F_trial), _ptr(state._particle_stress)
        a.d, a.vertex_force = _ptr(state._particle_d), _ptr(state._vertex_force)
        m = self._bound_model
        # damage / hardening mutate these (mpm_utils.py:250,287-292); a model the caller has touched since the last
        # import (_dirty) keeps the caller's values
        if m is not None and not m._dirty and state is self._bound_state:
            a.mu, a.lam, a.yield_stress = _ptr(m._mu), _ptr(m._lam), _ptr(m._yield_stress)
        state._stale = False  # before the call: the lazy properties used above must not recurse
        self._ck(self._libh.mpm_export_state(self._h, C.byref(a), self._stream()))

    def _export_grid(self):
        n = self.n_grid
        gm = torch.zeros(n, n, n, dtype=torch.float32, device=self.device)
        gvi = torch.zeros(n, n, n, 3, dtype=torch.float32, device=self.device)
        gvo = torch.zeros(n, n, n, 3, dtype=torch.float32, device=self.device)
        self._ck(self._libh.mpm_export_grid(self._h, _ptr(gm), _ptr(gvi), _ptr(gvo), self._stream()))
        return gm, gvi, gvo

    # ---- mpm_solver.py:229-536
    def p2g2p(self, mpm_model, mpm_state, dt, mesh_x=None, mesh_v=None, joint_traditional_v=None, joint_verts_v=None,
              joint_faces_v=None, device="cuda:0"):
        self.step(mpm_model, mpm_state, dt, 1, mesh_x, mesh_v, joint_traditional_v, joint_verts_v, joint_faces_v)

    def step(self, mpm_model, mpm_state, dt, num_substeps=1, mesh_x=None, mesh_v=None, joint_traditional_v=None,
             joint_verts_v=None, joint_faces_v=None):
        """num_substeps p2g2p substeps in one call; substep k sees body points
        mesh_x + dt*k*mesh_v, i.e. the caller's inner loop (train_material_params.py:622-626)."""
        self._bind(mpm_model, mpm_state)
        dev = self.device
        fi = _lib.MpmFrameInputs()

        def ready(t):  # the common case costs three attribute reads: an fp32 CUDA tensor of this device, dense
            if t is None or (type(t) is torch.Tensor and t.dtype is torch.float32 and t.device == dev and t.is_contiguous()):
                return t
            return _f32(t, dev)
        keep = [ready(q) for q in (mesh_x, mesh_v, joint_traditional_v, joint_verts_v, joint_faces_v)]
        fi.device_inputs = 1 if all(k is None or k.is_cuda for k in keep) else 0
        if keep[0] is not None and keep[0].shape[0] != self.num_mesh_v:
            raise ValueError("mesh_x does not match the body mesh given at construction")
        fi.mesh_x, fi.mesh_v = _ptr(keep[0]), _ptr(keep[1])
        fi.joint_traditional_v = _ptr(keep[2])
        fi.n_joint_t = 0 if keep[2] is None else int(keep[2].shape[0])
        fi.joint_verts_v, fi.joint_faces_v = _ptr(keep[3]), _ptr(keep[4])
        if keep[3] is not None and keep[3].shape[0] < self.num_joint_v:
            raise ValueError("joint_verts_v shorter than num_joint_v")
        if keep[4] is not None and keep[4].shape[0] < self.num_joint_f:
            raise ValueError("joint_faces_v shorter than num_joint_f")
        self._ck(self._libh.mpm_step(self._h, C.c_float(dt), int(num_substeps), C.byref(fi), self._stream()))
        if any(k is not None and not k.is_cuda for k in keep):
            torch.cuda.current_stream(dev).synchronize()
        self._keep = keep
        mpm_state._stale = True
        mpm_state._solver = self
        for _ in range(int(num_substeps)):
            self._time = self._time + dt

    def set_particles(self, mpm_model, mpm_state):
        """north_star alias: (re)bind the canonical particle arrays now instead of lazily."""
        mpm_state._dirty = True
        self._bind(mpm_model, mpm_state)

    # ---- profiling: mpm_solver.py:16, 538-541
    def enable_profiling(self, on=True):
        self._ck(self._libh.mpm_set_profiling(self._h, 1 if on else 0))

    def get_profile(self):
        p = _lib.MpmProfile()
        self._ck(self._libh.mpm_get_profile(self._h, C.byref(p)))
        d = {n: getattr(p, n) for n, _ in _lib.MpmProfile._fields_}
        self.time_profile = {k: [v] for k, v in d.items() if k.endswith("_ms")}
        return d

    def print_time_profile(self):
        self.get_profile()
        print("MPM Time profile:")
        for key, value in self.time_profile.items():
            print(key, sum(value))

    def stats(self):
        st = _lib.MpmStats()
        self._ck(self._libh.mpm_get_stats(self._h, C.byref(st), self._stream()))
        return {n: getattr(st, n) for n, _ in _lib.MpmStats._fields_}

    def set_debug(self, on=True):
        self._ck(self._libh.mpm_set_debug(self._h, 1 if on else 0))

    # ---- mpm_solver.py:543-561, mpm_utils.py:1108-1132
    def export_particle_cov_to_torch(self, mpm_state, device="cuda:0"):
        F = mpm_state.particle_F_trial
        c = mpm_state.particle_cov.reshape(-1, 6)
        S = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], -1).reshape(-1, 3, 3)
        cov = F @ S @ F.transpose(1, 2)
        return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], -1).reshape(-1)

    # ---- boundary conditions
    def add_surface_collider(self, point, normal, surface="sticky", friction=0.0, start_time=0.0, end_time=999.0):
        point = list(point)
        normal_scale = 1.0 / math.sqrt(float(sum(x ** 2 for x in normal)))
        normal = list(normal_scale * x for x in normal)
        if surface == "sticky" and friction != 0:
            raise ValueError("friction must be 0 on sticky surfaces.")
        st = {"sticky": 0, "slip": 1, "cut": 11}.get(surface, 2)
        self._ck(self._libh.mpm_add_surface_collider(self._h, _lib.f3(point), _lib.f3(normal), st, float(friction),
                                                    float(start_time), float(end_time)))
        self.collider_params.append(dict(point=point, normal=normal, surface_type=st, friction=friction))
        self.grid_postprocess.append("surface")
        self.modify_bc.append(None)

    def add_particle_mover(self, n_grid):
        self._ck(self._libh.mpm_add_particle_mover(self._h))
        self.particle_movers.append("mover")
        self.particle_mover_params.append(None)

    def add_mesh_collider(self, mesh_id, n_grid, friction=0.0):
        self._ck(self._libh.mpm_add_mesh_collider(self._h, float(friction)))
        self.mesh_colliders.append("mesh_collider")
        self.mesh_collider_params.append(dict(mesh_id=mesh_id, friction=friction))

    def set_velocity_on_cuboid(self, point, size, velocity, start_time=0.0, end_time=999.0, reset=0):
        self._ck(self._libh.mpm_set_velocity_on_cuboid(self._h, _lib.f3(point), _lib.f3(size), _lib.f3(velocity),
                                                      float(start_time), float(end_time), int(reset)))
        self.collider_params.append(dict(point=list(point), size=size, velocity=velocity))
        self.grid_postprocess.append("cuboid")
        self.modify_bc.append("device")

    def add_bounding_box(self, start_time=0.0, end_time=999.0):
        self._ck(self._libh.mpm_add_bounding_box(self._h, float(start_time), float(end_time)))
        self.collider_params.append(dict())
        self.grid_postprocess.append("bounding_box")
        self.modify_bc.append(None)

    def enforce_grid_velocity_by_mask(self, selection_mask):
        m = selection_mask.to(self.device, torch.int32).contiguous()
        self._ck(self._libh.mpm_enforce_grid_velocity_by_mask(self._h, _ptr(m), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        self.collider_params.append(dict(mask=m))
        self.grid_postprocess.append("mask")
        self.modify_bc.append(None)

    # ---- pre-P2G particle operations (mpm_solver.py:1058-1151, 1289-1328, 1360-1417)
    def _box_mask(self, mpm_state, point, size):
        off = mpm_state.particle_x - torch.tensor(list(point), dtype=torch.float32, device=self.device)
        sz = torch.tensor(list(size), dtype=torch.float32, device=self.device)
        return (off.abs() < sz).all(dim=1).to(torch.int32).contiguous()

    def _add_op(self, kind, vec, mask, start_time, end_time):
        mask = mask.to(self.device, torch.int32).contiguous()
        self._ck(self._libh.mpm_add_particle_op(self._h, kind, _lib.f3(vec), _ptr(mask), float(start_time),
                                               float(end_time), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def add_impulse_on_particles(self, mpm_state, force, dt, point=[1, 1, 1], size=[1, 1, 1], num_dt=1, start_time=0.0,
                                 device="cuda:0"):
        self._add_op(0, force, self._box_mask(mpm_state, point, size), start_time, start_time + dt * num_dt)
        self.pre_p2g_operations.append("apply_force")
        self.impulse_params.append(None)

    def add_impulse_on_particles_with_mask(self, mpm_state, force, dt, particle_mask, point=[1, 1, 1], size=[1, 1, 1],
                                           end_time=1, start_time=0.0, device="cuda:0"):
        assert len(particle_mask) == self.n_particles, "mask should have n_particles elements"
        # as the reference: the caller's mask tensor is aliased and then OVERWRITTEN by the box selection
        # (mpm_solver.py:1381-1398), so the mask that acts is the box and the caller's tensor holds it afterwards
        box = self._box_mask(mpm_state, point, size)
        if isinstance(particle_mask, torch.Tensor):
            particle_mask.copy_(box.to(particle_mask.device, particle_mask.dtype))
        self._add_op(1, force, box, start_time, end_time)
        self.pre_p2g_operations.append("apply_force")
        self.impulse_params.append(None)

    def enforce_particle_velocity_translation(self, mpm_state, point, size, velocity, start_time, end_time,
                                              device="cuda:0"):
        self._add_op(2, velocity, self._box_mask(mpm_state, point, size), start_time, end_time)
        self.particle_velocity_modifiers.append("translation")
        self.particle_velocity_modifier_params.append(None)

    def enforce_particle_velocity_by_mask(self, mpm_state, selection_mask, velocity, start_time, end_time):
        self._add_op(2, velocity, selection_mask, start_time, end_time)
        self.particle_velocity_modifiers.append("by_mask")
        self.particle_velocity_modifier_params.append(None)

    def release_particles_sequentially(self, mpm_state, normal, start_position, end_position, num_layers, start_time,
                                       end_time):
        num_layers = 50
        point, size, axis = [0, 0, 0], [0, 0, 0], -1
        for i in range(3):
            if normal[i] == 0:
                point[i] = 1
                size[i] = 1
            else:
                axis = i
                point[i] = end_position
        half_length_portion = abs(start_position - end_position) / num_layers
        end_time_portion = end_time / num_layers
        for i in range(num_layers):
            size[axis] = half_length_portion * (num_layers - i)
            self.enforce_particle_velocity_translation(mpm_state, point, size, [0, 0, 0], start_time,
                                                       end_time_portion * (i + 1))

    def enforce_particle_velocity_rotation(self, mpm_state, point, normal, half_height_and_radius, rotation_scale,
                                           translation_scale, start_time, end_time, device="cuda:0"):
        """mpm_solver.py:1156-1256: particles inside the cylinder (point, normal, half height, radius) AT CALL TIME rotate
        about the axis with angular velocity rotation_scale and move along it with translation_scale."""
        import math
        ns = 1.0 / math.sqrt(float(normal[0] ** 2 + normal[1] ** 2 + normal[2] ** 2))
        f32 = lambda v: torch.tensor([float(c) for c in v], dtype=torch.float32)
        n = f32([ns * c for c in normal])
        h1 = torch.ones(3, dtype=torch.float32)
        if abs(float(torch.dot(n, h1))) < 0.01:
            h1 = f32([0.72, 0.37, -0.67])
        h1 = h1 - torch.dot(h1, n) * n
        h1 = h1 * (1.0 / torch.linalg.norm(h1))
        h2 = torch.linalg.cross(h1, n)
        dev = self.device
        off = mpm_state.particle_x - f32(point).to(dev)
        nd = n.to(dev)
        along = off @ nd
        vert = along.abs()
        hor = torch.linalg.norm(off - along[:, None] * nd, dim=1)
        mask = ((vert < float(half_height_and_radius[0])) & (hor < float(half_height_and_radius[1]))).to(torch.int32).contiguous()
        self._ck(self._libh.mpm_add_particle_rotation(self._h, _lib.f3(point), _lib.f3(n.tolist()), _lib.f3(h1.tolist()),
                                                      _lib.f3(h2.tolist()), float(rotation_scale), float(translation_scale),
                                                      _ptr(mask), float(start_time), float(end_time), self._stream()))
        torch.cuda.current_stream(dev).synchronize()
        self.particle_velocity_modifiers.append("rotation")
        self.particle_velocity_modifier_params.append(None)


class MPMSolverWarp(MPMWARP):
    """Name used by BASELINE.json's north_star; identical to MPMWARP."""
