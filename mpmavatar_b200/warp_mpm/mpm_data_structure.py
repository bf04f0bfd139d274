"""MPMStateStruct / MPMModelStruct with the reference's method signatures
(/root/reference/warp_mpm/mpm_data_structure.py:13-530, 610-733) over torch tensors.

Arrays keep the reference's canonical layout ([elements | traditional | vertices] particle
order, row-major mat33, faces as float vec3).  The solver keeps its own cell-sorted copy in
HBM; the dynamic arrays below are therefore lazy views: reading one after a step triggers a
single un-permuting export (mpm_export_state), so `wp.to_torch(state.particle_x)`
(train_material_params.py:628) returns current positions in ORIGINAL order."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from torch import Tensor

from .warp_utils import from_torch_safe

_DYNAMIC = ("particle_x", "particle_v", "particle_C", "particle_F", "particle_F_trial", "particle_stress",
            "particle_d", "vertex_force")


def _dev(device):
    return torch.device(device if device is not None else "cuda:0")


def _lazy(name):
    key = "_" + name

    def get(self):
        self._pull()
        return getattr(self, key)

    def set_(self, value):
        # the solver may hold newer data than the canonical tensors: bring them up to date BEFORE the caller's value
        # goes in, so that the next (lazy) export cannot overwrite it
        self._pull()
        setattr(self, key, value)
        self._dirty = True

    return property(get, set_)


_MODEL_ARRAYS = ("E", "nu", "mu", "lam", "gamma", "kappa", "yield_stress")


def _model_array(name):
    """Per-particle model arrays.  The return maps mutate mu / lam / yield_stress in place in the reference
    (mpm_utils.py:250-252, 287-292); here they live in the solver's sorted records between exports, so a read pulls
    them first and a write pulls BEFORE the new tensor is bound (otherwise the next export would overwrite it)."""
    key = "_" + name

    def get(self):
        self._pull()
        return getattr(self, key)

    def set_(self, value):
        self._pull()
        setattr(self, key, value)
        self._dirty = True

    return property(get, set_)


class MPMStateStruct:
    for _n in _DYNAMIC:
        locals()[_n] = _lazy(_n)
    del _n

    def __init__(self):
        self._dirty = True      # canonical tensors changed since the solver last imported them
        self._stale = False     # the solver holds newer data than the canonical tensors
        self._solver = None
        self.n_grid = 10
        self.grid_lim = 1.0

    def _pull(self):
        if self._stale and self._solver is not None:
            self._solver._export_into(self)

    # ---- mpm_data_structure.py:51-134
    def init(self, n_particles: int, n_elements: int, n_vertices: int, device=None, requires_grad=False) -> None:
        dev = _dev(device)
        n_no_vertices = n_particles - n_vertices
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        zi = lambda n: torch.zeros(n, dtype=torch.int32, device=dev)
        self._particle_x = z(n_particles, 3)
        self._particle_v = z(n_particles, 3)
        self._particle_F = z(n_no_vertices, 3, 3)
        self._particle_d = z(n_elements, 3, 3)
        self.particle_cov = z(n_no_vertices * 6)
        self._particle_F_trial = z(n_no_vertices, 3, 3)
        self._particle_stress = z(n_no_vertices, 3, 3)
        self._particle_C = z(n_particles, 3, 3)
        self.particle_vol = z(n_particles)
        self.particle_mass = z(n_particles)
        self.particle_density = z(n_particles)
        self.particle_R_inv = z(n_elements, 3)
        self.particle_D_inv = z(n_elements, 3, 3)
        self.faces = z(n_elements, 3)
        self._vertex_force = z(n_vertices, 3)
        self.particle_traditional = zi(n_particles)
        self.particle_vertices = zi(n_particles)
        self.particle_elements = zi(n_particles)
        self.particle_selection = zi(n_particles)
        self.n_particles, self.n_elements, self.n_vertices = n_particles, n_elements, n_vertices
        self._dirty = True

    # ---- :136-156.  The reference allocates three dense n^3 arrays here; the B200 grid is a
    # sparse block pool owned by the solver, so only the resolution is recorded.
    def init_grid(self, grid_res: int, device=None, requires_grad=False):
        self.n_grid = grid_res

    # dense views for debugging (grid_m / grid_v_in / grid_v_out of the last substep)
    def export_grid(self):
        if self._solver is None:
            raise RuntimeError("state is not bound to a solver yet")
        return self._solver._export_grid()

    # ---- :158-260
    def from_torch(self, tensor_x: Tensor, tensor_volume: Tensor, tensor_D_inv: Tensor, tensor_R_inv: Tensor,
                   tensor_faces: Tensor, particle_traditional, particle_vertices, particle_elements,
                   tensor_cov: Optional[Tensor] = None, tensor_velocity: Optional[Tensor] = None, n_grid: int = 100,
                   grid_lim=1.0, device="cuda:0", requires_grad=True):
        dev = _dev(device)
        assert tensor_x.shape[0] == tensor_volume.shape[0]
        self.init_grid(grid_res=n_grid, device=device, requires_grad=requires_grad)
        self.grid_lim = grid_lim
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous().clone()
        if tensor_x is not None:
            self._particle_x = f32(tensor_x)
        if tensor_volume is not None:
            self.particle_vol = f32(tensor_volume)
        if tensor_D_inv is not None:
            self.particle_D_inv = f32(tensor_D_inv)
        if tensor_R_inv is not None:
            self.particle_R_inv = f32(tensor_R_inv)
        if tensor_faces is not None:
            self.faces = f32(tensor_faces)  # float vec3, as in the reference (:211-215)
        if tensor_cov is not None:
            self.particle_cov = f32(tensor_cov.reshape(-1))
        if tensor_velocity is not None:
            self._particle_v = f32(tensor_velocity)
        as_i = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int32).to(dev)
        self.particle_traditional = as_i(particle_traditional)
        self.particle_vertices = as_i(particle_vertices)
        self.particle_elements = as_i(particle_elements)
        self._dirty = True
        print("Particles initialized from torch data.")
        print("Total particles: ", tensor_x.shape[0])

    # ---- :262-374
    def reset_state(self, n_vertices, tensor_x: Tensor, tensor_d: Tensor, tensor_cov: Optional[Tensor] = None,
                    tensor_velocity: Optional[Tensor] = None, tensor_density: Optional[Tensor] = None,
                    selection_mask: Optional[Tensor] = None, tensor_R_inv: Optional[Tensor] = None, device="cuda:0",
                    requires_grad=True):
        self._pull()
        dev = _dev(device)
        n_particles = tensor_x.shape[0]
        n_no_vertices = n_particles - n_vertices
        if tensor_x is not None:
            self._particle_x = from_torch_safe(tensor_x)  # aliased, not cloned (:283-287)
        if tensor_d is not None:
            self._particle_d = from_torch_safe(tensor_d).clone()
        if tensor_R_inv is not None:
            self.particle_R_inv = from_torch_safe(tensor_R_inv).clone()
        if tensor_cov is not None:
            self.particle_cov = tensor_cov.reshape(-1).detach().clone().to(dev)
        if tensor_velocity is not None:
            self._particle_v = from_torch_safe(tensor_velocity).clone()
        if tensor_density is not None and selection_mask is not None:
            m = selection_mask.to(dev).type(torch.int) == 1
            self.particle_density = torch.where(m, tensor_density.to(dev).float(), self.particle_density)
        self._particle_C = torch.zeros(n_particles, 3, 3, dtype=torch.float32, device=dev)
        eye = torch.eye(3, dtype=torch.float32, device=dev)
        self._particle_F_trial = eye.repeat(n_no_vertices, 1, 1)
        self._particle_F = eye.repeat(n_no_vertices, 1, 1)
        self._particle_stress = torch.zeros(n_no_vertices, 3, 3, dtype=torch.float32, device=dev)
        self._vertex_force = torch.zeros(n_vertices, 3, dtype=torch.float32, device=dev)
        self._dirty = True
        self._stale = False

    # ---- :376-419
    def continue_from_torch(self, tensor_x: Tensor, tensor_velocity: Optional[Tensor] = None,
                            tensor_d: Optional[Tensor] = None, tensor_C: Optional[Tensor] = None,
                            tensor_R_inv: Optional[Tensor] = None, device="cuda:0", requires_grad=True):
        self._pull()
        if tensor_x is not None:
            self._particle_x = from_torch_safe(tensor_x)
        if tensor_velocity is not None:
            self._particle_v = from_torch_safe(tensor_velocity).clone()
        if tensor_d is not None:
            self._particle_d = from_torch_safe(tensor_d).clone()
        if tensor_C is not None:
            self._particle_C = from_torch_safe(tensor_C).clone()
        if tensor_R_inv is not None:
            self.particle_R_inv = from_torch_safe(tensor_R_inv).clone()
        self._dirty = True

    # ---- :421-432 (finite differences only: nothing to do)
    def set_require_grad(self, requires_grad=True):
        return None

    # ---- :434-467
    def reset_density(self, tensor_density: Tensor, selection_mask: Optional[Tensor] = None, device="cuda:0",
                      requires_grad=True, update_mass=False):
        self.particle_density = tensor_density.detach().to(self.particle_vol.device, torch.float32).contiguous().clone()
        if update_mass:
            self.particle_mass = self.particle_density * self.particle_vol
        self._dirty = True

    # ---- :469-486
    def reset_rest_dir(self, tensor_R_inv: Tensor, device="cuda:0"):
        self.particle_R_inv = from_torch_safe(tensor_R_inv).clone()
        self._dirty = True


class MPMModelStruct:
    """mpm_data_structure.py:610-733"""
    for _n in _MODEL_ARRAYS:
        locals()[_n] = _model_array(_n)
    del _n

    def __init__(self):
        self._dirty = True
        self._solver = None  # the MPMWARP this model was last bound to

    def _pull(self):
        sv = self._solver
        if sv is not None and sv._bound_model is self and sv._bound_state is not None:
            sv._bound_state._pull()

    def init(self, shape, device=None, requires_grad=False) -> None:
        dev = _dev(device)
        z = lambda: torch.zeros(shape, dtype=torch.float32, device=dev)
        self.E, self.nu, self.mu, self.lam = z(), z(), z(), z()
        self.gamma, self.kappa, self.yield_stress = z(), z(), z()
        self.n_particles = int(shape if isinstance(shape, int) else shape[0])
        self._dirty = True

    def finalize_mu_lam(self, n_particles=None, device="cuda:0"):
        # compute_mu_lam_from_E_nu_clean (:870-879)
        self.mu = self.E / (2.0 * (1.0 + self.nu))
        self.lam = self.E * self.nu / ((1.0 + self.nu) * (1.0 - 2.0 * self.nu))
        self._dirty = True

    def init_other_params(self, n_grid=100, grid_lim=1.0, device="cuda:0"):
        import math
        self.grid_lim = grid_lim
        self.n_grid = n_grid
        self.grid_dim_x = self.grid_dim_y = self.grid_dim_z = n_grid
        self.dx, self.inv_dx = self.grid_lim / self.n_grid, float(n_grid / grid_lim)
        self.update_cov_with_F = False
        self.material = 0
        self.plastic_viscosity = 0.0
        self.softening = 0.1
        self.hardening = 0.0
        self.xi = 0.0
        self.friction_angle = 0.0
        sin_phi = math.sin(self.friction_angle / 180.0 * 3.14159265)
        self.friction_coeff = math.tan(self.friction_angle / 180.0 * 3.14159265)
        self.alpha = math.sqrt(2.0 / 3.0) * 2.0 * sin_phi / (3.0 - sin_phi)
        self.gravitational_accelaration = (0.0, 0.0, 0.0)
        self.rpic_damping = 0.0
        self.grid_v_damping_scale = 1.1
        self._dirty = True

    def from_torch(self, tensor_E: Tensor, tensor_nu: Tensor, tensor_gamma: Tensor, tensor_kappa: Tensor,
                   device="cuda:0", requires_grad=False):
        self.E, self.nu = tensor_E.contiguous().float(), tensor_nu.contiguous().float()
        self.gamma, self.kappa = tensor_gamma.contiguous().float(), tensor_kappa.contiguous().float()
        self.finalize_mu_lam()

    def set_require_grad(self, requires_grad=True):
        return None
