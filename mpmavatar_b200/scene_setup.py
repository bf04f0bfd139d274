"""Builds (solver, model, state) from a synthetic Scene with the exact call sequence of the
reference's caller (train_material_params.py:403-506 setup_simulation, :589-610 rollout reset)."""
from __future__ import annotations

import numpy as np
import torch

from .warp_mpm.mpm_data_structure import MPMModelStruct, MPMStateStruct
from .warp_mpm.mpm_solver import MPMWARP


def build_from_scene(sc, device="cuda:0", resort_interval=0):
    dev = torch.device(device)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    N, Ne, Nv, Nt = sc.n_particles, sc.n_elements, sc.n_vertices, sc.n_traditional
    state = MPMStateStruct()
    state.init(N, Ne, Nv, device=device, requires_grad=True)
    trad = np.zeros(N, np.int32); trad[Ne:Ne + Nt] = 1
    vert = np.zeros(N, np.int32); vert[Ne + Nt:] = 1
    elem = np.zeros(N, np.int32); elem[:Ne] = 1
    d = T(sc.d) if Ne else torch.zeros(0, 3, 3, device=dev)
    R_inv = T(sc.R_inv) if Ne else torch.zeros(0, 3, device=dev)
    faces = T(sc.faces.astype(np.float32)) if Ne else torch.zeros(0, 3, device=dev)
    D_inv = torch.linalg.inv(d) if Ne else d
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        state.from_torch(T(sc.x), T(sc.vol), D_inv, R_inv, faces, trad, vert, elem,
                         torch.zeros(N - Nv, 6), device=device, requires_grad=True, n_grid=sc.n_grid,
                         grid_lim=sc.grid_lim)
    model = MPMModelStruct()
    model.init(N, device=device, requires_grad=True)
    model.init_other_params(n_grid=sc.n_grid, grid_lim=sc.grid_lim, device=device)
    solver = MPMWARP(N, Ne, Nv, n_grid=sc.n_grid, grid_lim=sc.grid_lim, mesh_vertices=sc.body_verts,
                     mesh_faces=sc.body_faces, num_joint_t=sc.num_joint_t, num_joint_v=sc.num_joint_v,
                     num_joint_f=sc.num_joint_f, device=device, resort_interval=resort_interval)
    solver.set_parameters_dict(model, state, {"material": sc.material, "g": list(sc.g), "density": 1.0,
                                              "grid_v_damping_scale": sc.grid_v_damping_scale,
                                              "friction_angle": sc.friction_angle, "rpic_damping": sc.rpic_damping})
    if sc.yield_stress is not None:
        model.yield_stress = T(sc.yield_stress)
    for b in sc.surface_colliders:
        solver.add_surface_collider(**b)
    if sc.body_verts is not None:
        solver.add_mesh_collider(solver.mesh.id, n_grid=sc.n_grid, friction=sc.mesh_friction)
    # (a rank of a sharded run needs the mover even without joint particles of its own: its grid blocks receive the
    # prescribed-velocity sums of the other ranks)
    if sc.num_joint_v or sc.num_joint_f or sc.num_joint_t or getattr(sc, "force_mover", False):
        solver.add_particle_mover(n_grid=sc.n_grid)
    reset_rollout(sc, solver, model, state, device)
    return solver, model, state


def reset_rollout(sc, solver, model, state, device="cuda:0"):
    """train_material_params.py:584-610."""
    dev = torch.device(device)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    Ne = sc.n_elements
    d = T(sc.d) if Ne else torch.zeros(0, 3, 3, device=dev)
    R_inv = T(sc.R_inv) if Ne else torch.zeros(0, 3, device=dev)
    state.reset_state(sc.n_vertices, T(sc.x).clone(), d, None, T(sc.v).clone(), tensor_R_inv=R_inv, device=device,
                      requires_grad=True)
    if sc.F_trial is not None:
        state.particle_F_trial = T(sc.F_trial)
    state.reset_density(T(sc.density), None, device, update_mass=True)
    solver.set_E_nu_from_torch(model, T(sc.E), T(sc.nu), T(sc.gamma), T(sc.kappa), device)
    solver.prepare_mu_lam(model, state, device)
    solver.time = 0.0


def frame_tensors(sc, i, device="cuda:0"):
    fi = sc.frame_inputs(i)
    dev = torch.device(device)
    return {k: (None if v is None else torch.as_tensor(v, dtype=torch.float32, device=dev)) for k, v in fi.items()}
