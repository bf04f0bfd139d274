"""ctypes binding of libmpm_b200.so (include/mpm_b200.h).  The product path has no CPU
fallback: a missing library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None


class MpmConfig(C.Structure):
    _fields_ = [("n_particles", C.c_int), ("n_elements", C.c_int), ("n_vertices", C.c_int), ("n_grid", C.c_int),
                ("grid_lim", C.c_float), ("n_mesh_v", C.c_int), ("n_mesh_f", C.c_int), ("num_joint_v", C.c_int),
                ("num_joint_f", C.c_int), ("device", C.c_int), ("resort_interval", C.c_int)]


class MpmModelParams(C.Structure):
    _fields_ = [("material", C.c_int), ("hardening", C.c_int), ("friction_coeff", C.c_float), ("alpha", C.c_float),
                ("g", C.c_float * 3), ("rpic_damping", C.c_float), ("grid_v_damping_scale", C.c_float),
                ("xi", C.c_float), ("plastic_viscosity", C.c_float), ("softening", C.c_float)]


_ARRAY_FIELDS = ["x", "v", "C", "F", "F_trial", "stress", "d", "R_inv", "faces", "vertex_force", "vol", "mass",
                 "mu", "lam", "gamma", "kappa", "yield_stress"]


class MpmParticleArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _ARRAY_FIELDS]


class MpmFrameInputs(C.Structure):
    _fields_ = [("mesh_x", C.c_void_p), ("mesh_v", C.c_void_p), ("joint_traditional_v", C.c_void_p),
                ("n_joint_t", C.c_int), ("joint_verts_v", C.c_void_p), ("joint_faces_v", C.c_void_p),
                ("device_inputs", C.c_int)]


class MpmStats(C.Structure):
    _fields_ = [("n_active_blocks", C.c_int), ("n_active_nodes", C.c_longlong), ("n_resorts", C.c_int),
                ("n_substeps", C.c_longlong), ("overflow", C.c_int), ("gpu_launches", C.c_int),
                ("sim_time", C.c_double)]


class MpmProfile(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("stress_ms", "p2g_ms", "collider_scatter_ms", "mover_scatter_ms", "grid_ms",
                                         "g2p_v_ms", "g2p_e_ms", "resort_ms")] + [("n_substeps", C.c_longlong)]


class MpmClothParticles(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x", "vol", "init_dir", "rest_dir", "rest_dir_inv")]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
REBUILD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
HOST_ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)

# every symbol include/mpm_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F3 = C.POINTER(C.c_float)
SYMBOLS = {
    "mpm_create": (C.c_int, [C.POINTER(MpmConfig), C.POINTER(_P)]),
    "mpm_destroy": (None, [_P]),
    "mpm_last_error": (C.c_char_p, [_P]),
    "mpm_set_model": (C.c_int, [_P, C.POINTER(MpmModelParams)]),
    "mpm_import_state": (C.c_int, [_P, C.POINTER(MpmParticleArrays), _P]),
    "mpm_export_state": (C.c_int, [_P, C.POINTER(MpmParticleArrays), _P]),
    "mpm_set_body_mesh": (C.c_int, [_P, _P, _P, _P]),
    "mpm_add_mesh_collider": (C.c_int, [_P, C.c_float]),
    "mpm_add_particle_mover": (C.c_int, [_P]),
    "mpm_add_surface_collider": (C.c_int, [_P, _F3, _F3, C.c_int, C.c_float, C.c_float, C.c_float]),
    "mpm_set_velocity_on_cuboid": (C.c_int, [_P, _F3, _F3, _F3, C.c_float, C.c_float, C.c_int]),
    "mpm_add_bounding_box": (C.c_int, [_P, C.c_float, C.c_float]),
    "mpm_enforce_grid_velocity_by_mask": (C.c_int, [_P, _P, _P]),
    "mpm_add_particle_op": (C.c_int, [_P, C.c_int, _F3, _P, C.c_float, C.c_float, _P]),
    "mpm_add_particle_rotation": (C.c_int, [_P, _F3, _F3, _F3, _F3, C.c_float, C.c_float, _P, C.c_float, C.c_float, _P]),
    "mpm_step": (C.c_int, [_P, C.c_float, C.c_int, C.POINTER(MpmFrameInputs), _P]),
    "mpm_step_scatter": (C.c_int, [_P, C.c_float, C.POINTER(MpmFrameInputs), _P]),
    "mpm_step_gather": (C.c_int, [_P, C.c_float, _P]),
    "mpm_step_sharded": (C.c_int, [_P, C.c_float, C.c_int, C.POINTER(MpmFrameInputs), _P, C.c_int, EXCHANGE_FN, REBUILD_FN, _P, _P]),
    "mpm_comm_unique_id": (C.c_int, [_P]),
    "mpm_attach_comm": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "mpm_attach_host_comm": (C.c_int, [_P, C.c_int, C.c_int, HOST_ALLGATHER_FN, _P]),
    "mpm_step_sharded_nccl": (C.c_int, [_P, C.c_float, C.c_int, C.POINTER(MpmFrameInputs), C.c_int, C.c_int, _P]),
    "mpm_shared_mode": (C.c_int, [_P]),
    "mpm_shared_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "mpm_get_active_blocks": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int), _P]),
    "mpm_get_potential_blocks": (C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(C.c_int), _P]),
    "mpm_set_shared_blocks": (C.c_int, [_P, _P, C.c_int, _P]),
    "mpm_shared_pack": (C.c_int, [_P, _P, _P]),
    "mpm_shared_unpack": (C.c_int, [_P, _P, _P]),
    "mpm_measure_timeline": (C.c_int, [_P, C.c_float, C.c_int, C.POINTER(MpmFrameInputs), _P, _P]),
    "mpm_measure_timeline_sharded": (C.c_int, [_P, C.c_float, C.c_int, C.POINTER(MpmFrameInputs), _P, _P]),
    "mpm_set_time": (C.c_int, [_P, C.c_double]),
    "mpm_export_grid": (C.c_int, [_P, _P, _P, _P, _P]),
    "mpm_set_debug": (C.c_int, [_P, C.c_int]),
    "mpm_set_profiling": (C.c_int, [_P, C.c_int]),
    "mpm_get_profile": (C.c_int, [_P, C.POINTER(MpmProfile)]),
    "mpm_get_stats": (C.c_int, [_P, C.POINTER(MpmStats), _P]),
    "mpm_force_resort": (C.c_int, [_P]),
    "mpm_debug_phase_clocks": (C.c_int, [_P, _P, C.c_int]),
    "mpm_cloth_normalisation": (C.c_int, [_P, C.c_int, _F3, _P]),
    "mpm_build_cloth_particles": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, C.c_float, _F3, C.POINTER(MpmClothParticles), _P]),
    "mpm_export_cloth_verts": (C.c_int, [_P, C.c_float, _F3, _P, _P, _P, _P, _P, _P]),
    "mpm_write_obj": (C.c_int, [C.c_char_p, _P, C.c_int, C.c_char_p, C.c_longlong]),
    "mpm_face_frames": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P]),
}


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load libmpm_b200.so; raises if it is absent and cannot be built."""
    global _lib
    if _lib is None:
        if build_if_missing and _build.needs_build():
            try:
                _build.build_cuda()
            except Exception as e:  # no nvcc on the box and no prebuilt library
                if not os.path.exists(_build.LIB):
                    raise RuntimeError(f"libmpm_b200.so is missing and could not be built: {e}") from e
        lib = C.CDLL(os.environ.get("MPM_B200_LIB", _build.LIB))  # MPM_B200_LIB: analysis builds (tools/)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))
