"""Caller-side glue around the hot path, on the device: what the reference's Trainer does
with torch ops and Python loops before and after the substep loop, as CUDA kernels behind the C-ABI
(include/mpm_b200.h, csrc/mpm_mesh.cuh).  Method names and return values follow the reference:

  compute_dir_vol / compute_rest_dir_inv / compute_rest_dir_inv_from_vf   train_material_params.py:508-553
  cloth_normalisation (wld2sim scale / shift)                            train_material_params.py:365-373
  build_cloth_particles (the arrays setup_simulation hands to from_torch) train_material_params.py:375-395
  export_cloth_verts (un-permute + sim2wld [+ scatter, + MSE])            train_material_params.py:628-631, 811-817
  write_obj (per-frame OBJ text)                                          train_material_params.py:819-821
  face_frames (MeshGaussianModel.set_mesh_by_verts)                       scene/mesh_gaussian_model.py:137-146
  load_split_idx (on-disk garment split)                                  preprocess/split_garments.py:84-94

No CPU fallback: the kernels need the CUDA library (write_obj is host code in the same library)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

SPLIT_IDX_KEYS = ("num_joint_v", "num_joint_f", "reordered_cloth_v_idx", "reordered_cloth_f_idx", "reordered_human_v_idx",
                  "reordered_human_f_idx", "new_cloth_faces", "new_human_faces")


def _ck(rc, h=None):
    if rc != 0:
        raise RuntimeError("libmpm_b200: " + _lib.load().mpm_last_error(h).decode())


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _cuda_f32(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{what}: a CUDA tensor is required (mpmavatar_b200 has no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def _faces_i32(faces, dev):
    return torch.as_tensor(faces, device=dev).to(torch.int32).contiguous()


def load_split_idx(path):
    """split_idx.npz written by preprocess/split_garments.py:84-94: joint counts, the cloth / human vertex and face
    re-orderings (joint vertices / faces first) and the re-indexed face lists."""
    z = np.load(path)
    missing = [k for k in SPLIT_IDX_KEYS if k not in z.files]
    if missing:
        raise KeyError(f"{path}: not a split_idx.npz, missing {missing}")
    out = {k: z[k] for k in SPLIT_IDX_KEYS}
    out["num_joint_v"], out["num_joint_f"] = int(out["num_joint_v"]), int(out["num_joint_f"])
    return out


def cloth_normalisation(verts_wld):
    """(scale, shift[1,3]) of setup_simulation: the garment's bounding box becomes unit-sized and centred at (1,1,1)."""
    v = _cuda_f32(verts_wld, "cloth_normalisation")
    out = (C.c_float * 4)()
    _ck(_lib.load().mpm_cloth_normalisation(_p(v), int(v.shape[0]), out, _stream(v.device)))
    return float(out[0]), torch.tensor([[out[1], out[2], out[3]]], dtype=torch.float32, device=v.device)


def _build(verts_wld, faces, thickness, scale, shift, want=("x", "vol", "init_dir", "rest_dir", "rest_dir_inv")):
    v = _cuda_f32(verts_wld, "build_cloth_particles")
    dev = v.device
    f = _faces_i32(faces, dev)
    Nv, Ne = int(v.shape[0]), int(f.shape[0])
    z = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    bufs = {"x": z(Ne + Nv, 3), "vol": z(Ne + Nv), "init_dir": z(Ne, 3, 3), "rest_dir": z(Ne, 3), "rest_dir_inv": z(Ne, 3)}
    st = _lib.MpmClothParticles()
    for k in bufs:
        if k in want or k == "x":
            setattr(st, k, bufs[k].data_ptr())
    sh = torch.as_tensor(shift, dtype=torch.float32).reshape(-1).tolist()
    _ck(_lib.load().mpm_build_cloth_particles(_p(v), _p(f), Nv, Ne, C.c_float(thickness), C.c_float(scale), _lib.f3(sh),
                                              C.byref(st), _stream(dev)))
    return bufs, Ne, Nv


def compute_dir_vol(vertices, faces, thickness):
    """train_material_params.py:533-553: (init_dir [Ne,3,3], rest_dir [Ne,3], element_vol [Ne], vertex_vol [Nv]) of
    vertices that are ALREADY in sim space."""
    b, Ne, Nv = _build(vertices, faces, thickness, 1.0, (0.0, 0.0, 0.0), want=("vol", "init_dir", "rest_dir"))
    return b["init_dir"], b["rest_dir"], b["vol"][:Ne], b["vol"][Ne:]


def compute_rest_dir_inv(rest_dir):
    """train_material_params.py:508-515 (three elementwise torch ops; no kernel needed)."""
    R11, R12, R22 = rest_dir[:, 0], rest_dir[:, 1], rest_dir[:, 2]
    iR11, iR22 = 1.0 / R11, 1.0 / R22
    return torch.stack([iR11, -R12 * iR11 * iR22, iR22], -1)


def compute_rest_dir_inv_from_vf(vertices, faces):
    """train_material_params.py:517-531."""
    b, _, _ = _build(vertices, faces, 0.0, 1.0, (0.0, 0.0, 0.0), want=("rest_dir_inv",))
    return b["rest_dir_inv"]


def build_cloth_particles(verts_wld, faces, thickness=1e-5):
    """Everything setup_simulation derives from the first tracked frame (train_material_params.py:360-395), in one pass on
    the device.  Returns a dict: scale, shift, x [Ne+Nv,3] ([elements | vertices]), vol, init_dir, rest_dir, rest_dir_inv,
    n_elements, n_vertices -- the arguments of MPMStateStruct.from_torch / reset_state."""
    scale, shift = cloth_normalisation(verts_wld)
    b, Ne, Nv = _build(verts_wld, faces, thickness, scale, shift)
    b.update(scale=scale, shift=shift, n_elements=Ne, n_vertices=Nv)
    return b


def export_cloth_verts(solver, scale, shift, *, out=None, scatter_idx=None, full_verts=None, target=None):
    """Cloth vertices of the bound state in ORIGINAL vertex order and WORLD coordinates, read from the solver's sorted
    records in one kernel (un-permute + sim2wld fused; train_material_params.py:628-630, 811-812).  scatter_idx /
    full_verts: also write them into a full-body vertex array (:814).  target: return F.mse_loss(cloth_verts, target) (:631)
    as a 0-d tensor.  Returns (cloth_verts, mse or None)."""
    dev = solver.device
    Nv = solver.n_vertices
    if out is None:
        out = torch.empty(Nv, 3, dtype=torch.float32, device=dev)
    sidx = None if scatter_idx is None else torch.as_tensor(scatter_idx, device=dev).to(torch.int64).contiguous()
    tgt = None if target is None else _cuda_f32(target, "export_cloth_verts")
    sse = None if target is None else torch.zeros((), dtype=torch.float64, device=dev)
    if full_verts is not None and not (full_verts.is_cuda and full_verts.dtype == torch.float32 and full_verts.is_contiguous()):
        raise RuntimeError("full_verts must be a contiguous fp32 CUDA tensor (it is written in place)")
    sh = torch.as_tensor(shift, dtype=torch.float32).reshape(-1).tolist()
    lib = solver._libh
    rc = lib.mpm_export_cloth_verts(solver._h, C.c_float(scale), _lib.f3(sh), _p(out), _p(sidx), _p(full_verts), _p(tgt), _p(sse),
                                    _stream(dev))
    _ck(rc, solver._h)
    solver._keep_export = (sidx, tgt)
    return out, (None if sse is None else (sse / (3 * Nv)).to(torch.float32))


def write_obj(path, verts, tail_lines=None):
    """train_material_params.py:819-821: 'v x y z' per vertex, then the vt / f lines verbatim.  verts: [n,3] tensor (any
    device) or array; every number is the shortest decimal that reads back as the same float32."""
    if isinstance(verts, torch.Tensor):
        verts = verts.detach().to("cpu", torch.float32).contiguous().numpy()
    v = np.ascontiguousarray(verts, dtype=np.float32)
    tail = "".join(tail_lines).encode() if tail_lines else b""
    _ck(_lib.load().mpm_write_obj(str(path).encode(), v.ctypes.data_as(C.c_void_p), int(v.shape[0]), tail, len(tail)))


def face_frames(verts, faces):
    """MeshGaussianModel.set_mesh_by_verts (scene/mesh_gaussian_model.py:137-146): (face_center [F,3], face_orien_mat
    [F,3,3], face_orien_quat wxyz [F,4], face_scaling [F,1]) of the simulated mesh, computed on the device."""
    v = _cuda_f32(verts, "face_frames")
    f = _faces_i32(faces, v.device)
    F = int(f.shape[0])
    z = lambda *s: torch.empty(*s, dtype=torch.float32, device=v.device)
    c, o, q, s = z(F, 3), z(F, 3, 3), z(F, 4), z(F, 1)
    _ck(_lib.load().mpm_face_frames(_p(v), _p(f), F, _p(c), _p(o), _p(q), _p(s), _stream(v.device)))
    return c, o, q, s
