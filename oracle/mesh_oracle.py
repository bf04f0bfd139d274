"""CPU restatement (numpy, fp32) of the caller-side glue the CUDA kernels of csrc/mpm_mesh.cuh replace.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).  Pinned by tests/golden/cloth_particles.npz, which
tests/golden/make_mesh_golden.py produces by executing the reference's own source text.
  compute_dir_vol, compute_rest_dir_inv     /root/reference/train_material_params.py:508-515, 533-553
  cloth_normalisation (wld2sim)             /root/reference/train_material_params.py:365-373
  face_frames                               /root/reference/utils/graphics_utils.py:79-107,
                                            /root/reference/scene/mesh_gaussian_model.py:137-146
"""
import numpy as np

f32 = np.float32


def cloth_normalisation(verts):
    v = verts.astype(f32)
    lo, hi = v.min(0), v.max(0)
    scale = f32(1.0) / (hi - lo).max()
    shift = (np.ones(3, f32) - ((lo + hi) / f32(2.0)) * scale).astype(f32)
    return f32(scale), shift


def compute_dir_vol(vertices, faces, thickness):
    v = vertices.astype(f32)
    d1 = v[faces[:, 1]] - v[faces[:, 0]]
    d2 = v[faces[:, 2]] - v[faces[:, 0]]
    c = np.cross(d1, d2).astype(f32)
    cn = np.sqrt((c * c).sum(1, dtype=f32)).astype(f32)
    init_dir = np.stack([d1, d2, c / cn[:, None]], -1).astype(f32)
    R11 = np.sqrt((d1 * d1).sum(1, dtype=f32))
    R12 = (d1 * d2).sum(1, dtype=f32) / R11
    u = d2 - (R12 / R11)[:, None] * d1
    R22 = np.sqrt((u * u).sum(1, dtype=f32))
    rest_dir = np.stack([R11, R12, R22], -1).astype(f32)
    area = f32(0.5) * cn
    element_vol = (f32(0.25) * f32(thickness) * area).astype(f32)
    vertex_vol = np.zeros(v.shape[0], f32)
    np.add.at(vertex_vol, faces.reshape(-1), np.repeat(element_vol, 3))
    return init_dir, rest_dir, element_vol, vertex_vol


def compute_rest_dir_inv(rest_dir):
    R11, R12, R22 = rest_dir[:, 0], rest_dir[:, 1], rest_dir[:, 2]
    iR11, iR22 = f32(1.0) / R11, f32(1.0) / R22
    return np.stack([iR11, -R12 * iR11 * iR22, iR22], -1).astype(f32)


def face_frames(verts, faces):
    v = verts.astype(np.float64)
    v0, v1, v2 = v[faces[:, 0]], v[faces[:, 1]], v[faces[:, 2]]
    length = lambda x: np.sqrt(np.maximum((x * x).sum(-1, keepdims=True), 1e-20))
    sn = lambda x: x / length(x)
    a0 = sn(v1 - v0)
    a1 = sn(np.cross(a0, v2 - v0))
    a2 = -sn(np.cross(a1, a0))
    orien = np.stack([a0, a1, a2], -1)
    scale = (length(v1 - v0) + np.abs((a2 * (v2 - v0)).sum(-1, keepdims=True))) / 2
    center = (v0 + v1 + v2) / 3
    return center, orien, scale


def quat_wxyz_to_rotmat(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                     np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                     np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
