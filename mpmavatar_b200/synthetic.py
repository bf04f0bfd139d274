"""Deterministic synthetic inputs for the MPM hot path.

Plain numpy containers in the reference's canonical particle layout
[elements | traditional | vertices] (train_material_params.py:387).  Used by
tests, bench.py and __graft_entry__.smoke(); there is no dataset or SMPL-X
model offline, so these stand in for the tracked garment / body meshes.

Particle construction follows the reference's own formulas:
  d, R_inv, vol      -> train_material_params.py:508-553 (compute_dir_vol,
                        compute_rest_dir_inv)
  joint ordering     -> first num_joint_v vertices / first num_joint_f faces
                        (preprocess/split_garments.py:72-76)
  physics defaults   -> arguments/__init__.py:81-97
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class Scene:
    name: str
    n_grid: int
    grid_lim: float
    dt: float
    material: str
    n_elements: int
    n_traditional: int
    n_vertices: int
    x: np.ndarray  # [N,3]
    v: np.ndarray  # [N,3]
    vol: np.ndarray  # [N]
    density: np.ndarray  # [N]
    E: np.ndarray
    nu: np.ndarray
    gamma: np.ndarray
    kappa: np.ndarray
    faces: Optional[np.ndarray] = None  # [Ne,3] int, vertex-local
    d: Optional[np.ndarray] = None  # [Ne,3,3]
    R_inv: Optional[np.ndarray] = None  # [Ne,3]
    F_trial: Optional[np.ndarray] = None  # [Nnv,3,3]
    yield_stress: Optional[np.ndarray] = None
    g: tuple = (0.0, -9.8, 0.0)
    friction_angle: float = 40.0
    grid_v_damping_scale: float = 1.1
    rpic_damping: float = 0.0
    substeps_per_frame: int = 400
    # body collider
    body_verts: Optional[np.ndarray] = None
    body_faces: Optional[np.ndarray] = None
    mesh_friction: float = 0.5
    body_motion: Optional[dict] = None
    # joints (particle mover)
    num_joint_v: int = 0
    num_joint_f: int = 0
    num_joint_t: int = 0
    surface_colliders: list = field(default_factory=list)

    @property
    def n_particles(self) -> int:
        return self.n_elements + self.n_traditional + self.n_vertices

    @property
    def n_no_vertices(self) -> int:
        return self.n_elements + self.n_traditional

    # ---- rigid body motion: rotation about y by w*t around a centre oscillating in x
    def _rigid(self, pts0: np.ndarray, t: float) -> np.ndarray:
        m = self.body_motion
        c0 = np.asarray(m["center"], dtype=np.float64)
        ang = m["omega"] * t
        ca, sa = math.cos(ang), math.sin(ang)
        r = pts0.astype(np.float64) - c0
        rot = np.stack([ca * r[:, 0] + sa * r[:, 2], r[:, 1], -sa * r[:, 0] + ca * r[:, 2]], 1)
        c = c0 + np.array([m["amp"] * math.sin(2.0 * math.pi * t / m["period"]), 0.0, 0.0])
        return rot + c

    def frame_inputs(self, i: int) -> dict:
        """Per-frame solver inputs, built like train_material_params.py:617-620:
        mesh_x at the frame start, velocities as finite differences to the next frame."""
        out = dict(mesh_x=None, mesh_v=None, joint_verts_v=None, joint_faces_v=None, joint_traditional_v=None)
        if self.body_verts is None:
            return out
        T = self.dt * self.substeps_per_frame
        x0 = self._rigid(self.body_verts, i * T)
        x1 = self._rigid(self.body_verts, (i + 1) * T)
        out["mesh_x"] = x0.astype(np.float32)
        out["mesh_v"] = ((x1 - x0) / T).astype(np.float32)
        if self.num_joint_v:
            Nnv = self.n_no_vertices
            jv0 = self.x[Nnv:Nnv + self.num_joint_v]
            j0 = self._rigid(jv0, i * T)
            j1 = self._rigid(jv0, (i + 1) * T)
            jv = ((j1 - j0) / T).astype(np.float32)
            out["joint_verts_v"] = jv
            out["joint_faces_v"] = jv[self.faces[: self.num_joint_f]].mean(1).astype(np.float32)
        return out


# --------------------------------------------------------------- cloth helpers
def compute_dir_vol(verts: np.ndarray, faces: np.ndarray, thickness: float = 1e-5):
    """train_material_params.py:533-553 in float32 numpy."""
    v = verts.astype(np.float32)
    d1 = v[faces[:, 1]] - v[faces[:, 0]]
    d2 = v[faces[:, 2]] - v[faces[:, 0]]
    d3 = np.cross(d1, d2)
    d3 = d3 / np.linalg.norm(d3, axis=1, keepdims=True)
    init_dir = np.stack([d1, d2, d3], -1).astype(np.float32)  # columns d1,d2,d3
    R11 = np.linalg.norm(d1, axis=1)
    R12 = (d1 * d2).sum(1) / R11
    R22 = np.linalg.norm(d2 - (R12 / R11)[:, None] * d1, axis=1)
    rest_dir = np.stack([R11, R12, R22], -1).astype(np.float32)
    area = 0.5 * np.linalg.norm(np.cross(d1, d2), axis=1)
    element_vol = (0.25 * thickness * area).astype(np.float32)
    vertex_vol = np.zeros(v.shape[0], np.float32)
    np.add.at(vertex_vol, faces.reshape(-1), np.repeat(element_vol, 3))
    return init_dir, rest_dir, element_vol, vertex_vol


def compute_rest_dir_inv(rest_dir: np.ndarray) -> np.ndarray:
    """train_material_params.py:508-515."""
    R11, R12, R22 = rest_dir[:, 0], rest_dir[:, 1], rest_dir[:, 2]
    iR11 = 1.0 / R11
    iR22 = 1.0 / R22
    iR12 = -R12 * iR11 * iR22
    return np.stack([iR11, iR12, iR22], -1).astype(np.float32)


def tube_mesh(Nu: int, Nr: int, radius: float, height: float, center, rng, noise: float):
    """Open tube around the y axis; ring 0 is the TOP ring so that the joint set
    (top rings) comes first in vertex and face order."""
    u = np.arange(Nu)
    r = np.arange(Nr)
    th = 2.0 * np.pi * u / Nu
    y = center[1] + 0.5 * height - height * r / (Nr - 1)
    X = center[0] + radius * np.cos(th)[None, :].repeat(Nr, 0)
    Z = center[2] + radius * np.sin(th)[None, :].repeat(Nr, 0)
    Y = y[:, None].repeat(Nu, 1)
    verts = np.stack([X, Y, Z], -1).reshape(-1, 3)
    verts = verts + rng.normal(0.0, noise, verts.shape)
    a = (r[:-1, None] * Nu + u[None, :]).reshape(-1)
    b = (r[:-1, None] * Nu + (u[None, :] + 1) % Nu).reshape(-1)
    c = ((r[:-1, None] + 1) * Nu + u[None, :]).reshape(-1)
    dd = ((r[:-1, None] + 1) * Nu + (u[None, :] + 1) % Nu).reshape(-1)
    faces = np.stack([np.stack([a, c, b], 1), np.stack([b, c, dd], 1)], 1).reshape(-1, 3)
    return verts.astype(np.float32), faces.astype(np.int64)


def capsule_mesh(segs: int, rings_cyl: int, rings_cap: int, radius: float, cyl_len: float, center):
    """Closed UV capsule along y: pole, cap rings, cylinder rings, cap rings, pole."""
    lat = []
    for k in range(1, rings_cap + 1):  # top cap down to the cylinder rim
        phi = 0.5 * np.pi * k / rings_cap
        lat.append((radius * np.sin(phi), 0.5 * cyl_len + radius * np.cos(phi)))
    for k in range(1, rings_cyl + 1):  # cylinder down to the bottom rim
        lat.append((radius, 0.5 * cyl_len - cyl_len * k / rings_cyl))
    for k in range(1, rings_cap):  # bottom cap, excluding the pole
        phi = 0.5 * np.pi * (rings_cap - k) / rings_cap
        lat.append((radius * np.sin(phi), -0.5 * cyl_len - radius * np.cos(phi)))
    th = 2.0 * np.pi * np.arange(segs) / segs
    pts = [[0.0, 0.5 * cyl_len + radius, 0.0]]
    for (rr, yy) in lat:
        for t in th:
            pts.append([rr * np.cos(t), yy, rr * np.sin(t)])
    pts.append([0.0, -0.5 * cyl_len - radius, 0.0])
    pts = np.asarray(pts) + np.asarray(center)[None, :]
    nl = len(lat)
    faces = []
    for s in range(segs):
        faces.append([0, 1 + (s + 1) % segs, 1 + s])
    for l in range(nl - 1):
        o0, o1 = 1 + l * segs, 1 + (l + 1) * segs
        for s in range(segs):
            s1 = (s + 1) % segs
            faces.append([o0 + s, o0 + s1, o1 + s])
            faces.append([o0 + s1, o1 + s1, o1 + s])
    last = 1 + nl * segs
    o = 1 + (nl - 1) * segs
    for s in range(segs):
        faces.append([last, o + s, o + (s + 1) % segs])
    return pts.astype(np.float32), np.asarray(faces, dtype=np.int64)


def _cloth_scene(name, seed, Nu, Nr, n_grid, with_body, E=100.0, nu=0.3, gamma=500.0, kappa=500.0, density=1.0,
                 radius=0.25, height=0.8):
    rng = np.random.default_rng(seed)
    center = (1.0, 1.0, 1.0)
    edge = 2.0 * np.pi * radius / Nu
    noise = min(1e-3, 0.15 * edge)
    verts, faces = tube_mesh(Nu, Nr, radius, height, center, rng, noise)
    init_dir, rest_dir, evol, vvol = compute_dir_vol(verts, faces, thickness=1e-5)
    R_inv = compute_rest_dir_inv(rest_dir)
    elts = verts[faces].mean(1).astype(np.float32)
    x = np.concatenate([elts, verts], 0).astype(np.float32)
    Ne, Nv = faces.shape[0], verts.shape[0]
    N = Ne + Nv
    ones = np.ones(N, np.float32)
    sc = Scene(name=name, n_grid=n_grid, grid_lim=2.0, dt=1e-4, material="cloth", n_elements=Ne, n_traditional=0,
               n_vertices=Nv, x=x, v=np.zeros_like(x), vol=np.concatenate([evol, vvol]).astype(np.float32),
               density=ones * density, E=ones * E, nu=ones * nu, gamma=ones * gamma, kappa=ones * kappa,
               faces=faces, d=init_dir, R_inv=R_inv)
    if with_body:
        bv, bf = capsule_mesh(segs=96, rings_cyl=62, rings_cap=24, radius=0.22, cyl_len=0.7, center=center)
        sc.body_verts, sc.body_faces = bv, bf
        sc.body_motion = dict(center=center, omega=1.0, amp=0.08, period=0.8)
        sc.mesh_friction = 0.5
        sc.num_joint_v = 2 * Nu
        sc.num_joint_f = 2 * Nu
    return sc


def scene_c1(n=10_000, n_grid=64, seed=0, material="jelly") -> Scene:
    """C1: traditional particles, fixed-corotated 'jelly', no collider."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.8, 1.2, (n, 3)).astype(np.float32)
    v = rng.normal(0.0, 0.1, (n, 3)).astype(np.float32)
    Ft = (np.eye(3)[None] + 0.05 * rng.normal(0.0, 1.0, (n, 3, 3))).astype(np.float32)
    ones = np.ones(n, np.float32)
    return Scene(name=f"C1_{material}_{n}_{n_grid}", n_grid=n_grid, grid_lim=2.0, dt=1e-4, material=material,
                 n_elements=0, n_traditional=n, n_vertices=0, x=x, v=v, vol=ones * (0.4 ** 3 / n),
                 density=ones, E=ones * 100.0, nu=ones * 0.3, gamma=ones * 500.0, kappa=ones * 500.0, F_trial=Ft,
                 yield_stress=ones * 5.0)


def scene_c2(Nu=256, Nr=130, n_grid=128, seed=1) -> Scene:
    """C2: ~100k cloth particles, 128^3, no body."""
    return _cloth_scene(f"C2_cloth_{Nu}x{Nr}_{n_grid}", seed, Nu, Nr, n_grid, with_body=False)


def scene_c3(Nu=576, Nr=290, n_grid=256, seed=2) -> Scene:
    """C3: ~500k cloth particles, 256^3, capsule body collider + joint rings."""
    return _cloth_scene(f"C3_cloth_body_{Nu}x{Nr}_{n_grid}", seed, Nu, Nr, n_grid, with_body=True)


def scene_c5(Nu=1152, Nr=580, n_grid=512, seed=3) -> Scene:
    """C5: ~2M cloth particles, 512^3, body + joints."""
    return _cloth_scene(f"C5_cloth_body_{Nu}x{Nr}_{n_grid}", seed, Nu, Nr, n_grid, with_body=True)


def scene_small_cloth_body(Nu=48, Nr=26, n_grid=48, seed=5) -> Scene:
    """Small C3-shaped case (cloth + body + joints) a CPU check finishes in seconds."""
    return _cloth_scene(f"small_cloth_body_{Nu}x{Nr}_{n_grid}", seed, Nu, Nr, n_grid, with_body=True)


def scene_demo_like(Nu=40, Nr=20, n_sand=3000, n_grid=48, seed=7) -> Scene:
    """run_demo.py-shaped case: cloth + 'sand' traditional particles pinned by
    joint_traditional_v + sticky floor plane (run_demo.py:309-379, 514-530)."""
    sc = _cloth_scene(f"demo_like_{Nu}x{Nr}_{n_sand}_{n_grid}", seed, Nu, Nr, n_grid, with_body=True)
    rng = np.random.default_rng(seed + 100)
    sand = np.stack([rng.uniform(0.8, 1.2, n_sand), rng.uniform(1.5, 1.56, n_sand), rng.uniform(0.9, 1.1, n_sand)], 1)
    Ne, Nv = sc.n_elements, sc.n_vertices
    svol = np.full(n_sand, 0.4 * 0.06 * 0.2 / n_sand, np.float32)
    sc.x = np.concatenate([sc.x[:Ne], sand.astype(np.float32), sc.x[Ne:]], 0)
    sc.v = np.zeros_like(sc.x)
    sc.vol = np.concatenate([sc.vol[:Ne], svol, sc.vol[Ne:]]).astype(np.float32)
    N = sc.x.shape[0]
    ones = np.ones(N, np.float32)
    sc.density = ones.copy()
    sc.density[Ne:Ne + n_sand] *= 0.1  # run_demo.py:481
    sc.E, sc.nu, sc.gamma, sc.kappa = ones * 100.0, ones * 0.3, ones * 500.0, ones * 500.0
    sc.n_traditional = n_sand
    sc.material = "sand"
    sc.F_trial = None
    sc.num_joint_t = n_sand // 2
    sc.surface_colliders = [dict(point=[0.0, 0.1, 0.0], normal=[0.0, 1.0, 0.0])]
    return sc
