"""Spatial-tile sharding of one MPM garment simulation over the GPUs of a box.

The reference is single-GPU; this is the new build's partition logic.  Pure numpy, no
CUDA: it is exercised on CPU by tests/test_sharding_cpu.py (gloo, world_size 2) and drives
mpmavatar_b200.sharded_solver on the GPUs.

Scheme
  * ELEMENTS are sorted along the longest axis of the garment's bounding box (ties: Morton key of their
    grid block, the key the solver sorts by) and cut into `world` contiguous, equally sized runs: rank r
    owns run r -- a slab-shaped spatial tile with a balanced particle count (equal-volume tiles would be
    badly unbalanced for a thin garment) and the shortest boundaries a 1-D cut of a garment offers.
    VERTICES are cut the same way, independently.
  * A rank simulates its owned elements, its owned vertices, and GHOST copies of the vertices its elements
    reference but another rank owns.  A ghost is an ordinary vertex particle with mass 0: its P2G
    contribution is then exactly dt * w * f, the (linear) force term of mpm_utils.py:520-524 for the partial
    vertex force accumulated from this rank's elements, so summing the ranks' grids reproduces the
    single-GPU grid and partial vertex forces need no exchange of their own.  Ghosts run G2P redundantly on
    the reduced grid, so elements find their corners' new positions locally (no second exchange).
  * Per substep ONE sum-reduction: the accumulators (P2G momentum/mass and joint-mover weights) of the grid
    blocks shared by >= 2 ranks.  The shared-block list is rebuilt every `refresh` substeps from each
    rank's active blocks dilated by one block, which covers every block a rank can activate before the next
    rebuild (CFL: <= 0.05 cell per substep, a block is 4 cells).
  * Prescribed-velocity (joint) particles are scattered by their owner only; they stay first in the local
    element / vertex order as the reference's mover kernels require (mpm_solver.py:450-472).
  * TRADITIONAL particles (the demo's sand, run_demo.py:309-379) have no mesh coupling: they are cut into equally
    sized runs along the same axis and need no ghosts.  A rank keeps its share in ascending global order, so the
    pinned TAIL of the global order (joint_traditional_v pins the last n particles, mpm_solver.py:283,446) is the tail
    of every rank's local order too: rank-local count = owned ids >= Nt - n (local_joint_traditional).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

import numpy as np


def _part1by2(x):
    x = x.astype(np.uint32) & 0x3FF
    x = (x | (x << 16)) & 0x30000FF
    x = (x | (x << 8)) & 0x300F00F
    x = (x | (x << 4)) & 0x30C30C3
    x = (x | (x << 2)) & 0x9249249
    return x


def morton_block_keys(x, n_grid, grid_lim):
    """Sort key of csrc/mpm_device.cuh sort_key(): Morton code of the 4^3 block of the stencil base cell."""
    inv_dx = np.float32(np.float64(n_grid) / np.float64(grid_lim))
    base = (x.astype(np.float32) * inv_dx - np.float32(0.5)).astype(np.int32)  # truncation toward zero
    base = np.clip(base, 0, n_grid - 1)
    b = base >> 2
    m = _part1by2(b[:, 0]) | (_part1by2(b[:, 1]) << 1) | (_part1by2(b[:, 2]) << 2)
    return (m.astype(np.uint64) << np.uint64(6)) | ((base[:, 0] & 3) << 4 | (base[:, 1] & 3) << 2 | (base[:, 2] & 3)).astype(np.uint64)


def _equal_cuts(keys, world, coord=None):
    order = np.argsort(keys, kind="stable") if coord is None else np.lexsort((keys, coord))
    owner = np.empty(len(keys), np.int32)
    owner[order] = (np.arange(len(keys), dtype=np.int64) * world // max(len(keys), 1)).astype(np.int32)
    return owner


@dataclass
class Part:
    rank: int
    world: int
    elems: np.ndarray        # global element ids owned by this rank (ascending: joint faces first)
    verts: np.ndarray        # global vertex ids simulated here: owned (ascending) then ghosts
    n_owned_v: int
    faces_local: np.ndarray  # [len(elems),3] indices into `verts`
    num_joint_v: int         # owned joint vertices (a prefix of `verts`)
    num_joint_f: int         # owned joint faces (a prefix of `elems`)
    trads: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.int64))  # owned traditional ids, ascending
    n_traditional_global: int = 0

    @property
    def n_ghost_v(self):
        return len(self.verts) - self.n_owned_v


def partition(x, faces, n_elements, n_vertices, n_grid, grid_lim, world, num_joint_v=0, num_joint_f=0):
    """x: canonical positions [elements | traditional | vertices]; faces: [Ne,3] vertex-local ints."""
    Ne, Nv = n_elements, n_vertices
    Nt = x.shape[0] - Ne - Nv
    if Nt < 0:
        raise ValueError("x is shorter than n_elements + n_vertices")
    xe, xt, xv = x[:Ne], x[Ne:Ne + Nt], x[Ne + Nt:]
    cloth = np.concatenate([xe, xv]) if Ne + Nv else xt
    axis = int(np.argmax(cloth.max(0) - cloth.min(0)))  # slabs across the longest extent of the garment
    e_owner = _equal_cuts(morton_block_keys(xe, n_grid, grid_lim), world, xe[:, axis]) if Ne else np.zeros(0, np.int32)
    v_owner = _equal_cuts(morton_block_keys(xv, n_grid, grid_lim), world, xv[:, axis]) if Nv else np.zeros(0, np.int32)
    t_owner = _equal_cuts(morton_block_keys(xt, n_grid, grid_lim), world, xt[:, axis]) if Nt else np.zeros(0, np.int32)
    parts = []
    for r in range(world):
        elems = np.nonzero(e_owner == r)[0]
        owned = np.nonzero(v_owner == r)[0]
        corners = np.unique(faces[elems].reshape(-1))
        ghosts = np.setdiff1d(corners, owned, assume_unique=True)
        verts = np.concatenate([owned, ghosts])
        g2l = np.full(Nv, -1, np.int64)
        g2l[verts] = np.arange(len(verts))
        parts.append(Part(rank=r, world=world, elems=elems, verts=verts, n_owned_v=len(owned),
                          faces_local=g2l[faces[elems]].astype(np.int64),
                          num_joint_v=int((owned < num_joint_v).sum()), num_joint_f=int((elems < num_joint_f).sum()),
                          trads=np.nonzero(t_owner == r)[0], n_traditional_global=Nt))
    return parts


def local_scene(sc, part: Part):
    """The rank's sub-scene in the canonical layout [owned elements | owned vertices, ghost vertices];
    ghosts get volume 0, hence mass 0 (mass = density * vol, mpm_data_structure.py:434-467)."""
    Ne, Nt = sc.n_elements, sc.n_traditional
    ids = np.concatenate([part.elems, Ne + part.trads, Ne + Nt + part.verts]).astype(np.int64)
    vol = sc.vol[ids].copy()
    vol[len(part.elems) + len(part.trads) + part.n_owned_v:] = 0.0
    pick = lambda a: None if a is None else np.ascontiguousarray(a[ids])
    nnv_ids = np.concatenate([part.elems, Ne + part.trads]).astype(np.int64)  # rows of the [Nnv, ...] arrays
    loc = dataclasses.replace(
        sc, name=f"{sc.name}_r{part.rank}of{part.world}", n_elements=len(part.elems), n_traditional=len(part.trads),
        n_vertices=len(part.verts), x=pick(sc.x), v=pick(sc.v), vol=vol, density=pick(sc.density), E=pick(sc.E),
        nu=pick(sc.nu), gamma=pick(sc.gamma), kappa=pick(sc.kappa), faces=part.faces_local,
        d=np.ascontiguousarray(sc.d[part.elems]), R_inv=np.ascontiguousarray(sc.R_inv[part.elems]),
        F_trial=None if sc.F_trial is None else np.ascontiguousarray(sc.F_trial[nnv_ids]),
        yield_stress=pick(sc.yield_stress), num_joint_v=part.num_joint_v, num_joint_f=part.num_joint_f,
        num_joint_t=local_joint_traditional(part, sc.num_joint_t)[0])
    # the mover must exist on every rank as soon as ANY rank has prescribed-velocity particles (scene_setup.build_from_scene)
    loc.force_mover = bool(sc.num_joint_v or sc.num_joint_f or sc.num_joint_t)
    return loc


def local_joint_traditional(part: Part, n_pinned: int):
    """(count, rows): how many of this rank's traditional particles lie in the pinned tail of the GLOBAL order (the last
    n_pinned traditional particles, mpm_solver.py:283,446) and which rows of the caller's joint_traditional_v they take.
    They are the tail of the rank's local order because `trads` is ascending."""
    first = part.n_traditional_global - int(n_pinned)
    mine = part.trads[part.trads >= first]
    return len(mine), (mine - first).astype(np.int64)


def local_frame_inputs(fi, part: Part):
    """Restrict the per-frame solver inputs to the rank's owned joints; the body mesh is replicated."""
    out = dict(fi)
    if fi.get("joint_traditional_v") is not None:
        out["joint_traditional_v"] = np.ascontiguousarray(fi["joint_traditional_v"][local_joint_traditional(part, len(fi["joint_traditional_v"]))[1]])
    if fi.get("joint_verts_v") is not None:
        out["joint_verts_v"] = np.ascontiguousarray(fi["joint_verts_v"][part.verts[:part.num_joint_v]])
        out["joint_faces_v"] = np.ascontiguousarray(fi["joint_faces_v"][part.elems[:part.num_joint_f]])
    return out


# ---- shared grid blocks ----------------------------------------------------------------------
def pack_coord(b):
    return (b[:, 0] | (b[:, 1] << 10) | (b[:, 2] << 20)).astype(np.int32)


def unpack_coord(c):
    c = np.asarray(c, np.int64)
    return np.stack([c & 1023, (c >> 10) & 1023, (c >> 20) & 1023], 1)


def dilate_blocks(coords, nb):
    """Active blocks plus their 26 neighbours (clipped to the grid), as sorted unique packed coordinates."""
    b = unpack_coord(coords)
    off = np.stack(np.meshgrid([-1, 0, 1], [-1, 0, 1], [-1, 0, 1], indexing="ij"), -1).reshape(-1, 3)
    d = (b[:, None, :] + off[None, :, :]).reshape(-1, 3)
    d = d[((d >= 0) & (d < nb)).all(1)]
    return np.unique(pack_coord(d))


def shared_blocks(dilated_per_rank):
    """Blocks that lie in the dilated active set of at least two ranks (sorted: same list on every rank)."""
    allc = np.concatenate(dilated_per_rank)
    u, cnt = np.unique(allc, return_counts=True)
    return u[cnt >= 2].astype(np.int32)


def blocks_of_particles(x, n_grid, grid_lim):
    """Blocks under the 3^3 stencils of particles at x (what the solver activates), packed and unique."""
    inv_dx = np.float32(np.float64(n_grid) / np.float64(grid_lim))
    base = np.clip((x.astype(np.float32) * inv_dx - np.float32(0.5)).astype(np.int32), 0, n_grid - 1)
    lo = base >> 2
    hi = np.minimum(base + 2, n_grid - 1) >> 2
    out = []
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                blk = np.stack([np.where(a, hi[:, 0], lo[:, 0]), np.where(b, hi[:, 1], lo[:, 1]),
                                np.where(c, hi[:, 2], lo[:, 2])], 1)
                out.append(pack_coord(blk))
    return np.unique(np.concatenate(out))
