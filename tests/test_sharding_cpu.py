"""CPU checks of the multi-GPU path: partition logic, shared-block lists and -- with two
gloo ranks driving the CPU oracle's phase functions -- the sharding scheme itself (mass-0 ghost vertices,
one sum-reduction of the grid per substep) against the unsharded run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpmavatar_b200 import sharding as sh
from mpmavatar_b200 import synthetic as S


def small_scene():
    sc = S._cloth_scene("shard_test", 11, 24, 12, 24, with_body=False)
    Ne = sc.n_elements
    rng = np.random.default_rng(2)
    verts = sc.x[Ne:] * np.array([1.0, 1.05, 1.0], np.float32) + np.array([0, -0.05, 0], np.float32)
    sc.x = sc.x.copy()
    sc.x[Ne:] = verts
    sc.x[:Ne] = verts[sc.faces].mean(1)
    sc.v = (0.5 * rng.normal(size=sc.x.shape)).astype(np.float32)
    d1 = verts[sc.faces[:, 1]] - verts[sc.faces[:, 0]]
    d2 = verts[sc.faces[:, 2]] - verts[sc.faces[:, 0]]
    d3 = np.cross(d1, d2)
    d3 /= np.linalg.norm(d3, axis=1, keepdims=True)
    sc.d = np.stack([d1, d2, d3 * 0.97], -1).astype(np.float32)
    return sc


def test_partition_covers_everything_once():
    sc = S.scene_small_cloth_body()
    for world in (2, 3, 8):
        parts = sh.partition(sc.x, sc.faces, sc.n_elements, sc.n_vertices, sc.n_grid, sc.grid_lim, world,
                             sc.num_joint_v, sc.num_joint_f)
        assert sorted(np.concatenate([p.elems for p in parts]).tolist()) == list(range(sc.n_elements))
        owned = np.concatenate([p.verts[:p.n_owned_v] for p in parts])
        assert sorted(owned.tolist()) == list(range(sc.n_vertices))
        sizes = [len(p.elems) for p in parts]
        assert max(sizes) - min(sizes) <= 1  # balanced by particle count, not by volume
        assert sum(p.num_joint_v for p in parts) == sc.num_joint_v
        assert sum(p.num_joint_f for p in parts) == sc.num_joint_f
        for p in parts:
            assert (p.faces_local >= 0).all() and (p.faces_local < len(p.verts)).all()
            assert (p.verts[p.faces_local] == sc.faces[p.elems]).all()  # same corners, local numbering
            assert (p.verts[:p.num_joint_v] < sc.num_joint_v).all() and (p.elems[:p.num_joint_f] < sc.num_joint_f).all()
            loc = sh.local_scene(sc, p)
            assert loc.n_particles == len(p.elems) + len(p.verts)
            assert (loc.vol[loc.n_elements + p.n_owned_v:] == 0).all() and (loc.vol[:loc.n_elements + p.n_owned_v] > 0).all()
            # tiles are spatially compact: far fewer ghosts than owned vertices
            assert p.n_ghost_v < 0.5 * p.n_owned_v + 64


def test_partition_with_traditional_particles_and_pinned_tail():
    """The run_demo.py scene shape: cloth + sand, the tail of the sand pinned by joint_traditional_v."""
    sc = S.scene_demo_like()
    Ne, Nt = sc.n_elements, sc.n_traditional
    for world in (2, 3):
        parts = sh.partition(sc.x, sc.faces, Ne, sc.n_vertices, sc.n_grid, sc.grid_lim, world, sc.num_joint_v, sc.num_joint_f)
        assert sorted(np.concatenate([p.trads for p in parts]).tolist()) == list(range(Nt))
        for n_pinned in (sc.num_joint_t, 37, 0):
            jt = np.arange(3 * n_pinned, dtype=np.float32).reshape(n_pinned, 3)
            tot, seen = 0, []
            for p in parts:
                k, rows = sh.local_joint_traditional(p, n_pinned)
                tot += k
                assert (np.diff(p.trads) > 0).all()
                assert (p.trads[len(p.trads) - k:] >= Nt - n_pinned).all() and (p.trads[: len(p.trads) - k] < Nt - n_pinned).all()
                seen.append(Nt - n_pinned + rows)  # global ids of the rows this rank takes
                if n_pinned:
                    lfi = sh.local_frame_inputs(dict(joint_traditional_v=jt), p)
                    assert (lfi["joint_traditional_v"] == jt[rows]).all()
            assert tot == n_pinned and sorted(np.concatenate(seen).tolist()) == list(range(Nt - n_pinned, Nt))
        loc = sh.local_scene(sc, parts[0])
        p = parts[0]
        assert loc.n_traditional == len(p.trads) and loc.n_particles == len(p.elems) + len(p.trads) + len(p.verts)
        assert (loc.x[loc.n_elements:loc.n_elements + loc.n_traditional] == sc.x[Ne + p.trads]).all()
        assert loc.num_joint_t == sh.local_joint_traditional(p, sc.num_joint_t)[0]


def test_morton_key_matches_block_order():
    x = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 1.01], [0.2, 1.9, 0.3]], np.float32)
    k = sh.morton_block_keys(x, 64, 2.0)
    assert k[0] != k[2] and (k[0] >> np.uint64(6)) == (k[1] >> np.uint64(6))


def test_shared_blocks_are_a_superset_that_survives_motion():
    sc = S._cloth_scene("shard_blocks", 1, 128, 64, 128, with_body=False)
    nb = (sc.n_grid + 3) // 4
    parts = sh.partition(sc.x, sc.faces, sc.n_elements, sc.n_vertices, sc.n_grid, sc.grid_lim, 2)
    def touched(x):
        out = []
        for p in parts:
            ids = np.concatenate([p.elems, sc.n_elements + p.verts])
            out.append(sh.blocks_of_particles(x[ids], sc.n_grid, sc.grid_lim))
        return out
    act = touched(sc.x)
    dil = [sh.dilate_blocks(a, nb) for a in act]
    shared = sh.shared_blocks(dil)
    dx = sc.grid_lim / sc.n_grid
    rng = np.random.default_rng(0)
    for step in (0.0, 1.5, 3.9):  # particles may drift up to (just under) one block between rebuilds
        moved = sc.x + rng.uniform(-step * dx, step * dx, sc.x.shape).astype(np.float32)
        t = touched(np.clip(moved, 2 * dx, sc.grid_lim - 2 * dx))
        u, cnt = np.unique(np.concatenate(t), return_counts=True)
        assert np.isin(u[cnt >= 2], shared).all()
    assert len(shared) < 0.5 * len(np.unique(np.concatenate(dil)))  # and it is a boundary layer, not the whole grid


def _worker(rank, world, port, nsub, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleSim
    sc = small_scene()
    part = sh.partition(sc.x, sc.faces, sc.n_elements, sc.n_vertices, sc.n_grid, sc.grid_lim, world)[rank]
    loc = sh.local_scene(sc, part)
    o = OracleSim.from_scene(loc, "f64", threads=1)
    R = o.real
    for _ in range(nsub):
        o.call("orc_zero_grid")
        o.vertex_force[:] = 0
        o.call("orc_compute_stress_from_F_trial", R(sc.dt))
        o.call("orc_p2g_apic_with_stress", R(sc.dt))
        for a in (o.grid_m, o.grid_v_in):  # the one exchange of the substep (dense here, shared blocks on GPU)
            t = torch.from_numpy(a)
            dist.all_reduce(t)
        o.call("orc_grid_normalization_and_gravity", R(sc.dt))
        o.call("orc_g2p_v", R(sc.dt))
        o.call("orc_g2p_e", R(sc.dt))
    Ne_l = loc.n_elements
    q.put((rank, part.elems, part.verts, part.n_owned_v, o.x.copy(), o.v.copy(), o.d.copy(), Ne_l))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_substeps_equal_unsharded_oracle():
    from oracle.oracle import OracleSim
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    nsub, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nsub, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = small_scene()
    full = OracleSim.from_scene(sc, "f64", threads=1)
    for _ in range(nsub):
        full.p2g2p(sc.dt)
    Ne = sc.n_elements
    x = np.full_like(full.x, np.nan)
    v = np.full_like(full.v, np.nan)
    for rank, elems, verts, n_owned, lx, lv, ld, Ne_l in res:
        x[elems], v[elems] = lx[:Ne_l], lv[:Ne_l]
        x[Ne + verts[:n_owned]] = lx[Ne_l:Ne_l + n_owned]
        v[Ne + verts[:n_owned]] = lv[Ne_l:Ne_l + n_owned]
        # ghost copies stay equal to the owner's result
        assert np.abs(lx[Ne_l + n_owned:] - full.x[Ne + verts[n_owned:]]).max() < 1e-12
        assert np.abs(ld - full.d[elems]).max() < 1e-10
    assert not np.isnan(x).any()
    assert np.abs(x - full.x).max() < 1e-12
    assert np.abs(v - full.v).max() / np.abs(full.v).max() < 1e-10
