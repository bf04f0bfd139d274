"""One MPM garment simulation sharded over the GPUs of a box: one process per GPU (torch.distributed,
NCCL over NVLink), one sum-reduction of the shared grid blocks per substep.

The reference is single-GPU, so this class has no reference counterpart; it drives the same C-ABI
(include/mpm_b200.h, mpm_step_scatter / mpm_shared_* / mpm_step_gather) and the same warp_mpm mirror on each
rank's sub-scene (mpmavatar_b200/sharding.py explains the scheme).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, sharding as sh
from .scene_setup import build_from_scene


class ShardedMPM:
    def __init__(self, sc, device, group=None, refresh=16, margin=1, resort_interval=0):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group, self.sc, self.device = group, sc, torch.device(device)
        self.refresh, self.margin = int(refresh), int(margin)
        self.part = sh.partition(sc.x, sc.faces, sc.n_elements, sc.n_vertices, sc.n_grid, sc.grid_lim, self.world,
                                 sc.num_joint_v, sc.num_joint_f)[self.rank]
        self.local = sh.local_scene(sc, self.part)
        self.solver, self.model, self.state = build_from_scene(self.local, device=device, resort_interval=resort_interval)
        self.solver._bind(self.model, self.state)
        self.lib, self.h = self.solver._libh, self.solver._h
        self.nb = (sc.n_grid + 3) // 4
        self.buf = None
        self.n_shared = 0
        self.k = 0  # substeps since the last shared-list rebuild
        self.stats = {"rebuilds": 0, "shared_blocks": 0, "exchange_bytes": 0}
        # The substep loop runs inside the library (mpm_step_sharded_nccl: captured windows, the shared blocks move with
        # the peer-to-peer exchange fused into the grid update).  NCCL backend: the solver gets its own communicator (the unique id travels
        # through torch.distributed).  Any other backend (gloo: CPU rendezvous, also two ranks SHARING one GPU, which NCCL
        # refuses): the library's set-up traffic goes through a host all-gather callback on this group.
        # MPM_B200_SHARD_GRAPH=0 keeps the older per-substep callback path (mpm_step_sharded).
        nccl = dist.get_backend(group) == "nccl"
        self.in_graph = os.environ.get("MPM_B200_SHARD_GRAPH", "1") != "0"
        if self.in_graph and nccl:
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                raw = (C.c_char * 128)()
                if self.lib.mpm_comm_unique_id(raw) != 0:
                    raise RuntimeError("libmpm_b200: " + self.lib.mpm_last_error(None).decode())
                uid = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
            uid = uid.to(self.device)
            dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self._uid = bytes(uid.cpu().numpy().tobytes())
            self._ck(self.lib.mpm_attach_comm(self.h, self._uid, self.rank, self.world))
        elif self.in_graph:
            self._cb_err = []

            def host_allgather(ctx, send, recv, nbytes):
                try:
                    mine = torch.frombuffer((C.c_ubyte * nbytes).from_address(send), dtype=torch.uint8).clone()
                    out = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.world)]
                    dist.all_gather(out, mine, group=self.group)
                    dst = torch.frombuffer((C.c_ubyte * (nbytes * self.world)).from_address(recv), dtype=torch.uint8)
                    dst.copy_(torch.cat(out))
                    return 0
                except Exception as e:  # noqa: BLE001 -- must not unwind through C
                    self._cb_err.append(e)
                    return 1
            self._host_ag = _lib.HOST_ALLGATHER_FN(host_allgather)  # keep the trampoline alive
            self._ck(self.lib.mpm_attach_host_comm(self.h, self.rank, self.world, self._host_ag, None))

    # collectives: NCCL works on device tensors; with gloo (CPU tests, or two ranks sharing one GPU) they are
    # staged through host memory
    def _host_staged(self):
        return dist.get_backend(self.group) != "nccl"

    def _all_gather(self, t):
        if self._host_staged():
            out = [torch.empty(t.shape, dtype=t.dtype) for _ in range(self.world)]
            dist.all_gather(out, t.cpu(), group=self.group)
            return [o.to(self.device) for o in out]
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return out

    def _all_reduce(self, t):
        if self._host_staged():
            h = t.cpu()
            dist.all_reduce(h, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, group=self.group)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("libmpm_b200: " + self.lib.mpm_last_error(self.h).decode())

    def rebuild_shared(self):
        cap = self.nb ** 3
        coords = np.empty(cap, np.int32)
        n = C.c_int(0)
        self._ck(self.lib.mpm_get_potential_blocks(self.h, self.margin, coords.ctypes.data_as(C.c_void_p), cap, C.byref(n), self._stream()))
        mine = torch.from_numpy(coords[: n.value].copy()).to(self.device)
        sizes = [int(s.item()) for s in self._all_gather(torch.tensor([n.value], dtype=torch.int64, device=self.device))]
        pad = max(sizes)
        send = torch.full((pad,), -1, dtype=torch.int32, device=self.device)
        send[: n.value] = mine
        recv = self._all_gather(send)
        lists = [r[:s].cpu().numpy() for r, s in zip(recv, sizes)]
        shared = np.ascontiguousarray(sh.shared_blocks(lists))
        self._shared_keep = torch.from_numpy(shared).to(self.device)
        self._ck(self.lib.mpm_set_shared_blocks(self.h, C.c_void_p(self._shared_keep.data_ptr()), len(shared), self._stream()))
        self.n_shared = len(shared)
        need = max(self.n_shared, 1) * 64 * 8
        if self.in_graph:
            self.buf = self.buf if self.buf is not None else torch.zeros(1, device=self.device)  # the library owns the exchange buffer
        elif self.buf is None:
            self.buf = torch.zeros(max(int(need * 2), 1 << 20), dtype=torch.float32, device=self.device)
        elif self.buf.numel() < need:  # the C loop holds the buffer's address: it cannot move mid-step
            raise RuntimeError("shared-block buffer too small; construct ShardedMPM with a larger reserve")
        self.k = 0
        self.stats["rebuilds"] += 1
        self.stats["shared_blocks"] = self.n_shared
        self.stats["exchange_bytes"] = self.n_shared * 64 * 8 * 4

    def step(self, dt, nsub, mesh_x=None, mesh_v=None, joint_verts_v=None, joint_faces_v=None, joint_traditional_v=None):
        """nsub substeps; substep k sees body points mesh_x + dt*k*mesh_v (the callers' inner loop,
        train_material_params.py:622-626).  Joint velocity arrays are the GLOBAL ones.  The substep loop
        runs in C (mpm_step_sharded); per substep it calls back once for the all-reduce."""
        dev = self.device
        T = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
        mesh_x, mesh_v = T(mesh_x), T(mesh_v)
        jv = jf = None
        if joint_verts_v is not None and joint_faces_v is not None:
            p = self.part
            if not hasattr(self, "_jidx"):
                self._jidx = (torch.as_tensor(p.verts[:p.num_joint_v], device=dev, dtype=torch.long),
                              torch.as_tensor(p.elems[:p.num_joint_f], device=dev, dtype=torch.long))
            jv = T(joint_verts_v)[self._jidx[0]].contiguous() if p.num_joint_v else torch.zeros(1, 3, device=dev)
            jf = T(joint_faces_v)[self._jidx[1]].contiguous() if p.num_joint_f else torch.zeros(1, 3, device=dev)
        jt, njt = None, 0
        if jv is not None and joint_traditional_v is not None and len(joint_traditional_v):
            # the caller pins the LAST len(joint_traditional_v) traditional particles (mpm_solver.py:283,446): this rank's share
            njt, rows = sh.local_joint_traditional(self.part, len(joint_traditional_v))
            if njt:
                jt = T(joint_traditional_v)[torch.as_tensor(rows, device=dev)].contiguous()
        if self.buf is None and not self.in_graph:
            self.rebuild_shared()
        fin = _lib.MpmFrameInputs()
        fin.mesh_x = None if mesh_x is None else C.c_void_p(mesh_x.data_ptr())
        fin.mesh_v = None if mesh_v is None else C.c_void_p(mesh_v.data_ptr())
        if jv is not None:
            fin.joint_verts_v, fin.joint_faces_v = C.c_void_p(jv.data_ptr()), C.c_void_p(jf.data_ptr())
        if jt is not None:
            fin.joint_traditional_v, fin.n_joint_t = C.c_void_p(jt.data_ptr()), njt
        err = []

        def exchange(ctx, buf, n):
            try:
                self._all_reduce(self.buf[:n])
                return 0
            except Exception as e:  # noqa: BLE001 -- must not unwind through C
                err.append(e)
                return 1

        def rebuild(ctx):
            try:
                self.rebuild_shared()
                return 0
            except Exception as e:  # noqa: BLE001
                err.append(e)
                return 1
        ex, rb = _lib.EXCHANGE_FN(exchange), _lib.REBUILD_FN(rebuild)
        if self.in_graph:
            rc = self.lib.mpm_step_sharded_nccl(self.h, C.c_float(dt), int(nsub), C.byref(fin), self.refresh, self.margin, self._stream())
        else:
            rc = self.lib.mpm_step_sharded(self.h, C.c_float(dt), int(nsub), C.byref(fin), C.c_void_p(self.buf.data_ptr()),
                                           self.refresh, ex, rb, None, self._stream())
        if err or getattr(self, "_cb_err", None):
            raise (err or self._cb_err)[0]
        self._ck(rc)
        if self.in_graph:
            self.refresh_stats()
        self._keep = (mesh_x, mesh_v, jv, jf, jt)
        self.state._stale = True
        self.state._solver = self.solver

    def refresh_stats(self):
        """in-graph path: the shared list lives in the library (synchronises)"""
        n, cap, rb = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.mpm_shared_info(self.h, C.byref(n), C.byref(cap), C.byref(rb), self._stream()))
        self.n_shared = n.value
        mode = {0: "callback", 1: "nccl all-reduce in graph", 2: "peer-to-peer, fused into the grid update, in graph",
                3: "host all-gather per substep"}.get(self.lib.mpm_shared_mode(self.h), "?")
        self.stats.update(rebuilds=rb.value, shared_blocks=n.value, exchange_bytes=cap.value * 64 * 8 * 4, exchange=mode)

    def gather_positions(self):
        """Full canonical particle_x / particle_v on every rank (original particle order)."""
        p, sc, dev = self.part, self.sc, self.device
        n_own = len(p.elems) + len(p.trads) + p.n_owned_v  # local order: [elements | traditional | owned vertices | ghosts]
        x, v = self.state.particle_x, self.state.particle_v
        ids = torch.as_tensor(np.concatenate([p.elems, sc.n_elements + p.trads,
                                              sc.n_elements + sc.n_traditional + p.verts[:p.n_owned_v]]), device=dev, dtype=torch.long)
        own = torch.cat([x[:n_own], v[:n_own]], 1).contiguous()
        sizes = [int(s.item()) for s in self._all_gather(torch.tensor([own.shape[0]], dtype=torch.int64, device=dev))]
        pad = max(sizes)
        send = torch.zeros(pad, 7, device=dev)
        send[: own.shape[0], :6] = own
        send[: own.shape[0], 6] = ids.to(torch.float32)  # ids < 2^24 are exact in fp32
        if sc.n_particles >= (1 << 24):
            raise NotImplementedError("gather_positions packs ids in fp32")
        recv = self._all_gather(send)
        X = torch.empty(sc.n_particles, 3, device=dev)
        V = torch.empty(sc.n_particles, 3, device=dev)
        for r, s in zip(recv, sizes):
            i = r[:s, 6].to(torch.long)
            X[i], V[i] = r[:s, :3], r[:s, 3:6]
        return X, V
