"""Timeline probe of the substep chain (C-ABI mpm_measure_timeline): per-kernel first-start / last-end stamps on the
GPU global timer.  The kernels of a substep overlap under programmatic dependent launch, which CUDA events around
them cannot resolve without serialising them; bench.py takes the P2G / G2P phase lengths for the roofline from here."""
import ctypes as C

import numpy as np
import torch

from . import _lib

NAMES = ["p2g_E", "p2g_T", "p2g_V", "scatter", "grid", "g2p_V", "g2p_T", "g2p_E", "push", "pull"]


def measure(solver, dt, ft, n=32):
    """ft: dict of device tensors mesh_x, mesh_v, joint_verts_v, joint_faces_v (any may be None)."""
    fin = _lib.MpmFrameInputs()
    for k in ("mesh_x", "mesh_v", "joint_verts_v", "joint_faces_v"):
        if ft.get(k) is not None:
            setattr(fin, k, ft[k].data_ptr())
    out = np.zeros((n, 8, 2), np.int64)
    rc = solver._libh.mpm_measure_timeline(solver._h, C.c_float(dt), n, C.byref(fin), out.ctypes.data_as(C.c_void_p),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("libmpm_b200: " + solver._libh.mpm_last_error(solver._h).decode())
    return out


def measure_sharded(sm, dt, ft, n=24):
    """Collective (every rank calls it): per-kernel stamps of n sharded substeps of a ShardedMPM, incl. the two exchange
    phases of the grid update ("push": its flagged stores into the members' receive areas; "pull": its pass over the
    nodes of shared blocks, which waits for the members' parts).  ft: GLOBAL frame tensors (the rank's joint rows are picked as ShardedMPM.step does)."""
    p = sm.part
    fin = _lib.MpmFrameInputs()
    keep = []
    for k in ("mesh_x", "mesh_v"):
        if ft.get(k) is not None:
            setattr(fin, k, ft[k].data_ptr())
    if ft.get("joint_verts_v") is not None and ft.get("joint_faces_v") is not None:
        dev = sm.device
        iv = torch.as_tensor(p.verts[:p.num_joint_v], device=dev, dtype=torch.long)
        jf_i = torch.as_tensor(p.elems[:p.num_joint_f], device=dev, dtype=torch.long)
        jv = ft["joint_verts_v"][iv].contiguous() if p.num_joint_v else torch.zeros(1, 3, device=dev)
        jf = ft["joint_faces_v"][jf_i].contiguous() if p.num_joint_f else torch.zeros(1, 3, device=dev)
        keep += [jv, jf]
        fin.joint_verts_v, fin.joint_faces_v = jv.data_ptr(), jf.data_ptr()
    out = np.zeros((n, 10, 2), np.int64)
    rc = sm.lib.mpm_measure_timeline_sharded(sm.h, C.c_float(dt), n, C.byref(fin), out.ctypes.data_as(C.c_void_p),
                                             C.c_void_p(torch.cuda.current_stream(sm.device).cuda_stream))
    if rc != 0:
        raise RuntimeError("libmpm_b200: " + sm.lib.mpm_last_error(sm.h).decode())
    sm.state._stale = True
    return out


def summarise(tl):
    """median over substeps (first 4 skipped): substep period, per-kernel start offset and duration, and the lengths of
    the P2G phase (first start to last end of kernels 0-2) and of the G2P phase (kernels 5-7)."""
    tl = tl[4:].astype(np.float64)
    ran = tl[0, :, 0] >= 0
    first = int(np.where(ran)[0][0])
    res = {"substep_us": float(np.median(np.diff(tl[:, first, 0]))) / 1e3, "kernels": {}}
    for k in np.where(ran)[0]:
        res["kernels"][NAMES[k]] = {"start_us": round(float(np.median(tl[:, k, 0] - tl[:, first, 0])) / 1e3, 2),
                                    "dur_us": round(float(np.median(tl[:, k, 1] - tl[:, k, 0])) / 1e3, 2)}

    def union(ids):
        ids = [i for i in ids if ran[i]]
        return float(np.median(tl[:, ids, 1].max(1) - tl[:, ids, 0].min(1))) / 1e3 if ids else 0.0
    res["p2g_union_us"] = union([0, 1, 2])
    res["g2p_union_us"] = union([5, 6, 7])
    return res
