#!/usr/bin/env python
"""Golden vectors for the caller-side glue, produced by EXECUTING THE REFERENCE'S OWN SOURCE TEXT:
the methods compute_dir_vol / compute_rest_dir_inv / compute_rest_dir_inv_from_vf and the wld2sim lines of
setup_simulation are cut out of /root/reference/train_material_params.py (the module itself cannot be imported offline:
dataset, SMPL-X, diff_gauss), compute_face_orientation & co. out of utils/graphics_utils.py, and run on a small synthetic
garment with torch on the CPU (`.cuda()` is a no-op here).  Output: tests/golden/cloth_particles.npz.

    python tests/golden/make_mesh_golden.py
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mpmavatar_b200 import synthetic as S  # noqa: E402

TRAINER = "/root/reference/train_material_params.py"
GRAPHICS = "/root/reference/utils/graphics_utils.py"


def cut_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for n in ast.walk(tree):
        if isinstance(n, ast.FunctionDef) and n.name in names:
            out[n.name] = textwrap.dedent(ast.get_source_segment(src, n))
    return out


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference allocates with .cuda(); this container has no GPU
    ns = {"torch": torch}
    for name, code in cut_functions(TRAINER, {"compute_dir_vol", "compute_rest_dir_inv", "compute_rest_dir_inv_from_vf"}).items():
        exec(code, ns)
    gns = {"torch": torch}
    for name, code in cut_functions(GRAPHICS, {"dot", "length", "safe_normalize", "compute_face_orientation"}).items():
        exec(code, gns)
    # the wld2sim lines of setup_simulation, verbatim (train_material_params.py:365-373)
    lines = open(TRAINER).read().splitlines()
    seg = textwrap.dedent("\n".join(lines[364:373]))
    assert "max_diff" in seg and "self.sim2wld" in seg, seg

    rng = np.random.default_rng(17)
    verts_sim0, faces = S.tube_mesh(18, 9, 0.25, 0.8, (1.0, 1.0, 1.0), rng, 2e-3)
    verts_wld = (verts_sim0 * 1.7 + np.array([0.3, -0.9, 2.1])).astype(np.float32)  # some world frame
    me = types.SimpleNamespace()
    env = {"torch": torch, "self": me, "verts": torch.from_numpy(verts_wld)}
    exec(seg, env)
    scale, shift = float(me.scale), me.shift.numpy()
    v_sim = me.wld2sim(torch.from_numpy(verts_wld))
    f = torch.from_numpy(faces)
    init_dir, rest_dir, evol, vvol = ns["compute_dir_vol"](me, v_sim, f, 1e-5)
    rinv = ns["compute_rest_dir_inv"](me, rest_dir)
    rinv_vf = ns["compute_rest_dir_inv_from_vf"](me, v_sim, f)
    elts = v_sim[f].mean(1)  # train_material_params.py:379
    # export side: sim2wld of moved vertices, scattered into the full-body array (:812-817), MSE (:631)
    moved = v_sim + 0.01 * torch.from_numpy(rng.normal(size=v_sim.shape).astype(np.float32))
    cloth_wld = me.sim2wld(moved)
    target = cloth_wld + 0.003 * torch.from_numpy(rng.normal(size=v_sim.shape).astype(np.float32))
    mse = torch.nn.functional.mse_loss(cloth_wld, target)
    orien, fscale = gns["compute_face_orientation"](cloth_wld, f, return_scale=True)
    center = cloth_wld[f].mean(dim=-2)
    path = os.path.join(HERE, "cloth_particles.npz")
    np.savez_compressed(path, verts_wld=verts_wld, faces=faces.astype(np.int32), thickness=np.float32(1e-5),
                        scale=np.float32(scale), shift=shift.astype(np.float32), verts_sim=v_sim.numpy(), elts=elts.numpy(),
                        init_dir=init_dir.numpy(), rest_dir=rest_dir.numpy(), element_vol=evol.numpy(), vertex_vol=vvol.numpy(),
                        rest_dir_inv=rinv.numpy(), rest_dir_inv_from_vf=rinv_vf.numpy(), moved_sim=moved.numpy(),
                        cloth_wld=cloth_wld.numpy(), target=target.numpy(), mse=np.float32(mse),
                        face_center=center.numpy(), face_orien=orien.numpy(), face_scale=fscale.numpy())
    print(f"cloth_particles: Nv={verts_wld.shape[0]} Ne={faces.shape[0]} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
