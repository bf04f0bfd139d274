"""Aggregate substeps/s of K concurrent C3 rollouts on one GPU (mpmavatar_b200/probes.py).
usage (GPU box): python tools/probe_time.py [K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.probes import ProbeBatch
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors

sc = S.scene_c3()
ft = frame_tensors(sc, 0)
args = (ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
for K in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    batch = ProbeBatch([build_from_scene(sc) for _ in range(K)])
    batch.step(sc.dt, 64, *args)
    batch.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    batch.step(sc.dt, 400, *args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"K={K}: 400 substeps of each probe in {ms:.1f} ms -> {K * 400 / ms * 1e3:.0f} substeps/s aggregate, "
          f"{400 / ms * 1e3:.0f} per probe", flush=True)
    del batch
    torch.cuda.empty_cache()
