"""Dev helper: time the C3 scene and print the per-phase profile (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 400
sc = getattr(S, "scene_" + name)()
solver, model, state = build_from_scene(sc)
ft = frame_tensors(sc, 0)
args = (ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
solver.step(model, state, sc.dt, 50, *args)
# optional third argument: whole frames (400 substeps each, body moving) to advance first -- bench.py times the rollout
# after 3 + k frames, when the garment is falling and has drifted off its sort order
nframes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for i in range(nframes):
    fi = frame_tensors(sc, i)
    solver.step(model, state, sc.dt, 400, fi["mesh_x"], fi["mesh_v"], None, fi["joint_verts_v"], fi["joint_faces_v"])
if nframes:
    ft = frame_tensors(sc, nframes)
    args = (ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
torch.cuda.synchronize()
print("stats", solver.stats())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
solver.step(model, state, sc.dt, nsub, *args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{name}: {nsub} substeps in {ms:.2f} ms -> {ms/nsub*1000:.1f} us/substep, {nsub/ms*1000:.0f} substeps/s")
solver.enable_profiling(True)
solver.step(model, state, sc.dt, 100, *args)
p = solver.get_profile()
n = p["n_substeps"]
print({k: round(v / n * 1000, 2) for k, v in p.items() if k.endswith("_ms")}, "us/substep")
solver.enable_profiling(False)
print("stats", solver.stats())
print("x finite", bool(torch.isfinite(state.particle_x).all()))
