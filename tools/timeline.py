"""Per-kernel timeline of the substep chain on the GPU global timer (mpmavatar_b200/timeline.py).
usage (GPU box): python tools/timeline.py [scene] [warm substeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
from mpmavatar_b200.timeline import measure, summarise

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sc = getattr(S, "scene_" + name)()
solver, model, state = build_from_scene(sc)
ft = frame_tensors(sc, 0)
solver.step(model, state, sc.dt, warm, ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
r = summarise(measure(solver, sc.dt, ft))
print(f"substep period {r['substep_us']:.1f} us; P2G phase {r['p2g_union_us']:.1f} us, G2P phase {r['g2p_union_us']:.1f} us")
for k, v in r["kernels"].items():
    print(f"  {k:8s} starts at {v['start_us']:7.1f} us  runs {v['dur_us']:6.1f} us  ends {v['start_us'] + v['dur_us']:7.1f}")
