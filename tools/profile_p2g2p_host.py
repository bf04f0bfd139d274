"""Host-side profile of the unchanged caller's per-substep loop (cProfile over 400 p2g2p calls on C3)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
sc = S.scene_c3()
solver, model, state = build_from_scene(sc)
ft = frame_tensors(sc, 0)
mesh_x, mesh_v, jv, jf = ft["mesh_x"], ft["mesh_v"], ft["joint_verts_v"], ft["joint_faces_v"]
def loop(n):
    for k in range(n):
        mx = mesh_x + sc.dt * k * mesh_v
        solver.p2g2p(model, state, sc.dt, mesh_x=mx, mesh_v=mesh_v, joint_traditional_v=None, joint_verts_v=jv, joint_faces_v=jf, device="cuda:0")
loop(100); torch.cuda.synchronize()
t0 = time.perf_counter(); loop(400); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host {1e6*(t1-t0)/400:.1f} us/call, with drain {1e6*(t2-t0)/400:.1f} us/call")
t0 = time.perf_counter()
for k in range(400): mx = mesh_x + sc.dt * k * mesh_v
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"caller's mesh_x arithmetic alone: {1e6*(t1-t0)/400:.1f} us/call")
pr = cProfile.Profile(); pr.enable(); loop(400); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
