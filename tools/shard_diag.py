"""Dev helper (torchrun, N>=2): wall-clock of sharded steps with / without host inputs and exports."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.sharded_solver import ShardedMPM
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
sc = S.scene_c3()
sm = ShardedMPM(sc, dev, refresh=16, margin=1)
keys = ("mesh_x", "mesh_v", "joint_verts_v", "joint_faces_v")
fr = [sc.frame_inputs(i) for i in range(12)]
devf = [{k: torch.as_tensor(f[k], device=dev) for k in keys} for f in fr]
pin = [{k: torch.as_tensor(f[k]).pin_memory() for k in keys} for f in fr]
n_own = len(sm.part.elems) + sm.part.n_owned_v
host_x = torch.empty(n_own, 3).pin_memory()
def run(i, host, export):
    f = pin[i] if host else devf[i]
    if host: f = {k: v.to(dev, non_blocking=True) for k, v in f.items()}
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    sm.step(sc.dt, 400, f["mesh_x"], f["mesh_v"], f["joint_verts_v"], f["joint_faces_v"])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if export: host_x.copy_(sm.state.particle_x[:n_own], non_blocking=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0: print(f"step {i} host={host} export={export}: step {1e3*(t1-t0):.1f} ms, export {1e3*(t2-t1):.1f} ms, launches {sm.solver.stats()['gpu_launches']}", flush=True)
for i in range(3): run(i, False, False)
for i in range(3, 6): run(i, False, True)
for i in range(6, 9): run(i, True, True)
for i in range(9, 12): run(i, False, False)
dist.destroy_process_group()
