"""Latency analysis: per-warp phase cycles of the particle kernels (needs the -DMPM_CLK build).
usage (GPU box): python tools/phase_clocks.py [scene] [nsub]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = os.path.join(ROOT, "mpmavatar_b200", "libmpm_b200_clk.so")
if not os.path.exists(lib) or "--build" in sys.argv:
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler",
                    "-fPIC", "-DMPM_CLK", "-o", lib, os.path.join(ROOT, "mpmavatar_b200", "csrc", "mpm_b200.cu")], check=True)
    if "--build" in sys.argv:
        sys.exit(0)
os.environ["MPM_B200_LIB"] = lib
import torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "c3"
nsub = int(args[1]) if len(args) > 1 else 64
sc = getattr(S, "scene_" + name)()
solver, model, state = build_from_scene(sc)
ft = frame_tensors(sc, 0)
a = (ft["mesh_x"], ft["mesh_v"], None, ft["joint_verts_v"], ft["joint_faces_v"])
solver.step(model, state, sc.dt, 32, *a)
buf = (C.c_ulonglong * 64)()
solver._libh.mpm_debug_phase_clocks(solver._h, buf, 1)
solver.step(model, state, sc.dt, nsub, *a)
solver._libh.mpm_debug_phase_clocks(solver._h, buf, 0)
names = {0: ("p2g_elements", ["slab load", "unpack+stress", "stage1", "stage2 rest", "s2 lookup", "s2 run sum", "s2 flush"]),
         1: ("p2g_traditional", ["slab load", "unpack", "stage1", "stage2 rest", "s2 lookup", "s2 run sum", "s2 flush"]),
         2: ("p2g_vertices", ["slab load", "unpack", "stage1", "stage2 rest", "s2 lookup", "s2 run sum", "s2 flush"]),
         3: ("g2p_vertices", ["slab load", "runs", "stage tile", "contract", "epilogue", "slab store"]),
         5: ("g2p_elements", ["slab load", "corners issue+runs", "stage tile", "corner consume", "contract", "epilogue+store"])}
for k, (nm, ph) in names.items():
    n = buf[k * 8 + 7]
    if not n:
        continue
    tot = sum(buf[k * 8 + i] for i in range(len(ph)))
    print(f"{nm}: {n} warps, {tot / n:.0f} cycles/warp")
    for i, p in enumerate(ph):
        print(f"    {p:22s} {buf[k * 8 + i] / n:8.0f} cycles  {buf[k * 8 + i] / tot * 100:5.1f}%")
