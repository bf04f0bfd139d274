"""Loads tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own source) back
into synthetic.Scene containers."""
import glob
import os

import numpy as np

from mpmavatar_b200 import synthetic as S

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    # cov_export.npz (tests/test_cov_export.py) and cloth_particles.npz (tests/test_cloth_io.py) are not scenes
    return [n for n in names if not n.startswith("cov_") and n != "cloth_particles"]


class FixedScene(S.Scene):
    """Scene whose per-frame inputs are the recorded ones."""
    _fi = None
    particle_ops = False
    grid_bcs = False
    material_variant = None

    def frame_inputs(self, i):
        assert i == 0
        return dict(self._fi)


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    sc_kw = {k[3:]: z[k][()] for k in z.files if k.startswith("sc_")}
    arr = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    for k in ("n_grid", "n_elements", "n_traditional", "n_vertices", "substeps_per_frame", "num_joint_v", "num_joint_f",
              "num_joint_t"):
        sc_kw[k] = int(sc_kw[k])
    for k in ("grid_lim", "dt", "friction_angle", "grid_v_damping_scale", "rpic_damping", "mesh_friction"):
        sc_kw[k] = float(sc_kw[k])
    for k in ("name", "material"):
        sc_kw[k] = str(sc_kw[k])
    sc = FixedScene(g=tuple(float(v) for v in z["g"]), **sc_kw, **arr)
    sc._fi = {k: (z["fi_" + k] if "fi_" + k in z.files else None)
              for k in ("mesh_x", "mesh_v", "joint_verts_v", "joint_faces_v", "joint_traditional_v")}
    if "plane_point" in z.files:
        sc.surface_colliders = [dict(point=list(z["plane_point"]), normal=list(z["plane_normal"]))]
    sc.particle_ops = "particle_ops" in z.files  # replay tests/golden/make_golden.py PARTICLE_OPS after the setup
    sc.grid_bcs = "grid_bcs" in z.files          # replay GRID_BCS
    sc.material_variant = name if "material_variant" in z.files else None  # replay MATERIAL_VARIANTS[name]
    ref64 = {k[6:]: z[k] for k in z.files if k.startswith("ref64_")}
    ref32 = {k[6:]: z[k] for k in z.files if k.startswith("ref32_")}
    return sc, int(z["nsub"]), ref64, ref32
