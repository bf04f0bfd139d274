"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/mpm_b200.h declares; the host mirror exposes the reference's names; the product path
refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from mpmavatar_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "mpm_b200.h")).read()
    declared = set(re.findall(r"\b(mpm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)


def test_struct_sizes_match_header_layout():
    from mpmavatar_b200 import _lib
    assert ctypes.sizeof(_lib.MpmConfig) == 11 * 4
    assert ctypes.sizeof(_lib.MpmModelParams) == 12 * 4
    assert ctypes.sizeof(_lib.MpmParticleArrays) == 17 * 8
    assert ctypes.sizeof(_lib.MpmFrameInputs) == 7 * 8


def test_reference_names_and_signatures():
    import inspect
    from mpmavatar_b200.warp_mpm.mpm_solver import MPMWARP, MPMSolverWarp
    from mpmavatar_b200.warp_mpm.mpm_data_structure import MPMStateStruct, MPMModelStruct
    sig = inspect.signature(MPMWARP.p2g2p)
    assert list(sig.parameters)[:10] == ["self", "mpm_model", "mpm_state", "dt", "mesh_x", "mesh_v",
                                         "joint_traditional_v", "joint_verts_v", "joint_faces_v", "device"]
    for m in ("set_parameters_dict", "set_E_nu_from_torch", "prepare_mu_lam", "add_mesh_collider",
              "add_particle_mover", "add_surface_collider", "set_velocity_on_cuboid", "add_bounding_box",
              "enforce_grid_velocity_by_mask", "add_impulse_on_particles", "enforce_particle_velocity_translation",
              "print_time_profile", "export_particle_cov_to_torch"):
        assert hasattr(MPMWARP, m), m
    for m in ("init", "from_torch", "reset_state", "continue_from_torch", "reset_density", "set_require_grad"):
        assert hasattr(MPMStateStruct, m), m
    for m in ("init", "init_other_params", "finalize_mu_lam", "from_torch"):
        assert hasattr(MPMModelStruct, m), m
    assert issubclass(MPMSolverWarp, MPMWARP)


def test_install_registers_reference_module_paths():
    import mpmavatar_b200
    mpmavatar_b200.install()
    from warp_mpm.mpm_solver import MPMWARP  # noqa: F401
    from warp_mpm.mpm_data_structure import MPMStateStruct, MPMModelStruct  # noqa: F401
    import warp as wp
    t = torch.zeros(3)
    assert wp.to_torch(t) is t
    wp.init()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from mpmavatar_b200.warp_mpm.mpm_solver import MPMWARP
    with pytest.raises(RuntimeError):
        MPMWARP(10, 0, 0, n_grid=16, grid_lim=2.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mpmavatar_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"(import|from)\s+oracle|liboracle|oracle/", src), f"{f} uses the oracle"


def test_header_is_plain_c_and_links(tmp_path):
    """include/mpm_b200.h is the boundary a non-Python host binds: it must compile as C (gcc -std=c99 -pedantic) and
    link against the library.  The program only calls entry points that need no GPU."""
    import shutil
    import subprocess
    from mpmavatar_b200 import build as b
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    lib = b.build_cuda()
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "mpm_b200.h"\n'
                   "int main(void) {\n"
                   "    MpmConfig cfg; MpmSolver *h = 0;\n"
                   "    memset(&cfg, 0, sizeof cfg);\n"
                   "    /* invalid particle counts: must be refused with a message, not crash */\n"
                   "    int rc = mpm_create(&cfg, &h);\n"
                   '    printf("%d %s\\n", rc, mpm_last_error(0));\n'
                   "    return (rc != 0 && h == 0 && mpm_shared_mode(0) == -1) ? 0 : 1;\n"
                   "}\n")
    exe = tmp_path / "abi"
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           lib, "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.run(cmd, check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[0] != "0" and len(out.stdout.split()) > 1  # an error code and a message
