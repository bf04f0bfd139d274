"""Value test of MPMWARP.export_particle_cov_to_torch against the reference's own compute_cov_from_F
(warp_mpm/mpm_solver.py:543-561, mpm_utils.py:1108-1132) run under oracle/warp_emu.py
(tests/golden/make_golden.py make_cov_fixture -> tests/golden/cov_export.npz)."""
import os

import numpy as np
import pytest
import torch

from mpmavatar_b200.warp_mpm.mpm_data_structure import MPMStateStruct
from mpmavatar_b200.warp_mpm.mpm_solver import MPMWARP

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cov_export.npz")


def _state(device):
    z = np.load(FIX)
    N, Ne, Nv = (int(v) for v in z["counts"])
    st = MPMStateStruct()
    st.init(N, Ne, Nv, device=device)
    st.particle_F_trial = torch.as_tensor(z["F_trial"], device=device)
    st.particle_cov = torch.as_tensor(z["cov"], device=device)
    return st, z["ref64_new_cov"]


def test_cov_export_matches_reference_source_cpu_tensors():
    st, ref = _state("cpu")
    out = MPMWARP.export_particle_cov_to_torch(None, st, device="cpu")  # the method only reads the state
    assert out.shape == (ref.shape[0],) and out.dtype == torch.float32
    assert np.abs(out.numpy() - ref).max() < 1e-5 * np.abs(ref).max()


def test_cov_export_matches_numpy_restatement():
    st, ref = _state("cpu")
    z = np.load(FIX)
    F = z["F_trial"].astype(np.float64)
    c = z["cov"].astype(np.float64).reshape(-1, 6)
    S = np.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], -1).reshape(-1, 3, 3)
    cov = F @ S @ F.transpose(0, 2, 1)
    flat = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], -1).reshape(-1)
    assert np.abs(flat - ref).max() < 1e-5 * np.abs(ref).max()  # fixture inputs are stored in fp32


@pytest.mark.gpu
def test_cov_export_through_solver_on_gpu():
    st, ref = _state("cuda:0")
    N, Ne, Nv = st.n_particles, st.n_elements, st.n_vertices
    solver = MPMWARP(N, Ne, Nv, n_grid=16, grid_lim=2.0)
    out = solver.export_particle_cov_to_torch(st)
    assert out.is_cuda
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-5 * np.abs(ref).max()
