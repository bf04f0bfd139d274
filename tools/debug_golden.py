import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.golden_util import load
from tests.test_golden import _run_cuda, run_oracle
sc, nsub, ref, _ = load("cloth_body_joints")
solver, state = _run_cuda(sc, nsub)
Ne = sc.n_elements
s = state.particle_stress.cpu().numpy()[:Ne]
r = ref["stress"][:Ne]
err = np.abs(s - r).reshape(Ne, -1).max(1)
bad = np.nonzero(err > 1e-3 * np.abs(r).max())[0]
print("Ne", Ne, "bad", len(bad), bad[:40])
o = run_oracle(sc, nsub, "f64")
print("oracle vs ref stress", np.abs(o.stress[:Ne]-r).max())
d = state.particle_d.cpu().numpy()
print("d err", np.abs(d - ref["d"]).max())
for b in bad[:3]:
    print(b, "cuda", s[b].ravel(), "\n   ref", r[b].ravel(), "\n   d", d[b].ravel())
print("zero rows", np.nonzero(np.abs(s).reshape(Ne,-1).max(1) == 0)[0][:50])
