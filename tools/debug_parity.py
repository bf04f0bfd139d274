import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpmavatar_b200 import synthetic as S
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_parity_gpu as T

sc = S.scene_small_cloth_body()
for nsub in (1, 2, 3, 5, 10):
    o = T.run_oracle(sc, nsub)
    _, _, st = T.run_cuda(sc, nsub)
    x = st.particle_x.cpu().numpy(); v = st.particle_v.cpu().numpy()
    Ne = sc.n_elements
    ex = np.abs(x - o.x).max(1); ev = np.abs(v - o.v).max(1)
    print(nsub, "x err E/V", ex[:Ne].max(), ex[Ne:].max(), "v err E/V", ev[:Ne].max(), ev[Ne:].max(), "vmax", np.abs(o.v).max(),
          "worst v idx", int(ev.argmax()), "is joint vert", (int(ev.argmax()) - Ne) < sc.num_joint_v)
    i = int(ev.argmax())
    print("   v cuda", v[i], "v ref", o.v[i], "x", o.x[i])
