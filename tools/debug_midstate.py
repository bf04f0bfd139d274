import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import numpy as np, torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
import test_parity_gpu as T
sc=S.scene_small_cloth_body(); k0=60
o=T.run_oracle(sc,k0,"f32",1)
solver,model,state=build_from_scene(sc)
solver.set_debug(True)
T._load_oracle_state(o,sc,solver,model,state)
fi=sc.frame_inputs(0); ft=frame_tensors(sc,0)
mx=fi["mesh_x"]+np.float32(sc.dt*k0)*fi["mesh_v"]
o.p2g2p(sc.dt,mx,fi["mesh_v"],None,fi["joint_verts_v"],fi["joint_faces_v"])
solver.p2g2p(model,state,sc.dt,mesh_x=torch.as_tensor(mx,device="cuda"),mesh_v=ft["mesh_v"],joint_verts_v=ft["joint_verts_v"],joint_faces_v=ft["joint_faces_v"])
gm,gvi,gvo=[a.cpu().numpy() for a in state.export_grid()]
gm=gm.reshape(-1); gvi=gvi.reshape(-1,3); gvo=gvo.reshape(-1,3)
has=o.grid_m>1e-15
print("grid_m rel", np.abs(gm-o.grid_m).max()/o.grid_m.max())
print("grid_v_in rel", np.abs(gvi-o.grid_v_in).max()/np.abs(o.grid_v_in).max())
e=np.abs(gvo[has]-o.grid_v_out[has]).max(1)
print("grid_v_out max abs", e.max(), "rel", e.max()/np.abs(o.grid_v_out[has]).max(), "p99", np.quantile(e,0.99))
# velocity-equivalent error of v_in: (gvi - ref)/m
ev=np.abs(gvi[has]-o.grid_v_in[has]).max(1)/o.grid_m[has]
print("v_in/m err max", ev.max(), "p99", np.quantile(ev,0.99), " mass min among has", o.grid_m[has].min(), o.grid_m.max())
i=np.argmax(ev); print("worst node mass", o.grid_m[has][i], "v_in", gvi[has][i], o.grid_v_in[has][i])
vf=state.vertex_force.cpu().numpy()
print("vertex_force rel", np.abs(vf-o.vertex_force).max()/np.abs(o.vertex_force).max())
st=state.particle_stress.cpu().numpy()[:sc.n_elements]
print("stress rel", np.abs(st-o.stress[:sc.n_elements]).max()/np.abs(o.stress[:sc.n_elements]).max())
v=state.particle_v.cpu().numpy(); evp=np.abs(v-o.v).max(1)
print("v err max", evp.max(), "p99.9", np.quantile(evp,0.999))
