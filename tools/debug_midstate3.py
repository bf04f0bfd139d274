import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import numpy as np, torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
import test_parity_gpu as T
scene,k0="demo_like",25
sc=getattr(S,"scene_"+scene)()
jt=np.zeros((sc.num_joint_t,3),np.float32) if sc.num_joint_t else None
o=T.run_oracle(sc,k0,"f32",1,jt)
d_before=o.d.copy()
solver,model,state=build_from_scene(sc)
solver.set_debug(True)
T._load_oracle_state(o,sc,solver,model,state)
fi=sc.frame_inputs(0); ft=frame_tensors(sc,0)
mx=fi["mesh_x"]+np.float32(sc.dt*k0)*fi["mesh_v"]
o.p2g2p(sc.dt,mx,fi["mesh_v"],jt,fi["joint_verts_v"],fi["joint_faces_v"])
jtt=None if jt is None else torch.as_tensor(jt,device="cuda")
solver.p2g2p(model,state,sc.dt,mesh_x=torch.as_tensor(mx,device="cuda"),mesh_v=ft["mesh_v"],joint_traditional_v=jtt,joint_verts_v=ft["joint_verts_v"],joint_faces_v=ft["joint_faces_v"])
Ne=sc.n_elements
st=state.particle_stress.cpu().numpy()[:Ne]; so=o.stress[:Ne]
es=np.abs(st-so).reshape(Ne,-1).max(1); sm=np.abs(so).max()
print("stress max",sm,"worst rel",es.max()/sm,"n>1e-3",(es>1e-3*sm).sum())
vf=state.vertex_force.cpu().numpy(); ef=np.abs(vf-o.vertex_force).max(1); fm=np.abs(o.vertex_force).max()
print("vforce max",fm,"worst rel",ef.max()/fm,"n>1e-3",(ef>1e-3*fm).sum())
from oracle.oracle import OracleSim
h=OracleSim(1,0,0,8,2.0,"f64")
for e in np.argsort(-es)[:6]:
    Q,R=h.qr3_signed(d_before[e].astype(np.float64))
    print("elem",e,"stress err rel",es[e]/sm,"r22-1",R[2,2]-1,"r02,r12",R[0,2],R[1,2],"|S_cuda|",np.abs(st[e]).max(),"|S_ref|",np.abs(so[e]).max())
for i in np.argsort(-ef)[:6]:
    print("vert",i,"f cuda",vf[i],"ref",o.vertex_force[i])
dd=state.particle_d.cpu().numpy(); print("d rel", np.abs(dd-o.d).max()/np.abs(o.d).max())
