import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import numpy as np, torch
from mpmavatar_b200 import synthetic as S
from mpmavatar_b200.scene_setup import build_from_scene, frame_tensors
import test_parity_gpu as T
for scene,k0 in (("small_cloth_body",60),("small_cloth_body",60),("demo_like",25),("demo_like",25)):
    sc=getattr(S,"scene_"+scene)()
    jt=np.zeros((sc.num_joint_t,3),np.float32) if sc.num_joint_t else None
    o=T.run_oracle(sc,k0,"f32",1,jt)
    solver,model,state=build_from_scene(sc)
    T._load_oracle_state(o,sc,solver,model,state)
    fi=sc.frame_inputs(0); ft=frame_tensors(sc,0)
    mx=fi["mesh_x"]+np.float32(sc.dt*k0)*fi["mesh_v"]
    o.p2g2p(sc.dt,mx,fi["mesh_v"],jt,fi["joint_verts_v"],fi["joint_faces_v"])
    jtt=None if jt is None else torch.as_tensor(jt,device="cuda")
    solver.p2g2p(model,state,sc.dt,mesh_x=torch.as_tensor(mx,device="cuda"),mesh_v=ft["mesh_v"],joint_traditional_v=jtt,joint_verts_v=ft["joint_verts_v"],joint_faces_v=ft["joint_faces_v"])
    v=state.particle_v.cpu().numpy(); ev=np.abs(v-o.v).max(1)/np.abs(o.v).max()
    Ne,Nt=sc.n_elements,sc.n_traditional
    print(scene,k0,"max",ev.max(),"p99.9",np.quantile(ev,0.999),"E",ev[:Ne].max(),"T",ev[Ne:Ne+Nt].max() if Nt else 0,"V",ev[Ne+Nt:].max(), "n>1e-4", (ev>1e-4).sum())
    w=np.argsort(-ev)[:5]
    for i in w: print("   idx",i,"cls","E" if i<Ne else ("T" if i<Ne+Nt else "V"),"x",o.x[i],"v",v[i],"ref",o.v[i])
