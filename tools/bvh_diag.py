import sys,tempfile,os
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import __graft_entry__ as ge
b=ge.load_package()
import scene_util
w,nt=scene_util.fixture_plus_mesh(b,tempfile.mkdtemp(),n=225)
ctx=b.Context(0); ctx.upload_scene(w)
rng=np.random.default_rng(1)
c=np.array([0.7,1.0,-0.5])
def run(name, rays):
    ctx.reset_stats(); g=ctx.intersect(rays, b.CAST_BVH); s=ctx.stats()
    t=ctx.intersect(rays, b.CAST_TWO_PHASE)
    print(name, "exact/cast", s["exact_confirms"]/len(rays), "walks", s["certify_fallbacks"], "hit frac", (g["prim_id"]>=0).mean(), "same ids", np.array_equal(g["prim_id"],t["prim_id"]))
n=4096
rays=np.zeros(n,dtype=b.RAY_DTYPE); rays["exclude_prim"]=-1
o=np.tile(c,(n,1)); o[:,:2]+=rng.uniform(-0.3,0.3,(n,2))
u=(o[:,0]-0.7)*3; v=(o[:,1]-1.0)*3
o[:,2]=-0.5+0.1*np.sin(6*np.pi*u)*np.cos(6*np.pi*v)/3+1e-4
rays["origin"]=o.astype(np.float32)
rays["direction"]=np.array([0.70710678,0.70710678,0.0],dtype=np.float32)
for face in (0,1,2):
    rays["face_direction"]=face
    run(f"sheet rays face {face}", rays)
d=np.array([0,10.0,0])-o; d/=np.linalg.norm(d,axis=1,keepdims=True)
rays["direction"]=d.astype(np.float32); rays["face_direction"]=1
run("to spot light (back)", rays)
eye=np.array([2,2.5,2.0]); tgt=c+rng.uniform(-0.3,0.3,(n,3))*np.array([1,1,0.1]); d=tgt-eye; d/=np.linalg.norm(d,axis=1,keepdims=True)
rays["origin"]=eye.astype(np.float32); rays["direction"]=d.astype(np.float32); rays["face_direction"]=0
run("camera rays", rays)
d=rng.normal(size=(n,3)); d/=np.linalg.norm(d,axis=1,keepdims=True)
rays["origin"]=rng.uniform(-1.5,1.5,(n,3)).astype(np.float32); rays["direction"]=d.astype(np.float32); rays["face_direction"]=2
run("random rays", rays)
