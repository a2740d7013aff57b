"""GPU diagnostic: iteration counts of the three solvers on a refined mesh (dev tool)."""
import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'fenics-eff-uptake_b200'))
import numpy as np, torch
from bench import build_mesh, H_CH
from sulcusfem import dofmap as dm
from sulcusfem.device import Context, ScalarProblem, StokesProblem
from sulcusfem.hierarchy import build_hierarchy
h=float(sys.argv[1]); r=int(sys.argv[2])
ctx=Context.get()
mr=build_mesh(h,r); mesh=mr['mesh']; bm=mr['bc_markers'].values
hier=build_hierarchy(mesh); print('levels',[m.num_vertices for m in hier.meshes])
st=StokesProblem(mesh,bm,hierarchy=hier,ctx=ctx)
X=dm.p2_dof_coordinates(mesh); d1=dm.dirichlet_dofs_p2(mesh,bm,1)
st.set_bcs({1:(4*X[d1,1]*(H_CH-X[d1,1]),0.0),4:(0.0,0.0),3:(0.0,0.0)})
st.assemble(1)
print('vel lmax',st.vel.mg.lambda_max())
# K-only CG
v=st.vel; f=v.fine
f.rhs.copy_(torch.rand(f.n,dtype=torch.float64,device=ctx.device)*(1-f.bc_flag.to(torch.float64)))
x=v.solve('cg',rtol=1e-12); print('K cg',v.last_info)
for rt in (1e-10,1e-12,1e-14):
    t=time.time(); st.solve(rtol=rt,maxit=800); torch.cuda.synchronize(); print('minres rtol',rt,st.last_info,'%.3fs'%(time.time()-t))
