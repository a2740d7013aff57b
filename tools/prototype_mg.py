"""Dev-only numpy prototype of the device multigrid/Krylov algorithms (not a product path).
Validates smoother degree / eigen-ratio / level choices on the CPU before they are written as CUDA."""
import sys, time
sys.path.insert(0, '/root/repo/fenics-eff-uptake_b200'); sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from sulcusfem import hostmesh as hm, hierarchy as hy, dofmap as dm
from oracle import cpu_oracle as co


def p1_ops(mesh, D, mu, ux=None, uy=None, upwind=True):
    x = mesh.coords; c = mesh.cells.astype(np.int64)
    p = x[c]
    det = (p[:,1,0]-p[:,0,0])*(p[:,2,1]-p[:,0,1]) - (p[:,2,0]-p[:,0,0])*(p[:,1,1]-p[:,0,1])
    g = np.empty((len(c),3,2))
    g[:,0,0]=p[:,1,1]-p[:,2,1]; g[:,0,1]=p[:,2,0]-p[:,1,0]
    g[:,1,0]=p[:,2,1]-p[:,0,1]; g[:,1,1]=p[:,0,0]-p[:,2,0]
    g[:,2,0]=p[:,0,1]-p[:,1,1]; g[:,2,1]=p[:,1,0]-p[:,0,0]
    g /= det[:,None,None]
    area = 0.5*np.abs(det)
    Dc = np.full(len(c), float(D))
    n = len(x)
    if ux is not None:
        uc = np.stack([ux[c].mean(1), uy[c].mean(1)],1)
        hc = np.sqrt(2*area)
        if upwind:
            pe = np.linalg.norm(uc,axis=1)*hc/(2*D)
            Dc = D*np.maximum(1.0, pe)       # crude artificial diffusion on coarse levels
    Ke = np.einsum('c,cid,cjd,c->cij', Dc, g, g, area)
    if ux is not None:
        # int (u.grad phi_j) phi_i with u const per cell: area/3 * (u.g_j)
        Ke = Ke + np.einsum('c,cj->cj', area/3.0, np.einsum('cd,cjd->cj', uc, g))[:,None,:].repeat(3,1)
    A = sp.coo_matrix((Ke.ravel(), (np.repeat(c[:,:,None],3,2).ravel(), np.repeat(c[:,None,:],3,1).ravel())), shape=(n,n)).tocsr()
    return A

def p1_robin(mesh, markers, mu):
    f, cells, loc = dm.boundary_facets(mesh, markers, 4)
    e = mesh.edges[f].astype(np.int64)
    L = np.linalg.norm(mesh.coords[e[:,0]]-mesh.coords[e[:,1]],axis=1)
    n = mesh.num_vertices
    rows = np.concatenate([e[:,0],e[:,0],e[:,1],e[:,1]]); cols=np.concatenate([e[:,0],e[:,1],e[:,0],e[:,1]])
    vals = np.concatenate([L/3,L/6,L/6,L/3])*mu
    return sp.coo_matrix((vals,(rows,cols)),shape=(n,n)).tocsr()

def sym_bc(A, b, dofs, vals):
    n = A.shape[0]
    g = np.zeros(n); g[dofs]=vals
    b = b - A@g
    keep = np.ones(n); keep[dofs]=0
    Dk = sp.diags(keep)
    A = Dk@A@Dk + sp.diags(1-keep)
    b[dofs]=vals
    return A.tocsr(), b

class Level: pass

def cheb_setup(A):
    dinv = 1.0/A.diagonal()
    # power iteration for lambda_max(D^-1 A)
    rng = np.random.default_rng(0); v = rng.random(A.shape[0])
    lam=1
    for _ in range(15):
        w = dinv*(A@v); lam = np.linalg.norm(w)/np.linalg.norm(v); v = w/np.linalg.norm(w)
    return dinv, 1.1*lam

def cheb(A, dinv, lmax, b, x, deg, ratio):
    lmin = lmax/ratio
    theta=0.5*(lmax+lmin); delta=0.5*(lmax-lmin); sigma=theta/delta; rho=1/sigma
    r = b - A@x if x is not None else b.copy()
    if x is None: x = np.zeros_like(b)
    d = dinv*r/theta
    for i in range(deg):
        x = x + d
        if i<deg-1:
            r = r - A@d
            rho_new = 1/(2*sigma-rho)
            d = rho_new*rho*d + 2*rho_new/delta*dinv*r
            rho = rho_new
    return x

def build_mg(H, fineA, fine_bc, D, mu, vel=None, deg=2, ratio=8, bc_ids=(1,2), robin=True, upwind=True):
    levels=[]
    L0=Level(); L0.A=fineA; L0.dinv,L0.lmax=cheb_setup(fineA); levels.append(L0)
    prev_bc = fine_bc
    for l, mesh in enumerate(H.meshes):
        mk = hy.level_markers(mesh)['bc_markers'].values
        ux=uy=None
        if vel is not None:
            X = mesh.coords; ux = vel[0](X); uy = vel[1](X)
        A = p1_ops(mesh, D, mu, ux, uy, upwind=upwind)
        if robin and mu: A = A + p1_robin(mesh, mk, mu)
        bc = np.unique(np.concatenate([dm.dirichlet_dofs_p1(mesh, mk, i) for i in bc_ids]))
        A,_ = sym_bc(A.tocsr(), np.zeros(A.shape[0]), bc, np.zeros(len(bc)))
        T = H.transfers[l]
        P = sp.csr_matrix((T.vals, T.cols, T.rowptr), shape=(T.n_fine, T.n_coarse))
        # zero rows of fine bc dofs, cols of coarse bc dofs
        kf = np.ones(T.n_fine); kf[prev_bc]=0; kc=np.ones(T.n_coarse); kc[bc]=0
        P = sp.diags(kf)@P@sp.diags(kc)
        Lv=Level(); Lv.A=A; Lv.P=P.tocsr(); Lv.R=P.T.tocsr(); Lv.dinv,Lv.lmax=cheb_setup(A); levels.append(Lv)
        prev_bc = bc
    levels[-1].inv = np.linalg.inv(levels[-1].A.toarray())
    def vcycle(l, b):
        Lv = levels[l]
        if l==len(levels)-1:
            return Lv.inv@b
        x = cheb(Lv.A, Lv.dinv, Lv.lmax, b, None, deg, ratio)
        r = b - Lv.A@x
        nxt = levels[l+1]
        xc = vcycle(l+1, nxt.R@r)
        x = x + nxt.P@xc
        x = cheb(Lv.A, Lv.dinv, Lv.lmax, b, x, deg, ratio)
        return x
    return levels, (lambda b: vcycle(0,b))

def pcg(A,b,M,x0,rtol,maxit=200):
    x=x0.copy(); r=b-A@x; z=M(r); p=z.copy(); rz=r@z; b0=np.linalg.norm(b); hist=[]
    for it in range(maxit):
        q=A@p; a=rz/(p@q); x+=a*p; r-=a*q
        rn=np.linalg.norm(r)/b0; hist.append(rn)
        if rn<rtol: break
        z=M(r); rz2=r@z; p=z+(rz2/rz)*p; rz=rz2
    return x,hist

def fgmres(A,b,M,x0,rtol,m=60,maxit=200):
    x=x0.copy(); b0=np.linalg.norm(b); hist=[]
    while True:
        r=b-A@x; beta=np.linalg.norm(r)
        V=[r/beta]; Z=[]; Hm=np.zeros((m+1,m)); 
        for j in range(m):
            z=M(V[j]); Z.append(z); w=A@z
            for _ in range(2):
                for i in range(j+1):
                    h=V[i]@w; Hm[i,j]+=h; w=w-h*V[i]
            Hm[j+1,j]=np.linalg.norm(w); V.append(w/Hm[j+1,j])
            e1=np.zeros(j+2); e1[0]=beta
            y,res,_,_=np.linalg.lstsq(Hm[:j+2,:j+1],e1,rcond=None)
            rn=np.linalg.norm(Hm[:j+2,:j+1]@y-e1)/b0; hist.append(rn)
            if rn<rtol or len(hist)>=maxit: break
        x=x+sum(yi*zi for yi,zi in zip(y,Z))
        if rn<rtol or len(hist)>=maxit: return x,hist

if __name__=='__main__':
    h=float(sys.argv[1]) if len(sys.argv)>1 else 0.04
    deg=int(sys.argv[2]) if len(sys.argv)>2 else 2
    ratio=float(sys.argv[3]) if len(sys.argv)>3 else 8
    mesh=hm.sulcus_mesh(10,1,0.5,1.0,h)
    H=hy.build_hierarchy(mesh)
    print('levels', [m.num_vertices for m in H.meshes])
    mk=hm.build_markers(mesh,10,1,4.75,5.25,'sulcus')
    om=co.Mesh(mesh.coords,mesh.cells); bm=mk['bc_markers'].values
    # --- pure diffusion
    for mu in (1.0,):
        c_ref,A,b=co.solve_concentration(om,bm,1.0,mu=mu)
        K=co.assemble_p2_stiffness(om)+co.assemble_p2_robin(om,np.flatnonzero((bm==4)&om.on_boundary),mu_const=mu)
        dofs,vals=co.concentration_bcs(om,bm)
        As,bs=sym_bc(K.tocsr(),np.zeros(om.n_p2),dofs,vals)
        lv,M=build_mg(H,As,dofs,1.0,mu,deg=deg,ratio=ratio)
        x0=np.zeros(om.n_p2); x0[dofs]=vals
        t=time.time(); x,hist=pcg(As,bs,M,x0,1e-13); 
        print('diffusion mu',mu,'pcg its',len(hist),'relL2 vs LU',np.linalg.norm(x-c_ref)/np.linalg.norm(c_ref), 'time',time.time()-t)
    # --- advection-diffusion with Poiseuille-like velocity (zero in cavity) Pe=40
    for Dv in (0.025,0.1):
        X=om.p2_dof_coords(); y=np.clip(X[:,1],0,1); ux=4*y*(1-y); uy=np.zeros_like(ux)
        c_ref,A,b=co.solve_concentration(om,bm,Dv,mu=1.0,ux=ux,uy=uy)
        Aa=Dv*co.assemble_p2_stiffness(om)+co.assemble_p2_advection(om,ux,uy)+co.assemble_p2_robin(om,np.flatnonzero((bm==4)&om.on_boundary),mu_const=1.0)
        As,bs=sym_bc(Aa.tocsr(),np.zeros(om.n_p2),dofs,vals)
        vel=(lambda X:4*np.clip(X[:,1],0,1)*(1-np.clip(X[:,1],0,1)), lambda X:np.zeros(len(X)))
        for up in (True,False):
            lv,M=build_mg(H,As,dofs,Dv,1.0,vel=vel,deg=deg,ratio=ratio,upwind=up)
            x,hist=fgmres(As,bs,M,x0,1e-13)
            print('advdiff D',Dv,'upwind',up,'gmres its',len(hist),'relL2 vs LU',np.linalg.norm(x-c_ref)/np.linalg.norm(c_ref))
