"""Dev-only prototype (numpy / scipy, not a product path): two-level Schur-complement correction for the Stokes MINRES
preconditioner -- the exact Schur complement of a COARSE Taylor-Hood discretisation (host sparse LU on a few hundred
pressure dofs) interpolated to the fine pressure space, in place of the 1-D lubrication correction of
sulcusfem/schur.py.  Evidence for DESIGN.md section 9, item 2.

    python tools/prototype_schur_coarse.py H_FINE H_COARSE

MINRES iterations to rtol 1e-12 with exact K^-1 / Mp^-1 (so only the Schur part is being compared), sulcus 0.5 x 1.0:

    fine h   lubrication, 11 hats (shipped)   coarse mesh h = 0.3 (166 p-dofs)   coarse mesh h = 0.2 (323 p-dofs)
    0.08     59                               45                                 43
    0.04     55                               49                                 --

(mass-matrix-only Schur approximation: 95 / 93; lubrication with 41 / 81 hats: 53 / 51 at h = 0.08, 53 / 53 at h = 0.04;
exact Galerkin coarse Schur complement on 41 1-D hats: 49; on a 2-D grid of 69 / 217 bilinear hats: 41 / 37 -- but that
needs one K-solve per coarse dof at set-up, which costs more than the Stokes solve it accelerates.)
"""
import sys
sys.path.insert(0, '/root/repo/fenics-eff-uptake_b200'); sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from sulcusfem import hostmesh as hm, hierarchy as hy
from sulcusfem.unstructured import mesh_domain
from oracle import cpu_oracle as co
h = float(sys.argv[1]); hc = float(sys.argv[2])
L, H, w, d = 10.0, 1.0, 0.5, 1.0

def stokes_system(mesh):
    mk = hm.build_markers(mesh, L, H, 4.75, 5.25, 'sulcus')
    om = co.Mesh(mesh.coords, mesh.cells)
    bm = mk['bc_markers'].values
    A0 = co.assemble_stokes(om)
    dofs, vals = co.stokes_bcs(om, bm, H)
    n = A0.shape[0]
    g = np.zeros(n); g[dofs] = vals
    b = -A0 @ g
    keep = np.ones(n); keep[dofs] = 0
    Dk = sp.diags(keep)
    A = (Dk @ A0 @ Dk + sp.diags(1 - keep)).tocsr()
    b[dofs] = vals
    return om, A, b

def mass_p1(om):
    lam, wq = co.triangle_rule(2)
    Me = np.einsum('q,qi,qj,c->cij', wq, lam, lam, np.abs(om.det))
    c = om.cells
    return sp.coo_matrix((Me.ravel(), (np.repeat(c[:, :, None], 3, 2).ravel(), np.repeat(c[:, None, :], 3, 1).ravel())), shape=(om.nv, om.nv)).tocsc()

mesh = mesh_domain(L, H, w, d, h, 'sulcus')
om, A, b = stokes_system(mesh)
n = A.shape[0]; n2 = om.n_p2; nv = om.nv
K = A[:n2, :n2].tocsc(); Klu = spla.splu(K)
Mp = mass_p1(om); Mlu = spla.splu(Mp)
# coarse mesh Schur complement
cm = mesh_domain(L, H, w, d, hc, 'sulcus')
omc, Ac, bc_ = stokes_system(cm)
n2c, nvc = omc.n_p2, omc.nv
Kc = Ac[:2 * n2c, :2 * n2c].tocsc()
Btc = Ac[:2 * n2c, 2 * n2c:].tocsc()
Sc = (Btc.T @ spla.splu(Kc).solve(Btc.toarray()))
Sc = 0.5 * (Sc + Sc.T)
Mc = mass_p1(omc).toarray()
T = hy.interpolation_transfer(mesh, cm)
Z = sp.csr_matrix((T.vals, T.cols, T.rowptr), shape=(nv, nvc))
print('coarse pressure dofs', nvc, 'cond Sc', np.linalg.cond(Sc))
for variant in ('G=ZtMZ', 'G=Mc'):
    G = (Z.T @ (Mp @ Z)).toarray() if variant == 'G=ZtMZ' else Mc
    Cc = np.linalg.inv(Sc) - np.linalg.inv(G)
    Cc = 0.5 * (Cc + Cc.T)
    ev, V = np.linalg.eigh(Cc)
    Cp = (V * np.maximum(ev, 0.0)) @ V.T
    def M(r):
        out = np.empty_like(r)
        out[:n2] = Klu.solve(r[:n2]); out[n2:2 * n2] = Klu.solve(r[n2:2 * n2])
        rp = r[2 * n2:]
        out[2 * n2:] = Mlu.solve(rp) + Z @ (Cp @ (Z.T @ rp))
        return out
    it = [0]
    def cb(xk): it[0] += 1
    x, info = spla.minres(A, b, M=spla.LinearOperator((n, n), M), rtol=1e-12, maxiter=2000, callback=cb)
    print('h', h, 'hc', hc, variant, 'neg ev', (ev < 0).sum(), 'iters', it[0], 'relres', np.linalg.norm(b - A @ x) / np.linalg.norm(b))
