"""Dev-only prototype: does a lubrication coarse correction of the pressure Schur-complement
preconditioner cut the MINRES iteration count on the long channel?  (numpy/scipy, not a product path)"""
import sys
sys.path.insert(0, '/root/repo/fenics-eff-uptake_b200'); sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from sulcusfem import hostmesh as hm
from sulcusfem.unstructured import mesh_domain
from oracle import cpu_oracle as co

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.08
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 11
L, H, w, d = 10.0, 1.0, 0.5, 1.0
mesh = mesh_domain(L, H, w, d, h, 'sulcus')
mk = hm.build_markers(mesh, L, H, 4.75, 5.25, 'sulcus')
om = co.Mesh(mesh.coords, mesh.cells)
bm = mk['bc_markers'].values
A0 = co.assemble_stokes(om)
dofs, vals = co.stokes_bcs(om, bm, H)
n = A0.shape[0]; n2 = om.n_p2; nv = om.nv
g = np.zeros(n); g[dofs] = vals
b = -A0 @ g
keep = np.ones(n); keep[dofs] = 0
Dk = sp.diags(keep)
A = (Dk @ A0 @ Dk + sp.diags(1 - keep)).tocsr()
b[dofs] = vals
K = A[:n2, :n2].tocsc()
Klu = spla.splu(K)
# pressure mass
lam, wq = co.triangle_rule(2)
Me = np.einsum('q,qi,qj,c->cij', wq, lam, lam, np.abs(om.det))
c = om.cells
Mp = sp.coo_matrix((Me.ravel(), (np.repeat(c[:, :, None], 3, 2).ravel(), np.repeat(c[:, None, :], 3, 1).ravel())), shape=(nv, nv)).tocsc()
Mlu = spla.splu(Mp)
X = om.x
# 1-D hats in x
zn = np.linspace(0, L, nz)
hz = zn[1] - zn[0]
Z = np.maximum(0.0, 1.0 - np.abs(X[:, 0][:, None] - zn[None, :]) / hz)      # [nv, nz]
G = Z.T @ (Mp @ Z)
# lubrication operator: E = int H(x)^3/12 z_i' z_j' dx, Dirichlet p=0 at the outlet (natural outflow), Neumann at the inlet
xs = np.linspace(0, L, 20001)
xm = 0.5 * (xs[1:] + xs[:-1]); dx = xs[1] - xs[0]
floor = np.where((xm > 4.75) & (xm < 5.25), -d * np.sin(np.pi * (xm - 4.75) / w), 0.0)
Hx = H - floor
dZ = (np.maximum(0.0, 1.0 - np.abs(xm[:, None] + 1e-9 - zn[None, :]) / hz) - np.maximum(0.0, 1.0 - np.abs(xm[:, None] - 1e-9 - zn[None, :]) / hz)) / 2e-9
E = (dZ * (Hx ** 3 / 12.0 * dx)[:, None]).T @ dZ
mode = sys.argv[3] if len(sys.argv) > 3 else 'lub'
if mode == 'exact':
    # exact Galerkin coarse Schur complement (reference point for what the correction can reach)
    Bt = A[:2 * n2, 2 * n2:].tocsc()
    BZ = Bt @ Z
    KiBZ = np.vstack([Klu.solve(BZ[:n2]), Klu.solve(BZ[n2:])])
    E = BZ.T @ KiBZ
# outlet Dirichlet-ish: the do-nothing outflow pins p ~ 0 at x = L: add a large penalty on the last hat
if mode != 'exact':
    E[-1, -1] += 1e3 * E.max()
Cc = np.linalg.inv(E) - np.linalg.inv(G)
Cc = 0.5 * (Cc + Cc.T)
ev, V = np.linalg.eigh(Cc)
print('coarse correction eigenvalues min/max', ev.min(), ev.max(), 'neg', (ev < 0).sum())
Cc = (V * np.maximum(ev, 0.0)) @ V.T


def make_M(corr):
    def M(r):
        out = np.empty_like(r)
        out[:n2] = Klu.solve(r[:n2]); out[n2:2 * n2] = Klu.solve(r[n2:2 * n2])
        rp = r[2 * n2:]
        zp = Mlu.solve(rp)
        if corr:
            zp = zp + Z @ (Cc @ (Z.T @ rp))
        out[2 * n2:] = zp
        return out
    return spla.LinearOperator((n, n), M)


for corr in (False, True):
    it = [0]
    def cb(xk): it[0] += 1
    x, info = spla.minres(A, b, M=make_M(corr), rtol=1e-12, maxiter=2000, callback=cb)
    print('h', h, 'nz', nz, 'mode', mode, 'corr', corr, 'iters', it[0], 'relres', np.linalg.norm(b - A @ x) / np.linalg.norm(b))

# pressure scaling experiment
for scale in (0.5, 0.7, 1.5, 2.0):
    def Ms(r, scale=scale):
        out = np.empty_like(r)
        out[:n2] = Klu.solve(r[:n2]); out[n2:2 * n2] = Klu.solve(r[n2:2 * n2])
        rp = r[2 * n2:]
        out[2 * n2:] = scale * (Mlu.solve(rp) + Z @ (Cc @ (Z.T @ rp)))
        return out
    it = [0]
    def cb(xk): it[0] += 1
    x, info = spla.minres(A, b, M=spla.LinearOperator((n, n), Ms), rtol=1e-12, maxiter=2000, callback=cb)
    print('scale', scale, 'iters', it[0])
