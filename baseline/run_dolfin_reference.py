#!/usr/bin/env python
"""Run the UNMODIFIED reference solvers under real dolfin and dump everything parity needs.

    python baseline/run_dolfin_reference.py --reference /root/reference --mesh mesh.xml --mode adv-diff --out ref.npz

Requires legacy FEniCS (dolfin 2019.1).  That stack is not installable in this image (no network, not in the
wheelhouse; the reference itself has no setup.py / pyproject.toml, so there is nothing to ``pip install``
into ``baseline/_ref``) -- the expected outcome here is the message below and exit code 3.  Where dolfin
exists, the script
  * reads a dolfin-XML mesh written by ``sulcusfem.hostmesh.write_dolfin_xml`` (identical vertex / cell arrays),
  * marks facets with the reference's own ``mesh.py`` predicates,
  * calls ``solvers.stokes_solver`` / ``advdiff_solver`` / ``pure_diffusion_solver`` from the reference tree,
  * stores DOF coordinates, cell dofs, the CSR of ``assemble(a)``, solution vectors, wall times and the
    flux / mass / mu_eff dictionaries in one ``.npz``;
``tests`` can then match DOFs by coordinates (dolfin's numbering is build dependent) and
compare fields at 1e-10.
"""
import argparse
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--mesh', required=False)
    ap.add_argument('--mode', default='adv-diff', choices=['adv-diff', 'no-adv'])
    ap.add_argument('--mu', type=float, default=1.0)
    ap.add_argument('--pe', type=float, default=40.0)
    ap.add_argument('--out', default='dolfin_reference.npz')
    args = ap.parse_args()
    try:
        import dolfin as df
    except Exception as e:                                   # the expected path in this image
        print(f"dolfin unavailable ({type(e).__name__}: {e}); the reference's FEniCS path cannot run here. "
              f"bench.py --impl reference therefore times the oracle port (oracle/cpu_oracle.py).")
        return 3
    import numpy as np
    sys.path.insert(0, args.reference)
    import solvers as ref_solvers                            # the reference's own module, unmodified
    import mesh as ref_mesh
    mesh = df.Mesh(args.mesh)
    L, H = 10.0, 1.0
    gen = ref_mesh.MeshGenerator.__new__(ref_mesh.MeshGenerator)
    gen.width, gen.height, gen.sulcus_left_x, gen.sulcus_right_x = L, H, 4.75, 5.25
    gen.TOLERANCE = ref_mesh.MeshGenerator.TOLERANCE
    bc = gen._create_boundary_markers(mesh, ['left', 'right', 'top', 'bottom']) if hasattr(gen, '_create_boundary_markers') else None
    mesh_results = {'mesh': mesh, 'bc_markers': bc}
    V = df.VectorFunctionSpace(mesh, 'P', 2)
    Q = df.FunctionSpace(mesh, 'P', 1)
    W = df.FunctionSpace(mesh, df.MixedElement([V.ufl_element(), Q.ufl_element()]))
    C = df.FunctionSpace(mesh, 'CG', 2)
    out = {}
    t0 = time.perf_counter()
    if args.mode == 'adv-diff':
        u, p = ref_solvers.stokes_solver(mesh_results, W, L, H, 'sulcus')
        t1 = time.perf_counter()
        c = ref_solvers.advdiff_solver(mesh_results, u, C, df.Constant(1.0 / args.pe), df.Constant(args.mu), 'sulcus')
        out.update(u=u.vector().get_local(), p=p.vector().get_local(), t_stokes=t1 - t0,
                   u_dof_coords=u.function_space().tabulate_dof_coordinates(),
                   p_dof_coords=p.function_space().tabulate_dof_coordinates())
    else:
        t1 = t0
        c = ref_solvers.pure_diffusion_solver(mesh_results, C, df.Constant(1.0), df.Constant(args.mu), 'sulcus')
    t2 = time.perf_counter()
    out.update(c=c.vector().get_local(), c_dof_coords=C.tabulate_dof_coordinates(), t_conc=t2 - t1,
               cell_dofs=np.array([C.dofmap().cell_dofs(i) for i in range(mesh.num_cells())]),
               coords=mesh.coordinates(), cells=mesh.cells())
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: stokes {t1 - t0:.2f} s, concentration {t2 - t1:.2f} s")
    return 0


if __name__ == '__main__':
    sys.exit(main())
