"""``MeshGenerator`` with the reference's constructor and result dictionary (reference mesh.py:29-599).

The reference writes a ``.geo`` file, shells out to Gmsh, converts with meshio and loads a
``dolfin.Mesh``.  None of these exist here; this class produces the same *interface* --
``generate_mesh()`` returns ``{'mesh', 'bc_markers', ['bottom_segment_markers', 'y0_markers',
'domain_markers'], 'mesh_info'}`` with the marker semantics of ``mesh.py:196-256,425-453`` -- from
the in-process meshers of :mod:`sulcusfem.unstructured` / :mod:`sulcusfem.hostmesh`, or from a mesh
file the reference pipeline already wrote (``mesh_file=`` dolfin-XML or Gmsh msh2).
``refinement_factor`` > 1 grades the Delaunay mesh towards the sulcus with the reference's Distance / Threshold size
field (``lc_fine = mesh_size / refinement_factor`` within w/10 of the sulcus nodes, ``mesh_size`` beyond w/2;
``mesh.py:266,330-337`` -> ``unstructured.threshold_size_field``); ``uniform_refinements`` (not a reference argument)
adds uniform red refinements of the whole mesh, the knob BASELINE config 5 (mesh convergence) turns.
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import hostmesh as hm
from .hostmesh import DOLFIN_EPS, MARKERS, TOLERANCE  # noqa: F401  (re-exported like the reference module)


class MeshGenerator:
    SULCUS_BASE, RECT_BASE = "sulcus_mesh", "rect_mesh"
    N_SULCUS_SEGMENTS = 20
    MARKERS = dict(MARKERS)
    TOLERANCE = TOLERANCE

    def __init__(self, width, height, sulcus_depth, sulcus_width, mesh_size, refinement_factor, domain_type,
                 output_dir=None, mesher='delaunay', uniform_refinements=0, mesh_file=None):
        self.output_dir = os.path.abspath(output_dir) if output_dir else None
        self.width, self.height = width, height
        self.sulcus_depth, self.sulcus_width = sulcus_depth, sulcus_width
        self.mesh_size, self.refinement_factor = mesh_size, refinement_factor
        self.sulcus_left_x = width / 2 - sulcus_width / 2
        self.sulcus_right_x = width / 2 + sulcus_width / 2
        self.domain_type = domain_type
        self.mesher, self.uniform_refinements, self.mesh_file = mesher, int(uniform_refinements), mesh_file
        self._validate_parameters()
        self.sulcus_mesh = self.rect_mesh = None

    def _validate_parameters(self):
        checks = [
            (self.height > 0, "Channel height must be positive"),
            (self.width > 0, "Channel width must be positive"),
            (self.mesh_size > 0, "Mesh size must be positive"),
            (self.sulcus_width > 0, "Sulcus width must be positive"),
            (self.sulcus_depth > 0, "Sulcus depth must be positive"),
            (self.refinement_factor > 0, "Refinement factor must be positive"),
            (self.sulcus_width < self.width, "Sulcus width must be less than channel width"),
            (self.domain_type in ('sulcus', 'rectangular'), "domain_type must be one of ['sulcus', 'rectangular']"),
        ]
        for ok, message in checks:
            if not ok:
                raise ValueError(message)

    # ------------------------------------------------------------------ geometry helpers
    def sulcus_points(self):
        """The 21 floor samples the reference hands to Gmsh (mesh.py:139-155), rounded like its '%.6f'."""
        n = self.N_SULCUS_SEGMENTS
        x_rel = np.arange(n + 1) / n
        x = self.sulcus_left_x + x_rel * self.sulcus_width
        y = -self.sulcus_depth * np.sin(np.pi * x_rel)
        y[0] = y[-1] = 0.0
        return np.round(np.stack([x, y], axis=1), 6)

    def _build_host_mesh(self):
        if self.mesh_file:
            ext = os.path.splitext(self.mesh_file)[1].lower()
            mesh = hm.read_dolfin_xml(self.mesh_file) if ext == '.xml' else hm.read_gmsh_msh2(self.mesh_file)
            mesh.geometry = {'domain_type': self.domain_type, 'L': float(self.width), 'H': float(self.height)}
            if self.domain_type == 'sulcus':
                mesh.geometry.update({'w': float(self.sulcus_width), 'd': float(self.sulcus_depth),
                                      'xL': float(self.sulcus_left_x), 'xR': float(self.sulcus_right_x)})
            return mesh
        h = self.mesh_size
        mesh = None
        if self.mesher == 'delaunay':
            from .unstructured import mesh_domain
            try:
                mesh = mesh_domain(self.width, self.height, self.sulcus_width, self.sulcus_depth, h, self.domain_type,
                                   refinement_factor=self.refinement_factor)
            except RuntimeError as e:                      # tiny cavities: fall back to the structured mesher
                logging.warning(f"unstructured mesher failed ({e}); using the structured mesher")
        if mesh is None:
            if self.domain_type == 'sulcus':
                mesh = hm.sulcus_mesh(self.width, self.height, self.sulcus_width, self.sulcus_depth, h)
            else:
                mesh = hm.rectangle_mesh(self.width, self.height, max(1, int(round(self.width / h))),
                                         max(1, int(round(self.height / h))))
        return hm.refine_n(mesh, self.uniform_refinements)

    def generate_mesh(self):
        mesh = self._build_host_mesh()
        markers = hm.build_markers(mesh, self.width, self.height, self.sulcus_left_x, self.sulcus_right_x, self.domain_type)
        info = {"num_vertices": int(mesh.num_vertices), "num_cells": int(mesh.num_cells),
                "hmin": mesh.hmin(), "hmax": mesh.hmax()}
        out = {"mesh": mesh, "bc_markers": markers['bc_markers'], "mesh_info": info}
        if self.domain_type == 'sulcus':
            self.sulcus_mesh = mesh
            out.update({k: markers[k] for k in ('bottom_segment_markers', 'y0_markers', 'domain_markers')})
            self.sulcus_bc_markers = markers['bc_markers']
        else:
            self.rect_mesh = mesh
            self.rect_bc_markers = markers['bc_markers']
        return out

    def save_mesh_pvd_files(self, pvd_output_dir):
        """ParaView export of meshes / normals (mesh.py:600-713) is visualisation only: not provided."""
        return None


def precompute_device_plans(mesh_results, hier, stokes=True):
    """Build (and memoise on the mesh objects) every CSR pattern + gather map the device problems of this geometry
    will ask for: the P2 system level and the P1 multigrid levels of the concentration problem (Robin id 4) and, with
    ``stokes``, of the velocity block (no Robin boundary), plus the Stokes divergence / pressure-mass plans.  Mirrors
    the level loop of ``device.ScalarProblem`` / ``device.StokesProblem``; a key that is not asked for later costs
    nothing but the time spent here."""
    from . import dofmap as dm
    from .hierarchy import level_markers
    mesh = mesh_results['mesh']
    bm = mesh_results['bc_markers'].values
    for robin in ((4, None) if stokes else (4,)):
        dm.scalar_level_plan(mesh, bm, 2, robin)
        for m in hier.meshes:
            mk = bm if m is mesh else level_markers(m)['bc_markers'].values
            dm.scalar_level_plan(m, mk, 1, robin)
    if stokes:
        dm.stokes_block_plans(mesh)


def generate_mesh_job(kwargs, with_hierarchy=True, plans=None):
    """Worker-process entry of ``simulation.prefetch_meshes``: mesh + markers (+ multigrid hierarchy, + the host plans
    of the device problems: ``plans`` = None / 'scalar' / 'stokes') of one geometry.  Pure host work (numpy / Qhull);
    nothing here touches CUDA or torch, so it is safe in spawned processes."""
    out = MeshGenerator(**kwargs).generate_mesh()
    hier = None
    if with_hierarchy and out:
        from .hierarchy import build_hierarchy
        hier = build_hierarchy(out['mesh'])
        if plans:
            precompute_device_plans(out, hier, stokes=(plans == 'stokes'))
    return out, hier


class Measure:
    """Descriptor of an integration measure: kind ('ds' | 'dS' | 'dx') + the marker set it reads."""

    def __init__(self, kind, markers=None, mesh=None):
        self.kind, self.markers, self.mesh = kind, markers, mesh

    def __call__(self, marker_id):
        return (self, int(marker_id))


def setup_sulcus_measures(mesh, bc_markers, bottom_segment_markers, y0_markers, domain_markers):
    """Same six-tuple as reference mesh.py:721-730."""
    return (Measure('ds', bc_markers, mesh), Measure('ds', bottom_segment_markers, mesh),
            Measure('dS', bottom_segment_markers, mesh), Measure('ds', y0_markers, mesh),
            Measure('dS', y0_markers, mesh), Measure('dx', domain_markers, mesh))


def setup_rectangular_measures(mesh, bc_markers):
    """Same pair as reference mesh.py:732-737."""
    return Measure('ds', bc_markers, mesh), Measure('dx', None, mesh)
