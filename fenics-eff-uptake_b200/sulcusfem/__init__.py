"""sulcusfem -- B200-native finite-element solve path for the sulcus transport model.

Host side (numpy): meshes, markers, DOF maps, sparsity and gather maps, multigrid hierarchy.
Device side: ``libsulcusfem.so`` (hand-written sm_100a CUDA behind a C ABI, ``include/sulcusfem.h``).
Front end: ``sulcusfem.solvers`` / ``sulcusfem.analysis`` mirror the reference's entry points.
"""
__version__ = "0.1.0"
