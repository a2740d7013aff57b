"""CPU: the C-ABI library loads and exports every symbol include/sulcusfem.h declares (no compute)."""
import ctypes
import os

import pytest

from sulcusfem import capi


def test_library_exports_header_symbols():
    lib = capi.load()
    names = capi.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sulcusfem.h but not exported"
    assert set(names) == set(capi.SIGNATURES), "ctypes prototypes out of sync with the header"
    assert lib.sfem_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sulcusfem.device import Context
    with pytest.raises(capi.SulcusFemError):
        Context()
    from sulcusfem import hostmesh as hm, solvers
    from sulcusfem.fem import Constant, FunctionSpace
    m = hm.rectangle_mesh(10.0, 1.0, 10, 2)
    mk = hm.build_markers(m, 10.0, 1.0, 4.75, 5.25, 'rectangular')
    with pytest.raises(capi.SulcusFemError):
        solvers.pure_diffusion_solver({'mesh': m, 'bc_markers': mk['bc_markers']}, FunctionSpace(m, 'CG', 2),
                                      Constant(1.0), Constant(1.0))


def test_product_code_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'fenics-eff-uptake_b200', 'sulcusfem')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            with open(os.path.join(pkg, fn)) as f:
                src = f.read()
            assert 'cpu_oracle' not in src and 'from oracle' not in src and 'import oracle' not in src, fn
