"""ctypes binding of ``libsulcusfem.so`` (C ABI declared in ``include/sulcusfem.h``).

The library is the product: there is no CPU fallback.  Importing this module works without a GPU
(the shared object only needs libcudart), but any compute entry point raises when no CUDA device
is present, and a missing / unbuilt library raises ``SulcusFemError`` at load time.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsulcusfem.so')
HEADER_PATH = os.path.normpath(os.path.join(_HERE, '..', '..', 'include', 'sulcusfem.h'))


class SulcusFemError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_d = C.c_double

# name -> (restype, argtypes); kept in the order of include/sulcusfem.h
SIGNATURES = {
    'sfem_last_error': (C.c_char_p, []),
    'sfem_version': (_i, []),
    'sfem_device_sms': (_i, []),
    'sfem_launch_count': (C.c_longlong, []),
    'sfem_launch_count_reset': (None, []),
    'sfem_profile_start': (_i, [_i]),
    'sfem_profile_stop': (_i, [_i, C.POINTER(_i), C.POINTER(_d), C.POINTER(C.c_float)]),
    'sfem_spmv_csr_f64': (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _i, _p]),
    'sfem_spmv_csr_f64_nb': (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    'sfem_staged_plan': (_i, [_i, _p, _i, _i, _p]),
    'sfem_staged_register': (_i, [_p, _i, _p, _i, _i, _i]),
    'sfem_staged_unregister': (None, [_p]),
    'sfem_staged_set_min_tiles': (_i, [_i]),
    'sfem_spmv_csr_f64_staged': (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    'sfem_sell_parts': (_i, []),
    'sfem_sell_register': (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p, C.c_longlong, _p, _i]),
    'sfem_sell_unregister': (None, [_p]),
    'sfem_sell_mark_dirty': (_i, [_p]),
    'sfem_sell_sync': (_i, [_p]),
    'sfem_sell_set_min_rows': (_i, [_i]),
    'sfem_spmv_csr_f64_sell': (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    'sfem_elem_p2_advdiff': (_i, [_i, _p, _p, _d, _p, _p, _p, _p]),
    'sfem_elem_p1_advdiff': (_i, [_i, _p, _p, _d, _p, _p, _i, _p, _p]),
    'sfem_elem_th_stokes': (_i, [_i, _p, _p, _p]),
    'sfem_elem_th_div': (_i, [_i, _p, _p, _p]),
    'sfem_csr_zero_flagged': (_i, [_i, _p, _p, _p, _p, _p, _p]),
    'sfem_elem_p1_mass': (_i, [_i, _p, _p, _p]),
    'sfem_facet_p2_robin': (_i, [_i, _p, _p, _d, _p, _i, _p, _p]),
    'sfem_facet_p1_robin': (_i, [_i, _p, _p, _d, _p, _i, _p, _p]),
    'sfem_gather_csr': (_i, [_i, _p, _p, _p, _p, _p]),
    'sfem_apply_dirichlet': (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _i, _p]),
    'sfem_csr_extract': (_i, [_i, _p, _p, _p, _p]),
    'sfem_vec_interleave2': (_i, [_i, _p, _p, _p, _p]),
    'sfem_vec_deinterleave2': (_i, [_i, _p, _p, _p, _p]),
    'sfem_vec_axpby': (_i, [_i, _d, _p, _d, _p, _p]),
    'sfem_vec_dot': (_i, [_i, _p, _p, C.POINTER(_d), _p]),
    'sfem_vec_set': (_i, [_i, _d, _p, _p]),
    'sfem_vec_pointwise_mul': (_i, [_i, _d, _p, _p, _p, _p]),
    'sfem_vec_select': (_i, [_i, _p, _p, _p, _p, _p]),
    'sfem_vec_copy': (_i, [_i, _p, _p, _p]),
    'sfem_dense_inverse_csr': (_i, [_i, _p, _p, _p, _p, _p]),
    'sfem_extract_diag_inv': (_i, [_i, _p, _p, _p, _p, _p]),
    'sfem_postprocess_concentration': (_i, [_i, _p, _i, C.POINTER(_d), _p]),
    'sfem_mg_create': (_p, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p),
                            C.POINTER(_i), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p),
                            C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), _p, _i, _d, _i]),
    'sfem_mg_setup': (_i, [_p, _p]),
    'sfem_mg_setup_fine': (_i, [_p, _p]),
    'sfem_mg_vcycle': (_i, [_p, _p, _p, _p]),
    'sfem_mg_set_tail': (_i, [_p, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p]),
    'sfem_mg_set_tail_rows': (_i, [_i]),
    'sfem_mg_lambda_max': (_i, [_p, C.POINTER(_d)]),
    'sfem_mg_destroy': (None, [_p]),
    'sfem_krylov_cg': (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _d, _i, C.POINTER(_d), _p]),
    'sfem_krylov_fgmres': (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _d, _i, _i, C.POINTER(_d), _p]),
    'sfem_krylov_cg_batch': (_i, [_i, _i, _p, _p, _i, C.POINTER(_p), C.POINTER(_p), _i, C.POINTER(_d), _d, _p, _p, _p, _p, _p,
                                  _d, _i, C.POINTER(_d), _p]),
    'sfem_batch_column': (_i, [_i, _i, _i, _p, _p, _p]),
    'sfem_stokes_create': (_p, [_i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p,
                                _i, _p, _p, _p, _p, _p, _p]),
    'sfem_stokes_solve': (_i, [_p, _p, _p, _d, _i, C.POINTER(_d), _p]),
    'sfem_stokes_solve_from': (_i, [_p, _p, _p, _p, _d, _i, C.POINTER(_d), _p]),
    'sfem_stokes_destroy': (None, [_p]),
    'sfem_dist_header_words': (C.c_longlong, [_i, C.c_longlong]),
    'sfem_dist_create': (_p, [_i, _i, C.c_longlong, C.c_longlong]),
    'sfem_dist_ipc_handle': (_i, [_p, _p]),
    'sfem_dist_open_peers': (_i, [_p, _p]),
    'sfem_dist_set_peer_pointer': (_i, [_p, _i, _p]),
    'sfem_dist_mailbox': (_p, [_p]),
    'sfem_dist_activate': (_i, [_p]),
    'sfem_dist_error': (_i, [_p]),
    'sfem_dist_destroy': (None, [_p]),
    'sfem_halo_create': (_p, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'sfem_halo_destroy': (None, [_p]),
    'sfem_halo_attach': (_i, [_p, _p]),
    'sfem_halo_exchange': (_i, [_p, _p, _i, _i, _p]),
    'sfem_dist_allreduce_vec': (_i, [_p, _p, _i, _i, _p]),
    'sfem_dist_allreduce_scalars': (_i, [_p, _p, _i, _p]),
    'sfem_facet_functionals': (_i, [_i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _d, _d, _p, _p, _p]),
    'sfem_cell_functionals': (_i, [_i, _p, _p, _p, _i, _p, _p, _p]),
    'sfem_stokes_create_part': (_p, [_i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p,
                                     _i, _p, _p, _p, _p, _p, _p, _i, C.c_longlong, C.c_longlong]),
    'sfem_eval_points': (_i, [_i, _i, _p, _i, _i, _d, _d, _d, _d, _p, _p, _p, _i, _p, _i, C.POINTER(_p), _d, _p, _p, _p]),
}

_lib = None


def header_symbols():
    """Function names declared in include/sulcusfem.h (used by the CPU symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sfem_[a-z0-9_]+)\s*\(', text)))


def load():
    """Load the shared library and attach prototypes; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SulcusFemError(
            f"{LIB_PATH} not found: build it with fenics-eff-uptake_b200/csrc/build.sh "
            "(or __graft_entry__.build()); there is no CPU fallback")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise SulcusFemError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SulcusFemError(f"libsulcusfem.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().sfem_last_error()
        raise SulcusFemError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
