"""ctypes binding of libfers_b200.so (include/fers_b200.h).  Fails loudly when the library is
missing: there is no CPU fallback anywhere in the product path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfers_b200.so")

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); void* is used for every device pointer
SIGNATURES = {
    "fers_b200_version": (C.c_char_p, []),
    "fers_b200_last_error": (C.c_char_p, []),
    "fers_b200_device_info": (_i, [_i, _ip, _ip, _ip, C.POINTER(_sz)]),
    "fers_b200_set_device": (_i, [_i]),
    "fers_b200_launch_count": (C.c_longlong, []),
    "fers_b200_pcg_profile": (_i, [_dp]),
    "fers_b200_fp64_peak": (_i, [_d, _dp, _dp, _vp, _sz, _vp]),
    "fers_b200_pattern_workspace": (_sz, [_i, _i, _i]),
    "fers_b200_pattern_build": (_i, [_i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _ip, _vp, _sz, _vp]),
    "fers_b200_bc_workspace": (_sz, [_i]),
    "fers_b200_json_records_bound": (_sz, [_sz, _i, C.c_longlong, _sz]),
    "fers_b200_json_records": (_i, [C.c_char_p, _sz, _i, C.c_longlong, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "fers_b200_host_pattern": (_i, [_i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _ip]),
    "fers_b200_bc_build": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ip, _ip, _vp, _sz, _vp]),
    "fers_b200_assemble_beams": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i,
                                      _vp, _vp, _i, _vp]),
    "fers_b200_assemble_plates": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_assemble_set_gather_kernel": (_i, [_i]),
    "fers_b200_zero": (_i, [_vp, _sz, _vp]),
    "fers_b200_assemble_warping": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_assemble_cables": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_cable_rhs": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "fers_b200_recover_warping": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_gather_map_workspace": (_sz, [_i, _i]),
    "fers_b200_gather_map_build": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _ip, _vp, _sz, _vp]),
    "fers_b200_assemble_beams_gather": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i,
                                             _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "fers_b200_load_case": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_combine": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "fers_b200_set_spmm_variant": (_i, [_i]),
    "fers_b200_spmm": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_block_jacobi": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_pcg_workspace": (_sz, [_i, _i]),
    "fers_b200_pcg_solve": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _i, _ip, _dp,
                                 _vp, _sz, _vp]),
    "fers_b200_dense_expand": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_dense_expand_batched": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_dense_cholesky": (_i, [_i, _i, _vp, _vp, _ip, _vp]),
    "fers_b200_dense_solve": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "fers_b200_frame_batch_factor_solve": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_frame_batch_step": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp, _vp, _vp]),
    "fers_b200_member_axial_masked": (_i, [_i, _i, _vp, _vp, _vp, _vp]),
    "fers_b200_recover_members": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_recover_sections": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp]),
    "fers_b200_reactions": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_recover_plates": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_member_axial": (_i, [_i, _i, _vp, _vp, _vp]),
    # sparse multifrontal Cholesky: host plan ([host] pointers are passed as integers / ctypes arrays)
    "fers_b200_chol_plan_create": (_i, [_i, _vp, _vp, _vp, _i, _vp, C.POINTER(_vp)]),
    "fers_b200_chol_plan_destroy": (None, [_vp]),
    "fers_b200_chol_plan_info": (_i, [_vp, C.POINTER(C.c_longlong), _i]),
    "fers_b200_chol_plan_perm": (_i, [_vp, _vp]),
    "fers_b200_chol_plan_levels": (_i, [_vp, C.POINTER(C.c_longlong), _i]),
    "fers_b200_chol_plan_export": (_i, [_vp, _vp, C.POINTER(C.c_longlong), _vp, _vp]),
    "fers_b200_chol_factor_bytes": (_sz, [_vp]),
    "fers_b200_chol_work_bytes": (_sz, [_vp, _i]),
    "fers_b200_chol_upload": (_i, [_vp, _vp, _vp]),
    "fers_b200_chol_factor": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _ip, _vp]),
    "fers_b200_chol_solve": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp]),
    "fers_b200_residual": (_i, [C.c_longlong, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_unity_generic": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fers_b200_chol_set_profile": (_i, [_i]),
    "fers_b200_chol_plan_partition": (_i, [_vp, _i, _i, C.POINTER(C.c_longlong), _i]),
    "fers_b200_chol_factor_levels": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _vp]),
    "fers_b200_chol_solve_phase": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _sz, _i, _i, _i, _vp]),
    "fers_b200_chol_set_variant": (_i, [_i]),
    "fers_b200_chol_set_outer": (_i, [_i, _i]),
    "fers_b200_chol_profile": (_i, [_dp, _i]),
    "fers_b200_chol_set_ozaki": (_i, [_i, _i]),
    "fers_b200_oz_syrk_work_bytes": (_sz, [_i, _i, _i]),
    "fers_b200_oz_syrk": (_i, [_vp, C.c_longlong, _vp, C.c_longlong, _i, _i, _i, _vp, _sz, _vp]),
}


class FersB200Error(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the declared signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FersB200Error(
            f"{LIB_PATH} is missing: build it with `python -m fers_core_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fers_b200_last_error().decode("utf-8", "replace")
        raise FersB200Error(f"{what} failed (status {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()
