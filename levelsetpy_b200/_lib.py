"""ctypes binding of the C-ABI in include/hjb200.h (levelsetpy_b200/_hjb200.so).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the product
path raises.  (`python -m levelsetpy_b200.build` / `__graft_entry__.build()` compiles it in-tree.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_hjb200.so")

HJ_MAX_DIM, HJ_MAX_PARAMS, HJ_MAX_TABLES, HJ_GHOST = 6, 96, 8, 3
HJ_HALO_DESC_BYTES = 512
HJ_OK, HJ_ERR_INVALID, HJ_ERR_CUDA, HJ_ERR_UNSUPPORTED, HJ_ERR_STATE, HJ_ERR_NAN = 0, -1, -2, -3, -4, -5
BC_EXTRAPOLATE, BC_PERIODIC, BC_HALO = 0, 1, 2
WENO_AS_SHIPPED, WENO_INTENDED, SCHEME_ENO3A, SCHEME_ENO2 = 0, 1, 2, 3
SYS_DUBINS_REL, SYS_DOUBLE_INT, SYS_FLOCK, SYS_DUBINS_REL_PAIR, SYS_DOUBLE_INT_PAIR = 1, 2, 3, 4, 5
SYS_GENERIC_DUBINS_CAR = 6
COMP_NONE, COMP_MIN_OVER_TIME, COMP_MAX_OVER_TIME, COMP_MIN_WITH_AUX, COMP_MAX_WITH_AUX = 0, 1, 2, 3, 4
FIELD_STATE, FIELD_AUX, FIELD_OBSTACLE = 0, 1, 2
BACKEND_AUTO, BACKEND_GATHER, BACKEND_TMA = 0, 1, 2

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pd, _pi, _pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol declared in include/hjb200.h
SIGNATURES = {
    "hj_version": (C.c_char_p, []),
    "hj_last_error": (C.c_char_p, []),
    "hj_launch_count": (_i64, []),
    "hj_create": (_i, [C.POINTER(_vp), _i, _i, _pi64, _pd, _pi, _pi, _i]),
    "hj_destroy": (_i, [_vp]),
    "hj_set_backend": (_i, [_vp, _i]),
    "hj_set_axis": (_i, [_vp, _i, _vp, _i64]),
    "hj_set_table": (_i, [_vp, _i, _vp, _i64]),
    "hj_set_system": (_i, [_vp, _i, _vp, _i]),
    "hj_upload": (_i, [_vp, _vp, _i, _vp, _i]),
    "hj_download": (_i, [_vp, _vp, _i, _vp, _i]),
    "hj_num_nodes": (_i64, [_vp]),
    "hj_field_elems": (_i64, [_vp]),
    "hj_state_ptr": (_i, [_vp, _i, C.POINTER(_vp)]),
    "hj_plane_elems": (_i64, [_vp]),
    "hj_deriv": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "hj_deriv_candidates": (_i, [_vp, _vp, _vp, _i, _vp]),
    "hj_add_ghost": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "hj_rhs": (_i, [_vp, _vp, _d, _vp, _vp, _pd, _pd]),
    "hj_ham": (_i, [_vp, _vp, _d, C.POINTER(_vp), _vp]),
    "hj_alpha": (_i, [_vp, _vp, _d, _i, _vp]),
    "hj_diss_glf": (_i, [_vp, _vp, _d, C.POINTER(_vp), C.POINTER(_vp), _vp, _pd, _pd]),
    "hj_alpha_max": (_i, [_vp, _vp, _d, _pd, _pd]),
    "hj_step": (_i, [_vp, _vp, _d, _d, _vp, _i, _i, _i]),
    "hj_step_reductions": (_i, [_vp, _vp, _pd]),
    "hj_stage": (_i, [_vp, _vp, _i, _d, _d, _vp, _i, _i, _i]),
    "hj_stage_pass": (_i, [_vp, _vp, _i, _i, _d, _d, _vp, _i, _i, _i]),
    "hj_is_split": (_i, [_vp]),
    "hj_split_cols": (_i, [_vp, _pi64, _pi]),
    "hj_stage_pass_cols": (_i, [_vp, _vp, _i, _i64, _i64, _d, _d, _vp, _i, _i, _i]),
    "hj_stage_range": (_i, [_vp, _vp, _i, _i64, _i64, _d, _d, _vp, _i, _i, _i]),
    "hj_stage_io": (_i, [_vp, _i, _pi, _pi]),
    "hj_eps_prepass": (_i, [_vp, _vp, _i, C.POINTER(_vp)]),
    "hj_fill_edge_halo": (_i, [_vp, _vp, _i, _i]),
    "hj_device_count": (_i, []),
    "hj_dev_alloc": (_i, [_i, _i64, C.POINTER(_vp)]),
    "hj_dev_free": (_i, [_vp]),
    "hj_memcpy": (_i, [_vp, _vp, _i64, _i, _vp, _i]),
    "hj_stream_sync": (_i, [_vp]),
    "hj_ode_cfl3_single": (_i, [_vp, _vp, _d, _d, _d, _d, _vp, _i, _i, _i, _pd, _pd]),
    "hj_ode_cfl3_step": (_i, [_vp, _vp, _d, _d, _d, _d, _vp, _vp, _i, _i, _i, _pd, _pd]),
    "hj_host_alloc": (_i, [_i64, C.POINTER(_vp)]),
    "hj_host_free": (_i, [_vp]),
    "hj_snapshot": (_i, [_vp, _vp]),
    "hj_change": (_i, [_vp, _vp, _pd, _pi]),
    "hj_discount": (_i, [_vp, _vp, _d, _i, _i, _d]),
    "hj_set_restrict": (_i, [_vp, _i]),
    "hj_set_pipeline_planes": (_i, [_vp, _i]),
    "hj_deriv_range": (_i, [_vp, _vp, _vp, _i, _pd, _pd]),
    "hj_step_rk2": (_i, [_vp, _vp, _d, _d, _vp, _i, _i, _i]),
    "hj_create_batch": (_i, [C.POINTER(_vp), _i, _i, _i, _pi64, _pd, _pi, _pi, _i]),
    "hj_step_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "hj_batch_size": (_i, [_vp]),
    "hj_halo_export": (_i, [_vp, _vp]),
    "hj_halo_attach": (_i, [_vp, _vp, _vp]),
    "hj_halo_detach": (_i, [_vp]),
    "hj_halo_attached": (_i, [_vp]),
    "hj_halo_set_fused": (_i, [_vp, _i]),
    "hj_halo_signal": (_i, [_vp, _vp, _i, _i]),
    "hj_halo_push": (_i, [_vp, _vp, _i, _i, _i64, _i64, _i64]),
    "hj_halo_wait": (_i, [_vp, _vp, _i, _i]),
}

_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "levelsetpy_b200: CUDA library %s not built (run `python -m levelsetpy_b200.build`); "
                "there is no CPU fallback" % SO_PATH)
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class HJError(RuntimeError):
    pass


def check(code):
    """Turn a negative hj_status into the exception the reference would raise (ValueError for bad
    arguments -- Utilities/matlab_utils.py `error()` -- RuntimeError otherwise)."""
    if code == HJ_OK:
        return
    msg = load().hj_last_error().decode("utf-8", "replace")
    if code in (HJ_ERR_INVALID, HJ_ERR_NAN):
        raise ValueError(msg)
    if code == HJ_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise HJError("hjb200 error %d: %s" % (code, msg))
