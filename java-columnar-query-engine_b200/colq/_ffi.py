"""ctypes binding of ``libcolq.so`` -- one declaration per symbol of ``include/colq.h``.

This is the Python twin of the java.lang.foreign binding in ``java/``: no glue logic, just
the downcall signatures.  Loading fails loudly when the shared library is missing; there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("COLQ_LIB", PKG_DIR / "lib" / "libcolq.so"))

# colq_status
OK, FAILURE, THROW_INDEX_OOB, THROW_NULL, THROW_ILLEGAL_STATE, THROW_ILLEGAL_ARG, ERR_DEVICE, ERR_CAPACITY = range(8)
# colq_placement
REPLICATED, SHARDED = 0, 1
# colq_option
OPT_LAZY_FK, OPT_PROFILE, OPT_PEER_EXCHANGE, OPT_FUSED_COMPACT, OPT_DEFER_CHAINS, OPT_PROMOTE, OPT_FUSED_GATHER, OPT_TAIL_PUBLISH, OPT_ROOT_FUSED, OPT_LAZY_GATHER_WAIT, OPT_PIPELINE = 0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11
ABI_VERSION = 2


class Timing(C.Structure):
    _fields_ = [("gpu_ms", C.c_double), ("kernel_launches", C.c_int32), ("collectives", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class Stage(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_double), ("rows", C.c_int64), ("bytes", C.c_int64)]


_p = C.c_void_p
_i32, _i64, _int = C.c_int32, C.c_int64, C.c_int

# every exported symbol of include/colq.h: name -> (restype, argtypes)
SIGNATURES = {
    "colq_abi_version": (_int, []),
    "colq_build_id": (C.c_char_p, []),
    "colq_create": (_int, [_int, C.POINTER(_p)]),
    "colq_destroy": (_int, [_p]),
    "colq_last_error": (C.c_char_p, [_p]),
    "colq_set_stream": (_int, [_p, _p]),
    "colq_get_stream": (_int, [_p, C.POINTER(_p)]),
    "colq_synchronize": (_int, [_p]),
    "colq_trim": (_int, [_p]),
    "colq_comm_unique_id": (_int, [_p, _p]),
    "colq_comm_init": (_int, [_p, _p, _int, _int]),
    "colq_comm_info": (_int, [_p, C.POINTER(_int), C.POINTER(_int)]),
    "colq_comm_init_local": (_int, [C.POINTER(_p), _int]),
    "colq_execute_group": (_int, [C.POINTER(_p), C.POINTER(_p), _int]),
    "colq_fetch_group": (_int, [C.POINTER(_p), C.POINTER(_p), _int, C.POINTER(_i64)]),
    "colq_table_create": (_int, [_p, _i64, _int, _i64, C.POINTER(_i32)]),
    "colq_register": (_int, [_p, C.c_char_p, _i32]),
    "colq_col_i32": (_int, [_p, _i32, _int, _p, _i64]),
    "colq_col_str": (_int, [_p, _i32, _int, _p, _p, _i64, _i64]),
    "colq_col_bool": (_int, [_p, _i32, _int, _p, _i64]),
    "colq_col_i32_device": (_int, [_p, _i32, _int, _p, _i64]),
    "colq_col_str_device": (_int, [_p, _i32, _int, _p, _i64, _p, _i64, _i64, _i64]),
    "colq_host_alloc": (_int, [_p, _i64, C.POINTER(_p)]),
    "colq_host_free": (_int, [_p, _p]),
    "colq_host_register": (_int, [_p, _p, _i64]),
    "colq_host_unregister": (_int, [_p, _p]),
    "colq_col_i32_host": (_int, [_p, _i32, _int, _p, _i64, _i64]),
    "colq_col_str_host": (_int, [_p, _i32, _int, _p, _i64, _p, _i64, _i64, _i64]),
    "colq_associate_fk_host": (_int, [_p, _i32, _int, _i32, _int, _p, _i64, _i64]),
    "colq_col_i32_dict": (_int, [_p, _i32, _int, _p, _i64, _p, _i64]),
    "colq_col_i32_dict_host": (_int, [_p, _i32, _int, _p, _i64, _i64, _p, _i64]),
    "colq_col_str_dict": (_int, [_p, _i32, _int, _p, _i64, _p, _p, _i64, _i64]),
    "colq_col_str_dict_device": (_int, [_p, _i32, _int, _p, _i64, _p, _p, _i64, _i64]),
    "colq_col_str_dict_host": (_int, [_p, _i32, _int, _p, _i64, _i64, _p, _p, _i64, _i64]),
    "colq_associate_fk": (_int, [_p, _i32, _int, _i32, _int, _p, _i64]),
    "colq_associate_csr": (_int, [_p, _i32, _int, _i32, _int, _p, _p, _i64, _i64]),
    "colq_associate_fk_device": (_int, [_p, _i32, _int, _i32, _int, _p, _i64]),
    "colq_associate": (_int, [_p, _i32, _int, _i32, _int, _p, _p, _i64, _i64, C.POINTER(_int)]),
    "colq_associate_device": (_int, [_p, _i32, _int, _i32, _int, _p, _p, _i64, _i64, C.POINTER(_int)]),
    "colq_col_str_encode": (_int, [_p, _i32, _int, C.POINTER(_i64)]),
    "colq_col_dict_str": (_int, [_p, _i32, _int, _p, _i64, _p, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "colq_table_partition": (_int, [_p, _i32, _p, _int]),
    "colq_associate_fk_global": (_int, [_p, _i32, _int, _i32, _int, _p, _i64]),
    "colq_associate_csr_global": (_int, [_p, _i32, _int, _i32, _int, _p, _p, _i64, _i64]),
    "colq_table_destroy": (_int, [_p, _i32]),
    "colq_table_size": (_int, [_p, _i32, C.POINTER(_i64)]),
    "colq_table_width": (_int, [_p, _i32, C.POINTER(_int)]),
    "colq_query_create": (_int, [_p, C.c_char_p, C.POINTER(_p)]),
    "colq_query_destroy": (_int, [_p]),
    "colq_query_child": (_int, [_p, _int, _int, C.POINTER(_int)]),
    "colq_query_criteria_i32_range": (_int, [_p, _int, _int, _i32, _i32]),
    "colq_query_criteria_str": (_int, [_p, _int, _int, _int, _p, _i32]),
    "colq_query_criteria_str_accept": (_int, [_p, _int, _int, _p, _i64]),
    "colq_query_criteria_i32_accept": (_int, [_p, _int, _int, _p, _i64]),
    "colq_query_criteria_bool": (_int, [_p, _int, _int, _int, _int]),
    "colq_query_set_option": (_int, [_p, _int, _int]),
    "colq_execute": (_int, [_p, _p, _p, _i64, _p, _i64, C.POINTER(_i64), C.POINTER(Timing)]),
    "colq_execute_async": (_int, [_p, _p]),
    "colq_fetch": (_int, [_p, _p, _p, _i64, _p, _i64, C.POINTER(_i64), C.POINTER(Timing)]),
    "colq_profile": (_int, [_p, C.POINTER(Stage), _int, C.POINTER(_int)]),
    "colq_result_count": (_int, [_p, _p, C.POINTER(_i64)]),
    "colq_result_i32": (_int, [_p, _p, _int, _p, _i64, C.POINTER(_i64)]),
    "colq_result_bool": (_int, [_p, _p, _int, _p, _i64, C.POINTER(_i64)]),
    "colq_result_str": (_int, [_p, _p, _int, _p, _i64, _p, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "colq_result_csr": (_int, [_p, _p, _int, _p, _i64, _p, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "colq_profile_hot": (_int, [_p, C.POINTER(Stage), C.POINTER(_int)]),
    "colq_node_cardinalities": (_int, [_p, _p, C.POINTER(_i64), _int, C.POINTER(_int)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libcolq.so and declare every signature.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"libcolq.so not found at {LIB_PATH}. Build it with `make -C {PKG_DIR}` (or python -c "
            "'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
