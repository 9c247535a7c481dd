"""ctypes binding of libdeepfwfm_sm100a.so (the C ABI declared in include/deepfwfm_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing this module
raises at import of the symbol table, and every compute entry point needs an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdeepfwfm_sm100a.so")

DFW_ABI_VERSION = 3
DFW_MAX_DEPTH = 8
DFW_MAX_FIELDS = 64
DFW_MAX_K = 32
DFW_MAX_RANKS = 8

USE_FWFM, USE_FWLW, USE_LW, USE_DEEP, CHECK_INDEX, XI_INT32 = 1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 8, 1 << 9
HINT_THROUGHPUT = 1 << 10
TABLE_PLAIN, TABLE_QR_MULT, TABLE_QR_ADD = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_FP32_CSR, PREC_BF16X3 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp32_csr": PREC_FP32_CSR, "bf16x3": PREC_BF16X3}


class FieldDesc(C.Structure):
    _fields_ = [("w2", C.c_void_p), ("w2_r", C.c_void_p), ("w1", C.c_void_p), ("w1_r", C.c_void_p),
                ("rows", C.c_int64), ("collisions", C.c_int32), ("qr_op", C.c_int32),
                ("qr1_op", C.c_int32), ("n_ranks", C.c_int32),
                ("w2_shard", C.c_void_p * DFW_MAX_RANKS)]


class PruneSpan(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("count", C.c_int64)]


class Csr(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
                ("nnz", C.c_int32), ("max_row_nnz", C.c_int32)]


class Model(C.Structure):
    _fields_ = [("struct_bytes", C.c_uint32), ("abi_version", C.c_uint32), ("flags", C.c_uint32),
                ("field_size", C.c_int32), ("numerical", C.c_int32), ("embedding_size", C.c_int32),
                ("depth", C.c_int32), ("widths", C.c_int32 * DFW_MAX_DEPTH),
                ("fields", C.c_void_p), ("fwfm_linear", C.c_void_p), ("fm_1st", C.c_void_p),
                ("field_cov", C.c_void_p), ("bias", C.c_void_p),
                ("W", C.c_void_p * DFW_MAX_DEPTH), ("b", C.c_void_p * DFW_MAX_DEPTH), ("fc", C.c_void_p),
                ("Wbf16", C.c_void_p * DFW_MAX_DEPTH), ("Wbf16_lo", C.c_void_p * DFW_MAX_DEPTH),
                ("csr", Csr * DFW_MAX_DEPTH),
                ("shallow_image", C.c_void_p), ("field_cov_host", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/deepfwfm_b200.h declares
_vp, _i64, _i32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t
_MP = C.POINTER(Model)
SYMBOLS = {
    "dfw_version": (C.c_int, []),
    "dfw_last_error_string": (C.c_char_p, []),
    "dfw_check_device": (C.c_int, [C.c_int]),
    "dfw_struct_bytes": (_sz, [C.c_int]),
    "dfw_launch_count": (_i64, []),
    "dfw_embed_fwfm": (C.c_int, [_MP, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "dfw_shallow_image_bytes": (_sz, [_MP]),
    "dfw_pack_shallow": (C.c_int, [_MP, _vp, _vp]),
    "dfw_mlp_workspace_bytes": (_sz, [_MP, _i64, C.c_int]),
    "dfw_mlp_fp32": (C.c_int, [_MP, _vp, _i64, _i64, _vp, _vp, _sz, _vp, _vp, _vp]),
    "dfw_mlp_csr": (C.c_int, [_MP, _vp, _i64, _i64, _vp, _vp, _sz, _vp, _vp, _vp]),
    "dfw_mlp_bf16": (C.c_int, [_MP, _vp, _i64, _i64, _vp, _vp, _sz, _vp, _vp, _vp]),
    "dfw_finish_shallow": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "dfw_pack_mlp_bf16_bytes": (_sz, [_i32, _i32]),
    "dfw_pack_mlp_bf16": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "dfw_pack_mlp_bf16_split": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "dfw_csr_count": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "dfw_csr_fill": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dfw_forward_workspace_bytes": (_sz, [_MP, _i64, C.c_int]),
    "dfw_forward": (C.c_int, [_MP, _vp, _i64, _i64, _vp, _i64, _i64, _i64, C.c_int, _vp, _sz, _vp, _vp, _vp, _vp]),
    "dfw_fused_supported": (C.c_int, [_MP, C.c_int]),
    "dfw_forward_fused": (C.c_int, [_MP, _vp, _i64, _i64, _vp, _i64, _i64, _i64, C.c_int, _vp, _vp, _vp, _vp]),
    "dfw_forward_host_workspace_bytes": (_sz, [_MP, _i64, C.c_int]),
    "dfw_forward_host": (C.c_int, [_MP, _vp, _vp, _i64, C.c_int, _vp, _sz, _vp, _vp, _vp]),
    "dfw_forward_host_stream_workspace_bytes": (_sz, [_MP, _i64, C.c_int]),
    "dfw_forward_host_stream": (C.c_int, [_MP, _vp, _vp, _i64, _i64, C.c_int, _vp, _sz, _vp, _vp, _vp]),
    "dfw_host_transport_is_mapped": (C.c_int, [_MP, C.c_int, _vp, _vp, _vp, _vp]),
    "dfw_pull_rows": (C.c_int, [_MP, C.POINTER(C.c_int32), C.c_int32, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "dfw_prune_workspace_bytes": (_sz, []),
    "dfw_prune_threshold": (C.c_int, [C.POINTER(PruneSpan), C.c_int, C.c_int, C.c_double, _i64, _vp, _sz, _vp, _vp, _vp]),
    "dfw_prune_apply": (C.c_int, [C.POINTER(PruneSpan), C.c_int, C.c_int, _vp, _vp, _vp]),
    "dfw_shard_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "dfw_shard_free": (C.c_int, [_vp]),
    "dfw_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "dfw_ipc_import": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "dfw_ipc_close": (C.c_int, [_vp]),
    "dfw_shard_rows": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "dfw_gather_rows": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp]),
}

_lib = None


class DfwError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DfwError(
            f"{LIB_PATH} is missing. Build it with `python -m xsdeepfwfm_deprecated_b200.build` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.dfw_version() != DFW_ABI_VERSION:
        raise DfwError(f"ABI mismatch: library v{lib.dfw_version()} vs binding v{DFW_ABI_VERSION}")
    for which, st in enumerate((Model, FieldDesc, Csr)):
        if lib.dfw_struct_bytes(which) != C.sizeof(st):
            raise DfwError(f"ABI mismatch: sizeof({st.__name__}) is {C.sizeof(st)} in the binding, "
                           f"{lib.dfw_struct_bytes(which)} in the library")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dfw_last_error_string().decode(errors="replace")
        raise DfwError(f"{what or 'libdeepfwfm_sm100a'} failed (rc={rc}): {msg}")


_checked_devices = set()


def require_device(index: int):
    """Fail loudly unless `index` is an sm_100 GPU."""
    if index in _checked_devices:
        return
    check(load().dfw_check_device(int(index)), "dfw_check_device")
    _checked_devices.add(index)
