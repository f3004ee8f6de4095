"""ctypes binding of libcsgpu.so (include/csgpu.h). No torch types cross this boundary.

The library is built in-tree (codesearch_b200/libcsgpu.so) by `make -C codesearch_b200/csrc`
or `__graft_entry__.build()`. There is no fallback: if the shared object is missing, import of
the product path fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsgpu.so")

OK, ERR_DIM, ERR_NOT_BUILT, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_ARG = range(7)
DTYPE_F32, DTYPE_BF16 = 0, 1
MAX_K = 1024
EXCHANGE_HANDLE_BYTES = 64
KEY_EMPTY = 0xFFFFFFFFFFFFFFFF


class CsgpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [
        ("live_rows", ctypes.c_uint64),
        ("pending_rows", ctypes.c_uint64),
        ("tombstones", ctypes.c_uint64),
        ("zero_norm_rows", ctypes.c_uint64),
        ("nonfinite_rows", ctypes.c_uint64),
        ("bytes_on_device", ctypes.c_uint64),
        ("dim", ctypes.c_uint32),
        ("dtype", ctypes.c_uint32),
        ("n_devices", ctypes.c_uint32),
        ("built", ctypes.c_uint32),
        ("last_search_us", ctypes.c_float),
        ("abi_version", ctypes.c_uint32),
        ("rows_per_device", ctypes.c_uint64 * 8),
        ("coalesced_passes", ctypes.c_uint64),
        ("coalesced_queries", ctypes.c_uint64),
        ("prefilter_rescored", ctypes.c_uint64),
        ("shadow_bytes", ctypes.c_uint64),
        ("byte_shadow_bytes", ctypes.c_uint64),
        ("byte_searches", ctypes.c_uint64),
        ("byte_fallbacks", ctypes.c_uint64),
        ("byte_candidates", ctypes.c_uint64),
        ("byte_rescored", ctypes.c_uint64),
        ("batch_route", ctypes.c_uint32),
        ("filter_max_err", ctypes.c_float),
    ]


class Predicate(ctypes.Structure):
    """csgpu_predicate_t (include/csgpu.h)."""
    _fields_ = [
        ("lang_mask", ctypes.c_uint32),
        ("file_lo", ctypes.c_uint32),
        ("file_hi", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
        ("file_bitmap", ctypes.c_void_p),
        ("n_file_bits", ctypes.c_uint64),
    ]


_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p

# name -> (restype, argtypes): every symbol include/csgpu.h declares
SIGNATURES = {
    "csgpu_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_uint32, ctypes.c_uint32, _i32p, ctypes.c_uint32]),
    "csgpu_destroy": (None, [_vp]),
    "csgpu_append": (ctypes.c_int, [_vp, _f32p, _u32p, ctypes.c_uint64]),
    "csgpu_remove": (ctypes.c_int, [_vp, _u32p, ctypes.c_uint64, _u64p]),
    "csgpu_reserve": (ctypes.c_int, [_vp, ctypes.c_uint64]),
    "csgpu_build": (ctypes.c_int, [_vp]),
    "csgpu_clear": (ctypes.c_int, [_vp]),
    "csgpu_save": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "csgpu_load": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "csgpu_search": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, _u32p, _f32p, _u32p]),
    "csgpu_set_coalescing": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32]),
    "csgpu_set_tensor_prefilter": (ctypes.c_int, [_vp, ctypes.c_uint32]),
    "csgpu_set_byte_prefilter": (ctypes.c_int, [_vp, ctypes.c_uint32]),
    "csgpu_search_batch": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p, _f32p, _u32p]),
    "csgpu_search_variants": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p, _f32p, _u32p]),
    "csgpu_search_filtered": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, _u64p, ctypes.c_uint64, _u32p, _f32p, _u32p]),
    "csgpu_append_tagged": (ctypes.c_int, [_vp, _f32p, _u32p, _u32p, ctypes.c_uint64]),
    "csgpu_search_tagged": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(Predicate), _u32p, _f32p, _u32p]),
    "csgpu_search_variants_tagged": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(Predicate), _u32p, _f32p, _u32p]),
    "csgpu_get_tags": (ctypes.c_int, [_vp, _u32p, ctypes.c_uint64, _u32p]),
    "csgpu_search_tagged_keys_device": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.POINTER(Predicate), ctypes.c_uint32, _vp, _vp]),
    "csgpu_append_synthetic_tagged": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]),
    "csgpu_search_keys_device": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, _vp, _vp]),
    "csgpu_merge_keys_device": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp]),
    "csgpu_merge_keys_batch_device": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp]),
    "csgpu_encode_keys": (None, [_u32p, _f32p, ctypes.c_uint32, ctypes.c_uint32, _u64p]),
    "csgpu_exchange_create": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    "csgpu_exchange_connect": (ctypes.c_int, [_vp, _vp]),
    "csgpu_exchange_connect_local": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "csgpu_search_keys_exchange_device": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, _vp, _vp]),
    "csgpu_search_exchange": (ctypes.c_int, [_vp, _f32p, ctypes.c_uint32, ctypes.c_uint32, _u32p, _f32p, _u32p]),
    "csgpu_exchange_status": (ctypes.c_int, [_vp, _u32p]),
    "csgpu_exchange_set_timeout_ms": (ctypes.c_int, [_vp, ctypes.c_uint32]),
    "csgpu_exchange_wait_stats": (ctypes.c_int, [_vp, _u64p, ctypes.c_uint32, _u32p]),
    "csgpu_exchange_destroy": (None, [_vp]),
    "csgpu_decode_keys": (None, [_u64p, ctypes.c_uint32, _u32p, _f32p, _u32p]),
    "csgpu_append_synthetic": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]),
    "csgpu_synth_rows_host": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, _f32p]),
    "csgpu_stats": (ctypes.c_int, [_vp, ctypes.POINTER(Stats)]),
    "csgpu_kernel_launches": (ctypes.c_uint64, []),
    "csgpu_last_error": (ctypes.c_char_p, []),
    "csgpu_abi_version": (ctypes.c_uint32, []),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libcsgpu.so and bind every declared symbol. Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C codesearch_b200/csrc` "
                "(or __graft_entry__.build()). There is no CPU/PyTorch fallback for the search path."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().csgpu_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise CsgpuError(rc, last_error())
