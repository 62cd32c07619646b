"""ctypes binding of ``libdeft_b200.so`` (the C ABI declared in ``include/deft_b200.h``).

The product path has no fallback: if the shared library is missing or lacks a symbol, importing
this module raises.  Build it with ``python -m deft_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEFT_B200_LIB") or os.path.join(_HERE, "lib", "libdeft_b200.so")   # (override: A/B of two builds)

ABI_VERSION = 5
T_NAMES = ["node_q", "node_kv", "node_q_len", "node_kv_len", "node_q_offset", "node_kv_offset",
           "block_q", "block_q_cnts", "block_q_offset", "block_bitmasks", "block_kv", "block_lens",
           "flat_items", "flat_groups", "flat_csr_off", "flat_csr_rows",
           "node_items", "node_groups", "node_csr_off", "node_csr_rows",
           "u_units", "u_csr_off", "u_csr_rows", "u_kv", "u_mask", "u_q", "u_job_off", "u_jobs", "u_blk"]
T_COUNT = len(T_NAMES)
STAGE1_AUTO, STAGE1_FMA, STAGE1_UMMA = 0, 1, 2
ITEM_BYTES = 24
GROUP_BYTES = 24
UNIT_BYTES = 80
JOB_BYTES = 96
N_SCALARS = 11


class Plan(C.Structure):
    """``deft_plan_t``"""
    _fields_ = [("items", C.c_void_p), ("groups", C.c_void_p), ("csr_off", C.c_void_p), ("csr_rows", C.c_void_p),
                ("n_items", C.c_int32), ("n_groups", C.c_int32), ("n_part_rows", C.c_int32), ("n_units", C.c_int32),
                ("units", C.c_void_p), ("u_csr_off", C.c_void_p), ("u_csr_rows", C.c_void_p), ("u_kv", C.c_void_p),
                ("u_blk", C.c_void_p), ("u_mask", C.c_void_p), ("u_q", C.c_void_p), ("u_job_off", C.c_void_p), ("u_jobs", C.c_void_p),
                ("n_unit_slots", C.c_int32), ("n_ctas", C.c_int32), ("hkv", C.c_int32), ("paired", C.c_int32),
                ("fresh", C.c_int32), ("pad", C.c_int32)]


class Append(C.Structure):
    """``deft_append_t``"""
    _fields_ = [("new_k", C.c_void_p), ("new_v", C.c_void_p), ("new_row_stride", C.c_int64), ("new_head_stride", C.c_int64),
                ("cache_loc", C.c_void_p)]


class DeftError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension with `python -m deft_b200.build` "
                          "(there is no CPU or PyTorch fallback for the tree-attention path)")
    lib = C.CDLL(LIB_PATH)
    i32, i64, vp, sz = C.c_int32, C.c_int64, C.c_void_p, C.c_size_t
    sig = {
        "deft_b200_abi_version": (C.c_int, []),
        "deft_b200_last_error": (C.c_char_p, []),
        "deft_b200_set_stages": (None, [i32]),
        "deft_b200_set_stage1_impl": (None, [i32]),
        "deft_b200_set_debug_buffer": (None, [vp]),
        "deft_b200_set_trace_buffer": (None, [vp]),
        "deft_b200_set_tma": (None, [i32]),
        "deft_b200_set_pdl": (None, [i32]),
        "deft_b200_set_gather4": (None, [i32]),
        "deft_b200_set_experiment": (None, [i32]),
        "deft_b200_flatten_workspace_bytes": (sz, [i32, i32, i32, i32, i64, i64, C.POINTER(Plan)]),
        "deft_b200_flatten_fwd": (C.c_int, [vp, i64, i64, vp, vp, i64, i64, i64, vp, i64, i64, i32, i32, i32, i32,
                                            i32, vp, i64, vp, vp, vp, i64, vp, vp, C.POINTER(Plan), vp, sz, vp]),
        "deft_b200_node_workspace_bytes": (sz, [i32, i32, i32, i32, i64, i64, i64, C.POINTER(Plan)]),
        "deft_b200_node_fwd": (C.c_int, [vp, i64, i64, vp, vp, i64, i64, i64, vp, i64, i64, i32, i32, i32, i32,
                                         vp, i32, vp, vp, vp, i64, vp, vp, i64, i64, C.POINTER(Plan), vp, sz, vp]),
        "deft_b200_kv_append": (C.c_int, [vp, vp, i64, i64, vp, vp, i64, i64, vp, i32, i32, i32, vp]),
        "deft_b200_build_tables": (vp, [i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
        "deft_b200_flatten_fwd_append": (C.c_int, [vp, i64, i64, vp, vp, i64, i64, i64, vp, i64, i64, i32, i32, i32, i32,
                                                   i32, vp, i64, vp, vp, vp, i64, vp, vp, C.POINTER(Plan), C.POINTER(Append), vp, sz, vp]),
        "deft_b200_node_fwd_append": (C.c_int, [vp, i64, i64, vp, vp, i64, i64, i64, vp, i64, i64, i32, i32, i32, i32,
                                                vp, i32, vp, vp, vp, i64, vp, vp, i64, i64, C.POINTER(Plan), C.POINTER(Append), vp, sz, vp]),
        "deft_b200_tree_new": (vp, []),
        "deft_b200_tree_free": (None, [vp]),
        "deft_b200_tree_set": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, i32]),
        "deft_b200_tree_append": (C.c_int, [vp, i32, vp, vp]),
        "deft_b200_tree_pages": (i64, [vp]),
        "deft_b200_build_tables_trees": (vp, [vp, i32, i64, i32, i32, i32, i32, i32, i32, vp, vp]),
        "deft_b200_layout_new": (vp, []),
        "deft_b200_layout_free": (None, [vp]),
        "deft_b200_layout_version": (i64, [vp]),
        "deft_b200_layout_set_native_only": (None, [vp, C.c_int]),
        "deft_b200_tables_data": (vp, [vp]),
        "deft_b200_tables_bytes": (sz, [vp]),
        "deft_b200_tables_directory": (C.c_int, [vp, vp]),
        "deft_b200_tables_scalars": (C.c_int, [vp, vp]),
        "deft_b200_tables_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    got = lib.deft_b200_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libdeft_b200.so ABI {got} != binding ABI {ABI_VERSION}: rebuild")
    return lib


lib = _load()
EXPORTS = ["deft_b200_abi_version", "deft_b200_last_error", "deft_b200_set_stages", "deft_b200_set_stage1_impl",
           "deft_b200_set_debug_buffer", "deft_b200_set_trace_buffer", "deft_b200_set_tma", "deft_b200_set_pdl", "deft_b200_set_gather4", "deft_b200_set_experiment",
           "deft_b200_flatten_workspace_bytes",
           "deft_b200_flatten_fwd", "deft_b200_node_workspace_bytes", "deft_b200_node_fwd", "deft_b200_flatten_fwd_append", "deft_b200_node_fwd_append", "deft_b200_kv_append",
           "deft_b200_build_tables", "deft_b200_tree_new", "deft_b200_tree_free", "deft_b200_tree_set", "deft_b200_tree_append",
           "deft_b200_tree_pages", "deft_b200_build_tables_trees", "deft_b200_layout_new", "deft_b200_layout_free", "deft_b200_layout_version", "deft_b200_layout_set_native_only",
           "deft_b200_tables_data", "deft_b200_tables_bytes",
           "deft_b200_tables_directory", "deft_b200_tables_scalars", "deft_b200_tables_free"]


def check(rc: int) -> None:
    if rc != 0:
        raise DeftError(f"libdeft_b200 error {rc}: {lib.deft_b200_last_error().decode()}")


def last_error() -> str:
    return lib.deft_b200_last_error().decode()
