"""ctypes binding of libpathnet_b200.so (the C ABI in include/pathnet_b200.h).

There is no CPU fallback: if the library is missing it is built (nvcc), and if
that fails importing the op raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
_LIB = None


class PnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pathnet_b200 error {code}: {msg}")
        self.code = code


class PaggCfg(C.Structure):
    _fields_ = [("N", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("C", C.c_int32),
                ("S", C.c_int32), ("nw", C.c_int32), ("L", C.c_int32), ("variant", C.c_int32),
                ("math_mode", C.c_int32), ("training", C.c_int32), ("dropout_p", C.c_float),
                ("dropout_seed", C.c_uint64), ("no_backward", C.c_int32), ("row_chunk", C.c_int32),
                ("dropout_seed_dev", C.c_void_p), ("x16", C.c_void_p), ("cell", C.c_int32), ("pool", C.c_int32)]


PARAM_NAMES = ("fc0_w", "fc0_b", "nets_w", "nets_b", "w_ih", "w_hh", "b_ih", "b_hh", "attw_w",
               "attw_b", "fc2_w", "fc2_b")


class AdamCfg(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float)]


class PaggParams(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in PARAM_NAMES]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = PKG / "libpathnet_b200.so"
    if not so.exists():
        from . import build as _b
        _b.build()
    L = C.CDLL(str(so))
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.pn_version.restype = C.c_int
    L.pn_last_error.restype = C.c_char_p
    L.pn_launch_count.restype = u64
    L.pn_dropout_hash.restype = C.c_uint32
    L.pn_dropout_hash.argtypes = [u64, C.c_uint32, u64]
    L.pn_graph_from_edges.argtypes = [i32, i64, vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.pn_graph_from_edges_uniform.argtypes = [i32, i64, vp, vp, C.c_int, C.POINTER(vp)]
    L.pn_graph_free.argtypes = [vp]
    L.pn_graph_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    L.pn_graph_cells_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pn_graph_prepare_labels.argtypes = [vp, i32, C.c_size_t, vp]
    L.pn_merw_sample.argtypes = [vp, i32, i32, i64, i64, i32, i32, i32, u64, vp, vp, vp]
    L.pn_merw_plan_host.argtypes = [i32, i32, i32, i64, i64, i32, i64, u64, C.POINTER(i32),
                                    C.POINTER(i64), vp]
    L.pn_merw_probabilities.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, C.c_double, i32, C.POINTER(i32), vp]
    L.pn_paths_format_txt.restype = i64
    L.pn_paths_format_txt.argtypes = [vp, vp, i64, i32, vp, i64]
    L.pn_paths_parse_txt.restype = i64
    L.pn_paths_parse_txt.argtypes = [vp, i64, i32, vp, vp, i64]
    L.pn_edges_parse_txt.restype = i64
    L.pn_edges_parse_txt.argtypes = [vp, i64, i32, i64, C.POINTER(i32), C.POINTER(i64), vp, vp, vp]
    L.pn_debug_gemm_tc.argtypes = [vp, vp, vp, vp, i64, i64, i64, C.c_int, vp]
    L.pn_debug_gemm_tc2.argtypes = [vp, vp, vp, i64, i64, i64, C.c_int, C.c_int, vp]
    L.pn_profile_enable.argtypes = [C.c_int]
    L.pn_profile_read.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(i64)]
    if hasattr(L, "pn_pagg_forward"):
        L.pn_pagg_canonicalize.argtypes = [i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, vp, vp, vp, vp]
        L.pn_pagg_workspace_bytes.restype = C.c_size_t
        L.pn_pagg_workspace_bytes.argtypes = [C.POINTER(PaggCfg)]
        L.pn_pagg_forward.argtypes = [C.POINTER(PaggCfg), C.POINTER(PaggParams), vp, vp, vp, vp, vp,
                                      vp, vp, vp]
        L.pn_pagg_backward.argtypes = [C.POINTER(PaggCfg), C.POINTER(PaggParams), vp, vp, vp, vp,
                                       vp, vp, vp, C.POINTER(PaggParams), vp]
        L.pn_pagg_x16_bytes.restype = C.c_size_t
        L.pn_pagg_x16_bytes.argtypes = [i32, i32]
        L.pn_pagg_pack_x16.argtypes = [i32, i32, vp, vp, vp]
        L.pn_debug_gemm_tma.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, C.c_int, C.c_float, vp, vp]
        L.pn_step_advance.argtypes = [vp, vp]
        L.pn_ce_loss_grad.argtypes = [i64, i32, vp, vp, i32, C.c_float, vp, vp, vp]
        L.pn_adam_step.argtypes = [i64, vp, vp, vp, vp, C.POINTER(AdamCfg), vp, vp]
    _LIB = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise PnError(rc, lib().pn_last_error().decode(errors="replace"))
    return rc


# symbols include/pathnet_b200.h declares (tests assert the .so exports each)
EXPORTS = ("pn_version", "pn_last_error", "pn_graph_from_edges", "pn_graph_from_edges_uniform", "pn_graph_free", "pn_graph_info",
           "pn_graph_cells_host", "pn_graph_prepare_labels", "pn_merw_sample", "pn_merw_probabilities",
           "pn_paths_format_txt", "pn_paths_parse_txt", "pn_edges_parse_txt", "pn_pagg_canonicalize",
           "pn_pagg_workspace_bytes", "pn_pagg_forward", "pn_pagg_backward", "pn_pagg_x16_bytes", "pn_pagg_pack_x16",
           "pn_step_advance", "pn_ce_loss_grad",
           "pn_adam_step", "pn_dropout_hash",
           "pn_launch_count", "pn_profile_enable", "pn_profile_read")


def profile_enable(on: bool) -> None:
    lib().pn_profile_enable(1 if on else 0)


def profile_read(name: str):
    ms, n = C.c_double(), C.c_int64()
    lib().pn_profile_read(name.encode(), C.byref(ms), C.byref(n))
    return ms.value, n.value
