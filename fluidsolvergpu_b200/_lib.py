"""ctypes binding of libfsg.so (include/fsg.h).  There is no fallback: if the library is missing the
import of the compute classes fails loudly, and on a box without a CUDA device fsg_create returns
FSG_E_NO_DEVICE."""
from __future__ import annotations

import ctypes as C
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "libfsg.so"

FSG_OK, FSG_E_INVALID, FSG_E_NO_DEVICE, FSG_E_CUDA, FSG_E_NOMEM, FSG_E_STATE, FSG_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
FSG_MODEL_BASE, FSG_MODEL_UNIDYN = 0, 1
FSG_AOS_STRIDE = 340


class FsgConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("grid", C.c_int32), ("origin", C.c_float), ("cellsize", C.c_double), ("h", C.c_double),
        ("dt", C.c_double), ("gravity", C.c_double), ("sound", C.c_double), ("alpha_fluid", C.c_double),
        ("alpha_boundary", C.c_double), ("neighbour_cap", C.c_int32), ("bin_cap", C.c_int32), ("capacity", C.c_int64),
        ("device", C.c_int32), ("pair_fp64", C.c_int32), ("collect_stats", C.c_int32), ("rank", C.c_int32),
        ("world", C.c_int32), ("slab_x0", C.c_int32), ("slab_x1", C.c_int32),
        ("pair_mode", C.c_int32), ("unidyn_open_box", C.c_int32), ("unidyn_adapt", C.c_int32),
        ("unidyn_merge_distance", C.c_double), ("unidyn_split_mass_min", C.c_double),
    ]


class FsgSoa(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("pos", C.c_void_p), ("vel", C.c_void_p), ("acc", C.c_void_p), ("dens", C.c_void_p),
        ("press", C.c_void_p), ("delpress", C.c_void_p), ("newdens", C.c_void_p), ("newdelpress", C.c_void_p),
        ("index", C.c_void_p), ("cell", C.c_void_p), ("boundary", C.c_void_p), ("solid", C.c_void_p), ("fluid", C.c_void_p),
        ("stress_tensor", C.c_void_p), ("stress_rate", C.c_void_p), ("mass", C.c_void_p),
    ]


class FsgStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n", "n_live", "occupied_bins", "pairs_tested", "pairs_in_range", "dropped",
                                           "steps", "kernel_launches")]


# every symbol include/fsg.h declares: name -> (restype, argtypes)
P = C.c_void_p
SIGNATURES = {
    "fsg_version": (C.c_int, []),
    "fsg_device_count": (C.c_int, []),
    "fsg_config_default": (C.c_int, [C.POINTER(FsgConfig), C.c_int]),
    "fsg_create": (C.c_int, [C.POINTER(FsgConfig), C.POINTER(P)]),
    "fsg_destroy": (C.c_int, [P]),
    "fsg_last_error": (C.c_char_p, [P]),
    "fsg_set_stream": (C.c_int, [P, P]),
    "fsg_get_stream": (P, [P]),
    "fsg_upload_aos": (C.c_int, [P, P, C.c_int64]),
    "fsg_upload_soa": (C.c_int, [P, C.POINTER(FsgSoa)]),
    "fsg_download_aos": (C.c_int, [P, P, C.c_int64]),
    "fsg_download_soa": (C.c_int, [P, C.POINTER(FsgSoa)]),
    "fsg_step": (C.c_int, [P, C.c_int]),
    "fsg_sync": (C.c_int, [P]),
    "fsg_export_viz": (C.c_int, [P, P, P, P]),
    "fsg_get_tables": (C.c_int, [P, P, P, P]),
    "fsg_get_split": (C.c_int, [P, P]),
    "fsg_get_stats": (C.c_int, [P, C.POINTER(FsgStats)]),
    "fsg_set_collect_stats": (C.c_int, [P, C.c_int]),
    "fsg_set_profiling": (C.c_int, [P, C.c_int]),
    "fsg_get_phase_ms": (C.c_int, [P, C.POINTER(C.c_double * 4), C.POINTER(C.c_int64)]),
    "fsg_scene_plume": (C.c_int, [P, C.c_double, C.c_double, C.c_uint64, C.POINTER(C.c_int64)]),
    "fsg_scene_plume_host": (C.c_int, [C.POINTER(FsgConfig), C.c_double, C.c_double, C.c_uint64, P, P, C.c_int64,
                                       C.POINTER(C.c_int64)]),
    "fsg_scene_plume_hist": (C.c_int, [C.POINTER(FsgConfig), C.c_double, P]),
    "fsg_device_ptr": (C.c_int, [P, C.c_int, C.POINTER(P)]),
    "fsg_write_point_mesh": (C.c_int, [C.c_char_p, C.c_int, C.c_int, P, C.c_int, P, P, P]),
    "fsg_write_frame": (C.c_int, [P, C.c_char_p, C.c_int]),
    "fsg_write_frame_async": (C.c_int, [P, C.c_char_p, C.c_int]),
    "fsg_frame_wait": (C.c_int, [P, C.POINTER(C.c_int64)]),
    "fsg_unidyn_adapt_counts": (C.c_int, [P, C.POINTER(C.c_int64 * 3), C.POINTER(C.c_int64 * 3)]),
    "fsg_slab_pack": (C.c_int, [P, P, P, C.c_int64, C.c_int64]),
    "fsg_slab_unpack": (C.c_int, [P, P, P, C.c_int64, C.c_int64]),
    "fsg_slab_check": (C.c_int, [P, C.POINTER(C.c_int64 * 9)]),
    "fsg_slab_keep_foreign": (C.c_int, [P, C.c_int]),
    "fsg_slab_message_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "fsg_slab_message_bytes_model": (C.c_int64, [C.c_int, C.c_int64, C.c_int64]),
    "fsg_slab_alloc_messages": (C.c_int, [P, C.c_int64, C.c_int64]),
    "fsg_slab_alloc_messages2": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int]),
    "fsg_slab_mode": (C.c_int, [P]),
    "fsg_slab_set_split_step": (C.c_int, [P, C.c_int]),
    "fsg_slab_step_finish": (C.c_int, [P]),
    "fsg_slab_get_ghost_ms": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fsg_slab_inbox_handle": (C.c_int, [P, C.c_int, C.c_int, P]),
    "fsg_slab_open_peer": (C.c_int, [P, C.c_int, C.c_int, P]),
    "fsg_slab_pack_send": (C.c_int, [P]),
    "fsg_slab_unpack_recv": (C.c_int, [P]),
    "fsg_slab_close_peers": (C.c_int, [P]),
    "fsg_slab_set_peer": (C.c_int, [P, C.c_int, C.c_int, P]),
    "fsg_slab_inbox_ptr": (P, [P, C.c_int, C.c_int]),
    "fsg_slab_set_overlap": (C.c_int, [P, C.c_int]),
    "fsg_stage_sort": (C.c_int, [P, P, P, C.c_int64]),
    "fsg_stage_findneighbours": (C.c_int, [P, P, P, P, C.c_int64]),
    "fsg_stage_mykernel": (C.c_int, [P, P, P, P, P, C.c_int64]),
    "fsg_stage_mykernel2": (C.c_int, [P, P, P, P, P, C.c_int64, P, P, P]),
    "fsg_stage_unidyn_count_after_merge": (C.c_int, [P, P, C.c_int64, P]),
    "fsg_stage_unidyn_findneighbours": (C.c_int, [P, P, P, P, P, C.c_int64, C.c_int32]),
    "fsg_stage_unidyn_mykernel": (C.c_int, [P, P, P, P, P, P, P, C.c_int64]),
    "fsg_stage_unidyn_mykernel3": (C.c_int, [P, P, P, P, P, C.c_int64]),
    "fsg_stage_unidyn_mykernel2": (C.c_int, [P, P, P, P, P, P, P, P, C.c_int64, C.c_int32, C.c_int32, P, P, P]),
    "fsg_stage_unidyn_cell_calc": (C.c_int, [P, P, P, C.c_int64]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libfsg.so from the package directory (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C fluidsolvergpu_b200/csrc).  fluidsolvergpu_b200 has no CPU fallback.")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class FsgError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str = ""):
        self.code = code
        super().__init__(f"{where} failed with code {code}" + (f": {msg}" if msg else ""))
