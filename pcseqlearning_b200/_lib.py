"""ctypes binding of libpcseq_b200.so (the C ABI declared in include/pcseq_b200.h).

There is NO CPU fallback: if the library is missing it is built with nvcc; if that fails, or if a call
is made without a CUDA device, an exception is raised.
"""
import ctypes
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
c_double = ctypes.c_double

# name -> (restype, argtypes); mirrors include/pcseq_b200.h one to one
_SIGNATURES = {
    "pcs_version": (c_int, []),
    "pcs_last_error": (ctypes.c_char_p, []),
    "pcs_launch_count": (c_int64, []),
    "pcs_reset_launch_count": (None, []),
    "pcs_bounds_init": (c_int, [c_void_p, c_void_p, c_int]),
    "pcs_bounds_update": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "pcs_grid_params": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "pcs_voxel_keys": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "pcs_hash_build": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "pcs_hash_build_sorted_ws_bytes": (c_int64, [c_int64, c_int64]),
    "pcs_hash_build_sorted": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_int64]),
    "pcs_radius_search": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_int64]),
    "pcs_self_search_uf": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_void_p, c_void_p, c_int64]),
    "pcs_exclusive_scan": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64]),
    "pcs_exclusive_scan_tmp_bytes": (c_int64, [c_int64]),
    "pcs_lists_to_edges": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                   c_void_p]),
    "pcs_uf_init": (c_int, [c_void_p, c_void_p, c_int64]),
    "pcs_uf_union_edges": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "pcs_uf_labels": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64]),
    "pcs_uf_labels_tmp_bytes": (c_int64, [c_int64, c_int]),
    "pcs_point_segments": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "pcs_voxelize_params": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pcs_voxelize_insert": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                    c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcs_sort_pairs_tmp_bytes": (c_int64, [c_int64]),
    "pcs_sort_pairs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64]),
    "pcs_voxelize_finish": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcs_ground_ransac": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcs_l1_heightfield": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                   c_float, c_int, c_float, c_int, c_void_p, c_void_p]),
    "pcs_plane_prune": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                c_void_p]),
    "pcs_smooth_velo": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                c_float, c_int, c_float, c_void_p]),
    "pcs_group_minmax": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                 c_void_p]),
    "pcs_gather_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "pcs_group_median": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p]),
    "pcs_trk_cell_keys": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_double, c_void_p]),
    "pcs_trk_grid_fill": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pcs_trk_table_clear": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pcs_trk_bounds_reset": (c_int, [c_void_p, c_void_p, c_int]),
    "pcs_trk_group_bounds": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pcs_trk_sampler_init": (c_int, [c_void_p, c_void_p]),
    "pcs_trk_sample": (c_int, [c_void_p, c_void_p]),
    "pcs_trk_icp": (c_int, [c_void_p, c_void_p]),
    "pcs_trk_icp_timing": (None, [c_int]),
    "pcs_trk_icp_elapsed": (c_int, [c_void_p]),
    "pcs_trk_dir_init": (c_int, [c_void_p, c_void_p]),
    "pcs_trk_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "pcs_trk_finish": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pcs_trk_run": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pcs_trk_group_grid": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_double, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcs_trk_group_nn": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "pcs_compat_hash_insert": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                       c_void_p]),
    "pcs_compat_radius_degree": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pcs_compat_radius_fill": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                       c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p]),
    "pcs_nn_correspondence": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "pcs_points_in_radius": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p, c_int64, c_void_p, c_void_p, c_float, c_void_p]),
    "pcs_radius_graph": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                 c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p, c_void_p, c_int64]),
    "pcs_connected_components": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int,
                                         c_void_p, c_void_p, c_void_p, c_int64]),
    "pcs_voxelize": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                             c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcs_register_icp": (c_int, [c_void_p, c_void_p]),
    "pcs_box_prep": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pcs_points_in_boxes": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def library_path():
    return _build.SO


def lib():
    """Load (building first if necessary) libpcseq_b200.so.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


PCS_ERR_BAD_ARG, PCS_ERR_TABLE_FULL, PCS_ERR_KEY_RANGE = -2, -3, -4  # include/pcseq_b200.h


class PcsError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().pcs_last_error()
        raise PcsError(f"{what}: {msg.decode() if msg else rc}")
