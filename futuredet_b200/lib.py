"""ctypes binding of libfuturedet_b200.so (the C ABI declared in include/futuredet_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  The oracle under oracle/ is test infrastructure only and is
never imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfuturedet_b200.so")

c_int_p = C.POINTER(C.c_int32)
c_float_p = C.POINTER(C.c_float)


class ConvDesc(C.Structure):
    """Mirror of `struct fd_conv_desc`."""
    _fields_ = [
        ("d_in", C.c_void_p), ("in_stride", C.c_int32), ("cin", C.c_int32),
        ("in_format", C.c_int32), ("in_ctot", C.c_int32),
        ("d_w", C.c_void_p), ("cout", C.c_int32), ("K", C.c_int32),
        ("d_w_packed", C.c_void_p),
        ("d_scale", C.c_void_p), ("d_shift", C.c_void_p),
        ("d_residual", C.c_void_p), ("res_stride", C.c_int32),
        ("res_format", C.c_int32), ("res_ctot", C.c_int32),
        ("relu", C.c_int32),
        ("d_out", C.c_void_p), ("out_stride", C.c_int32),
        ("out_format", C.c_int32), ("out_ctot", C.c_int32),
        ("d_n_out", C.c_void_p), ("n_out_cap", C.c_int32),
        ("mode", C.c_int32),
        ("d_nbr", C.c_void_p), ("nbr_stride", C.c_int32),
        ("d_tile_mask", C.c_void_p),
        ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("ph", C.c_int32), ("pw", C.c_int32),
        ("out_map", C.c_int32),
        ("d_out_coords4", C.c_void_p), ("bevD", C.c_int32), ("bevH", C.c_int32), ("bevW", C.c_int32),
        ("precision", C.c_int32),
        ("n_in_cap", C.c_int32),
        ("d_in_split", C.c_void_p), ("d_out_split", C.c_void_p),
        ("d_row_perm", C.c_void_p),
    ]


GATHER_TABLE, GATHER_CONV2D, GATHER_CONVT2D, GATHER_CONV2D_DGRAD = 0, 1, 2, 3
OUTMAP_IDENTITY, OUTMAP_BEV, OUTMAP_BEV_DMAJOR = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}

# name -> (restype, argtypes); must list every symbol of include/futuredet_b200.h
SIGNATURES = {
    "fd_version": (C.c_int, []),
    "fd_last_error": (C.c_char_p, []),
    "fd_launch_count": (C.c_int64, []),
    "fd_voxelize_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "fd_voxelize_vfe": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   c_float_p, c_float_p, c_int_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fd_vfe_mean": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fd_coord_index_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_int_p, C.c_void_p,
                                        C.c_int64, C.c_void_p]),
    "fd_scan_tmp_bytes": (C.c_size_t, [C.c_int64]),
    "fd_rulebook_out_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p,
                                          c_int_p, c_int_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_void_p]),
    "fd_rulebook_neighbors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                                         c_int_p, c_int_p, c_int_p, c_int_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "fd_rulebook_neighbors_bitmap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, c_int_p, c_int_p,
                                                c_int_p, c_int_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p]),
    "fd_rulebook_neighbors_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, c_int_p, c_int_p,
                                                 c_int_p, c_int_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_void_p]),
    "fd_rulebook_count_pairs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fd_rulebook_sort_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "fd_rulebook_sort_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fd_rulebook_to_pairs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "fd_conv_forward": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "fd_conv_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "fd_conv_pack_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fd_sparse_to_dense_ncdhw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fd_convert_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "fd_center_loss_workspace_bytes": (C.c_size_t, []),
    "fd_center_head_loss": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fd_fill_i32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    # ---- multi-sweep assembly
    "fd_sweeps_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "fd_assemble_sweeps": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    # ---- target assignment
    "fd_assign_center_targets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    # ---- predict
    "fd_center_predict_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "fd_center_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_float_p, C.c_float, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fd_boxes_iou_bev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    # ---- training
    "fd_rulebook_transpose": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p]),
    "fd_conv_wgrad": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p]),
    "fd_conv_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(ConvDesc)]),
    "fd_conv_wgrad_det": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fd_bn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "fd_bn_train_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fd_affine_act": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "fd_bn_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fd_col_sum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fd_add_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "fd_rows_to_bev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fd_bev_to_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fd_center_head_loss_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_float, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "futuredet_b200: %s is missing -- build it with `python -m futuredet_b200.build` "
            "(there is no CPU/PyTorch fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().fd_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, msg))


def launch_count():
    return int(load().fd_launch_count())


def i32x3(vals):
    return (C.c_int32 * 3)(*[int(v) for v in vals])


def f32(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])
