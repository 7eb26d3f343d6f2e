"""ctypes binding of the C-ABI library `libscan_b200.so` (include/scan_b200.h).

The library is the product: there is NO fallback.  If it is missing or fails to load, importing an op
raises; if a call returns a negative code, `check()` raises RuntimeError (the reference turns AT_ERROR /
AT_ASSERTM into RuntimeError the same way, csrc/SigmoidFocalLoss.h:20-23).
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscan_b200.so")

SCAN_MAX_LEVELS = 8
SCAN_MAX_CLASSES = 16


class ScanLevels(ctypes.Structure):
    _fields_ = [("n_levels", c_int32), ("n_images", c_int32),
                ("h", c_int32 * SCAN_MAX_LEVELS), ("w", c_int32 * SCAN_MAX_LEVELS),
                ("stride", c_int32 * SCAN_MAX_LEVELS)]


class ScanSampleMeta(ctypes.Structure):
    _fields_ = [("n_nodes", c_int32), ("n_neg_nodes", c_int32), ("error", c_int32), ("reserved", c_int32),
                ("n_pos", c_int32 * SCAN_MAX_LEVELS), ("n_neg", c_int32 * SCAN_MAX_LEVELS),
                ("n_neg_sel", c_int32 * SCAN_MAX_LEVELS), ("neg_off", c_int32 * SCAN_MAX_LEVELS),
                ("pos_off", c_int32 * SCAN_MAX_LEVELS)]


_P = c_void_p
_LV = ctypes.POINTER(ScanLevels)

# name -> (restype, argtypes): every symbol include/scan_b200.h declares
SIGNATURES = {
    "scan_abi_version": (c_int32, []),
    "scan_strerror": (c_char_p, [c_int32]),
    "scan_last_cuda_error": (c_char_p, []),
    "scan_init": (c_int32, [c_int32]),
    "scan_upload_small": (c_int32, [_P, _P, c_int64, _P]),
    "scan_pack_rows": (c_int32, [_LV, _P, c_int32, _P, _P]),
    "scan_unpack_rows": (c_int32, [_LV, _P, c_int32, _P, c_int32, _P]),
    "scan_unpack_levels": (c_int32, [_LV, _P, c_int32, _P, _P]),
    "scan_gn_workspace_bytes": (c_int64, [_LV]),
    "scan_gn_relu_fwd": (c_int32, [_LV, _P, _P, _P, _P, c_float, _P, _P, _P, c_int64, _P]),
    "scan_gn_relu_apply": (c_int32, [_LV, _P, _P, _P, _P, _P, _P, _P]),
    "scan_gn_relu_bwd": (c_int32, [_LV, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_add_relu_fwd": (c_int32, [_LV, _P, _P, _P, _P, _P]),
    "scan_add_relu_bwd": (c_int32, [_LV, _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_fcos_assign": (c_int32, [_LV, _P, _P, _P, c_int32, _P, _P]),
    "scan_fcos_assign_reg": (c_int32, [_LV, _P, _P, _P, c_int32, _P, _P, _P]),
    "scan_fcos_loss_num_partials": (c_int32, []),
    "scan_fcos_loss_fwd": (c_int32, [_LV, _P, _P, _P, _P, _P, c_int32, c_float, c_float, _P, _P, _P, _P]),
    "scan_fcos_loss_bwd": (c_int32, [_LV, _P, _P, _P, _P, _P, c_int32, c_float, c_float, _P, _P, _P, _P, _P, _P]),
    "scan_sample_workspace_bytes": (c_int64, [c_int64]),
    "scan_sample_nodes": (c_int32, [_LV, c_int32, c_int32, _P, _P, _P, _P, _P, c_int32, _P, _P, c_int64, _P]),
    "scan_gather_rows": (c_int32, [_P, _P, c_int32, c_int32, _P, _P]),
    "scan_scatter_add_rows": (c_int32, [_P, _P, c_int32, c_int32, _P, _P]),
    "scan_condconv_num_partials": (c_int32, []),
    "scan_condconv_fwd": (c_int32, [_LV, _P, _P, _P, c_int32, c_int32, _P, _P, _P, _P, c_int32, _P]),
    "scan_condconv_bwd_workspace_bytes": (c_int64, [_LV, c_int32]),
    "scan_condconv_bwd": (c_int32, [_LV, _P, _P, c_int32, c_int32, _P, _P, _P, c_float, _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_condconv_bwd2": (c_int32, [_LV, _P, _P, c_int32, c_int32, _P, _P, _P, c_float, _P, _P, c_int32, _P, _P, _P, c_int64, _P]),
    "scan_manifest_rnn_saved_floats": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "scan_manifest_rnn_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "scan_manifest_rnn_fwd": (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, c_int32] + [_P] * 13),
    "scan_manifest_rnn_bwd": (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, c_int32] + [_P] * 16 + [c_int64, _P]),
    "scan_attn_workspace_bytes": (c_int64, [c_int32]),
    "scan_attn_fwd": (c_int32, [_P, _P, _P, c_int32, c_float, c_float, c_uint64, _P, _P, _P, c_int64, _P]),
    "scan_attn_bwd_workspace_bytes": (c_int64, [c_int32]),
    "scan_attn_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, c_int32, c_float, c_float, c_uint64, _P, _P, _P, _P, _P, c_int64,
                                _P]),
    "scan_graph_workspace_bytes": (c_int64, [c_int32]),
    "scan_qkv_fwd": (c_int32, [_P, _P, _P, c_int32, _P, _P]),
    "scan_qkv_bwd": (c_int32, [_P, _P, _P, c_int32, c_int32, _P, _P, _P, _P, c_int64, _P]),
    "scan_attn_out_ln_fwd": (c_int32, [_P, _P, _P, _P, _P, _P, c_int32, c_float, c_float, c_uint64, _P, _P, _P, _P]),
    "scan_attn_out_ln_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, c_int32, c_float, c_uint64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_node_cls_fwd": (c_int32, [_P, _P, _P, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P, _P, _P, c_int64, _P]),
    "scan_node_cls_bwd": (c_int32, [_P, _P, _P, _P, _P, c_int32, c_int32, c_int32, c_float, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_class_mean_bwd": (c_int32, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P]),
    "scan_gemm_nt_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "scan_gemm_nt": (c_int32, [_P, c_int64, _P, c_int64, c_int32, c_int32, c_int32, _P, c_int32, c_int32, _P, c_int64, _P, c_int64, _P]),
    "scan_transpose": (c_int32, [_P, c_int32, c_int32, c_int64, _P, c_int64, _P]),
    "scan_linear_wgrad_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "scan_linear_wgrad": (c_int32, [_P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, c_int64, _P]),
    "scan_rows_softmax": (c_int32, [_P, c_int32, c_int32, c_int64, _P]),
    "scan_rows_l2normalize": (c_int32, [_P, c_int32, c_float, _P, _P]),
    "scan_gcn_act_fwd": (c_int32, [_P, _P, c_int32, c_int32, _P, _P, _P]),
    "scan_gcn_act_bwd": (c_int32, [_P, _P, c_int32, c_int32, _P, _P]),
    "scan_rows_linear_fwd": (c_int32, [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "scan_rows_linear_bwd": (c_int32, [_P, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "scan_rows_gn_relu_fwd": (c_int32, [_P, _P, _P, c_int32, c_int32, c_int32, c_float, _P, _P, _P]),
    "scan_rows_gn_relu_bwd": (c_int32, [_P, _P, _P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "scan_conv3x3_packed_floats": (c_int64, [c_int32, c_int32]),
    "scan_conv3x3_pack_weights": (c_int32, [_P, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, _P, _P, _P]),
    "scan_tf32_residual": (c_int32, [_P, c_int64, _P, _P]),
    "scan_conv3x3_gn_workspace_bytes": (c_int64, [_LV]),
    "scan_conv3x3_rows_gn": (c_int32, [_LV, _P, _P, c_int32, _P, _P, _P, c_float, _P, _P, c_int32, _P, c_int64, _P]),
    "scan_conv3x3_rows": (c_int32, [_LV, _P, _P, c_int32, _P, _P, c_int32, _P, _P, c_int32, _P, c_int32, c_int32, _P]),
    "scan_conv3x3_rows2": (c_int32, [_LV, _P, _P, c_int32, _P, _P, c_int32, _P, _P, c_int32, _P, _P, _P, c_int32, _P, c_int32, c_int32, _P]),
    "scan_conv1x1_rows": (c_int32, [_LV, _P, _P, c_int32, _P, _P, c_int32, _P, c_int32, _P, c_int32, c_int32, _P]),
    "scan_conv3x3_wgrad_workspace_bytes": (c_int64, [_LV, c_int32, c_int32, c_int32]),
    "scan_conv3x3_wgrad": (c_int32, [_LV, _P, _P, c_int32, _P, _P, c_int32, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P]),
    "scan_thin_wgrad_workspace_bytes": (c_int64, [_LV, c_int32]),
    "scan_thin_wgrad": (c_int32, [_LV, _P, _P, _P, c_int32, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P]),
    "scan_thin_pack": (c_int32, [_LV, _P, c_int32, c_int32, c_int32, _P, c_int32, _P]),
    "scan_thin_unpack": (c_int32, [_LV, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P]),
    "scan_scale": (c_int32, [_P, c_int64, c_float, _P, _P]),
    "scan_thin_gather": (c_int32, [_LV, _P, c_int32, c_int32, _P, _P]),
    "scan_colsum_workspace_bytes": (c_int64, [c_int64, c_int32]),
    "scan_colsum": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, c_int64, _P]),
    "scan_cka_bce_workspace_bytes": (c_int64, []),
    "scan_cka_bce_fwd": (c_int32, [_P, c_int32, _P, c_int32, c_int64, c_int32, c_float, _P, _P, _P, c_int64, _P]),
    "scan_cka_bce_bwd": (c_int32, [_P, c_int32, _P, c_int32, c_int64, c_int32, c_float, _P, _P, _P, _P, c_int32, _P]),
    "scan_postprocess_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "scan_postprocess": (c_int32, [_LV, _P, _P, _P, c_int32, _P, c_float, c_int32, c_float, c_int32, c_float, _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_transfer_nodes_num_partials": (c_int32, []),
    "scan_transfer_nodes_fwd": (c_int32, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "scan_transfer_nodes_bwd": (c_int32, [_P, _P, c_int32, _P, _P]),
    "scan_transfer_proto": (c_int32, [_P, _P, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "scan_class_sums": (c_int32, [_P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "scan_proto_update": (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P, _P]),
    "scan_dbscan_workspace_bytes": (c_int64, [c_int64]),
    "scan_dbscan_level": (c_int32, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_float, c_double, c_int32, c_int32,
                                    _P, _P, _P, _P, _P, c_int64, _P]),
    "scan_dbscan_points": (c_int32, [_P, c_int32, c_int32, c_double, c_int32, _P, _P, _P, c_int64, _P]),
    "scan_sigmoid_focal_fwd": (c_int32, [_P, _P, c_int64, c_int32, c_float, c_float, _P, _P]),
    "scan_sigmoid_focal_bwd": (c_int32, [_P, _P, _P, c_int64, c_int32, c_float, c_float, _P, _P]),
    "scan_ensemble_levels": (c_int32, [_LV, _P, _P, c_int32, c_int32, _P, _P]),
}

_lib = None
# kernels each entry point launches (memsets excluded): bench.py's gpu_launches is counted from this table
LAUNCHES = {"scan_manifest_rnn_fwd": 6, "scan_manifest_rnn_bwd": 7, "scan_gn_relu_fwd": 3, "scan_gn_relu_apply": 1, "scan_conv3x3_rows_gn": 2, "scan_gn_relu_bwd": 4, "scan_add_relu_fwd": 1, "scan_add_relu_bwd": 2, "scan_upload_small": 1, "scan_pack_rows": 1, "scan_unpack_rows": 1, "scan_unpack_levels": 1, "scan_fcos_assign": 1, "scan_fcos_assign_reg": 1, "scan_fcos_loss_fwd": 2, "scan_fcos_loss_bwd": 1, "scan_sample_nodes": 4, "scan_gather_rows": 1,
            "scan_scatter_add_rows": 1, "scan_condconv_fwd": 1, "scan_condconv_bwd": 3, "scan_condconv_bwd2": 3, "scan_attn_fwd": 5, "scan_attn_bwd": 4,
            "scan_class_sums": 1, "scan_proto_update": 1, "scan_dbscan_level": 20, "scan_dbscan_points": 15,
            "scan_qkv_fwd": 1, "scan_qkv_bwd": 9, "scan_attn_out_ln_fwd": 1, "scan_attn_out_ln_bwd": 9, "scan_node_cls_fwd": 3,
            "scan_node_cls_bwd": 9, "scan_class_mean_bwd": 1,
            "scan_gemm_nt": 2, "scan_transpose": 1, "scan_linear_wgrad": 5, "scan_rows_softmax": 1, "scan_rows_l2normalize": 1,
            "scan_gcn_act_fwd": 1, "scan_gcn_act_bwd": 1,
            "scan_rows_linear_fwd": 1, "scan_rows_linear_bwd": 2, "scan_rows_gn_relu_fwd": 1, "scan_rows_gn_relu_bwd": 2,
            "scan_postprocess": 4, "scan_conv3x3_pack_weights": 1, "scan_tf32_residual": 1, "scan_conv3x3_rows": 1, "scan_conv3x3_rows2": 1, "scan_conv1x1_rows": 1, "scan_thin_pack": 1, "scan_thin_unpack": 1, "scan_scale": 1, "scan_thin_gather": 1,
            "scan_colsum": 2, "scan_thin_wgrad": 3, "scan_cka_bce_fwd": 2, "scan_cka_bce_bwd": 1, "scan_conv3x3_wgrad": 2,
            "scan_transfer_nodes_fwd": 2, "scan_transfer_nodes_bwd": 1, "scan_transfer_proto": 1,
            "scan_sigmoid_focal_fwd": 1, "scan_sigmoid_focal_bwd": 1, "scan_ensemble_levels": 1}
CALLS = {"n": 0, "launches": 0}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "scan_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or scan_b200/csrc/build.sh); there is no fallback path" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.scan_abi_version() != 1:
            raise RuntimeError("scan_b200: ABI version mismatch")
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        L = lib()
        msg = L.scan_strerror(code).decode()
        if code == -2:
            msg += ": " + L.scan_last_cuda_error().decode()
        raise RuntimeError("scan_b200.%s failed: %s" % (what, msg))


def call(name, *args):
    CALLS["n"] += 1
    CALLS["launches"] += LAUNCHES.get(name, 0)
    check(getattr(lib(), name)(*args), name)
