"""ctypes binding of libfvgn_b200.so (include/fvgn_b200.h).  There is NO fallback: if the library is missing or the
device is not sm_100, every op raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FVGN_B200_LIB: an experimental build of the same ABI (build.py --variant ...), for A/B measurements of kernel variants
LIB_PATH = os.environ.get("FVGN_B200_LIB") or os.path.join(_HERE, "libfvgn_b200.so")

MATH_FP32_SIMT, MATH_BF16_TC, MATH_BF16X3_TC = 0, 1, 2
MATH_MODES = {"fp32": MATH_FP32_SIMT, "bf16": MATH_BF16_TC, "bf16x3": MATH_BF16X3_TC}


class MlpT(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p),
                ("b3", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p), ("k_in", C.c_int), ("n_out", C.c_int), ("packed", C.c_void_p),
                ("packed_x3", C.c_void_p), ("z_save", C.c_void_p), ("packed_x3_bwd", C.c_void_p),
                ("da_in", C.c_void_p), ("n_products", C.c_int)]


class TrainTablesT(C.Structure):
    _fields_ = [("n_cells", C.c_int), ("n_faces", C.c_int), ("n_nodes", C.c_int), ("neighbour_cell", C.c_void_p), ("face_node", C.c_void_p),
                ("cells_node_T", C.c_void_p), ("node_ptr", C.c_void_p), ("node_items", C.c_void_p), ("cell_from_face_ptr", C.c_void_p),
                ("cell_from_face_items", C.c_void_p), ("node_from_cell_ptr", C.c_void_p), ("node_from_cell_items", C.c_void_p)]


EXCHANGE_MAX_PEERS = 8


class ExchangePlanT(C.Structure):  # fvgn_exchange_plan_t
    _M = EXCHANGE_MAX_PEERS
    _fields_ = [("n_peers", C.c_int), ("row_bytes", C.c_int), ("n_if", C.c_int), ("self_pos", C.c_int), ("local_parity_stride", C.c_longlong),
                ("seq", C.c_void_p), ("counters", C.c_void_p), ("status", C.c_void_p), ("local_buf", C.c_void_p), ("if_rows", C.c_void_p),
                ("remote_buf", C.c_void_p * _M), ("remote_off", C.c_longlong * _M), ("remote_parity_stride", C.c_longlong * _M),
                ("remote_flag", C.c_void_p * _M), ("local_flag", C.c_void_p * _M), ("send_rows", C.c_void_p * _M), ("n_send", C.c_int * _M),
                ("recv_rows", C.c_void_p * _M), ("n_recv", C.c_int * _M), ("local_off", C.c_longlong * _M)]


_P, _I, _F, _S = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_SIGS = {
    "fvgn_abi_version": (C.c_int, []),
    "fvgn_last_error": (C.c_char_p, []),
    "fvgn_device_check": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "fvgn_status_read": (_I, [C.POINTER(C.c_uint), _I]),
    "fvgn_x3_pair_enable": (_I, [_I]),
    "fvgn_connectivity_workspace_bytes": (_S, [_I]),
    "fvgn_connectivity_faces": (_I, [_P, _I, _I, _P, C.POINTER(C.c_int), _P, _S, _P]),
    "fvgn_connectivity_tables": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S, _P]),
    "fvgn_connectivity_poly_workspace_bytes": (_S, [_I]),
    "fvgn_connectivity_poly_faces": (_I, [_P, _I, _I, _P, C.POINTER(C.c_int), _P, _S, _P]),
    "fvgn_connectivity_poly_tables": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S, _P]),
    "fvgn_cell_mean_poly": (_I, [_P, _P, _I, _P, _P, _P]),
    "fvgn_interp_cells_poly": (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P]),
    "fvgn_integrator_fwd_poly": (_I, [_P, _P, _P, _P, _I, _F, _F, _P, _P, _P]),
    "fvgn_processor_train_fwd": (_I, [C.POINTER(MlpT), C.POINTER(MlpT), _I, _I, C.POINTER(TrainTablesT), _P, _P, _P, _P, _P, _P, _P, _P]),
    "fvgn_processor_train_bwd_scratch_floats": (_S, [_I, _I, _I]),
    "fvgn_processor_train_bwd": (_I, [C.POINTER(MlpT), C.POINTER(MlpT), _I, _I, _I, _I, C.POINTER(TrainTablesT), _P, _P, _P, _P, _P, _P, _P, _P,
                                     _P, _P, C.POINTER(C.c_void_p), _P, _S, _P, _P]),
    "fvgn_csr_workspace_bytes": (_S, [_I]),
    "fvgn_build_csr": (_I, [_P, _I, _I, _P, _P, _P, _S, _P]),
    "fvgn_interp_nodes": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P, _P]),
    "fvgn_assemble_edge_attr": (_I, [_P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P]),
    "fvgn_create_next_graph": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "fvgn_normalizer_accumulate": (_I, [_P, _I, _I, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "fvgn_norm_workspace_floats": (_S, [_I]),
    "fvgn_normalizer_apply": (_I, [_P, _I, _I, _P, _I, _I, _I, _P, _P, _P, _I, _P, _I, _P, _I, _P]),
    "fvgn_mlp_packed_bytes": (_S, [_I]),
    "fvgn_mlp_pack_bf16": (_I, [C.POINTER(MlpT), _P, _S, _P]),
    "fvgn_mlp_packed_x3_bytes": (_S, [_I]),
    "fvgn_mlp_pack_bf16x3": (_I, [C.POINTER(MlpT), _P, _S, _P]),
    "fvgn_mlp_pack_bf16x3_many": (_I, [C.POINTER(MlpT), _I, _P]),
    "fvgn_mlp_packed_x3_bwd_bytes": (_S, [_I]),
    "fvgn_mlp_pack_bf16x3_bwd": (_I, [C.POINTER(MlpT), _P, _S, _P]),
    "fvgn_mlp_pack_bf16x3_bwd_many": (_I, [C.POINTER(MlpT), _I, _P]),
    "fvgn_ln_bwd": (_I, [_P, _I, _P, _P, _I, _P, _P, _P]),
    "fvgn_ln_bwd_sums_floats": (_S, [_I]),
    "fvgn_mlp_wgrad_x3_workspace_floats": (_S, [_I]),
    "fvgn_mlp_wgrad_x3": (_I, [C.POINTER(MlpT), _I, _P, _I, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P]),
    "fvgn_mlp_dgrad_x3": (_I, [C.POINTER(MlpT), _P, _I, _I, _P, _P, _I, _P, _I, _P]),
    "fvgn_mlp_rows_fwd": (_I, [C.POINTER(MlpT), _P, _I, _I, _P, _I, _I, _P]),
    "fvgn_node_aggregate_fwd": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "fvgn_cell_block_fwd": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _P, _P, _I, _P]),
    "fvgn_edge_block_fwd": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _P, _P, _I, _P]),
    "fvgn_cell_block_fwd_xbf16": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _P, _I, _P]),
    "fvgn_edge_block_fwd_xbf16": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _I, _P]),
    "fvgn_cell_mean_split": (_I, [_P, _P, _I, _P, _P]),
    "fvgn_cell_mean": (_I, [_P, _P, _I, _P, _P]),
    "fvgn_cell_block_fwd_split": (_I, [C.POINTER(MlpT), _P, _P, _I, _P, _P]),
    "fvgn_edge_block_fwd_split": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _P]),
    "fvgn_face_area_bn_fwd": (_I, [_P, _P, _P, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fvgn_face_area_sums": (_I, [_P, _P, _P, _I, _F, _P, _P, _P, _P]),
    "fvgn_integrator_fwd": (_I, [_P, _P, _P, _P, _I, _F, _F, _P, _P, _P]),
    "fvgn_loss_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P, _P, _P]),
    "fvgn_loss_workspace_floats": (_S, [_I]),
    "fvgn_loss_partials": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "fvgn_loss_finish": (_I, [_I, _I, _F, _F, _F, _P, _P, _P]),
    "fvgn_mlp_grad_floats": (_S, [_I]),
    "fvgn_mlp_bwd_workspace_floats": (_S, [_I]),
    "fvgn_mlp_rows_bwd": (_I, [C.POINTER(MlpT), _P, _I, _I, _P, _I, _P, _I, _P, _P, _P]),
    "fvgn_cell_block_bwd": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P]),
    "fvgn_edge_block_bwd": (_I, [C.POINTER(MlpT), _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "fvgn_segment_sum_rows": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _F, _P, _I, _P]),
    "fvgn_edge_grad_combine": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "fvgn_loss_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    "fvgn_integrator_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P]),
    "fvgn_exchange_push": (_I, [C.POINTER(ExchangePlanT), _P, C.c_longlong, _P]),
    "fvgn_exchange_wait_scatter": (_I, [C.POINTER(ExchangePlanT), _P, C.c_longlong, _P]),
    "fvgn_exchange_wait_node_sum": (_I, [C.POINTER(ExchangePlanT), _P, _P]),
    "fvgn_integrator_bwd_poly": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P]),
    "fvgn_face_area_bn_bwd": (_I, [_P, _P, _P, _I, _P, _P, _P]),
}

_lib = None


def declared_symbols():
    return sorted(_SIGS.keys())


def lib():
    """Loads the shared library once.  Raises (never degrades) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python finite-volume-graph-network_b200/build.py` "
                "(there is no CPU / eager fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.fvgn_abi_version() != 11:
            raise RuntimeError("libfvgn_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().fvgn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libfvgn_b200 {what} failed (code {rc}): {msg}")
