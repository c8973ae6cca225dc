"""ctypes binding of libgmeta_b200.so (the C ABI declared in include/gmeta_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an
exception is raised -- nothing on the product path computes on the CPU or through torch ops.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GMETA_B200_LIB: another build of the same library (e.g. one compiled with per-role cycle counters for triage)
LIB_PATH = os.environ.get("GMETA_B200_LIB") or os.path.join(_HERE, "libgmeta_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

MAX_LAYERS = 3
TILE_ROWS = 128
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, IMPL_TCPAIR = 0, 1, 2, 3
AGG_GCN, AGG_MEAN, AGG_SUM = 0, 1, 2
AGGREGATIONS = {"gcn": AGG_GCN, "mean": AGG_MEAN, "sum": AGG_SUM}

i32, i64, f32, f64, vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class PackedSet(C.Structure):
    _fields_ = [(n, i32) for n in ("n_nodes", "n_edges", "n_tiles", "n_tasks", "n_subgraphs",
                                    "centres_per_subgraph")] + \
               [(n, vp) for n in ("indptr", "indices", "t_indptr", "t_indices", "tile_row0", "tile_nrows",
                                   "tile_task", "task_row_ptr", "task_sub_ptr", "centre_row", "feat_row",
                                   "labels", "norm", "norm_dst", "class_pos", "class_occ", "n_classes")] + \
               [("n_act", i32 * MAX_LAYERS), ("n_act_tiles", i32 * MAX_LAYERS)] + \
               [(n, vp * MAX_LAYERS) for n in ("act_rows", "act_task_ptr", "act_tile_row0", "act_tile_nrows",
                                               "act_tile_task", "row_pos")] + \
               [("centre_pos", vp)]


class Model(C.Structure):
    _fields_ = [("n_layers", i32), ("f_in", i32 * MAX_LAYERS), ("f_out", i32 * MAX_LAYERS),
                ("n_out", i32), ("link_pred", i32), ("n_params_padded", i32),
                ("w_off", i32 * MAX_LAYERS), ("b_off", i32 * MAX_LAYERS), ("wlin_off", i32),
                ("blin_off", i32), ("aggregation", i32)]


class StepArgs(C.Structure):
    _fields_ = [("model", Model), ("spt", PackedSet), ("qry", PackedSet), ("feat_table", vp),
                ("ld_feat", i32), ("theta", vp), ("update_step", i32), ("n_support", i32),
                ("max_classes", i32), ("spt_max_rows_per_task", i32), ("qry_max_rows_per_task", i32),
                ("update_lr", f32), ("grad_scale", f32), ("compute_meta_grad", i32), ("dense_backward", i32),
                ("impl", i32), ("pruned_forward", i32),
                ("meta_grad", vp), ("loss_q", vp), ("acc_q", vp), ("loss_s", vp), ("logits_spt0", vp),
                ("workspace", vp), ("workspace_bytes", i64), ("feat_rowmax", vp), ("aux_stream", vp),
                ("step_stats", vp)]


_SIGNATURES = {
    "gmeta_version": (C.c_int, []),
    "gmeta_error_string": (C.c_char_p, [C.c_int]),
    "gmeta_degree_norm": (C.c_int, [vp, i32, vp, vp]),
    "gmeta_gcn_layer_fwd": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i64, i32, i32, vp, i64,
                                      i32, i32, i32, vp, vp, i32, i32, vp, i64, vp]),
    "gmeta_gcn_layer_fwd_workspace_bytes": (i64, [i32, i64, i32, i32, i32]),
    "gmeta_gcn_layer_fwd_ex": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i64, i32, i32, vp, i64,
                                         i32, i32, i32, vp, vp, i32, i32, vp, i64, i32, i32, vp, vp, vp, vp]),
    "gmeta_gcn_layer_fwd_nd": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i64, i32, i32, vp, i64,
                                         i32, i32, i32, vp, vp, i32, i32, vp, i64, i32, i32, vp, vp, vp, vp]),
    "gmeta_aggregate_rows_nd": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, vp]),
    "gmeta_gcn_layer_wgrad_nd": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, i64, vp, i64,
                                           vp, i64, vp]),
    "gmeta_aggregation_norms": (C.c_int, [vp, i32, i32, vp, vp, vp]),
    "gmeta_layer_plan_bytes": (i64, [i32, i32, i32, i32]),
    "gmeta_layer_plan_build": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "gmeta_gcn_layer_fwd_ex_workspace_bytes": (i64, [i32, i64, i32, i32, i32, i32, i32, i32]),
    "gmeta_row_absmax": (C.c_int, [vp, i32, i32, i32, vp, vp]),
    "gmeta_host_pack_csr": (C.c_int, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
    "gmeta_host_pack_feat_rows": (C.c_int, [i32, vp, vp, vp, vp, vp, vp, i32]),
    "gmeta_host_active_in_neighbours": (i64, [vp, vp, vp, i64, i64, vp, vp]),
    "gmeta_host_pack_small": (C.c_int, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "gmeta_host_active_rows": (i64, [vp, vp, i64, vp, i64, vp, i32, i32, vp, vp, vp, i64, vp, vp, vp]),
    "gmeta_host_validate_labels": (C.c_int, [i32, vp, vp, vp, vp, i32]),
    "gmeta_aggregate_rows": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, vp]),
    "gmeta_gcn_layer_wgrad_workspace_bytes": (i64, [i32, i32, i32]),
    "gmeta_gcn_layer_wgrad": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, i64, vp, i64,
                                        vp, i64, vp]),
    "gmeta_readout_linear_fwd": (C.c_int, [vp, i32, i32, vp, i32, vp, i32, i32, vp, i64, vp, i64, i32, vp, vp]),
    "gmeta_readout_linear_bwd": (C.c_int, [vp, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp, i64, i32, vp, vp, i64,
                                           vp, i64, vp, vp]),
    "gmeta_build_row_pos": (C.c_int, [vp, i32, i32, vp, vp]),
    "gmeta_proto_label_prep": (C.c_int, [vp, vp, i32, vp, vp, vp, vp]),
    "gmeta_proto_loss_spt": (C.c_int, [vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, i32, vp, vp]),
    "gmeta_proto_loss_qry": (C.c_int, [vp, i32, vp, i32, vp, vp, vp, i32, i32, f32, vp, vp, i32, vp, vp, vp]),
    "gmeta_proto_grad_to_support": (C.c_int, [vp, i32, i32, vp, i32, vp, vp, i32, i32, vp, vp]),
    "gmeta_sgd_update": (C.c_int, [vp, i64, vp, f32, i32, i32, vp, vp]),
    "gmeta_sum_over_tasks": (C.c_int, [vp, vp, i32, i32, vp, vp]),
    "gmeta_adam_update": (C.c_int, [vp, vp, vp, vp, i32, f64, f64, f64, f64, i32, f32, vp, vp, vp]),
    "gmeta_adam_step": (C.c_int, [vp, vp, vp, vp, i32, f64, f64, f64, f64, vp, f32, vp, f32, vp, i32, vp, vp]),
    "gmeta_khop_workspace_bytes": (i64, [i32, i32, i32]),
    "gmeta_khop_select": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, C.c_uint64, vp, vp, vp, vp, i64, vp]),
    "gmeta_khop_build": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "gmeta_packed_set_finish_workspace_bytes": (i64, [i32, i32, i32]),
    "gmeta_packed_set_finish": (C.c_int, [vp, vp, i32, i32, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp,
                                          C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                          vp, vp, vp, i64, vp]),
    "gmeta_maml_step_workspace_bytes": (i64, [C.POINTER(StepArgs)]),
    "gmeta_maml_step": (C.c_int, [C.POINTER(StepArgs), vp]),
    "gmeta_step_graph_create": (C.c_int, [C.POINTER(vp)]),
    "gmeta_step_graph_destroy": (None, [vp]),
    "gmeta_step_graph_prepare": (C.c_int, [vp, C.POINTER(StepArgs), vp]),
    "gmeta_step_graph_launch": (C.c_int, [vp, vp]),
    "gmeta_step_graph_stats": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32)]),
    "gmeta_last_launch_count": (C.c_int, []),
    "gmeta_debug_set_tc_profile": (None, [vp]),
    "gmeta_debug_set_tc_flags": (None, [C.c_int]),
    "gmeta_debug_set_pair_flags": (None, [C.c_int]),
    "gmeta_debug_set_pair_profile": (None, [vp]),
}

_lib = None


class GMetaError(RuntimeError):
    pass


def build_library(verbose=False):
    """Compile csrc/*.cu for sm_100a into libgmeta_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise GMetaError("building libgmeta_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded shared library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise GMetaError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(or make -C gmeta_b200/csrc); gmeta_b200 has no CPU fallback" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise GMetaError("%s failed: %s (code %d)" % (what or "gmeta call", lib().gmeta_error_string(rc).decode(), rc))


def exported_symbols():
    return sorted(_SIGNATURES)
