"""ctypes binding of libgsd_b200.so (C ABI: include/gsd.h).

There is NO CPU fallback: if the shared library is missing or a call fails the product raises.  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a) — the .so is kept in-tree.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSD_LIB_PATH") or os.path.join(_HERE, "libgsd_b200.so")   # GSD_LIB_PATH: a debug build of the same sources

GSD_STATUS_WORDS = 8


class GsdError(RuntimeError):
    pass


class GsdRasterFwd(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("n_sets", C.c_int32),
        ("capacity", C.c_int64),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("bg0", C.c_void_p), ("bg1", C.c_void_p),
        ("means3D", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("colors0", C.c_void_p), ("colors1", C.c_void_p),
        ("out_color", C.c_void_p), ("out_depth", C.c_void_p), ("radii", C.c_void_p),
        ("geom_ws", C.c_void_p), ("binning_ws", C.c_void_p), ("image_ws", C.c_void_p), ("status", C.c_void_p),
        ("sticky", C.c_void_p), ("unnorm_rotations", C.c_void_p),
    ]


class GsdRasterBwd(C.Structure):
    _fields_ = [
        ("fwd", GsdRasterFwd),
        ("dL_dcolor", C.c_void_p), ("partial_ws", C.c_void_p),
        ("dL_dmeans3D", C.c_void_p), ("dL_dmeans2D", C.c_void_p), ("dL_dcolors0", C.c_void_p),
        ("dL_dcolors1", C.c_void_p), ("dL_dopacities", C.c_void_p), ("dL_dscales", C.c_void_p),
        ("dL_drotations", C.c_void_p), ("prefix_done", C.c_int32),
    ]


GSD_ADAM_MAX_TENSORS = 16


class GsdTrackLosses(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("Gf", C.c_int32), ("K", C.c_int32), ("Gb", C.c_int32),
        ("means3D", C.c_void_p), ("rotations", C.c_void_p), ("fg_index", C.c_void_p), ("prev_inv_rot", C.c_void_p),
        ("neighbor_indices", C.c_void_p), ("neighbor_weight", C.c_void_p), ("neighbor_dist", C.c_void_p),
        ("prev_offset", C.c_void_p), ("in_ptr", C.c_void_p), ("in_edge", C.c_void_p), ("edge_records", C.c_void_p),
        ("bg_index", C.c_void_p),
        ("init_bg_pts", C.c_void_p), ("init_bg_rot", C.c_void_p),
        ("w_rigid", C.c_float), ("w_rot", C.c_float), ("w_iso", C.c_float), ("w_floor", C.c_float), ("w_bg", C.c_float),
        ("ws", C.c_void_p), ("losses", C.c_void_p), ("grad_means3D", C.c_void_p), ("grad_rotations", C.c_void_p),
        ("rotations_unnormalized", C.c_int32),
    ]


class GsdPhotometric(C.Structure):
    _fields_ = [
        ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_sets", C.c_int32),
        ("x", C.c_void_p), ("y", C.c_void_p), ("affine_log_scale", C.c_void_p), ("affine_shift", C.c_void_p),
        ("w_l1", C.c_float), ("w_ssim", C.c_float), ("set_weight", C.c_float * 2), ("ws", C.c_void_p),
        ("y_mu", C.c_void_p), ("y_s22", C.c_void_p),
    ]


class GsdTrackUpdate(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("lr_means", C.c_float),
        ("lr_rot", C.c_float),
        ("means3D", C.c_void_p), ("unnorm_rotations", C.c_void_p), ("g_means_a", C.c_void_p), ("g_means_b", C.c_void_p),
        ("g_rot_a", C.c_void_p), ("g_rot_b", C.c_void_p), ("m_means", C.c_void_p), ("v_means", C.c_void_p),
        ("m_rot", C.c_void_p), ("v_rot", C.c_void_p), ("step_means", C.c_void_p), ("step_rot", C.c_void_p),
        ("radii", C.c_void_p), ("max_2D_radius", C.c_void_p), ("seen", C.c_void_p), ("block_counter", C.c_void_p),
    ]


class GsdAdam(C.Structure):
    _fields_ = [
        ("n_tensors", C.c_int32), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("param", C.c_void_p * GSD_ADAM_MAX_TENSORS), ("grad", C.c_void_p * GSD_ADAM_MAX_TENSORS),
        ("exp_avg", C.c_void_p * GSD_ADAM_MAX_TENSORS), ("exp_avg_sq", C.c_void_p * GSD_ADAM_MAX_TENSORS),
        ("step", C.c_void_p * GSD_ADAM_MAX_TENSORS), ("lr", C.c_float * GSD_ADAM_MAX_TENSORS),
        ("numel", C.c_int64 * GSD_ADAM_MAX_TENSORS),
    ]


class GsdDensifyPlan(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("do_densify", C.c_int32), ("grad_thresh", C.c_float), ("clone_limit", C.c_float), ("prune_opacity", C.c_float),
        ("prune_big", C.c_float), ("grad_accum", C.c_void_p), ("denom", C.c_void_p), ("log_scales", C.c_void_p),
        ("logit_opacities", C.c_void_p), ("dst", C.c_void_p), ("totals", C.c_void_p),
    ]


class GsdDensifyApply(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("samples_scaled", C.c_int32), ("reset_opacity", C.c_int32), ("dst", C.c_void_p), ("totals", C.c_void_p),
        ("samples", C.c_void_p), ("p_src", C.c_void_p * 6), ("m_src", C.c_void_p * 6), ("v_src", C.c_void_p * 6),
        ("p_dst", C.c_void_p * 6), ("m_dst", C.c_void_p * 6), ("v_dst", C.c_void_p * 6), ("width", C.c_int32 * 6),
    ]


class GsdGnnEdges(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("n_tool", C.c_int32), ("topk", C.c_int32), ("connect_all", C.c_int32),
        ("capacity", C.c_int32),
        ("states", C.c_void_p), ("mask", C.c_void_p), ("tool_mask", C.c_void_p), ("adj_thresh", C.c_void_p),
        ("adj_thresh_sq_scalar", C.c_float),
        ("ws", C.c_void_p), ("row_ptr", C.c_void_p), ("n_edges", C.c_void_p), ("receivers", C.c_void_p), ("senders", C.c_void_p),
    ]


_lib = None

# every symbol include/gsd.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "gsd_last_error", "gsd_version", "gsd_launch_count", "gsd_raster_backward_stage",
    "gsd_raster_workspace_bytes", "gsd_raster_count_instances", "gsd_raster_forward", "gsd_raster_backward",
    "gsd_raster_mark_visible",
    "gsd_photometric_workspace_bytes", "gsd_photometric_forward", "gsd_photometric_backward", "gsd_photometric_stats",
    "gsd_photometric_reduce",
    "gsd_track_losses_workspace_bytes", "gsd_track_losses_fwd_bwd", "gsd_track_pack_edges", "gsd_adam_step", "gsd_track_update_radii",
    "gsd_track_normalize_rotations", "gsd_track_update", "gsd_track_backward_update", "gsd_photometric_target_stats", "gsd_track_unpack_target_u8",
    "gsd_gnn_edges_workspace_bytes", "gsd_gnn_build_edges", "gsd_gnn_edge_inputs", "gsd_gnn_aggregate_workspace_bytes",
    "gsd_gnn_aggregate", "gsd_gnn_aggregate_bwd_workspace_bytes", "gsd_gnn_aggregate_bwd", "gsd_gnn_edge_inputs_bwd", "gsd_fps", "gsd_tf32_pack", "gsd_skin_bone_transforms", "gsd_skin_apply", "gsd_knn",
    "gsd_tf32_split", "gsd_linear_tf32x3", "gsd_linear_small", "gsd_densify_plan", "gsd_densify_apply",
    "gsd_gnn_rollout_pre", "gsd_gnn_rollout_post",
]


def lib():
    """Loads the shared library; raises GsdError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsdError("libgsd_b200.so not built (%s). Run __graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    l.gsd_last_error.restype = C.c_char_p
    l.gsd_version.restype = C.c_int
    l.gsd_raster_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_size_t)]
    l.gsd_raster_count_instances.argtypes = [C.POINTER(GsdRasterFwd), C.c_void_p]
    l.gsd_raster_forward.argtypes = [C.POINTER(GsdRasterFwd), C.c_void_p]
    l.gsd_raster_backward.argtypes = [C.POINTER(GsdRasterBwd), C.c_void_p]
    l.gsd_raster_backward_stage.argtypes = [C.POINTER(GsdRasterBwd), C.c_int32, C.c_void_p]
    l.gsd_launch_count.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    l.gsd_launch_count.restype = None
    l.gsd_raster_mark_visible.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_photometric_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    l.gsd_photometric_forward.argtypes = [C.POINTER(GsdPhotometric), C.c_void_p, C.c_void_p]
    l.gsd_photometric_stats.argtypes = [C.POINTER(GsdPhotometric), C.c_void_p]
    l.gsd_photometric_reduce.argtypes = [C.POINTER(GsdPhotometric), C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_photometric_backward.argtypes = [C.POINTER(GsdPhotometric), C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_photometric_target_stats.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_track_unpack_target_u8.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_track_normalize_rotations.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_track_update.argtypes = [C.POINTER(GsdTrackUpdate), C.c_void_p]
    l.gsd_track_backward_update.argtypes = [C.POINTER(GsdRasterBwd), C.POINTER(GsdTrackUpdate), C.c_void_p]
    l.gsd_track_losses_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    l.gsd_track_losses_fwd_bwd.argtypes = [C.POINTER(GsdTrackLosses), C.c_void_p]
    l.gsd_track_pack_edges.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 6
    l.gsd_adam_step.argtypes = [C.POINTER(GsdAdam), C.c_void_p]
    l.gsd_track_update_radii.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_gnn_edges_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    l.gsd_gnn_build_edges.argtypes = [C.POINTER(GsdGnnEdges), C.c_void_p]
    l.gsd_gnn_edge_inputs.argtypes = [C.c_int32] * 7 + [C.c_void_p] * 7
    l.gsd_gnn_aggregate_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    l.gsd_gnn_aggregate.argtypes = [C.c_int32] * 5 + [C.c_void_p] * 7
    l.gsd_gnn_aggregate_bwd_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
    l.gsd_gnn_aggregate_bwd.argtypes = [C.c_int32] * 5 + [C.c_void_p] * 11
    l.gsd_gnn_edge_inputs_bwd.argtypes = [C.c_int32] * 6 + [C.c_void_p] * 6
    l.gsd_tf32_pack.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_skin_bone_transforms.argtypes = [C.c_int32] + [C.c_void_p] * 7
    l.gsd_skin_apply.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 8
    l.gsd_gnn_rollout_pre.argtypes = [C.c_int32] * 8 + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 3
    l.gsd_gnn_rollout_post.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 3 + [C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    l.gsd_tf32_split.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_linear_tf32x3.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    l.gsd_linear_small.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    l.gsd_densify_plan.argtypes = [C.POINTER(GsdDensifyPlan), C.c_void_p]
    l.gsd_densify_apply.argtypes = [C.POINTER(GsdDensifyApply), C.c_void_p]
    l.gsd_knn.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.gsd_fps.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = l
    return l


def launch_count():
    """(own kernels, CUB kernels) launched by the library so far (host-side count; graph replays are not counted)."""
    a, b = C.c_longlong(), C.c_longlong()
    lib().gsd_launch_count(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def check(rc, what):
    if rc != 0:
        raise GsdError("%s failed (%d): %s" % (what, rc, lib().gsd_last_error().decode()))
