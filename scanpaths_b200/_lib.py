"""ctypes binding of libscanpaths_b200.so (the C ABI of include/scanpaths_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this
raises.  PyTorch is only used by callers for device memory and streams; the
library itself has no torch dependency.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libscanpaths_b200.so")


class SpbError(RuntimeError):
    pass


class ScanMatchCfg(C.Structure):
    _fields_ = [("Xres", C.c_int32), ("Yres", C.c_int32), ("Xbin", C.c_int32), ("Ybin", C.c_int32),
                ("Threshold", C.c_double), ("GapValue", C.c_double), ("TempBin", C.c_double),
                ("OffsetX", C.c_double), ("OffsetY", C.c_double)]


class ScoreCfg(C.Structure):
    _fields_ = [("sm", ScanMatchCfg),
                ("sed_height", C.c_int32), ("sed_width", C.c_int32), ("sed_n", C.c_int32), ("reserved", C.c_int32),
                ("stde_max_dim", C.c_double), ("dur_scale", C.c_double), ("max_sub", C.c_double),
                ("d_sub_delta", C.c_void_p), ("d_xlut", C.c_void_p), ("d_ylut", C.c_void_p), ("d_mask", C.c_void_p)]


class ReduceArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_scores", "d_valid", "d_pair_h", "d_pair_s", "d_len_h", "d_len_s",
                                          "d_group_count", "d_out", "d_reward", "d_group_valid", "d_acc")] + \
               [("acc_bytes", C.c_int64), ("n_groups", C.c_int64), ("group_size", C.c_int32), ("n_images", C.c_int32),
                ("min_len_valid", C.c_int32), ("acc_blocks", C.c_int32), ("mean_over_kept", C.c_int32),
                ("reserved", C.c_int32)]


class PathPack(C.Structure):
    _fields_ = [("d_sym", C.c_void_p), ("d_run", C.c_void_p), ("d_nwd", C.c_void_p), ("d_sed", C.c_void_p),
                ("d_xyn", C.c_void_p), ("d_len", C.c_void_p), ("n_paths", C.c_int64), ("lmax", C.c_int32),
                ("reserved", C.c_int32)]


class SampleGeom(C.Structure):
    _fields_ = [("map_width", C.c_int32), ("map_height", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("min_length", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/scanpaths_b200.h declares: name -> (restype, argtypes)
P, I32, I64, U64, F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
SYMBOLS = {
    "spb_version": (C.c_int, []),
    "spb_last_error": (C.c_char_p, []),
    "spb_kernel_launches": (C.c_int64, []),
    "spb_profile_enable": (C.c_int, [C.c_int32]),
    "spb_profile_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "spb_scanmatch_tables": (C.c_int, [C.POINTER(ScanMatchCfg), P, P, P, P, P]),
    "spb_prep_paths": (C.c_int, [P, P, I64, I32, C.POINTER(ScoreCfg), P, P, P, P, P, P]),
    "spb_score_workspace_bytes": (I64, [I64]),
    "spb_scanmatch_matrix": (C.c_int, [P, I32, P, I32, C.POINTER(ScoreCfg), P, P]),
    "spb_tde_work_bytes": (I64, [I32, I32]),
    "spb_tde_distances": (C.c_int, [P, I32, P, I32, P, I64, P, P]),
    "spb_score_pairs": (C.c_int, [C.POINTER(PathPack), C.POINTER(PathPack), P, P, I64, C.POINTER(ScoreCfg), P, P,
                                  I64, P, P]),
    "spb_reduce_pairs_eval": (C.c_int, [P, P, I64, I32, P, P, P]),
    "spb_reduce_acc_bytes": (I64, []),
    "spb_reduce_pairs": (C.c_int, [C.POINTER(ReduceArgs), P]),
    "spb_sample_paths": (C.c_int, [P, P, P, P, P, U64, I32, I32, I32, I32, C.POINTER(SampleGeom), P, P, P, P, P, P,
                                   P, P, P]),
    "spb_generate_scanpaths": (C.c_int, [P, P, I64, I32, C.POINTER(SampleGeom), P, P, P, P, P, P]),
}


class DecoderWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("wx_hi", "wx_lo", "wh_hi", "wh_lo", "wp_hi", "wp_lo", "ww_hi", "ww_lo", "wwx_hi", "wwx_lo", "wwx2_hi", "wwx2_lo", "bias_gate", "bias_p", "wm", "wm_hi", "wm_lo", "w2", "w3", "wd1",
                 "wd2", "w_spatial_embed", "b_spatial_embed", "w_semantic_embed", "b_semantic_embed", "wse_hi", "wse_lo",
                 "w23_hi", "w23_lo", "b23_eff", "wd_eff", "bd_eff", "w_eff_spatial", "u_semantic")] + \
               [(n, C.c_float) for n in ("b2", "b3", "bd1", "bd2_mu", "bd2_sigma", "inv_scale_x", "inv_scale_h",
                                         "inv_scale_p", "inv_scale_w", "inv_scale_wx", "inv_scale_23", "inv_scale_m", "inv_scale_se",
                                         "inv_scale_wx2")] + \
               [(n, C.c_int32) for n in ("n_streams", "n_heads", "n_weight_sets", "reserved")]


class DecoderIO(C.Structure):
    _fields_ = [("n_images", C.c_int32), ("steps", C.c_int32), ("use_tensor_cores", C.c_int32),
                ("reserved", C.c_int32), ("d_vf", C.c_void_p), ("d_att", C.c_void_p), ("d_w_row_base", C.c_void_p),
                ("d_workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("d_probs", C.c_void_p),
                ("d_mu", C.c_void_p), ("d_sigma2", C.c_void_p), ("d_action_map", C.c_void_p)]


SYMBOLS.update({
    "spb_decoder_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "spb_decode": (C.c_int, [C.POINTER(DecoderWeights), C.POINTER(DecoderIO), C.c_void_p]),
    "spb_conv_gemm": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                                   C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "spb_wino_gemm": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "spb_sal_conv_workspace_bytes": (C.c_int64, [C.c_int32]),
    "spb_sal_conv": (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "spb_get_acc_trunc_fix": (C.c_float, []),
    "spb_set_acc_trunc_fix": (C.c_int, [C.c_float]),
    "spb_get_acc_trunc_fix_fine": (C.c_float, []),
    "spb_set_acc_trunc_fix_fine": (C.c_int, [C.c_float]),
    "spb_split_fp16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_float, C.c_void_p]),
})

class ScstArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_probs", "d_mu", "d_sigma2", "d_actions", "d_dur", "d_action_mask",
                                          "d_duration_mask", "d_reward", "d_group_valid", "d_extra_adv", "d_loss",
                                          "d_adv", "d_log_actions", "d_log_durations", "d_mask_sums", "d_trial_used")] + \
               [(n, C.c_int32) for n in ("N", "T", "A", "K", "k_use", "reserved")]


SYMBOLS.update({
    "spb_loglik_rows": (C.c_int, [P] * 8 + [I32] * 4 + [P] * 4),
    "spb_loglik_rows_backward": (C.c_int, [P] * 9 + [I32] * 3 + [P] * 5),
    "spb_cross_entropy": (C.c_int, [P, P, P, I64, I32, P, P, P, P, P, P]),
    "spb_lognormal_nll": (C.c_int, [P, P, P, P, I64, P, P, P, P, P, P, P]),
    "spb_scst_loss": (C.c_int, [C.POINTER(ScstArgs), P]),
    "spb_scst_loss_backward": (C.c_int, [C.POINTER(ScstArgs), P, P, P, P, P]),
})

_lib = None


def load():
    """Loads the shared library (building is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpbError("libscanpaths_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; "
                       "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().spb_last_error().decode("utf-8", "replace")
        raise SpbError("%s failed with status %d: %s" % (what or "scanpaths_b200 call", status, msg))


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        if not t.is_contiguous():
            raise SpbError("non-contiguous tensor passed to the C ABI")
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SpbError("scanpaths_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
