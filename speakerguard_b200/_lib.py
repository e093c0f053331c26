"""ctypes binding of libsgb200.so (C-ABI declared in include/sgb200.h).

There is no CPU or pure-PyTorch fallback: if the shared library is missing or a call fails, this
module raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C speakerguard_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgb200.so")

SG_OK, SG_EINVAL, SG_ECUDA, SG_ESTATE, SG_EUNSUPPORTED = 0, -1, -2, -3, -4
PREC_FP32, PREC_TF32, PREC_BF16 = 0, 1, 2
DITHER_OFF, DITHER_TENSOR, DITHER_PHILOX = 0, 1, 2
OPT_POOL_FUSION, OPT_FEAT_STASH, OPT_L1_TAP_FORM, OPT_UTT_OFFSET, OPT_CUDA_GRAPH, OPT_CMVN_FUSION, OPT_ROW_COMPACTION = 1, 2, 3, 4, 5, 6, 7
LOSS_CE, LOSS_MARGIN = 0, 1
PROF_COUNT = 16
IV_STAGE_POST, IV_STAGE_STATS, IV_STAGE_IVECTOR = 0, 1, 2
TASK_CSI, TASK_SV, TASK_OSI = 0, 1, 2
TASKS = {"CSI": TASK_CSI, "SV": TASK_SV, "OSI": TASK_OSI}
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16}
DITHERS = {"off": DITHER_OFF, "tensor": DITHER_TENSOR, "philox": DITHER_PHILOX}

_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


class XvWeights(C.Structure):
    _fields_ = [("tdnn_w", _vp * 5), ("tdnn_b", _vp * 5), ("bn_mean", _vp * 5), ("bn_var", _vp * 5),
                ("fc1_w", _vp), ("fc1_b", _vp), ("emb_mean", _vp), ("lda", _vp), ("plda_mean", _vp),
                ("plda_transform", _vp), ("plda_psi", _vp), ("enroll", _vp),
                ("L", C.c_int), ("S", C.c_int), ("bn_eps", C.c_float)]


class IvWeights(C.Structure):
    _fields_ = [("C", C.c_int), ("F", C.c_int), ("D", C.c_int), ("L", C.c_int), ("S", C.c_int),
                ("gmm_gconsts", _vp), ("gmm_means_invcovars", _vp), ("gmm_invcovars", _vp), ("ive_T", _vp),
                ("ive_sigma_inv", _vp), ("ive_offset", C.c_float), ("emb_mean", _vp), ("lda", _vp), ("plda_mean", _vp),
                ("plda_transform", _vp), ("plda_psi", _vp), ("enroll", _vp)]


class LossParams(C.Structure):
    _fields_ = [("loss", C.c_int), ("task", C.c_int), ("targeted", C.c_int), ("clip_max", C.c_int),
                ("confidence", C.c_float), ("threshold", C.c_float)]


class AudioNetWeights(C.Structure):
    _fields_ = [("conv1_w", _vp), ("conv1_b", _vp), ("conv_w", _vp * 7), ("conv_b", _vp * 7), ("bn_mean", _vp * 8),
                ("bn_var", _vp * 8), ("bn_gamma", _vp * 8), ("bn_beta", _vp * 8), ("fc_w", _vp), ("fc_b", _vp),
                ("num_class", C.c_int), ("bn_eps", C.c_float)]


class AudioNetTrainTensors(C.Structure):
    _fields_ = [("conv1_w", _vp), ("conv1_b", _vp), ("conv_w", _vp * 7), ("conv_b", _vp * 7), ("bn_gamma", _vp * 8),
                ("bn_beta", _vp * 8), ("bn_mean", _vp * 8), ("bn_var", _vp * 8), ("fc_w", _vp), ("fc_b", _vp)]


class Cw2Params(C.Structure):
    _fields_ = [("binary_search_steps", C.c_int), ("max_iter", C.c_int), ("stop_early", C.c_int),
                ("stop_early_iter", C.c_int), ("lr", C.c_float), ("initial_const", C.c_float), ("loss", LossParams),
                ("decision_threshold", C.c_float)]


class PgdParams(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("epsilon", C.c_float), ("step_size", C.c_float), ("eot_size", C.c_int),
                ("dither_mode", C.c_int), ("seed", C.c_uint64), ("loss", LossParams),
                ("decision_threshold", C.c_float), ("grad_sign", C.c_float), ("eot_batch", C.c_int),
                ("feco_ratio", C.c_float), ("feco_max_iter", C.c_int), ("feco_tol", C.c_float)]


# name -> (restype, argtypes); must list every symbol include/sgb200.h declares
PROTOTYPES = {
    "sg_version": (C.c_int, []),
    "sg_last_error": (C.c_char_p, []),
    "sg_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "sg_destroy": (None, [_vp]),
    "sg_set_precision": (C.c_int, [_vp, C.c_int]),
    "sg_get_precision": (C.c_int, [_vp]),
    "sg_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "sg_load_xv": (C.c_int, [_vp, C.POINTER(XvWeights)]),
    "sg_num_frames": (C.c_int, [C.c_int]),
    "sg_mfcc_fwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_uint64, C.c_uint64, _vp, C.c_int, _vp]),
    "sg_mfcc_bwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_uint64, C.c_uint64, _vp, C.c_int, _vp,
                              C.c_float, C.c_int, _vp]),
    "sg_dither_fill": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _vp, _vp]),
    "sg_cmvn_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "sg_cmvn_bwd": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "sg_xv_ws_bytes": (C.c_size_t, [_vp, C.c_int, C.c_int]),
    "sg_xv_embed_fwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_xv_embed_bwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_plda_score_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_float, _vp, _vp, _vp]),
    "sg_plda_score_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "sg_loss_fwd_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(LossParams), _vp, _vp, _vp]),
    "sg_step_linf": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, C.c_float, C.c_float, C.c_float, _vp]),
    "sg_pgd_ws_bytes": (C.c_size_t, [_vp, C.c_int, C.c_int]),
    "sg_pgd_run": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(PgdParams), _vp, _vp, _vp, _vp, _vp]),
    "sg_xv_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_uint64, C.c_uint64, C.c_float, _vp, _vp,
                                _vp, _vp, _vp]),
    "sg_load_audionet": (C.c_int, [_vp, C.POINTER(AudioNetWeights)]),
    "sg_audionet_num_frames": (C.c_int, [C.c_int]),
    "sg_audionet_num_class_padded": (C.c_int, [_vp]),
    "sg_audionet_ws_bytes": (C.c_size_t, [_vp, C.c_int, C.c_int, C.c_int]),
    "sg_audionet_logmel_fwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "sg_audionet_logmel_bwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_float, C.c_int, _vp]),
    "sg_audionet_cnn_fwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_audionet_cnn_bwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_audionet_emb_fwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_audionet_emb_bwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_audionet_fc_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "sg_audionet_fc_bwd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "sg_argmax_decide": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp]),
    "sg_cw2_audionet_run": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(Cw2Params), _vp, _vp, _vp, _vp, _vp]),
    "sg_cw2_last_iterations": (C.c_longlong, [_vp]),
    "sg_feco_kmeans": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_float, _vp, _vp]),
    "sg_feco_kmeans_keyed": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_float, _vp,
                                       C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "sg_feco_means_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_feco_means_bwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "sg_load_iv": (C.c_int, [_vp, C.POINTER(IvWeights)]),
    "sg_iv_ws_bytes": (C.c_size_t, [_vp, C.c_int, C.c_int]),
    "sg_iv_embed_fwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_iv_embed_bwd": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "sg_iv_stage_read": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "sg_add_delta_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "sg_add_delta_bwd": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "sg_cmvn_cols": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "sg_audionet_train_ws_bytes": (C.c_size_t, [_vp, C.c_int, C.c_int]),
    "sg_audionet_train_fwd": (C.c_int, [_vp, C.POINTER(AudioNetTrainTensors), _vp, C.c_int, C.c_int, C.c_float, C.c_float, _vp, _vp,
                                        _vp]),
    "sg_audionet_train_bwd": (C.c_int, [_vp, C.POINTER(AudioNetTrainTensors), _vp, _vp, C.c_int, C.c_int, _vp, _vp,
                                        C.POINTER(AudioNetTrainTensors), _vp]),
    "sg_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                               _vp]),
    "sg_pcm16_quantize": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "sg_wav_write_batch": (C.c_int, [C.POINTER(C.c_char_p), _vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sg_wav_read_batch": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int]),
    "sg_comm_unique_id": (C.c_int, [_vp]),
    "sg_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "sg_allreduce_metrics": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "sg_comm_destroy": (C.c_int, [_vp]),
    "sg_comm_nccl_version": (C.c_int, []),
    "sg_debug_conv": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "sg_profile_enable": (C.c_int, [_vp, C.c_int]),
    "sg_profile_read": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "sg_profile_name": (C.c_char_p, [C.c_int]),
    "sg_profile_dump": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "sg_launch_count": (C.c_longlong, [_vp]),
    "sg_reset_launch_count": (None, [_vp]),
}

_lib = None


class SgError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libsgb200.so and bind every prototype; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgError(f"{LIB_PATH} not found: build it (python -c 'import __graft_entry__ as g; g.build()'); "
                      "speakerguard_b200 has no CPU / pure-PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != SG_OK:
        msg = load().sg_last_error().decode("utf-8", "replace")
        raise SgError(f"{what or 'libsgb200'} failed (code {rc}): {msg}")
