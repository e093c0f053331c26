"""Thin Python owner of one libsgb200 handle (one per device).

PyTorch is used for device memory and streams only: every method takes contiguous fp32 CUDA
tensors, passes raw ``data_ptr()``s to the C-ABI on the current stream, and returns tensors
allocated with the caching allocator.  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import AudioNetTrainTensors, AudioNetWeights, Cw2Params, IvWeights, LossParams, PgdParams, XvWeights, check

FLD = 32  # internal feature row stride


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if t.device != dev:
        raise _lib.SgError(f"tensor on {t.device}, engine on {dev}")
    return t.detach().to(torch.float32).contiguous()


def make_loss_params(loss_name: str = "Entropy", targeted: bool = False, task: str = "CSI", confidence: float = 0.0,
                     threshold: Optional[float] = None, clip_max: bool = False) -> LossParams:
    """Loss selection rule of the reference's resolve_loss (attack/utils.py:104-116)."""
    if task not in _lib.TASKS:
        raise ValueError(f"task must be one of {list(_lib.TASKS)}")
    if loss_name not in ("Entropy", "Margin"):
        raise ValueError("loss must be 'Entropy' or 'Margin'")
    margin = task in ("SV", "OSI") or loss_name == "Margin"
    if threshold is None:
        # the reference's margin loss does arithmetic on the threshold for SV / OSI (attack/utils.py:48-61, :73-91) and would
        # raise a TypeError on None; CSI never reads it
        if margin and task in ("SV", "OSI"):
            raise ValueError(f"task {task} needs the model's threshold for the margin loss (got None)")
        thr = 0.0
    else:
        thr = float(threshold)                       # -inf (a model built without a threshold) is passed through as is
        if math.isnan(thr):
            raise ValueError("threshold is NaN")
    return LossParams(_lib.LOSS_MARGIN if margin else _lib.LOSS_CE, _lib.TASKS[task], int(bool(targeted)),
                      int(bool(clip_max)), float(confidence), thr)


def grad_sign_of(loss_name: str, targeted: bool) -> float:
    """The update direction resolve_loss pairs with a loss (attack/utils.py:114).  It follows the loss *name*: SV / OSI with
    loss='Entropy' run the margin loss but keep the cross-entropy sign."""
    return float(1 - 2 * int(bool(targeted))) if loss_name == "Entropy" else -1.0


class _BoundLib:
    """libsgb200's entry points, called with the engine's device current: a handle is bound to one device and its kernels,
    events and per-device function attributes must land there whatever ``torch.cuda.current_device()`` is."""

    def __init__(self, lib, device: torch.device):
        self._lib, self._device = lib, device

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        dev = self._device

        def call(*args):
            if torch.cuda.current_device() == dev.index:
                return fn(*args)
            with torch.cuda.device(dev):
                return fn(*args)

        call.__name__ = name
        setattr(self, name, call)
        return call


class Engine:
    def __init__(self, device="cuda:0", precision: str = "fp32"):
        lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.SgError(f"speakerguard_b200 runs on CUDA devices only (got '{device}'); there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.lib = _BoundLib(lib, dev)
        self._h = C.c_void_p()
        check(self.lib.sg_create(C.byref(self._h), dev.index), "sg_create")
        self.L = self.S = 0
        self.set_precision(precision)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.sg_destroy(h)
            self._h = C.c_void_p()

    # ---- configuration -----------------------------------------------------------------------
    def set_precision(self, precision: str) -> None:
        check(self.lib.sg_set_precision(self._h, _lib.PRECISIONS[precision]), "sg_set_precision")
        self.precision = precision

    def set_option(self, option: int, value: int) -> None:
        """Engine options of include/sgb200.h (SG_OPT_*)."""
        check(self.lib.sg_set_option(self._h, int(option), int(value)), "sg_set_option")

    def set_utt_offset(self, offset: int) -> None:
        """Global index of the first utterance of the batches this engine is given (SG_OPT_UTT_OFFSET): the philox dither is
        keyed on (seed, pass, global utterance, frame, sample), so rank r of a G-way contiguous split that sets r * B / G
        draws exactly the noise the unsharded run draws for those utterances."""
        if offset != getattr(self, "_utt_offset", 0):
            self.set_option(_lib.OPT_UTT_OFFSET, offset)
            self._utt_offset = int(offset)

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self) -> int:
        return int(self.lib.sg_launch_count(self._h))

    def reset_launch_count(self) -> None:
        self.lib.sg_reset_launch_count(self._h)

    def profile(self, enable: bool) -> None:
        check(self.lib.sg_profile_enable(self._h, int(enable)), "sg_profile_enable")

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """{category: (total device ms, launches)} since profile(True); synchronises."""
        out = {}
        for c in range(_lib.PROF_COUNT):
            ms, n = C.c_double(), C.c_longlong()
            check(self.lib.sg_profile_read(self._h, c, C.byref(ms), C.byref(n)), "sg_profile_read")
            out[self.lib.sg_profile_name(c).decode()] = (ms.value, n.value)
        return out

    def profile_dump(self):
        """[(category name, tag, ms)] for every launch recorded since profile(True), in launch order; synchronises."""
        n = C.c_int()
        check(self.lib.sg_profile_dump(self._h, None, None, None, 0, C.byref(n)), "sg_profile_dump")
        cats, tags, ms = (C.c_int * n.value)(), (C.c_int * n.value)(), (C.c_float * n.value)()
        check(self.lib.sg_profile_dump(self._h, cats, tags, ms, n.value, C.byref(n)), "sg_profile_dump")
        names = [self.lib.sg_profile_name(c).decode() for c in range(_lib.PROF_COUNT)]
        return [(names[cats[i]], int(tags[i]), float(ms[i])) for i in range(n.value)]

    # ---- the metric all-reduce (sg_comm.cu) -----------------------------------------------------------
    def comm_init(self, rank: int, world: int, exchange) -> None:
        """One NCCL communicator for this engine's GPU.  ``exchange(bytes_or_None) -> bytes``: a host channel that returns
        rank 0's 128-byte id on every rank (dist.py uses torch.distributed's object broadcast)."""
        buf = (C.c_char * 128)()
        if rank == 0:
            check(self.lib.sg_comm_unique_id(buf), "sg_comm_unique_id")
        ident = exchange(bytes(buf) if rank == 0 else None)
        assert len(ident) == 128
        ibuf = (C.c_char * 128).from_buffer_copy(ident)
        check(self.lib.sg_comm_init(self._h, ibuf, int(rank), int(world)), "sg_comm_init")
        self.comm_world = world

    def allreduce_metrics(self, v: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a device fp64 vector (NCCL over NVLink / NVSwitch)."""
        assert v.is_cuda and v.dtype == torch.float64 and v.is_contiguous() and v.device == self.device
        check(self.lib.sg_allreduce_metrics(self._h, _ptr(v), v.numel(), self.stream), "sg_allreduce_metrics")
        return v

    def load_xv(self, p: Dict[str, torch.Tensor], bn_eps: float = 1e-5) -> None:
        """p: 'tdnn{1..5}.weight/.bias', 'bn{1..5}.mean/.var', 'fc1.weight/.bias', 'emb_mean',
        'lda' [L,513], 'plda.mean/.transform/.psi', 'enroll' [S,L] (any device; copied to host)."""
        keep = []

        def host(name):
            a = np.ascontiguousarray(p[name].detach().cpu().numpy().astype(np.float32))
            keep.append(a)
            return a.ctypes.data

        w = XvWeights()
        for i in range(5):
            w.tdnn_w[i] = host(f"tdnn{i + 1}.weight")
            w.tdnn_b[i] = host(f"tdnn{i + 1}.bias")
            w.bn_mean[i] = host(f"bn{i + 1}.mean")
            w.bn_var[i] = host(f"bn{i + 1}.var")
        w.fc1_w, w.fc1_b = host("fc1.weight"), host("fc1.bias")
        w.emb_mean, w.lda = host("emb_mean"), host("lda")
        w.plda_mean, w.plda_transform, w.plda_psi = host("plda.mean"), host("plda.transform"), host("plda.psi")
        w.enroll = host("enroll")
        w.L, w.S, w.bn_eps = int(p["plda.mean"].shape[0]), int(p["enroll"].shape[0]), float(bn_eps)
        if tuple(p["lda"].shape) != (w.L, 513):
            raise ValueError(f"lda must be [L, 513], got {tuple(p['lda'].shape)}")
        with torch.cuda.device(self.device):
            check(self.lib.sg_load_xv(self._h, C.byref(w)), "sg_load_xv")
        self.L, self.S = w.L, w.S

    # ---- i-vector system -----------------------------------------------------------------------
    def load_iv(self, p: Dict[str, torch.Tensor]) -> None:
        """p: 'gmm.gconsts' [C], 'gmm.means_invcovars' [C,F], 'gmm.invcovars' [C,F,F], 'ive.T' [C,F,D],
        'ive.sigma_inv' [C,F,F], 'ive.offset', 'emb_mean' [D], 'lda' [L,D+1], 'plda.mean/.transform/.psi',
        'enroll' [S,L] (any device; copied to host)."""
        keep = []

        def host(name):
            a = np.ascontiguousarray(p[name].detach().cpu().numpy().astype(np.float32))
            keep.append(a)
            return a.ctypes.data

        w = IvWeights()
        Cn, F, D = (int(v) for v in p["ive.T"].shape)
        w.C, w.F, w.D = Cn, F, D
        w.L, w.S = int(p["plda.mean"].shape[0]), int(p["enroll"].shape[0])
        if tuple(p["gmm.invcovars"].shape) != (Cn, F, F) or tuple(p["ive.sigma_inv"].shape) != (Cn, F, F):
            raise ValueError("gmm.invcovars / ive.sigma_inv must be [C,F,F] matching ive.T [C,F,D]")
        if tuple(p["lda"].shape) != (w.L, D + 1):
            raise ValueError(f"lda must be [L, D+1] = [{w.L}, {D + 1}], got {tuple(p['lda'].shape)}")
        w.gmm_gconsts, w.gmm_means_invcovars = host("gmm.gconsts"), host("gmm.means_invcovars")
        w.gmm_invcovars, w.ive_T, w.ive_sigma_inv = host("gmm.invcovars"), host("ive.T"), host("ive.sigma_inv")
        w.ive_offset = float(p["ive.offset"])
        w.emb_mean, w.lda = host("emb_mean"), host("lda")
        w.plda_mean, w.plda_transform, w.plda_psi = host("plda.mean"), host("plda.transform"), host("plda.psi")
        w.enroll = host("enroll")
        with torch.cuda.device(self.device):
            check(self.lib.sg_load_iv(self._h, C.byref(w)), "sg_load_iv")
        self.L, self.S = w.L, w.S
        self.iv_dims = (Cn, F, D)

    def add_delta(self, feat: torch.Tensor, backward: bool = False) -> torch.Tensor:
        """[B,T,F] -> [B,T,3F] (or the adjoint, [B,T,3F] -> [B,T,F])."""
        feat = _f32c(feat, self.device)
        B, T, ld = feat.shape
        F = ld // 3 if backward else ld
        out = torch.empty(B, T, F if backward else 3 * F, device=self.device, dtype=torch.float32)
        fn = self.lib.sg_add_delta_bwd if backward else self.lib.sg_add_delta_fwd
        check(fn(self._h, _ptr(feat), ld, _ptr(out), out.shape[2], B, T, F, self.stream), "sg_add_delta")
        return out

    def cmvn_cols(self, feat: torch.Tensor, backward: bool = False) -> torch.Tensor:
        feat = _f32c(feat, self.device)
        B, T, ld = feat.shape
        out = torch.empty_like(feat)
        check(self.lib.sg_cmvn_cols(self._h, _ptr(feat), ld, _ptr(out), ld, ld, B, T, int(backward), self.stream), "sg_cmvn_cols")
        return out

    def iv_embed_fwd(self, feat: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """feat [B,T,F] CMVN'd MFCC+deltas -> (emb [B,L], workspace for iv_embed_bwd / iv_stage)."""
        feat = _f32c(feat, self.device)
        B, T, ld = feat.shape
        ws = self.alloc_ws(self.lib.sg_iv_ws_bytes(self._h, B, T))
        emb = torch.empty(B, self.L, device=self.device, dtype=torch.float32)
        check(self.lib.sg_iv_embed_fwd(self._h, _ptr(feat), ld, B, T, _ptr(ws), _ptr(emb), self.stream), "sg_iv_embed_fwd")
        return emb, ws

    def iv_embed_bwd(self, demb: torch.Tensor, ws: torch.Tensor, B: int, T: int) -> torch.Tensor:
        demb = _f32c(demb, self.device)
        F = self.iv_dims[1]
        dfeat = torch.empty(B, T, F, device=self.device, dtype=torch.float32)
        check(self.lib.sg_iv_embed_bwd(self._h, _ptr(demb), B, T, _ptr(ws), _ptr(dfeat), F, self.stream), "sg_iv_embed_bwd")
        return dfeat

    def iv_stage(self, ws: torch.Tensor, B: int, T: int, stage: str) -> torch.Tensor:
        Cn, F, D = self.iv_dims
        shape, code = {"post": ((B, T, Cn), _lib.IV_STAGE_POST), "stats": ((B, F + 1, Cn), _lib.IV_STAGE_STATS),
                       "ivector": ((B, D), _lib.IV_STAGE_IVECTOR)}[stage]
        out = torch.empty(*shape, device=self.device, dtype=torch.float32)
        check(self.lib.sg_iv_stage_read(self._h, _ptr(ws), B, T, code, _ptr(out), self.stream), "sg_iv_stage_read")
        return out

    # ---- AudioNet, training mode ---------------------------------------------------------------
    @staticmethod
    def an_train_struct(conv1_w, conv1_b, conv_w, conv_b, bn_gamma, bn_beta, fc_w, fc_b, bn_mean=None, bn_var=None):
        """Pack device tensors (PyTorch layouts) into sg_audionet_train_tensors."""
        t = AudioNetTrainTensors()
        t.conv1_w, t.conv1_b, t.fc_w, t.fc_b = _ptr(conv1_w), _ptr(conv1_b), _ptr(fc_w), _ptr(fc_b)
        for i in range(7):
            t.conv_w[i], t.conv_b[i] = _ptr(conv_w[i]), _ptr(conv_b[i])
        for i in range(8):
            t.bn_gamma[i], t.bn_beta[i] = _ptr(bn_gamma[i]), _ptr(bn_beta[i])
            t.bn_mean[i] = _ptr(bn_mean[i]) if bn_mean is not None else None
            t.bn_var[i] = _ptr(bn_var[i]) if bn_var is not None else None
        return t

    def an_train_fwd(self, tensors, feat: torch.Tensor, N: int, momentum: float = 0.1, eps: float = 1e-5):
        """feat [B,T,32] -> (logits [B,C], workspace) with batch-statistics BatchNorm; running statistics behind
        ``tensors`` are updated in place when momentum > 0."""
        feat = _f32c(feat, self.device)
        B = feat.shape[0]
        ws = self.alloc_ws(self.lib.sg_audionet_train_ws_bytes(self._h, B, N))
        Cp = self.lib.sg_audionet_num_class_padded(self._h)
        logits = torch.empty(B, Cp, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_train_fwd(self._h, C.byref(tensors), _ptr(feat), B, N, float(momentum), float(eps), _ptr(ws),
                                             _ptr(logits), self.stream), "sg_audionet_train_fwd")
        return logits, ws

    def an_train_bwd(self, tensors, feat: torch.Tensor, dlogits: torch.Tensor, N: int, ws: torch.Tensor, grads=None,
                     want_dfeat: bool = True):
        feat, dlogits = _f32c(feat, self.device), _f32c(dlogits, self.device)
        B = feat.shape[0]
        dfeat = torch.empty_like(feat) if want_dfeat else None
        check(self.lib.sg_audionet_train_bwd(self._h, C.byref(tensors), _ptr(feat), _ptr(dlogits), B, N, _ptr(ws), _ptr(dfeat),
                                             C.byref(grads) if grads is not None else None, self.stream), "sg_audionet_train_bwd")
        return dfeat

    def adam_step(self, param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
                  lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
        for t in (param, grad, exp_avg, exp_avg_sq):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        check(self.lib.sg_adam_step(self._h, _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr),
                                    float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), self.stream),
              "sg_adam_step")

    # ---- AudioNet ------------------------------------------------------------------------------
    AN_CONVS = ["conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv8"]

    def load_audionet(self, p: Dict[str, torch.Tensor], bn_eps: float = 1e-5) -> None:
        """p: 'conv1.weight/.bias', 'conv{2..8}.weight/.bias', '{name}.bn_mean/_var/_gamma/_beta' for
        conv1..conv8, 'fc.weight/.bias'."""
        keep = []

        def host(name):
            a = np.ascontiguousarray(p[name].detach().cpu().numpy().astype(np.float32))
            keep.append(a)
            return a.ctypes.data

        w = AudioNetWeights()
        w.conv1_w, w.conv1_b = host("conv1.weight"), host("conv1.bias")
        for i, n in enumerate(self.AN_CONVS):
            w.conv_w[i], w.conv_b[i] = host(f"{n}.weight"), host(f"{n}.bias")
        for i, n in enumerate(["conv1"] + self.AN_CONVS):
            w.bn_mean[i], w.bn_var[i] = host(f"{n}.bn_mean"), host(f"{n}.bn_var")
            w.bn_gamma[i], w.bn_beta[i] = host(f"{n}.bn_gamma"), host(f"{n}.bn_beta")
        w.fc_w, w.fc_b = host("fc.weight"), host("fc.bias")
        w.num_class, w.bn_eps = int(p["fc.bias"].shape[0]), float(bn_eps)
        with torch.cuda.device(self.device):
            check(self.lib.sg_load_audionet(self._h, C.byref(w)), "sg_load_audionet")
        self.an_classes = w.num_class
        self.an_cp = int(self.lib.sg_audionet_num_class_padded(self._h))

    def an_num_frames(self, N: int) -> int:
        return int(self.lib.sg_audionet_num_frames(N))

    def an_ws(self, B: int, N: int, for_cw2: bool = False) -> torch.Tensor:
        return self.alloc_ws(self.lib.sg_audionet_ws_bytes(self._h, B, N, int(for_cw2)))

    def an_logmel_fwd(self, x: torch.Tensor) -> torch.Tensor:
        x = _f32c(x, self.device)
        B, N = x.shape
        feat = torch.empty(B, self.an_num_frames(N), 32, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_logmel_fwd(self._h, _ptr(x), B, N, _ptr(feat), self.stream), "sg_audionet_logmel_fwd")
        return feat

    def an_logmel_bwd(self, x: torch.Tensor, dfeat: torch.Tensor, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
        x, dfeat = _f32c(x, self.device), _f32c(dfeat, self.device)
        B, N = x.shape
        if ws is None:
            ws = self.an_ws(B, N)
        dx = torch.empty(B, N, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_logmel_bwd(self._h, _ptr(x), B, N, _ptr(dfeat), _ptr(ws), _ptr(dx), 1.0, 0, self.stream),
              "sg_audionet_logmel_bwd")
        return dx

    def an_cnn_fwd(self, feat: torch.Tensor, N: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """feat [B,T,32] (T = an_num_frames(N)) -> (logits [B,C], workspace for an_cnn_bwd)."""
        feat = _f32c(feat, self.device)
        B = feat.shape[0]
        ws = self.an_ws(B, N)
        logits = torch.empty(B, self.an_cp, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_cnn_fwd(self._h, _ptr(feat), B, N, _ptr(ws), _ptr(logits), self.stream), "sg_audionet_cnn_fwd")
        return logits[:, :self.an_classes], ws

    def an_cnn_bwd(self, dlogits: torch.Tensor, ws: torch.Tensor, B: int, N: int) -> torch.Tensor:
        dl = torch.zeros(B, self.an_cp, device=self.device, dtype=torch.float32)
        dl[:, :self.an_classes] = dlogits
        dfeat = torch.empty(B, self.an_num_frames(N), 32, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_cnn_bwd(self._h, _ptr(dl), B, N, _ptr(ws), _ptr(dfeat), self.stream), "sg_audionet_cnn_bwd")
        return dfeat

    def an_emb_fwd(self, feat: torch.Tensor, N: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """feat [B,T,32] -> (emb [B,32] = extract_emb, workspace for an_emb_bwd)."""
        feat = _f32c(feat, self.device)
        B = feat.shape[0]
        ws = self.an_ws(B, N)
        emb = torch.empty(B, 32, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_emb_fwd(self._h, _ptr(feat), B, N, _ptr(ws), _ptr(emb), self.stream), "sg_audionet_emb_fwd")
        return emb, ws

    def an_emb_bwd(self, demb: torch.Tensor, ws: torch.Tensor, B: int, N: int) -> torch.Tensor:
        demb = _f32c(demb, self.device)
        dfeat = torch.empty(B, self.an_num_frames(N), 32, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_emb_bwd(self._h, _ptr(demb), B, N, _ptr(ws), _ptr(dfeat), self.stream), "sg_audionet_emb_bwd")
        return dfeat

    def an_fc_fwd(self, emb: torch.Tensor) -> torch.Tensor:
        emb = _f32c(emb, self.device)
        logits = torch.empty(emb.shape[0], self.an_cp, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_fc_fwd(self._h, _ptr(emb), emb.shape[0], _ptr(logits), self.stream), "sg_audionet_fc_fwd")
        return logits[:, :self.an_classes]

    def an_fc_bwd(self, dlogits: torch.Tensor) -> torch.Tensor:
        B = dlogits.shape[0]
        dl = torch.zeros(B, self.an_cp, device=self.device, dtype=torch.float32)
        dl[:, :self.an_classes] = dlogits
        demb = torch.empty(B, 32, device=self.device, dtype=torch.float32)
        check(self.lib.sg_audionet_fc_bwd(self._h, _ptr(dl), B, _ptr(demb), self.stream), "sg_audionet_fc_bwd")
        return demb

    def cw2_audionet_run(self, x: torch.Tensor, y: torch.Tensor, *, lp: LossParams, binary_search_steps: int, max_iter: int,
                         stop_early: bool, stop_early_iter: int, lr: float, initial_const: float,
                         decision_threshold: float = -math.inf):
        """x [B,N] -> (best adversarial x [B,N], success [B] int64, final c [B])."""
        x = _f32c(x, self.device)
        y = y.to(device=self.device, dtype=torch.int64).contiguous()
        B, N = x.shape
        ws = self.an_ws(B, N, for_cw2=True)
        best = torch.empty_like(x)
        suc = torch.empty(B, device=self.device, dtype=torch.int64)
        cst = torch.empty(B, device=self.device, dtype=torch.float32)
        pp = Cw2Params(int(binary_search_steps), int(max_iter), int(bool(stop_early)), int(stop_early_iter), float(lr),
                       float(initial_const), lp, float(decision_threshold))
        check(self.lib.sg_cw2_audionet_run(self._h, _ptr(x), _ptr(y), B, N, C.byref(pp), _ptr(ws), _ptr(best), _ptr(suc),
                                           _ptr(cst), self.stream), "sg_cw2_audionet_run")
        return best, suc, cst

    def cw2_iterations(self) -> int:
        """Gradient iterations executed by the last cw2_audionet_run (summed over the binary-search steps; early stop
        shortens a search step).  Synchronises."""
        return int(self.lib.sg_cw2_last_iterations(self._h))

    # ---- FeCo ------------------------------------------------------------------------------------
    def feco_kmeans(self, feat: torch.Tensor, k: int, seed: int = 0, max_iter: int = 100, tol: float = 0.01,
                    pass_: int = 0, utt_offset: int = 0, copy_rows: int = 0) -> torch.Tensor:
        """feat [B,n,dim] -> cluster ids [B,n] int32.  pass_ / utt_offset / copy_rows key the random stream of a row the way
        sg_pgd_run does (sg_feco_kmeans_keyed): the clustering of pass `pass_`, rows = `copy_rows` utterances repeated, utterance
        indices offset by `utt_offset`; all zero: the plain per-row stream."""
        feat = _f32c(feat, self.device)
        B, n, dim = feat.shape
        ids = torch.empty(B, n, device=self.device, dtype=torch.int32)
        if pass_ or utt_offset or copy_rows:
            check(self.lib.sg_feco_kmeans_keyed(self._h, _ptr(feat), dim, B, n, dim, k, seed, max_iter, tol, _ptr(ids),
                                                int(pass_), int(utt_offset), int(copy_rows), self.stream), "sg_feco_kmeans_keyed")
        else:
            check(self.lib.sg_feco_kmeans(self._h, _ptr(feat), dim, B, n, dim, k, seed, max_iter, tol, _ptr(ids), self.stream),
                  "sg_feco_kmeans")
        return ids

    def feco_means_fwd(self, feat: torch.Tensor, ids: torch.Tensor, k: int, force: bool = True):
        feat = _f32c(feat, self.device)
        ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
        B, n, dim = feat.shape
        out = torch.empty(B, k, dim, device=self.device, dtype=torch.float32)
        counts = torch.empty(B, k, device=self.device, dtype=torch.int32)
        check(self.lib.sg_feco_means_fwd(self._h, _ptr(feat), dim, _ptr(ids), B, n, dim, k, int(force), _ptr(out), _ptr(counts),
                                         self.stream), "sg_feco_means_fwd")
        return out, counts

    def feco_means_bwd(self, dout: torch.Tensor, ids: torch.Tensor, counts: torch.Tensor, n: int, force: bool = True):
        dout = _f32c(dout, self.device)
        B, k, dim = dout.shape
        dfeat = torch.empty(B, n, dim, device=self.device, dtype=torch.float32)
        check(self.lib.sg_feco_means_bwd(self._h, _ptr(dout), _ptr(ids), _ptr(counts), B, n, dim, k, int(force), _ptr(dfeat),
                                         self.stream), "sg_feco_means_bwd")
        return dfeat

    # ---- stage ops -----------------------------------------------------------------------------
    def num_frames(self, N: int) -> int:
        return int(self.lib.sg_num_frames(N))

    def mfcc_fwd(self, x: torch.Tensor, dither_mode: int = _lib.DITHER_OFF, dither: Optional[torch.Tensor] = None,
                 seed: int = 0, pass_: int = 0, ld: int = 30) -> torch.Tensor:
        x = _f32c(x, self.device)
        B, N = x.shape
        raw = torch.empty(B, self.num_frames(N), ld, device=self.device, dtype=torch.float32)
        d = None if dither is None else _f32c(dither, self.device)
        check(self.lib.sg_mfcc_fwd(self._h, _ptr(x), B, N, dither_mode, _ptr(d), seed, pass_, _ptr(raw), ld, self.stream),
              "sg_mfcc_fwd")
        return raw

    def mfcc_bwd(self, x: torch.Tensor, draw: torch.Tensor, dither_mode: int = _lib.DITHER_OFF,
                 dither: Optional[torch.Tensor] = None, seed: int = 0, pass_: int = 0,
                 grad: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
        x, draw = _f32c(x, self.device), _f32c(draw, self.device)
        B, N = x.shape
        acc = grad is not None
        if grad is None:
            grad = torch.empty(B, N, device=self.device, dtype=torch.float32)
        d = None if dither is None else _f32c(dither, self.device)
        check(self.lib.sg_mfcc_bwd(self._h, _ptr(x), B, N, dither_mode, _ptr(d), seed, pass_, _ptr(draw), draw.shape[2],
                                   _ptr(grad), scale, int(acc), self.stream), "sg_mfcc_bwd")
        return grad

    def dither_fill(self, B: int, N: int, seed: int, pass_: int) -> torch.Tensor:
        out = torch.empty(B, self.num_frames(N), 400, device=self.device, dtype=torch.float32)
        check(self.lib.sg_dither_fill(self._h, B, N, seed, pass_, _ptr(out), self.stream), "sg_dither_fill")
        return out

    def cmvn(self, feat: torch.Tensor, ld_out: int = 30, backward: bool = False) -> torch.Tensor:
        feat = _f32c(feat, self.device)
        B, T, ld = feat.shape
        out = torch.empty(B, T, ld_out, device=self.device, dtype=torch.float32)
        fn = self.lib.sg_cmvn_bwd if backward else self.lib.sg_cmvn_fwd
        check(fn(self._h, _ptr(feat), ld, _ptr(out), ld_out, B, T, self.stream), "sg_cmvn")
        return out

    def alloc_ws(self, nbytes: int) -> torch.Tensor:
        return torch.empty(nbytes, device=self.device, dtype=torch.uint8)

    def embed_fwd(self, feat32: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """feat32 [B,T,32] -> (emb [B,L], workspace holding the activations for embed_bwd)."""
        feat32 = _f32c(feat32, self.device)
        B, T, ld = feat32.shape
        assert ld == FLD
        ws = self.alloc_ws(self.lib.sg_xv_ws_bytes(self._h, B, T))
        emb = torch.empty(B, self.L, device=self.device, dtype=torch.float32)
        check(self.lib.sg_xv_embed_fwd(self._h, _ptr(feat32), B, T, _ptr(ws), _ptr(emb), self.stream), "sg_xv_embed_fwd")
        return emb, ws

    def embed_bwd(self, demb: torch.Tensor, ws: torch.Tensor, B: int, T: int) -> torch.Tensor:
        demb = _f32c(demb, self.device)
        dfeat = torch.empty(B, T, FLD, device=self.device, dtype=torch.float32)
        check(self.lib.sg_xv_embed_bwd(self._h, _ptr(demb), B, T, _ptr(ws), _ptr(dfeat), self.stream), "sg_xv_embed_bwd")
        return dfeat

    def score_fwd(self, emb: torch.Tensor, enroll: Optional[torch.Tensor] = None, threshold: float = -math.inf,
                  want_decisions: bool = True):
        emb = _f32c(emb, self.device)
        B = emb.shape[0]
        en = None if enroll is None else _f32c(enroll, self.device)
        S = self.S if en is None else en.shape[0]
        scores = torch.empty(B, S, device=self.device, dtype=torch.float32)
        dec = torch.empty(B, device=self.device, dtype=torch.int64) if want_decisions else None
        check(self.lib.sg_plda_score_fwd(self._h, _ptr(emb), B, _ptr(en), S, float(threshold), _ptr(scores), _ptr(dec),
                                         self.stream), "sg_plda_score_fwd")
        return scores, dec

    def score_bwd(self, emb: torch.Tensor, dscores: torch.Tensor, enroll: Optional[torch.Tensor] = None) -> torch.Tensor:
        emb, dscores = _f32c(emb, self.device), _f32c(dscores, self.device)
        en = None if enroll is None else _f32c(enroll, self.device)
        B, S = dscores.shape
        demb = torch.empty_like(emb)
        check(self.lib.sg_plda_score_bwd(self._h, _ptr(emb), _ptr(dscores), B, _ptr(en), S, _ptr(demb), self.stream),
              "sg_plda_score_bwd")
        return demb

    def loss(self, scores: torch.Tensor, y: torch.Tensor, lp: LossParams, want_grad: bool = True):
        scores = _f32c(scores, self.device)
        y = y.to(device=self.device, dtype=torch.int64).contiguous()
        B, S = scores.shape
        loss = torch.empty(B, device=self.device, dtype=torch.float32)
        ds = torch.empty_like(scores) if want_grad else None
        check(self.lib.sg_loss_fwd_bwd(self._h, _ptr(scores), _ptr(y), B, S, C.byref(lp), _ptr(loss), _ptr(ds), self.stream),
              "sg_loss_fwd_bwd")
        return loss, ds

    def step_linf(self, x: torch.Tensor, x0: torch.Tensor, grad: torch.Tensor, step: float, grad_sign: float,
                  eps: float) -> None:
        """In place on x (must be contiguous fp32)."""
        assert x.is_contiguous() and x.dtype == torch.float32
        check(self.lib.sg_step_linf(self._h, _ptr(x), _ptr(_f32c(x0, self.device)), _ptr(_f32c(grad, self.device)),
                                    x.numel(), step, grad_sign, eps, self.stream), "sg_step_linf")

    # ---- fused paths -------------------------------------------------------------------------
    def xv_forward(self, x: torch.Tensor, dither_mode: int, dither: Optional[torch.Tensor], seed: int, pass_: int,
                   threshold: float = -math.inf, ws: Optional[torch.Tensor] = None):
        x = _f32c(x, self.device)
        B, N = x.shape
        if ws is None:
            ws = self.alloc_ws(self.lib.sg_pgd_ws_bytes(self._h, B, N))
        scores = torch.empty(B, self.S, device=self.device, dtype=torch.float32)
        dec = torch.empty(B, device=self.device, dtype=torch.int64)
        emb = torch.empty(B, self.L, device=self.device, dtype=torch.float32)
        d = None if dither is None else _f32c(dither, self.device)
        check(self.lib.sg_xv_forward(self._h, _ptr(x), B, N, dither_mode, _ptr(d), seed, pass_, float(threshold), _ptr(ws),
                                     _ptr(scores), _ptr(dec), _ptr(emb), self.stream), "sg_xv_forward")
        return scores, dec, emb

    def pgd_ws(self, B: int, N: int) -> torch.Tensor:
        """Scratch for one fused attack / forward call.  It only lives inside that call (the handle is not re-entrant and its
        work is stream-ordered), so the engine keeps the largest one it has handed out instead of asking the allocator for
        ~10 KB per frame on every ``attack()``."""
        n = int(self.lib.sg_pgd_ws_bytes(self._h, B, N))
        sid = torch.cuda.current_stream(self.device).cuda_stream    # like the caching allocator: never shared across streams
        cached = getattr(self, "_pgd_ws_cache", None)
        if cached is None or cached[0] != sid or cached[1].numel() < n:
            self._pgd_ws_cache = None                      # release the old block before asking for a larger one
            cached = (sid, self.alloc_ws(n))
            self._pgd_ws_cache = cached
        return cached[1][:n]

    def pgd_run(self, x_adv: torch.Tensor, x0: torch.Tensor, y: torch.Tensor, *, max_iter: int, epsilon: float,
                step_size: float, lp: LossParams, dither_mode: int = _lib.DITHER_PHILOX,
                dither: Optional[torch.Tensor] = None, seed: int = 0, eot_size: int = 1,
                decision_threshold: float = -math.inf, ws: Optional[torch.Tensor] = None, want_loss_hist: bool = False,
                grad_sign: float = 0.0, utt_offset: int = 0, eot_batch: int = 1, feco_ratio: float = 0.0,
                feco_max_iter: int = 100, feco_tol: float = 0.01):
        """x_adv [B,N] is updated in place.  Returns (decisions [B] i64, scores [B,S], loss_hist or None).
        grad_sign: the sign resolve_loss returned (0 derives it from ``lp``); utt_offset: global index of this shard's first
        utterance (keys the philox dither so that a sharded run reproduces the unsharded one bit for bit).
        eot_batch: EOT copies run as batch rows (B * eot_batch rows per pass); feco_ratio > 0: FeCo k-means compression of
        the raw features inside the loop (sg_pgd_params)."""
        assert x_adv.is_contiguous() and x_adv.dtype == torch.float32 and x_adv.device == self.device
        x0 = _f32c(x0, self.device)
        y = y.to(device=self.device, dtype=torch.int64).contiguous()
        B, N = x_adv.shape
        eot_batch = max(1, int(eot_batch))
        if ws is None:
            ws = self.pgd_ws(B * eot_batch, N)
        scores = torch.empty(B, self.S, device=self.device, dtype=torch.float32)
        dec = torch.empty(B, device=self.device, dtype=torch.int64)
        hist = torch.empty(max_iter + 1, B, device=self.device, dtype=torch.float32) if want_loss_hist else None
        d = None if dither is None else _f32c(dither, self.device)
        pp = PgdParams(int(max_iter), float(epsilon), float(step_size), int(eot_size), int(dither_mode), int(seed), lp,
                       float(decision_threshold), float(grad_sign), eot_batch, float(feco_ratio), int(feco_max_iter),
                       float(feco_tol))
        self.set_utt_offset(utt_offset)
        check(self.lib.sg_pgd_run(self._h, _ptr(x_adv), _ptr(x0), _ptr(y), _ptr(d), B, N, C.byref(pp), _ptr(ws), _ptr(dec),
                                  _ptr(scores), _ptr(hist), self.stream), "sg_pgd_run")
        return dec, scores, hist


_DEFAULT_ENGINES: Dict[int, "Engine"] = {}


def default_engine(device) -> "Engine":
    """A weight-less engine per device for the stateless entry points (losses, sign step)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.SgError(f"speakerguard_b200 ops need CUDA tensors (got a tensor on '{dev}'); no CPU fallback")
    idx = torch.cuda.current_device() if dev.index is None else dev.index
    if idx not in _DEFAULT_ENGINES:
        _DEFAULT_ENGINES[idx] = Engine(torch.device("cuda", idx))
    return _DEFAULT_ENGINES[idx]


def debug_conv(eng: "Engine", precision: str, A: torch.Tensor, W: torch.Tensor, bias, rows: int, N: int, cin: int,
               taps: int, tap_step: int, epilogue: int, mask=None, T: int = 1, t_valid: int = 0,
               op_bf16: bool = False, out_bf16: bool = False) -> torch.Tensor:
    """One conv-as-GEMM launch through sg_debug_conv.  A [rows, cin], W [taps*cin, N] (fp32 on entry;
    converted to bf16 here when op_bf16)."""
    A, W = _f32c(A, eng.device), _f32c(W, eng.device)
    Wk = W.t().contiguous()
    if op_bf16:
        A, Wk = A.to(torch.bfloat16).contiguous(), Wk.to(torch.bfloat16).contiguous()
    out = torch.empty(rows, N, device=eng.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    b = None if bias is None else _f32c(bias, eng.device)
    mk = None if mask is None else _f32c(mask, eng.device)
    check(eng.lib.sg_debug_conv(eng._h, _lib.PRECISIONS[precision], _ptr(A), A.shape[1], _ptr(W), _ptr(Wk), _ptr(b),
                                _ptr(out), N, rows, N, cin, taps, tap_step, epilogue, _ptr(mk),
                                0 if mk is None else mk.shape[1], T, t_valid, int(op_bf16), int(out_bf16), eng.stream),
          "sg_debug_conv")
    return out
