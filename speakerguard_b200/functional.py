"""torch.autograd bridges over the libsgb200 stage entry points.

They let unchanged host code (``loss.backward()`` in an EOT wrapper, ``torch.optim.Adam`` in CW2,
feature-level defenses sitting between the stages) differentiate through the CUDA kernels.  Each
Function's backward calls the hand-written adjoint kernel of its stage; nothing is recomputed or
approximated in PyTorch.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .engine import FLD, Engine


class MfccFn(torch.autograd.Function):
    """x [B,N] -> raw MFCC [B,m,30]  (sg_mfcc_fwd / sg_mfcc_bwd)."""

    @staticmethod
    def forward(ctx, x, eng: Engine, mode: int, dither: Optional[torch.Tensor], seed: int, pass_: int):
        x = x.contiguous()
        ctx.eng, ctx.mode, ctx.dither, ctx.seed, ctx.pass_ = eng, mode, dither, seed, pass_
        ctx.save_for_backward(x)
        return eng.mfcc_fwd(x, mode, dither, seed, pass_, ld=30)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        grad = ctx.eng.mfcc_bwd(x, g.contiguous(), ctx.mode, ctx.dither, ctx.seed, ctx.pass_)
        return grad, None, None, None, None, None


class CmvnFn(torch.autograd.Function):
    """[B,T,30] -> sliding-window mean normalised [B,T,30]  (sg_cmvn_fwd / sg_cmvn_bwd)."""

    @staticmethod
    def forward(ctx, feat, eng: Engine):
        ctx.eng = eng
        return eng.cmvn(feat.contiguous(), ld_out=30)

    @staticmethod
    def backward(ctx, g):
        return ctx.eng.cmvn(g.contiguous(), ld_out=30, backward=True), None


class EmbedFn(torch.autograd.Function):
    """CMVN features [B,T,30] -> PLDA-space embedding [B,L]  (sg_xv_embed_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, feat, eng: Engine):
        B, T, F = feat.shape
        f32 = torch.zeros(B, T, FLD, device=feat.device, dtype=torch.float32)
        f32[:, :, :F] = feat
        emb, ws = eng.embed_fwd(f32)
        ctx.eng, ctx.ws, ctx.shape = eng, ws, (B, T, F)
        return emb

    @staticmethod
    def backward(ctx, g):
        B, T, F = ctx.shape
        dfeat = ctx.eng.embed_bwd(g.contiguous(), ctx.ws, B, T)
        return dfeat[:, :, :F].contiguous(), None


class ScoreFn(torch.autograd.Function):
    """embedding [B,L] x enrolled [S,L] -> PLDA LLR scores [B,S]  (sg_plda_score_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, emb, enroll, eng: Engine):
        emb = emb.contiguous()
        en = None if enroll is None else enroll.detach().contiguous()
        scores, _ = eng.score_fwd(emb, en, want_decisions=False)
        ctx.eng, ctx.en = eng, en
        ctx.save_for_backward(emb)
        return scores

    @staticmethod
    def backward(ctx, g):
        (emb,) = ctx.saved_tensors
        return ctx.eng.score_bwd(emb, g.contiguous(), ctx.en), None, None


DITHER_MODE = {"off": _lib.DITHER_OFF, "philox": _lib.DITHER_PHILOX, "torch": _lib.DITHER_TENSOR,
               "tensor": _lib.DITHER_TENSOR}


class AnLogMelFn(torch.autograd.Function):
    """x [B,N] -> AudioNet log-mel [B,T,32]  (sg_audionet_logmel_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, eng: Engine):
        x = x.contiguous()
        ctx.eng = eng
        ctx.save_for_backward(x)
        return eng.an_logmel_fwd(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ctx.eng.an_logmel_bwd(x, g.contiguous()), None


class AnCnnFn(torch.autograd.Function):
    """log-mel [B,T,32] -> logits [B,C]  (sg_audionet_cnn_fwd / _bwd); N = samples of the waveform."""

    @staticmethod
    def forward(ctx, feat, eng: Engine, N: int):
        logits, ws = eng.an_cnn_fwd(feat.contiguous(), N)
        ctx.eng, ctx.ws, ctx.dims = eng, ws, (feat.shape[0], N)
        return logits.contiguous()

    @staticmethod
    def backward(ctx, g):
        B, N = ctx.dims
        return ctx.eng.an_cnn_bwd(g.contiguous(), ctx.ws, B, N), None, None


class AnEmbFn(torch.autograd.Function):
    """log-mel [B,T,32] -> AudioNet embedding [B,32]  (sg_audionet_emb_fwd / _bwd; extract_emb of the reference)."""

    @staticmethod
    def forward(ctx, feat, eng: Engine, N: int):
        emb, ws = eng.an_emb_fwd(feat.contiguous(), N)
        ctx.eng, ctx.ws, ctx.dims = eng, ws, (feat.shape[0], N)
        return emb

    @staticmethod
    def backward(ctx, g):
        B, N = ctx.dims
        return ctx.eng.an_emb_bwd(g.contiguous(), ctx.ws, B, N), None, None


class AnFcFn(torch.autograd.Function):
    """AudioNet embedding [B,32] -> logits [B,C]  (sg_audionet_fc_fwd / _bwd; predict_from_embeddings)."""

    @staticmethod
    def forward(ctx, emb, eng: Engine):
        ctx.eng = eng
        return eng.an_fc_fwd(emb.contiguous()).contiguous()

    @staticmethod
    def backward(ctx, g):
        return ctx.eng.an_fc_bwd(g.contiguous()), None


class Mfcc24Fn(torch.autograd.Function):
    """x [B,N] -> the first ``num_ceps`` MFCCs [B,m,num_ceps] (iv_plda.raw asks kaldi.mfcc for 24 of the
    30 cepstra, model/iv_plda.py:203-237; lifter and DCT columns do not depend on num_ceps)."""

    @staticmethod
    def forward(ctx, x, eng: Engine, mode: int, dither: Optional[torch.Tensor], seed: int, pass_: int, num_ceps: int):
        x = x.contiguous()
        ctx.eng, ctx.mode, ctx.dither, ctx.seed, ctx.pass_, ctx.nc = eng, mode, dither, seed, pass_, num_ceps
        ctx.save_for_backward(x)
        return eng.mfcc_fwd(x, mode, dither, seed, pass_, ld=30)[:, :, :num_ceps].contiguous()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g30 = torch.zeros(g.shape[0], g.shape[1], 30, device=g.device, dtype=torch.float32)
        g30[:, :, :ctx.nc] = g
        grad = ctx.eng.mfcc_bwd(x, g30, ctx.mode, ctx.dither, ctx.seed, ctx.pass_)
        return grad, None, None, None, None, None, None


class DeltaFn(torch.autograd.Function):
    """[B,T,F] -> [B,T,3F] delta + delta-delta features  (sg_add_delta_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, feat, eng: Engine):
        ctx.eng = eng
        return eng.add_delta(feat.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ctx.eng.add_delta(g.contiguous(), backward=True), None


class CmvnColsFn(torch.autograd.Function):
    """Sliding-window CMVN over any number of columns  (sg_cmvn_cols)."""

    @staticmethod
    def forward(ctx, feat, eng: Engine):
        ctx.eng = eng
        return eng.cmvn_cols(feat.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ctx.eng.cmvn_cols(g.contiguous(), backward=True), None


class IvEmbedFn(torch.autograd.Function):
    """CMVN features [B,T,72] -> PLDA-space i-vector embedding [B,L]  (sg_iv_embed_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, feat, eng: Engine):
        B, T, F = feat.shape
        emb, ws = eng.iv_embed_fwd(feat.contiguous())
        ctx.eng, ctx.ws, ctx.shape = eng, ws, (B, T, F)
        return emb

    @staticmethod
    def backward(ctx, g):
        B, T, F = ctx.shape
        return ctx.eng.iv_embed_bwd(g.contiguous(), ctx.ws, B, T), None


_INPUT_GRAD_ONLY = [0]


class input_grad_only:
    """``with input_grad_only():`` - backward passes started inside the block only need the gradient with respect to the
    waveform / features (an attack step on a trainable model): parameter-gradient kernels are skipped and the corresponding
    autograd outputs are None."""

    def __enter__(self):
        _INPUT_GRAD_ONLY[0] += 1
        return self

    def __exit__(self, *exc):
        _INPUT_GRAD_ONLY[0] -= 1
        return False


class AnCnnTrainFn(torch.autograd.Function):
    """AudioNet CNN in training mode: log-mel [B,T,32] + the 34 parameter tensors -> logits [B,C]
    (sg_audionet_train_fwd / _bwd).  BatchNorm uses batch statistics; ``running`` = (list of 8 running means, list of 8
    running variances) is updated in place with ``momentum``.  Parameter order: conv1.w, conv1.b, conv2..8 w (7),
    conv2..8 b (7), BN gamma (8), BN beta (8), fc.w, fc.b."""

    @staticmethod
    def forward(ctx, feat, eng: Engine, N: int, running, momentum: float, eps: float, *params):
        assert len(params) == 34
        ps = [p.detach().contiguous() for p in params]
        feat_c = feat.detach().contiguous()
        t = eng.an_train_struct(ps[0], ps[1], ps[2:9], ps[9:16], ps[16:24], ps[24:32], ps[32], ps[33],
                                running[0] if running else None, running[1] if running else None)
        logits, ws = eng.an_train_fwd(t, feat_c, N, momentum if running else 0.0, eps)
        ctx.eng, ctx.N, ctx.ws, ctx.ps, ctx.feat = eng, N, ws, ps, feat_c
        ctx.C = ps[33].shape[0]
        return logits[:, :ctx.C].contiguous()

    @staticmethod
    def backward(ctx, g):
        eng, ps = ctx.eng, ctx.ps
        Cp = eng.lib.sg_audionet_num_class_padded(eng._h)
        dl = torch.zeros(g.shape[0], Cp, device=g.device, dtype=torch.float32)
        dl[:, :ctx.C] = g
        want_params = any(ctx.needs_input_grad[6:]) and not _INPUT_GRAD_ONLY[0]
        grads = [torch.empty_like(p) for p in ps] if want_params else None
        t = eng.an_train_struct(ps[0], ps[1], ps[2:9], ps[9:16], ps[16:24], ps[24:32], ps[32], ps[33])
        gt = None
        if want_params:
            gt = eng.an_train_struct(grads[0], grads[1], grads[2:9], grads[9:16], grads[16:24], grads[24:32], grads[32], grads[33])
        dfeat = eng.an_train_bwd(t, ctx.feat, dl, ctx.N, ctx.ws, gt, want_dfeat=ctx.needs_input_grad[0] or not want_params)
        out = [dfeat if ctx.needs_input_grad[0] else None, None, None, None, None, None]
        out += grads if want_params else [None] * 34
        return tuple(out)
