"""FGSM and the shared batched step loop (reference attack/FGSM.py:10-98).

``attack(x, y)`` keeps the reference's contract.  Two execution paths:

* fused: when the model is this package's ``xv_plda`` (bare, or inside a ``defended_model`` with
  no defense), the whole loop - ``max_iter`` gradient passes, the sign step with clipping, and the
  final evaluation pass - runs on the device through ``sg_pgd_run`` with a single host sync at
  the end (the reference syncs >= 4 times per pass);
* generic: any other model (e.g. one wrapped with feature-level defenses) goes through the EOT
  wrapper and autograd over the per-stage CUDA kernels, mirroring FGSM.attack_batch step by step;
  the update itself is the ``sg_step_linf`` kernel.
"""
import math

import numpy as np
import torch

from ..adaptive_attack.EOT import EOT
from ..model.utils import known_input_range
from ..engine import default_engine, grad_sign_of, make_loss_params
from .Attack import Attack
from .. import _lib
from .utils import resolve_loss, resolve_prediction


def fused_target(model):
    """(xv, feco): the engine-backed xv_plda behind ``model`` if the fused loop applies (else None), and the FeCoDefense the
    loop has to run between the raw MFCC and CMVN (None for an undefended model).  Fusable defended models: no defense, or
    exactly one FeCoDefense (k-means, L2) at the raw-feature level, combined sequentially."""
    from ..defense.feature_level import FeCoDefense
    from ..model.defended_model import defended_model, sequential
    from ..model.xv_plda import xv_plda
    feco = None
    if isinstance(model, defended_model):
        if model.defense is not None:
            d = model.defense
            if not (len(d) == 1 and d[0][0] == 1 and isinstance(d[0][1], FeCoDefense) and d[0][1].fusable()
                    and getattr(model, "order", sequential) == sequential):
                return None, None
            feco = d[0][1]
        model = model.base_model
    if isinstance(model, xv_plda) and model.engine_enroll_current():
        return model, feco
    return None, None


class FGSM(Attack):

    def __init__(self, model, task='CSI', epsilon=0.002, loss='Entropy', targeted=False,
                 batch_size=1, EOT_size=1, EOT_batch_size=1, verbose=1):
        self.model = model                      # remember to call model.eval()
        self.task = task
        self.epsilon = epsilon
        self.loss_name = loss
        self.targeted = targeted
        self.batch_size = batch_size
        EOT_size, EOT_batch_size = max(1, EOT_size), max(1, EOT_batch_size)
        assert EOT_size % EOT_batch_size == 0, 'EOT size should be divisible by EOT batch size'
        self.EOT_size, self.EOT_batch_size = EOT_size, EOT_batch_size
        self.verbose = verbose
        self.max_iter = 1                       # single step, same loop as PGD
        self.step_size = epsilon
        self._setup()

    def _setup(self):
        self.threshold = None
        if self.task in ['SV', 'OSI']:
            self.threshold = self.model.threshold
            print('Running white box attack for {} task, directly using the true threshold {}'.format(
                self.task, self.threshold))
        self.loss, self.grad_sign = resolve_loss(loss_name=self.loss_name, targeted=self.targeted, task=self.task,
                                                 threshold=self.threshold, clip_max=False)
        self.EOT_wrapper = EOT(self.model, self.loss, self.EOT_size, self.EOT_batch_size, True)
        self.use_fused = True                   # set False to force the generic autograd path
        self.utt_offset = 0                     # global index of x[0] when the utterance axis is sharded over GPUs (dist.py)

    # ---- generic path (reference attack/FGSM.py:38-70) -----------------------------------------
    def attack_batch(self, x_batch, y_batch, lower, upper, batch_id, x0_batch=None, epsilon=None):
        x_batch = x_batch.detach().clone().contiguous()
        eng = default_engine(x_batch.device)
        success = None
        # bounds as tensors: sg_step_linf recomputes min(max(x, lower), upper) from them
        for it in range(self.max_iter + 1):
            last = it == self.max_iter
            n_b = 1 if last else int(self.EOT_size // self.EOT_batch_size)
            e_b = 1 if last else self.EOT_batch_size
            # decisions are only read back (a host sync) when they are used: the last pass, or every pass when verbose;
            # the iterate is inside [-1, 1] by construction (_check + the clipping below), so the model's range detection
            # (another host sync per pass) is told the answer
            with known_input_range("scale"):
                scores, loss, grad, decisions = self.EOT_wrapper(x_batch, y_batch, n_b, e_b, not last,
                                                                 need_decisions=last or bool(self.verbose))
            if decisions is not None:
                predict = resolve_prediction(decisions)
                target = y_batch.detach().cpu().numpy()
                success = self.compare(target, predict, self.targeted)
                if self.verbose:
                    print("batch:{} iter:{} loss: {} predict: {}, target: {}".format(
                        batch_id, it, (loss / n_b).cpu().numpy().tolist(), predict, target))
            if not last:
                if x0_batch is not None and x_batch.dtype == torch.float32:
                    # sign step + eps-ball + [-1,1] box in one kernel (bounds recomputed from x0: attack/PGD.py:48-49)
                    eng.step_linf(x_batch, x0_batch, grad, self.step_size, self.grad_sign, epsilon)
                else:
                    grad = grad / n_b
                    x_batch = x_batch + self.step_size * torch.sign(grad) * self.grad_sign
                    x_batch = torch.min(torch.max(x_batch, lower), upper).contiguous()
        return x_batch, success

    # ---- fused path -----------------------------------------------------------------------------
    def _fused_batch(self, xv, x_batch, x0_batch, y_batch, epsilon, batch_id, batch_offset=0, feco=None):
        eng = xv.engine
        B, _, N = x_batch.shape
        xa = x_batch[:, 0, :].detach().to(torch.float32).contiguous().clone()
        x0 = x0_batch[:, 0, :].detach().to(torch.float32).contiguous()
        E = self.EOT_size
        mode, dither, seed = xv.fused_dither(self.max_iter * E + 1, B, N)
        lp = make_loss_params(self.loss_name, self.targeted, self.task, 0.0, self.threshold, False)
        # EOT copies as batch rows (EOT.py:30-42): the largest divisor of E that is <= EOT_batch_size and keeps a pass at
        # <= 2048 rows (the workspace is ~25 KB per frame and row); one copy per pass with a dither tensor or a loss history
        eb = 1
        if E > 1 and mode != _lib.DITHER_TENSOR and not self.verbose:
            eb = max(d for d in range(1, self.EOT_batch_size + 1) if E % d == 0 and (d == 1 or B * d <= 2048))
        fk = {} if feco is None else dict(feco_ratio=feco.param, feco_max_iter=feco.max_iter, feco_tol=feco.tol)
        dec, scores, hist = eng.pgd_run(xa, x0, y_batch, max_iter=self.max_iter, epsilon=epsilon,
                                        step_size=self.step_size, lp=lp, dither_mode=mode, dither=dither, seed=seed,
                                        eot_size=E, eot_batch=eb, decision_threshold=xv.decision_threshold,
                                        want_loss_hist=bool(self.verbose), grad_sign=float(self.grad_sign),
                                        utt_offset=self.utt_offset + batch_offset, **fk)
        predict = dec.cpu().numpy()                              # the only host sync of the attack
        target = y_batch.detach().cpu().numpy()
        if self.verbose:
            h = hist.cpu().numpy()
            for it in range(h.shape[0]):
                print("batch:{} iter:{} loss: {}".format(batch_id, it, h[it].tolist()))
            print("batch:{} predict: {}, target: {}".format(batch_id, predict, target))
        return xa.unsqueeze(1), self.compare(target, predict, self.targeted)

    def _check(self, x, y):
        lower, upper = -1, 1
        assert lower <= x.max() < upper, 'generating adversarial examples should be done in [-1, 1) float domain'
        n_audios, n_channels, _ = x.size()
        assert n_channels == 1, 'Only Support Mono Audio'
        assert y.shape[0] == n_audios, 'The number of x and y should be equal'
        # what the reference's losses raise on (IndexError for a label >= num_spks, the SV assert of attack/utils.py:50);
        # the device loss kernel cannot raise, it would return NaN losses and leave those utterances unchanged
        if y.numel():
            y_lo, y_hi = int(y.min()), int(y.max())
            if self.task == 'SV':
                assert y_lo >= -1 and y_hi <= 0, 'SV task should not have labels out of 0 and -1'
            else:
                n_spk = getattr(self.model, 'num_spks', None) or getattr(getattr(self.model, 'base_model', None), 'num_spks', None)
                assert y_lo >= -1 and (n_spk is None or y_hi < n_spk), \
                    'labels must be -1 (imposter) or an enrolled speaker index < {}'.format(n_spk)
        return n_audios

    def _run(self, x, x0, y, lower, upper, epsilon, tag=""):
        """One sweep over the mini-batches (shared by FGSM and PGD)."""
        n_audios = x.shape[0]
        batch_size = min(self.batch_size, n_audios)
        n_batches = int(np.ceil(n_audios / float(batch_size)))
        xv, feco = fused_target(self.model) if self.use_fused else (None, None)
        adver, success = [], []
        for b in range(n_batches):
            sl = slice(b * batch_size, (b + 1) * batch_size)
            bid = '{}{}'.format(tag, b)
            if xv is not None and (feco is None or x[sl].shape[0] >= 2):      # FeCo on a batch of one drops empty clusters: generic path
                a, s = self._fused_batch(xv, x[sl], x0[sl], y[sl], epsilon, bid, batch_offset=b * batch_size, feco=feco)
            else:
                a, s = self.attack_batch(x[sl], y[sl], lower[sl], upper[sl], bid, x0_batch=x0[sl].detach().contiguous(),
                                         epsilon=epsilon)
            adver.append(a)
            success += s
        return torch.cat(adver, 0), success

    def attack(self, x, y):
        self._check(x, y)
        lower = torch.full_like(x, -1.0)
        upper = torch.full_like(x, 1.0)
        return self._run(x, x, y, lower, upper, math.inf)
