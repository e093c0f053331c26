"""CW2 (reference attack/CW2.py:9-137): L2 attack in tanh space with Adam and a per-utterance
binary search on the trade-off constant.

* fused: against this package's ``audionet_csine`` (bare or in a defense-less ``defended_model``)
  the whole attack runs on the device through ``sg_cw2_audionet_run`` (the reference copies scores
  and losses to the host and loops over the batch in Python every iteration).
* generic: any other model goes through autograd over the stage kernels with ``torch.optim.Adam``;
  the bookkeeping (best-L2 tracking, binary search) is vectorised on the device with the
  reference's comparisons and the -2 sentinel, and only the early-stop test touches the host.
"""
import numpy as np
import torch

from ..engine import make_loss_params
from .FGSM import FGSM
from .utils import SEC4SR_MarginLoss


def _fused_audionet(model):
    from ..model.audionet_csine import audionet_csine
    from ..model.defended_model import defended_model
    if isinstance(model, defended_model) and model.defense is None:
        model = model.base_model
    # the device loop runs the inference-mode network (running statistics); a model in train() mode takes the generic path
    return model if isinstance(model, audionet_csine) and not model.training else None


class CW2(FGSM):

    def __init__(self, model, task='CSI', targeted=False, confidence=0., initial_const=1e-3, binary_search_steps=9,
                 max_iter=10000, stop_early=True, stop_early_iter=1000, lr=1e-2, batch_size=1, verbose=1):
        for name, value in dict(model=model, task=task, targeted=targeted, confidence=confidence, initial_const=initial_const,
                                binary_search_steps=binary_search_steps, max_iter=max_iter, stop_early=stop_early,
                                stop_early_iter=stop_early_iter, lr=lr, batch_size=batch_size, verbose=verbose,
                                threshold=None).items():
            setattr(self, name, value)
        if self.task in ['SV', 'OSI']:
            self.threshold = self.model.threshold
            print('Running white box attack for {} task, directly using the true threshold {}'.format(
                self.task, self.threshold))
        self.loss = SEC4SR_MarginLoss(targeted=self.targeted, confidence=self.confidence, task=self.task,
                                      threshold=self.threshold, clip_max=True)
        self.use_fused = True

    def _fused_batch(self, an, x_batch, y_batch):
        lp = make_loss_params("Margin", self.targeted, self.task, self.confidence, self.threshold, True)
        an._sync_engine()
        best, suc, cst = an.engine.cw2_audionet_run(
            x_batch[:, 0, :], y_batch, lp=lp, binary_search_steps=self.binary_search_steps, max_iter=self.max_iter,
            stop_early=self.stop_early, stop_early_iter=self.stop_early_iter, lr=self.lr, initial_const=self.initial_const)
        if self.verbose:
            print("final const:", cst.cpu().numpy())
        return best.unsqueeze(1), [bool(v) for v in suc.cpu().tolist()]

    def attack_batch(self, x_batch, y_batch, lower, upper, batch_id):
        an = _fused_audionet(self.model) if self.use_fused else None
        if an is not None:
            return self._fused_batch(an, x_batch, y_batch)
        n_audios = x_batch.shape[0]
        dev = x_batch.device
        const = torch.full((n_audios,), self.initial_const, dtype=torch.float, device=dev)
        lower_bound = torch.zeros(n_audios, device=dev)
        upper_bound = torch.full((n_audios,), 1e10, device=dev)
        inf = torch.full((n_audios,), float("inf"), device=dev)
        global_best_l2 = inf.clone()
        global_best_adver_x = x_batch.clone()
        global_best_score = torch.full((n_audios,), -2, dtype=torch.int64, device=dev)
        for _ in range(self.binary_search_steps):
            self.modifier = torch.zeros_like(x_batch, dtype=torch.float, requires_grad=True)
            self.optimizer = torch.optim.Adam([self.modifier], lr=self.lr)
            best_l2 = inf.clone()
            best_score = torch.full((n_audios,), -2, dtype=torch.int64, device=dev)
            continue_flag, prev_loss = True, np.inf
            for n_iter in range(self.max_iter + 1):
                if not continue_flag:
                    break
                input_x = torch.tanh(self.modifier + torch.atanh(x_batch * 0.999999))
                decisions, scores = self.model.make_decision(input_x)
                loss1 = self.loss(scores, y_batch)
                loss2 = torch.sum(torch.square(input_x - x_batch), dim=(1, 2))
                loss = const * loss1 + loss2
                if n_iter < self.max_iter:
                    loss.backward(torch.ones_like(loss))
                    self.optimizer.step()
                    self.modifier.grad.zero_()
                l1, l2 = loss1.detach(), loss2.detach()
                if self.verbose:
                    print("batch: {}, c: {}, iter: {}, loss: {}, loss1: {}, loss2: {}, y_pred: {}, y: {}".format(
                        batch_id, const.cpu().numpy(), n_iter, loss.detach().cpu().numpy().tolist(), l1.cpu().numpy().tolist(),
                        l2.cpu().numpy().tolist(), decisions.cpu().numpy(), y_batch.cpu().numpy()))
                if self.stop_early and n_iter % self.stop_early_iter == 0:
                    mean_loss = float(loss.detach().double().mean())
                    if mean_loss > 0.9999 * prev_loss:
                        print("Early Stop ! ")
                        continue_flag = False
                    prev_loss = mean_loss
                ok = l1 <= 0
                hit = ok & (l2 < best_l2)                              # IF-BRANCH-1
                best_l2 = torch.where(hit, l2, best_l2)
                best_score = torch.where(hit, decisions, best_score)
                ghit = ok & (l2 < global_best_l2)                      # IF-BRANCH-2
                global_best_l2 = torch.where(ghit, l2, global_best_l2)
                global_best_score = torch.where(ghit, decisions, global_best_score)
                global_best_adver_x = torch.where(ghit.view(-1, 1, 1), input_x.detach(), global_best_adver_x)
            succeeded = best_score != -2
            upper_bound = torch.where(succeeded, torch.minimum(upper_bound, const), upper_bound)
            lower_bound = torch.where(succeeded, lower_bound, torch.maximum(lower_bound, const))
            mid = (lower_bound + upper_bound) / 2
            const = torch.where(upper_bound < 1e9, mid, torch.where(succeeded, const, const * 10))
            if self.verbose:
                print(const.cpu().numpy(), best_l2.cpu().numpy().tolist(), global_best_l2.cpu().numpy().tolist())
        success = (global_best_score != -2).cpu().tolist()
        return global_best_adver_x, success

    def attack(self, x, y):
        self._check(x, y)
        lower = torch.full_like(x, -1.0)
        upper = torch.full_like(x, 1.0)
        n_audios = x.shape[0]
        batch_size = min(self.batch_size, n_audios)
        adver, success = [], []
        for b in range(int(np.ceil(n_audios / float(batch_size)))):
            sl = slice(b * batch_size, (b + 1) * batch_size)
            a, s = self.attack_batch(x[sl], y[sl], lower[sl], upper[sl], b)
            adver.append(a)
            success += s
        return torch.cat(adver, 0), success
