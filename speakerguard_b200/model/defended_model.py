"""Defense plumbing around a base model (reference model/defended_model.py:11-172).

``defense`` is a list of ``[flag, callable]`` pairs; flag is the input level the callable works
at (0 waveform, 1 raw features, 2 CMVN features, ...).  'sequential' chains them level by level,
always walking every level up to the deepest one that has a defense (reference quirk Q12);
'average' averages the model outputs over the defenses.
"""
import warnings

import torch
import torch.nn as nn

sequential = 'sequential'   # model(d_n(...d_2(d_1(x))))
average = 'average'         # mean_i model(d_i(x))


class defended_model(nn.Module):

    def __init__(self, base_model, defense=None, order=sequential):
        super().__init__()
        self.base_model = base_model
        self.threshold = base_model.threshold
        if defense is not None:
            assert isinstance(defense, (list, tuple))
            assert order in [sequential, average]
            by_flag = {flag: [] for flag in self.base_model.allowed_flags}
            prev = -1
            for item in defense:
                assert isinstance(item, (list, tuple)) and len(item) == 2
                flag, method = item
                if flag not in self.base_model.allowed_flags:
                    warnings.warn('Unsupported Input Level Flag. Ignore the Defense!')
                    continue
                by_flag[flag].append(method)
                if order == sequential:
                    if flag < prev:
                        warnings.warn('Defenses given out of level order for sequential combination; re-ordered.')
                    prev = flag
            self.order = order
            self.flag2defense = by_flag
        self.defense = defense

    def _levels(self):
        return sorted(self.flag2defense.keys())

    def process_sequential(self, x):
        """x [B,1,T] -> input of the base model at its deepest level (e.g. CMVN features for xv_plda)."""
        if self.defense is None:
            return x
        xx = x
        for flag in self._levels():
            if flag == 0:
                xx = x.clone()
            elif flag == 1:
                xx = self.base_model.compute_feat(xx, flag=1)
            else:
                xx = self.base_model.comput_feat_from_feat(xx, ori_flag=flag - 1, des_flag=flag)
            for d in self.flag2defense[flag]:
                xx = d(xx)
        return xx

    def _averaged(self, x, fn):
        """mean over all defenses of fn(defended input, flag) (tuple outputs are averaged member-wise).

        Like the reference (model/defended_model.py:79-91, :108-124, :143-153) the members after the first are added through
        ``.data`` and the division is done on ``.data``: the returned *values* are the mean, but autograd only sees the first
        member's graph, with weight 1.  An adaptive attack through an 'average' ensemble therefore follows the first
        defense's gradient - kept as is, because sign steps would differ otherwise."""
        acc = None
        for flag in self._levels():
            xx = x.clone() if flag == 0 else self.base_model.compute_feat(x, flag=flag)
            for d in self.flag2defense[flag]:
                out = fn(d(xx), flag)
                out = out if isinstance(out, tuple) else (out,)
                if acc is None:
                    acc = list(out)
                else:
                    for a, o in zip(acc, out):
                        a.data += o.data
        for a in acc:
            a.data /= len(self.defense)
        return acc[0] if len(acc) == 1 else tuple(acc)

    def embedding(self, x):
        if self.defense is None:
            return self.base_model.embedding(x, flag=0)
        if self.order == sequential:
            return self.base_model.embedding(self.process_sequential(x), flag=self._levels()[-1])
        return self._averaged(x, lambda z, flag: self.base_model.embedding(z, flag=flag))

    def forward(self, x, return_emb=False, enroll_embs=None):
        if self.defense is None:
            return self.base_model(x, flag=0, return_emb=return_emb, enroll_embs=enroll_embs)
        if self.order == sequential:
            return self.base_model(self.process_sequential(x), flag=self._levels()[-1], return_emb=return_emb,
                                   enroll_embs=enroll_embs)
        logits, emb = self._averaged(x, lambda z, flag: self.base_model(z, flag=flag, return_emb=True,
                                                                        enroll_embs=enroll_embs))
        return (logits, emb) if return_emb else logits

    def score(self, x, enroll_embs=None):
        if self.defense is None:
            return self.base_model.score(x, flag=0, enroll_embs=enroll_embs)
        if self.order == sequential:
            return self.base_model.score(self.process_sequential(x), flag=self._levels()[-1], enroll_embs=enroll_embs)
        return self._averaged(x, lambda z, flag: self.base_model.score(z, flag=flag, enroll_embs=enroll_embs))

    def make_decision(self, x, enroll_embs=None):
        scores = self.score(x, enroll_embs=enroll_embs)
        decisions = torch.argmax(scores, dim=1)
        max_scores = torch.max(scores, dim=1)[0]
        decisions = torch.where(max_scores > self.base_model.threshold, decisions, torch.full_like(decisions, -1))
        return decisions, scores
