"""Common base of the attack classes (interface of the reference's attack/Attack.py:5-15).

``attack(x, y)`` takes a batch of waveforms ``x`` [n, 1, N] (float, in [-1, 1)) with labels ``y`` [n] (int64; the
true speaker for untargeted attacks, the wanted one for targeted attacks, -1 = reject for SV / OSI) and returns
``(adver_x, success)``: the adversarial batch and one bool per utterance.
"""
import abc

import numpy as np


class Attack(abc.ABC):

    @abc.abstractmethod
    def attack(self, x, y, verbose=1, EOT_size=1, EOT_batch_size=1):
        raise NotImplementedError

    @staticmethod
    def _as_array(v):
        return v if isinstance(v, np.ndarray) else np.asarray(v)

    def compare(self, y, y_pred, targeted):
        """Per-utterance success flags: a targeted attack succeeds when the decision equals ``y``, an untargeted one
        when it differs (host arrays, as produced by ``resolve_prediction``)."""
        same = np.equal(self._as_array(y_pred), self._as_array(y))
        return [bool(s) if targeted else not bool(s) for s in same]
