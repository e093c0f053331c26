"""Abstract attack interface (reference attack/Attack.py:5-15)."""
from abc import ABCMeta, abstractmethod


class Attack(metaclass=ABCMeta):

    @abstractmethod
    def attack(self, x, y, verbose=1, EOT_size=1, EOT_batch_size=1):
        """x [n,1,N] float in [-1,1), y [n] int64 -> (adver_x [n,1,N], success list[bool])."""

    def compare(self, y, y_pred, targeted):
        """Success predicate: targeted attacks must hit y, untargeted ones must leave it."""
        hit = (y_pred == y)
        return hit.tolist() if targeted else (~hit).tolist()
