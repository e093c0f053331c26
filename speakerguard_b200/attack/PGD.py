"""PGD (reference attack/PGD.py:11-79): FGSM's step loop inside an epsilon ball, optional
random restarts drawn with numpy on the host (kept for parity: reference quirk Q10)."""
import numpy as np
import torch

from .FGSM import FGSM


class PGD(FGSM):

    def __init__(self, model, task='CSI', epsilon=0.002, step_size=0.0004, max_iter=10, num_random_init=0,
                 loss='Entropy', targeted=False, batch_size=1, EOT_size=1, EOT_batch_size=1, verbose=1):
        eot, eot_bs = max(1, EOT_size), max(1, EOT_batch_size)
        if eot % eot_bs:
            raise AssertionError('EOT size should be divisible by EOT batch size')
        for name, value in dict(model=model, task=task, epsilon=epsilon, step_size=step_size, max_iter=max_iter,
                                num_random_init=num_random_init, loss_name=loss, targeted=targeted, batch_size=batch_size,
                                EOT_size=eot, EOT_batch_size=eot_bs, verbose=verbose).items():
            setattr(self, name, value)
        self._setup()                      # loss object, grad sign, EOT wrapper, threshold (shared with FGSM)

    def attack(self, x, y):
        self._check(x, y)
        eps = self.epsilon
        x0 = x.clone()
        box = (torch.clamp(x - eps, min=-1), torch.clamp(x + eps, max=1))      # the eps-ball cut by the [-1, 1] range
        restarts = max(1, self.num_random_init)
        winner = (-1.0, None, None)                                             # (success rate, success list, adversarial batch)
        for trial in range(restarts):
            start = x0
            if self.num_random_init > 0:                                        # host numpy draw, like the reference (quirk Q10)
                start = x0 + torch.as_tensor(np.random.uniform(-eps, eps, tuple(x.shape)), device=x.device, dtype=x.dtype)
            adv, ok = self._run(start, x0, y, box[0], box[1], eps, tag=f'{trial}-')
            rate = float(sum(ok)) / len(ok)
            if rate > winner[0]:
                winner = (rate, ok, adv)
        return winner[2], winner[1]
