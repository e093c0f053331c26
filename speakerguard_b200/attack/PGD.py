"""PGD (reference attack/PGD.py:11-79): FGSM's step loop inside an epsilon ball, optional
random restarts drawn with numpy on the host (kept for parity: reference quirk Q10)."""
import numpy as np
import torch

from .FGSM import FGSM


class PGD(FGSM):

    def __init__(self, model, task='CSI', epsilon=0.002, step_size=0.0004, max_iter=10, num_random_init=0,
                 loss='Entropy', targeted=False, batch_size=1, EOT_size=1, EOT_batch_size=1, verbose=1):
        self.model = model
        self.task = task
        self.epsilon = epsilon
        self.step_size = step_size
        self.max_iter = max_iter
        self.num_random_init = num_random_init
        self.loss_name = loss
        self.targeted = targeted
        self.batch_size = batch_size
        EOT_size, EOT_batch_size = max(1, EOT_size), max(1, EOT_batch_size)
        assert EOT_size % EOT_batch_size == 0, 'EOT size should be divisible by EOT batch size'
        self.EOT_size, self.EOT_batch_size = EOT_size, EOT_batch_size
        self.verbose = verbose
        self._setup()

    def attack(self, x, y):
        n_audios = self._check(x, y)
        _, n_channels, max_len = x.size()
        upper = torch.clamp(x + self.epsilon, max=1)
        lower = torch.clamp(x - self.epsilon, min=-1)
        x_ori = x.clone()
        best_rate, best_success, best_adver = -1, None, None
        for init in range(max(1, self.num_random_init)):
            x_start = x_ori
            if self.num_random_init > 0:
                noise = np.random.uniform(-self.epsilon, self.epsilon, (n_audios, n_channels, max_len))
                x_start = x_ori + torch.tensor(noise, device=x.device, dtype=x.dtype)
            adver_x, success = self._run(x_start, x_ori, y, lower, upper, self.epsilon, tag='{}-'.format(init))
            rate = sum(success) / len(success)
            if rate > best_rate:
                best_rate, best_success, best_adver = rate, success, adver_x
        return best_adver, best_success
