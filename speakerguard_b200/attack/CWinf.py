"""CW-inf (reference attack/CWinf.py:5-16): PGD with the loss forced to the margin loss."""
from .PGD import PGD


class CWinf(PGD):

    def __init__(self, model, task='CSI', epsilon=0.002, step_size=0.0004, max_iter=10, num_random_init=0,
                 loss='Margin', targeted=False, batch_size=1, EOT_size=1, EOT_batch_size=1, verbose=1):
        super().__init__(model, task=task, epsilon=epsilon, step_size=step_size, max_iter=max_iter,
                         num_random_init=num_random_init, loss='Margin', targeted=targeted, batch_size=batch_size,
                         EOT_size=EOT_size, EOT_batch_size=EOT_batch_size, verbose=verbose)
