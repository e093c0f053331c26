"""CW-inf: the PGD sign-step loop driven by the margin loss instead of cross-entropy.

Interface of the reference's ``attack.CWinf.CWinf`` (attack/CWinf.py:5-16): same positional order and defaults as
``PGD`` (model, task, epsilon, step_size, max_iter, num_random_init, loss, targeted, batch_size, EOT_size,
EOT_batch_size, verbose); whatever ``loss`` the caller passes, the margin loss is used.  Against the engine's
``xv_plda`` the whole loop runs on the device (``sg_pgd_run`` with ``SG_LOSS_MARGIN``).
"""
from .PGD import PGD

_ORDER = ("task", "epsilon", "step_size", "max_iter", "num_random_init", "loss", "targeted", "batch_size", "EOT_size",
          "EOT_batch_size", "verbose")


class CWinf(PGD):

    def __init__(self, model, *args, **options):
        if len(args) > len(_ORDER):
            raise TypeError(f"CWinf takes at most {len(_ORDER) + 1} positional arguments ({len(args) + 1} given)")
        for name, value in zip(_ORDER, args):
            if name in options:
                raise TypeError(f"CWinf got multiple values for argument '{name}'")
            options[name] = value
        options["loss"] = "Margin"
        PGD.__init__(self, model, **options)
