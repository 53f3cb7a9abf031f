"""Single-site DMRG optimizer with the reference's interface (trmps/optimizers/singlesiteOptimizer.py:7-237)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

from optimizers.optimizer import *


class SingleSiteMPSOptimizer(MPSOptimizer):
    """One-site variant of the sweep: the special node at the current site is optimised alone (C is (B, d, Dl, Dr),
    singlesiteOptimizer.py:140-152), split as a (d*Dl) x (L*Dr) matrix (:202-237), and the neighbouring node absorbs
    S V (:124-125, :76-77).  Same construction, `train(data_source, batch_size, n_step, params)` and sweep schedule as
    MPSOptimizer; the Hessian variant is not available (it is not implemented in the reference either, :161-172).

    On the device the special node is treated as a two-site bond whose far site is the identity, so the update runs on
    the same contraction, Armijo and Jacobi kernels (`trmps_opt.single_site = 1`)."""

    def _opt(self, rate_of_change):
        if self.use_hessian:
            raise ValueError("SingleSiteMPSOptimizer: use_hessian is not supported (singlesiteOptimizer.py:161-172)")
        opt = super(SingleSiteMPSOptimizer, self)._opt(rate_of_change)
        opt.single_site = 1
        return opt

    def _setup_optimization(self):
        """Initial environments C1s[0..loc], C2s[loc..N-1] (singlesiteOptimizer.py:9-47)."""
        self._state.init_env(self.MPS._special_node_loc)
