"""Two-site DMRG optimizer with the reference's interface (trmps/optimizers/optimizer.py:6-321)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

from optimizers.baseoptimizer import *


class MPSOptimizer(BaseOptimizer):
    """Implements the two-site sweep of Stoudenmire & Schwab on the GPU.  Usage is the reference's:

        network = MPS(d_feature, d_output, input_size); network.prepare(data_source)
        optimizer = MPSOptimizer(network, max_size, MPSOptimizerParameters(...))
        optimizer.train(data_source, batch_size, n_step, MPSTrainingParameters(rate_of_change=...))

    The methods below expose the pieces of one bond update for step-by-step use (and for the parity
    tests); each maps to the reference method of the same name.
    """

    def _setup_optimization(self):
        """Initial environments C1s[0..loc], C2s[loc+1..N-1] (optimizer.py:49-87)."""
        self._state.init_env(self.MPS._special_node_loc)

    def _update_right(self, counter):
        """Bond (counter, counter+1): merge, Armijo-guarded update, split, C1s[counter+1] (optimizer.py:144-195)."""
        self._state.bond_step(counter, +1, self._opt(self.rate_of_change))
        return counter + 1

    def _update_left(self, counter):
        """Bond (counter-1, counter) (optimizer.py:89-142)."""
        self._state.bond_step(counter, -1, self._opt(self.rate_of_change))
        return counter - 1

    def _sweep_right(self, from_index, to_index):
        self._state.sweep(from_index, to_index, +1, self._opt(self.rate_of_change))   # baseoptimizer.py:276-297

    def _sweep_left(self, from_index=None, to_index=0):
        if from_index is None:
            from_index = self.MPS.input_size - 1
        self._state.sweep(from_index, to_index, -1, self._opt(self.rate_of_change))   # baseoptimizer.py:246-274
