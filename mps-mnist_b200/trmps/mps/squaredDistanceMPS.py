"""sqMPS: squared-distance loss variant.  Reference: trmps/mps/squaredDistanceMPS.py:9-30."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np

import trmps_native as nat
from mps.mps import MPS, _to_numpy


class sqMPS(MPS):
    """Use exactly like MPS; the optimizer then descends 0.5*mean((f-y)^2) with residual (y - f)
    (baseoptimizer.py:347-352)."""

    loss_kind = nat.LOSS_SQUARED

    def cost(self, f, labels):
        f, y = _to_numpy(f).astype(np.float32), _to_numpy(labels).astype(np.float32)
        return np.float32(0.5 * np.mean(np.square(f - y)))     # squaredDistanceMPS.py:25-27

    # the linear pre-training descends 0.5 * sum((z - y)^2) (squaredDistanceMPS.py:29-30): loss_kind selects it in
    # trmps_linreg_pretrain
