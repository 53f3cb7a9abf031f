"""Seeded synthetic MNIST-shaped datasource (no reference counterpart for the data itself: MNIST cannot be
downloaded here; the pattern follows trmps/preprocessing/randomsequencepreprocessing.py:15-34 and the
feature map is the reference's, MNISTpreprocessing.py:55)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np

from preprocessing.preprocessing import MPSDatasource

SQRT3 = np.float32(np.sqrt(3.0))


def mnist_feature_map(pixels):
    """[1, sqrt(3) (2x - 1)] per pixel (MNISTpreprocessing.py:55).  pixels (..., N) -> (..., N, 2) float32."""
    x = np.asarray(pixels, dtype=np.float32)
    return np.stack([np.ones_like(x), SQRT3 * (2 * x - 1)], axis=-1)


class SyntheticMNISTDatasource(MPSDatasource):
    """Pixels ~ U[0,1) and uniform labels; shrink=True gives 14x14 (N=196), else 28x28 (N=784), mirroring
    MNISTDatasource(shrink=...) (MNISTpreprocessing.py:90-105).  Labels are correlated with the mean of a
    label-specific pixel subset so that training has something to learn."""

    _use_cache_dir = False

    def __init__(self, shrink=True, shuffled=False, n_train=1000, n_test=200, n_classes=10, seed=0, learnable=True):
        n = 196 if shrink else 784
        rng = np.random.default_rng(seed)

        def make(count):
            pix = rng.random((count, n), dtype=np.float32)
            if learnable:
                cls = rng.integers(0, n_classes, count)
                seg = n // n_classes
                for c in range(n_classes):
                    rows = np.nonzero(cls == c)[0]
                    pix[rows, c * seg:(c + 1) * seg] = 0.5 + 0.5 * pix[rows, c * seg:(c + 1) * seg]
            else:
                cls = rng.integers(0, n_classes, count)
            onehot = np.zeros((count, n_classes), dtype=np.float32)
            onehot[np.arange(count), cls] = 1
            return mnist_feature_map(pix), onehot

        super().__init__((n, 2), shuffled, _training_data=make(n_train), _test_data=make(n_test))
