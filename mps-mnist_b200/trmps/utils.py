"""Helpers kept from the reference's trmps/utils.py that the hot path's callers use."""
import numpy as np


def check_nan(tensor, name, replace_nan=True):
    """NaN -> 0 guard of utils.py:9-20.  The native split never divides by a singular value, so this is
    only used on host arrays (e.g. weights loaded from disk)."""
    arr = np.asarray(tensor)
    if np.isnan(arr.sum()):
        print('{} is not finite'.format(name))
    if replace_nan:
        arr = np.where(np.isnan(arr), np.zeros_like(arr), arr)
    return arr


def list_from(array_like, length):
    """utils.py:33-50: the first `length` entries as a list."""
    return [array_like[i] for i in range(length)]


def convert_to_onehot(vector, num_classes=None):
    """utils.py:84-111."""
    vector = np.asarray(vector)
    assert vector.ndim == 1 and len(vector) > 0
    if num_classes is None:
        num_classes = int(vector.max()) + 1
    out = np.zeros((len(vector), num_classes), dtype=int)
    out[np.arange(len(vector)), vector] = 1
    return out
