"""MNIST datasource with the reference's interface (trmps/preprocessing/MNISTpreprocessing.py:14-160), without TensorFlow.

The reference pushes every image through its own sess.run (2x2 average pool, flatten, feature map: 70 000 graph
executions) and reads MNIST with TensorFlow's tutorial loader.  Here the idx files are read directly and the whole
set goes through ONE launch of the fused preprocessing kernel (`trmps_preprocess_batch`, csrc/preprocess.cu); the
result is stored in the reference's layout ([batch, input_size, 2], [batch, classes]) so that everything the base
class does (npy cache, shuffle, batches) is unchanged.  No download is attempted: there is no network path here.
"""
import gzip
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np
import torch

import trmps_native as nat
import trmps_device as dev
from preprocessing.preprocessing import MPSDatasource

_FILES = {"train": ("train-images-idx3-ubyte.gz", "train-labels-idx1-ubyte.gz"),
          "test": ("t10k-images-idx3-ubyte.gz", "t10k-labels-idx1-ubyte.gz")}
VALIDATION_SIZE = 5000      # read_data_sets(..., validation_size=5000) of the TF-1.2 loader: train = images[5000:]


def read_idx(images_path, labels_path):
    """idx3-ubyte images -> (count, rows*cols) uint8, idx1-ubyte labels -> (count,) uint8 (gzip or plain)."""
    def raw(path):
        if not os.path.isfile(path):
            raise FileNotFoundError("MNIST file %s not found; place the idx files there (no download is attempted)" % path)
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rb") as fh:
            return fh.read()
    img, lab = raw(images_path), raw(labels_path)
    magic, count, rows, cols = (int.from_bytes(img[i:i + 4], "big") for i in (0, 4, 8, 12))
    if magic != 2051 or int.from_bytes(lab[0:4], "big") != 2049:
        raise ValueError("not idx3/idx1 ubyte files: %s, %s" % (images_path, labels_path))
    images = np.frombuffer(img, dtype=np.uint8, offset=16).reshape(count, rows * cols)
    labels = np.frombuffer(lab, dtype=np.uint8, offset=8)
    return images, labels, rows, cols


def one_hot(labels, n_classes=10):
    out = np.zeros((len(labels), n_classes), dtype=np.float64)       # the TF loader's one-hot labels are float64
    out[np.arange(len(labels)), labels] = 1
    return out


def _preprocess_images(images, labels, size, shrink=True, rows=28, cols=28, feature_map=nat.FM_ONE_SQRT3):
    """(data points, results) in the format ([batch, MPS input size, 2], [batch, classifications]) for the first `size`
    samples, wrapping around like 20 x next_batch(size / 20, shuffle=False) does (MNISTpreprocessing.py:14-78).
    images: (count, rows*cols) uint8 or float32.  One kernel launch on the current CUDA device."""
    nat.require_gpu(torch.cuda.current_device())
    size = 20 * int(size / 20)                                       # MNISTpreprocessing.py:68
    device = torch.device("cuda", torch.cuda.current_device())
    arr = np.array(images, dtype=np.uint8 if np.asarray(images).dtype == np.uint8 else np.float32)   # writable copy for torch
    phi = dev.preprocess_batch(torch.from_numpy(arr).to(device), rows, cols, shrink, 0, size, out_map=feature_map)
    data = phi.permute(1, 0, 2).contiguous().cpu().numpy()           # storage layout of the reference: (batch, N, 2)
    picked = np.arange(size) % len(labels)
    return data, np.asarray(labels)[picked]


class MNISTDatasource(MPSDatasource):
    """Use as any MPSDatasource.  `data_dir` (default 'MNIST_data', the directory the reference hands to
    read_data_sets) must hold the four idx(.gz) files."""

    def __init__(self, shrink=True, permuted=False, shuffled=False, add_random=False, data_dir="MNIST_data"):
        expected_shape = (196, 2) if shrink else (784, 2)
        self.shrink = shrink
        self.add_random = add_random
        self.data_dir = data_dir
        super().__init__(expected_shape, shuffled)
        if permuted:                                                 # MNISTpreprocessing.py:106-121
            print("permuting")
            permutation = np.random.permutation(self._training_data[0].shape[1])
            self._training_data = self._training_data[0][:, permutation], self._training_data[1]
            self._test_data = self._test_data[0][:, permutation], self._test_data[1]
            self._swapped_training_data = self._swapped_test_data = None

    def _read(self, which):
        names = _FILES[which]
        paths = [os.path.join(self.data_dir, n) for n in names]
        paths = [p if os.path.isfile(p) or not os.path.isfile(p[:-3]) else p[:-3] for p in paths]
        images, labels, rows, cols = read_idx(*paths)
        if which == "train":
            images, labels = images[VALIDATION_SIZE:], labels[VALIDATION_SIZE:]
        return images, one_hot(labels), rows, cols

    def _load_test_data(self):
        images, labels, rows, cols = self._read("test")
        self._test_data = _preprocess_images(images, labels, 10000, self.shrink, rows, cols)
        if self.add_random:
            self._test_data = self._add_random(*self._test_data)
        super()._load_test_data()

    def _load_training_data(self):
        images, labels, rows, cols = self._read("train")
        self._training_data = _preprocess_images(images, labels, 60000, self.shrink, rows, cols)
        if self.add_random:
            self._training_data = self._add_random(*self._training_data)
        super()._load_training_data()

    def _add_random(self, data, labels):
        # MNISTpreprocessing.py:146-160: half as many uniform-noise samples in an extra class
        count = int(data.shape[0] / 2)
        noise = np.random.random((count, data.shape[1]))
        noise_phi = np.stack([np.ones_like(noise), np.sqrt(3) * (2 * noise - 1)], axis=2)
        data = np.concatenate((data, noise_phi), axis=0)
        labels = np.concatenate((labels, np.zeros((labels.shape[0], 1))), axis=1)
        noise_labels = np.zeros((count, labels.shape[1]))
        noise_labels[:, -1] = 1
        return data, np.concatenate((labels, noise_labels), axis=0)
