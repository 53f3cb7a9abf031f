"""Datasource base class with the reference's interface (trmps/preprocessing/preprocessing.py:6-228).

Data is held as ([batch, input_size, d_feature], [batch, classes]); the model-facing accessors return
the features with the first two axes swapped, (input_size, batch, d_feature), as the MPS expects.
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np


class MPSDatasource(object):
    """Use directly by passing _training_data / _test_data, or subclass and implement
    _load_training_data / _load_test_data (then call the base method to cache to <ClassName>/*.npy)."""

    _use_cache_dir = True     # reference behaviour: mkdir <ClassName> and cache .npy files there

    def __init__(self, _expected_shape=None, shuffled=False, _training_data=None, _test_data=None):
        cls_dir = type(self).__name__
        self._training_data = _training_data
        self._test_data = _test_data
        self._expected_shape = _expected_shape
        self._training_data_path = os.path.join(cls_dir, "training_data.npy")
        self._training_labels_path = os.path.join(cls_dir, "training_labels.npy")
        self._test_data_path = os.path.join(cls_dir, "testing_data.npy")
        self._test_labels_path = os.path.join(cls_dir, "testing_labels.npy")
        if self._use_cache_dir and not os.path.isdir(cls_dir):
            os.mkdir(cls_dir)                                              # preprocessing.py:36-37
        if self._training_data is None:
            self._training_data = self._cached(self._training_data_path, self._training_labels_path)
        if self._test_data is None:
            self._test_data = self._cached(self._test_data_path, self._test_labels_path)
        self.current_index = 0
        if self._test_data is None:
            self._load_test_data()
        if self._training_data is None:
            self._load_training_data()
        if shuffled:
            self.shuffle()
        self._swapped_test_data = None
        self._swapped_training_data = None
        self._device_training = None

    def to_device(self, device=None):
        """Keep the training set resident on the GPU in its storage layout ([batch, input_size, d], [batch, classes]).
        next_training_data_batch then returns CUDA tensors produced by one trmps_batch_gather launch (slice, wrap-around
        and the (batch, N, d) -> (N, batch, d) swap of preprocessing.py:165-208) instead of a NumPy copy per step."""
        import torch
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        data, labels = self._training_data
        self._device_training = (torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(device),
                                 torch.from_numpy(np.ascontiguousarray(labels, dtype=np.float32)).to(device))
        return self

    def _cached(self, data_path, labels_path):
        # preprocessing.py:39-53: reuse the cache only if an element has the expected shape
        if self._use_cache_dir and os.path.isfile(data_path):
            data, labels = np.load(data_path), np.load(labels_path)
            if data[0].shape == self._expected_shape:
                return data, labels
        return None

    # ---- hooks for subclasses (preprocessing.py:84-97, :142-157) ----
    def _load_test_data(self):
        self._save_test_data()

    def _load_training_data(self):
        self._save_training_data()

    def _save_test_data(self):
        np.save(self._test_data_path, self._test_data[0])
        np.save(self._test_labels_path, self._test_data[1])

    def _save_training_data(self):
        np.save(self._training_data_path, self._training_data[0])
        np.save(self._training_labels_path, self._training_data[1])

    # ---- model-facing accessors ----
    @property
    def test_data(self):
        if self._test_data is None:
            self._load_test_data()
        if self._swapped_test_data is None:
            self._swapped_test_data = np.swapaxes(self._test_data[0], 0, 1)
        return self._swapped_test_data, self._test_data[1]

    @property
    def training_data(self):
        if self._training_data is None:
            self._load_training_data()
        if self._swapped_training_data is None:
            self._swapped_training_data = np.swapaxes(self._training_data[0], 0, 1)
        return self._swapped_training_data, self._training_data[1]

    def shuffle(self):
        data, labels = self._training_data
        order = np.random.permutation(len(data))
        self._training_data = data[order], labels[order]
        self._swapped_training_data = None
        if getattr(self, "_device_training", None) is not None:
            self.to_device(self._device_training[0].device)

    def next_training_data_batch(self, batch_size):
        """Next batch_size samples in dataset order, wrapping around at the end and carrying the read
        position across calls (preprocessing.py:165-208)."""
        if self._training_data is None:
            self._load_training_data()
        all_data, all_labels = self._training_data
        total = len(all_data)
        if batch_size > total:
            print("Probably shouldn't do this; your batch size is greater than the size of the dataset")
        if self._device_training is not None:
            import trmps_device as dev
            start = self.current_index % total
            feature, labels = dev.batch_gather(self._device_training[0], self._device_training[1], start, batch_size)
            # where the reference's loop leaves the read position (it stays at `total` after an exact fit)
            self.current_index = (start + batch_size) % total or (total if batch_size else start)
            return feature, labels
        pieces_x, pieces_y = [], []
        remaining = batch_size
        while remaining > 0:
            take = min(remaining, total - self.current_index)
            if take > 0:
                pieces_x.append(all_data[self.current_index:self.current_index + take])
                pieces_y.append(all_labels[self.current_index:self.current_index + take])
            remaining -= take
            self.current_index += take
            if remaining > 0:
                self.current_index = 0          # wrap
        data = np.concatenate(pieces_x, axis=0) if len(pieces_x) > 1 else np.array(pieces_x[0])
        labels = np.concatenate(pieces_y, axis=0) if len(pieces_y) > 1 else np.array(pieces_y[0])
        return np.swapaxes(data, 0, 1), labels

    @property
    def num_train_samples(self):
        return len(self._training_data[0])

    @property
    def num_test_samples(self):
        return len(self._test_data[0])

    def _save_binary(self):
        """Little-endian dump: version, N, d, classes, batch (uint16 each), then features and labels
        (preprocessing.py:99-122)."""
        base = type(self).__name__
        for name, (feat, lab), swapped in (("training_data_bin", self._training_data, self.training_data[0]),
                                           ("test_data_bin", self._test_data, self.test_data[0])):
            blob = b"".join(int(v).to_bytes(2, byteorder='little')
                            for v in (1, feat.shape[1], feat.shape[2], lab.shape[1], feat.shape[0]))
            blob += swapped.tobytes(order='C') + lab.tobytes(order='C')
            with open(os.path.join(base, name), "w+b") as fh:
                fh.write(blob)
