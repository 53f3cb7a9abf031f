"""Fashion-MNIST datasource with the reference's interface (trmps/preprocessing/FashionMNISTpreprocessing.py:21-100).

The idx files are expected in the directory named after the class ('FashionMNISTDatasource/', as load_mnist is called
at :59); pixels are divided by 255, shrunk with a 2x2 block MAXIMUM when `shrink` (:61-65, not the average pool of the
MNIST datasource) and mapped to [1, x] (:69-74).  All of it is one launch of the fused preprocessing kernel.  The
reference keeps the result in float64; here it is float32, the precision the model is fed with either way."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np
import torch

import trmps_native as nat
import trmps_device as dev
from preprocessing.MNISTpreprocessing import MNISTDatasource, read_idx
from utils import convert_to_onehot


class FashionMNISTDatasource(MNISTDatasource):
    """Drop-in replacement for MNISTDatasource on the Fashion-MNIST files."""

    def _load_with_correct_shape(self, which):
        kind = "train" if which == "train" else "t10k"
        base = type(self).__name__
        paths = [os.path.join(base, "%s-images-idx3-ubyte.gz" % kind), os.path.join(base, "%s-labels-idx1-ubyte.gz" % kind)]
        if not all(os.path.isfile(p) for p in paths):
            raise Exception("Please place the MNIST Data in directory\n" + base + "\nWith the names\n" +
                            os.path.basename(paths[0]) + "\nAnd\n" + os.path.basename(paths[1]))      # :44-48
        images, labels, rows, cols = read_idx(*paths)
        nat.require_gpu(torch.cuda.current_device())
        device = torch.device("cuda", torch.cuda.current_device())
        phi = dev.preprocess_batch(torch.from_numpy(np.array(images)).to(device), rows, cols,
                                   nat.POOL_MAX if self.shrink else nat.POOL_NONE, 0, len(images),
                                   out_map=nat.FM_ONE_X, u8_divide=True)
        return phi.permute(1, 0, 2).contiguous().cpu().numpy(), convert_to_onehot(np.array(labels))     # :76

    def _load_test_data(self):
        self._test_data = self._load_with_correct_shape("test")
        self._save_test_data()

    def _load_training_data(self):
        self._training_data = self._load_with_correct_shape("train")
        self._save_training_data()
