"""Runs the REFERENCE's own hot-path code (trmps/mps/mps.py, trmps/optimizers/*.py, unmodified, imported from
/root/reference) under the NumPy TF1 stand-in `oracle/tf1_shim.py` on small seeded inputs and commits what it
produced as golden vectors:  tests/golden/ref_*.npz.

    python oracle/make_golden_from_reference.py          (only works where /root/reference exists)

Each file holds the inputs (pixels, labels, configuration) and the reference's outputs after
`MPSOptimizer.train(data_source, batch_size, n_step=1, ...)`: the N updated weight tensors' bond dimensions, the
logits `MPS.predict` gives with the initial and with the updated weights, and the training cost.  tests/ then check
(a) the oracle restatement and (b) the CUDA path against these vectors.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/trmps"
OUT = os.path.join(HERE, "..", "tests", "golden")

sys.path.insert(0, HERE)
import tf1_shim
tf = tf1_shim.install()
import trmps_oracle as o          # only for the seeded synthetic inputs and the feature map

sys.path.insert(0, REF)
from optimizers.optimizer import MPSOptimizer                      # noqa: E402  (reference code)
from optimizers.singlesiteOptimizer import SingleSiteMPSOptimizer  # noqa: E402
from optimizers.parameterObjects import MPSOptimizerParameters, MPSTrainingParameters   # noqa: E402
from mps.mps import MPS                                            # noqa: E402
from mps.squaredDistanceMPS import sqMPS                           # noqa: E402

CASES = {
    # name: (N, B, L, max_size, loc, lr, loss, updates_per_step, seed)
    "ref_xent_n8": (8, 96, 10, 24, None, 0.05, 0, 1, 7),
    "ref_xent_n10_loc0": (10, 64, 10, 32, 0, 0.05, 0, 1, 8),
    "ref_squared_n6": (6, 80, 10, 24, 2, 0.02, 1, 2, 9),
    # single-site optimizer (singlesiteOptimizer.py): loss code 2 = cross-entropy, single site
    "ref_single_n8": (8, 96, 10, 24, None, 0.05, 2, 1, 11),
}


class FixedBatch(object):
    """Duck-typed MPSDatasource: one fixed batch, also used as the test set."""

    def __init__(self, phi, labels):
        self.phi, self.labels = phi, labels

    @property
    def test_data(self):
        return self.phi, self.labels

    def next_training_data_batch(self, batch_size):
        assert batch_size == self.phi.shape[1]
        return self.phi, self.labels


def reference_predict(network, weights, phi):
    feature = tf.placeholder(tf.float32, shape=[network.input_size, None, network.d_feature])
    f = network.predict(feature)
    feed = network.create_feed_dict(weights)
    feed[feature] = phi
    with tf.Session() as sess:
        return sess.run(f, feed_dict=feed)


def run_case(N, B, L, max_size, loc, lr, loss, updates, seed):
    pix, lab = o.synthetic_batch(N, B, L, seed=seed)
    phi = o.feature_map(pix, o.FM_ONE_SQRT3)
    single = loss == 2
    cls = sqMPS if loss == 1 else MPS
    network = cls(2, L, N, special_node_loc=loc)
    network.prepare()                                   # default linear-model MPS (mps.py:211-219)
    f_init = reference_predict(network, None, phi)
    lr_reg = 0.99
    params = MPSOptimizerParameters(path="/tmp/ref_MPSconfig", updates_per_step=updates, lr_reg=lr_reg)
    optimizer = (SingleSiteMPSOptimizer if single else MPSOptimizer)(network, max_size, params)
    # train() scales the rate by 1/sqrt(1 - lr_reg^(i+1)) (baseoptimizer.py:109): undo it so that step 0 runs at `lr`
    rate0 = lr * np.sqrt(1 - lr_reg)
    optimizer.train(FixedBatch(phi, lab), B, 1, MPSTrainingParameters(rate_of_change=rate0, verbose_save=False))
    weights = [np.asarray(w) for w in optimizer.test]
    # after MPSOptimizer.__init__ the network's node graph is the post-sweep graph (baseoptimizer.py:221-238), so
    # evaluate the updated weights the way a user would: a fresh MPS prepared from them (mps.py:205-206, :425-443)
    fresh = cls(2, L, N, special_node_loc=int(network._special_node_loc))
    fresh.prepare(_weights=weights)
    f_final = reference_predict(fresh, None, phi)
    with tf.Session() as sess:
        cost = sess.run(network.cost(tf1_shim.convert_to_tensor(f_final), tf1_shim.convert_to_tensor(lab)))
    return dict(pixels=pix, labels=lab, loc=np.int32(network._special_node_loc), f_init=f_init, f_final=f_final,
                cost_final=np.float32(cost), dims=np.array([w.shape[-1] for w in weights]),
                cfg=np.array([x if x is not None else -1 for x in (N, B, L, max_size, loc, lr, loss, updates, seed)], dtype=np.float64))


class FakeMnistSplit(object):
    """Stands in for `read_data_sets(...).train`: next_batch(n, shuffle=False) of the TF-1.2 loader (consecutive samples,
    wrapping to the start of the set), on seeded uint8 images scaled by 1/255 like DataSet.__init__ does."""

    def __init__(self, images_u8, labels):
        self.images = o.scale_uint8(images_u8)
        self.labels = labels
        self.pos = 0

    def next_batch(self, n, shuffle=True):
        assert shuffle is False                      # MNISTpreprocessing.py:68
        rows = (self.pos + np.arange(n)) % len(self.images)
        self.pos = (self.pos + n) % len(self.images)
        return self.images[rows], self.labels[rows]


def run_preprocess_case(seed=21, count=26, size=40):
    """The reference's own _preprocess_images (MNISTpreprocessing.py:14-78), shrink on and off, on `count` seeded 28x28
    uint8 images; size > count so the loader wraps around."""
    from preprocessing.MNISTpreprocessing import _preprocess_images       # reference code
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, (count, 784), dtype=np.uint8)
    images[0, :28] = 0
    images[1, :28] = 255
    labels = np.eye(10)[rng.integers(0, 10, count)]
    rec = dict(images_u8=images, labels=labels, size=np.int32(size))
    for shrink in (True, False):
        data, lab = _preprocess_images(FakeMnistSplit(images, labels), size, shrink=shrink)
        assert data.dtype == np.float32 and np.all(data[..., 0] == 1)
        rec["phi1_shrink" if shrink else "phi1_full"] = data[..., 1]       # phi[..., 0] is the constant 1
        rec["labels_out"] = lab
    return rec


def run_linreg_case(loss, N=9, L=4, total=230, iterations=5, lr=0.05, seed=31):
    """The reference's own MPS._lin_reg (mps.py:331-399; sqMPS._cost_for_lin_reg squaredDistanceMPS.py:29-30) fed by the
    reference's own MPSDatasource.next_training_data_batch(100) (preprocessing.py:165-208) on `total` seeded samples, so the
    third minibatch wraps around the end of the set.  tf.train.GradientDescentOptimizer is the shim's numerical derivative
    of the reference's cost graph (tf1_shim._GradientDescentOptimizer), i.e. nothing of the oracle's closed-form gradient."""
    from preprocessing.preprocessing import MPSDatasource                # reference code
    pix, lab = o.synthetic_batch(N, total, L, seed=seed)                 # (N, total), (total, L)
    data = np.ascontiguousarray(np.swapaxes(o.feature_map(pix, o.FM_ONE_X), 0, 1))      # (total, N, 2)
    ds = MPSDatasource.__new__(MPSDatasource)
    ds._training_data, ds._test_data = (data, lab), (data[:40], lab[:40])
    ds._swapped_test_data = ds._swapped_training_data = None
    ds.current_index = 0
    network = (sqMPS if loss == 1 else MPS)(2, L, N)
    network._lin_reg(ds, iterations, lr)
    return dict(data=data, labels=lab, weight=np.asarray(network.weight), bias=np.asarray(network.bias),
                loc=np.int32(network._special_node_loc), next_index=np.int32(ds.current_index),
                cfg=np.array([N, L, total, iterations, lr, loss, seed], dtype=np.float64))


def write_idx(dirname, name, payload, dims, magic):
    import gzip
    header = magic.to_bytes(4, "big") + b"".join(int(x).to_bytes(4, "big") for x in dims)
    with gzip.open(os.path.join(dirname, name), "wb") as fh:
        fh.write(header + payload.tobytes())


def run_fashion_case(seed=23, count=24):
    """The reference's own FashionMNISTDatasource._load_with_correct_shape (FashionMNISTpreprocessing.py:56-77), shrink on
    and off, on seeded idx files in a scratch directory (it reads '<ClassName>/train-images-idx3-ubyte.gz')."""
    import tempfile
    from preprocessing.FashionMNISTpreprocessing import FashionMNISTDatasource, data_type      # reference code
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, (count, 784), dtype=np.uint8)
    images[0, :56] = np.arange(56) * 4                    # every residue class of u / 255 rounding shows up
    labels = rng.integers(0, 10, count).astype(np.uint8)
    labels[0] = 9
    rec = dict(images_u8=images, labels_u8=labels)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.mkdir("FashionMNISTDatasource")
            write_idx("FashionMNISTDatasource", "train-images-idx3-ubyte.gz", images, (count, 28, 28), 2051)
            write_idx("FashionMNISTDatasource", "train-labels-idx1-ubyte.gz", labels, (count,), 2049)
            for shrink in (True, False):
                ds = FashionMNISTDatasource.__new__(FashionMNISTDatasource)        # only the loader is exercised
                ds.shrink = shrink
                data, lab = ds._load_with_correct_shape(data_type.training)
                assert data.dtype == np.float64 and np.all(data[..., 0] == 1)
                rec["phi1_shrink" if shrink else "phi1_full"] = data[..., 1].astype(np.float32)   # what the float32 graph is fed
                rec["labels_out"] = lab
        finally:
            os.chdir(cwd)
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        rec = run_fashion_case()
    np.savez_compressed(os.path.join(OUT, "ref_fashion.npz"), **rec)
    print("ref_fashion", rec["phi1_shrink"].shape, rec["phi1_full"].shape, rec["labels_out"].shape)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        rec = run_preprocess_case()
    np.savez_compressed(os.path.join(OUT, "ref_preprocess.npz"), **rec)
    print("ref_preprocess", rec["phi1_shrink"].shape, rec["phi1_full"].shape)
    for loss, name in ((0, "ref_linreg_xent"), (1, "ref_linreg_squared")):
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            rec = run_linreg_case(loss)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "weight", rec["weight"].shape, "loc", int(rec["loc"]), "next", int(rec["next_index"]))
    for name, cfg in CASES.items():
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):            # the reference prints a lot
            rec = run_case(*cfg)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "dims", rec["dims"].tolist(), "cost", float(rec["cost_final"]))


if __name__ == "__main__":
    # the lazy graph is evaluated recursively and the reference chains thousands of TensorArray writes
    import threading
    sys.setrecursionlimit(1000000)
    threading.stack_size(1 << 30)
    t = threading.Thread(target=main)
    t.start()
    t.join()
