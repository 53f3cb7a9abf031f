"""MPS model object -- same construction / predict surface as the reference's trmps/mps/mps.py,
with the TensorFlow graph replaced by calls into libtrmps_b200.so (sm_100a kernels).

Differences a reference user will notice (there is no TensorFlow here):
  * predict() is eager: it returns the logits (numpy in -> numpy out, torch in -> torch CUDA out)
    instead of a lazy tf.Tensor, so no Session / feed_dict is needed.
  * create_feed_dict(weights) installs `weights` as the model's current weights and returns a
    plain {site: array} dict.
  * d_feature must be 2 and d_output <= 16 (every reference configuration on the hot path).
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np
import torch

import trmps_native as nat
import trmps_device as dev


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class MPS(object):
    """Matrix product state classifier (Stoudenmire & Schwab).  Reference: trmps/mps/mps.py:4-590.

    Attributes kept from the reference: input_size, d_matrix, d_feature, d_output, _special_node_loc,
    round, weight, bias, nodes, nodes_list, start_node, end_node.
    """

    loss_kind = nat.LOSS_SOFTMAX_XENT

    def __init__(self, d_feature, d_output, input_size, special_node_loc=None, round=False):
        # mps.py:162-187
        self.input_size = int(input_size)
        self.d_matrix = int(d_output) * 2
        self.d_feature = int(d_feature)
        self.d_output = int(d_output)
        self._special_node_loc = special_node_loc
        self.round = round
        self.nodes = None
        self.nodes_list = None
        self.start_node = None
        self.end_node = None
        self.weight = None
        self.bias = None
        self._device_weights = None
        if self.d_feature != 2:
            raise ValueError("this build supports d_feature == 2 only (got %d)" % self.d_feature)
        if not 1 <= self.d_output <= 16:
            raise ValueError("this build supports 1 <= d_output <= 16 (got %d)" % self.d_output)

    # ------------------------------------------------------------------ preparation (mps.py:189-219)
    def prepare(self, data_source=None, iterations=1000, learning_rate=0.05, _weights=None):
        if data_source is not None:
            self._lin_reg(data_source, iterations, learning_rate)
            self._setup_nodes()
        elif _weights is not None:
            self._prepare_from_previous(_weights)
        else:
            self._random_prepare()
            self._setup_nodes()

    def _random_prepare(self):
        # mps.py:211-219: uniform linear model, zero bias, special node in the middle
        n_in = self.input_size * (self.d_feature - 1)
        self.weight = np.full((self.input_size, self.d_feature - 1, self.d_output), 1.0 / n_in)
        self.bias = np.zeros((self.d_output,))
        if self._special_node_loc is None:
            self._special_node_loc = self.input_size // 2

    def _lin_reg(self, data_source, iterations, learning_rate):
        """Softmax-regression pre-training on phi[..., 1:] with 100-image minibatches (mps.py:331-399): plain gradient
        descent on the mean cross-entropy (squared error for sqMPS), zero init.  All iterations run inside one kernel
        launch (trmps_linreg_pretrain) on the training set kept on the GPU; the datasource's read position ends where the
        reference's `iterations` calls of next_training_data_batch(100) leave it."""
        if self.d_feature != 2:
            raise ValueError("this build has d_feature = 2 (SURVEY.md section 8)")
        nat.require_gpu(torch.cuda.current_device())
        resident = getattr(data_source, "_device_training", None)
        if resident is None:
            data, labels = data_source._training_data
            tdev = torch.device('cuda', torch.cuda.current_device())
            resident = (torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(tdev),
                        torch.from_numpy(np.ascontiguousarray(labels, dtype=np.float32)).to(tdev))
        total = resident[0].shape[0]
        start = data_source.current_index % total
        W, b = dev.linreg_pretrain(resident[0], resident[1], start, int(iterations), learning_rate, self.loss_kind, batch=100)
        if iterations > 0:
            data_source.current_index = (start + 100 * int(iterations)) % total or total
        # (n_in, L) -> (N, d-1, L), matching the reshapes at mps.py:358-361
        self.weight = W.t().reshape(self.d_output, self.input_size, self.d_feature - 1).permute(1, 2, 0).cpu().numpy()
        self.bias = b.cpu().numpy()
        if self._special_node_loc is None:
            site = int(np.unravel_index(self.weight.argmax(), self.weight.shape)[0])
            site = min(max(site, 1), self.input_size - 2)      # mps.py:392-399
            self._special_node_loc = site

    def _setup_nodes(self):
        # mps.py:401-423
        self.start_node = self._make_start_node()
        self.end_node = self._make_end_node()
        nodes = []
        for i in range(self.input_size):
            nodes.append(self._make_special_node(i) if i == self._special_node_loc else self._make_middle_node(i))
        self._install(nodes)

    def _prepare_from_previous(self, weights):
        # mps.py:425-443.  Like the reference this never sets _special_node_loc: pass it to the constructor.
        if self._special_node_loc is None:
            raise ValueError("prepare(_weights=...) needs special_node_loc given to the constructor (mps.py:425-443)")
        self.start_node = self._make_start_node()
        self.end_node = self._make_end_node()
        self._install([np.asarray(w, dtype=np.float32) for w in weights])

    def _install(self, nodes):
        if len(nodes) != self.input_size:
            raise ValueError("expected %d node tensors, got %d" % (self.input_size, len(nodes)))
        self.nodes_list = list(nodes)
        self.nodes = self.nodes_list
        self._device_weights = None

    def _make_start_node(self):
        v = np.zeros((1, self.d_matrix), dtype=np.float32)      # mps.py:446-454
        v[0, self.d_output:] = 1
        return v

    def _make_end_node(self):
        v = np.zeros((1, self.d_matrix), dtype=np.float32)      # mps.py:456-466
        v[0, :self.d_output] = 1
        return v

    def _make_special_node(self, index):
        # mps.py:468-482: one (d, 2L, 2L) block per label
        L, D = self.d_output, self.d_matrix
        node = np.zeros((L, self.d_feature, D, D), dtype=np.float32)
        lab = np.arange(L)
        node[lab, 0, lab, lab] = 1
        node[lab, 0, lab + L, lab + L] = 1
        node[lab, 0, lab + L, lab] = self.bias
        node[lab, 1:, lab + L, lab] = self.weight[index].T
        return node

    def _make_middle_node(self, index):
        # mps.py:485-497
        L, D = self.d_output, self.d_matrix
        node = np.zeros((self.d_feature, D, D), dtype=np.float32)
        node[0] = np.eye(D, dtype=np.float32)
        lab = np.arange(L)
        for k in range(1, self.d_feature):
            node[k, lab + L, lab] = self.weight[index, k - 1]
        return node

    # ------------------------------------------------------------------ inference (mps.py:521-560)
    def _weights_on_device(self):
        if self.nodes_list is None:
            raise RuntimeError("call prepare() before using the MPS (mps.py:189)")
        if self._device_weights is None:
            nat.require_gpu(torch.cuda.current_device())
            Ds = dev.bond_stride(self.nodes_list, self._special_node_loc, self.d_output)
            self._device_weights = dev.DeviceWeights(self.nodes_list, self._special_node_loc, self.d_output, Ds,
                                                     torch.device('cuda', torch.cuda.current_device()))
        return self._device_weights

    def predict(self, feature, feature_map=nat.FM_PRECOMPUTED):
        """feature: (input_size, batch, d_feature) phi (reference contract), or (input_size, batch) raw
        pixels together with feature_map != FM_PRECOMPUTED (fused feature map).  Returns (batch, d_output)."""
        w = self._weights_on_device()
        kind, _ = dev.feature_layout(feature, self.input_size, self.d_feature)
        if kind == "phi":
            feature_map = nat.FM_PRECOMPUTED
        elif feature_map == nat.FM_PRECOMPUTED:
            raise ValueError("raw pixels need a feature_map (FM_ONE_SQRT3, FM_ONE_X or FM_COS_SIN)")
        was_numpy = not isinstance(feature, torch.Tensor)
        f = dev._as_device_f32(feature, w.device)
        out = dev.predict(w, f, feature_map, self.round)
        return out.cpu().numpy() if was_numpy else out

    def create_feed_dict(self, weights):
        # mps.py:562-590 (documented break: installs the weights, returns a plain dict)
        if weights is not None:
            self._install([np.asarray(w, dtype=np.float32) for w in weights])
            return {i: w for i, w in enumerate(self.nodes_list)}
        return {}

    # ------------------------------------------------------------------ metrics (mps.py:267-325)
    def cost(self, f, labels):
        f, y = _to_numpy(f).astype(np.float32), _to_numpy(labels).astype(np.float32)
        z = f - f.max(axis=1, keepdims=True)
        logp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
        return np.float32(np.mean(-(y * logp).sum(axis=1)))

    def accuracy(self, f, labels):
        f, y = _to_numpy(f), _to_numpy(labels)
        return np.float32(np.mean(f.argmax(axis=1) == y.argmax(axis=1)))

    def confusion_matrix(self, f, labels):
        f, y = _to_numpy(f), _to_numpy(labels)
        cm = np.zeros((self.d_output, self.d_output), dtype=np.int64)
        np.add.at(cm, (y.argmax(axis=1), f.argmax(axis=1)), 1)
        return cm

    def f1score(self, f, labels, _confusion_matrix=None):
        cm = self.confusion_matrix(f, labels) if _confusion_matrix is None else np.asarray(_confusion_matrix)
        with np.errstate(divide='ignore', invalid='ignore'):
            per_class = 2.0 * np.diag(cm) / (cm.sum(axis=0) + cm.sum(axis=1))
        return float(np.sum(per_class) / self.d_output)

    def test(self, test_feature, test_label):
        # mps.py:222-246
        f = self.predict(test_feature)
        print("\n\n\n\n Accuracy is:" + str(self.accuracy(f, test_label)))
        print("\n\n\n\n Cost is:" + str(self.cost(f, test_label)))
        print("\n\n\n\n" + str(self.confusion_matrix(f, test_label)))
        print("\n\n\n\n Sample prediction: " + str(_to_numpy(f)[0]))

    # ------------------------------------------------------------------ checkpoint (mps.py:80-160)
    def save(self, weights=None, path="MPSconfig", verbose=False):
        """Byte-compatible with the reference file format (Notes for developement.md:44-73)."""
        blob = bytearray()
        blob += bytes([1 if weights is not None else 0])
        for v in (self.input_size, self._special_node_loc, self.d_feature, self.d_output):
            blob += int(v).to_bytes(2, byteorder='big')
        if weights is None:
            print("contains_weights False")
        else:
            n_params = 0
            for i, w in enumerate(weights):
                blob += int(w.shape[-1]).to_bytes(2, byteorder='big')
                n_params += int(np.prod(w.shape))
            if verbose:
                print("The number of parameters are:", n_params)
            for w in weights:
                blob += np.ascontiguousarray(w, dtype=np.float32).tobytes(order='C')
        with open(path, "w+b") as fh:
            fh.write(bytes(blob))

    @classmethod
    def from_file(cls, path="MPSconfig"):
        with open(path, "rb") as fh:
            raw = fh.read()
        has_weights = bool(raw[0])
        n, loc, d, L = (int.from_bytes(raw[1 + 2 * k:3 + 2 * k], byteorder='big') for k in range(4))
        self = cls(d, L, n, special_node_loc=loc)
        if not has_weights:
            print("Please remember to prepare the MPS")
            return self
        print("No need to prepare: previous weights will be used")
        right = [int.from_bytes(raw[9 + 2 * i:11 + 2 * i], byteorder='big') for i in range(n - 1)] + [self.d_matrix]
        left = [self.d_matrix] + right[:-1]
        pos = 9 + 2 * n
        weights = []
        for i in range(n):
            shape = (d, left[i], right[i]) if i != loc else (L, d, left[i], right[i])
            cnt = int(np.prod(shape))
            weights.append(np.frombuffer(raw, dtype=np.float32, count=cnt, offset=pos).reshape(shape).copy())
            pos += 4 * cnt
        self.prepare(_weights=weights)
        return self
