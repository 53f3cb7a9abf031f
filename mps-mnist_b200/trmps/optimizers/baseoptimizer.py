"""Sweep driver with the reference's interface (trmps/optimizers/baseoptimizer.py:15-244).

The reference builds one static TensorFlow graph for a whole training step and runs it with one
sess.run per step.  Here a step is one trmps_train_step call that enqueues every kernel of the three
half-sweeps on the current CUDA stream; the host never waits inside the step.
"""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np
import torch

import trmps_native as nat
import trmps_device as dev
from optimizers.parameterObjects import MPSOptimizerParameters, MPSTrainingParameters


class BaseOptimizer(object):

    def __init__(self, MPSNetwork, max_size, optional_parameters=MPSOptimizerParameters()):
        # baseoptimizer.py:17-56
        self.parameters = optional_parameters
        self.path = self.parameters.path
        self.MPS = MPSNetwork
        self.use_hessian = self.parameters.use_hessian
        if self.use_hessian and self.MPS.loss_kind == nat.LOSS_SQUARED:
            print("Using the hessian for squared distance error is not yet supported, sorry!")
            self.use_hessian = False
        self._armijo_iterations = self.parameters.armijo_iterations
        self.rate_of_change = None
        self.reg = self.parameters.reg
        self.lr_reg = self.parameters.lr_reg
        self.max_size = max_size
        self.armijo_coeff = self.parameters.armijo_coeff
        self.cutoff = self.parameters.cutoff
        self.min_singular_value = self.parameters.min_singular_value
        self.verbosity = self.parameters.verbosity
        self.updates_per_step = self.parameters.updates_per_step
        self.feature_map = nat.FM_PRECOMPUTED
        self.test = None
        self.feed_dict = None
        self._state = None
        self._feature = None
        self._label = None
        self._comm = None
        self._world = (0, 1)      # (rank, world) of the data-parallel job; see attach_process_group
        if self.MPS.nodes_list is None:
            raise RuntimeError("prepare() the MPS before constructing the optimizer")

    # ------------------------------------------------------------------ knobs -> native struct
    def _opt(self, rate_of_change):
        return dev.make_opt(lr=rate_of_change, max_size=self.max_size, reg=self.reg, armijo_coeff=self.armijo_coeff,
                            min_singular_value=self.min_singular_value, updates_per_step=self.updates_per_step,
                            use_hessian=self.use_hessian)

    def _ensure_state(self, batch, feature_kind, weights):
        L, N = self.MPS.d_output, self.MPS.input_size
        Ds = dev.bond_stride(weights, self.MPS._special_node_loc, L, self.max_size)
        fm = nat.FM_PRECOMPUTED if feature_kind == "phi" else self.feature_map
        st = self._state
        rank, world = self._world
        if (st is None or st.B != batch or st.Ds != Ds or st.feature_map != fm or st.N != N):
            self._state = None
            torch.cuda.empty_cache()
            device = torch.device('cuda', torch.cuda.current_device())
            self._state = dev.SweepState(N, batch, L, Ds, fm, self.MPS.loss_kind, device,
                                         B_total=batch * world, comm=self._comm)
        return self._state

    # ------------------------------------------------------------------ training (baseoptimizer.py:58-143)
    def train(self, data_source, batch_size, n_step, optional_parameters=MPSTrainingParameters()):
        """One n_step = right half-sweep loc->N-1, left sweep N-1->0, right half-sweep 0->loc
        (baseoptimizer.py:213-238).  After each step the updated weights are in self.test (list of
        numpy arrays) and saved to self.path, as in the reference."""
        self.test = optional_parameters.initial_weights
        if self.test is None:
            self.test = list(self.MPS.nodes_list)
        initial_lr = optional_parameters.rate_of_change
        test_feature, test_label = data_source.test_data
        # the test set goes to the device once, and the test metrics are evaluated with the weights the sweep left on the device
        # (the reference feeds the NumPy test set and every weight tensor to the graph each step)
        test_kind, _ = dev.feature_layout(test_feature, self.MPS.input_size, self.MPS.d_feature)
        test_fm = nat.FM_PRECOMPUTED if test_kind == "phi" else self.feature_map
        test_dev = None
        for i in range(n_step):
            start = time.time()
            rate_of_change = initial_lr / np.sqrt(1 - self.lr_reg ** (i + 1))      # baseoptimizer.py:109
            print(rate_of_change)
            batch_feature, batch_label = data_source.next_training_data_batch(batch_size)
            self.bind_batch(batch_feature, batch_label, weights=self.test)
            self.rate_of_change = rate_of_change
            train_acc = self.train_step()
            self.test = self._state.get_nodes()                  # NumPy copies: the reference's self.test, and what save() writes
            self.MPS.create_feed_dict(self.test)
            self._state_holds = self.test                        # the device already holds these weights: bind_batch skips the upload
            train_f = dev.predict(self._state.weights_view(), self._state.feat, self._state.feature_map, self.MPS.round)
            train_c = self.MPS.cost(train_f, batch_label)
            if test_dev is None:
                test_dev = dev._as_device_f32(test_feature, self._state.device)
            test_prediction = dev.predict(self._state.weights_view(), test_dev, test_fm, self.MPS.round).cpu().numpy()
            test_c = self.MPS.cost(test_prediction, test_label)
            test_acc = self.MPS.accuracy(test_prediction, test_label)
            test_conf = self.MPS.confusion_matrix(test_prediction, test_label)
            test_f1score = self.MPS.f1score(test_prediction, test_label, test_conf)
            self.MPS.save(self.test, path=self.path, verbose=optional_parameters.verbose_save)
            end = time.time()
            print(time.strftime("%Y-%m-%d %H:%M:%S"))
            print('step {}, training cost {}, accuracy {}. Took {} s'.format(i, train_c, train_acc, end - start))
            print('step {}, testing cost {}, accuracy {}'.format(i, test_c, test_acc))
            print('f1 score: ', test_f1score)
            print('sample prediction: ', np.asarray(test_prediction)[0])
            print('sample truth: ', np.asarray(test_label)[0])
            print('confusion matrix: \n' + str(test_conf))

    def bind_batch(self, feature, label, weights=None):
        """Feed one batch: the placeholders self._feature / self._label of the reference (baseoptimizer.py:44-45)."""
        weights = self.MPS.nodes_list if weights is None else weights
        kind, batch = dev.feature_layout(feature, self.MPS.input_size, self.MPS.d_feature)
        prev = self._state
        st = self._ensure_state(batch, kind, weights)
        if not (st is prev and getattr(self, "_state_holds", None) is weights and weights is not None
                and st.loc == self.MPS._special_node_loc):
            st.set_nodes(weights, self.MPS._special_node_loc)
        self._state_holds = None
        st.set_batch(feature, label)
        self._feature, self._label = st.feat, st.labels

    def train_step(self):
        """One training step on the bound batch (baseoptimizer.py:203-244); returns the training accuracy
        of the updated MPS, like the reference's train_step graph output."""
        if self._state is None:
            raise RuntimeError("bind_batch(feature, label) before train_step()")
        if self.rate_of_change is None:
            raise RuntimeError("set optimizer.rate_of_change before train_step()")
        st = self._state
        st.train_step(self._opt(self.rate_of_change), loc=self.MPS._special_node_loc)
        f = dev.predict(st.weights_view(), st.feat, st.feature_map, self.MPS.round)
        return self.MPS.accuracy(f, st.labels)

    def updated_nodes(self):
        """Weights after the last train_step (the TensorArray self.updated_nodes of the reference)."""
        return self._state.get_nodes()

    def _test_step(self, feature, label):
        # baseoptimizer.py:145-158
        f = self.MPS.predict(feature)
        cost = self.MPS.cost(f, label)
        accuracy = self.MPS.accuracy(f, label)
        confusion = self.MPS.confusion_matrix(f, label)
        f1 = self.MPS.f1score(f, label, confusion)
        return cost, accuracy, confusion, f1, f

    # ------------------------------------------------------------------ data-parallel (new)
    def attach_process_group(self, group=None):
        """Shard the batch over the ranks of an initialised torch.distributed job: every rank keeps a
        replicated MPS and its own images/environments; only the per-bond gradient (and the cost sums)
        are all-reduced with NCCL inside the bond step."""
        import ctypes as C
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self._world = (rank, world)
        if world == 1:
            return
        lib = nat.load()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            nat.check(lib.trmps_comm_unique_id(buf), "trmps_comm_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            uid = uid.cuda()
        dist.broadcast(uid, src=0, group=group)
        raw = bytes(uid.cpu().numpy().tobytes())
        comm = C.c_void_p()
        nat.check(lib.trmps_comm_create(raw, rank, world, torch.cuda.current_device(), C.byref(comm)), "trmps_comm_create")
        self._comm = comm
        self._state = None
