"""Device-side state behind the Python mirror: packs the reference's list-of-arrays weights into the
slot layout of include/trmps_b200.h, owns the torch CUDA buffers, and enqueues the C-ABI calls.

torch is plumbing here (device memory, streams, host<->device copies); all arithmetic on the path
happens in libtrmps_b200.so.
"""
import ctypes as C

import numpy as np
import torch

import trmps_native as nat


def align8(x):
    return (int(x) + 7) // 8 * 8


def bond_stride(nodes, loc, L, max_size=0):
    """Padded bond stride Ds: multiple of 8 covering every bond dimension, 2L and max_size."""
    mx = max(2 * L, int(max_size))
    for i, w in enumerate(nodes):
        mx = max(mx, w.shape[-1], w.shape[-2])
    return align8(mx)


def pack_nodes(nodes, loc, L, Ds):
    """list of (d,Dl,Dr) arrays (+ one (L,d,Dl,Dr) at loc) -> (slots (N,2,Ds,Ds), special (L,2,Ds,Ds), dims (N+1,))"""
    N = len(nodes)
    slots = np.zeros((N, 2, Ds, Ds), dtype=np.float32)
    special = np.zeros((L, 2, Ds, Ds), dtype=np.float32)
    dims = np.zeros((N + 1,), dtype=np.int32)
    for i, w in enumerate(nodes):
        w = np.asarray(w, dtype=np.float32)
        if i == loc:
            if w.ndim != 4 or w.shape[0] != L or w.shape[1] != 2:
                raise ValueError("special node %d must have shape (L=%d, 2, Dl, Dr); got %r" % (i, L, w.shape))
            special[:, :, :w.shape[2], :w.shape[3]] = w
        else:
            if w.ndim != 3 or w.shape[0] != 2:
                raise ValueError("node %d must have shape (2, Dl, Dr); got %r (only d_feature=2 is built)" % (i, w.shape))
            slots[i, :, :w.shape[1], :w.shape[2]] = w
        dl, dr = w.shape[-2], w.shape[-1]
        if i > 0 and dims[i] != dl:
            raise ValueError("bond mismatch between node %d and %d: %d vs %d" % (i - 1, i, dims[i], dl))
        dims[i] = dl
        dims[i + 1] = dr
    return slots, special, dims


def unpack_nodes(slots, special, dims, loc):
    nodes = []
    for i in range(slots.shape[0]):
        dl, dr = int(dims[i]), int(dims[i + 1])
        if i == loc:
            nodes.append(np.ascontiguousarray(special[:, :, :dl, :dr]))
        else:
            nodes.append(np.ascontiguousarray(slots[i, :, :dl, :dr]))
    return nodes


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_device_f32(x, device, pinned_stage=None):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    arr = np.ascontiguousarray(x, dtype=np.float32)
    if pinned_stage is not None and pinned_stage.numel() >= arr.size:
        host = pinned_stage[:arr.size].view(arr.shape)
        host.copy_(torch.from_numpy(arr))
        return host.to(device, non_blocking=True)
    return torch.from_numpy(arr).to(device)


def feature_layout(feature, N, d):
    """Decide whether `feature` is phi (N,B,d) or raw pixels (N,B)."""
    shp = tuple(feature.shape)
    if len(shp) == 3 and shp[0] == N and shp[2] == d:
        return "phi", shp[1]
    if len(shp) == 2 and shp[0] == N:
        return "pixels", shp[1]
    raise ValueError("feature must be (N=%d, B, d=%d) phi or (N, B) pixels; got %r" % (N, d, shp))


def preprocess_batch(images, rows, cols, shrink, start, batch, out_map=nat.FM_PRECOMPUTED, out=None, u8_divide=False):
    """Fused datasource preprocessing + batch fetch on the device (MNISTpreprocessing.py:44-55, preprocessing.py:165-208).
    images: CUDA tensor (total, rows*cols), float32 in [0,1] or uint8 (scaled like the TF MNIST loader, or divided by 255
    like FashionMNISTpreprocessing.py:60 when u8_divide).  shrink: False / True (2x2 average, MNIST) or a nat.POOL_* code
    (POOL_MAX = the Fashion-MNIST datasource).  Returns (N, batch) pooled pixels when out_map == FM_PRECOMPUTED, else
    (N, batch, 2) phi of that map; the batch is images[(start + t) % total]."""
    pool = int(shrink)
    if not (isinstance(images, torch.Tensor) and images.is_cuda and images.is_contiguous()):
        raise ValueError("preprocess_batch: images must be a contiguous CUDA tensor")
    if images.dtype not in (torch.float32, torch.uint8):
        raise ValueError("preprocess_batch: images must be float32 or uint8")
    total = images.shape[0]
    if images.numel() != total * rows * cols:
        raise ValueError("preprocess_batch: images must be (total, rows*cols)")
    n_sites = (rows // 2) * (cols // 2) if pool else rows * cols
    shape = (n_sites, batch) if out_map == nat.FM_PRECOMPUTED else (n_sites, batch, 2)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=images.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("preprocess_batch: out must be a contiguous float32 tensor of shape %r" % (shape,))
    dt = nat.DT_F32 if images.dtype == torch.float32 else (nat.DT_U8_DIV if u8_divide else nat.DT_U8)
    if batch == 0:
        return out
    nat.check(nat.load().trmps_preprocess_batch(images.data_ptr(), dt, total, rows, cols, pool, int(start),
                                                int(batch), int(out_map), out.data_ptr(), _stream_ptr()),
              "trmps_preprocess_batch")
    return out


def batch_gather(data, labels, start, batch, feat_out=None, labels_out=None):
    """MPSDatasource.next_training_data_batch (preprocessing.py:165-208) on a device-resident dataset:
    data (total, N, c) or (total, N) float32 CUDA, labels (total, L) float32 CUDA or None.
    Returns feat (N, batch[, c]) and labels (batch, L) for rows (start + t) % total."""
    if not (isinstance(data, torch.Tensor) and data.is_cuda and data.is_contiguous() and data.dtype == torch.float32):
        raise ValueError("batch_gather: data must be a contiguous float32 CUDA tensor")
    total, N = data.shape[0], data.shape[1]
    c = data.shape[2] if data.dim() == 3 else 1
    shape = (N, batch, c) if data.dim() == 3 else (N, batch)
    if feat_out is None:
        feat_out = torch.empty(shape, dtype=torch.float32, device=data.device)
    elif tuple(feat_out.shape) != shape or not feat_out.is_contiguous():
        raise ValueError("batch_gather: feat_out must be contiguous with shape %r" % (shape,))
    L, lab_ptr, lab_out_ptr = 0, None, None
    if labels is not None:
        if not (labels.is_cuda and labels.is_contiguous() and labels.dtype == torch.float32 and labels.shape[0] == total):
            raise ValueError("batch_gather: labels must be a contiguous float32 CUDA tensor (total, L)")
        L = labels.shape[1]
        if labels_out is None:
            labels_out = torch.empty((batch, L), dtype=torch.float32, device=data.device)
        lab_ptr, lab_out_ptr = labels.data_ptr(), labels_out.data_ptr()
    if batch == 0:
        return feat_out, labels_out
    nat.check(nat.load().trmps_batch_gather(data.data_ptr(), lab_ptr, total, N, c, L, int(start), int(batch),
                                            feat_out.data_ptr(), lab_out_ptr, _stream_ptr()), "trmps_batch_gather")
    return feat_out, labels_out


def linreg_pretrain(data, labels, start, iterations, learning_rate, loss_kind=nat.LOSS_SOFTMAX_XENT, batch=100):
    """MPS._lin_reg (mps.py:331-399) on a device-resident dataset: data (total, N, 2) phi, labels (total, L), float32 CUDA.
    Returns weight (N, L) and bias (L) after `iterations` minibatch steps starting at dataset row `start`."""
    for name, x in (("data", data), ("labels", labels)):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.is_contiguous() and x.dtype == torch.float32):
            raise ValueError("linreg_pretrain: %s must be a contiguous float32 CUDA tensor" % name)
    if data.dim() != 3 or data.shape[2] != 2 or labels.dim() != 2 or labels.shape[0] != data.shape[0]:
        raise ValueError("linreg_pretrain: data must be (total, N, 2) and labels (total, L)")
    total, N, L = data.shape[0], data.shape[1], labels.shape[1]
    weight = torch.empty((N, L), dtype=torch.float32, device=data.device)
    bias = torch.empty((L,), dtype=torch.float32, device=data.device)
    nat.check(nat.load().trmps_linreg_pretrain(data.data_ptr(), labels.data_ptr(), total, N, L, int(start) % total, int(iterations),
                                               int(batch), float(learning_rate), int(loss_kind), weight.data_ptr(),
                                               bias.data_ptr(), _stream_ptr()), "trmps_linreg_pretrain")
    return weight, bias


class DeviceWeights(object):
    """MPS weights resident on the GPU in slot layout."""

    def __init__(self, nodes, loc, L, Ds, device):
        slots, special, dims = pack_nodes(nodes, loc, L, Ds)
        self.N, self.L, self.Ds, self.loc = len(nodes), L, Ds, loc
        self.device = device
        self.slots = torch.from_numpy(slots).to(device)
        self.special = torch.from_numpy(special).to(device)
        self.dims = torch.from_numpy(dims).to(device)

    def to_host(self, loc=None):
        loc = self.loc if loc is None else loc
        return unpack_nodes(self.slots.cpu().numpy(), self.special.cpu().numpy(), self.dims.cpu().numpy(), loc)


def predict(weights, feature, feature_map, round_output=False, out=None):
    """MPS.predict on the device (mps.py:521-560).  feature: CUDA float32 tensor, (N,B,2) phi or (N,B) pixels."""
    lib = nat.load()
    N, L, Ds = weights.N, weights.L, weights.Ds
    B = feature.shape[1]
    dev = weights.device
    work = torch.empty(int(lib.trmps_predict_work_floats(B, L, Ds, N)), dtype=torch.float32, device=dev)
    logits = out if out is not None else torch.empty((B, L), dtype=torch.float32, device=dev)
    p = nat.PredictDesc(N=N, d=2, L=L, Ds=Ds, B=B, loc=weights.loc, feature_map=feature_map,
                        round_output=int(bool(round_output)), reserved=0,
                        feat=feature.data_ptr(), nodes=weights.slots.data_ptr(), special=weights.special.data_ptr(),
                        dims=weights.dims.data_ptr(), logits=logits.data_ptr(), work=work.data_ptr())
    nat.check(lib.trmps_predict(C.byref(p), _stream_ptr()), "trmps_predict")
    return logits


class SweepState(object):
    """One rank's buffers for the two-site sweep: what BaseOptimizer.__init__ and
    MPSOptimizer._setup_optimization build as placeholders/TensorArrays (baseoptimizer.py:37-47, optimizer.py:49-87)."""

    def __init__(self, N, B, L, Ds, feature_map, loss_kind, device, B_total=None, comm=None):
        nat.require_gpu(device.index or 0)
        self.lib = nat.load()
        self.N, self.B, self.L, self.Ds = int(N), int(B), int(L), int(Ds)
        self.B_total = int(B_total if B_total is not None else B)
        self.feature_map, self.loss_kind = int(feature_map), int(loss_kind)
        self.device = device
        self.comm = comm
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        self.layout = nat.sweep_layout(self.N, self.B, self.L, self.Ds, sms)
        f32 = dict(dtype=torch.float32, device=device)
        self.slots = torch.zeros((self.N, 2, self.Ds, self.Ds), **f32)
        self.special = torch.zeros((self.L, 2, self.Ds, self.Ds), **f32)
        self.dims = torch.zeros((self.N + 1,), dtype=torch.int32, device=device)
        self.envL = torch.zeros((self.N, self.B, self.Ds), **f32)
        self.envR = torch.zeros((self.N, self.B, self.Ds), **f32)
        self.work = torch.zeros((self.layout.total,), **f32)
        self.iwork = torch.zeros((nat.I_COUNT,), dtype=torch.int32, device=device)
        fdim = 2 if self.feature_map == nat.FM_PRECOMPUTED else 1
        self.feat = torch.zeros((self.N, self.B, fdim) if fdim == 2 else (self.N, self.B), **f32)
        self.labels = torch.zeros((self.B, self.L), **f32)
        self.loc = None
        self._desc = nat.Sweep(N=self.N, d=2, L=self.L, Ds=self.Ds, B=self.B, B_total=self.B_total,
                               feature_map=self.feature_map, loss_kind=self.loss_kind,
                               feat=self.feat.data_ptr(), labels=self.labels.data_ptr(), nodes=self.slots.data_ptr(),
                               special=self.special.data_ptr(), dims=self.dims.data_ptr(), envL=self.envL.data_ptr(),
                               envR=self.envR.data_ptr(), work=self.work.data_ptr(), iwork=self.iwork.data_ptr(),
                               comm=comm, work_floats=int(self.work.numel()))

    # ---- views into the library-defined work buffer -------------------------------------------
    @property
    def R(self):
        return 2 * self.Ds

    @property
    def Cc(self):
        return 2 * self.L * self.Ds

    def _view(self, off, *shape):
        n = int(np.prod(shape))
        return self.work[off:off + n].view(*shape)

    def bond_matrix(self):
        return self._view(self.layout.bondM, self.R, self.Cc)

    def grad_matrix(self):
        return self._view(self.layout.gradM, self.R, self.Cc)

    def logits0(self):
        return self._view(self.layout.f0, self.B, self.L)

    def logits_delta(self):
        return self._view(self.layout.fd, self.B, self.L)

    def scalars(self):
        return self._view(self.layout.scal, nat.S_COUNT)

    def singular_values(self):
        return self._view(self.layout.sv, self.R)

    def bond_tensor(self, M=None):
        """bond matrix (R, Cc) -> reference layout (L, d, d, Dl_pad, Dr_pad)."""
        M = self.bond_matrix() if M is None else M
        t = M.view(2, self.Ds, self.L, 2, self.Ds)          # m, i, l, n, k
        return t.permute(2, 0, 3, 1, 4).contiguous()        # l, m, n, i, k

    # ---- data in --------------------------------------------------------------------------------
    def set_nodes(self, nodes, loc):
        slots, special, dims = pack_nodes(nodes, loc, self.L, self.Ds)
        self.slots.copy_(torch.from_numpy(slots), non_blocking=True)
        self.special.copy_(torch.from_numpy(special), non_blocking=True)
        self.dims.copy_(torch.from_numpy(dims), non_blocking=True)
        self.loc = int(loc)

    def set_batch(self, feature, labels):
        f = feature if isinstance(feature, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(feature, dtype=np.float32))
        y = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels, dtype=np.float32))
        if tuple(f.shape) != tuple(self.feat.shape):
            raise ValueError("feature shape %r does not match the sweep state %r" % (tuple(f.shape), tuple(self.feat.shape)))
        self.feat.copy_(f, non_blocking=True)
        self.labels.copy_(y.to(torch.float32), non_blocking=True)

    def get_nodes(self):
        return unpack_nodes(self.slots.cpu().numpy(), self.special.cpu().numpy(), self.dims.cpu().numpy(), self.loc)

    def weights_view(self):
        w = DeviceWeights.__new__(DeviceWeights)
        w.N, w.L, w.Ds, w.loc, w.device = self.N, self.L, self.Ds, self.loc, self.device
        w.slots, w.special, w.dims = self.slots, self.special, self.dims
        return w

    # ---- C-ABI calls ------------------------------------------------------------------------------
    def _d(self):
        return C.byref(self._desc)

    def init_env(self, loc=None):
        loc = self.loc if loc is None else loc
        nat.check(self.lib.trmps_sweep_init_env(self._d(), loc, _stream_ptr()), "trmps_sweep_init_env")

    def bond_step(self, site, direction, opt):
        nat.check(self.lib.trmps_bond_step(self._d(), site, direction, C.byref(opt), _stream_ptr()), "trmps_bond_step")
        self.loc = site + direction

    def sweep(self, frm, to, direction, opt):
        nat.check(self.lib.trmps_sweep_run(self._d(), frm, to, direction, C.byref(opt), _stream_ptr()), "trmps_sweep_run")
        self.loc = to

    def train_step(self, opt, loc=None):
        loc = self.loc if loc is None else loc
        nat.check(self.lib.trmps_train_step(self._d(), loc, C.byref(opt), _stream_ptr()), "trmps_train_step")
        self.loc = loc

    def bond_merge(self, site, direction):
        nat.check(self.lib.trmps_bond_merge(self._d(), site, direction, _stream_ptr()), "trmps_bond_merge")

    def bond_forward(self, site, direction, use_delta=0):
        nat.check(self.lib.trmps_bond_forward(self._d(), site, direction, use_delta, _stream_ptr()), "trmps_bond_forward")

    def bond_gradient(self, site, direction, opt):
        nat.check(self.lib.trmps_bond_gradient(self._d(), site, direction, C.byref(opt), _stream_ptr()), "trmps_bond_gradient")

    def bond_split(self, site, direction, opt):
        nat.check(self.lib.trmps_bond_split(self._d(), site, direction, C.byref(opt), _stream_ptr()), "trmps_bond_split")
        self.loc = site + direction


def make_opt(lr, max_size, reg=1e-3, armijo_coeff=1e-4, min_singular_value=1e-4, updates_per_step=1,
             use_hessian=False, min_size=3, max_svd_sweeps=0, single_site=False, split_impl=0):
    return nat.Opt(lr=float(lr), reg=float(reg), armijo_coeff=float(armijo_coeff),
                   min_singular_value=float(min_singular_value), max_size=int(max_size), min_size=int(min_size),
                   updates_per_step=int(updates_per_step), use_hessian=int(bool(use_hessian)),
                   max_svd_sweeps=int(max_svd_sweeps), single_site=int(bool(single_site)),
                   split_impl=int(split_impl), reserved=0)
