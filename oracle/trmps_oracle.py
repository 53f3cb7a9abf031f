"""CPU oracle for the TrMPS hot path (MPS forward contraction + two-site DMRG sweep).

TEST INFRASTRUCTURE ONLY.  This file restates, in NumPy, the arithmetic that the
reference (TrMPS/MPS-MNIST, a TensorFlow-1.2 graph) performs on the path named in
BASELINE.json.  It is the checker for the CUDA path, never the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.

Parity status: PINNED against the reference's own code.  The reference holds no golden
vectors for this path (its single unit test is stale) and
TensorFlow 1.2 cannot be installed here, so the reference's hot-path modules are executed
UNMODIFIED under a NumPy stand-in for the ~45 TF symbols they use (``oracle/tf1_shim.py``,
driver ``oracle/make_golden_from_reference.py``); their outputs are committed as
``tests/golden/ref_*.npz`` and this restatement reproduces them bit for bit
(``tests/test_oracle.py``).  What is NOT pinned is TensorFlow's own kernels (Eigen
einsum/tensordot summation order, its SVD): they are replaced by NumPy/LAPACK.

Every function cites the reference file:line it follows (paths relative to the
reference root).  ``dtype`` is float32 to match the reference graph; pass
``np.float64`` for the error-budget variant.

Conventions (reference ``Notes for developement.md``):
  feature  (N, B, d)      labels (B, L)
  node     (d, Dl, Dr)    special node (L, d, Dl, Dr)
  C1s[i]   (B, Dl_i)  everything left of site i,   C2s[i] (B, Dr_i) everything right of i
  bond     (L, d, d, Dl, Dr)
"""
import numpy as np

SOFTMAX_XENT = 0
SQUARED = 1

FM_PRECOMPUTED = 0   # feature already holds phi
FM_ONE_SQRT3 = 1     # [1, sqrt(3)(2x-1)]   trmps/preprocessing/MNISTpreprocessing.py:55
FM_ONE_X = 2         # [1, x]               trmps/preprocessing/FashionMNISTpreprocessing.py:73
FM_COS_SIN = 3       # [cos(pi x/2), sin(pi x/2)]  (north_star benchmark map; not in the reference)


# --------------------------------------------------------------------------- feature map
def feature_map(pixels, kind, dtype=np.float32):
    """pixels (N, B) in [0,1] -> phi (N, B, 2).  MNISTpreprocessing.py:55 for FM_ONE_SQRT3."""
    x = np.asarray(pixels, dtype=dtype)
    if kind == FM_ONE_SQRT3:
        return np.stack([np.ones_like(x), (np.sqrt(3) * (2 * x - 1)).astype(dtype)], axis=-1)
    if kind == FM_ONE_X:
        return np.stack([np.ones_like(x), x], axis=-1)
    if kind == FM_COS_SIN:
        half_pi = dtype(np.pi / 2)
        return np.stack([np.cos(half_pi * x), np.sin(half_pi * x)], axis=-1).astype(dtype)
    raise ValueError("unknown feature map %r" % (kind,))


# --------------------------------------------------------------------------- datasource side
def scale_uint8(images_u8):
    """uint8 idx pixels -> float32 in [0,1] as the TF-1.2 MNIST loader does it (tensorflow/contrib/learn/python/learn/
    datasets/mnist.py, DataSet.__init__: astype(float32) then multiply by 1.0 / 255.0; un-vendored dependency of
    MNISTpreprocessing.py:12, restated from its published source)."""
    return np.multiply(np.asarray(images_u8).astype(np.float32), 1.0 / 255.0)


def avg_pool_2x2(images, rows, cols):
    """tf.nn.avg_pool(ksize=[1,2,2,1], strides=[1,2,2,1], 'SAME') + reshape to [rows/2 * cols/2]
    (MNISTpreprocessing.py:44-48) for even rows/cols.  (B, rows*cols) -> (B, rows*cols/4), float32.
    Summation order inside a window: (top-left + top-right) + (bottom-left + bottom-right); TensorFlow's own order is
    not pinned (at most one ulp apart)."""
    x = np.asarray(images, dtype=np.float32).reshape(-1, rows // 2, 2, cols // 2, 2)
    top = x[:, :, 0, :, 0] + x[:, :, 0, :, 1]
    bottom = x[:, :, 1, :, 0] + x[:, :, 1, :, 1]
    return ((top + bottom) * np.float32(0.25)).reshape(x.shape[0], -1)


def max_pool_2x2(images, rows, cols):
    """skimage.measure.block_reduce(image, (2, 2), np.max) per image, then flattened (FashionMNISTpreprocessing.py:61-68)."""
    x = np.asarray(images).reshape(-1, rows // 2, 2, cols // 2, 2)
    return x.max(axis=(2, 4)).reshape(x.shape[0], -1)


def preprocess_fashion(images_u8, rows, cols, shrink):
    """FashionMNISTDatasource._load_with_correct_shape (FashionMNISTpreprocessing.py:56-77): u8 / 255.0 in float64,
    optional 2x2 maximum, phi = [1, x].  (B, rows*cols) uint8 -> (B, N, 2) float64, as the reference stores it."""
    x = np.asarray(images_u8).reshape(len(images_u8), rows * cols) / 255.0
    if shrink:
        x = max_pool_2x2(x, rows, cols)
    return np.stack([np.ones_like(x), x], axis=-1)


def feature_map_f32(x, kind):
    """The feature maps evaluated the way a float32 TF graph does (constants cast to float32 first):
    MNISTpreprocessing.py:55, FashionMNISTpreprocessing.py:73.  (...,) -> (..., 2)."""
    x = np.asarray(x, dtype=np.float32)
    if kind == FM_ONE_SQRT3:
        return np.stack([np.ones_like(x), np.float32(np.sqrt(3)) * (np.float32(2) * x - np.float32(1))], axis=-1)
    return feature_map(x, kind)


def preprocess_images(images, rows, cols, shrink, kind=FM_ONE_SQRT3):
    """_preprocess_images (MNISTpreprocessing.py:14-78) for a whole array of images instead of one sess.run per image:
    (B, rows*cols) float32 -> (B, N, 2) phi in the datasource's storage layout; kind = FM_PRECOMPUTED returns the
    (B, N) pooled pixels."""
    x = np.asarray(images, dtype=np.float32).reshape(len(images), rows * cols)
    if shrink:
        x = avg_pool_2x2(x, rows, cols)
    return x if kind == FM_PRECOMPUTED else feature_map_f32(x, kind)


def next_batch(data, labels, start, batch_size):
    """MPSDatasource.next_training_data_batch (preprocessing.py:165-208): batch_size consecutive samples from `start`,
    wrapping to index 0 at the end of the dataset; features with the first two axes swapped.
    -> ((N, B, ...), (B, L), next start)."""
    total = len(data)
    rows = (start + np.arange(batch_size)) % total
    return np.swapaxes(data[rows], 0, 1), labels[rows], int((start + batch_size) % total)


def lin_reg(data, labels, start, iterations, learning_rate, loss_kind=SOFTMAX_XENT, batch=100, dtype=np.float64):
    """MPS._lin_reg (mps.py:331-399): gradient descent from zero on z = x W + b with x = phi[..., 1] (mps.py:345-349),
    cost = mean softmax cross-entropy (mps.py:327-329) or 0.5 * sum of squares (squaredDistanceMPS.py:29-30), one minibatch
    of `batch` consecutive samples per iteration (next_training_data_batch(100), wrapping).  data (total, N, 2), labels
    (total, L).  Returns weight (N, L), bias (L) and the next read position.  TensorFlow's autodiff is restated as the
    closed-form gradient; pinned against the reference's own _lin_reg run under the shim, whose tf.train optimizer
    differentiates the reference's cost graph numerically (tests/golden/ref_linreg_*.npz, agreement 7e-8)."""
    data, labels = np.asarray(data, dtype=dtype), np.asarray(labels, dtype=dtype)
    total, N = data.shape[0], data.shape[1]
    W, b = np.zeros((N, labels.shape[1]), dtype=dtype), np.zeros(labels.shape[1], dtype=dtype)
    lr = dtype(learning_rate)
    for it in range(iterations):
        rows = (start + it * batch + np.arange(batch)) % total
        x, y = data[rows, :, 1], labels[rows]
        z = x @ W + b
        g = (z - y) if loss_kind == SQUARED else (softmax(z) - y) / dtype(batch)
        W -= lr * (x.T @ g)
        b -= lr * g.sum(axis=0)
    return W, b, int((start + iterations * batch) % total)


# --------------------------------------------------------------------------- initial MPS
def start_vector(L, dtype=np.float32):
    """trmps/mps/mps.py:446-454  [0_L, 1_L]"""
    v = np.zeros((1, 2 * L), dtype=dtype)
    v[0, L:] = 1
    return v


def end_vector(L, dtype=np.float32):
    """trmps/mps/mps.py:456-466  [1_L, 0_L]"""
    v = np.zeros((1, 2 * L), dtype=dtype)
    v[0, :L] = 1
    return v


def default_linear_model(N, d, L):
    """trmps/mps/mps.py:211-219 (_random_prepare): uniform weights, zero bias, loc=N//2."""
    weight = np.ones((N, d - 1, L)) / (N * (d - 1))
    bias = np.zeros((L,))
    return weight, bias, int(np.floor(N / 2))


def middle_node(index, weight, d, L, dtype=np.float32):
    """trmps/mps/mps.py:485-497"""
    D = 2 * L
    node = np.zeros((d, D, D), dtype=dtype)
    node[0] = np.identity(D)
    for k in range(1, d):
        node[k, L:, :L] = np.diag(weight[index, k - 1])
    return node


def special_node(index, weight, bias, d, L, dtype=np.float32):
    """trmps/mps/mps.py:468-482"""
    D = 2 * L
    out = np.zeros((L, d, D, D), dtype=dtype)
    for i in range(L):
        out[i, 0, i, i] = 1
        out[i, 0, i + L, i + L] = 1
        out[i, 0, i + L, i] = bias[i]
        out[i, 1:, i + L, i] = weight[index, :, i]
    return out


def initial_nodes(N, d, L, loc=None, weight=None, bias=None, dtype=np.float32):
    """trmps/mps/mps.py:401-423 (_setup_nodes) on top of prepare():189-209."""
    if weight is None:
        weight, bias, default_loc = default_linear_model(N, d, L)
        if loc is None:
            loc = default_loc
    nodes = []
    for i in range(N):
        if i == loc:
            nodes.append(special_node(i, weight, bias, d, L, dtype))
        else:
            nodes.append(middle_node(i, weight, d, L, dtype))
    return nodes, loc


COS_SIN_SCALE = 0.7893698     # exp(-E[log(cos(pi x/2) + sin(pi x/2))]), x ~ U[0,1)


def random_full_bond_mps(N, d, L, D, loc=0, seed=1, noise=None, dtype=np.float32):
    """Full-bond initial MPS for fixed-D throughput runs: COS_SIN_SCALE * identity on both
    feature components plus N(0, s^2) noise, s = 0.1/sqrt(D); outer dims stay 2L (mps.py:451,:462).
    Not a reference function: the reference can only reach bond D by training."""
    rng = np.random.default_rng(seed)
    if noise is None:
        noise = 0.1 / np.sqrt(D)
    dims = [2 * L] + [D] * (N - 1) + [2 * L]
    nodes = []
    for i in range(N):
        Dl, Dr = dims[i], dims[i + 1]
        base = np.zeros((d, Dl, Dr))
        k = min(Dl, Dr)
        base[:, :k, :k] = np.eye(k) * COS_SIN_SCALE
        if i == loc:
            node = base[None] + rng.normal(0.0, noise, size=(L, d, Dl, Dr))
        else:
            node = base + rng.normal(0.0, noise, size=(d, Dl, Dr))
        nodes.append(node.astype(dtype))
    return nodes


# --------------------------------------------------------------------------- forward pass
def chain_right(C1, phi_i, node):
    """trmps/mps/mps.py:500-508; baseoptimizer.py:160-180 (_find_C1).
    Literal: materialise T[t,i,j] = sum_m phi[t,m] node[m,i,j], then C1'[t,j] = sum_i C1[t,i] T[t,i,j]."""
    T = np.tensordot(phi_i, node, axes=([1], [0]))
    return np.einsum('ti,tij->tj', C1, T)


def chain_left(C2, phi_i, node):
    """trmps/mps/mps.py:510-519; baseoptimizer.py:182-201 (_find_C2)."""
    T = np.tensordot(phi_i, node, axes=([1], [0]))
    return np.einsum('tij,tj->ti', T, C2)


def predict(nodes, loc, feature, L, round_output=False):
    """trmps/mps/mps.py:521-560."""
    dtype = nodes[0].dtype
    N, B = feature.shape[0], feature.shape[1]
    C1 = np.tile(start_vector(L, dtype), (B, 1))
    for c in range(0, loc):
        C1 = chain_right(C1, feature[c], nodes[c])
    C2 = np.tile(end_vector(L, dtype), (B, 1))
    for c in range(N - 1, loc, -1):
        C2 = chain_left(C2, feature[c], nodes[c])
    sp = np.einsum('lnij,tn->tlij', nodes[loc], feature[loc])
    C1 = np.einsum('ti,tlij->tlj', C1, sp)
    f = np.einsum('tli,ti->tl', C1, C2)
    if round_output:
        f = np.round(f)
    return f


# --------------------------------------------------------------------------- metrics
def softmax(f):
    z = f - f.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def cost_softmax_xent(f, labels):
    """trmps/mps/mps.py:267-280: mean softmax cross-entropy with logits."""
    z = f - f.max(axis=1, keepdims=True)
    logsm = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
    return (-(labels * logsm).sum(axis=1)).mean(dtype=f.dtype)


def cost_squared(f, labels):
    """trmps/mps/squaredDistanceMPS.py:25-27: 0.5 * mean over all B*L elements."""
    return f.dtype.type(0.5) * np.square(f - labels).mean(dtype=f.dtype)


def cost(f, labels, loss_kind=SOFTMAX_XENT):
    return cost_softmax_xent(f, labels) if loss_kind == SOFTMAX_XENT else cost_squared(f, labels)


def accuracy(f, labels):
    """trmps/mps/mps.py:282-297."""
    return np.mean((f.argmax(axis=1) == labels.argmax(axis=1)).astype(np.float32))


def confusion_matrix(f, labels, L):
    """trmps/mps/mps.py:299-314 (rows = truth, columns = prediction)."""
    cm = np.zeros((L, L), dtype=np.int64)
    np.add.at(cm, (labels.argmax(axis=1), f.argmax(axis=1)), 1)
    return cm


def f1score(cm):
    """trmps/mps/mps.py:316-325."""
    diag = np.diag(cm).astype(np.float64)
    return float(np.sum(2 * diag / (cm.sum(0) + cm.sum(1))) / cm.shape[0])


# --------------------------------------------------------------------------- optimiser
class SweepParams(object):
    """Knobs of MPSOptimizerParameters (trmps/optimizers/parameterObjects.py:1-49) that reach the
    arithmetic, plus max_size (baseoptimizer.py:17) and the loss switch (baseoptimizer.py:347-352)."""

    def __init__(self, max_size, reg=0.001, min_singular_value=1e-4, armijo_coeff=1e-4,
                 use_hessian=False, updates_per_step=1, loss_kind=SOFTMAX_XENT, min_size=3):
        self.max_size = max_size
        self.reg = reg
        self.min_singular_value = min_singular_value
        self.armijo_coeff = armijo_coeff
        self.use_hessian = use_hessian and loss_kind == SOFTMAX_XENT   # baseoptimizer.py:32-35
        self.updates_per_step = updates_per_step
        self.loss_kind = loss_kind
        self.min_size = min_size


def setup_environments(nodes, loc, feature, L):
    """trmps/optimizers/optimizer.py:49-87: C1s[0..loc], C2s[loc+1..N-1]."""
    dtype = nodes[0].dtype
    N, B = feature.shape[0], feature.shape[1]
    C1s, C2s = {}, {}
    C1 = np.tile(start_vector(L, dtype), (B, 1))
    C1s[0] = C1
    for c in range(0, loc):
        C1 = chain_right(C1, feature[c], nodes[c])
        C1s[c + 1] = C1
    C2 = np.tile(end_vector(L, dtype), (B, 1))
    C2s[N - 1] = C2
    for c in range(N - 1, loc + 1, -1):
        C2 = chain_left(C2, feature[c], nodes[c])
        C2s[c - 1] = C2
    return C1s, C2s


def calculate_C(C1, C2, input1, input2):
    """trmps/optimizers/optimizer.py:197-220: C[t,m,n,i,k] = C1[t,i] C2[t,k] in1[t,m] in2[t,n], materialised."""
    B = C1.shape[0]
    a = C1.reshape(B, 1, 1, -1, 1) * C2.reshape(B, 1, 1, 1, -1)
    b = input1.reshape(B, -1, 1, 1, 1) * input2.reshape(B, 1, -1, 1, 1)
    return a * b


def get_f_and_cost(bond, C, labels, p):
    """optimizer.py:222-227 + baseoptimizer.py:347-352: returns (h, cost) where h is softmax(f)
    for the cross-entropy model and f itself for the squared-distance model."""
    f = np.tensordot(C, bond, axes=([1, 2, 3, 4], [1, 2, 3, 4]))
    c = cost(f, labels, p.loss_kind)
    h = softmax(f) if p.loss_kind == SOFTMAX_XENT else f
    return h, c, f


def calculate_hessian(h, C, p):
    """optimizer.py:229-237."""
    B, L = h.shape
    f_part = (h * (1 - h)).reshape(B, L, 1, 1, 1, 1)
    C_sq = np.square(C).reshape((B, 1) + C.shape[1:])
    return (f_part * C_sq).sum(axis=0) + 2 * p.reg


def armijo_loop(bond, C, labels, lr, cost0, delta, gdc, p):
    """baseoptimizer.py:299-322.  Counter starts at 1, bound 20 -> at most 19 trials; lr halves
    after every trial whether or not it was accepted."""
    counter, cond = 1, True
    updated = bond
    dt = bond.dtype.type
    trials = 0
    while cond and counter < 20:
        updated = bond + dt(lr) * delta
        _, c_new, _ = get_f_and_cost(updated, C, labels, p)
        target = cost0 - dt(p.armijo_coeff) * dt(lr) * gdc
        cond = bool(c_new > target)
        if cond:
            updated = bond
        counter += 1
        lr = lr * 0.5
        trials += 1
    return lr, updated, trials


def update_bond(bond, C, labels, lr, p, trace=None):
    """optimizer.py:239-264."""
    B = C.shape[0]
    dt = bond.dtype.type
    h, cost0, f0 = get_f_and_cost(bond, C, labels, p)
    hess = dt(1.0)
    if p.use_hessian:
        hess = calculate_hessian(h, C, p)
    gradient = np.tensordot(labels - h, C, axes=([0], [0])) - dt(2 * p.reg) * bond
    delta = gradient / hess
    gdc = np.tensordot(gradient, delta, axes=([0, 1, 2, 3, 4], [0, 1, 2, 3, 4])) / dt(B)
    _, updated, trials = armijo_loop(bond, C, labels, lr, cost0, delta, gdc, p)
    _, cost1, _ = get_f_and_cost(updated, C, labels, p)
    accepted = bool(cost1 < cost0)
    if not accepted:
        updated = bond
    if trace is not None:
        trace.append(dict(cost0=float(cost0), cost1=float(cost1), gdc=float(gdc), trials=trials,
                          accepted=accepted, grad_norm=float(np.linalg.norm(gradient.astype(np.float64))),
                          f0=f0, gradient=gradient))
    return updated


def repeatedly_update_bond(bond, C, labels, lr, p, trace=None):
    """baseoptimizer.py:324-332."""
    for _ in range(p.updates_per_step):
        bond = update_bond(bond, C, labels, lr, p, trace)
    return bond


def bond_decomposition(bond, p, trace=None):
    """optimizer.py:266-321 + utils.check_nan (utils.py:9-20).
    Returns a_j (d, Dl, m) and a_j1 (m, L, d, Dr)."""
    br = np.transpose(bond, (1, 3, 0, 2, 4))
    dims = br.shape
    flat = br.reshape(dims[0] * dims[1], dims[2] * dims[3] * dims[4])
    u, s, vh = np.linalg.svd(flat, full_matrices=False)
    u = np.where(np.isnan(u), 0, u)
    vh = np.where(np.isnan(vh), 0, vh)
    s_size = int(np.count_nonzero(s > p.min_singular_value))
    if s_size < p.min_size:
        m = p.min_size
    elif s_size > p.max_size:
        m = p.max_size
    else:
        m = s_size
    a_j = u[:, :m].reshape(dims[0], dims[1], -1)
    sv = np.diag(s[:m]) @ vh[:m]
    a_j1 = sv.reshape(-1, dims[2], dims[3], dims[4])
    if trace is not None:
        trace[-1]['s'] = s.copy()
        trace[-1]['m'] = a_j.shape[2]
    return a_j, a_j1


def update_right(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace=None):
    """optimizer.py:144-195: bond (c, c+1); special node moves c -> c+1; writes C1s[c+1]."""
    n2 = nodes[c + 1]
    bond = np.einsum('lmij,njk->lmnik', n1, n2)
    C = calculate_C(C1s[c], C2s[c + 1], feature[c], feature[c + 1])
    if trace is not None:
        trace.append_site(c, +1)
    bond = repeatedly_update_bond(bond, C, labels, lr, p, trace)
    a_j, a_j1 = bond_decomposition(bond, p, trace)
    a_j1 = np.transpose(a_j1, (1, 2, 0, 3))
    C1s[c + 1] = chain_right(C1s[c], feature[c], a_j)
    return a_j, a_j1


def update_left(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace=None):
    """optimizer.py:89-142: bond (c-1, c) in mirrored index order; special node moves c -> c-1; writes C2s[c-1]."""
    n2 = nodes[c - 1]
    bond = np.einsum('nkj,lmji->lmnik', n2, n1)
    C = calculate_C(C2s[c], C1s[c - 1], feature[c], feature[c - 1])
    if trace is not None:
        trace.append_site(c, -1)
    bond = repeatedly_update_bond(bond, C, labels, lr, p, trace)
    a_j, a_j1 = bond_decomposition(bond, p, trace)
    a_j = np.transpose(a_j, (0, 2, 1))
    a_j1 = np.transpose(a_j1, (1, 2, 3, 0))
    C2s[c - 1] = chain_left(C2s[c], feature[c], a_j)
    return a_j, a_j1


class Trace(list):
    """Per-bond-update record (one dict per _update_bond call) used for golden vectors."""

    def __init__(self):
        super().__init__()
        self.sites = []

    def append_site(self, c, direction):
        self.sites.append((c, direction))


def sweep_right(nodes, frm, to, C1s, C2s, feature, labels, lr, p, trace=None):
    """baseoptimizer.py:276-297: for c = frm .. to-1; final special node written at `to`."""
    nodes = list(nodes)
    n1 = nodes[frm]
    for c in range(frm, to):
        a_j, n1 = update_right(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace)
        nodes[c] = a_j
    nodes[to] = n1
    return nodes


def sweep_left(nodes, frm, to, C1s, C2s, feature, labels, lr, p, trace=None):
    """baseoptimizer.py:246-274: for c = frm .. to+1 downwards; final special node written at `to`."""
    nodes = list(nodes)
    n1 = nodes[frm]
    for c in range(frm, to, -1):
        a_j, n1 = update_left(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace)
        nodes[c] = a_j
    nodes[to] = n1
    return nodes


def train_step(nodes, loc, feature, labels, lr, p, L, trace=None):
    """baseoptimizer.py:203-244 on top of optimizer.py:49-87: right half-sweep loc->N-1, full left
    sweep N-1->0, right half-sweep 0->loc.  Returns the N updated nodes (special node back at loc)."""
    N = feature.shape[0]
    dtype = nodes[0].dtype
    feature = feature.astype(dtype, copy=False)
    labels = labels.astype(dtype, copy=False)
    C1s, C2s = setup_environments(nodes, loc, feature, L)
    nodes = sweep_right(nodes, loc, N - 1, C1s, C2s, feature, labels, lr, p, trace)
    nodes = sweep_left(nodes, N - 1, 0, C1s, C2s, feature, labels, lr, p, trace)
    nodes = sweep_right(nodes, 0, loc, C1s, C2s, feature, labels, lr, p, trace)
    return nodes


# --------------------------------------------------------------------------- single-site DMRG
def setup_environments_single(nodes, loc, feature, L):
    """singlesiteOptimizer.py:9-47: C1s[0..loc], C2s[loc..N-1] (one more right environment than the two-site setup)."""
    dtype = nodes[0].dtype
    N, B = feature.shape[0], feature.shape[1]
    C1s, C2s = {}, {}
    C1 = np.tile(start_vector(L, dtype), (B, 1))
    C1s[0] = C1
    for c in range(0, loc):
        C1 = chain_right(C1, feature[c], nodes[c])
        C1s[c + 1] = C1
    C2 = np.tile(end_vector(L, dtype), (B, 1))
    C2s[N - 1] = C2
    for c in range(N - 1, loc, -1):
        C2 = chain_left(C2, feature[c], nodes[c])
        C2s[c - 1] = C2
    return C1s, C2s


def calculate_C_single(C1, C2, inp):
    """singlesiteOptimizer.py:140-152: C[t,m,i,k] = C1[t,i] C2[t,k] phi[t,m]."""
    return np.einsum('ti,tk,tm->tmik', C1, C2, inp)


def get_f_and_cost_single(bond, C, labels, p):
    """singlesiteOptimizer.py:154-159 + baseoptimizer.py:347-352."""
    f = np.tensordot(C, bond, axes=([1, 2, 3], [1, 2, 3]))
    c = cost(f, labels, p.loss_kind)
    h = softmax(f) if p.loss_kind == SOFTMAX_XENT else f
    return h, c, f


def update_bond_single(bond, C, labels, lr, p, trace=None):
    """singlesiteOptimizer.py:174-200 (Hessian: not implemented in the reference either, :161-172)."""
    B = C.shape[0]
    dt = bond.dtype.type
    h, cost0, f0 = get_f_and_cost_single(bond, C, labels, p)
    gradient = np.tensordot(labels - h, C, axes=([0], [0])) - dt(2 * p.reg) * bond
    delta = gradient
    gdc = np.tensordot(gradient, delta, axes=([0, 1, 2, 3], [0, 1, 2, 3])) / dt(B)
    counter, cond, lr_t, trials, updated = 1, True, lr, 0, bond
    while cond and counter < 20:                       # baseoptimizer.py:299-322
        updated = bond + dt(lr_t) * delta
        _, c_new, _ = get_f_and_cost_single(updated, C, labels, p)
        cond = bool(c_new > cost0 - dt(p.armijo_coeff) * dt(lr_t) * gdc)
        if cond:
            updated = bond
        counter += 1
        lr_t = lr_t * 0.5
        trials += 1
    _, cost1, _ = get_f_and_cost_single(updated, C, labels, p)
    accepted = bool(cost1 < cost0)
    if not accepted:
        updated = bond
    if trace is not None:
        trace.append(dict(cost0=float(cost0), cost1=float(cost1), gdc=float(gdc), trials=trials, accepted=accepted,
                          f0=f0, gradient=gradient))
    return updated


def bond_decomposition_single(bond, p, trace=None):
    """singlesiteOptimizer.py:202-237: bond (L, d, Dl, Dr) -> a_j (d, Dl, m), a_j1 (m, L, Dr)."""
    br = np.transpose(bond, (1, 2, 0, 3))
    dims = br.shape
    flat = br.reshape(dims[0] * dims[1], dims[2] * dims[3])
    u, s, vh = np.linalg.svd(flat, full_matrices=False)
    u = np.where(np.isnan(u), 0, u)
    vh = np.where(np.isnan(vh), 0, vh)
    s_size = int(np.count_nonzero(s > p.min_singular_value))
    m = p.min_size if s_size < p.min_size else p.max_size if s_size > p.max_size else s_size
    a_j = u[:, :m].reshape(dims[0], dims[1], -1)
    a_j1 = (np.diag(s[:m]) @ vh[:m]).reshape(-1, dims[2], dims[3])
    if trace is not None:
        trace[-1]['s'] = s.copy()
        trace[-1]['m'] = a_j.shape[2]
    return a_j, a_j1


def update_right_single(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace=None):
    """singlesiteOptimizer.py:95-138: site c; the special node is split and its S V part absorbed by node c+1."""
    C = calculate_C_single(C1s[c], C2s[c], feature[c])
    if trace is not None:
        trace.append_site(c, +1)
    bond = n1
    for _ in range(p.updates_per_step):
        bond = update_bond_single(bond, C, labels, lr, p, trace)
    a_j, a_j1 = bond_decomposition_single(bond, p, trace)
    new_node = np.transpose(np.tensordot(a_j1, nodes[c + 1], axes=([2], [1])), (1, 2, 0, 3))     # (L, d, m, Dr')
    C1s[c + 1] = chain_right(C1s[c], feature[c], a_j)
    return a_j, new_node


def update_left_single(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace=None):
    """singlesiteOptimizer.py:49-93: site c moving left (bond transposed [0,1,3,2]); S V absorbed by node c-1."""
    C = calculate_C_single(C2s[c], C1s[c], feature[c])
    if trace is not None:
        trace.append_site(c, -1)
    bond = np.transpose(n1, (0, 1, 3, 2))
    for _ in range(p.updates_per_step):
        bond = update_bond_single(bond, C, labels, lr, p, trace)
    a_j, a_j1 = bond_decomposition_single(bond, p, trace)
    a_j = np.transpose(a_j, (0, 2, 1))
    a_j1 = np.transpose(a_j1, (2, 1, 0))                                                          # (Dl, L, m)
    new_node = np.transpose(np.tensordot(a_j1, nodes[c - 1], axes=([0], [2])), (0, 2, 3, 1))       # (L, d, Dl', m)
    C2s[c - 1] = chain_left(C2s[c], feature[c], a_j)
    return a_j, new_node


def train_step_single(nodes, loc, feature, labels, lr, p, L, trace=None):
    """baseoptimizer.py:203-244 with the single-site updates: loc -> N-1, N-1 -> 0, 0 -> loc."""
    N = feature.shape[0]
    dtype = nodes[0].dtype
    feature = feature.astype(dtype, copy=False)
    labels = labels.astype(dtype, copy=False)
    C1s, C2s = setup_environments_single(nodes, loc, feature, L)
    nodes = list(nodes)

    def right(frm, to, nodes):
        n1 = nodes[frm]
        for c in range(frm, to):
            a_j, n1 = update_right_single(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace)
            nodes[c] = a_j
        nodes[to] = n1
        return nodes

    def left(frm, to, nodes):
        n1 = nodes[frm]
        for c in range(frm, to, -1):
            a_j, n1 = update_left_single(c, n1, nodes, C1s, C2s, feature, labels, lr, p, trace)
            nodes[c] = a_j
        nodes[to] = n1
        return nodes

    nodes = right(loc, N - 1, nodes)
    nodes = left(N - 1, 0, nodes)
    nodes = right(0, loc, nodes)
    return nodes


def learning_rate(initial_lr, lr_reg, step):
    """baseoptimizer.py:109."""
    return initial_lr / np.sqrt(1 - lr_reg ** (step + 1))


# --------------------------------------------------------------------------- factored variant
def forward_factored(bond, C1, C2, in1, in2):
    """Same f as get_f_and_cost without materialising C: f[t,l] = sum P[t,(m,i)] bond[l,m,n,i,k] Q[t,(n,k)]."""
    L, d1, d2, Dl, Dr = bond.shape
    P = (in1[:, :, None] * C1[:, None, :]).reshape(-1, d1 * Dl)
    Q = (in2[:, :, None] * C2[:, None, :]).reshape(-1, d2 * Dr)
    Bm = np.transpose(bond, (1, 3, 0, 2, 4)).reshape(d1 * Dl, L, d2 * Dr)
    X = (P @ Bm.reshape(d1 * Dl, L * d2 * Dr)).reshape(-1, L, d2 * Dr)
    return np.einsum('tlc,tc->tl', X, Q)


def gradient_factored(resid, C1, C2, in1, in2):
    """G[l,m,n,i,k] = sum_t resid[t,l] P[t,(m,i)] Q[t,(n,k)] without materialising C."""
    B, L = resid.shape
    d1, d2, Dl, Dr = in1.shape[1], in2.shape[1], C1.shape[1], C2.shape[1]
    P = (in1[:, :, None] * C1[:, None, :]).reshape(B, d1 * Dl)
    Q = (in2[:, :, None] * C2[:, None, :]).reshape(B, d2 * Dr)
    W = (resid[:, :, None] * Q[:, None, :]).reshape(B, L * d2 * Dr)
    G = (P.T @ W).reshape(d1, Dl, L, d2, Dr)
    return np.transpose(G, (2, 0, 3, 1, 4))


def chain_right_factored(C1, phi_i, node):
    d, Dl, Dr = node.shape
    P = (phi_i[:, :, None] * C1[:, None, :]).reshape(-1, d * Dl)
    return P @ node.reshape(d * Dl, Dr)


def chain_left_factored(C2, phi_i, node):
    d, Dl, Dr = node.shape
    P = (phi_i[:, :, None] * C2[:, None, :]).reshape(-1, d * Dr)
    return P @ np.transpose(node, (0, 2, 1)).reshape(d * Dr, Dl)


# --------------------------------------------------------------------------- checkpoint format
def save_bytes(nodes, loc, d, L):
    """trmps/mps/mps.py:126-160: big-endian header, N right-bond dims (int16), float32 C-order tensors."""
    N = len(nodes)
    out = (True).to_bytes(1, 'big') + N.to_bytes(2, 'big') + loc.to_bytes(2, 'big')
    out += d.to_bytes(2, 'big') + L.to_bytes(2, 'big')
    for i, w in enumerate(nodes):
        out += int(w.shape[3] if i == loc else w.shape[2]).to_bytes(2, 'big')
    for w in nodes:
        out += np.ascontiguousarray(w, dtype=np.float32).tobytes(order='C')
    return out


def load_bytes(raw):
    """trmps/mps/mps.py:80-124: reads N-1 dims starting at byte 9; weights start at byte 9+2N."""
    has_w = bool(raw[0])
    N = int.from_bytes(raw[1:3], 'big')
    loc = int.from_bytes(raw[3:5], 'big')
    d = int.from_bytes(raw[5:7], 'big')
    L = int.from_bytes(raw[7:9], 'big')
    if not has_w:
        return dict(N=N, loc=loc, d=d, L=L, nodes=None, dims=None)
    dims = [int.from_bytes(raw[i:i + 2], 'big') for i in range(9, 9 + (N - 1) * 2, 2)]
    pos = 9 + N * 2
    nodes = []
    for i in range(N):
        l_dim = 2 * L if i == 0 else dims[i - 1]
        r_dim = 2 * L if i == N - 1 else dims[i]
        if i == loc:
            n = l_dim * r_dim * d * L
            w = np.frombuffer(raw[pos:pos + 4 * n], dtype=np.float32).reshape(L, d, l_dim, r_dim)
        else:
            n = l_dim * r_dim * d
            w = np.frombuffer(raw[pos:pos + 4 * n], dtype=np.float32).reshape(d, l_dim, r_dim)
        nodes.append(w)
        pos += 4 * n
    return dict(N=N, loc=loc, d=d, L=L, nodes=nodes, dims=dims)


# --------------------------------------------------------------------------- synthetic data
def synthetic_batch(N, B, L, seed=0, dtype=np.float32):
    """pixels ~ U[0,1) site-major (N, B); labels one-hot (B, L)."""
    rng = np.random.default_rng(seed)
    pixels = rng.random((N, B), dtype=np.float32).astype(dtype)
    cls = rng.integers(0, L, B)
    labels = np.zeros((B, L), dtype=dtype)
    labels[np.arange(B), cls] = 1
    return pixels, labels
