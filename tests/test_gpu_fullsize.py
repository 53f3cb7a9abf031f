"""-m gpu tests at BASELINE.json's full sizes (cfg3: D=120, 60 000 images; cfg5: N=784, D=32, 1 M images).

The oracle cannot walk a whole 60k-image sweep in seconds, so these cases use what is independent of the size:
  * everything the path computes per image (environments, logits) is checked on a random sample of the images against
    the fp64 oracle, and must not depend on which other images share the batch;
  * what is summed over the batch (the gradient, the cost) is checked in full against fp64 matrix products, which NumPy
    does in about a second at this size;
  * exact power-of-two scaling of the weights scales the logits exactly;
  * the full-size split (240 x 2400) is checked against LAPACK, with orthonormal a_j and a_j . a_j1 = rank-m truncation;
  * the linearity the Armijo loop relies on, f(bond + lr delta) = f(bond) + lr f(delta).
Tolerances: north_star's 1e-4 relative for logits, gradients and singular values; argmax exact on non-tied rows."""
import numpy as np
import pytest
import torch

import trmps_oracle as o
import trmps_native as nat
import trmps_device as dev

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def check_inference(N, D, B, sample, seed):
    L, loc = 10, N // 2
    device = torch.device("cuda", 0)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=seed)
    w = dev.DeviceWeights(nodes, loc, L, dev.bond_stride(nodes, loc, L), device)
    gen = torch.Generator(device=device).manual_seed(seed)
    pix = torch.rand((N, B), device=device, generator=gen)
    logits = dev.predict(w, pix, nat.FM_COS_SIN)
    assert bool(torch.isfinite(logits).all())
    pick = torch.from_numpy(np.random.default_rng(seed).choice(B, sample, replace=False)).to(device)
    sub = pix[:, pick].contiguous()
    # (1) a random sample of the images against the fp64 oracle
    phi64 = o.feature_map(sub.cpu().numpy().astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    want = o.predict([n.astype(np.float64) for n in nodes], loc, phi64, L)
    got = logits[pick].cpu().numpy()
    assert rel(got, want) < RTOL
    top2 = np.sort(want, axis=1)
    clear = (top2[:, -1] - top2[:, -2]) > 1e-4 * np.abs(top2[:, -1])
    assert clear.sum() > sample // 2
    assert np.array_equal(got.argmax(1)[clear], want.argmax(1)[clear])
    # (2) an image's logits do not depend on the rest of the batch
    alone = dev.predict(w, sub, nat.FM_COS_SIN).cpu().numpy()
    assert rel(alone, got) < 2e-6
    # (3) exact power-of-two scaling of the special node scales every logit exactly
    scaled = [n if i != loc else n * np.float32(4) for i, n in enumerate(nodes)]
    w4 = dev.DeviceWeights(scaled, loc, L, w.Ds, device)
    assert torch.equal(dev.predict(w4, sub, nat.FM_COS_SIN), 4 * torch.from_numpy(alone).to(device))


def test_cfg5_inference_one_million_images():
    check_inference(784, 32, 1000000, 192, seed=5)


def test_cfg3_inference_sixty_thousand_images():
    check_inference(196, 120, 60000, 96, seed=6)


@pytest.mark.parametrize("direction", [+1, -1])
@pytest.mark.parametrize("D,B", [(120, 60000), (64, 60000), (20, 20000)], ids=["cfg3", "cfg4", "cfg2"])
def test_bond_update_pieces_at_full_size(direction, D, B):
    """One bond at the bond dimension and batch of cfg3 (D=120: bond matrix 240 x 2400, 60 000 images), cfg4 (D=64) and
    cfg2 (D=20, 20 000 images) on a short chain: forward, cost, gradient and <G, delta> against fp64 matrix products over
    the whole batch; linearity of the forward; the split against LAPACK."""
    N, L, loc = 5, 10, 2
    pix, lab = o.synthetic_batch(N, B, L, seed=70)
    phi = o.feature_map(pix, o.FM_COS_SIN)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=71)
    Ds = dev.bond_stride(nodes, loc, L, D)
    st = dev.SweepState(N, B, L, Ds, nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc)
    st.set_batch(pix, lab)
    st.init_env(loc)
    n64 = [n.astype(np.float64) for n in nodes]
    phi64 = o.feature_map(pix.astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    C1s, C2s = o.setup_environments(n64, loc, phi64, L)
    if direction > 0:
        bond = np.einsum('lmij,njk->lmnik', n64[loc], n64[loc + 1])
        Cn, Cf, pn, pf = C1s[loc], C2s[loc + 1], phi64[loc], phi64[loc + 1]
    else:
        lib = nat.load()
        nat.check(lib.trmps_env_step(st.feat[loc + 1].data_ptr(), nat.FM_COS_SIN, B, st.Ds, -1, st.envR[loc + 1].data_ptr(),
                                     st.slots[loc + 1].data_ptr(), st.dims[loc + 2:].data_ptr(), st.envR[loc].data_ptr(), None), "env_step")
        C2 = o.chain_left(C2s[loc + 1], phi64[loc + 1], n64[loc + 1])
        bond = np.einsum('nkj,lmji->lmnik', n64[loc - 1], n64[loc])
        Cn, Cf, pn, pf = C2, C1s[loc - 1], phi64[loc], phi64[loc - 1]
    reg = 1e-3
    opt = dev.make_opt(lr=1e-3, max_size=D, reg=reg, min_singular_value=0.0)
    st.bond_merge(loc, direction)
    st.bond_gradient(loc, direction, opt)          # forward, loss and gradient of the merged bond
    torch.cuda.synchronize()
    # ---- per-image quantity: logits of every image
    f64 = o.forward_factored(bond, Cn, Cf, pn, pf)
    f0 = st.logits0().cpu().numpy()
    assert rel(f0, f64) < RTOL
    # ---- batch sums: cost, gradient, <G, delta>/B
    h64 = o.softmax(f64)
    cost64 = o.cost_softmax_xent(f64, lab.astype(np.float64))
    sc = st.scalars().cpu().numpy()
    assert abs(sc[nat.S_COST0] - cost64) < 1e-5 * abs(cost64)
    G64 = o.gradient_factored(lab - h64, Cn, Cf, pn, pf) - 2 * reg * bond
    dn, df = bond.shape[3], bond.shape[4]
    gotG = st.bond_tensor(st.grad_matrix()).cpu().numpy()[:, :, :, :dn, :df]
    assert rel(gotG, G64) < RTOL
    gdc = np.sum(G64 * G64) / B
    assert abs(sc[nat.S_GDC] - gdc) < 1e-4 * gdc
    # ---- the linearity the Armijo loop relies on (baseoptimizer.py:299-322 evaluates f(bond + lr delta) per trial)
    st.bond_forward(loc, direction, 1)              # f(delta), delta = the gradient just computed
    fd = st.logits_delta().cpu().numpy()
    lr = 1e-3
    f_sum = o.forward_factored(bond + lr * G64, Cn, Cf, pn, pf)
    assert rel(f0 + np.float32(lr) * fd, f_sum) < RTOL
    # ---- the split of the 240 x 2400 bond against LAPACK
    st.bond_split(loc, direction, opt)
    torch.cuda.synchronize()
    iw = st.iwork.cpu().numpy()
    m = int(iw[nat.I_M])
    assert m == D
    M = np.transpose(bond, (1, 3, 0, 2, 4)).reshape(2 * dn, -1)
    s_ref = np.linalg.svd(M, compute_uv=False)
    sv = st.singular_values().cpu().numpy()[:len(s_ref)]
    big = s_ref > 1e-3 * s_ref[0]
    assert np.max(np.abs(sv[big] - s_ref[big]) / s_ref[big]) < RTOL
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    want = ((U[:, :m] * s[:m]) @ Vt[:m]).reshape(2, dn, L, 2, df).transpose(2, 0, 3, 1, 4)      # rank-m truncation
    dims = st.dims.cpu().numpy()
    slots, special = st.slots.cpu().numpy(), st.special.cpu().numpy()
    if direction > 0:
        a_j, a_j1 = slots[loc][:, :dims[loc], :m], special[:, :, :m, :dims[loc + 2]]
        assert dims[loc + 1] == m
        got = np.einsum('miq,lnqk->lmnik', a_j, a_j1)
        ortho = np.einsum('miq,mir->qr', a_j, a_j)
    else:
        a_j, a_j1 = slots[loc][:, :m, :dims[loc + 1]], special[:, :, :dims[loc - 1], :m]
        assert dims[loc] == m
        got = np.einsum('mqi,lnkq->lmnik', a_j, a_j1)
        ortho = np.einsum('mqi,mri->qr', a_j, a_j)
    assert rel(got, want) < RTOL
    assert np.max(np.abs(ortho - np.eye(m))) < 3e-5


def test_long_wide_chain_keeps_the_tolerance():
    """N=784 at D=120 is longer than any BASELINE configuration: the fused chain's round-toward-zero accumulation would
    shrink the logits by ~2.4e-4 there, so trmps_predict takes the per-site FP32 path (trmps_set_option(2, 2) overrides)."""
    N, D, B, L = 784, 120, 256, 10
    loc = N // 2
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=9)
    w = dev.DeviceWeights(nodes, loc, L, dev.bond_stride(nodes, loc, L), torch.device("cuda", 0))
    pix = torch.rand((N, B), device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    phi64 = o.feature_map(pix.cpu().numpy().astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    want = o.predict([n.astype(np.float64) for n in nodes], loc, phi64, L)
    assert rel(dev.predict(w, pix, nat.FM_COS_SIN).cpu().numpy(), want) < RTOL
    try:
        nat.set_option(nat.OPT_CHAIN_IMPL, nat.CHAIN_TCGEN05_ALWAYS)
        forced = rel(dev.predict(w, pix, nat.FM_COS_SIN).cpu().numpy(), want)
    finally:
        nat.set_option(nat.OPT_CHAIN_IMPL, nat.CHAIN_TCGEN05)
    assert 1e-4 < forced < 5e-4          # the documented bias; argmax is unaffected (a common factor)


def test_cfg3_half_sweep_invariants():
    """The benchmark's workload itself (N=196, D=120, 60 000 images, 195 bond updates left to right) and what must hold
    after it, whatever the size: every node left behind is an isometry; the bond dimension is max_size wherever the rank allows it; the
    updated MPS evaluated by the inference path (fused tensor-core chains) and by the training path (cached environments of
    the sweep + the bond contraction) are the same function, and both match the fp64 oracle on a sample of the images."""
    from bench import synthetic_pixels, full_bond_mps
    N, D, L, B, loc = 196, 120, 10, 60000, 0
    pix, lab = synthetic_pixels(N, B, L, seed=100)
    nodes = full_bond_mps(N, 2, L, D, loc)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc)
    st.set_batch(pix, lab)
    opt = dev.make_opt(lr=1e-4, max_size=D, reg=1e-3, min_singular_value=0.0)
    st.init_env(loc)
    st.sweep(loc, N - 1, +1, opt)
    st.loc = N - 1
    torch.cuda.synchronize()
    iw = st.iwork.cpu().numpy()
    assert int(iw[nat.I_N_UPDATES]) == N - 1
    dims = st.dims.cpu().numpy()
    want_dims = [2 * L]
    for i in range(N - 1):
        want_dims.append(min(D, 2 * want_dims[-1]))           # min_singular_value = 0: m = min(max_size, rank bound d * Dl)
    assert dims[:N].tolist() == want_dims and dims[N] == 2 * L
    new_nodes = st.get_nodes()
    assert all(np.isfinite(w).all() for w in new_nodes)
    worst = 0.0
    for i in range(N - 1):                                   # a_j of every split: orthonormal columns (optimizer.py:303-306)
        a = new_nodes[i].astype(np.float64)
        g = np.einsum('miq,mir->qr', a, a)
        worst = max(worst, float(np.max(np.abs(g - np.eye(g.shape[0])))))
    assert worst < 5e-5
    # the same function through both kernel families, against fp64 on a sample
    sample = np.random.default_rng(3).choice(B, 48, replace=False)
    phi64 = o.feature_map(pix[:, sample].astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    want = o.predict([w.astype(np.float64) for w in new_nodes], N - 1, phi64, L)
    inference = dev.predict(st.weights_view(), st.feat, nat.FM_COS_SIN).cpu().numpy()
    assert rel(inference[sample], want) < RTOL
    st.bond_merge(N - 1, -1)                                  # bond (N-2, N-1) from the swept nodes; C1s[N-2] is the sweep's own cache
    st.bond_forward(N - 1, -1, 0)
    training = st.logits0().cpu().numpy()
    assert rel(training[sample], want) < RTOL
    top2 = np.sort(want, axis=1)
    print("cfg3 half-sweep: inference vs fp64 %.2e, training vs fp64 %.2e (sample of 48), training vs inference %.2e (all 60 000 images)"
          % (rel(inference[sample], want), rel(training[sample], want), rel(training, inference)))
    # the two kernel families agree inside north_star's tolerance, and on the class of every image whose two largest logits
    # are further apart than that tolerance
    assert rel(training, inference) < RTOL
    srt = np.sort(inference, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 2 * RTOL * np.abs(srt).max(axis=1)
    assert clear.mean() > 0.95
    assert np.array_equal(training.argmax(1)[clear], inference.argmax(1)[clear])


def test_cfg3_single_site_update_at_full_size():
    """One single-site update (singlesiteOptimizer.py:49-138) at cfg3's bond dimension with all 60 000 images: cost and
    <G, delta> against fp64 matrix products over the whole batch, the singular values of the updated node against LAPACK,
    and the function the two touched sites represent afterwards against the rank-m truncation."""
    N, B, L, D, loc = 5, 60000, 10, 120, 2
    pix, lab = o.synthetic_batch(N, B, L, seed=80)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=81)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc)
    st.set_batch(pix, lab)
    reg, lr = 1e-3, 1e-4
    opt = dev.make_opt(lr=lr, max_size=D, reg=reg, min_singular_value=0.0, single_site=True)
    st.init_env(loc)
    st.bond_step(loc, +1, opt)
    torch.cuda.synchronize()
    n64 = [n.astype(np.float64) for n in nodes]
    phi64 = o.feature_map(pix.astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    C1s, C2s = o.setup_environments_single(n64, loc, phi64, L)
    S = n64[loc]                                                                   # (L, d, Dl, Dr)
    Dl, Dr = S.shape[2], S.shape[3]
    P = (phi64[loc][:, :, None] * C1s[loc][:, None, :]).reshape(B, 2 * Dl)       # [t, (m, i)]
    Smat = np.transpose(S, (1, 2, 0, 3)).reshape(2 * Dl, L * Dr)                   # rows (m, i), columns (l, k)
    C2 = C2s[loc]
    f = np.einsum('tlk,tk->tl', (P @ Smat).reshape(B, L, Dr), C2)
    y = lab.astype(np.float64)
    cost0 = o.cost_softmax_xent(f, y)
    W = ((y - o.softmax(f))[:, :, None] * C2[:, None, :]).reshape(B, L * Dr)
    G = P.T @ W - 2 * reg * Smat
    sc, iw = st.scalars().cpu().numpy(), st.iwork.cpu().numpy()
    assert abs(sc[nat.S_COST0] - cost0) < 1e-5 * abs(cost0)
    gdc = np.sum(G * G) / B
    assert abs(sc[nat.S_GDC] - gdc) < 1e-4 * gdc
    lr_star = float(sc[nat.S_LR_STAR])
    assert (lr_star > 0) == bool(iw[nat.I_ACCEPT])
    M = Smat + lr_star * G                                                        # the node the split saw
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    m = int(iw[nat.I_M])
    assert m == D
    sv = st.singular_values().cpu().numpy()[:len(s)]
    big = s > 1e-3 * s[0]
    assert np.max(np.abs(sv[big] - s[big]) / s[big]) < RTOL
    # isometry left at `loc`, S V absorbed by node loc + 1 (singlesiteOptimizer.py:124-125): compare the two-site product
    want_aj1 = ((s[:m, None] * Vt[:m]).reshape(m, L, Dr))                          # (q, l, k)
    want = np.einsum('miq,qlk,nkj->lmnij', U[:, :m].reshape(2, Dl, m), want_aj1, n64[loc + 1])
    dims = st.dims.cpu().numpy()
    slots, special = st.slots.cpu().numpy(), st.special.cpu().numpy()
    a_j = slots[loc][:, :dims[loc], :m]
    new_special = special[:, :, :m, :dims[loc + 2]]                               # (l, n, q, j)
    got = np.einsum('miq,lnqj->lmnij', a_j, new_special)
    assert rel(got, want) < RTOL
    assert np.max(np.abs(np.einsum('miq,mir->qr', a_j, a_j) - np.eye(m))) < 3e-5


def test_cfg3_full_training_step_leaves_a_canonical_mps():
    """BaseOptimizer.train_step (baseoptimizer.py:203-244) at cfg3's size: right half-sweep, left sweep, right half-sweep =
    390 bond updates on 60 000 images.  Afterwards the MPS must be in mixed canonical form around the special node (every
    node to its left an isometry over (feature, left bond), every node to its right over (feature, right bond)), and the
    updated weights evaluated by the inference path must match the fp64 oracle on a sample of the images."""
    from bench import synthetic_pixels, full_bond_mps
    N, D, L, B = 196, 120, 10, 60000
    loc = N // 2
    pix, lab = synthetic_pixels(N, B, L, seed=100)
    nodes = full_bond_mps(N, 2, L, D, loc)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc)
    st.set_batch(pix, lab)
    st.train_step(dev.make_opt(lr=1e-4, max_size=D, reg=1e-3, min_singular_value=0.0), loc=loc)
    torch.cuda.synchronize()
    assert int(st.iwork.cpu().numpy()[nat.I_N_UPDATES]) == 2 * (N - 1)
    new_nodes = st.get_nodes()
    assert all(np.isfinite(w).all() for w in new_nodes)
    assert new_nodes[loc].ndim == 4
    worst = 0.0
    for i, w in enumerate(new_nodes):
        a = w.astype(np.float64)
        if i < loc:
            g = np.einsum('miq,mir->qr', a, a)
        elif i > loc:
            g = np.einsum('mqk,mrk->qr', a, a)
        else:
            continue
        worst = max(worst, float(np.max(np.abs(g - np.eye(g.shape[0])))))
    assert worst < 5e-5
    sample = np.random.default_rng(4).choice(B, 48, replace=False)
    phi64 = o.feature_map(pix[:, sample].astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    want = o.predict([w.astype(np.float64) for w in new_nodes], loc, phi64, L)
    got = dev.predict(st.weights_view(), st.feat, nat.FM_COS_SIN).cpu().numpy()
    assert rel(got[sample], want) < RTOL
