"""CPU tests of the oracle itself: golden vectors, self-consistency (literal vs factored, fp32 vs fp64) and the
invariants SURVEY.md section 8c lists."""
import glob
import os

import numpy as np
import pytest

import trmps_oracle as o

ALL_NPZ = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in ALL_NPZ if not os.path.basename(p).startswith("ref_")]     # written by oracle/make_golden.py
REF_GOLDEN = [p for p in ALL_NPZ if os.path.basename(p).startswith("ref_")      # written by the REFERENCE code under oracle/tf1_shim.py
              and os.path.basename(p) not in ("ref_preprocess.npz", "ref_fashion.npz", "ref_linreg_xent.npz", "ref_linreg_squared.npz")]   # (sweep cases; the others have their own tests)
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def golden_cfg(g):
    N, B, L, ms, loc, lr, loss, upd, seed = g['cfg']
    return int(N), int(B), int(L), int(ms), (None if loc < 0 else int(loc)), float(lr), int(loss), int(upd), int(seed)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    N, B, L, ms, loc, lr, loss, upd, seed = golden_cfg(g)
    pix, lab = o.synthetic_batch(N, B, L, seed=seed)
    assert np.array_equal(pix, g['pixels']) and np.array_equal(lab, g['labels'])
    phi = o.feature_map(pix, o.FM_ONE_SQRT3)
    nodes, loc = o.initial_nodes(N, 2, L, loc=loc)
    assert rel(o.predict(nodes, loc, phi, L), g['f_init']) < 1e-6
    tr = o.Trace()
    new = o.train_step(nodes, loc, phi, lab, lr, o.SweepParams(max_size=ms, updates_per_step=upd, loss_kind=loss), L, tr)
    assert [t['trials'] for t in tr] == g['trials'].tolist()
    assert [t['m'] for t in tr if 'm' in t] == g['m'].tolist()
    assert np.allclose([t['cost0'] for t in tr], g['cost0'], rtol=1e-4)
    assert rel(o.predict(new, loc, phi, L), g['f_final']) < 1e-3
    assert rel(g['f_final'], g['f_final_f64']) < 1e-3           # fp32 error budget of the whole step


def test_golden_files_present():
    assert len(GOLDEN) >= 3 and len(REF_GOLDEN) >= 3


@pytest.mark.parametrize("path", REF_GOLDEN, ids=[os.path.basename(p) for p in REF_GOLDEN])
def test_oracle_matches_reference_code_run_under_tf_shim(path):
    """tests/golden/ref_*.npz were produced by the reference's own trmps/mps/mps.py + trmps/optimizers/*.py executed
    under the NumPy TF1 stand-in (oracle/make_golden_from_reference.py).  The restatement must reproduce them."""
    g = np.load(path)
    N, B, L, ms, loc, lr, loss, upd, seed = golden_cfg(g)
    pix, lab = o.synthetic_batch(N, B, L, seed=seed)
    assert np.array_equal(pix, g['pixels']) and np.array_equal(lab, g['labels'])
    phi = o.feature_map(pix, o.FM_ONE_SQRT3)
    nodes, loc = o.initial_nodes(N, 2, L, loc=loc)
    assert loc == int(g['loc'])
    assert rel(o.predict(nodes, loc, phi, L), g['f_init']) < 1e-6
    single = loss == 2                      # generator code 2 = cross-entropy with the reference's SingleSiteMPSOptimizer
    loss = 0 if single else loss
    step = o.train_step_single if single else o.train_step
    new = step(nodes, loc, phi, lab, lr, o.SweepParams(max_size=ms, updates_per_step=upd, loss_kind=loss), L)
    assert [w.shape[-1] for w in new] == g['dims'].tolist()
    f = o.predict(new, loc, phi, L)
    assert rel(f, g['f_final']) < 1e-5
    assert abs(o.cost(f, lab, loss) - float(g['cost_final'])) < 1e-5 * abs(float(g['cost_final']))


@pytest.mark.skipif(not os.path.isdir("/root/reference/trmps"), reason="reference sources not present on this machine")
def test_reference_code_runs_under_the_shim_and_reproduces_committed_vectors(tmp_path):
    import subprocess, sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.setrecursionlimit(1000000); import threading; threading.stack_size(1<<30)\n"
            "import make_golden_from_reference as m\n"
            "out = {}\n"
            "def go():\n"
            "    import io, contextlib\n"
            "    with contextlib.redirect_stdout(io.StringIO()):\n"
            "        out['r'] = m.run_case(*m.CASES['ref_xent_n8'])\n"
            "t = threading.Thread(target=go); t.start(); t.join()\n"
            "g = np.load(%r)\n"
            "assert np.array_equal(out['r']['f_final'], g['f_final']) and np.array_equal(out['r']['dims'], g['dims'])\n"
            "print('ok')\n") % (os.path.join(root, "oracle"), os.path.join(root, "tests", "golden", "ref_xent_n8.npz"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stderr[-2000:]


def test_initial_mps_is_the_linear_model():
    N, B, L = 20, 50, 10
    rng = np.random.default_rng(0)
    weight, bias = rng.normal(size=(N, 1, L)), rng.normal(size=(L,))
    pix, _ = o.synthetic_batch(N, B, L, seed=1)
    phi = o.feature_map(pix, o.FM_ONE_SQRT3)
    for loc in (0, 7, N - 1):
        nodes, _ = o.initial_nodes(N, 2, L, loc=loc, weight=weight, bias=bias)
        lin = bias[None] + np.einsum('ntk,nkl->tl', phi[:, :, 1:], weight)
        assert rel(o.predict(nodes, loc, phi, L), lin) < 1e-6


def test_factored_forms_equal_literal():
    N, B, L, D = 6, 40, 10, 12
    pix, lab = o.synthetic_batch(N, B, L, seed=2)
    phi = o.feature_map(pix, o.FM_COS_SIN)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=2, seed=3)
    C1s, C2s = o.setup_environments(nodes, 2, phi, L)
    bond = np.einsum('lmij,njk->lmnik', nodes[2], nodes[3])
    C = o.calculate_C(C1s[2], C2s[3], phi[2], phi[3])
    p = o.SweepParams(max_size=D)
    h, c, f = o.get_f_and_cost(bond, C, lab, p)
    assert rel(o.forward_factored(bond, C1s[2], C2s[3], phi[2], phi[3]), f) < 1e-5
    G = np.tensordot(lab - h, C, axes=([0], [0]))
    assert rel(o.gradient_factored(lab - h, C1s[2], C2s[3], phi[2], phi[3]), G) < 1e-5
    assert rel(o.chain_right_factored(C1s[2], phi[2], nodes[3][:, :D, :]), o.chain_right(C1s[2], phi[2], nodes[3][:, :D, :])) < 1e-5
    assert rel(o.chain_left_factored(C2s[3], phi[3], nodes[3]), o.chain_left(C2s[3], phi[3], nodes[3])) < 1e-5


def test_forward_is_linear_in_the_bond():
    # f(B + lr*delta) = f(B) + lr*f(delta): what lets one forward on delta serve every Armijo trial
    rng = np.random.default_rng(4)
    B, L, D = 30, 10, 6
    C = rng.normal(size=(B, 2, 2, D, D)).astype(np.float32)
    b1 = rng.normal(size=(L, 2, 2, D, D)).astype(np.float32)
    b2 = rng.normal(size=(L, 2, 2, D, D)).astype(np.float32)
    f = lambda b: np.tensordot(C, b, axes=([1, 2, 3, 4], [1, 2, 3, 4]))
    assert rel(f(b1 + np.float32(0.3) * b2), f(b1) + np.float32(0.3) * f(b2)) < 1e-5


def test_split_properties():
    rng = np.random.default_rng(5)
    bond = rng.normal(size=(10, 2, 2, 9, 7)).astype(np.float32)
    for max_size, expect in ((5, 5), (50, 18)):
        a, a1 = o.bond_decomposition(bond, o.SweepParams(max_size=max_size, min_singular_value=1e-4))
        m = a.shape[2]
        assert m == expect
        u = a.reshape(-1, m)
        assert np.allclose(u.T @ u, np.eye(m), atol=1e-5)
        flat = np.transpose(bond, (1, 3, 0, 2, 4)).reshape(18, -1)
        uu, ss, vv = np.linalg.svd(flat.astype(np.float64), full_matrices=False)
        trunc = (uu[:, :m] * ss[:m]) @ vv[:m]
        assert rel(u @ a1.reshape(m, -1), trunc) < 1e-5
    low = np.zeros_like(bond)      # rank-deficient bond -> clamped to min_size
    a, a1 = o.bond_decomposition(low, o.SweepParams(max_size=30))
    assert a.shape[2] == 3


def test_armijo_halves_and_reverts():
    rng = np.random.default_rng(6)
    B, L, D = 50, 10, 4
    C = rng.normal(size=(B, 2, 2, D, D)).astype(np.float32)
    lab = np.eye(L, dtype=np.float32)[rng.integers(0, L, B)]
    bond = (0.1 * rng.normal(size=(L, 2, 2, D, D))).astype(np.float32)
    p = o.SweepParams(max_size=8)
    tr = o.Trace()
    new = o.update_bond(bond, C, lab, 1e3, p, tr)          # far too large a rate: needs several halvings
    assert tr[0]['trials'] > 1 and tr[0]['cost1'] <= tr[0]['cost0']
    tr2 = o.Trace()
    kept = o.update_bond(bond, C, lab, 1e30, p, tr2)         # never acceptable within 19 trials -> bond kept
    assert tr2[0]['trials'] == 19 and not tr2[0]['accepted'] and np.array_equal(kept, bond)


def test_learning_rate_schedule_and_checkpoint_bytes():
    assert abs(o.learning_rate(1.0, 0.99, 0) - 10.0) < 1e-9       # baseoptimizer.py:109
    nodes, loc = o.initial_nodes(5, 2, 10)
    raw = o.save_bytes(nodes, loc, 2, 10)
    assert raw[0] == 1 and int.from_bytes(raw[1:3], 'big') == 5 and int.from_bytes(raw[3:5], 'big') == loc
    back = o.load_bytes(raw)
    assert all(np.array_equal(a, b) for a, b in zip(back['nodes'], nodes))
    assert len(raw) == 9 + 2 * 5 + 4 * sum(n.size for n in nodes)


def test_feature_maps():
    x = np.array([[0.0, 0.5, 1.0]], dtype=np.float32)
    assert np.allclose(o.feature_map(x, o.FM_ONE_SQRT3)[0], [[1, -np.sqrt(3)], [1, 0], [1, np.sqrt(3)]], atol=1e-6)
    assert np.allclose(o.feature_map(x, o.FM_ONE_X)[0], [[1, 0], [1, 0.5], [1, 1]])
    cs = o.feature_map(x, o.FM_COS_SIN)[0]
    assert np.allclose(cs, [[1, 0], [np.sqrt(0.5), np.sqrt(0.5)], [0, 1]], atol=1e-6)
    assert np.allclose((cs ** 2).sum(-1), 1, atol=1e-6)


def test_datasource_oracle_matches_reference_preprocessing_code():
    """o.preprocess_images / o.next_batch against what the reference's own _preprocess_images (MNISTpreprocessing.py:14-78,
    run under the TF shim on a stand-in for the TF loader) returned: bit for bit, wrap-around included."""
    g = np.load(os.path.join(GOLDEN_DIR, "ref_preprocess.npz"))
    images = o.scale_uint8(g["images_u8"])
    size = int(g["size"])
    rows = np.arange(size) % len(images)
    for shrink, key in ((True, "phi1_shrink"), (False, "phi1_full")):
        phi = o.preprocess_images(images, 28, 28, shrink, o.FM_ONE_SQRT3)
        assert phi.dtype == np.float32 and np.all(phi[..., 0] == 1)
        assert np.array_equal(phi[rows][..., 1], g[key])
        feat, lab, nxt = o.next_batch(phi, g["labels"], 0, size)
        assert np.array_equal(feat[..., 1].T, g[key]) and np.array_equal(lab, g["labels_out"])
        assert nxt == size % len(images)
    pooled = o.preprocess_images(images, 28, 28, True, o.FM_PRECOMPUTED)
    assert np.allclose(pooled, images.reshape(-1, 14, 2, 14, 2).mean(axis=(2, 4)).reshape(-1, 196), atol=1e-7)


def test_next_batch_oracle_follows_the_reference_datasource():
    """o.next_batch against the mirror of MPSDatasource.next_training_data_batch (preprocessing.py:165-208) over a
    sequence of batch sizes that wraps, fits exactly and exceeds the dataset."""
    from preprocessing.preprocessing import MPSDatasource

    class DS(MPSDatasource):
        _use_cache_dir = False

    rng = np.random.default_rng(0)
    data, labels = rng.normal(size=(23, 6, 2)).astype(np.float32), np.eye(4)[rng.integers(0, 4, 23)]
    ds = DS((6, 2), False, _training_data=(data, labels), _test_data=(data, labels))
    start = 0
    for batch in (5, 18, 10, 23, 50, 1):
        got_f, got_l = ds.next_training_data_batch(batch)
        want_f, want_l, start = o.next_batch(data, labels, start, batch)
        assert np.array_equal(got_f, want_f) and np.array_equal(got_l, want_l)
        assert ds.current_index % 23 == start


def test_fashion_oracle_matches_reference_loader_code():
    """o.preprocess_fashion against FashionMNISTDatasource._load_with_correct_shape of the reference
    (FashionMNISTpreprocessing.py:56-77, run on seeded idx files): /255 in float64, 2x2 maximum, [1, x]."""
    g = np.load(os.path.join(GOLDEN_DIR, "ref_fashion.npz"))
    for shrink, key in ((True, "phi1_shrink"), (False, "phi1_full")):
        phi = o.preprocess_fashion(g["images_u8"], 28, 28, shrink)
        assert phi.dtype == np.float64 and np.all(phi[..., 0] == 1)
        assert np.array_equal(phi[..., 1].astype(np.float32), g[key])
    assert np.array_equal(g["labels_out"].argmax(1), g["labels_u8"])
    # the float32 division the kernel uses is the same function of the byte as float64 division followed by the cast
    u = np.arange(256)
    assert np.array_equal((u / 255.0).astype(np.float32), u.astype(np.float32) / np.float32(255))
    assert np.any((u / 255.0).astype(np.float32) != o.scale_uint8(u.astype(np.uint8)))       # and differs from the TF loader's


@pytest.mark.parametrize("name,loss", [("ref_linreg_xent", o.SOFTMAX_XENT), ("ref_linreg_squared", o.SQUARED)])
def test_lin_reg_oracle_matches_reference_code_run_under_tf_shim(name, loss):
    """tests/golden/ref_linreg_*.npz: the reference's own MPS._lin_reg / sqMPS._cost_for_lin_reg (mps.py:327-399,
    squaredDistanceMPS.py:29-30) fed by its own MPSDatasource.next_training_data_batch, with tf.train.GradientDescentOptimizer
    standing for a float64 central-difference derivative of the reference's cost graph -- so the oracle's closed-form
    gradient, the batch of 100, the wrap-around and the weight layout (N, d-1, L) are pinned by reference code."""
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, L, total, iters, lr = int(g["cfg"][0]), int(g["cfg"][1]), int(g["cfg"][2]), int(g["cfg"][3]), float(g["cfg"][4])
    assert total < iters * 100                                          # the set is read more than once (wrap-around)
    W, b, nxt = o.lin_reg(g["data"], g["labels"], 0, iters, lr, loss)
    assert g["weight"].shape == (N, 1, L) and nxt == int(g["next_index"])
    assert np.abs(W - g["weight"][:, 0, :]).max() < 2e-7 * np.abs(W).max()        # float32 storage of the reference's variables
    assert np.abs(b - g["bias"]).max() < 2e-7 * max(np.abs(b).max(), np.abs(W).max())
    site = int(np.unravel_index(g["weight"].argmax(), g["weight"].shape)[0])     # mps.py:392-399
    assert int(g["loc"]) == min(max(site, 1), N - 2)
