"""-m gpu parity tests of the datasource-side kernels (csrc/preprocess.cu) against the oracle and against what the
reference's own _preprocess_images produced (tests/golden/ref_preprocess.npz).  Byte/index work and the float32
arithmetic of pooling and of the linear feature maps are bit-exact; the cos/sin map (not in the reference) is held to
1e-6 absolute."""
import gzip
import os

import numpy as np
import pytest
import torch

import trmps_oracle as o
import trmps_native as nat
import trmps_device as dev

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_preprocess.npz")


def seeded_images(count, rows=28, cols=28, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (count, rows * cols), dtype=np.uint8)


def want_batch(images_f32, rows, cols, shrink, start, B, kind):
    """oracle: preprocess, then the batch slice with wrap-around and swapped axes"""
    data = o.preprocess_images(images_f32, rows, cols, shrink, kind)
    feat, _, _ = o.next_batch(data, np.zeros((len(data), 1)), start, B)
    return np.ascontiguousarray(feat)


def test_matches_the_reference_codes_own_output():
    g = np.load(GOLDEN)
    images, size = g["images_u8"], int(g["size"])
    dev_images = torch.from_numpy(images).cuda()
    for shrink, key in ((True, "phi1_shrink"), (False, "phi1_full")):
        got = dev.preprocess_batch(dev_images, 28, 28, shrink, 0, size, out_map=nat.FM_ONE_SQRT3).cpu().numpy()
        assert got.shape == (g[key].shape[1], size, 2)
        assert np.all(got[..., 0] == 1)
        assert np.array_equal(got[..., 1].T, g[key])                 # bit-exact, including the wrapped-around samples


def test_fashion_matches_the_reference_codes_own_output():
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "ref_fashion.npz"))
    d = torch.from_numpy(g["images_u8"]).cuda()
    n = len(g["images_u8"])
    for pool, key in ((nat.POOL_MAX, "phi1_shrink"), (nat.POOL_NONE, "phi1_full")):
        got = dev.preprocess_batch(d, 28, 28, pool, 0, n, out_map=nat.FM_ONE_X, u8_divide=True).cpu().numpy()
        assert np.all(got[..., 0] == 1)
        assert np.array_equal(got[..., 1].T, g[key])


@pytest.mark.parametrize("kind", [o.FM_PRECOMPUTED, o.FM_ONE_X])
@pytest.mark.parametrize("pool", [nat.POOL_MAX, nat.POOL_NONE])
def test_fashion_pipeline_matches_oracle(pool, kind):
    total, start, B = 257, 100, 700
    u8 = seeded_images(total, seed=13)
    got = dev.preprocess_batch(torch.from_numpy(u8).cuda(), 28, 28, pool, start, B, out_map=kind, u8_divide=True).cpu().numpy()
    phi = o.preprocess_fashion(u8, 28, 28, pool == nat.POOL_MAX).astype(np.float32)
    data = phi[..., 1] if kind == o.FM_PRECOMPUTED else phi
    want, _, _ = o.next_batch(data, np.zeros((total, 1)), start, B)
    assert np.array_equal(got, want)
    f32 = (u8 / 255.0).astype(np.float32)                            # float32 input takes the same pooling path
    got32 = dev.preprocess_batch(torch.from_numpy(f32).cuda(), 28, 28, pool, start, B, out_map=kind).cpu().numpy()
    assert np.array_equal(got32, want)


@pytest.mark.parametrize("shrink", [True, False])
def test_fashion_datasource_from_idx_files(tmp_path, monkeypatch, shrink):
    from trmps import FashionMNISTDatasource
    monkeypatch.chdir(tmp_path)
    os.mkdir("FashionMNISTDatasource")
    rng = np.random.default_rng(8)
    n_train, n_test = 300, 90
    tr, te = seeded_images(n_train, seed=16), seeded_images(n_test, seed=17)
    trl, tel = rng.integers(0, 10, n_train).astype(np.uint8), rng.integers(0, 10, n_test).astype(np.uint8)
    trl[0] = tel[0] = 9
    for name, arr, dims, magic in (("train-images-idx3-ubyte.gz", tr, (n_train, 28, 28), 2051), ("train-labels-idx1-ubyte.gz", trl, (n_train,), 2049),
                                   ("t10k-images-idx3-ubyte.gz", te, (n_test, 28, 28), 2051), ("t10k-labels-idx1-ubyte.gz", tel, (n_test,), 2049)):
        write_idx("FashionMNISTDatasource", name, arr, dims, magic)
    ds = FashionMNISTDatasource(shrink=shrink)
    data, labels = ds._training_data
    assert np.array_equal(data, o.preprocess_fashion(tr, 28, 28, shrink).astype(np.float32))
    assert labels.dtype.kind == "i" and np.array_equal(labels.argmax(1), trl)
    assert np.array_equal(ds._test_data[0], o.preprocess_fashion(te, 28, 28, shrink).astype(np.float32))
    feat, lab = ds.next_training_data_batch(64)
    assert feat.shape == ((196 if shrink else 784), 64, 2)
    os.remove(os.path.join("FashionMNISTDatasource", "train-images-idx3-ubyte.gz"))
    assert np.array_equal(FashionMNISTDatasource(shrink=shrink)._training_data[0], data)      # from the .npy cache
    os.remove(os.path.join("FashionMNISTDatasource", "training_data.npy"))
    with pytest.raises(Exception, match="Please place the MNIST Data"):
        FashionMNISTDatasource(shrink=shrink)


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("shrink", [True, False])
@pytest.mark.parametrize("kind", [o.FM_PRECOMPUTED, o.FM_ONE_SQRT3, o.FM_ONE_X, o.FM_COS_SIN])
def test_preprocess_batch_matches_oracle(dtype, shrink, kind):
    total, start, B = 300, 170, 421                                  # ragged tile, wraps twice
    u8 = seeded_images(total, seed=3)
    f32 = o.scale_uint8(u8)
    src = u8 if dtype == np.uint8 else f32
    got = dev.preprocess_batch(torch.from_numpy(src).cuda(), 28, 28, shrink, start, B, out_map=kind).cpu().numpy()
    want = want_batch(f32, 28, 28, shrink, start, B, kind)
    assert got.shape == want.shape
    if kind == o.FM_COS_SIN:
        assert np.max(np.abs(got - want)) < 1e-6
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("rows,cols", [(2, 4), (6, 8), (14, 64), (3, 4), (5, 3), (1, 1)])
def test_other_image_shapes(rows, cols):
    total, B = 37, 130
    rng = np.random.default_rng(rows * cols)
    images = rng.random((total, rows * cols), dtype=np.float32)
    for shrink in ((True, False) if rows % 2 == 0 else (False,)):
        got = dev.preprocess_batch(torch.from_numpy(images).cuda(), rows, cols, shrink, 5, B).cpu().numpy()
        assert np.array_equal(got, want_batch(images, rows, cols, shrink, 5, B, o.FM_PRECOMPUTED))
    u8 = (images * 255).astype(np.uint8)                             # rows that are not 4-byte aligned
    got = dev.preprocess_batch(torch.from_numpy(u8).cuda(), rows, cols, False, 36, B, out_map=nat.FM_ONE_X).cpu().numpy()
    assert np.array_equal(got, want_batch(o.scale_uint8(u8), rows, cols, False, 36, B, o.FM_ONE_X))


def test_full_size_mnist_set():
    """60 000 uint8 images, both the pooled and the full-resolution pass, against the oracle (NumPy does this in a second)."""
    u8 = seeded_images(60000, seed=9)
    f32 = o.scale_uint8(u8)
    d = torch.from_numpy(u8).cuda()
    for shrink in (True, False):
        got = dev.preprocess_batch(d, 28, 28, shrink, 0, 60000).cpu().numpy()
        assert np.array_equal(got, o.preprocess_images(f32, 28, 28, shrink, o.FM_PRECOMPUTED).T)


def test_empty_batch_and_bad_arguments():
    d = torch.from_numpy(seeded_images(8)).cuda()
    assert dev.preprocess_batch(d, 28, 28, True, 0, 0).shape == (196, 0)
    lib = nat.load()
    out = torch.empty(196 * 8, device="cuda")
    bad = [
        (d.data_ptr(), nat.DT_U8, 8, 28, 28, 1, 8, 4, 0, out.data_ptr(), None),      # start == total
        (d.data_ptr(), 7, 8, 28, 28, 1, 0, 4, 0, out.data_ptr(), None),              # dtype
        (d.data_ptr(), nat.DT_U8, 8, 28, 30, 1, 0, 4, 0, out.data_ptr(), None),      # cols % 4 with shrink
        (d.data_ptr(), nat.DT_U8, 8, 27, 28, 1, 0, 4, 0, out.data_ptr(), None),      # odd rows with shrink
        (d.data_ptr(), nat.DT_U8, 8, 28, 28, 1, 0, 4, 9, out.data_ptr(), None),      # feature map
        (d.data_ptr(), nat.DT_U8, 8, 28, 28, 3, 0, 4, 0, out.data_ptr(), None),      # pool code
        (None, nat.DT_U8, 8, 28, 28, 1, 0, 4, 0, out.data_ptr(), None),              # null images
    ]
    for args in bad:
        assert lib.trmps_preprocess_batch(*args) == -1
        assert "trmps_preprocess_batch" in nat.last_error()
    with pytest.raises(ValueError):
        dev.preprocess_batch(d.cpu(), 28, 28, True, 0, 4)


@pytest.mark.parametrize("c", [1, 2])
@pytest.mark.parametrize("N,total,start,B", [(196, 500, 0, 500), (196, 500, 431, 700), (10, 33, 32, 5), (7, 40, 1, 300), (784, 129, 7, 300)])
def test_batch_gather_matches_next_training_data_batch(c, N, total, start, B):
    rng = np.random.default_rng(N + B)
    data = rng.normal(size=(total, N, c) if c == 2 else (total, N)).astype(np.float32)
    labels = np.eye(10, dtype=np.float32)[rng.integers(0, 10, total)]
    feat, lab = dev.batch_gather(torch.from_numpy(data).cuda(), torch.from_numpy(labels).cuda(), start, B)
    want_f, want_l, _ = o.next_batch(data, labels, start, B)
    assert np.array_equal(feat.cpu().numpy(), want_f)
    assert np.array_equal(lab.cpu().numpy(), want_l)
    only_feat, none = dev.batch_gather(torch.from_numpy(data).cuda(), None, start, B)
    assert none is None and np.array_equal(only_feat.cpu().numpy(), want_f)


def test_device_resident_datasource_hands_out_the_same_batches():
    from trmps import SyntheticMNISTDatasource
    host = SyntheticMNISTDatasource(shrink=True, n_train=230, n_test=10, seed=4)
    devd = SyntheticMNISTDatasource(shrink=True, n_train=230, n_test=10, seed=4).to_device()
    for batch in (100, 100, 30, 230, 77, 500, 0 + 1):
        hf, hl = host.next_training_data_batch(batch)
        df, dl = devd.next_training_data_batch(batch)
        assert df.is_cuda and np.array_equal(df.cpu().numpy(), hf) and np.array_equal(dl.cpu().numpy(), hl)
        assert devd.current_index == host.current_index


def write_idx(dirname, name, payload, dims, magic):
    header = magic.to_bytes(4, "big") + b"".join(int(x).to_bytes(4, "big") for x in dims)
    with gzip.open(os.path.join(dirname, name), "wb") as fh:
        fh.write(header + payload.tobytes())


@pytest.mark.parametrize("shrink", [True, False])
def test_mnist_datasource_from_idx_files(tmp_path, monkeypatch, shrink):
    """MNISTDatasource on synthetic idx files: same arrays as the oracle pipeline (loader scaling, validation split,
    wrap-around to 60 000 samples, pooling, feature map), cached as .npy like the reference does."""
    from trmps import MNISTDatasource
    monkeypatch.chdir(tmp_path)
    os.mkdir("MNIST_data")
    rng = np.random.default_rng(5)
    n_train, n_test = 5600, 700                      # after the 5000-image validation split: 600 training images
    tr, te = seeded_images(n_train, seed=6), seeded_images(n_test, seed=7)
    trl, tel = rng.integers(0, 10, n_train).astype(np.uint8), rng.integers(0, 10, n_test).astype(np.uint8)
    write_idx("MNIST_data", "train-images-idx3-ubyte.gz", tr, (n_train, 28, 28), 2051)
    write_idx("MNIST_data", "train-labels-idx1-ubyte.gz", trl, (n_train,), 2049)
    write_idx("MNIST_data", "t10k-images-idx3-ubyte.gz", te, (n_test, 28, 28), 2051)
    write_idx("MNIST_data", "t10k-labels-idx1-ubyte.gz", tel, (n_test,), 2049)
    ds = MNISTDatasource(shrink=shrink)
    data, labels = ds._training_data
    n = 196 if shrink else 784
    assert data.shape == (60000, n, 2) and labels.shape == (60000, 10)
    want = o.preprocess_images(o.scale_uint8(tr[5000:]), 28, 28, shrink, o.FM_ONE_SQRT3)
    rows = np.arange(60000) % 600
    assert np.array_equal(data, want[rows])
    assert np.array_equal(labels.argmax(1), trl[5000:][rows])
    tdata, tlabels = ds._test_data
    assert tdata.shape == (10000, n, 2)
    assert np.array_equal(tdata, o.preprocess_images(o.scale_uint8(te), 28, 28, shrink, o.FM_ONE_SQRT3)[np.arange(10000) % n_test])
    feat, lab = ds.next_training_data_batch(50)
    assert feat.shape == (n, 50, 2)
    assert os.path.isfile(os.path.join("MNISTDatasource", "training_data.npy"))
    again = MNISTDatasource(shrink=shrink)                            # second construction reads the cache
    assert np.array_equal(again._training_data[0], data)
    with pytest.raises(FileNotFoundError):
        MNISTDatasource(shrink=not shrink, data_dir="nowhere")


# --------------------------------------------------------------------------------------------- linear pre-training
@pytest.mark.parametrize("loss", [o.SOFTMAX_XENT, o.SQUARED])
@pytest.mark.parametrize("N,total,start,iters", [(20, 230, 0, 57), (196, 1000, 950, 200), (784, 400, 3, 40), (300, 128, 100, 25)])
def test_linreg_pretrain_matches_oracle(loss, N, total, start, iters):
    """trmps_linreg_pretrain (all iterations in one launch) against the oracle restatement of MPS._lin_reg
    (mps.py:331-399) in fp64, with minibatches that wrap around the dataset and chains longer than one 256-site chunk."""
    rng = np.random.default_rng(N + iters)
    L = 10
    cls = rng.integers(0, L, total)
    pix = rng.random((total, N), dtype=np.float32)
    pix[np.arange(total), cls % N] += 1.0                              # something to learn
    data = o.feature_map_f32(pix, o.FM_ONE_SQRT3)                      # (total, N, 2)
    labels = np.eye(L, dtype=np.float32)[cls]
    lr = 0.05 if loss == o.SOFTMAX_XENT else 2e-4
    W, b = dev.linreg_pretrain(torch.from_numpy(data).cuda(), torch.from_numpy(labels).cuda(), start, iters, lr, loss)
    wantW, wantb, _ = o.lin_reg(data, labels, start, iters, lr, loss)
    assert np.isfinite(wantW).all() and np.abs(wantW).max() > 1e-3
    assert rel(W.cpu().numpy(), wantW) < 1e-4 and rel(b.cpu().numpy(), wantb) < 1e-4
    W32, b32, _ = o.lin_reg(data, labels, start, iters, lr, loss, dtype=np.float32)
    assert rel(W.cpu().numpy(), wantW) < max(20 * rel(W32, wantW), 2e-6)       # no worse than fp32 NumPy


@pytest.mark.parametrize("name,loss", [("ref_linreg_xent", o.SOFTMAX_XENT), ("ref_linreg_squared", o.SQUARED)])
def test_linreg_pretrain_matches_reference_code_golden(name, loss):
    """trmps_linreg_pretrain and MPS.prepare(data_source) against what the reference's own MPS._lin_reg produced under the
    shim (tests/golden/ref_linreg_*.npz, oracle/make_golden_from_reference.py::run_linreg_case)."""
    from trmps import MPS, sqMPS, MPSDatasource
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    N, L, iters, lr = int(g["cfg"][0]), int(g["cfg"][1]), int(g["cfg"][3]), float(g["cfg"][4])
    data, labels = g["data"].astype(np.float32), g["labels"].astype(np.float32)
    W, b = dev.linreg_pretrain(torch.from_numpy(data).cuda(), torch.from_numpy(labels).cuda(), 0, iters, lr, loss)
    assert rel(W.cpu().numpy(), g["weight"][:, 0, :]) < 2e-6 and rel(b.cpu().numpy(), g["bias"]) < 2e-6
    class InMemory(MPSDatasource):
        _use_cache_dir = False
    ds = InMemory(_training_data=(data, labels), _test_data=(data[:40], labels[:40]))
    net = (sqMPS if loss == o.SQUARED else MPS)(2, L, N)
    net.prepare(ds, iterations=iters, learning_rate=lr)
    assert rel(net.weight, g["weight"]) < 2e-6 and rel(net.bias, g["bias"]) < 2e-6
    assert net._special_node_loc == int(g["loc"]) and ds.current_index == int(g["next_index"])


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_prepare_with_datasource_follows_the_reference_semantics():
    """MPS.prepare(data_source): zero iterations leave a zero model; the read position of the datasource advances by
    100 * iterations like the reference's loop; the special node location comes from the largest weight (mps.py:390-399)."""
    from trmps import MPS, sqMPS, SyntheticMNISTDatasource
    ds = SyntheticMNISTDatasource(shrink=True, n_train=730, n_test=50, seed=3)
    net = MPS(2, 10, 196)
    net.prepare(ds, iterations=37, learning_rate=0.05)
    assert ds.current_index == (37 * 100) % 730
    data, labels = ds._training_data
    W, b, _ = o.lin_reg(data, labels, 0, 37, 0.05)
    assert rel(net.weight[:, 0, :], W) < 1e-4 and rel(net.bias, b) < 1e-4
    site = int(np.unravel_index(net.weight.argmax(), net.weight.shape)[0])
    assert net._special_node_loc == min(max(site, 1), 194)
    # continuing from the advanced read position, on a device-resident datasource, squared loss
    ds.to_device()
    sq = sqMPS(2, 10, 196)
    sq.prepare(ds, iterations=5, learning_rate=1e-4)
    W2, b2, nxt = o.lin_reg(data, labels, (37 * 100) % 730, 5, 1e-4, o.SQUARED)
    assert rel(sq.weight[:, 0, :], W2) < 1e-4 and ds.current_index == nxt
    zero = MPS(2, 10, 196)
    zero.prepare(ds, iterations=0)
    assert not zero.weight.any() and not zero.bias.any()
