"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/trmps_b200.h declares,
buffer-layout arithmetic, slot packing, the datasource and checkpoint mirrors, and the bench reference arm."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import trmps_native as nat
import trmps_device as dev
import trmps_oracle as o

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    text = open(os.path.join(ROOT, "include", "trmps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trmps_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(nat.exported_symbols()) == declared       # the ctypes table and the header agree
    assert lib.trmps_abi_version() == 1


def test_struct_sizes_match_the_header_layout():
    import ctypes as C
    assert C.sizeof(nat.Opt) == 48
    assert C.sizeof(nat.Layout) == 20 * 8 + 6 * 4
    assert C.sizeof(nat.Sweep) == 4 * 4 + 2 * 8 + 2 * 4 + 10 * 8 + 8
    assert C.sizeof(nat.PredictDesc) == 4 * 4 + 8 + 4 * 4 + 6 * 8


def test_no_gpu_is_reported_not_faked():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = nat.load()
    assert lib.trmps_device_check(0) == -3
    assert "CUDA" in nat.last_error()
    with pytest.raises(RuntimeError):
        nat.require_gpu(0)
    from mps.mps import MPS
    net = MPS(2, 10, 8)
    net.prepare()
    with pytest.raises(RuntimeError):           # the product path fails loudly: no CPU fallback
        net.predict(np.zeros((8, 4, 2), np.float32))


def test_layout_arithmetic():
    lay = nat.sweep_layout(196, 60000, 10, 120, 148)
    R, Cc = 240, 2400
    assert lay.gradM - lay.bondM == R * Cc
    assert lay.gpart - lay.hessM >= R * Cc
    assert lay.f0 - lay.gpart == max(lay.gsplit, lay.gsplit_tc) * R * Cc
    assert lay.gsplit == 7 and lay.gsplit_tc == 14 and lay.svd_block == 8 and 1 <= lay.nsplit <= 8
    assert lay.total - lay.bimg >= 30 * 8 * 2 * 16 * 160      # hi/lo bond images of the tensor-core forward
    assert lay.total * 4 < 240e6
    for off in ("bondM", "gradM", "deltaM", "hessM", "gpart", "f0", "fd", "fpart", "resid", "sres", "hres", "phi2", "Ut", "sv", "red", "scal", "bimg", "split_ws"):
        assert getattr(lay, off) % 4 == 0
    with pytest.raises(RuntimeError):
        nat.sweep_layout(10, 100, 10, 20, 148)      # Ds must be a multiple of 8
    with pytest.raises(RuntimeError):
        nat.sweep_layout(10, 100, 17, 40, 148)      # d_output <= 16
    small = nat.sweep_layout(8, 100, 3, 8, 148)
    assert small.gsplit == 1


def test_pack_unpack_roundtrip_and_padding():
    nodes = o.random_full_bond_mps(7, 2, 10, 13, loc=3, seed=0)
    Ds = dev.bond_stride(nodes, 3, 10, 0)
    assert Ds == 24
    slots, special, dims = dev.pack_nodes(nodes, 3, 10, Ds)
    assert dims.tolist() == [20, 13, 13, 13, 13, 13, 13, 20]
    assert slots.shape == (7, 2, 24, 24) and special.shape == (10, 2, 24, 24)
    assert np.all(slots[1, :, 13:, :] == 0) and np.all(slots[1, :, :, 13:] == 0)
    back = dev.unpack_nodes(slots, special, dims, 3)
    assert all(np.array_equal(a, b) for a, b in zip(back, nodes))
    with pytest.raises(ValueError):
        dev.pack_nodes(nodes[:3] + [nodes[4]] * 4, 3, 10, Ds)
    assert dev.bond_stride(nodes, 3, 10, 120) == 120 and dev.bond_stride(nodes, 3, 10, 30) == 32


def test_mps_mirror_matches_oracle_construction(tmp_path):
    from mps.mps import MPS
    from mps.squaredDistanceMPS import sqMPS
    net = MPS(2, 10, 12)
    net.prepare()
    ref_nodes, loc = o.initial_nodes(12, 2, 10)
    assert net._special_node_loc == loc == 6
    assert all(np.array_equal(a, b) for a, b in zip(net.nodes_list, ref_nodes))
    assert np.array_equal(net.start_node, o.start_vector(10)) and np.array_equal(net.end_node, o.end_vector(10))
    path = str(tmp_path / "MPSconfig")
    net.save(net.nodes_list, path=path)
    assert open(path, "rb").read() == o.save_bytes(ref_nodes, loc, 2, 10)       # byte-compatible with mps.py:126-160
    again = MPS.from_file(path)
    assert again._special_node_loc == loc and all(np.array_equal(a, b) for a, b in zip(again.nodes_list, ref_nodes))
    net.save(None, path=path)
    assert len(open(path, "rb").read()) == 9
    # metrics agree with the oracle's
    rng = np.random.default_rng(0)
    f = rng.normal(size=(40, 10)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 40)]
    assert abs(net.cost(f, y) - o.cost_softmax_xent(f, y)) < 1e-6
    assert abs(sqMPS(2, 10, 12).cost(f, y) - o.cost_squared(f, y)) < 1e-6
    assert net.accuracy(f, y) == o.accuracy(f, y)
    cm = net.confusion_matrix(f, y)
    assert np.array_equal(cm, o.confusion_matrix(f, y, 10)) and abs(net.f1score(f, y, cm) - o.f1score(cm)) < 1e-9
    with pytest.raises(ValueError):
        MPS(3, 10, 12)
    with pytest.raises(ValueError):
        MPS(2, 10, 12).prepare(_weights=ref_nodes)         # special_node_loc must be given, as in the reference


def test_datasource_batching_wraps_like_the_reference():
    from preprocessing.preprocessing import MPSDatasource

    class Plain(MPSDatasource):
        _use_cache_dir = False

    x = np.arange(10 * 3 * 2, dtype=np.float32).reshape(10, 3, 2)
    y = np.eye(10, dtype=np.float32)
    ds = Plain((3, 2), False, _training_data=(x, y), _test_data=(x[:4], y[:4]))
    f, l = ds.next_training_data_batch(4)
    assert f.shape == (3, 4, 2) and np.array_equal(f, np.swapaxes(x[:4], 0, 1)) and ds.current_index == 4
    f, l = ds.next_training_data_batch(8)               # wraps: items 4..9 then 0..1
    assert np.array_equal(l, np.concatenate([y[4:], y[:2]])) and ds.current_index == 2
    f, l = ds.next_training_data_batch(23)              # wraps twice
    assert l.shape[0] == 23 and np.array_equal(l[:8], y[2:]) and ds.current_index == (2 + 23) % 10
    assert ds.num_train_samples == 10 and ds.num_test_samples == 4
    tf_, tl = ds.test_data
    assert tf_.shape == (3, 4, 2) and ds.training_data[0].shape == (3, 10, 2)


def test_synthetic_datasource_and_feature_map():
    from preprocessing.syntheticpreprocessing import SyntheticMNISTDatasource, mnist_feature_map
    ds = SyntheticMNISTDatasource(shrink=True, n_train=30, n_test=10, seed=3)
    f, l = ds.next_training_data_batch(30)
    assert f.shape == (196, 30, 2) and l.shape == (30, 10) and np.all(f[:, :, 0] == 1)
    pix = np.random.default_rng(0).random((5, 7), dtype=np.float32)
    assert np.allclose(mnist_feature_map(pix), o.feature_map(pix, o.FM_ONE_SQRT3))
    assert not os.path.isdir("SyntheticMNISTDatasource")


def test_optimizer_parameter_defaults_match_reference():
    from optimizers.parameterObjects import MPSOptimizerParameters, MPSTrainingParameters
    p, t = MPSOptimizerParameters(), MPSTrainingParameters()
    assert (p.cutoff, p.reg, p.lr_reg, p.min_singular_value, p.verbosity, p.armijo_coeff, p.use_hessian,
            p.armijo_iterations, p.path, p.updates_per_step) == (1000, 0.001, 0.99, 1e-4, 0, 1e-4, False, 10, "MPSconfig", 1)
    assert (t.rate_of_change, t.initial_weights, t.verbose_save, t._logging_enabled) == (1000, None, True, False)


def test_bench_reference_arm_prints_contract_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-images", "64", "--cpu-sample-bonds", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "sweep_images_bonds_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_checkpoint_fixture_parses_with_boundary_one():
    """The one trained-weights file the reference ships (trmps/generation/MPSconfig) was written by its generation module:
    a shortMPS whose outer bonds are 1, not 2L.  Its bytes follow MPS.save (mps.py:126-160): flag, int16 N / loc / d / L, N
    int16 right-bond dimensions, float32 C-order tensors.  Parsed here with boundary 1 (our from_file keeps the reference's
    boundary 2L, mps.py:80-124, under which the reference's own loader cannot read this file either): the byte count has to
    come out exactly, and the weights must be the mixed-canonical MPS SURVEY.md section 2 describes."""
    path = "/root/reference/trmps/generation/MPSconfig"
    if not os.path.isfile(path):
        pytest.skip("reference tree not present (GPU box)")
    raw = open(path, "rb").read()
    assert len(raw) == 731041
    be = lambda a, b: int.from_bytes(raw[a:b], "big")
    assert raw[0] == 1
    N, loc, d, L = be(1, 3), be(3, 5), be(5, 7), be(7, 9)
    assert (N, loc, d, L) == (196, 91, 2, 11)
    dims = [be(9 + 2 * i, 11 + 2 * i) for i in range(N)]
    assert dims[-1] == 1 and max(dims) == 40 and dims[0] == 2
    pos, nodes = 9 + 2 * N, []
    for i in range(N):
        dl, dr = (1 if i == 0 else dims[i - 1]), dims[i]
        n = dl * dr * d * (L if i == loc else 1)
        shape = (L, d, dl, dr) if i == loc else (d, dl, dr)
        nodes.append(np.frombuffer(raw[pos:pos + 4 * n], dtype="<f4").astype(np.float64).reshape(shape))     # the tensors are native (little-endian) float32, only the header is big-endian
        pos += 4 * n
    assert pos == len(raw)                                   # every byte accounted for
    for i, w in enumerate(nodes):
        if i < loc:        # left-canonical: sum_{m,a} A[m,a,b] A[m,a,b'] = delta
            g = np.einsum('mab,mac->bc', w, w)
        elif i > loc:      # right-canonical: sum_{m,b} A[m,a,b] A[m,a',b] = delta
            g = np.einsum('mab,mcb->ac', w, w)
        else:
            continue
        assert np.abs(g - np.eye(g.shape[0])).max() < 2e-5, i
    # the same bytes go through our writer unchanged (format parity, whatever the boundary)
    out = (True).to_bytes(1, "big") + b"".join(int(v).to_bytes(2, "big") for v in (N, loc, d, L)) + b"".join(int(v).to_bytes(2, "big") for v in dims)
    out += b"".join(np.ascontiguousarray(w, dtype="<f4").tobytes() for w in nodes)
    assert out == raw
