#!/usr/bin/env python
"""bench.py -- headline benchmark of the TrMPS hot path on B200.

Metric (BASELINE.json): sweep images*bonds/sec at N=196, D=120 (cfg3: 14x14 synthetic MNIST shape, d=2,
10 labels, 60k images per GPU, full-bond initial MPS, feature map [cos, sin]).  One "step" = one full
left-to-right half-sweep of N-1 = 195 two-site bond updates (merge, forward, gradient, Armijo, truncated
Jacobi SVD split, environment advance) including the initial environment build.

    python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference ...                     the reference arithmetic on the host CPU (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps"))

if "reference" in sys.argv:
    # the CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1, and the BLAS pools are sized when NumPy is
    # imported, so this has to happen first
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

# BASELINE.json configs (SURVEY.md section 8 shorthand).  init "full_bond": full-bond random MPS, every internal bond = D
# (fixed-D throughput runs); "reference": MPS.prepare() default = the linear model in bond dimension 2L
# (mps.py:211-219), bonds then grow to max_size = D as the sweep proceeds.  kind "l2r": one left-to-right half-sweep of
# N-1 bonds from special_node_loc = 0 (weights reset every step); "train": BaseOptimizer.train_step, 2(N-1) bond updates
# (baseoptimizer.py:213-238), weights carried from step to step with the environments rebuilt per step like the reference.
CONFIGS = {
    "cfg1": dict(N=196, D=10, B=1000, kind="train", init="reference", fm="ONE_SQRT3", lr=1e-2),
    "cfg2": dict(N=784, D=20, B=20000, kind="train", init="reference", fm="ONE_SQRT3", lr=1e-3),
    "cfg3": dict(N=196, D=120, B=60000, kind="l2r", init="full_bond", fm="COS_SIN", lr=1e-4),
    "cfg4": dict(N=784, D=64, B=60000, kind="train", init="reference", fm="ONE_SQRT3", lr=1e-3),
    "cfg5": dict(N=784, D=32, B=1000000, kind="inference", init="full_bond", fm="COS_SIN", lr=0.0),
}
L_OUT = 10
METRIC = "sweep_images_bonds_per_sec"
UNIT = "images*bonds/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    bf16_burst=p["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, bf16_sustained=1400.0, bf16_burst=1590.0, source="fallback")


# ------------------------------------------------------------------------------------------ synthetic workload
def synthetic_pixels(N, B, L, seed):
    rng = np.random.default_rng(seed)
    pixels = rng.random((N, B), dtype=np.float32)
    cls = rng.integers(0, L, B)
    seg = max(N // L, 1)
    for c in range(L):          # make the labels learnable: class c brightens its own segment of sites
        cols = np.nonzero(cls == c)[0]
        pixels[c * seg:(c + 1) * seg, cols] = 0.5 + 0.5 * pixels[c * seg:(c + 1) * seg, cols]
    labels = np.zeros((B, L), dtype=np.float32)
    labels[np.arange(B), cls] = 1
    return pixels, labels


COS_SIN_SCALE = 0.7893698     # exp(-E[log(cos(pi x/2) + sin(pi x/2))]), x ~ U[0,1): keeps the chain product O(1)


def full_bond_mps(N, d, L, D, loc, seed=1):
    """Full-bond random MPS for fixed-D runs: every internal bond = D, outer bonds 2L.  Each site
    tensor is COS_SIN_SCALE * identity on BOTH feature components plus N(0, (0.1/sqrt(D))^2) noise, so that with the
    [cos, sin] feature map the 196-site product neither under- nor overflows and the logits are O(1)."""
    rng = np.random.default_rng(seed)
    noise = 0.1 / np.sqrt(D)
    dims = [2 * L] + [D] * (N - 1) + [2 * L]
    nodes = []
    for i in range(N):
        dl, dr = dims[i], dims[i + 1]
        base = np.zeros((d, dl, dr))
        k = min(dl, dr)
        base[:, :k, :k] = np.eye(k) * COS_SIN_SCALE
        if i == loc:
            node = base[None] + rng.normal(0.0, noise, size=(L, d, dl, dr))
        else:
            node = base + rng.normal(0.0, noise, size=(d, dl, dr))
        nodes.append(node.astype(np.float32))
    return nodes


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------ CPU arm
class _LightTrace(list):
    """keeps the Armijo trial counts of the oracle's per-update records, drops their arrays"""

    def __init__(self):
        super().__init__()
        self.sites = []

    def append(self, rec):
        super().append(dict(trials=rec.get("trials", 0), accepted=rec.get("accepted", False)))

    def append_site(self, c, direction):
        self.sites.append((c, direction))


def benchmark_mps(cfg, loc, probe_predict, default_nodes):
    """Initial MPS of a sweep configuration, identical for the CPU and the GPU arm.  reference: the arm's own default
    prepare() (default_nodes()).  full_bond: the special node is rescaled so that the initial logits of a fixed 2048-image
    probe have unit spread (keeps the softmax un-saturated and the Armijo search in its working range);
    probe_predict(nodes, pixels) -> logits is the arm's own inference."""
    N, D = cfg["N"], cfg["D"]
    if cfg["init"] == "reference":
        return [np.asarray(w, dtype=np.float32) for w in default_nodes()], 1.0
    nodes = full_bond_mps(N, 2, L_OUT, D, loc)
    probe_pix, _ = synthetic_pixels(N, 2048, L_OUT, seed=4242)
    spread = float(np.std(probe_predict(nodes, probe_pix)))
    nodes[loc] = (nodes[loc] / max(spread, 1e-30)).astype(np.float32)
    return nodes, spread


def cpu_sweep_rate(cfg, B_s, bonds, lr, seed=100):
    """Times the oracle (literal restatement of the reference graph: materialised C, four contractions per update plus one
    per extra Armijo trial, LAPACK SVD; oracle/trmps_oracle.py) on the host cores: the SAME configuration as the GPU arm
    (same chain, same initial MPS, same feature map, same rate_of_change), bounded to a B_s-image sample and the first
    `bonds` bond updates of the sweep.  The environment build runs over the whole chain for the sample and is charged pro
    rata (bonds / (N-1)), as every bond update of the full sweep has one environment step behind it.
    Returns (images*bonds/s, seconds charged, detail dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import trmps_oracle as o
    N, D = cfg["N"], cfg["D"]
    fm = getattr(o, "FM_" + cfg["fm"])
    loc = 0 if cfg["kind"] == "l2r" else N // 2
    pix, lab = synthetic_pixels(N, B_s, L_OUT, seed)
    phi = o.feature_map(pix, fm)
    nodes, _ = benchmark_mps(cfg, loc, lambda nd, px: o.predict(nd, loc, o.feature_map(px, fm), L_OUT),
                             lambda: o.initial_nodes(N, 2, L_OUT, loc=loc)[0])
    p = o.SweepParams(max_size=D, min_singular_value=0.0)
    bonds = min(bonds, N - 1 - loc)
    t0 = time.perf_counter()
    C1s, C2s = o.setup_environments(nodes, loc, phi, L_OUT)
    t_env = time.perf_counter() - t0
    tr = _LightTrace()
    t0 = time.perf_counter()
    o.sweep_right(nodes, loc, loc + bonds, C1s, C2s, phi, lab, lr, p, tr)
    t_bonds = time.perf_counter() - t0
    charged = t_bonds + t_env * bonds / (N - 1)
    trials = [r["trials"] for r in tr]
    detail = dict(env_build_s=t_env, bond_updates_s=t_bonds, armijo_trials_per_update=float(np.mean(trials)) if trials else 0.0,
                  accepted=int(sum(r["accepted"] for r in tr)), updates=len(tr))
    return B_s * bonds / charged, charged, detail


def cpu_inference_rate(B_s, N, D, L, seed=0):
    """Times the oracle's MPS.predict (mps.py:521-560 restated: one batched chain step per site) on the host cores
    for a B_s-image sample.  Returns (images/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import trmps_oracle as o
    pix, _ = synthetic_pixels(N, B_s, L, seed)
    phi = o.feature_map(pix, o.FM_COS_SIN)
    nodes = full_bond_mps(N, 2, L, D, N // 2)
    o.predict(nodes, N // 2, phi[:, :64], L)
    t0 = time.perf_counter()
    o.predict(nodes, N // 2, phi, L)
    dt = time.perf_counter() - t0
    return B_s / dt, dt


def cpu_fetch_rate(B_s, seed=0):
    """The datasource side on the host: oracle average pool + batch slice + the swapaxes copy the reference makes per step."""
    import trmps_oracle as o
    images = np.random.default_rng(seed).random((B_s, 784), dtype=np.float32)
    t0 = time.perf_counter()
    pooled = o.preprocess_images(images, 28, 28, True, o.FM_PRECOMPUTED)
    feat, _, _ = o.next_batch(pooled, np.zeros((B_s, 1), np.float32), 0, B_s)
    np.ascontiguousarray(feat)
    dt = time.perf_counter() - t0
    return B_s / dt, dt


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(x.get("num_threads", 1)) for x in threadpool_info()] or [1])
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count()))


def workload_name(cfg_name, cfg, B_total, world, scaling):
    if cfg["kind"] == "inference":
        return "%s: inference N=%d D=%d L=10, %d images in total" % (cfg_name, cfg["N"], cfg["D"], B_total)
    what = ("one L->R half-sweep of %d bond updates from special_node_loc=0, environment build included; full-bond random MPS (m = D at "
            "every bond)" % (cfg["N"] - 1)) if cfg["kind"] == "l2r" else \
           ("one training step = %d bond updates (loc->N-1, N-1->0, 0->loc; baseoptimizer.py:213-238), environment build included; "
            "reference default initial MPS (linear model, bond 2L growing to max_size=D), weights carried between steps" % (2 * (cfg["N"] - 1)))
    return "%s: N=%d D=%d L=10 d=2, %d images in total %s over %d GPU(s); %s; feature map %s, min_sv=0" % (
        cfg_name, cfg["N"], cfg["D"], B_total, "sharded" if scaling == "strong" else "(weak: per-GPU batch fixed)", world, what, cfg["fm"])


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    if cfg["kind"] == "inference":
        rates, times = [], []
        for _ in range(max(args.warmup, 1)):
            cpu_inference_rate(64, cfg["N"], cfg["D"], L_OUT)
        for _ in range(args.steps):
            r, dt = cpu_inference_rate(args.cpu_sample_images, cfg["N"], cfg["D"], L_OUT)
            rates.append(r); times.append(dt)
        value = args.cpu_sample_images * len(times) / sum(times)
        sample = "%d images per step through the oracle's predict (NumPy chain)" % args.cpu_sample_images
        metric, unit, detail = "inference_images_per_sec", "images/s", {}
    else:
        lr = args.lr if args.lr > 0 else cfg["lr"]
        B_s, bonds = args.cpu_sample_images, args.cpu_sample_bonds
        for _ in range(args.warmup):
            cpu_sweep_rate(cfg, min(B_s, 128), 1, lr)
        times, detail = [], {}
        for _ in range(args.steps):
            r, dt, detail = cpu_sweep_rate(cfg, B_s, bonds, lr)
            times.append(dt)
        value = B_s * min(bonds, cfg["N"] - 1) * len(times) / sum(times)
        sample = ("%d images x the first %d bond updates of the sweep per step (same chain, initial MPS, feature map and rate_of_change=%g "
                  "as the GPU arm; environment build charged pro rata), literal reference graph in NumPy; Armijo trials per update %.1f"
                  % (B_s, bonds, lr, detail["armijo_trials_per_update"]))
        metric, unit = METRIC, UNIT
    cores = blas_threads()
    out = dict(impl="reference", metric=metric, value=value, unit=unit, n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling="strong",
               vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload=workload_name(args.config, cfg, cfg["B"], args.gpus, "strong"), sample=sample, **detail),
               cpu_baseline=dict(value=value, unit=unit, cores=cores, host_cores=os.cpu_count(), kind="port", sample=sample),
               e2e=dict(value=value, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def measure_tf32_peak(torch, device, seconds=0.6):
    """Dense TF32 GEMM throughput of this GPU (cuBLAS, 8192^3, back to back for `seconds`): the denominator of the tensor
    rooflines, measured in the same run instead of derived from the bf16 peak.  A library GEMM is used only here."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device=device); b = torch.randn((n, n), device=device)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, t_total, n_total = 8, 0.0, 0
        while t_total < seconds * 1e3:
            e0.record()
            for _ in range(reps):
                torch.matmul(a, b)
            e1.record(); torch.cuda.synchronize()
            t_total += e0.elapsed_time(e1); n_total += reps
        return 2.0 * n ** 3 * n_total / (t_total * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import trmps_native as nat
    import trmps_device as dev
    from trmps import MPS, MPSOptimizer, MPSOptimizerParameters

    nat.require_gpu(local_rank)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg_name = args.config
    cfg = CONFIGS[cfg_name]
    L = L_OUT
    N, D = cfg["N"], cfg["D"]
    fm = getattr(nat, "FM_" + cfg["fm"])
    lr = args.lr if args.lr > 0 else cfg["lr"]   # the gradient is a SUM over the batch (optimizer.py:249); Armijo halves from here
    B_total = args.images if args.images > 0 else cfg["B"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x, dtype=torch.float32):
        t = torch.tensor([x], device=device, dtype=dtype)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = measured_peaks()

    # ---------------------------------------------------------------------------------------- inference
    def inference_rate(Ni, Di, Bi, label):
        inodes = full_bond_mps(Ni, 2, L, Di, Ni // 2)
        w = dev.DeviceWeights(inodes, Ni // 2, L, dev.bond_stride(inodes, Ni // 2, L), device)
        gen = torch.Generator(device=device).manual_seed(7 + rank)
        ipix = torch.rand((Ni, Bi), device=device, generator=gen)
        iout = torch.empty((Bi, L), device=device)
        for _ in range(3):
            dev.predict(w, ipix, nat.FM_COS_SIN, out=iout)
        barrier()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        i0.record()
        for _ in range(reps):
            dev.predict(w, ipix, nat.FM_COS_SIN, out=iout)
        i1.record()
        barrier()
        ims = max_over_ranks(i0.elapsed_time(i1) / reps)
        iflops = Bi * ((Ni - 1) * 2.0 * 2 * Di * Di + 2.0 * L * 2 * Di * Di)
        return dict(metric="inference_images_per_sec", value=world * Bi / (ims * 1e-3), unit="images/s", ms_per_pass=ims,
                    config="%s: N=%d D=%d L=10, %d images per GPU, raw pixels resident in HBM, fused cos/sin map, "
                           "fp16 hi/lo split chain on tcgen05 (FP32-equivalent accuracy)" % (label, Ni, Di, Bi),
                    algorithmic_tflops=world * iflops / (ims * 1e-3) / 1e12,
                    pixel_stream_gbs=world * 4.0 * Ni * Bi / (ims * 1e-3) / 1e9)

    if cfg["kind"] == "inference":
        # cfg5: images sharded over the ranks, no collective; e2e = predict through the Python class with HOST pixels
        Bi = B_total // world
        tf32_peak = measure_tf32_peak(torch, device)
        l0 = nat.launch_count()
        res = inference_rate(N, D, Bi, cfg_name)
        launches = (nat.launch_count() - l0) * 5 // 8
        inodes = full_bond_mps(N, 2, L, D, N // 2)
        net = MPS(2, L, N, special_node_loc=N // 2)
        net.prepare(_weights=inodes)
        hb = min(Bi, 200000)
        hpix = torch.rand((N, hb)).pin_memory()
        net.predict(hpix, feature_map=nat.FM_COS_SIN)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            logits = net.predict(hpix, feature_map=nat.FM_COS_SIN)
            logits = logits.cpu().numpy() if hasattr(logits, "cpu") else np.asarray(logits)      # D2H of the result
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0, torch.float64)
        if rank == 0:
            frac = res["algorithmic_tflops"] / world / tf32_peak
            out = dict(metric="inference_images_per_sec", value=res["value"], unit="images/s", n_gpus=world, steps=5, warmup=3,
                       ms_per_step=res["ms_per_pass"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                       data="synthetic", config=dict(workload=workload_name(cfg_name, cfg, B_total, world, "strong"), images_per_gpu=Bi,
                                                     l2="pixel stream %.1f GB per GPU, larger than L2" % (4.0 * N * Bi / 1e9)),
                       gpu_launches=int(launches),
                       e2e=dict(value=world * hb * args.e2e_steps / e2e_s, unit="images/s", h2d_bytes_per_step=int(4 * N * hb),
                                d2h_bytes_per_step=int(4 * L * hb), steps=args.e2e_steps, images_per_step=hb),
                       roofline=dict(kernel="chain_tc_kernel<32> (fused on-chip inference chain, tcgen05 kind::f16 hi/lo)", bound="tensor",
                                     achieved=res["algorithmic_tflops"] / world, peak=tf32_peak, unit="TFLOP/s", frac=frac, traffic=None,
                                     peak_source="TF32 dense GEMM measured in this run (cuBLAS 8192^3)"))
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------------------------------- sweeps
    strong = args.scaling == "strong"
    B = B_total // world if strong else B_total
    loc = 0 if cfg["kind"] == "l2r" else N // 2
    bonds = (N - 1) if cfg["kind"] == "l2r" else 2 * (N - 1)

    def gpu_predict(nd, px):
        probe = MPS(2, L, N, special_node_loc=loc)
        probe.prepare(_weights=nd)
        return probe.predict(px, feature_map=fm)

    def gpu_default_nodes():
        ref = MPS(2, L, N, special_node_loc=loc)
        ref.prepare()                                     # mps.py:211-219: the linear model in bond dimension 2L
        return ref.nodes_list

    nodes, spread = benchmark_mps(cfg, loc, gpu_predict, gpu_default_nodes)

    def build(B_):
        net = MPS(2, L, N, special_node_loc=loc)
        net.prepare(_weights=nodes)
        params = MPSOptimizerParameters(min_singular_value=0.0, reg=1e-3, updates_per_step=1)
        optim = MPSOptimizer(net, D, params)
        optim.feature_map = fm
        optim.rate_of_change = lr
        if world > 1:
            optim.attach_process_group()
        pix, lab = synthetic_pixels(N, B_, L, seed=100 + rank)
        pix_pin = torch.from_numpy(pix).pin_memory()
        lab_pin = torch.from_numpy(lab).pin_memory()
        optim.bind_batch(pix_pin, lab_pin, weights=nodes)
        return optim, pix_pin, lab_pin

    def timed_steps(optim, steps, warmup, profile, graph=False):
        st = optim._state
        opt = optim._opt(lr)
        slots0, special0, dims0 = st.slots.clone(), st.special.clone(), st.dims.clone()

        def reset():
            st.slots.copy_(slots0); st.special.copy_(special0); st.dims.copy_(dims0)
            st.loc = loc

        def device_step():
            # inputs already resident in HBM
            if cfg["kind"] == "l2r":
                reset()
                st.init_env(loc)
                st.sweep(loc, N - 1, +1, opt)
            else:
                st.train_step(opt, loc=loc)

        for _ in range(warmup):
            device_step()
        run_step = device_step
        if graph:
            # the whole step (every kernel of the three half-sweeps, the NCCL all-reduces included) as ONE CUDA graph: possible
            # because no kernel of the sweep is a cooperative launch any more and nothing on the path synchronises with the host
            barrier()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                device_step()
            run_step = cg.replay
            run_step()
        if cfg["kind"] != "l2r":
            reset()                       # the timed steps start from the initial MPS again
        barrier()
        if profile:
            nat.profile_begin(400000)
        l0 = nat.launch_count()
        sweeps0 = int(st.iwork[nat.I_N_SWEEPS].item())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            run_step()
        ev1.record()
        barrier()
        launches = nat.launch_count() - l0
        prof = nat.profile_end() if profile else None
        ms_total = max_over_ranks(ev0.elapsed_time(ev1))
        jacobi_sweeps = int(st.iwork[nat.I_N_SWEEPS].item()) - sweeps0
        return ms_total, launches, prof, jacobi_sweeps

    optim, pix_pin, lab_pin = build(B)
    st = optim._state
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, prof, jacobi_sweeps = timed_steps(optim, args.steps, args.warmup, True)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * bonds * args.steps / (ms_total * 1e-3)
    iw = st.iwork.cpu().numpy()
    dims_end = st.dims.cpu().numpy()

    # the replicated MPS must be bit-identical on every rank after the timed sweeps (the split is replicated)
    def checksum():
        acc = torch.zeros((), dtype=torch.int64, device=device)
        for t in (st.slots, st.special):
            acc = acc + t.view(torch.int32).to(torch.int64).sum()
        return acc + st.dims.to(torch.int64).sum()

    cs = checksum()
    cs_min, cs_max = cs.clone(), cs.clone()
    if world > 1:
        dist.all_reduce(cs_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(cs_max, op=dist.ReduceOp.MAX)
    ranks_bit_identical = bool(cs_min.item() == cs_max.item())

    # end-to-end through the public Python API with HOST buffers: H2D of the batch and the weights, sweep, D2H of the weights
    def e2e_step():
        optim.bind_batch(pix_pin, lab_pin, weights=nodes)
        if cfg["kind"] == "l2r":
            optim._setup_optimization()
            optim._sweep_right(loc, N - 1)
        else:
            st.train_step(optim._opt(lr), loc=loc)
        out_nodes = optim.updated_nodes()
        cost = float(st.scalars()[nat.S_COST1].item())
        return out_nodes, cost

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        out_nodes, last_cost = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0, torch.float64)
    e2e_value = world * B * bonds * args.e2e_steps / e2e_s
    h2d = pix_pin.numel() * 4 + lab_pin.numel() * 4 + sum(w.nbytes for w in nodes)
    d2h = sum(w.nbytes for w in out_nodes) + 4
    Dp = st.Ds

    # secondary number: the same steps replayed from one CUDA graph (launch-bound configurations gain; cfg3 does not)
    graph_line = None
    if args.graph == "on" or (args.graph == "auto" and world == 1):
        try:
            gms, _, _, _ = timed_steps(optim, args.steps, 1, False, graph=True)
            graph_line = dict(value=world * B * bonds * args.steps / (gms * 1e-3), unit=UNIT, ms_per_step=gms / args.steps, steps=args.steps,
                              note="the same timed steps replayed from ONE captured CUDA graph per step (cudaGraphLaunch instead of %d kernel "
                                   "launches); `value` above is the plain stream-launch number the breakdown belongs to"
                                   % (launches // max(args.steps, 1)))
        except Exception as exc:                      # a failed capture must not cost the headline line
            graph_line = dict(error=str(exc)[:300])
            torch.cuda.synchronize()

    # secondary number at N > 1: the same step with the per-GPU batch held at the full batch (weak scaling)
    weak = None
    if world > 1 and strong and not args.no_weak:
        del optim, st
        torch.cuda.empty_cache()
        optim, pix_pin, lab_pin = build(B_total)
        st = optim._state
        wms, _, _, _ = timed_steps(optim, 2, 1, False)
        weak = dict(value=world * B_total * bonds * 2 / (wms * 1e-3), unit=UNIT, images_per_gpu=B_total, global_batch=B_total * world,
                    ms_per_step=wms / 2, steps=2, warmup=1,
                    note="same step with %d images on EVERY GPU (the replicated split costs the same, the contractions do not shrink)" % B_total)

    inference = inference_cfg5 = None
    extras = cfg_name == "cfg3" and not args.no_inference
    tf32_peak = measure_tf32_peak(torch, device)
    if extras:
        del optim, st
        torch.cuda.empty_cache()
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        inference = inference_rate(N, D, 60000, "cfg3 (60 000 images)")
        inference_waves = inference_rate(N, D, 3 * sms * 256, "cfg3 (three full waves of the chain kernel)")
        inference["three_full_waves"] = dict(value=inference_waves["value"], images_per_gpu=3 * sms * 256, ms_per_pass=inference_waves["ms_per_pass"])
        torch.cuda.empty_cache()
        inference_cfg5 = inference_rate(784, 32, args.inference_images, "cfg5")

    # datasource side of the path: one fused launch per batch fetch -- slice, 2x2 average pool and
    # (B, pixels) -> (N, B) transposition -- from raw float32 images resident in HBM.  HBM bound: bytes = images read + batch written.
    def fetch_rate(Bi, shrink, label):
        gen = torch.Generator(device=device).manual_seed(11 + rank)
        imgs = torch.rand((Bi, 784), device=device, generator=gen)
        n_sites = 196 if shrink else 784
        outp = torch.empty((n_sites, Bi), device=device)
        for _ in range(3):
            dev.preprocess_batch(imgs, 28, 28, shrink, 0, Bi, out=outp)
        barrier()
        reps = 20
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for r_ in range(reps):
            dev.preprocess_batch(imgs, 28, 28, shrink, (r_ * 977) % Bi, Bi, out=outp)
        p1.record()
        barrier()
        pms = max_over_ranks(p0.elapsed_time(p1) / reps)
        nbytes = Bi * 4.0 * (784 + n_sites)
        return dict(kernel="preprocess_%s_kernel<float>" % ("pool" if shrink else "flat"), bound="hbm",
                    workload="%s: %d float32 28x28 images per GPU -> (%d, B) site-major pixels%s" %
                             (label, Bi, n_sites, ", 2x2 average pool" if shrink else ""),
                    images_per_sec=world * Bi / (pms * 1e-3), ms_per_launch=pms, bytes_per_launch=nbytes,
                    achieved=nbytes / (pms * 1e-3) / 1e9, unit="GB/s",
                    l2="%.0f MB per launch, larger than L2" % (nbytes / 1e6),
                    # ncu --set full of the same launches (profiles/r1m_datasource_kernels_ncu.txt): dram read + write
                    traffic={(60000, True): 0.188178e9 + 0.022715e9, (1000000, False): 3.136020e9 + 3.086650e9}.get((Bi, shrink)),
                    traffic_source="ncu dram__bytes_read+write per launch (profiles/r1m_datasource_kernels_ncu.txt): the algorithmic "
                                   "bytes, no re-reads; part of a 60k-image batch's output is still in L2 when the kernel ends")

    datasource = None
    if extras:
        torch.cuda.empty_cache()
        datasource = [fetch_rate(60000, True, "cfg3 batch fetch"), fetch_rate(args.inference_images, False, "cfg5 batch fetch")]
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if datasource:
        for d_ in datasource:
            d_.update(peak=peaks["hbm_gbs"], frac=d_["achieved"] / peaks["hbm_gbs"], peak_source="%s HBM copy bandwidth" % peaks["source"])

    # ---- rooflines.  Algorithmic work per image*bond: 6 L d^2 Dl Dr + 2 d Dl m flops (three 2 L K contractions
    #      f(bond), G, f(delta) + one environment advance), 4 (Dl + Dr + m + 3) bytes.  Bond dimensions: the realised ones for
    #      the full-bond MPS (outer bonds 2L), nominal D for the reference-initialised runs whose bonds grow during the sweep.
    if cfg["init"] == "full_bond":
        dl = np.array([2 * L] + [D] * (N - 2), dtype=np.float64)[:N - 1]          # left bond of bond i
        dr = np.array([D] * (N - 2) + [2 * L], dtype=np.float64)[:N - 1]          # right bond of bond i
        flops_ib = 6.0 * L * 4 * dl * dr + 2.0 * 2 * dl * D
        bytes_ib = 4.0 * (dl + dr + D + 3)
        flops_step = float(B * world * flops_ib.sum()) * (1 if cfg["kind"] == "l2r" else 2)
        bytes_step = float(B * world * bytes_ib.sum()) * (1 if cfg["kind"] == "l2r" else 2)
        dims_note = "realised bond dimensions (all %d, outer bonds %d)" % (D, 2 * L)
    else:
        flops_step = float(B * world * bonds) * (6.0 * L * 4 * D * D + 2.0 * 2 * D * D)
        bytes_step = float(B * world * bonds) * 4.0 * (3 * D + 3)
        dims_note = "nominal D=%d at every bond (the realised bonds grow from 2L=%d to max_size during the first sweep: an upper bound on the work)" % (D, 2 * L)
    step_s = ms_total * 1e-3 / args.steps
    step_tflops = flops_step / step_s / 1e12 / world              # per GPU
    breakdown = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
    counts = {k: v[1] for k, v in prof.items()}
    fwd_ms, fwd_n = prof["forward"]
    grad_ms, grad_n = prof["gradient"]
    env_ms, env_n = prof["env"]
    env_launches = args.steps * ((N - 1) + bonds)       # one timed region holds the whole initial build
    svd_ms, svd_n = prof["svd"]
    flops_launch = 2.0 * B * L * 4 * D * D                        # one 2 L d^2 Dl Dr contraction over this rank's images
    kernels = dict(
        forward=dict(kernel="forward_h_kernel (+ absmax, image prep): f(bond) and f(delta), tcgen05 kind::f16 with fp16 hi/lo operands "
                            "(3 products per FP32 product); `achieved` counts the ALGORITHMIC 2LK flops of both launches -- f(bond) is "
                            "evaluated in its single-site form (special node x cached environment of the same site), which executes half "
                            "of them (`executed_fraction`)", bound="tensor", avg_launch_ms=fwd_ms / max(fwd_n, 1), launches=fwd_n,
                     achieved=flops_launch * fwd_n / max(fwd_ms * 1e-3, 1e-12) / 1e12, peak=tf32_peak, unit="TFLOP/s", executed_fraction=0.75,
                     traffic=61.7e6 if (B == 60000 and D == 120) else None, share_of_step=fwd_ms / ms_total),
        gradient=dict(kernel="grad_scale_kernel + grad_h_kernel + sum_parts_scaled_kernel: G = P^T W over the batch, tcgen05 kind::f16 with fp16 hi/lo operands "
                             "(3 products per FP32 product; 3xTF32 grad_tc_kernel with trmps_set_option(0, 0))", bound="tensor",
                      avg_launch_ms=grad_ms / max(grad_n, 1), launches=grad_n,
                      achieved=flops_launch * grad_n / max(grad_ms * 1e-3, 1e-12) / 1e12, peak=tf32_peak, unit="TFLOP/s",
                      traffic=62.9e6 if (B == 60000 and D == 120) else None, share_of_step=grad_ms / ms_total),
        env=dict(kernel="chain_tc_kernel from / to HBM (tcgen05; env_step_kernel when the chain is not supported): one environment step with the row written back (initial build + one advance per bond)",
                 bound="hbm", avg_launch_ms=env_ms / max(env_launches, 1), launches=env_launches,
                 achieved=4.0 * (2 * D + 1) * B * env_launches / max(env_ms * 1e-3, 1e-12) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                 traffic=None, share_of_step=env_ms / ms_total),
        split=dict(kernel="split_gram_kernel + split_chol_cluster_kernel + split_jacobi_kernel + split_rank_kernel + split_project_kernel "
                          "(FP64 Gram -> Cholesky spread over a thread-block cluster -> one-sided Jacobi on the factor in one cluster of up to 16 CTAs -> U^T M)",
                   bound="latency", avg_split_ms=svd_ms / max(svd_n, 1), splits=svd_n,
                   jacobi_sweeps_per_split=jacobi_sweeps / max(1.0, float(svd_n)), share_of_step=svd_ms / ms_total,
                   note="replicated on every rank and strictly sequential (bond i+1 needs its result): no bandwidth or flop roofline "
                        "applies; %d dependent rotation steps per Jacobi sweep on an n x n factor, n = 2 D = %d; per-kernel times "
                        "in profiles/" % (2 * D - 1, 2 * D)))
    for k_ in ("forward", "gradient", "env"):
        kernels[k_]["frac"] = kernels[k_]["achieved"] / kernels[k_]["peak"]
    dom = max(("forward", "gradient", "env", "split"), key=lambda k_: kernels[k_]["share_of_step"])
    roofline = dict(kernel="whole step (every kernel of the sweep; SURVEY.md 8d counts three 2LK contractions + one environment "
                           "advance per image*bond)", bound="tensor", achieved=step_tflops, peak=tf32_peak, unit="TFLOP/s",
                    frac=step_tflops / tf32_peak, traffic=None,
                    algorithmic_flops_per_step=flops_step, algorithmic_bytes_per_step=bytes_step, dims=dims_note,
                    hbm_frac=bytes_step / world / step_s / 1e9 / peaks["hbm_gbs"],
                    peak_source="TF32 dense GEMM measured in this run (cuBLAS 8192^3, %.0f TFLOP/s); bf16 sustained from MEASURED_PEAKS.json "
                                "is %.0f" % (tf32_peak, peaks["bf16_sustained"]),
                    dominant_kernel=dom, dominant_share_of_step=kernels[dom]["share_of_step"])
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms_total / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f32",
               data="synthetic",
               config=dict(workload=workload_name(cfg_name, cfg, B * world if not strong else B_total, world, args.scaling),
                           images_per_gpu=B, global_batch=B * world, bonds_per_step=bonds, parallelism="dp%d" % world,
                           rate_of_change=lr, bond_stride=int(Dp),
                           l2="inputs larger than L2 (environment caches %.1f GB per GPU)" % (2 * N * B * Dp * 4 / 1e9),
                           realised_m=int(iw[nat.I_M]), largest_bond=int(dims_end.max()), updates=int(iw[nat.I_N_UPDATES]),
                           reverts=int(iw[nat.I_N_REVERT]), armijo_exhausted=int(iw[nat.I_N_EXHAUST]), initial_logit_spread=spread),
               clocks=clocks, gpu_launches=int(launches), ranks_bit_identical=ranks_bit_identical,
               e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                        steps=args.e2e_steps, last_cost=last_cost),
               roofline=roofline, roofline_kernels=kernels, breakdown_ms_per_step=breakdown, breakdown_regions=counts,
               peaks=dict(tf32_tflops_measured=tf32_peak, hbm_gbs=peaks["hbm_gbs"], bf16_tflops_sustained=peaks["bf16_sustained"],
                          source=peaks["source"]))
    if weak is not None:
        out["weak_scaling"] = weak
    if graph_line is not None:
        out["cuda_graph"] = graph_line
    if inference is not None:
        out.update(inference=inference, inference_cfg5=inference_cfg5, roofline_datasource=datasource)
    if world == 1 and not args.no_cpu_baseline:
        r, dt, detail = cpu_sweep_rate(cfg, args.cpu_sample_images, args.cpu_sample_bonds, lr)
        out["cpu_baseline"] = dict(value=r, unit=UNIT, cores=blas_threads(), host_cores=os.cpu_count(), kind="port",
                                   sample="%d images x the first %d bond updates of the same sweep (%.1f s charged; same chain, initial "
                                          "MPS, feature map and rate_of_change as the GPU arm; environment build pro rata), literal "
                                          "reference graph in NumPy; %.1f Armijo trials per update"
                                          % (args.cpu_sample_images, min(args.cpu_sample_bonds, N - 1), dt, detail["armijo_trials_per_update"]))
        if extras:
            ri, dti = cpu_inference_rate(4000, N, D, L)
            out["cpu_baseline"].update(inference_value=ri, inference_unit="images/s",
                                       inference_sample="cfg3 predict on 4000 images (%.1f s), NumPy chain" % dti)
            rf, dtf = cpu_fetch_rate(60000)
            out["cpu_baseline"].update(fetch_value=rf, fetch_unit="images/s",
                                       fetch_sample="cfg3 batch fetch of 60000 images (%.2f s): NumPy average pool + slice + swapaxes copy" % dtf)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: the contract is ONE JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration; cfg3 (N=196, D=120, 60k images) is the one the metric is quoted on")
    ap.add_argument("--images", type=int, default=0, help="images in total (default: the configuration's)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-images", type=int, default=6000)
    ap.add_argument("--cpu-sample-bonds", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling measurement at N > 1")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="also time the steps replayed from a captured CUDA graph (secondary key cuda_graph); auto = on one GPU only")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default, BASELINE.json: '60k images, batch-sharded at 1/2/4/8'): the configuration's images in total, "
                         "sharded over the GPUs; weak: that many images on every GPU")
    ap.add_argument("--lr", type=float, default=0.0, help="rate_of_change of the sweep (MPSTrainingParameters); 0 = the configuration's")
    ap.add_argument("--inference-images", type=int, default=1000000, help="cfg5 images per GPU (extra line of the cfg3 run)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Everything libraries print (NCCL banners, reference-style progress prints) goes to stderr; stdout carries
    # exactly one JSON line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    captured = io.StringIO()
    with contextlib.redirect_stdout(captured):
        if args.impl == "reference":
            run_reference(args, rank)
        else:
            run_ours(args, rank, local_rank, world)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    lines = [l for l in captured.getvalue().splitlines() if l.startswith("{")]
    for l in captured.getvalue().splitlines():
        if not l.startswith("{"):
            print(l, file=sys.stderr)
    if lines:
        os.write(1, (lines[-1] + "\n").encode())


if __name__ == "__main__":
    main()
