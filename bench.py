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

import numpy as np

CFG = dict(N=196, D=120, L=10, d=2, B=60000)          # configs[2] of BASELINE.json ("cfg3")
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
def cpu_sweep_rate(B_s, bonds, D, L, seed=0, reps=1):
    """Times the oracle (literal restatement of the reference graph: materialised C, four contractions per
    update, LAPACK SVD; oracle/trmps_oracle.py) on the host cores for `bonds` full-size bond updates on a
    B_s-image sample.  Returns (images*bonds/s, seconds, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import trmps_oracle as o
    Nc = bonds + 3
    loc = 1
    pix, lab = synthetic_pixels(Nc, B_s, L, seed)
    phi = o.feature_map(pix, o.FM_COS_SIN)
    nodes = full_bond_mps(Nc, 2, L, D, loc)
    p = o.SweepParams(max_size=D, min_singular_value=0.0)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        C1s, C2s = o.setup_environments(nodes, loc, phi, L)
        o.sweep_right(nodes, loc, loc + bonds, C1s, C2s, phi, lab, 1e-5, p)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return B_s * bonds / best, best, os.cpu_count()


def cpu_inference_rate(B_s, N, D, L, seed=0):
    """Times the oracle's MPS.predict (mps.py:521-560 restated: one batched chain step per site) on the host cores
    for a B_s-image sample at the cfg3 shape.  Returns (images/s, seconds)."""
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


def run_reference(args, rank):
    if rank != 0:
        return
    B_s, bonds = args.cpu_sample_images, args.cpu_sample_bonds
    for _ in range(args.warmup):
        cpu_sweep_rate(min(B_s, 256), 1, CFG["D"], CFG["L"])
    times, rates = [], []
    for _ in range(args.steps):
        r, dt, cores = cpu_sweep_rate(B_s, bonds, CFG["D"], CFG["L"])
        times.append(dt); rates.append(r)
    value = B_s * bonds * len(times) / sum(times)
    sample = "%d images x %d full-size (D=%d) bond updates per step, literal reference graph in NumPy" % (B_s, bonds, CFG["D"])
    out = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload="cfg3 N=196 D=120 B=60000/GPU L->R half-sweep (bounded CPU sample)", sample=sample),
               cpu_baseline=dict(value=value, unit=UNIT, cores=os.cpu_count(), kind="port", sample=sample),
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
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
    N, D, L, B = CFG["N"], CFG["D"], CFG["L"], args.images_per_gpu
    if args.scaling == "strong":
        B = B // world
    bonds = N - 1
    loc = 0
    lr = args.lr       # the gradient is a SUM over the batch (optimizer.py:249); the Armijo search halves from here
    pix, lab = synthetic_pixels(N, B, L, seed=100 + rank)
    nodes = full_bond_mps(N, 2, L, D, loc)
    # rescale the special node so that the initial logits have unit spread (same fixed sample on every rank, so the
    # replicated MPS stays identical): keeps the softmax un-saturated and the Armijo search in its working range
    probe_pix, _ = synthetic_pixels(N, 2048, L, seed=4242)
    probe = MPS(2, L, N, special_node_loc=loc)
    probe.prepare(_weights=nodes)
    spread = float(np.std(probe.predict(probe_pix, feature_map=nat.FM_COS_SIN)))
    nodes[loc] = (nodes[loc] / max(spread, 1e-30)).astype(np.float32)
    del probe

    net = MPS(2, L, N, special_node_loc=loc)
    net.prepare(_weights=nodes)
    params = MPSOptimizerParameters(min_singular_value=0.0, reg=1e-3, updates_per_step=1)
    optim = MPSOptimizer(net, D, params)
    optim.feature_map = nat.FM_COS_SIN
    optim.rate_of_change = lr
    if world > 1:
        optim.attach_process_group()
    pix_pin = torch.from_numpy(pix).pin_memory()
    lab_pin = torch.from_numpy(lab).pin_memory()
    optim.bind_batch(pix_pin, lab_pin, weights=nodes)
    st = optim._state
    opt = optim._opt(lr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def device_step():
        # inputs already resident in HBM: reset weights from a device copy, build environments, sweep L->R
        st.slots.copy_(slots0); st.special.copy_(special0); st.dims.copy_(dims0)
        st.loc = loc
        st.init_env(loc)
        st.sweep(loc, N - 1, +1, opt)

    slots0, special0, dims0 = st.slots.clone(), st.special.clone(), st.dims.clone()
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sweeps0 = int(st.iwork[nat.I_N_SWEEPS].item())
    nat.profile_begin(200000)
    l0 = nat.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    launches = nat.launch_count() - l0
    prof = nat.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * bonds * args.steps / (ms_total * 1e-3)
    iw = st.iwork.cpu().numpy()

    # end-to-end through the public Python API with HOST buffers: H2D of the batch and the weights, sweep, D2H of the weights
    def e2e_step():
        optim.bind_batch(pix_pin, lab_pin, weights=nodes)
        optim._setup_optimization()
        optim._sweep_right(loc, N - 1)
        out_nodes = optim.updated_nodes()
        cost = float(st.scalars()[nat.S_COST1].item())
        return out_nodes, cost

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        out_nodes, last_cost = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * bonds * args.e2e_steps / float(e2e_s.item())
    h2d = pix.nbytes + lab.nbytes + sum(w.nbytes for w in nodes)
    d2h = sum(w.nbytes for w in out_nodes) + 4

    # second headline metric: inference images/s (MPS.predict on raw pixels resident in HBM, fused cos/sin map):
    # at the headline shape cfg3 (N=196, D=120) and at cfg5 (N=784, D=32, 1M images per GPU)
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
        ims = torch.tensor([i0.elapsed_time(i1) / reps], device=device)
        if world > 1:
            dist.all_reduce(ims, op=dist.ReduceOp.MAX)
        ims = float(ims.item())
        iflops = Bi * ((Ni - 1) * 2.0 * 2 * Di * Di + 2.0 * L * 2 * Di * Di)
        return dict(metric="inference_images_per_sec", value=world * Bi / (ims * 1e-3), unit="images/s", ms_per_pass=ims,
                    config="%s: N=%d D=%d L=10, %d images per GPU, raw pixels resident in HBM, fused cos/sin map, "
                           "fp16 hi/lo split chain on tcgen05 (FP32-equivalent accuracy)" % (label, Ni, Di, Bi),
                    algorithmic_tflops=world * iflops / (ims * 1e-3) / 1e12,
                    pixel_stream_gbs=world * 4.0 * Ni * Bi / (ims * 1e-3) / 1e9)

    inference = inference_cfg5 = None
    if not args.no_inference:
        del optim, st
        torch.cuda.empty_cache()
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        b3 = args.inference_images_cfg3 if args.inference_images_cfg3 > 0 else 3 * sms * 256
        inference = inference_rate(N, D, b3, "cfg3")
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
        pms = torch.tensor([p0.elapsed_time(p1) / reps], device=device)
        if world > 1:
            dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        pms = float(pms.item())
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
    if not args.no_inference:
        torch.cuda.empty_cache()
        datasource = [fetch_rate(B, True, "cfg3 batch fetch"), fetch_rate(args.inference_images, False, "cfg5 batch fetch")]
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    if datasource:
        for d_ in datasource:
            d_.update(peak=peaks["hbm_gbs"], frac=d_["achieved"] / peaks["hbm_gbs"], peak_source="%s HBM copy bandwidth" % peaks["source"])
    tf32_peak = peaks["bf16_sustained"] / 2.0       # TF32 dense = bf16 / 2, kernel timed inside a long step
    fwd_ms, fwd_n = prof["forward"]
    flops_per_launch = 2.0 * B * L * 4 * D * D     # one 2*L*d^2*Dl*Dr contraction per image
    # the first bond has Dl = 2L < D; count its launches at their true size
    fwd_flops_total = (fwd_n - 2 * args.steps) * flops_per_launch + 2 * args.steps * 2.0 * B * L * 4 * (2 * L) * D
    achieved = fwd_flops_total / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else 0.0
    breakdown = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
    fp16_peak = peaks["bf16_sustained"]
    roofline_fwd = dict(kernel="forward_h_kernel, tcgen05 kind::f16 with fp16 hi/lo operands (f(bond) and f(delta): 2 of the 3 dense "
                               "contractions per update; achieved counts algorithmic FP32 flops, the tensor pipe executes 3x as "
                               "many 16-bit flops; time includes the absmax and image-prep kernels)",
                        bound="tensor", achieved=achieved, peak=fp16_peak, unit="TFLOP/s", frac=achieved / fp16_peak,
                        traffic=61.7e6 if (B == 60000 and D == 120) else None,
                        traffic_source="ncu --set full, dram__bytes_read+write per launch at B=60000, D=120 (profiles/r1g_tensor_kernels_ncu.txt): "
                                       "the two environment reads, no re-reads",
                        pipe_frac=3.0 * achieved / fp16_peak,
                        peak_source="%s bf16 sustained (16-bit dense tensor peak)" % peaks["source"], avg_launch_ms=fwd_ms / max(fwd_n, 1),
                        launches=fwd_n, share_of_step=fwd_ms / ms_total)
    # dominant kernel of the step: the one-sided Jacobi split.  Its streaming requirement is the row exchange: every
    # round each of the R/16 row-block pairs is loaded and stored once (R x (Cc + R) floats each way per round,
    # R/8 - 1 rounds per sweep); the working set (2.9 MB) lives in L2, so this is L2 traffic, not DRAM.
    Dp = dev.bond_stride(nodes, loc, L, D)
    R_, Cc_ = 2 * Dp, 2 * L * Dp
    svd_ms, svd_n = prof["svd"]
    total_sweeps = int(iw[nat.I_N_SWEEPS]) - sweeps0
    rounds_per_sweep = (2 * D + 7) // 8 - 1
    svd_bytes = total_sweeps * rounds_per_sweep * 2.0 * (2 * D) * (Cc_ + R_) * 4
    svd_gbs = svd_bytes / (svd_ms * 1e-3) / 1e9 if svd_ms > 0 else 0.0
    roofline_svd = dict(kernel="svd_jacobi_cluster_kernel (+ norms/finalize/scatter): one-sided block Jacobi, %d splits, %.1f sweeps "
                               "per split; bound by the %d dependent rounds per sweep (flag hand-off, cluster Gram reduction, "
                               "rotation recurrences: ~21k cycles per round), not by bandwidth or flops" %
                               (svd_n, total_sweeps / max(1.0, float(svd_n)), rounds_per_sweep),
                        bound="hbm", achieved=svd_gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=svd_gbs / peaks["hbm_gbs"],
                        traffic=None, traffic_source="ncu cannot replay the cooperative cluster kernel (LaunchFailed); the working set (2.9 MB) is L2 resident",
                        peak_source="%s HBM copy bandwidth" % peaks["source"],
                        avg_launch_ms=svd_ms / max(1.0, float(svd_n)), share_of_step=svd_ms / ms_total,
                        note="algorithmic bytes = row-exchange traffic through L2; DRAM traffic is the 2.9 MB bond once")
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms_total / args.steps, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f32",
               data="synthetic",
               config=dict(workload="cfg3: N=196 D=120 L=10 d=2, %d images per GPU, one L->R half-sweep (195 bond updates, "
                                    "init env build included), feature map cos/sin, full-bond random MPS, min_sv=0" % B,
                           images_per_gpu=B, global_batch=B * world, bonds_per_step=bonds, parallelism="dp%d" % world,
                           l2="inputs larger than L2 (environment caches %.1f GB per GPU)" % (2 * N * B * D * 4 / 1e9),
                           realised_m=int(iw[nat.I_M]), updates=int(iw[nat.I_N_UPDATES]), reverts=int(iw[nat.I_N_REVERT]),
                           armijo_exhausted=int(iw[nat.I_N_EXHAUST]), initial_logit_spread=spread),
               clocks=clocks, gpu_launches=int(launches),
               e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                        steps=args.e2e_steps, last_cost=last_cost),
               roofline=roofline_svd, roofline_forward=roofline_fwd,
               breakdown_ms_per_step=breakdown, inference=inference, inference_cfg5=inference_cfg5,
               roofline_datasource=datasource)
    if world == 1 and not args.no_cpu_baseline:
        r, dt, cores = cpu_sweep_rate(args.cpu_sample_images, args.cpu_sample_bonds, D, L)
        out["cpu_baseline"] = dict(value=r, unit=UNIT, cores=cores, kind="port",
                                   sample="%d images x %d full-size bond updates (%.1f s), literal reference graph in NumPy"
                                          % (args.cpu_sample_images, args.cpu_sample_bonds, dt))
        if not args.no_inference:
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
    ap.add_argument("--images-per-gpu", type=int, default=CFG["B"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-images", type=int, default=6000)
    ap.add_argument("--cpu-sample-bonds", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --images-per-gpu images on every GPU (default); strong: that many images in total, sharded")
    ap.add_argument("--lr", type=float, default=1e-4, help="rate_of_change of the sweep (MPSTrainingParameters)")
    ap.add_argument("--inference-images", type=int, default=1000000, help="cfg5 images per GPU")
    ap.add_argument("--inference-images-cfg3", type=int, default=0,
                    help="cfg3 (N=196, D=120) inference images per GPU; 0 = three full waves of the chain kernel (SMs x 256 images x 3)")
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
