import sys, time
sys.path.insert(0, "mps-mnist_b200/trmps")
import torch
import trmps_native as nat, trmps_device as dev
for N in (196, 784):
    total, L = 60000, 10
    data = torch.rand((total, N, 2), device="cuda"); lab = torch.eye(L, device="cuda")[torch.randint(0, L, (total,), device="cuda")].contiguous()
    dev.linreg_pretrain(data, lab, 0, 10, 0.05); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.linreg_pretrain(data, lab, 0, 1000, 0.05); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    # the same loop as separate torch ops (what the mirror did before)
    W = torch.zeros((N, L), device="cuda"); bb = torch.zeros((L,), device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for it in range(1000):
        rows = (it * 100 + torch.arange(100, device="cuda")) % total
        x = data[rows, :, 1]; y = lab[rows]
        g = (torch.softmax(x @ W + bb, 1) - y) / 100
        W -= 0.05 * (x.t() @ g); bb -= 0.05 * g.sum(0)
    torch.cuda.synchronize()
    print("N=%d: 1000 iterations in one launch %.2f ms (%.2f us per iteration); torch op loop %.1f ms" % (N, ms, ms, (time.perf_counter() - t0) * 1e3))
