"""Data-parallel parity check (run under torchrun, one process per GPU):
every rank holds a replicated MPS and a shard of the images; after one full training step the weights must be
(a) bit-identical on all ranks and (b) equal (to fp32 round-off) to a single-rank run on the concatenated batch."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import trmps_native as nat, trmps_device as dev
from trmps import MPS, MPSOptimizer, MPSOptimizerParameters
from bench import synthetic_pixels, full_bond_mps

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, L, D, B = 12, 10, 40, 1024 * world
pix, lab = synthetic_pixels(N, B, L, seed=5)
loc, lr = 0, 1e-4
nodes = full_bond_mps(N, 2, L, D, loc)

def run(pix_s, lab_s, attach):
    net = MPS(2, L, N, special_node_loc=loc); net.prepare(_weights=nodes)
    opt = MPSOptimizer(net, D, MPSOptimizerParameters(min_singular_value=0.0))
    opt.feature_map = nat.FM_COS_SIN; opt.rate_of_change = lr
    if attach: opt.attach_process_group()
    opt.bind_batch(np.ascontiguousarray(pix_s), np.ascontiguousarray(lab_s), weights=nodes)
    opt.train_step()
    torch.cuda.synchronize()
    return opt.updated_nodes(), opt._state.iwork.cpu().numpy()

per = B // world
got, iw = run(pix[:, rank * per:(rank + 1) * per], lab[rank * per:(rank + 1) * per], True)
flat = torch.from_numpy(np.concatenate([w.ravel() for w in got])).cuda()
ref = flat.clone(); dist.broadcast(ref, src=0)
same = torch.tensor([int(torch.equal(flat, ref))], device="cuda"); dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    want, iw1 = run(pix, lab, False)
    net = MPS(2, L, N, special_node_loc=loc)
    net.prepare(_weights=got); fa = net.predict(pix, feature_map=nat.FM_COS_SIN)
    net.prepare(_weights=want); fb = net.predict(pix, feature_map=nat.FM_COS_SIN)
    err = np.linalg.norm(fa - fb) / np.linalg.norm(fb)
    print("world=%d  ranks bit-identical: %s  updates %d/%d reverts %d/%d  logits vs single-rank rel err %.3e  argmax agree %.4f"
          % (world, bool(same.item()), iw[nat.I_N_UPDATES], iw1[nat.I_N_UPDATES], iw[nat.I_N_REVERT], iw1[nat.I_N_REVERT], err,
             float(np.mean(fa.argmax(1) == fb.argmax(1)))))
    assert bool(same.item()) and err < 5e-3
dist.destroy_process_group()
