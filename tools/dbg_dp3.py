"""2-rank repro loop: the all-reduced gradients of the sharded step, step after step, against the 1-GPU full step."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench
from gpzoo_b200 import functional
from gpzoo_b200.distributed import FlatGradReducer, shard_range
functional.set_sync_checks(False)
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev, dt = torch.device('cuda', local), torch.float32
dist.init_process_group("nccl", device_id=dev)
c = dict(bench.CONFIGS[2])
prob = bench.make_problem(c, c["N"], c["seed"], dt)
names = ["Z", "sigma", "lengthscale", "mu", "Lu", "W"]
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
ref = None
if rank == 0:
    m1, sh1 = bench.build_model(c, prob, dt, dev)
    e1 = m1.elbo(prob["X"].to(dev), prob["y"].to(dev), E=1, eps=prob["eps"].to(dev))
    (-e1).backward()
    ref = [p.grad.clone() for p in sh1]
    del m1, sh1, e1
lo, hi = shard_range(c["N"], world, rank)
sl = slice(lo, hi)
p = dict(prob); p["V"] = prob["V"][sl].contiguous()
model, shared = bench.build_model(c, p, dt, dev)
X, y, eps = prob["X"][sl].to(dev), prob["y"][:, sl].contiguous().to(dev), prob["eps"][:, :, sl].contiguous().to(dev)
red = FlatGradReducer(shared, device=dev, dtype=dt)
bad = 0
from gpzoo_b200 import _cabi
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    if mode == "profile":
        _cabi.profile = {} if (it // 10) % 2 == 1 else None
    if mode == "barrier" and it % 10 == 0:
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([1.0], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if mode == "check" and it % 10 == 9:
        functional.check_cholesky_info()
    for q in model.parameters():
        q.grad = None
    e = model.elbo(X, y, E=1, eps=eps, kl_weight=1.0 / world)
    (-e).backward()
    loc = [q.grad.clone() for q in shared]
    red.all_reduce(e)
    if rank == 0:
        errs = {n: rel(q.grad, r) for n, q, r in zip(names, shared, ref)}
        if max(errs.values()) > 1e-4:
            bad += 1
            print(f"it {it} BAD", {k: "%.1e" % v for k, v in errs.items()}, "local norms", {n: "%.3e" % float(g.norm()) for n, g in zip(names, loc)}, flush=True)
if rank == 0:
    print("bad iterations:", bad, flush=True)
dist.barrier()
dist.destroy_process_group()
