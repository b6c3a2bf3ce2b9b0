import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gan-error-avoidance_b200")): sys.path.insert(0, p)
import torch, torch.distributed as dist
import common.model as pm
from glis_b200 import dp
from glis_b200.trainer import GLISTrainer, GraphedStep
rank, world, local = dp.init_from_env("nccl"); torch.cuda.set_device(local); dev = torch.device("cuda", local)
torch.manual_seed(77)
g = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 2, "fractional").to(dev)
d = pm.build_discriminator(32, 32, 16, 3, "weight", 0).to(dev)
sync = dp.OverlappedGradSync(world, bucket_mb=0.05)
tr = GLISTrainer(g, d, lr=2e-5, grad_sync=sync)
gs = GraphedStep(tr, 4, 32, 32, 32, dev, warmup=1)
gen = torch.Generator().manual_seed(5 + rank)
def cmp(name, t):
    ref = t.clone(); dist.broadcast(ref, 0)
    diff = (t - ref).abs()
    if rank == 1: print(name, "max diff vs rank0 %.3e" % diff.max().item(), "first bad idx", int(diff.gt(0).float().argmax()) if diff.max() > 0 else -1, "numel", t.numel(), flush=True)
for it, dep in enumerate([(2, 1), (2, 1), (0, 2), (2, 1)]):
    real = torch.rand(4, 3, 32, 32, generator=gen).to(dev); zd = torch.randn(4, 32, generator=gen).to(dev); zg = torch.randn(4, 32, generator=gen).to(dev)
    gs.step(real, zd, zg, *dep)
    torch.cuda.synchronize()
    if rank == 1: print("iter", it, dep, "buckets gen", sync.sets["gen"]["buckets"], "dis", sync.sets["dis"]["buckets"][:3], flush=True)
    cmp("  gen.g", tr.gen_flat.g); cmp("  dis.g", tr.dis_flat.g); cmp("  gen.p", tr.gen_flat.p); cmp("  dis.p", tr.dis_flat.p)
dist.destroy_process_group()
