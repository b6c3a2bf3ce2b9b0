import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, oracle
import common.model as pm
from glis_b200 import _lib
from util import copy_params, rel_err
import torch.nn.functional as F

W = H = 32; nf = 64; nl = 3; code = 256; B = 32
torch.manual_seed(11)
og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
copy_params(pd, od)
og, od = og.double(), od.double(); pd = pd.cuda()
gen = torch.Generator().manual_seed(5)
real = torch.rand(B, 3, H, W, generator=gen); zd = torch.randn(B, code, generator=gen)
with torch.no_grad():
    fake, _ = og(zd.double(), n_execute_lis_layers=1)
for name, x, t in (("real", real.double(), 1.0), ("fake", fake, 0.0)):
    for p in od.parameters(): p.grad = None
    lo = F.binary_cross_entropy(od(x), torch.full((B, 1), t, dtype=torch.float64)); lo.backward()
    for mode in ("bf16x3",):
        _lib.set_precision(mode)
        for p in pd.parameters(): p.grad = None
        lp = F.binary_cross_entropy(pd(x.float().cuda()), torch.full((B, 1), t, device="cuda")); lp.backward()
        print(name, mode, "loss", lp.item(), lo.item())
        for (n, p), (_, q) in zip(pd.named_parameters(), od.named_parameters()):
            print("   %-28s rel_err %.3e  |g|max %.3e" % (n, rel_err(p.grad, q.grad), q.grad.abs().max().item()))
