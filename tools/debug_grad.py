import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, oracle
import common.model as pm
from glis_b200 import _lib
from util import copy_params, rel_err
import torch.nn.functional as F

torch.manual_seed(11)
od = oracle.build_discriminator(32, 32, 64, 3, "weight", 0)
pd = pm.build_discriminator(32, 32, 64, 3, "weight", 0)
copy_params(pd, od)
od = od.double(); pd = pd.cuda()
x = torch.rand(32, 3, 32, 32)
lo = F.binary_cross_entropy(od(x.double()), torch.ones(32, 1, dtype=torch.float64)); lo.backward()
ref = {n: p.grad.clone() for n, p in od.named_parameters()}
for mode in ("fp32", "bf16x3", "bf16"):
    _lib.set_precision(mode)
    for p in pd.parameters(): p.grad = None
    lp = F.binary_cross_entropy(pd(x.cuda()), torch.ones(32, 1, device="cuda")); lp.backward()
    print(mode, "loss", lp.item(), lo.item())
    for (n, p), (_, q) in zip(pd.named_parameters(), od.named_parameters()):
        print("   %-28s rel_err %.3e  |g|max %.3e" % (n, rel_err(p.grad, q.grad), q.grad.abs().max().item()))
