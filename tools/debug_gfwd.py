import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, oracle
import common.model as pm
from glis_b200 import _lib
from util import copy_params, rel_err

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
_lib.set_precision(mode)
W = H = 32; nf = 64; nl = 3; code = 256; B = 32
torch.manual_seed(11)
og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
copy_params(pg, og)
og = og.double(); pg = pg.cuda()
z = torch.randn(B, code, generator=torch.Generator().manual_seed(5))
with torch.no_grad():
    xo, xp = z.double(), z.cuda()
    xo = xo + og.lis_layers[0](xo); xp = xp + pg.lis_layers[0](xp)
    print("lis", rel_err(xp, xo))
    for lo_, lp_ in zip(og.initial_linear, pg.initial_linear):
        xo, xp = lo_(xo), lp_(xp)
        print(type(lp_).__name__, tuple(xp.shape), "%.3e" % rel_err(xp, xo), "amax %.3e" % xo.abs().max().item())
    for lo_, lp_ in zip(og.conv_layers, pg.conv_layers):
        xin_o, xin_p = xo, xp
        xo, xp = lo_(xo), lp_(xp)
        # error of this layer alone: feed the product layer the oracle's input
        alone = lp_(xin_o.float().cuda())
        print(type(lp_).__name__, tuple(xp.shape), "chain %.3e" % rel_err(xp, xo), "alone %.3e" % rel_err(alone, xo), "amax %.3e" % xo.abs().max().item())
