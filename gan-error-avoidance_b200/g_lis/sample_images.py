"""Sample images from a trained G-LIS generator on the B200 kernels — the forward-only consumers of the reference's
g_lis/sample_images.py: grids per LIS depth, R-separate repairs, chains across depths, interpolations,
perturbations and the embedding of given images into the latent space.

    python g_lis/sample_images.py --image_size 80 --code_size 256 --norm weight --r_iterations 1 \\
        --load_path_g /ckpt/exp01/net_archive/last_gen.pt --save_path /tmp/samples [--load_path_r .../last_r.pt]

Flags as in the reference (g_lis/sample_images.py:49-97).  Added: ``--rounds`` (the reference hard-codes 20 of each
kind), ``--seed``, ``--precision``, ``--embed_steps`` / ``--embed_lr`` (:299: 100000 Adam steps at 1e-4) and
``--real_images`` (a ``.pt`` tensor (N, 3, H, W) in [0, 1] instead of the reference's hard-coded jpg folder).
Everything here is inference through the product's modules: G forward (every LIS depth), R forward, and for the
embedding G's data-gradient-only backward (parameters frozen).
"""
from __future__ import print_function

import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402


def _to_uint8(images, output_scale):
    """(N, 3, H, W) in [0, 1] -> uint8 (N, H, W, 3) exactly as :262-263 (optionally x*2-1 first, :260-261)."""
    if output_scale:
        images = images * 2 - 1
    return (images.detach().float().cpu() * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def generate_images(gen, r, code, n_execute_lis_layers, batch_size, output_scale=False):
    """Images of ``code`` at the given LIS depth, in batches, and (with a reverser) the same images after one
    round trip G(R(G(z))) (:246-266).  Returns (uint8 (N,H,W,3), uint8 (N,H,W,3) or None)."""
    plain, fixed = [], []
    with torch.no_grad():
        for i in range(0, code.size(0), batch_size):
            images, _ = gen(code[i:i + batch_size], n_execute_lis_layers=n_execute_lis_layers)
            if r is not None:
                again, _ = gen(r(images), n_execute_lis_layers=n_execute_lis_layers)
                fixed.append(_to_uint8(again, output_scale))
            plain.append(_to_uint8(images, output_scale))
    return torch.cat(plain), (torch.cat(fixed) if fixed else None)


def generate_interpolations(gen, code_start, code_end, nb_steps, n_execute_lis_layers, output_scale=False):
    """start, start + i*(end-start)/nb_steps for i < nb_steps, end (:268-281: nb_steps + 2 images; the start
    appears twice, as in the reference)."""
    step = (code_end - code_start) / nb_steps
    vectors = [code_start] + [code_start + i * step for i in range(nb_steps)] + [code_end]
    with torch.no_grad():
        images, _ = gen(torch.stack(vectors), n_execute_lis_layers=n_execute_lis_layers)
    return _to_uint8(images, output_scale)


def generate_perturbations(gen, code, seed, stddev, nb_images, n_execute_lis_layers, output_scale=False):
    """``nb_images`` noisy copies of one code (:283-297).  As in the reference the first row of the noise is
    REPLACED by the code itself before it is added (``rands[0] = codes[0]``), so image 0 shows 2*code."""
    if code.dim() == 1:
        code = code.unsqueeze(0)
    assert code.size(0) == 1
    g = torch.Generator().manual_seed(int(seed))
    rands = (torch.randn(nb_images, code.size(1), generator=g) * stddev).to(code.device, code.dtype)
    codes = code.expand(nb_images, code.size(1))
    rands[0] = codes[0]
    with torch.no_grad():
        images, _ = gen(codes + rands, n_execute_lis_layers=n_execute_lis_layers)
    return _to_uint8(images, False)


def embed_real_images(gen, r, images, lr=1e-4, test_steps=1000, log=None):
    """Latent codes whose images approximate ``images`` (:299-335): start from R(images), then Adam on the codes
    with G frozen (forward at the default LIS depth of eval mode: all modules)."""
    flags = [p.requires_grad for p in gen.parameters()]
    for p in gen.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        code = r(images).detach().clone()
    code.requires_grad_(True)
    opt = torch.optim.Adam([code], lr=lr)
    for j in range(test_steps):
        generated, _ = gen(code)
        loss = torch.nn.functional.mse_loss(generated, images)
        loss.backward()
        opt.step()
        code.grad.zero_()
        if log is not None and j % 100 == 0:
            log("Embedding real images... iter %d with loss %.08f and lr %.08f" % (j, loss.item(), lr))
    for p, f in zip(gen.parameters(), flags):
        p.requires_grad_(f)
    return code.detach()


def _grid(images_u8, cols):
    import torchvision
    t = images_u8.permute(0, 3, 1, 2).float() / 255
    return torchvision.utils.make_grid(t, nrow=cols, padding=1)


def _save(t, path):
    import torchvision
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torchvision.utils.save_image(t, path)


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    a = p.add_argument
    a("--batch_size", type=int, default=32)
    a("--image_size", type=int, default=-1)
    a("--width", type=int, default=-1)
    a("--height", type=int, default=-1)
    a("--code_size", type=int, default=128)
    a("--nfeature", type=int, default=64)
    a("--nlayer", type=int, default=-1)
    a("--norm", default="none")
    a("--load_path_g", required=True, help="state dict of the generator (…_gen.pt)")
    a("--load_path_r", default=None, help="state dict of an R-separate reverser (…_r.pt)")
    a("--save_path", required=True)
    a("--output_scale", action="store_true", default=False)
    a("--r_iterations", type=int, default=3)
    a("--spatial_dropout_r", type=float, default=0)
    a("--with_real_images", action="store_true", default=False)
    a("--g_upscaling", default="fractional")
    a("--rounds", type=int, default=20, help="how many files of each kind (the reference writes 20)")
    a("--seed", type=int, default=1234)
    a("--precision", default=None, choices=["fp32", "bf16x3", "bf16"])
    a("--real_images", default=None, help=".pt tensor (N,3,H,W) in [0,1] to embed (--with_real_images)")
    a("--embed_steps", type=int, default=100000)
    a("--embed_lr", type=float, default=1e-4)
    return p


def main(argv=None):
    opt = build_parser().parse_args(argv)
    if opt.width <= 0 or opt.height <= 0:
        if opt.image_size <= 0:
            raise ValueError("must specify valid image size")
        opt.width = opt.height = opt.image_size
    if opt.nlayer < 0:
        opt.nlayer, s = 0, max(opt.width, opt.height)
        while s >= 8:
            s = (s + 1) // 2
            opt.nlayer += 1
    if not torch.cuda.is_available():
        raise SystemExit("this path needs a CUDA device; there is no CPU fallback")
    from common.model import GeneratorLearnedInputSpace, build_reverser
    from glis_b200 import _lib
    if opt.precision:
        _lib.set_precision(opt.precision)
    torch.manual_seed(opt.seed)
    dev = torch.device("cuda")
    gen = GeneratorLearnedInputSpace(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.code_size, opt.norm,
                                     n_lis_layers=opt.r_iterations, upscaling=opt.g_upscaling)
    gen.load_state_dict(torch.load(opt.load_path_g, map_location="cpu"))
    gen = gen.to(dev).eval()
    r = None
    if opt.load_path_r is not None:
        r = build_reverser(opt.width, opt.height, opt.nfeature // 2, opt.nlayer, opt.code_size, opt.norm,
                           opt.spatial_dropout_r)
        r.load_state_dict(torch.load(opt.load_path_r, map_location="cpu"))
        r = r.to(dev).eval()
    depths = range(1 + opt.r_iterations)
    out = opt.save_path
    for i in range(opt.rounds):
        rows = cols = 16
        code = torch.randn(rows * cols, opt.code_size, device=dev)
        by_r = []
        for k in depths:
            images, fixed = generate_images(gen, r, code, k, opt.batch_size, opt.output_scale)
            by_r.append(images)
            _save(_grid(images, cols), os.path.join(out, "sampled_images_r%d" % k, "r%d_full_%04d.jpg" % (k, i)))
            _save(_grid(images[:64], 8), os.path.join(out, "sampled_images_r%d" % k, "r%d_small_%04d.jpg" % (k, i)))
            if fixed is not None:
                _save(_grid(fixed, cols), os.path.join(out, "sampled_images_rsep_r%d_after" % k, "rsep_r%d_full_%04d.jpg" % (k, i)))
                both = torch.cat([images, fixed], dim=2)        # before | after, side by side
                _save(_grid(both, cols), os.path.join(out, "sampled_images_rsep_r%d_both" % k, "rsep_r%d_chain_full_%04d.jpg" % (k, i)))
        chains = torch.cat(by_r, dim=2)                          # one image per depth, left to right
        _save(_grid(chains[:64], 4), os.path.join(out, "sampled_images_chains", "chain_full_%04d.jpg" % i))
        _save(_grid(chains[:16], 4), os.path.join(out, "sampled_images_chains", "chain_small_%04d.jpg" % i))
    for i in range(opt.rounds):
        codes = torch.randn(2, opt.code_size, device=dev)
        rows_ = []
        for k in depths:
            strip = torch.cat(list(generate_interpolations(gen, codes[0], codes[1], 8, k, opt.output_scale)), dim=1)
            _save(strip.permute(2, 0, 1).float() / 255, os.path.join(out, "sampled_images_interpolations_r%d" % k, "interp_r%d_%04d.jpg" % (k, i)))
            rows_.append(strip)
        _save(torch.cat(rows_, dim=0).permute(2, 0, 1).float() / 255, os.path.join(out, "sampled_images_interpolations_all", "interp_all_%04d.jpg" % i))
    for i in range(opt.rounds):
        code = torch.randn(1, opt.code_size, device=dev)
        seed = int(torch.randint(0, 10 ** 6, (1,)).item())
        for k in depths:
            images = generate_perturbations(gen, code, seed, 1.0, 64, k, opt.output_scale)
            _save(_grid(images, 8), os.path.join(out, "sampled_images_perturbations_r%d" % k, "pert_r%d_%04d.jpg" % (k, i)))
    if opt.with_real_images and r is not None and opt.real_images:
        real = torch.load(opt.real_images, map_location="cpu").float().to(dev)
        codes = embed_real_images(gen, r, real, lr=opt.embed_lr, test_steps=opt.embed_steps, log=print)
        for i in range(codes.size(0)):
            images = generate_perturbations(gen, codes[i], 42, 1.0, 49, opt.r_iterations, opt.output_scale)
            _save(_grid(images, 7), os.path.join(out, "sampled_images_real_images_perturbations", "pert_real_%04d.jpg" % i))


if __name__ == "__main__":
    main()
