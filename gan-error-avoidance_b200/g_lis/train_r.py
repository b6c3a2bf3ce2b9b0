"""Train a reverser R alone against a trained (frozen) G-LIS and D on the B200 kernels — the command line of
the reference's g_lis/train_r.py.

    python g_lis/train_r.py --synthetic --image_size 80 --code_size 256 --norm weight --r_iterations 1 \\
        --load_path /ckpt/exp01 --save_path_r /ckpt/exp01_r --niter 10000

Flags of the reference script (g_lis/train_r.py:40-144) with its defaults: ``--load_path`` (the G-LIS experiment
whose ``net_archive/{net}_gen.pt`` / ``_dis.pt`` are loaded, required), ``--save_path_r``, ``--load_path_r``,
``--net last``, ``--save_interval 500``, ``--vis_interval 100``, ``--test_interval 1000``.  One iteration =
``RSeparateTrainer.step`` (:406-436).  Checkpoints: ``net_archive/{iter}_{r,r_opt,state}.pt`` (:254-267).
"""
from __future__ import print_function

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))

import torch  # noqa: E402

_spec = importlib.util.spec_from_file_location("glis_b200_g_lis_main_for_r", os.path.join(HERE, "main.py"))
glis_main = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(glis_main)


def build_parser():
    p = glis_main.build_parser()
    p.description = __doc__
    p.set_defaults(test_interval=1000, vis_interval=100, save_interval=500, net="last")
    absent = ("--save_path", "--lambda_r", "--always_train_all", "--load_tolerant", "--nb_cache_total",
              "--nb_cache_lists", "--cache_p_drop", "--augment", "--d_dropout")
    for action in list(p._actions):
        if any(o in absent for o in action.option_strings):
            p._remove_action(action)
            for o in action.option_strings:
                p._option_string_actions.pop(o, None)
    p.add_argument("--save_path_r", default=None, help="path to save R's files")
    p.add_argument("--load_path_r", default=None, help="continue an existing R experiment")
    return p


def new_history():
    """g_lis/train_r.py:226-233."""
    from common.plotting import History
    h = History()
    h.add_group("loss-r", ["train"], increasing=False)
    h.add_group("loss-stage1", ["train"], increasing=False)
    h.add_group("loss-stage2", ["train"], increasing=True)
    h.add_group("loss-stage-mix", ["train-stage1", "train-stage2"], increasing=True)
    return h


def main(argv=None):
    opt = build_parser().parse_args(argv)
    opt.augment, opt.d_dropout, opt.save_path = "none", 0, None
    opt = glis_main.resolve_geometry(opt)
    if opt.norm not in ("weight", "weight-affine"):
        raise SystemExit("g_lis/train_r.py (B200 path): --norm must be weight or weight-affine")
    if opt.load_path is None:
        raise SystemExit("--load_path (the trained G-LIS experiment) is required")
    if opt.save_path_r is None and opt.load_path_r is None:
        raise ValueError("must specify --save_path_r if not continuing an R experiment")
    if opt.save_path_r is None:
        opt.save_path_r = opt.load_path_r
    if not torch.cuda.is_available():
        raise SystemExit("this path needs a CUDA device; there is no CPU fallback")

    from common.model import GeneratorLearnedInputSpace, build_discriminator, build_reverser
    from common.plotting import History
    from glis_b200 import _lib, dp, ops
    from glis_b200.trainer import RSeparateTrainer

    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if opt.precision:
        _lib.set_precision(opt.precision)
    data_seed = dp.seed_everything(opt.seed, rank)

    gen = GeneratorLearnedInputSpace(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.code_size, opt.norm,
                                     n_lis_layers=opt.r_iterations, upscaling=opt.g_upscaling)
    rev = build_reverser(opt.width, opt.height, opt.nfeature // 2, opt.nlayer, opt.code_size, opt.norm,
                         opt.spatial_dropout_r)
    dis = build_discriminator(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.norm)
    arch_in = os.path.join(opt.load_path, "net_archive")
    gen.load_state_dict(torch.load(os.path.join(arch_in, "{0}_gen.pt".format(opt.net)), map_location="cpu"))
    dis.load_state_dict(torch.load(os.path.join(arch_in, "{0}_dis.pt".format(opt.net)), map_location="cpu"))
    gen, rev, dis = gen.to(device), rev.to(device), dis.to(device)
    sync = dp.OverlappedGradSync(world) if world > 1 else None
    tr = RSeparateTrainer(gen, rev, dis, lr=opt.lr, r_iterations=opt.r_iterations, grad_sync=sync, ls=opt.ls)

    if rank == 0:
        for sub in ("", "samples", "samples_r", "samples_both", "net_archive"):
            os.makedirs(os.path.join(opt.save_path_r, sub), exist_ok=True)
    state = {"current_iter": 0, "best_iter": 0, "min_loss": 1e100}
    history = new_history()
    if opt.load_path_r is not None:
        arch_r = os.path.join(opt.load_path_r, "net_archive")
        loaded = torch.load(os.path.join(arch_r, "{0}_state.pt".format(opt.net)), map_location="cpu", weights_only=False)
        state.update({k: loaded[k] for k in ("current_iter", "best_iter", "min_loss") if k in loaded})
        if isinstance(loaded.get("history"), (bytes, str)):
            history = History.from_string(loaded["history"])
        r_fp = os.path.join(arch_r, "{0}_r.pt".format(opt.net))
        if os.path.isfile(r_fp):
            rev.load_state_dict(torch.load(r_fp, map_location="cpu"))
            tr.rev_flat.load_optimizer_state_dict(torch.load(os.path.join(arch_r, "{0}_r_opt.pt".format(opt.net)),
                                                             map_location="cpu"))
    vis_path = os.path.join(opt.load_path, "samples", "vis_code.pt")
    vis_code = (torch.load(vis_path) if os.path.exists(vis_path)
                else torch.randn(opt.vis_row * opt.vis_col, opt.code_size)).to(device)[:opt.vis_row * opt.vis_col]

    def visualize(it):
        """G(z), G(R(G(z))) and both side by side (g_lis/train_r.py:318-355)."""
        import torchvision
        modes = [(m, m.training) for m in (gen, rev)]
        for m, _ in modes:
            m.eval()
        with torch.no_grad(), glis_main.private_depth_rng(gen):
            a, _ = gen(vis_code, n_execute_lis_layers=opt.r_iterations)
            b, _ = gen(rev(a), n_execute_lis_layers=opt.r_iterations)
        for m, was in modes:
            m.train(was)
        scale = (lambda t: t * 2 - 1) if opt.output_scale else (lambda t: t)
        d = opt.save_path_r
        torchvision.utils.save_image(scale(a), os.path.join(d, "samples", "sample_{0}.jpg".format(it)), nrow=opt.vis_row)
        torchvision.utils.save_image(scale(b), os.path.join(d, "samples_r", "sample_{0}_r.jpg".format(it)), nrow=opt.vis_row)
        both = torch.stack([a, b], dim=1).reshape(-1, *a.shape[1:])
        torchvision.utils.save_image(scale(both), os.path.join(d, "samples_both", "sample_{0}_both.jpg".format(it)),
                                     nrow=2 * opt.vis_row)

    def save(prefix, it):
        d = os.path.join(opt.save_path_r, "net_archive")
        torch.save(rev.state_dict(), os.path.join(d, "{0}_r.pt".format(prefix)))
        torch.save(tr.rev_flat.optimizer_state_dict(opt.lr), os.path.join(d, "{0}_r_opt.pt".format(prefix)))
        out = dict(state, current_iter=it, index_shuffle=torch.zeros(0, dtype=torch.long), current_sample=0,
                   history=history.to_string())
        torch.save(out, os.path.join(d, "{0}_state.pt".format(prefix)))

    z = torch.empty(opt.batch_size, opt.code_size, device=device)
    it = state["current_iter"]
    while it < opt.niter:
        it += 1
        ops.randn_(z, data_seed + 7, it * glis_main.NoiseSource.STRIDE)
        out = tr.step(z)
        if rank == 0 and it % opt.log_interval == 0:
            v = {k: t.item() for k, t in out.items()}
            history.add_value("loss-r", "train", it, v["r"])
            history.add_value("loss-stage1", "train", it, v["stage1"])
            history.add_value("loss-stage2", "train", it, v["stage2"])
            history.add_value("loss-stage-mix", "train-stage1", it, v["stage1"])
            history.add_value("loss-stage-mix", "train-stage2", it, v["stage2"])
            print("{0} | loss: r:{1} dis-stage1:{2} dis-stage2:{3}".format(it, v["r"], v["stage1"], v["stage2"]))
        if rank == 0 and it % opt.vis_interval == 0:
            visualize(it)
        if rank == 0 and it % opt.save_interval == 0:
            save(it, it)


if __name__ == "__main__":
    main()
