"""Train a GAN with the R-iterative scheme on the B200 kernels — the command line of the reference's
r_iterative/main.py.

    python r_iterative/main.py --synthetic --image_size 80 --code_size 256 --norm weight --lr 0.00002 \\
        --r_iterations 3 --always_train_all --batch_size 64 --niter 100 --save_path /tmp/exp

Every flag of the reference script (r_iterative/main.py:27-134) is accepted with the same name, type and default
(``--test_interval 1000 --vis_interval 100 --save_interval 2000``; no cache / augment / upscaling / dropout-in-D /
tolerant-load flags).  Added, as in g_lis/main.py: ``--synthetic``, ``--seed``, ``--precision``, ``--no_graph``,
``--log_interval``; under ``torchrun`` the batch is sharded over the ranks.

One outer iteration = ``RIterTrainer.step`` (r_iterative/main.py:428-535): a chain of 1 + r_iterations hops
code -> G -> images -> R -> code ..., each trained hop updating G, R (hops > 0) and D.  Checkpoints carry the
reference's file names (``net_archive/{prefix}_{gen,gen_opt,r,r_opt,dis,dis_opt,state}.pt``, :236-254).
"""
from __future__ import print_function

import importlib.util
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))

import torch  # noqa: E402

_spec = importlib.util.spec_from_file_location("glis_b200_g_lis_main", os.path.join(HERE, "..", "g_lis", "main.py"))
glis_main = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(glis_main)


def build_parser():
    """g_lis/main.py's parser with this script's defaults and without the flags it does not have."""
    p = glis_main.build_parser()
    p.description = __doc__
    p.set_defaults(test_interval=1000, vis_interval=100, save_interval=2000)
    absent = ("--load_tolerant", "--nb_cache_total", "--nb_cache_lists", "--cache_p_drop", "--augment",
              "--g_upscaling", "--d_dropout")
    for action in list(p._actions):
        if any(o in absent for o in action.option_strings):
            p._remove_action(action)
            for o in action.option_strings:
                p._option_string_actions.pop(o, None)
    return p


def new_history(r_iterations):
    """The loss-history groups of r_iterative/main.py:216-220."""
    from common.plotting import History
    h = History()
    h.add_group("loss-r-mix", ["train-r%d" % i for i in range(r_iterations)], increasing=False)
    h.add_group("loss-g-mix", ["train-g%d" % i for i in range(1 + r_iterations)], increasing=False)
    h.add_group("loss-d-mix", ["train-d-real%d" % i for i in range(1 + r_iterations)]
                + ["train-d-fake%d" % i for i in range(1 + r_iterations)], increasing=False)
    return h


def save_state(path, prefix, tr, state):
    d = os.path.join(path, "net_archive")
    for tag, net, flat in (("gen", tr.gen, tr.gen_flat), ("r", tr.rev, tr.rev_flat), ("dis", tr.dis, tr.dis_flat)):
        torch.save(net.state_dict(), os.path.join(d, "{0}_{1}.pt".format(prefix, tag)))
        torch.save(flat.optimizer_state_dict(tr.lr), os.path.join(d, "{0}_{1}_opt.pt".format(prefix, tag)))
    torch.save(state, os.path.join(d, "{0}_state.pt".format(prefix)))


def load_state(path, prefix, tr):
    d = os.path.join(path, "net_archive")
    for tag, net, flat in (("gen", tr.gen, tr.gen_flat), ("r", tr.rev, tr.rev_flat), ("dis", tr.dis, tr.dis_flat)):
        net.load_state_dict(torch.load(os.path.join(d, "{0}_{1}.pt".format(prefix, tag)), map_location="cpu"))
        flat.load_optimizer_state_dict(torch.load(os.path.join(d, "{0}_{1}_opt.pt".format(prefix, tag)),
                                                  map_location="cpu"))
    return torch.load(os.path.join(d, "{0}_state.pt".format(prefix)), map_location="cpu")


def main(argv=None):
    opt = glis_main.resolve_geometry(build_parser().parse_args(argv))
    if opt.norm not in ("weight", "weight-affine"):
        raise SystemExit("r_iterative/main.py (B200 path): --norm must be weight or weight-affine")
    if not opt.synthetic and (opt.dataset is None or opt.dataroot is None):
        raise SystemExit("--dataset and --dataroot are required unless --synthetic is given")
    if opt.save_path is None and opt.load_path is None:
        raise ValueError("must specify save path if not continue training")
    if opt.save_path is None:
        opt.save_path = opt.load_path
    if not torch.cuda.is_available():
        raise SystemExit("this path needs a CUDA device; there is no CPU fallback")
    opt.augment, opt.final_test = "none", getattr(opt, "final_test", False)

    from common.model import build_discriminator, build_generator, build_reverser
    from glis_b200 import _lib, dp, ops
    from glis_b200.trainer import GraphedRIter, RIterTrainer

    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if opt.precision:
        _lib.set_precision(opt.precision)
    data_seed = dp.seed_everything(opt.seed, rank)

    # nets as r_iterative/main.py:195-205 (the reference passes 5 arguments to build_discriminator: no dropout in D)
    gen = build_generator(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.code_size, opt.norm).to(device)
    rev = build_reverser(opt.width, opt.height, opt.nfeature // 2, opt.nlayer, opt.code_size, opt.norm,
                         opt.spatial_dropout_r).to(device)
    dis = build_discriminator(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.norm).to(device)
    if rank == 0:
        print(gen)
        print(rev)
        print(dis)
    sync = dp.OverlappedGradSync(world) if world > 1 else None
    # the do_train schedule must agree across ranks: its own identically seeded generator
    tr = RIterTrainer(gen, rev, dis, lr=opt.lr, lambda_r=opt.lambda_r, r_iterations=opt.r_iterations,
                      rng=random.Random(opt.seed), grad_sync=sync, ls=opt.ls)
    hops = 1 + opt.r_iterations

    if rank == 0:
        for sub in ("", "samples", "net_archive", "log", "running_test"):
            os.makedirs(os.path.join(opt.save_path, sub), exist_ok=True)
    graphed = None if opt.no_graph else GraphedRIter(tr, opt.batch_size, opt.height, opt.width, opt.code_size, device)
    data = glis_main.SyntheticData(opt, device, data_seed) if opt.synthetic \
        else glis_main.DatasetData(opt, device, rank, world)
    first_code = graphed.first_code if graphed is not None else torch.empty(opt.batch_size, opt.code_size, device=device)

    state = {"current_iter": 0, "best_iter": 0, "min_loss": 1e100}
    history = new_history(opt.r_iterations)
    if opt.load_path is not None:
        loaded = load_state(opt.load_path, opt.net if opt.final_test else "last", tr)
        state.update({k: loaded[k] for k in ("current_iter", "best_iter", "min_loss") if k in loaded})
        h = loaded.get("history")
        if isinstance(h, (bytes, str)):
            from common.plotting import History
            history = History.from_string(h)
        data.restore(loaded)
        tr.rng = random.Random(opt.seed * 1000003 + int(state["current_iter"]))
        vis_code = torch.load(os.path.join(opt.load_path, "samples", "vis_code.pt")).to(device)
    else:
        vis_code = torch.randn(opt.vis_row * opt.vis_col, opt.code_size).to(device)
        if rank == 0:
            torch.save(vis_code.cpu(), os.path.join(opt.save_path, "samples", "vis_code.pt"))

    def visualize(it):
        """Chain of reconstructions from vis_code (r_iterative/main.py:256-283): G(z), G(R(G(z))), ..."""
        import torchvision
        modes = [(m, m.training) for m in (gen, rev)]
        for m, _ in modes:
            m.eval()
        with torch.no_grad():
            imgs = [gen(vis_code)]
            for _ in range(opt.r_iterations):
                imgs.append(gen(rev(imgs[-1])))
        for m, was in modes:
            m.train(was)
        for r_idx, img in enumerate(imgs):
            name = "sample_{0}.jpg".format(it) if r_idx == 0 else "sample_{0}_r{1}.jpg".format(it, r_idx)
            torchvision.utils.save_image(img * 2 - 1 if opt.output_scale else img,
                                         os.path.join(opt.save_path, "samples", name), nrow=opt.vis_row)

    def checkpoint(prefix, it):
        state["current_iter"] = it
        out = {k: state[k] for k in ("current_iter", "best_iter", "min_loss")}
        out.update(data.position())
        out["history"] = history.to_string()
        save_state(opt.save_path, prefix, tr, out)

    it = state["current_iter"]
    while it < opt.niter:
        t0 = time.time()
        it += 1
        flags = tr.draw_train_flags(opt.always_train_all)
        ops.randn_(first_code, data_seed + 7, it * glis_main.NoiseSource.STRIDE)
        reals = [data.next_batch().clone() if opt.synthetic else data.next_batch() for _ in range(sum(flags))]
        out = graphed.step(None, reals, flags) if graphed is not None else tr.step(first_code, reals, flags)
        if rank == 0 and it % opt.log_interval == 0:
            msg = ["%d |" % it]
            for r_idx, rec in enumerate(out):
                if rec is None:
                    msg.append("hop%d: -" % r_idx)
                    continue
                vals = {k: v.item() for k, v in rec.items()}
                msg.append("hop%d: g %.4f%s d-real %.4f d-fake %.4f" % (
                    r_idx, vals["g"], (" r %.4f" % vals["r"]) if "r" in vals else "", vals["d_real"], vals["d_fake"]))
                history.add_value("loss-g-mix", "train-g%d" % r_idx, it, vals["g"])
                history.add_value("loss-d-mix", "train-d-real%d" % r_idx, it, vals["d_real"])
                history.add_value("loss-d-mix", "train-d-fake%d" % r_idx, it, vals["d_fake"])
                if "r" in vals:
                    history.add_value("loss-r-mix", "train-r%d" % (r_idx - 1), it, min(max(vals["r"], 0.0), 1.0))
            msg.append("t:%.4fs" % (time.time() - t0))
            print(" ".join(msg))
        if rank == 0 and it % opt.vis_interval == 0:
            visualize(it)
        if rank == 0 and it % opt.save_interval == 0:
            checkpoint(it, it)
    if rank == 0:
        checkpoint("last", it)
    return hops


if __name__ == "__main__":
    main()
