"""Train a G-LIS GAN on the B200 kernels — the command line of the reference's g_lis/main.py.

    python g_lis/main.py --dataset folder --dataroot /data/celeba --crop_size 160 --image_size 80 \\
        --code_size 256 --norm weight --lr 0.00002 --r_iterations 1 --niter 300000 --save_path /ckpt/exp01
    python g_lis/main.py --synthetic --image_size 80 --code_size 256 --norm weight --lr 0.00002 \\
        --r_iterations 1 --batch_size 64 --niter 100 --save_path /tmp/exp

Every flag of the reference script (g_lis/main.py:41-166) is accepted with the same name, type
and default.  Added: ``--synthetic`` (uniform [0,1) images generated on the device; no dataset
needed), ``--seed``, ``--precision fp32|bf16x3|bf16``, ``--no_graph``, ``--log_interval``; under
``torchrun`` the batch is sharded over the ranks (``--batch_size`` is per GPU; every rank draws its own latent
codes and images, the LIS depths agree across ranks).  ``--ls`` and ``--d_dropout`` run on the kernels too.

What runs on the device is one `GLISTrainer.step` per iteration (g_lis/main.py:526-589).  Host-side
extras of the reference that are out of this path's scope (loss plots, t-SNE, Inception score)
are not reproduced; sample grids, checkpoints (`net_archive/{prefix}_{gen,gen_opt,dis,dis_opt,
state}.pt`, reference key names) and the latent-reconstruction test are.
"""
from __future__ import print_function

import argparse
import contextlib
import os
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    a = p.add_argument
    a("--dataset", default=None, help="cifar10 | lsun | imagenet | folder | lfw (not needed with --synthetic)")
    a("--lsun_class", default="bedroom", help="class of lsun dataset to use")
    a("--dataroot", default=None, help="path to dataset")
    a("--batch_size", type=int, default=32, help="input batch size (per GPU)")
    a("--image_size", type=int, default=-1, help="image size")
    a("--width", type=int, default=-1, help="image width")
    a("--height", type=int, default=-1, help="image height")
    a("--crop_size", type=int, default=-1, help="crop size before scaling")
    a("--crop_width", type=int, default=-1, help="crop width before scaling")
    a("--crop_height", type=int, default=-1, help="crop height before scaling")
    a("--code_size", type=int, default=128, help="size of latent code")
    a("--nfeature", type=int, default=64, help="number of features of first conv layer")
    a("--nlayer", type=int, default=-1, help="number of down/up conv layers")
    a("--norm", default="none", help="type of normalization: none | batch | weight | weight-affine")
    a("--save_path", default=None, help="path to save generated files")
    a("--load_path", default=None, help="load to continue existing experiment")
    a("--lr", type=float, default=0.0001, help="learning rate")
    a("--test_interval", type=int, default=10000, help="how often to test reconstruction")
    a("--test_lr", type=float, default=0.01, help="learning rate for reconstruction test")
    a("--test_steps", type=int, default=50, help="number of steps in running reconstruction test")
    a("--vis_interval", type=int, default=2000, help="how often to save generated samples")
    a("--vis_size", type=int, default=10, help="size of visualization grid")
    a("--vis_row", type=int, default=-1, help="height of visualization grid")
    a("--vis_col", type=int, default=-1, help="width of visualization grid")
    a("--save_interval", type=int, default=5000, help="how often to save network")
    a("--niter", type=int, default=50000, help="number of iterations to train")
    a("--final_test", action="store_true", default=False, help="do final test")
    a("--ls", action="store_true", default=False, help="use LSGAN")
    a("--output_scale", action="store_true", default=False, help="save x*2-1 instead of x when saving image")
    a("--net", default="best", help="network to load for final test: best | last | <niter>")
    a("--lambda_r", type=float, default=0.9, help="strength of MSE on R")
    a("--spatial_dropout_r", type=float, default=0, help="Spatial dropout applied to R")
    a("--r_iterations", type=int, default=3, help="how many LIS modules to use in G")
    a("--always_train_all", action="store_true", default=False, help="always train with all LIS modules")
    a("--load_tolerant", action="store_true", default=False, help="load G's state dict tolerantly")
    a("--nb_cache_total", type=int, default=0, help="size of the dataset cache")
    a("--nb_cache_lists", type=int, default=1, help="number of caches to use")
    a("--cache_p_drop", type=float, default=0.1, help="chance to drop a data entry from the cache")
    a("--augment", default="none", help="name of augmentation set to use")
    a("--g_upscaling", default="fractional", help="upscaling method to use in G: fractional|nearest|bilinear")
    a("--d_dropout", type=float, default=0, help="dropout probability to use in D before the last layer")
    # additive
    a("--synthetic", action="store_true", default=False, help="train on device-generated U[0,1) images")
    a("--seed", type=int, default=1234, help="seed of weights / LIS depth (shared by ranks) and data (per rank)")
    a("--precision", default=None, choices=["fp32", "bf16x3", "bf16"], help="contraction arithmetic")
    a("--no_graph", action="store_true", default=False, help="do not replay the step as a CUDA graph")
    a("--log_interval", type=int, default=1, help="print (and synchronise for) the losses every N iterations")
    a("--workers", type=int, default=min(8, os.cpu_count() or 1), help="decoder processes of the input pipeline")
    return p


def resolve_geometry(opt):
    """--image_size / --width / --height / --crop_* and the automatic level count (g_lis/main.py:169-251)."""
    if opt.width <= 0:
        opt.width = opt.image_size
    if opt.height <= 0:
        opt.height = opt.image_size
    if opt.width <= 0 or opt.height <= 0:
        raise ValueError("must specify valid image size")
    if opt.crop_width <= 0:
        opt.crop_width = opt.crop_size
    if opt.crop_height <= 0:
        opt.crop_height = opt.crop_size
    if opt.vis_row <= 0 or opt.vis_col <= 0:
        opt.vis_row = opt.vis_col = opt.vis_size
    if opt.nlayer < 0:
        opt.nlayer, s = 0, max(opt.width, opt.height)
        while s >= 8:
            s = (s + 1) // 2
            opt.nlayer += 1
    return opt


class NoiseSource(object):
    """The two latent batches of an iteration (``torch.randn(B, code)`` at g_lis/main.py:561,576) from the device
    Philox generator, keyed by the PER-RANK data seed and the iteration number: data-parallel ranks draw different
    codes (N ranks x B samples are N*B distinct samples), nothing crosses PCIe, and a resumed run continues the
    stream where the checkpoint left it."""

    STRIDE = 1 << 22          # Philox counter blocks reserved per iteration (4 floats each: B*code <= 16 M)

    def __init__(self, batch, code, device, data_seed):
        from glis_b200 import ops
        self.ops, self.seed = ops, int(data_seed)
        self.z_d = torch.empty(batch, code, device=device)
        self.z_g = torch.empty(batch, code, device=device)

    @staticmethod
    def streams(data_seed, iteration):
        """((seed, offset) of z_d, (seed, offset) of z_g) for one iteration."""
        off = int(iteration) * NoiseSource.STRIDE
        return (int(data_seed) + 7, off), (int(data_seed) + 13, off)

    def draw(self, iteration):
        (sd, od), (sg, og) = self.streams(self.seed, iteration)
        self.ops.randn_(self.z_d, sd, od)
        self.ops.randn_(self.z_g, sg, og)
        return self.z_d, self.z_g


@contextlib.contextmanager
def private_depth_rng(gen):
    """Run forward-only side work (sample grids, the reconstruction test) without touching the training run's
    LIS-depth draws: ``GeneratorLearnedInputSpace.lis_depth`` consumes one ``rng.random()`` per module even when
    the depth is forced (as the reference does, common/model.py:286-294), and rank 0 alone draws sample grids — on
    the shared stream that would shift rank 0's later depths against the other ranks'."""
    saved, gen.rng = gen.rng, random.Random(0)
    try:
        yield
    finally:
        gen.rng = saved


class SyntheticData(object):
    """U[0,1) image batches from the device Philox generator (the benchmark's data source)."""

    def __init__(self, opt, device, seed, out=None):
        from glis_b200 import ops
        self.ops, self.seed, self.calls = ops, seed, 0
        self.buf = out if out is not None else torch.empty(
            opt.batch_size, 3, opt.height, opt.width, device=device).contiguous(memory_format=torch.channels_last)
        self.test = torch.rand(min(64, opt.vis_row * opt.vis_col), 3, opt.height, opt.width,
                               generator=torch.Generator().manual_seed(seed)).to(device)

    def next_batch(self):
        self.calls += 1
        return self.ops.uniform_(self.buf, self.seed, self.calls * (1 << 24))

    def position(self):
        return {"index_shuffle": torch.zeros(0, dtype=torch.long), "current_sample": self.calls}

    def restore(self, state):
        self.calls = int(state.get("current_sample", 0))


class DatasetData(object):
    """The reference's data source (g_lis/main.py:169-297, :543-554) — CenterCrop / Scale / CenterCrop / ToTensor,
    the `data_index.pt` split, the `index_shuffle` sample order — behind `glis_b200.data.PrefetchLoader`: decoded by
    `--workers` processes, copied through pinned memory on a copy stream, augmented (`--augment`, horizontal flip
    included) on the device, sharded by rank."""

    def __init__(self, opt, device, rank, world, seed, out=None):
        import torchvision.datasets as datasets
        import torchvision.transforms as transforms
        from glis_b200.data import PrefetchLoader
        tl = []
        if opt.crop_height > 0 and opt.crop_width > 0:
            tl.append(transforms.CenterCrop((opt.crop_height, opt.crop_width)))
        tl += [transforms.Resize((opt.height, opt.width)), transforms.CenterCrop((opt.height, opt.width)),
               transforms.ToTensor()]
        tf = transforms.Compose(tl)
        if opt.dataset == "cifar10":
            parts = [datasets.CIFAR10(root=opt.dataroot, download=False, transform=tf),
                     datasets.CIFAR10(root=opt.dataroot, train=False, transform=tf)]
            self.dataset = torch.utils.data.ConcatDataset(parts)
        elif opt.dataset in ("imagenet", "folder", "lfw"):
            self.dataset = datasets.ImageFolder(root=opt.dataroot, transform=tf)
        elif opt.dataset == "lsun":
            self.dataset = datasets.LSUN(opt.dataroot, classes=[opt.lsun_class + "_train"], transform=tf)
        else:
            raise ValueError("unknown --dataset %r" % (opt.dataset,))
        index = torch.load(os.path.join(opt.dataroot, "data_index.pt"))
        self.train_index = index["train"]
        self.test_index = index["final_test" if opt.final_test else "running_test"]
        self.make = lambda shuffle, current: PrefetchLoader(
            self.dataset, self.train_index, opt.batch_size, device, rank, world, workers=opt.workers,
            augment_set=opt.augment, seed=seed, shuffle=shuffle, current=current, out=out)
        self.loader = None
        self._resume = (None, 0)
        n_test = min(self.test_index.size(0), opt.vis_row * opt.vis_col)
        self.test = torch.stack([self.dataset[int(self.test_index[i])][0] for i in range(n_test)]).to(device)

    def next_batch(self):
        if self.loader is None:
            self.loader = self.make(*self._resume)
        return self.loader.next_batch()

    def position(self):
        """``index_shuffle`` / ``current_sample`` of the reference's state file (g_lis/main.py:350-357)."""
        if self.loader is None:
            self.loader = self.make(*self._resume)
        return self.loader.position()

    def restore(self, state):
        sh = state.get("index_shuffle")
        if sh is not None and sh.numel() > 0:        # (a permutation of THIS rank's share; one of another size is ignored)
            self._resume = (sh.clone(), int(state.get("current_sample", 0)))


def reconstruction_test(gen, targets, opt):
    """Latent-reconstruction error of held-out images (g_lis/main.py:398-453): 50 RMSprop steps on
    the code with G frozen; returns the mean MSE."""
    was_training = gen.training
    gen.eval()
    for p in gen.parameters():
        p.requires_grad_(False)
    total = 0.0
    with private_depth_rng(gen):
        total = _reconstruction_loop(gen, targets, opt)
    for p in gen.parameters():
        p.requires_grad_(True)
    gen.train(was_training)
    return total / max(1, targets.size(0))


def _reconstruction_loop(gen, targets, opt):
    total = 0.0
    for i in range(0, targets.size(0), opt.batch_size):
        tgt = targets[i:i + opt.batch_size]
        code = torch.zeros(tgt.size(0), opt.code_size, device=tgt.device, dtype=tgt.dtype, requires_grad=True)
        test_opt = torch.optim.RMSprop([code], lr=opt.test_lr, eps=1e-6, alpha=0.9)
        for _ in range(opt.test_steps):
            out, _ = gen(code)
            loss = torch.nn.functional.mse_loss(out, tgt)
            test_opt.zero_grad()
            loss.backward()
            test_opt.step()
        with torch.no_grad():
            out, _ = gen(code)
            total += torch.nn.functional.mse_loss(out, tgt).item() * tgt.size(0)
    return total


def new_history(r_iterations):
    """The loss-history groups of g_lis/main.py:316-319."""
    from common.plotting import History
    h = History()
    h.add_group("loss-r-mix", ["train-r%d" % i for i in range(r_iterations)], increasing=False)
    h.add_group("loss-g-mix", ["train-g%d" % i for i in range(1 + r_iterations)], increasing=False)
    h.add_group("loss-d-mix", ["train-d-real"] + ["train-d-fake%d" % i for i in range(1 + r_iterations)], increasing=False)
    return h


def state_for_saving(state, history, data):
    """The reference's ``*_state.pt`` (g_lis/main.py:350-357): index_shuffle, current_iter, best_iter, min_loss,
    current_sample and the PICKLED history."""
    out = {k: state[k] for k in ("current_iter", "best_iter", "min_loss")}
    out.update(data.position())
    out["history"] = history.to_string()
    return out


def history_from_state(state, r_iterations):
    """History of a loaded state file: the reference's pickled string (python-2 pickles included), a History
    object, or the plain list round-1 checkpoints of this repository carried."""
    from common.plotting import History
    h = state.get("history")
    if isinstance(h, (bytes, str)):
        return History.from_string(h)
    if isinstance(h, History):
        return h
    hist = new_history(r_iterations)
    for rec in (h or []):
        it, d_real, d_fake, g, r = rec
        hist.add_value("loss-d-mix", "train-d-real", it, d_real)
        hist.add_value("loss-d-mix", "train-d-fake0", it, d_fake)
        hist.add_value("loss-g-mix", "train-g0", it, g)
        for i, v in enumerate(r):
            hist.add_value("loss-r-mix", "train-r%d" % i, it, min(max(v, 0.0), 1.0))
    return hist


def save_state(path, prefix, trainer, state):
    d = os.path.join(path, "net_archive")
    torch.save(trainer.gen.state_dict(), os.path.join(d, "{0}_gen.pt".format(prefix)))
    torch.save(trainer.gen_flat.optimizer_state_dict(trainer.lr), os.path.join(d, "{0}_gen_opt.pt".format(prefix)))
    torch.save(trainer.dis.state_dict(), os.path.join(d, "{0}_dis.pt".format(prefix)))
    torch.save(trainer.dis_flat.optimizer_state_dict(trainer.lr), os.path.join(d, "{0}_dis_opt.pt".format(prefix)))
    torch.save(state, os.path.join(d, "{0}_state.pt".format(prefix)))


def load_state(path, prefix, trainer, tolerant=False):
    d = os.path.join(path, "net_archive")
    gen_sd = torch.load(os.path.join(d, "{0}_gen.pt".format(prefix)), map_location="cpu")
    if tolerant:
        own = trainer.gen.state_dict()
        own.update({k: v for k, v in gen_sd.items() if k in own})
        gen_sd = own
    trainer.gen.load_state_dict(gen_sd)
    trainer.dis.load_state_dict(torch.load(os.path.join(d, "{0}_dis.pt".format(prefix)), map_location="cpu"))
    trainer.gen_flat.load_optimizer_state_dict(torch.load(os.path.join(d, "{0}_gen_opt.pt".format(prefix)),
                                                          map_location="cpu"))
    trainer.dis_flat.load_optimizer_state_dict(torch.load(os.path.join(d, "{0}_dis_opt.pt".format(prefix)),
                                                          map_location="cpu"))
    return torch.load(os.path.join(d, "{0}_state.pt".format(prefix)), map_location="cpu")


def main(argv=None):
    opt = resolve_geometry(build_parser().parse_args(argv))
    if opt.norm not in ("weight", "weight-affine"):
        raise SystemExit("g_lis/main.py (B200 path): --norm must be weight or weight-affine")
    if not opt.synthetic and (opt.dataset is None or opt.dataroot is None):
        raise SystemExit("--dataset and --dataroot are required unless --synthetic is given")
    if opt.save_path is None and opt.load_path is None:
        raise ValueError("must specify save path if not continue training")
    if opt.save_path is None:
        opt.save_path = opt.load_path
    if not torch.cuda.is_available():
        raise SystemExit("this path needs a CUDA device; there is no CPU fallback")

    from common.model import GeneratorLearnedInputSpace, build_discriminator
    from glis_b200 import _lib, dp
    from glis_b200.trainer import GLISTrainer, GraphedStep

    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if opt.precision:
        _lib.set_precision(opt.precision)
    data_seed = dp.seed_everything(opt.seed, rank)

    gen = GeneratorLearnedInputSpace(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.code_size, opt.norm,
                                     n_lis_layers=opt.r_iterations, upscaling=opt.g_upscaling).to(device)
    dis = build_discriminator(opt.width, opt.height, opt.nfeature, opt.nlayer, opt.norm, opt.d_dropout).to(device)
    # the stochastic LIS depth must agree across ranks: its draws come from a generator seeded identically
    # everywhere that nothing but the training loop touches (private_depth_rng shields it from sample grids / tests)
    gen.rng = random.Random(opt.seed)
    if rank == 0:
        print(gen)
        print(dis)
    sync = dp.OverlappedGradSync(world) if world > 1 else None
    trainer = GLISTrainer(gen, dis, lr=opt.lr, lambda_r=opt.lambda_r, grad_sync=sync, ls=opt.ls)

    if rank == 0:
        for sub in ("", "samples", "net_archive", "log", "running_test"):
            os.makedirs(os.path.join(opt.save_path, sub), exist_ok=True)
    use_graph = not opt.no_graph
    graphed = GraphedStep(trainer, opt.batch_size, opt.height, opt.width, opt.code_size, device) if use_graph else None
    # synthetic images are generated straight into the step's static input (the first half of D's 2B batch)
    static_in = graphed.real if graphed is not None else None
    data = SyntheticData(opt, device, data_seed, out=static_in) if opt.synthetic \
        else DatasetData(opt, device, rank, world, data_seed, out=static_in)
    noise = NoiseSource(opt.batch_size, opt.code_size, device, data_seed)

    state = {"current_iter": 0, "best_iter": 0, "min_loss": 1e100}
    history = new_history(opt.r_iterations)
    if opt.load_path is not None:
        loaded = load_state(opt.load_path, opt.net if opt.final_test else "last", trainer, opt.load_tolerant)
        state.update({k: loaded[k] for k in ("current_iter", "best_iter", "min_loss") if k in loaded})
        history = history_from_state(loaded, opt.r_iterations)
        data.restore(loaded)
        gen.rng = random.Random(opt.seed * 1000003 + int(state["current_iter"]))   # same on every rank
        vis_code = torch.load(os.path.join(opt.load_path, "samples", "vis_code.pt")).to(device)
    else:
        vis_code = torch.randn(opt.vis_row * opt.vis_col, opt.code_size).to(device)
        if rank == 0:
            torch.save(vis_code.cpu(), os.path.join(opt.save_path, "samples", "vis_code.pt"))
    if opt.final_test:
        print("loss = {0}".format(reconstruction_test(gen, data.test, opt)))
        return

    forced = opt.r_iterations if opt.always_train_all else None   # the reference's `opr` typo, fixed (App. D)

    def visualize(it):
        import torchvision
        was = gen.training
        gen.eval()
        with torch.no_grad(), private_depth_rng(gen):
            img, _ = gen(vis_code, n_execute_lis_layers="all")
        gen.train(was)
        torchvision.utils.save_image(img * 2 - 1 if opt.output_scale else img,
                                     os.path.join(opt.save_path, "samples", "sample_{0}.jpg".format(it)),
                                     nrow=opt.vis_row)

    def checkpoint(prefix, it):
        state["current_iter"] = it
        save_state(opt.save_path, prefix, trainer, state_for_saving(state, history, data))

    it = state["current_iter"]
    while it < opt.niter:
        t0 = time.time()
        it += 1
        real = data.next_batch()
        z_d, z_g = noise.draw(it)
        depth_d, depth_g = gen.lis_depth(forced), gen.lis_depth(forced)
        if graphed is not None:
            in_place = real.data_ptr() == graphed.real.data_ptr()
            out = graphed.step(None if in_place else real, z_d, z_g, depth_d, depth_g)
        else:
            out = trainer.step(real, z_d, z_g, depth_d, depth_g)
        if rank == 0 and it % opt.log_interval == 0:
            vals = {k: out[k].item() for k in ("d_real", "d_fake", "g")}   # the only host synchronisation
            r = [v.item() for v in out["r"]]
            msg = ["%d |" % it, "d-real: %.4f" % vals["d_real"], "d-fake%d: %.4f" % (depth_d, vals["d_fake"]),
                   "g%d: %.4f" % (depth_g, vals["g"])]
            msg += ["r%d: %.4f" % (i, v) for i, v in enumerate(r)]
            msg.append("t:%.4fs" % (time.time() - t0))
            print(" ".join(msg))
            # the reference's history lines (g_lis/main.py:597-613): the index of a fake / g line is the LIS depth
            history.add_value("loss-d-mix", "train-d-real", it, vals["d_real"])
            history.add_value("loss-d-mix", "train-d-fake%d" % depth_d, it, vals["d_fake"])
            history.add_value("loss-g-mix", "train-g%d" % depth_g, it, vals["g"])
            for i, v in enumerate(r):
                history.add_value("loss-r-mix", "train-r%d" % i, it, min(max(v, 0.0), 1.0))
        if rank == 0 and it % opt.vis_interval == 0:
            visualize(it)
        if it % opt.test_interval == 0:
            loss = reconstruction_test(gen, data.test, opt)
            if rank == 0:
                print("Testing ... loss = {0}".format(loss))
                if loss < state["min_loss"]:
                    state["min_loss"], state["best_iter"] = loss, it
                    checkpoint("best", it)
                checkpoint("last", it)
        if rank == 0 and it % opt.save_interval == 0:
            checkpoint(it, it)
    if rank == 0:
        checkpoint("last", it)


if __name__ == "__main__":
    main()
