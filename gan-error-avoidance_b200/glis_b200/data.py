"""Input pipeline for real datasets: decoded, cropped / scaled, augmented ``[0, 1]`` image batches on the device.

The reference fills every batch with a synchronous per-sample loop in the training thread
(g_lis/main.py:543-554: ``true_sample[i].copy_(get_data(train_index[index_shuffle[current_sample]]))``, PIL
decode + torchvision transforms + imgaug, then one ``.cuda()``).  With the device iteration at ~1.9 ms that loop
would be the whole run time.  Here:

* the SAME sample order — ``index_shuffle = randperm(N)``, consumed front to back, redrawn when it runs out
  (:494, :550-553) — drives a ``torch.utils.data.DataLoader`` with worker processes, so decoding runs ahead of
  training on as many cores as the host has; under data parallelism rank r owns ``train_index[r::world]``;
* batches land in pinned memory and cross PCIe on a copy stream into one of two device slots while the previous
  iteration computes (``next_batch`` hands out a batch whose copy is already in flight);
* augmentation (``--augment``, :176-231) runs on the DEVICE as one kernel per batch (``glis_augment``) from a
  per-image parameter row drawn on the host: horizontal flip (every set, also ``none``), and per set some of
  additive Gaussian noise, brightness multiply, contrast normalisation and a random affine map — the imgaug
  operators of the reference with their probabilities (0.5 each) and ranges.  Differences, stated: the operators
  apply in one fixed order (affine, multiply, contrast, noise) where imgaug shuffles them per batch, the affine
  map samples bilinearly (imgaug: order 3 for three of the four sets), contrast pivots on 0.5 like imgaug's
  ``ContrastNormalization`` (128/255).
* ``position()`` / ``restore()`` carry ``index_shuffle`` and ``current_sample`` of the reference's state file.
"""
import math
import random

import torch

from . import _lib as L

# per set: (noise sigma max or 0, multiply range, contrast range, affine: (scale range, rotate deg, translate frac, border))
AUGMENT_SETS = {
    "none": (0.0, None, None, None),
    "flowers102": (0.035, (0.9, 1.1), (0.9, 1.1), ((0.9, 1.1), 15.0, 0.0, 1)),
    "cifar10": (0.035, (0.9, 1.1), (0.9, 1.1), None),
    "10kcats": (0.0, (0.9, 1.1), (0.9, 1.1), ((0.9, 1.1), 15.0, 0.0, 1)),
    "lsun_churches": (0.035, (0.9, 1.1), (0.9, 1.1), ((0.9, 1.1), 1.0, 0.1, 0)),
}


def augment_params(name, n, height, width, rng):
    """(n, 12) float32 parameter rows of ``glis_augment`` for augmentation set ``name`` (g_lis/main.py:185-229)."""
    if name not in AUGMENT_SETS:
        raise Exception("--augment must be 'flowers102' or 'cifar10' or '10kcats' or 'lsun_churches' or 'none'")
    sigma_max, mul, contrast, affine = AUGMENT_SETS[name]
    rows = torch.zeros(n, 12)
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    for i in range(n):
        a = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
        border = 0.0
        if affine is not None and rng.random() < 0.5:
            (s_lo, s_hi), rot, trans, border = affine
            sx, sy = rng.uniform(s_lo, s_hi), rng.uniform(s_lo, s_hi)
            th = math.radians(rng.uniform(-rot, rot))
            tx, ty = rng.uniform(-trans, trans) * width, rng.uniform(-trans, trans) * height
            # forward map about the image centre: p' = R S (p - c) + c + t; the kernel wants its inverse
            c, s = math.cos(th), math.sin(th)
            m00, m01, m10, m11 = c * sx, -s * sy, s * sx, c * sy
            det = m00 * m11 - m01 * m10
            i00, i01, i10, i11 = m11 / det, -m01 / det, -m10 / det, m00 / det
            a = [i00, i01, cx - i00 * (cx + tx) - i01 * (cy + ty), i10, i11, cy - i10 * (cx + tx) - i11 * (cy + ty)]
        m = rng.uniform(*mul) if (mul is not None and rng.random() < 0.5) else 1.0
        al = rng.uniform(*contrast) if (contrast is not None and rng.random() < 0.5) else 1.0
        sg = rng.uniform(0.0, sigma_max) if (sigma_max > 0 and rng.random() < 0.5) else 0.0
        flip = 1.0 if rng.random() < 0.5 else 0.0          # transforms.RandomHorizontalFlip(): every set
        rows[i] = torch.tensor(a + [m, al, sg, flip, float(border), 0.0])
    return rows


def augment(batch_nchw, params, seed, out=None):
    """``glis_augment`` on a device batch (N, C, H, W) in [0, 1]; returns the NHWC-dense (channels_last) result."""
    n, c, h, w = batch_nchw.shape
    if out is None:
        out = torch.empty((n, c, h, w), device=batch_nchw.device, dtype=torch.float32).contiguous(
            memory_format=torch.channels_last)
    L.call("glis_augment", L.ptr(batch_nchw.contiguous()), L.ptr(out), L.ptr(params.contiguous()), n, c, h, w, int(seed),
           L.stream())
    return out


class ShuffledOrder(object):
    """The reference's sample order (g_lis/main.py:494, :550-553) as an endless batch sampler: a permutation of the
    (rank's) training indices consumed front to back and redrawn when exhausted.  Remembers the permutation every
    batch came from, so the position of the batch the TRAINER holds — not the one a worker is decoding — can be
    written into a checkpoint."""

    def __init__(self, n, batch_size, seed, shuffle=None, current=0):
        self.n, self.batch_size = int(n), int(batch_size)
        self.gen = torch.Generator().manual_seed(int(seed))
        usable = shuffle is not None and shuffle.numel() == self.n
        self.perm = shuffle.clone() if usable else torch.randperm(self.n, generator=self.gen)
        self.current = int(current) % self.n if usable else 0
        self.initial = (self.perm, self.current)
        self.log = []                   # (perm, position after the batch) per issued batch

    def __iter__(self):
        while True:
            idx = []
            for _ in range(self.batch_size):
                idx.append(int(self.perm[self.current]))
                self.current += 1
                if self.current == self.n:
                    self.current, self.perm = 0, torch.randperm(self.n, generator=self.gen)
            self.log.append((self.perm, self.current))
            yield idx


class _Indexed(torch.utils.data.Dataset):
    def __init__(self, dataset, index):
        self.dataset, self.index = dataset, index

    def __len__(self):
        return len(self.index)

    def __getitem__(self, k):
        item = self.dataset[int(self.index[k])]
        return item[0] if isinstance(item, (tuple, list)) else item


class PrefetchLoader(object):
    """Endless stream of device batches from ``dataset`` (items: (3, H, W) float tensors in [0, 1], or (image, label)).

    ``train_index``: the global training indices (``data_index.pt['train']``); rank r of ``world`` owns every
    world-th one.  ``workers`` decoder processes fill pinned batches; two device slots; ``next_batch()`` returns the
    next (augmented, NHWC) batch and immediately starts the copy of the following one."""

    def __init__(self, dataset, train_index, batch_size, device, rank=0, world=1, workers=4, augment_set="none",
                 seed=0, shuffle=None, current=0, out=None):
        self.device, self.batch_size, self.augment_set = device, batch_size, augment_set
        self.index = train_index[rank::world]
        self.order = ShuffledOrder(len(self.index), batch_size, seed * 7919 + rank, shuffle, current)
        self.loader = torch.utils.data.DataLoader(
            _Indexed(dataset, self.index), batch_sampler=self.order, num_workers=workers, pin_memory=True,
            persistent_workers=workers > 0, prefetch_factor=4 if workers > 0 else None)
        self.it = iter(self.loader)
        self.rng = random.Random(seed * 104729 + rank)
        self.seed, self.count = seed * 15485863 + rank, 0
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots, self.ready, self.free = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]
        self.out = out                   # optional fixed NHWC destination (the graph's static input)
        self.delivered = 0               # batches handed to the trainer
        self._stage(0)

    def _stage(self, k):
        """Fetch the next decoded batch (blocks only if the workers fell behind) and start its H2D copy into slot k."""
        host = next(self.it)
        if self.free[k] is not None:
            self.copy_stream.wait_event(self.free[k])          # the iteration that read slot k has been enqueued
        with torch.cuda.stream(self.copy_stream):
            self.slots[k] = host.to(self.device, non_blocking=True)
            params = augment_params(self.augment_set, host.shape[0], host.shape[2], host.shape[3], self.rng)
            self.slots[k] = (self.slots[k], params.to(self.device, non_blocking=True))
            self.ready[k].record(self.copy_stream)

    def next_batch(self):
        k = self.count & 1
        main = torch.cuda.current_stream()
        main.wait_event(self.ready[k])
        raw, params = self.slots[k]
        self.count += 1
        batch = augment(raw, params, self.seed + self.count, out=self.out)       # also NCHW -> NHWC
        done = torch.cuda.Event()
        done.record(main)
        self.free[k] = done
        self.delivered += 1
        self._stage(k ^ 1)
        return batch

    def position(self):
        """``index_shuffle`` / ``current_sample`` (g_lis/main.py:350-357) as of the last batch handed out."""
        perm, cur = self.order.initial if self.delivered == 0 else self.order.log[self.delivered - 1]
        return {"index_shuffle": perm.clone(), "current_sample": cur}
