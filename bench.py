#!/usr/bin/env python
"""Benchmark of the G-LIS training step (BASELINE.json metric):

    G-LIS train images/sec at 80x80, batch 64 per GPU (config 2: nfeature 64, 4 levels,
    code 256, 1 LIS module, lr 2e-5, lambda_r 0.9), synthetic data, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full iteration of g_lis/main.py:526-589 (D-real, D-fake, D update, G+LIS
update).  Prints ONE JSON line (rank 0).  `value` is measured with the inputs already in
HBM; `e2e` drives the same step from pinned host buffers (H2D of the image batch and both
noise batches every step, D2H of the three losses).  `roofline` describes the dominant
kernel family timed live with CUDA events; `cpu_baseline` is the oracle's step on the
host cores (N=1 only).  `--impl reference` times that CPU oracle alone (the reference's
python-2.7 / PyTorch@065c5986 pin cannot be installed offline; torch 2.11 CPU is the
nearest installable build — SURVEY.md §8c).
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gan-error-avoidance_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# BASELINE.json `configs`, with the algorithmic GFLOP per step of SURVEY.md §8d (F_step = 8 F_D + 4 F_G + 4 F_LIS;
# config 5a: 21 F_G + 38 F_D + 9 F_R per outer iteration).  The driver's line is config 2 (the default); the others
# are reachable with --config for the configs table of DESIGN.md.
CONFIGS = {
    "1": dict(W=32, H=32, B=32, nfeature=64, nlayer=3, code=256, n_lis=1, gflop=13.79,
              workload="G-LIS 1 LIS module, 32x32 CIFAR-shaped synthetic, batch 32 (BASELINE configs[0])"),
    "2": dict(W=80, H=80, B=64, nfeature=64, nlayer=4, code=256, n_lis=1, gflop=250.9,
              workload="G-LIS 1 LIS module, 80x80 CelebA-shaped synthetic, batch 64/GPU (BASELINE configs[1])"),
    "3": dict(W=80, H=80, B=64, nfeature=64, nlayer=4, code=256, n_lis=3, gflop=251.0,
              workload="G-LIS 3 LIS modules, 80x80 synthetic, batch 64/GPU (BASELINE configs[2])"),
    "4": dict(W=160, H=160, B=32, nfeature=64, nlayer=5, code=256, n_lis=1, gflop=661.1,
              workload="G-LIS 1 LIS module at 160x160 (extra G/D layer), batch 32/GPU (BASELINE configs[3])"),
    "5a": dict(W=80, H=80, B=64, nfeature=64, nlayer=4, code=256, n_lis=0, gflop=1283.7, riter=3,
               workload="R-iterative 3 iterations (--always_train_all), 80x80 synthetic, batch 64/GPU (BASELINE configs[4], first half)"),
    "5b": dict(W=64, H=64, B=64, nfeature=64, nlayer=3, code=256, n_lis=1, gflop=153.1, upscaling="nearest", d_dropout=0.2,
               workload="G-LIS 64x64 3-layer NN-upsampling G/D with dropout 0.2 in D, batch 64/GPU (BASELINE configs[4], second half)"),
}
for _c in CONFIGS.values():
    _c.setdefault("upscaling", "fractional")
    _c.setdefault("d_dropout", 0)
    _c.setdefault("riter", 0)
    _c.update(lr=2e-5, lambda_r=0.9)
CFG = dict(CONFIGS["2"])
WORKLOAD = CFG["workload"]
GFLOP_PER_STEP = CFG["gflop"]
SEED = 1234


def select_config(name):
    global WORKLOAD, GFLOP_PER_STEP
    CFG.clear()
    CFG.update(CONFIGS[name])
    WORKLOAD, GFLOP_PER_STEP = CFG["workload"], CFG["gflop"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                              ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_oracle_pair():
    import oracle
    torch.manual_seed(SEED)
    g = oracle.GeneratorLearnedInputSpace(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], CFG["code"], "weight",
                                          CFG["n_lis"], CFG["upscaling"])
    d = oracle.build_discriminator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], "weight", CFG["d_dropout"])
    return g, d


def build_oracle_riter():
    import oracle
    torch.manual_seed(SEED)
    g = oracle.build_generator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], CFG["code"], "weight")
    r = oracle.build_reverser(CFG["W"], CFG["H"], CFG["nfeature"] // 2, CFG["nlayer"], CFG["code"], "weight", 0)
    d = oracle.build_discriminator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], "weight", 0)
    return g, r, d


def time_cpu_oracle(steps, warmup, budget_s=150.0):
    """(seconds per step, cores, images per step) of the oracle's iteration on all host cores.

    A step is one full config-2 iteration at batch 64; if `steps + warmup` of those would not fit
    `budget_s`, the per-step batch is cut (never below 8) so that the run stays bounded — images/s
    is then per-step images over per-step time, which is what the metric means."""
    from oracle.step import GLISOracleTrainer, riter_iteration
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(SEED + 1)
    if CFG["riter"]:
        g, r, d = build_oracle_riter()
        states = ({}, {}, {})

        def one(B):
            z = torch.randn(B, CFG["code"], generator=gen)
            reals = [torch.rand(B, 3, CFG["H"], CFG["W"], generator=gen) for _ in range(1 + CFG["riter"])]
            t0 = time.perf_counter()
            riter_iteration(g, r, d, states[0], states[1], states[2], z, reals, CFG["lr"], CFG["lambda_r"], CFG["riter"])
            return time.perf_counter() - t0
    else:
        g, d = build_oracle_pair()
        tr = GLISOracleTrainer(g, d, lr=CFG["lr"], lambda_r=CFG["lambda_r"])

        def one(B):
            real = torch.rand(B, 3, CFG["H"], CFG["W"], generator=gen)
            zd, zg = torch.randn(B, CFG["code"], generator=gen), torch.randn(B, CFG["code"], generator=gen)
            t0 = time.perf_counter()
            tr.step(real, zd, zg, CFG["n_lis"], CFG["n_lis"])
            return time.perf_counter() - t0

    B = CFG["B"]
    t_first = one(B)                      # doubles as the first warm-up step
    total = (steps + max(warmup, 1)) * t_first
    if total > budget_s:
        B = max(8, int(B * budget_s / total) // 8 * 8)
    times = []
    for i in range(max(warmup, 1) - 1 + steps):
        t = one(B)
        if i >= max(warmup, 1) - 1:
            times.append(t)
    return sum(times) / len(times), cores, B


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def bench_config():
    """The `config` object of BOTH arms' JSON lines (identical by construction: the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": CFG["B"], "lis_depth": "all" if not CFG["riter"] else "n/a",
            "gflop_per_step": GFLOP_PER_STEP,
            "l2": "inputs larger than L2: every step touches > 200 MB of activations and all weights (126 MB L2)"}


def metric_name():
    if CFG["workload"] == CONFIGS["2"]["workload"]:
        return "G-LIS train images/sec at 80x80 bs64/GPU"
    return "train images/sec, " + CFG["workload"]


def run_reference(args, rank):
    if rank != 0:
        return
    sec, cores, b_step = time_cpu_oracle(args.steps, args.warmup)
    ips = b_step / sec
    sample = "%d full iterations of the workload (batch %d per step) after %d warm-up, oracle port on torch %s CPU, %s" % (
        args.steps, b_step, args.warmup, torch.__version__, cpu_model())
    print(json.dumps({
        "impl": "reference", "metric": metric_name(), "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "note": "python2.7/PyTorch@065c5986 pin not installable offline; nearest installable PyTorch CPU build used "
                "(oracle port of the reference step); batch per timed step: %d" % b_step,
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def probe_dominant_kernel(dev):
    """Time the dominant kernel alone: `tc_conv_kernel` on the discriminator's level-1 convolution
    of the 2B-image D pass (64 -> 128 channels, 40x40 -> 20x20, 128 images: M=51200, N=128, K=1024).
    20 launches replayed as one CUDA graph, CUDA events on the launching stream."""
    from glis_b200 import _lib as L, ops
    if L.default_precision == L.PREC_FP32:
        return None
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    n, ci, co, hi, ho = 128, 64, 128, 40, 20
    x = torch.rand(n, ci, hi, hi, device=dev).contiguous(memory_format=torch.channels_last)
    ops.attach_planes(x, ops.split_bf16(x, L.default_precision == L.PREC_BF16X3))
    w = torch.nn.Parameter((torch.rand(co, ci, 4, 4, device=dev) - 0.5) * 0.06)
    pw = ops.PackedWeights(w, None, spec)

    def run():
        return ops.launch(spec, L.CONV, x, (n, co, ho, ho), pw, forward_pack=True)[0]
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    reps = 20
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            run()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * n * ho * ho * co * ci * 16
    passes = 3 if L.default_precision == L.PREC_BF16X3 else 1
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic, traffic_src = tj.get("tc_conv_kernel_d1_2B_dram_bytes"), tj.get("captured_at")
    return {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "traffic": traffic,
            "traffic_source": traffic_src,
            "kernel": "tc_conv_halo_kernel<64> (csrc/tc_conv2.cu)  D level-1 conv 64->128, 40x40->20x20, 128 images "
                      "(M=51200 N=128 K=1024)",
            "ms_per_launch": ms, "algorithmic_gflop_per_launch": flop / 1e9,
            "mma_passes": passes, "mma_pipe_tflops": passes * flop / (ms * 1e-3) / 1e12,
            "note": "bf16x3 issues 3 MMAs per algorithmic product (fp32-faithful split); "
                    "`achieved` counts algorithmic FLOPs only"}


def time_gpu_library(dev, steps=10):
    """Secondary yardstick: the ORACLE's modules — stock PyTorch ops, i.e. cuDNN / cuBLAS / ATen
    kernels — running the same iteration on this GPU, fp32 (TF32 off) and with TF32 allowed, launched eagerly and
    as one captured CUDA graph.  Not the target and not the reference arm: it tells how the hand-written path
    compares with the vendor libraries at the same fidelity."""
    import oracle
    from oracle.step import rmsprop_update
    import torch.nn.functional as F
    if CFG["riter"]:
        return None
    out = {}
    B, H, W, code, depth = CFG["B"], CFG["H"], CFG["W"], CFG["code"], CFG["n_lis"]
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        g, d = build_oracle_pair()
        g, d = g.to(dev), d.to(dev)
        if CFG["d_dropout"]:
            d.eval()         # (graph-safe: stock dropout inside a captured graph needs the philox plumbing of torch's RNG)
        gs, ds = {}, {}
        real = torch.rand(B, 3, H, W, device=dev)
        zd, zg = torch.randn(B, code, device=dev), torch.randn(B, code, device=dev)
        ones, zeros = torch.ones(B, 1, device=dev), torch.zeros(B, 1, device=dev)

        def step():
            for p in d.parameters():
                p.requires_grad_(True)
                if p.grad is not None:
                    p.grad.zero_()
            F.binary_cross_entropy(d(real), ones).backward()
            with torch.no_grad():
                fake, _ = g(zd, n_execute_lis_layers=depth)
            F.binary_cross_entropy(d(fake), zeros).backward()
            rmsprop_update(list(d.parameters()), ds, CFG["lr"])
            for p in d.parameters():
                p.requires_grad_(False)
            for p in g.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            fake, lis = g(zg, n_execute_lis_layers=depth)
            total = F.binary_cross_entropy(d(fake), ones)
            for i, u in enumerate(lis):
                total = total + F.mse_loss(u, zg) * (CFG["lambda_r"] ** (i + 1))
            total.backward()
            rmsprop_update(list(g.parameters()), gs, CFG["lr"])

        def timed(fn, n):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        key = "tf32" if tf32 else "fp32"
        try:
            for _ in range(3):
                step()
            out[key + "_eager_ms"] = timed(step, steps)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            graph.replay()
            out[key + "_graph_ms"] = timed(graph.replay, steps)
            del graph
        except Exception as exc:           # a yardstick must never take the bench line down
            out[key + "_error"] = repr(exc)[:200]
        del g, d
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    best = min([v for k, v in out.items() if k.endswith("_ms") and k.startswith("fp32")] or [float("nan")])
    out.update({"images_per_s_fp32": B / (best * 1e-3) if best == best else None, "unit": "ms per step",
                "what": "oracle modules on torch %s + cuDNN %s on this GPU, same iteration, batch %d" % (
                    torch.__version__, torch.backends.cudnn.version(), B)})
    return out


class GlisWorkload(object):
    """G-LIS iteration (configs 1-4, 5b): inputs from the device Philox generator into the graph's static buffers."""

    def __init__(self, dev, world, data_seed, use_graph):
        import common.model as pm
        from glis_b200 import dp, ops
        from glis_b200.trainer import GLISTrainer, GraphedStep
        self.ops, self.seed = ops, data_seed
        gen = pm.GeneratorLearnedInputSpace(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], CFG["code"], "weight",
                                            CFG["n_lis"], CFG["upscaling"]).to(dev)
        dis = pm.build_discriminator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], "weight", CFG["d_dropout"]).to(dev)
        sync = dp.OverlappedGradSync(world) if world > 1 else None
        self.tr = GLISTrainer(gen, dis, lr=CFG["lr"], lambda_r=CFG["lambda_r"], grad_sync=sync)
        B, H, W, code = CFG["B"], CFG["H"], CFG["W"], CFG["code"]
        self.depth = CFG["n_lis"]   # LIS depth forced to "all": fixed work per step
        self.graphed = GraphedStep(self.tr, B, H, W, code, dev) if use_graph else None
        if self.graphed is not None:
            self.real, self.zd, self.zg = self.graphed.real, self.graphed.z_d, self.graphed.z_g
        else:
            self.real = torch.empty(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
            self.zd, self.zg = torch.empty(B, code, device=dev), torch.empty(B, code, device=dev)
        self.counter = 0
        self.h2d = 4 * (self.real.numel() + self.zd.numel() + self.zg.numel())
        self.d2h = 4 * (3 + CFG["n_lis"])

    def fill(self):
        off = self.counter * (1 << 22)
        self.counter += 1
        self.ops.uniform_(self.real, self.seed, off)
        self.ops.randn_(self.zd, self.seed + 7, off)
        self.ops.randn_(self.zg, self.seed + 13, off)

    def device_step(self):
        self.fill()
        if self.graphed is not None:
            return self.graphed.step(None, None, None, self.depth, self.depth)
        return self.tr.step(self.real, self.zd, self.zg, self.depth, self.depth)

    def eager_step(self):
        self.fill()
        return self.tr.step(self.real, self.zd, self.zg, self.depth, self.depth)

    def make_e2e(self):
        """(step(), drain() -> last losses): pinned host batches -> H2D -> iteration -> D2H of the losses, through
        trainer.HostFedStepper (copy of iteration i+1 overlaps iteration i; losses read one iteration late)."""
        B, H, W, code = CFG["B"], CFG["H"], CFG["W"], CFG["code"]
        h_real = torch.rand(B, 3, H, W).pin_memory()
        h_zd, h_zg = torch.randn(B, code).pin_memory(), torch.randn(B, code).pin_memory()
        last = [None]
        if self.graphed is not None:
            from glis_b200.trainer import HostFedStepper
            feeder = HostFedStepper(self.graphed)

            def step():
                last[0] = feeder.submit(h_real, h_zd, h_zg, self.depth, self.depth)

            def drain():
                last[0] = feeder.flush()
                return last[0]
            return step, drain, "glis_b200.trainer.HostFedStepper.submit(pinned real, z_d, z_g) -> losses"
        h_loss = torch.empty(3).pin_memory()

        def step():
            self.real.copy_(h_real, non_blocking=True)
            self.zd.copy_(h_zd, non_blocking=True)
            self.zg.copy_(h_zg, non_blocking=True)
            out = self.tr.step(self.real, self.zd, self.zg, self.depth, self.depth)
            h_loss.copy_(torch.stack([out["d_real"], out["d_fake"], out["g"]]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            last[0] = h_loss.tolist()
        return step, (lambda: last[0]), "glis_b200.trainer.GLISTrainer.step(real, z_d, z_g) from pinned host batches"


class RIterWorkload(object):
    """R-iterative outer iteration (config 5a): 1 + R hops, every hop trained (--always_train_all)."""

    def __init__(self, dev, world, data_seed, use_graph):
        import common.model as pm
        from glis_b200 import dp, ops
        from glis_b200.trainer import GraphedRIter, RIterTrainer
        self.ops, self.seed = ops, data_seed
        gen = pm.build_generator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], CFG["code"], "weight").to(dev)
        rev = pm.build_reverser(CFG["W"], CFG["H"], CFG["nfeature"] // 2, CFG["nlayer"], CFG["code"], "weight", 0).to(dev)
        dis = pm.build_discriminator(CFG["W"], CFG["H"], CFG["nfeature"], CFG["nlayer"], "weight").to(dev)
        sync = dp.OverlappedGradSync(world) if world > 1 else None
        self.tr = RIterTrainer(gen, rev, dis, lr=CFG["lr"], lambda_r=CFG["lambda_r"], r_iterations=CFG["riter"],
                               grad_sync=sync)
        B, H, W, code = CFG["B"], CFG["H"], CFG["W"], CFG["code"]
        self.hops = 1 + CFG["riter"]
        self.flags = [True] * self.hops
        self.graphed = GraphedRIter(self.tr, B, H, W, code, dev) if use_graph else None
        if self.graphed is not None:
            self.first, self.reals = self.graphed.first_code, self.graphed.reals
        else:
            self.first = torch.empty(B, code, device=dev)
            self.reals = [torch.empty(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
                          for _ in range(self.hops)]
        self.counter = 0
        self.h2d = 4 * (self.first.numel() + sum(r.numel() for r in self.reals))
        self.d2h = 4 * (4 * self.hops - 1)

    def fill(self):
        off = self.counter * (1 << 24)
        self.counter += 1
        self.ops.randn_(self.first, self.seed + 7, off)
        for i, r in enumerate(self.reals):
            self.ops.uniform_(r, self.seed, off + (i << 21))

    def device_step(self):
        self.fill()
        if self.graphed is not None:
            return self.graphed.step(None, None, self.flags)
        return self.tr.step(self.first, self.reals, self.flags)

    def eager_step(self):
        self.fill()
        return self.tr.step(self.first, self.reals, self.flags)

    def make_e2e(self):
        B, H, W, code = CFG["B"], CFG["H"], CFG["W"], CFG["code"]
        h_first = torch.randn(B, code).pin_memory()
        h_reals = [torch.rand(B, 3, H, W).pin_memory() for _ in range(self.hops)]
        h_loss = torch.empty(4 * self.hops).pin_memory()
        last = [None]

        def step():
            if self.graphed is not None:
                out = self.graphed.step(h_first, h_reals, self.flags)
            else:
                self.first.copy_(h_first, non_blocking=True)
                for d_, s_ in zip(self.reals, h_reals):
                    d_.copy_(s_, non_blocking=True)
                out = self.tr.step(self.first, self.reals, self.flags)
            vals = [v.reshape(()) for rec in out for v in rec.values()]
            h_loss[:len(vals)].copy_(torch.stack(vals), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            last[0] = h_loss[:len(vals)].tolist()
        return step, (lambda: last[0]), "glis_b200.trainer.GraphedRIter.step(pinned first_code, reals) -> losses"


def run_ours(args, rank, world, local):
    import torch.distributed as dist
    from glis_b200 import _lib, dp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the product path needs a CUDA device (there is no CPU fallback)")
    if args.precision:
        _lib.set_precision(args.precision)
    prec_name = [k for k, v in _lib.PRECISION_NAMES.items() if v == _lib.default_precision][0]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    data_seed = dp.seed_everything(SEED, rank)
    wl = (RIterWorkload if CFG["riter"] else GlisWorkload)(dev, world, data_seed, not args.no_graph)
    B = CFG["B"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.2 s to produce its first sample: start it before the warm-up
    for _ in range(args.warmup):
        wl.device_step()
    ms_total = timed_region(wl.device_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = B * world / (ms_step * 1e-3)

    # ---- kernels per iteration: a replayed graph launches what was captured once, so count one eager iteration
    # through the C ABI (every launch of the library goes through _lib.call) and tally the tensor-core launches by tag
    _lib.launch_tags.clear()
    l0 = _lib.launch_count
    wl.eager_step()
    torch.cuda.synchronize()
    per_step = _lib.launch_count - l0
    tags = dict(_lib.launch_tags)
    launches = args.steps * (per_step + 3)       # + the three input generators
    if args.kernel_table and rank == 0:
        with open(args.kernel_table, "w") as fh:
            json.dump({"launches_per_iteration": per_step, "tagged": tags}, fh, indent=1, sort_keys=True)

    probe = probe_dominant_kernel(dev) if rank == 0 else None

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the losses, through the public host-fed API
    e2e_step, e2e_drain, e2e_api = wl.make_e2e()
    for _ in range(max(3, args.warmup // 2)):
        e2e_step()
    e2e_drain()
    e2e_losses = [None]

    def e2e_region():
        for _ in range(args.steps):
            e2e_step()
        e2e_losses[0] = e2e_drain()   # the last iteration's losses are read inside the timed region too

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    e2e_region()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e_value = B * world / (ms_e2e * 1e-3)

    if rank != 0:
        return
    peaks = measured_peaks()
    roofline = probe
    if roofline is not None:
        def flops(tag):
            dims = dict(kv.split("=") for kv in tag.split() if "=" in kv)
            return 2.0 * int(dims["M"]) * int(dims["N"]) * int(dims["K"])
        tc = {t: c for t, c in tags.items() if t.endswith(" tc") or " tc " in t}
        roofline["peak"] = peaks["tf_burst"]
        roofline["frac"] = roofline["achieved"] / peaks["tf_burst"]
        roofline["frac_mma_pipe"] = roofline["mma_pipe_tflops"] / peaks["tf_burst"]
        roofline["peak_source"] = peaks["source"] + " bf16 dense cuBLAS, burst (kernel timed alone)"
        roofline["tensor_core_launches_per_step"] = sum(tc.values())
        roofline["tensor_core_gflop_per_step"] = sum(flops(t) * c for t, c in tc.items()) / 1e9
    cpu = lib = None
    if world == 1 and not args.no_cpu_baseline:
        sec, cores, b_step = time_cpu_oracle(3, 1)
        cpu = {"value": b_step / sec, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "3 full iterations of the workload (batch %d) after 1 warm-up; oracle port, torch %s CPU, %s; "
                         "reference pin (py2.7/PyTorch@065c5986) not installable offline" % (
                             b_step, torch.__version__, cpu_model())}
        lib = time_gpu_library(dev)
    print(json.dumps({
        "metric": metric_name(), "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (split-bf16 tcgen05, fp32-faithful) + f32", "bf16": "bf16",
                  "tf32": "tf32 gradients + bf16x3 forward"}.get(prec_name, prec_name),
        "data": "synthetic", "config": bench_config(),
        "details": {"global_batch": B * world, "parallelism": "dp%d" % world, "precision": prec_name,
                    "cuda_graph": wl.graphed is not None, "step_tflops": GFLOP_PER_STEP * world / ms_step,
                    "launches_per_iteration": per_step},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h,
                "ms_per_step": ms_e2e, "host_wall_ms_per_step": wall_ms / args.steps, "last_losses": e2e_losses[0],
                "api": e2e_api},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_library_baseline": lib,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS),
                    help="BASELINE.json config: 1, 2 (default, the driver's line), 3, 4, 5a (R-iterative), 5b")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python (no CUDA graph)")
    ap.add_argument("--kernel-table", default=None, help="write the per-iteration launch tally (JSON) here")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "bf16"])
    args = ap.parse_args()
    select_config(args.config)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        from glis_b200 import dp
        dp.init_from_env("nccl")
    try:
        run_ours(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
